timeout 600 python bench.py --workload c2 --skip-cpu --skip-ess --no-profile --configs "" > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/e2e_probe.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['whole_call_value'])
print(json.dumps(d['e2e']['host_phases_s']))
print(d['e2e']['chunk_log'][:3])
PY
