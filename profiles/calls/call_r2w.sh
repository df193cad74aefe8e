timeout 600 python bench.py --workload c2 --skip-cpu --skip-ess --no-profile --configs "" > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err; echo rc=$?
tail -c 300 gpurun_out/e2e_probe.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/e2e_probe.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
print(json.dumps(d['e2e']))
PY
