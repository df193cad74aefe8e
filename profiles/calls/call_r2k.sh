mkdir -p gpurun_out
for c in f7b4c0b 35135a3 96352ee ee52c1d; do
  (cd _snap/$c && python -m pymc3_b200.build > /dev/null 2>&1
   B2_TC_FUSED=0 B2_TC_EPI=0 B2_TC_FLUSH=100000 timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --skip-cpu --skip-ess --no-profile --no-clocks > ../../gpurun_out/s_$c.json 2> ../../gpurun_out/s_$c.err
   python - <<PY
import json
try:
    d=json.loads([l for l in open('../../gpurun_out/s_$c.json') if l.startswith('{')][-1]); print('$c', 'value %.3fM'%(d['value']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'failed', d['failed_chains'])
except Exception as e: print('$c', 'ERR', e)
PY
  )
done
