mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --skip-cpu --skip-ess --skip-e2e --no-profile --no-clocks > gpurun_out/v_$label.json 2> gpurun_out/v_$label.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/v_$label.json') if l.startswith('{')][-1]); print('$label', 'value %.3fM'%(d['value']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'failed', d['failed_chains'])
except Exception as e: print('$label', 'ERR', e)
PY
}
run base_epi0_noref B2_TC_EPI=0 B2_TC_NOREF=1
B2_NVCC_EXTRA="-DTC_V_NOETA" python -m pymc3_b200.build --force > /dev/null 2>&1
run noeta_epi0 B2_TC_EPI=0 B2_TC_NOREF=1
run noeta_epi1 B2_TC_EPI=1 B2_TC_NOREF=1
B2_NVCC_EXTRA="-DTC_V_NOETA -DTC_V_SYNCTHREADS" python -m pymc3_b200.build --force > /dev/null 2>&1
run noeta_sync_epi0 B2_TC_EPI=0 B2_TC_NOREF=1
