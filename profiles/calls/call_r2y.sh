run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload c4 --skip-cpu --skip-e2e --no-clocks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.3f M grad-evals/s  device %.3f s  failed %s  min_ess/s %s  mean_tree %.1f' % (d['value']/1e6, d.get('job_seconds_device',0), d.get('failed_chains'), d.get('min_bulk_ess_per_sec'), d.get('mean_tree_size',0)))"; }
run B2_PBLOCK_CTAS=2
run B2_PBLOCK_CTAS=1
run B2_PBLOCK_CTAS=2
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "vol or sv or persistent or block or c4 or hostsim or parity" 2>&1 | tail -4
python profiles/sv_ncu_target.py > gpurun_out/sv_plain.log 2>&1 && tail -1 gpurun_out/sv_plain.log && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_persistent_block -s 1 -c 1 -o gpurun_out/prof_sv_r2 -f python profiles/sv_ncu_target.py > gpurun_out/ncu_sv_r2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_sv_r2.log
