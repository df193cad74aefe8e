run() { echo "== $*"; env "$@" timeout 300 python bench.py --workload c4 --skip-cpu --skip-e2e --no-clocks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.3f M grad-evals/s  device %.3f s  failed %s  min_ess %s' % (d['value']/1e6, d.get('job_seconds_device',0), d.get('failed_chains'), (d.get('ess') or {}).get('min_bulk_ess')))"; }
run B2_PBLOCK_NT=384
run B2_PBLOCK_NT=256
run B2_PBLOCK_NT=384
