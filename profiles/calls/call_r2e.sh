set -x
mkdir -p gpurun_out
timeout 600 python profiles/tc_ab.py > gpurun_out/tc_ab_r2e.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/tc_ab_r2e.txt
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/gputest_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2e.log
tail -8 gpurun_out/gputest_r2e.log
(time timeout 900 python bench.py --steps 6 --warmup 3 --skip-cpu --configs c5) > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2e.err
