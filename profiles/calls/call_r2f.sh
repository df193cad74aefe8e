set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gputest_r2f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2f.log
tail -5 gpurun_out/gputest_r2f.log
(time timeout 900 python bench.py --steps 17 --warmup 3 --cpu-budget 5 --configs c1,c4) > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2f.err
B2_GRAPH=0 timeout 300 python bench.py --steps 17 --warmup 3 --skip-cpu --skip-e2e --skip-ess --workload c2 > gpurun_out/bench_r2f_nograph.json 2> gpurun_out/bench_r2f_nograph.err; echo "nograph rc=$?"
