set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/gputest_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2c.log
tail -12 gpurun_out/gputest_r2c.log
(time timeout 900 python bench.py --steps 6 --warmup 3 --cpu-budget 5) > gpurun_out/bench_r2c_all.json 2> gpurun_out/bench_r2c_all.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r2c_all.err
