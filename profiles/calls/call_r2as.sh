set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gputest_r2as.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2as.log
tail -3 gpurun_out/gputest_r2as.log
(time timeout 1500 python bench.py) > gpurun_out/bench_r2as_default.json 2> gpurun_out/bench_r2as_default.err; echo "bench rc=$?"
grep real gpurun_out/bench_r2as_default.err
