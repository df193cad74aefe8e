set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gputest_r2af.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2af.log
tail -4 gpurun_out/gputest_r2af.log
(time timeout 1500 python bench.py) > gpurun_out/bench_r2af_default.json 2> gpurun_out/bench_r2af_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r2af_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2af_reference.json 2> gpurun_out/bench_r2af_reference.err; echo "ref rc=$?"
