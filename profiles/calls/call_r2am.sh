set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gputest_r2am.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2am.log
tail -3 gpurun_out/gputest_r2am.log
(time timeout 1500 python bench.py) > gpurun_out/bench_r2am_default.json 2> gpurun_out/bench_r2am_default.err; echo "bench rc=$?"
grep real gpurun_out/bench_r2am_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2am_reference.json 2> gpurun_out/bench_r2am_reference.err; echo "ref rc=$?"
