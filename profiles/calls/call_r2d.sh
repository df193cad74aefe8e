set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/gputest_r2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2d.log
tail -8 gpurun_out/gputest_r2d.log
(time timeout 900 python bench.py --steps 6 --warmup 3 --cpu-budget 5 --configs c1,c3,c4,c5) > gpurun_out/bench_r2d_all.json 2> gpurun_out/bench_r2d_all.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2d_all.err
python profiles/tc_ncu_target.py > gpurun_out/plain_ncu_target.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_glm_tc_main -s 3 -c 2 -o gpurun_out/prof_main_r2 python profiles/tc_ncu_target.py > gpurun_out/ncu_main_r2.log 2>&1; echo "ncu rc=$?"
