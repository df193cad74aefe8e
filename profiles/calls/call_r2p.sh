set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gputest_r2p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2p.log
tail -4 gpurun_out/gputest_r2p.log
(time timeout 1200 python bench.py) > gpurun_out/bench_r2p_default.json 2> gpurun_out/bench_r2p_default.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r2p_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2p_reference.json 2> gpurun_out/bench_r2p_reference.err; echo "ref rc=$?"
python bench.py --workload c2 --steps 3 --warmup 3 --iters-per-step 25 --skip-cpu --skip-ess --skip-e2e --no-profile > gpurun_out/plain_launches_r2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/launches_r2.csv python bench.py --workload c2 --steps 3 --warmup 3 --iters-per-step 25 --skip-cpu --skip-ess --skip-e2e --no-profile > gpurun_out/ncu_launches_r2.log 2>&1; echo "ncu launches rc=$?"
python profiles/tc_ncu_target.py > gpurun_out/plain_ncu_target.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_glm_tc_main -s 3 -c 2 -o gpurun_out/prof_main_r2_final python profiles/tc_ncu_target.py > gpurun_out/ncu_main_r2_final.log 2>&1; echo "ncu rc=$?"
