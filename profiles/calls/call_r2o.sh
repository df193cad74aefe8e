set -x
mkdir -p gpurun_out
echo "== 4 stages"; B2_TC_STAGES_UNFUSED=4 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -1
echo "== 5 stages (default)"; timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/gputest_r2o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2o.log
tail -4 gpurun_out/gputest_r2o.log
python profiles/hier_ncu_target.py > gpurun_out/plain_hier_target.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_hier_slab -s 2 -c 2 -o gpurun_out/prof_hier_r2 python profiles/hier_ncu_target.py > gpurun_out/ncu_hier_r2.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/plain_hier_target.log | tail -1
