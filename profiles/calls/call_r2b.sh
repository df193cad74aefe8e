set -x
mkdir -p gpurun_out
timeout 600 python profiles/tc_fused_probe.py > gpurun_out/tc_fused_probe_r2b.txt 2>&1; echo "probe rc=$?"
tail -20 gpurun_out/tc_fused_probe_r2b.txt
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/gputest_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2b.log
tail -15 gpurun_out/gputest_r2b.log
