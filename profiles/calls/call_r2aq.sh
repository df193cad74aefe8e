timeout 600 python profiles/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_smoke.py > gpurun_out/sanitize_r2.log 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/sanitize_r2.log
