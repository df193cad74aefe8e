set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench profiles/ubench_pipes.cu && timeout 120 /tmp/ubench > gpurun_out/ubench_pipes_r2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 600 > gpurun_out/gputest_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r2a.log
tail -5 gpurun_out/gputest_r2a.log
for cfg in "0 0" "0 1" "1 0" "1 1"; do
  set -- $cfg
  B2_TC_FUSED=$1 B2_TC_EPI=$2 timeout 300 python bench.py --steps 6 --warmup 3 --skip-cpu --skip-ess --no-profile > gpurun_out/bench_c2_r2a_f$1_e$2.json 2> gpurun_out/bench_c2_r2a_f$1_e$2.err
  echo "fused=$1 epi=$2 rc=$?"; head -c 400 gpurun_out/bench_c2_r2a_f$1_e$2.json; echo
done
