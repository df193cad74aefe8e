set -x
mkdir -p gpurun_out
cd _r1_snapshot && python -m pymc3_b200.build > /dev/null 2>&1; ls -la pymc3_b200/libb200nuts.so
for i in 1 2; do
timeout 400 python bench.py --steps 17 --warmup 3 --skip-cpu --skip-ess --no-profile > ../gpurun_out/bench_r2h_r1code_$i.json 2> ../gpurun_out/bench_r2h_r1code_$i.err; echo "r1 code rc=$?"
cd ..
timeout 400 python bench.py --steps 17 --warmup 3 --skip-cpu --skip-ess --skip-e2e --no-profile --workload c2 > gpurun_out/bench_r2h_r2code_$i.json 2> gpurun_out/bench_r2h_r2code_$i.err; echo "r2 code rc=$?"
cd _r1_snapshot
done
