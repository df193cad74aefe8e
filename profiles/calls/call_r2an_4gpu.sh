set -x
mkdir -p gpurun_out
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 4) > gpurun_out/bench_r2an_4gpu.json 2> gpurun_out/bench_r2an_4gpu.err; echo "bench rc=$?"
grep real gpurun_out/bench_r2an_4gpu.err
