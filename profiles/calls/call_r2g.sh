set -x
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --configs c3,c4,c5) > gpurun_out/bench_r2g_2gpu.json 2> gpurun_out/bench_r2g_2gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2g_2gpu.err
