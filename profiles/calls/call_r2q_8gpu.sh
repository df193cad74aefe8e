mkdir -p gpurun_out
nvidia-smi -L | wc -l
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 6 --warmup 3 --configs c3,c4,c5) > gpurun_out/bench_r2q_8gpu.json 2> gpurun_out/bench_r2q_8gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r2q_8gpu.err
