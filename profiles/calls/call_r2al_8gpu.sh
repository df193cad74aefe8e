set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8) > gpurun_out/bench_r2al_8gpu.json 2> gpurun_out/bench_r2al_8gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r2al_8gpu.err
