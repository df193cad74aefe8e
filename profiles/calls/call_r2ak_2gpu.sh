set -x
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2) > gpurun_out/bench_r2ak_2gpu.json 2> gpurun_out/bench_r2ak_2gpu.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/bench_r2ak_2gpu.err
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 10 --warmup 3) > gpurun_out/bench_r2ak_2gpu_ref.json 2> gpurun_out/bench_r2ak_2gpu_ref.err; echo "ref rc=$?"
