mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --skip-knn 2>&1 | tail -4) > gpurun_out/bench_2gpu.log
cat gpurun_out/bench_2gpu.log | cut -c1-1500
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -2) > gpurun_out/bench_2gpu_ref.log
cat gpurun_out/bench_2gpu_ref.log | cut -c1-600
