mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 1500 python tools/fused_debug.py 1001 2>&1 | tail -20) > gpurun_out/fused_debug.log
cat gpurun_out/fused_debug.log
(timeout 600 python -m pytest tests/test_gpu_encoder.py -q 2>&1 | tail -8) > gpurun_out/t_enc_fused.log
cat gpurun_out/t_enc_fused.log
(timeout 600 python bench.py --steps 5 --warmup 3 --skip-knn --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/bench_fused.log
cat gpurun_out/bench_fused.log
