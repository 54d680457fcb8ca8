mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
export PYTHONUNBUFFERED=1
(timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -25) > gpurun_out/t_gemm.log
(timeout 600 python -m pytest tests/test_gpu_mel.py -q 2>&1 | tail -40) > gpurun_out/t_mel.log
(timeout 900 python -m pytest tests/test_gpu_knn.py -q 2>&1 | tail -60) > gpurun_out/t_knn.log
(timeout 600 python -m pytest tests/test_gpu_kmeans.py -q 2>&1 | tail -40) > gpurun_out/t_kmeans.log
(timeout 900 python -m pytest tests/test_gpu_encoder.py -q 2>&1 | tail -60) > gpurun_out/t_enc.log
(AM_GEMM_IMPL=simt timeout 900 python -m pytest tests/test_gpu_encoder.py -q 2>&1 | tail -60) > gpurun_out/t_enc_simt.log
for f in gemm mel knn kmeans enc enc_simt; do echo "=== $f"; tail -12 gpurun_out/t_$f.log; done
