#!/bin/bash
# k-means (balanced accumulate), fused k-NN: tests + timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_knn.py tests/test_gpu_gemm.py -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c_pytest.log
timeout 600 python tools/kmeans_bench.py > gpurun_out/c_kmeans.log 2>&1
timeout 600 python tools/bench_scale.py > gpurun_out/c_scale.log 2>&1
tail -15 gpurun_out/c_pytest.log; tail -2 gpurun_out/c_kmeans.log; tail -30 gpurun_out/c_scale.log
