#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_knn.py -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/d_pytest.log
timeout 600 python tools/kmeans_bench.py > gpurun_out/d_kmeans.log 2>&1
timeout 600 python tools/knn_bench.py > gpurun_out/d_knn.log 2>&1
tail -4 gpurun_out/d_pytest.log; tail -2 gpurun_out/d_kmeans.log; tail -2 gpurun_out/d_knn.log
