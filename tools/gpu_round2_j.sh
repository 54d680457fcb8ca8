#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_gemm.py tests/test_gpu_ref_trace.py -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/j_pytest.log
timeout 600 python tools/knn_bench.py > gpurun_out/j_knn.log 2>&1
timeout 900 python tools/e2e_from_wav.py > gpurun_out/j_e2e_wav.log 2>&1
tail -4 gpurun_out/j_pytest.log; tail -1 gpurun_out/j_knn.log; tail -1 gpurun_out/j_e2e_wav.log
