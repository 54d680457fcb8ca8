#!/bin/bash
# k-means tensor-core path: parity tests, config-4 timing, ncu capture of the assignment kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_mel.py -x -q -s > gpurun_out/b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b_pytest.log
timeout 600 python tools/kmeans_bench.py > gpurun_out/b_kmeans.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assign_tc_kernel -c 2 -o gpurun_out/b_kmeans_assign python tools/kmeans_bench.py --iters 1 --check 0 > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:accumulate_sorted -c 2 -o gpurun_out/b_kmeans_accum python tools/kmeans_bench.py --iters 1 --check 0 >> gpurun_out/b_ncu.log 2>&1
tail -15 gpurun_out/b_pytest.log; cat gpurun_out/b_kmeans.log | tail -3
