#!/bin/bash
# full single-GPU validation: every -m gpu test, smoke(), the default bench line and the reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/k_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/k_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/k_bench_ref.log 2>&1
timeout 900 python tools/e2e_from_wav.py > gpurun_out/k_e2e_wav.log 2>&1
tail -5 gpurun_out/k_pytest.log; tail -2 gpurun_out/k_smoke.log; tail -1 gpurun_out/k_bench.log; tail -1 gpurun_out/k_bench_ref.log; tail -1 gpurun_out/k_e2e_wav.log
