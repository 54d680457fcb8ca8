#!/bin/bash
# first GPU call of round 2: whole GPU suite, mel error report, short bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
timeout 300 python tools/mel_error_report.py > gpurun_out/a_mel_err.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.log 2>&1
tail -5 gpurun_out/a_pytest.log; cat gpurun_out/a_mel_err.log; tail -2 gpurun_out/a_bench.log
