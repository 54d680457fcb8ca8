#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_onnx.py -x -q -s > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.log 2>&1
bash tools/gpu_trace.sh > gpurun_out/f_trace.log 2>&1
grep -E "parity|passed|failed|exit" gpurun_out/f_pytest.log | tail -30; python - <<'PY'
import json
l=[x for x in open('gpurun_out/f_bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['kernel_ms_per_step'])
PY
grep -A3 "fused trace" gpurun_out/f_trace.log | cut -c1-230 | head -40
