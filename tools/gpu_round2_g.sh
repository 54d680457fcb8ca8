#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py -x -q -k "not config2_edge" > gpurun_out/g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-scale --no-cpu-baseline --skip-knn > gpurun_out/g_bench.log 2>&1
bash tools/gpu_trace.sh > gpurun_out/g_trace.log 2>&1
tail -3 gpurun_out/g_pytest.log; python - <<'PY'
import json
l=[x for x in open('gpurun_out/g_bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['kernel_ms_per_step'])
PY
grep -A2 "fused trace" gpurun_out/g_trace.log | cut -c1-150 | tail -24
