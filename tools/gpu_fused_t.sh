#!/bin/bash
# channel-per-lane fused block: parity against the unfused path block by block (hang-safe), then timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python tools/fused_debug.py 1001 > gpurun_out/ft_debug.log 2>&1
cat gpurun_out/ft_debug.log | tail -14
timeout 600 python bench.py --skip-knn --skip-scale --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ft_bench.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/ft_bench.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
else:
    print(open('gpurun_out/ft_bench.log').read()[-2000:])
PY
