#!/bin/bash
# channel-per-lane fused block: parity against the unfused path block by block (hang-safe), timing, phase trace
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python tools/fused_debug.py 1001 > gpurun_out/ft_debug.log 2>&1
cat gpurun_out/ft_debug.log | tail -14
for mode in default small; do
  if [ $mode = small ]; then export AM_FUSEDT_SMALL_TILES=1; fi
  timeout 600 python bench.py --skip-knn --skip-scale --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ft_bench_$mode.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/ft_bench_$mode.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$mode', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])
else:
    print(open('gpurun_out/ft_bench_$mode.log').read()[-2000:])
PY
done
unset AM_FUSEDT_SMALL_TILES
TRACE_LINES=70 bash tools/gpu_trace.sh > gpurun_out/ft_trace.log 2>&1
grep -A12 "fused-t" gpurun_out/ft_trace.log | tail -60
