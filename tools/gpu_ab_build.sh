#!/bin/bash
# A/B of compile-time switches of fused_block_t.cu: each line of AB_FLAGS = extra nvcc flags (or "none")
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
while IFS= read -r cfg; do
  [ -z "$cfg" ] && continue
  touch audiomuse-ai_b200/csrc/fused_block_t.cu audiomuse-ai_b200/csrc/fused_block.cu
  if [ "$cfg" = none ]; then flags=""; else flags="$cfg"; fi
  AM_EXTRA_NVCC_FLAGS="$flags" python audiomuse-ai_b200/build_native.py > /dev/null 2> gpurun_out/ab_build.log || { echo "$cfg BUILD FAILED"; tail -5 gpurun_out/ab_build.log; continue; }
  timeout 600 python bench.py --skip-knn --skip-scale --skip-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab.log 2>&1
  python - "$cfg" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/ab.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); k=d['kernel_ms_per_step']
    print(sys.argv[1], '->', round(d['ms_per_step'],3), {a: k[a] for a in k if 'fused' in a})
else:
    print(sys.argv[1], 'FAILED', open('gpurun_out/ab.log').read()[-800:])
PY
done <<< "$AB_FLAGS"
