#!/bin/bash
# N-GPU validation: every -m gpu test (the 2-rank NCCL test included) and bench.py at N = 2 (and 4, 8 when the box has them)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_pytest.log
tail -6 gpurun_out/m_pytest.log
for N in 2 4 8; do
  if [ "$N" -le "$NG" ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m_bench_${N}gpu.log 2>&1
    python - <<PY
import json
l=[x for x in open('gpurun_out/m_bench_${N}gpu.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print($N, d['value'], d['e2e']['value']); print(json.dumps({k:d.get(k) for k in ('strong_scaling','knn_sharded','kmeans_sharded')}))
else:
    print(open('gpurun_out/m_bench_${N}gpu.log').read()[-2500:])
PY
  fi
done
