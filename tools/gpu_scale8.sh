#!/bin/bash
# 8-GPU check of the driver's scaling command (one box, NCCL over NVLink): bench.py at N = 8 and N = 4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/s_bench_${N}gpu.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/s_bench_${N}gpu.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print($N, round(d['value']), round(d['e2e']['value'])); print(json.dumps({k:{a:b for a,b in (d.get(k) or {}).items() if a in ('tracks_per_s_whole_job','queries_per_s','ids_equal_single_rank_answer','ms_per_iteration_device','allreduce_share')} for k in ('strong_scaling','knn_sharded','kmeans_sharded')}))
else:
    print(open('gpurun_out/s_bench_${N}gpu.log').read()[-2500:])
PY
done
