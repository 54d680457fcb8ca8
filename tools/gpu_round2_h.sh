#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_ref_trace.py -x -q -s > gpurun_out/h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/h_pytest.log
timeout 900 python tools/e2e_from_wav.py > gpurun_out/h_e2e_wav.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/h_bench.log 2>&1
grep -E "resample|ref trace|passed|failed|Error|exit" gpurun_out/h_pytest.log | tail -20; tail -2 gpurun_out/h_e2e_wav.log; python - <<'PY'
import json
l=[x for x in open('gpurun_out/h_bench.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['e2e']['value']); print(json.dumps({k:d.get(k) for k in ('strong_scaling','knn_sharded','kmeans_sharded')}, indent=1))
else:
    print(open('gpurun_out/h_bench.log').read()[-3000:])
PY
