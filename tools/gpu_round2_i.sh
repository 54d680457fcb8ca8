#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_frontend.py tests/test_gpu_ref_trace.py -x -q > gpurun_out/i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/i_pytest.log
timeout 600 python tools/knn_bench.py > gpurun_out/i_knn.log 2>&1
timeout 900 python tools/e2e_from_wav.py > gpurun_out/i_e2e_wav.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/i_bench.log 2>&1
tail -5 gpurun_out/i_pytest.log; tail -1 gpurun_out/i_knn.log; tail -1 gpurun_out/i_e2e_wav.log; python - <<'PY'
import json
l=[x for x in open('gpurun_out/i_bench.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['knn'])
else:
    print(open('gpurun_out/i_bench.log').read()[-3000:])
PY
