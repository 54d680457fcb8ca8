mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 900 python -m pytest tests/test_gpu_knn.py -q -x 2>&1 | tail -6) > gpurun_out/t_knn.log; cat gpurun_out/t_knn.log
(timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e 2>&1 | tail -2) > gpurun_out/bench_knn.log
python3 -c "
import json
l=[x for x in open('gpurun_out/bench_knn.log') if x.startswith('{')][-1]; d=json.loads(l)
print(d['value'], d['ms_per_step']); print(d['knn'])"
