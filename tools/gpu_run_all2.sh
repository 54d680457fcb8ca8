mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/t_all.log; cat gpurun_out/t_all.log
(timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2) > gpurun_out/bench3.log
python3 -c "
import json
l=[x for x in open('gpurun_out/bench3.log') if x.startswith('{')][-1]; d=json.loads(l)
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['kernel_ms_per_step']); print(d['knn']); print(d['roofline']['frac'], d['roofline_mel']['frac'], d['cpu_baseline'])"
