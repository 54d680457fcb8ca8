mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 600 python -m pytest tests/test_gpu_encoder.py -q 2>&1 | tail -4) > gpurun_out/t_enc_fused.log; cat gpurun_out/t_enc_fused.log
(timeout 600 python bench.py --steps 5 --warmup 3 --skip-knn --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/bench_fused.log
python3 -c "
import json
l=[x for x in open('gpurun_out/bench_fused.log') if x.startswith('{')][-1]; d=json.loads(l)
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['kernel_ms_per_step'])"
