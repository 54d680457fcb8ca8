mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_encoder.py tests/test_gpu_knn.py -q -x 2>&1 | tail -6) > gpurun_out/t_gemm.log; cat gpurun_out/t_gemm.log
(timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/bench_fused.log
python3 -c "
import json
l=[x for x in open('gpurun_out/bench_fused.log') if x.startswith('{')][-1]; d=json.loads(l)
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['kernel_ms_per_step']); print(d['knn'])"
python tools/gemm_bench.py 2>&1 | tail -9
