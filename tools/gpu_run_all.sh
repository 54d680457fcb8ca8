mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25) > gpurun_out/t_all.log
(timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3) > gpurun_out/bench2.log
echo "=== tests"; cat gpurun_out/t_all.log
echo "=== bench"; cat gpurun_out/bench2.log
