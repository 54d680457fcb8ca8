mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/t_all.log
cat gpurun_out/t_all.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused.csv \
  python bench.py --steps 1 --warmup 1 --profile-mode --tracks 64 --skip-knn --skip-e2e --no-cpu-baseline > gpurun_out/ncu_list_f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_block -s 5 -c 5 -o gpurun_out/prof_fused \
  python bench.py --steps 1 --warmup 1 --profile-mode --tracks 64 --skip-knn --skip-e2e --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
tail -n 2 gpurun_out/ncu_fused.log
