mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 600 python -m pytest tests/test_gpu_mel.py tests/test_gpu_kmeans.py -q 2>&1 | tail -15) > gpurun_out/t_mel_km.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/smoke.log
(timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5) > gpurun_out/bench1.log
# launch list (cold-cache, serialised: shares only) and full captures of the two roofline kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --profile-mode --tracks 64 --skip-knn --skip-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 20 -c 18 -o gpurun_out/prof_gemm \
  python bench.py --steps 1 --warmup 1 --profile-mode --tracks 64 --skip-knn --skip-e2e --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 1 -c 1 -o gpurun_out/prof_mel \
  python bench.py --steps 1 --warmup 1 --profile-mode --tracks 64 --skip-knn --skip-e2e --no-cpu-baseline > gpurun_out/ncu_mel.log 2>&1
for f in t_mel_km smoke bench1; do echo "=== $f"; cat gpurun_out/$f.log; done
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_gemm.log gpurun_out/ncu_mel.log
ls -la gpurun_out
