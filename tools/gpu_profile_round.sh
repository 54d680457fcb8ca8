TAG=${TAG:-r01i}
# one command, several captures (profiles/${TAG}_*): the launch list and the --set full captures of the kernels
# bench.py reports rooflines for.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 1 --warmup 1 --profile-mode --tracks 64 --skip-knn --skip-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_block -s 5 -c 5 -o gpurun_out/${TAG}_fused $CMD > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 8 -c 8 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 1 -c 1 -o gpurun_out/${TAG}_mel $CMD > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_kernel -s 1 -c 1 -o gpurun_out/${TAG}_stem $CMD > gpurun_out/ncu6.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:select_rerank -c 1 -o gpurun_out/${TAG}_knn_select python - > gpurun_out/ncu5.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from audiomuse_ai_b200 import corpus, voyager_compat as vc
x = corpus.knn_library(100_000, 512, 1234); q = corpus.knn_queries(x, 256, 0, 1)
idx = vc.Index(vc.Space.Cosine, num_dimensions=512); idx.add_items(x)
idx.query(q, 50)
PY
ls -la gpurun_out | tail -8
(timeout 600 python -m pytest tests/test_gpu_kmeans.py -q 2>&1 | tail -4)
