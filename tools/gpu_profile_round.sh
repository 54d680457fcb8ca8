TAG=${TAG:-r02}
# one command, several captures (profiles/${TAG}_*): the launch list and the --set full captures of the kernels
# bench.py reports rooflines for.  Numbers printed by runs under ncu are never bench values.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 1 --warmup 1 --profile-mode --tracks 64 --skip-knn --skip-e2e --skip-scale --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_block_t -s 4 -c 4 -o gpurun_out/${TAG}_fusedt $CMD > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_block_kernel -s 1 -c 1 -o gpurun_out/${TAG}_fused0 $CMD > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 8 -c 8 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:mel_kernel -s 1 -c 1 -o gpurun_out/${TAG}_mel $CMD > gpurun_out/ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:select_cm -c 2 -o gpurun_out/${TAG}_knn_select python - > gpurun_out/ncu6.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from audiomuse_ai_b200 import corpus, voyager_compat as vc
x = corpus.knn_library(100_000, 512, 1234); q = corpus.knn_queries(x, 4096, 0, 1)
idx = vc.Index(vc.Space.Cosine, num_dimensions=512); idx.add_items(x)
idx.query(q, 50)
idx.query(q, 50)
PY
ls -la gpurun_out | tail -10
