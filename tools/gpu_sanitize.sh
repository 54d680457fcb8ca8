#!/bin/bash
# compute-sanitizer over smoke() (small student, fused kernels) and one full-size window: memcheck + synccheck
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_memcheck_smoke.txt 2>&1
tail -4 gpurun_out/r02_memcheck_smoke.txt
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_synccheck_smoke.txt 2>&1
tail -3 gpurun_out/r02_synccheck_smoke.txt
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python - > gpurun_out/r02_memcheck_full_window.txt 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from audiomuse_ai_b200 import clap_analyzer as ca, weights
sess = ca.B200Session.from_state_dict(weights.random_state_dict(0))
mel = (np.random.default_rng(0).standard_normal((2, 1, 128, 1001)) * 12 - 30).astype(np.float32)
out = sess.run(None, {"mel_spectrogram": mel})[0]
print("embedding norms", np.linalg.norm(out, axis=1))
PY
tail -4 gpurun_out/r02_memcheck_full_window.txt
