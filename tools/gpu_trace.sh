export PYTHONUNBUFFERED=1
# the phase trace is compiled out of the shipped library: rebuild the fused kernel with it (box-local copy)
touch audiomuse-ai_b200/csrc/fused_block.cu audiomuse-ai_b200/csrc/fused_block_t.cu
AM_EXTRA_NVCC_FLAGS=-DAM_FUSED_TRACE_BUILD python audiomuse-ai_b200/build_native.py > /dev/null
AM_FUSED_TRACE=${TRACE_MODE:-1} AM_CLAP_SUB_BATCH=64 timeout 300 python - <<'PY' 2>&1 | grep -A 16 "fused.* trace" | tail -${TRACE_LINES:-120}
import numpy as np, sys
sys.path.insert(0, ".")
from audiomuse_ai_b200 import clap_analyzer as ca, weights
sess = ca.B200Session.from_state_dict(weights.random_state_dict(0))
mel = (np.random.default_rng(0).standard_normal((64, 1, 128, 1001)) * 12 - 30).astype(np.float32)
sess.run(None, {"mel_spectrogram": mel})
sess.run(None, {"mel_spectrogram": mel})
PY
