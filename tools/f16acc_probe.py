"""How does TMEM hold fp16 accumulators of tcgen05.mma kind::f16 (am_probe_mn_major, swap bit 1)?  Prints the raw 32-bit
words of a few accumulator cells next to the expected sums.  GPU box only."""
import sys

import numpy as np

sys.path.insert(0, ".")
import audiomuse_ai_b200 as am  # noqa: E402,F401
from audiomuse_ai_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
rng = np.random.default_rng(0)
N, K = 64, 64
a = (rng.integers(-3, 4, (128, K))).astype(np.float16)
b = (rng.integers(-3, 4, (N, K))).astype(np.float16)
want = a.astype(np.float32) @ b.astype(np.float32).T
for flag, name in ((0, "f32 accumulators"), (2, "f16 accumulators")):
    d = np.zeros((128, N), dtype=np.float32)
    try:
        _lib.check_debug(lib.am_probe_mn_major(a.ctypes.data, b.ctypes.data, N, K, 16384, 1024, 128, flag, d.ctypes.data))
    except Exception as e:  # noqa: BLE001
        print(name, "ERROR", e)
        continue
    raw = d.view(np.uint32)
    print(name)
    for r in (0, 1, 37):
        words = [f"{w:08x}" for w in raw[r, :8]]
        halves = [tuple(np.array([w & 0xffff, w >> 16], dtype=np.uint16).view(np.float16).tolist()) for w in raw[r, :8]]
        print(f"  row {r}: words {words}\n          as f16 pairs {halves}\n          as f32 {d[r, :8].tolist()}\n          want columns 0..15 {want[r, :16].tolist()}")
