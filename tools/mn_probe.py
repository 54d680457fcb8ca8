"""MN-major A operand probe (am_probe_mn_major): which shared-memory layout / descriptor fields does tcgen05.mma take.  GPU box only."""
import sys

import numpy as np

sys.path.insert(0, ".")
import audiomuse_ai_b200 as am  # noqa: E402,F401
from audiomuse_ai_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
rng = np.random.default_rng(0)
for N, K in ((80, 128), (144, 64)):
    for name, lbo, sbo, rows, swap in (("interleaved lbo=1024 sbo=2048", 1024, 2048, 128, 0),
                                       ("interleaved, fields swapped", 1024, 2048, 128, 1),
                                       ("separate lbo=16384 sbo=1024", 16384, 1024, 128, 0),
                                       ("alias lbo=0 sbo=1024 (64 rows)", 0, 1024, 64, 0)):
        a = rng.standard_normal((rows, K)).astype(np.float16)
        b = rng.standard_normal((N, K)).astype(np.float16)
        d = np.zeros((128, N), dtype=np.float32)
        try:
            _lib.check_debug(lib.am_probe_mn_major(a.ctypes.data, b.ctypes.data, N, K, lbo, sbo, rows, swap, d.ctypes.data))
        except Exception as e:  # noqa: BLE001
            print(f"N={N} K={K} {name}: ERROR {e}", flush=True)
            continue
        want = a.astype(np.float32) @ b.astype(np.float32).T
        if rows == 64:
            want = np.concatenate([want, want], 0)
        err = np.abs(d - want).max()
        print(f"N={N} K={K} {name}: max |diff| {err:.3e} (|want| max {np.abs(want).max():.2f})", flush=True)
