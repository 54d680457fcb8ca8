"""tcgen05.mma cost probe (am_probe_mma).  GPU box only."""
import ctypes as C
import sys

sys.path.insert(0, ".")
import audiomuse_ai_b200 as am  # noqa: E402,F401
from audiomuse_ai_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
for traffic in (0, 1, 2):
    for N in (64, 128, 256):
        for tiles in (1, 2, 3):
            if tiles * N > 512:
                continue
            a, b = C.c_double(0), C.c_double(0)
            _lib.check_debug(lib.am_probe_mma(N, 960, tiles, traffic, C.byref(a), C.byref(b)))
            print(f"traffic={traffic} N={N} d_tiles={tiles}: issue {a.value:.1f} cyc/MMA, to-commit {b.value:.1f} cyc/MMA", flush=True)
