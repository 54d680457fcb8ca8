"""TMEM read bandwidth probe (am_probe_tmem_ld), alone and beside an MMA stream.  GPU box only."""
import ctypes as C
import sys

sys.path.insert(0, ".")
import audiomuse_ai_b200 as am  # noqa: E402,F401
from audiomuse_ai_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
for warps, n_mma in ((4, 0), (8, 0), (16, 0), (16, 2000), (16, 8000)):
    for cols, depth in ((16, 1), (16, 2), (32, 1)):
        bpc, cpm = C.c_double(0), C.c_double(0)
        _lib.check_debug(lib.am_probe_tmem_ld(warps, cols, depth, 960, n_mma, C.byref(bpc), C.byref(cpm)))
        print(f"warps={warps} x{cols} depth={depth} n_mma={n_mma}: {bpc.value:.1f} B/cycle/SM, {cpm.value:.1f} cyc/MMA", flush=True)
