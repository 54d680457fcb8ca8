"""Issue-rate probe of the fp16x2 instruction kinds (am_probe_pipe).  GPU box only."""
import ctypes as C
import sys

sys.path.insert(0, ".")
import audiomuse_ai_b200 as am  # noqa: E402,F401
from audiomuse_ai_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
names = ["HFMA2 reg", "HFMA2 imm", "HFMA2.SAT", "HMNMX2 x2", "PRMT", "F2FP+add", "HFMA2+PRMT"]
for warps in (4, 8, 16, 18):
    for op, n in enumerate(names):
        v = C.c_double(0)
        _lib.check_debug(lib.am_probe_pipe(op, warps, 2000, C.byref(v)))
        print(f"warps={warps:2d} {n:11s}: {v.value:.2f} cycles / warp-instruction / SMSP", flush=True)
