"""Compare the fused-block encoder path against the unfused one, block by block (GPU debug aid)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, numpy as np
sys.path.insert(0, %r)
from audiomuse_ai_b200 import clap_analyzer as ca, weights
sd = weights.random_state_dict(0)
sess = ca.B200Session.from_state_dict(sd)
rng = np.random.default_rng(0)
mel = (rng.standard_normal((3, 1, 128, %d)) * 12 - 30).astype(np.float32)
out = sess.run(None, {"mel_spectrogram": mel})[0]
np.save(sys.argv[1], out)
"""


def run(mask, T, path):
    env = dict(os.environ, AM_FUSED_BLOCKS=str(mask))
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, T), path], env=env, capture_output=True, text=True,
                       timeout=180)
    return r.returncode, r.stderr[-600:]


def main():
    import numpy as np
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
    rc, err = run(0, T, "/tmp/fd_ref.npy")
    assert rc == 0, err
    ref = np.load("/tmp/fd_ref.npy")
    for mask in [1, 2, 4, 8, 16, 32, 64, 128, 256, 0x1f, 0xffffffff]:
        try:
            rc, err = run(mask, T, "/tmp/fd_out.npy")
        except subprocess.TimeoutExpired:
            print(f"mask {mask:#x}: TIMEOUT (hang)")
            continue
        if rc != 0:
            print(f"mask {mask:#x}: rc={rc} {err}")
            continue
        out = np.load("/tmp/fd_out.npy")
        cos = [float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))) for a, b in zip(out, ref)]
        print(f"mask {mask:#x}: min cosine vs unfused = {min(cos):.6f} max|diff| = {np.abs(out - ref).max():.2e}")


if __name__ == "__main__":
    main()
