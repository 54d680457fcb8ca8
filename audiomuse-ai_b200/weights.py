"""Student CLAP audio-encoder weights: architecture plan, BatchNorm folding, "AMW1" blob.

The reference ships its encoder as an ONNX file that is not available offline
(Dockerfile:254-282); the in-repo definition is ``StudentCLAPAudio``
(student_clap/models/student_onnx_model.py:166-287) built on micromind's PhiNet with
student_clap/config.yaml:15-24.  This module turns a ``StudentCLAPAudio.state_dict()``
(PyTorch checkpoint naming) into the flat blob ``am_clap_load`` reads:

    "AMW1" u32 version u32 n_mels u32 emb_dim u32 n_records
    record := u32 type, i32 params[8], arrays...      array := u64 count, f32 data[count]
      type 0 stem       params (cout, pad_t, pad_b, pad_l, pad_r)
                        arrays bn0_scale[n_mels] bn0_shift[n_mels] dw[9] pw_scale[cout] pw_shift[cout]
      type 1 pointwise  params (cin, cout, relu6, residual, block_start)  arrays W[cout,cin] bias[cout]
      type 2 depthwise  params (c, stride, -, -, block_start)             arrays W[c,9] bias[c]
      type 3 head       params (cin, trunk, emb, stride, ln_eps bits)
                        arrays pn_w[trunk,cin] pn_b[trunk] lin1[emb,trunk] lin2[emb,emb] ln_g ln_b

BatchNorm layers (eval mode) are folded into the preceding convolution:
scale = gamma / sqrt(running_var + eps), W' = W * scale, b' = beta - running_mean * scale.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Dict, List, Mapping, Tuple

import numpy as np

T_STEM, T_PW, T_DW, T_HEAD = 0, 1, 2, 3


@dataclass(frozen=True)
class StudentConfig:
    """student_clap/config.yaml:4-24."""
    n_mels: int = 128
    alpha: float = 3.0
    beta: float = 0.75
    t_zero: float = 6.0
    num_layers: int = 8
    trunk_dim: int = 2048
    embedding_dim: int = 512
    input_hw: Tuple[int, int] = (640, 128)
    bn_eps: float = 1e-3      # micromind BatchNorm2d(eps=1e-3)
    bn0_eps: float = 1e-5     # nn.BatchNorm2d default (student_onnx_model.py:34)
    ln_eps: float = 1e-5


@dataclass(frozen=True)
class BlockPlan:
    index: int        # index in phinet._layers
    block_id: int     # 0 = no expansion conv
    cin: int
    cmid: int
    cout: int
    stride: int
    residual: bool


def correct_pad(input_shape, kernel_size=3):
    """micromind.utils.correct_pad on the (C, H, W) build-time input shape -> (l, r, t, b)."""
    adjust = (1 - input_shape[0] % 2, 1 - input_shape[1] % 2)
    correct = (kernel_size // 2, kernel_size // 2)
    return (int(correct[1] - adjust[1]), int(correct[1]), int(correct[0] - adjust[0]), int(correct[0]))


def expansion_factor(t_zero, beta, block_id, num_blocks):
    return (t_zero * beta) * block_id / num_blocks + t_zero * (num_blocks - block_id) / num_blocks


def block_plan(cfg: StudentConfig) -> Tuple[int, List[BlockPlan]]:
    """Channel / stride plan of micromind PhiNet(compatibility=True): returns (stem_cout, blocks)."""
    a, N = cfg.alpha, cfg.num_layers
    c0 = int(48 * a)
    spec = [(0, 1.0, 1, int(24 * a)),
            (1, expansion_factor(cfg.t_zero, cfg.beta, 1, N), 2, int(24 * a)),
            (2, expansion_factor(cfg.t_zero, cfg.beta, 2, N), 1, int(24 * a)),
            (3, expansion_factor(cfg.t_zero, cfg.beta, 3, N), 2, int(48 * a))]
    bf = 48
    for bid in range(4, N + 1):
        ds = bid in (5, 7)
        if ds:
            bf *= 2
        spec.append((bid, expansion_factor(cfg.t_zero, cfg.beta, bid, N), 2 if ds else 1, int(bf * a)))
    blocks, cin = [], c0
    for i, (bid, exp, stride, filt) in enumerate(spec):
        cmid = int(exp * cin) if bid else cin
        blocks.append(BlockPlan(i + 2, bid, cin, cmid, filt, stride, cin == filt and stride == 1))
        cin = filt
    return c0, blocks


def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def _fold(sd: Mapping, prefix: str, eps: float):
    g, b = _np(sd[prefix + ".weight"]), _np(sd[prefix + ".bias"])
    mu, var = _np(sd[prefix + ".running_mean"]), _np(sd[prefix + ".running_var"])
    scale = g / np.sqrt(var + eps)
    return scale, b - mu * scale


def _arr(a) -> bytes:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    return struct.pack("<Q", a.size) + a.tobytes()


def _rec(rtype: int, params, arrays) -> bytes:
    p = list(params) + [0] * (8 - len(params))
    return struct.pack("<I8i", rtype, *p) + b"".join(_arr(a) for a in arrays)


def export_blob(state_dict: Mapping, cfg: StudentConfig = StudentConfig()) -> bytes:
    """StudentCLAPAudio.state_dict() -> AMW1 blob (see module docstring)."""
    sd = state_dict
    c0, blocks = block_plan(cfg)
    recs = []
    # ---- stem: bn0, ZeroPad2d(correct_pad), SeparableConv2d(dw 3x3 s2, pw 1x1, BN, ReLU6)
    s0, sh0 = _fold(sd, "phinet.bn0", cfg.bn0_eps)
    pl, pr, pt, pb = correct_pad((1,) + tuple(cfg.input_hw))
    dw = _np(sd["phinet._layers.1._layers.0.weight"]).reshape(9)
    pw = _np(sd["phinet._layers.1._layers.1.weight"]).reshape(-1)
    bs, bsh = _fold(sd, "phinet._layers.1._layers.2", cfg.bn_eps)
    if pw.shape[0] != c0:
        raise ValueError(f"stem has {pw.shape[0]} channels, plan says {c0}")
    recs.append(_rec(T_STEM, (c0, pt, pb, pl, pr), (s0, sh0, dw, pw * bs, bsh)))
    # ---- inverted-residual blocks
    for blk in blocks:
        base = f"phinet._layers.{blk.index}._layers."
        j = 0
        first = True
        if blk.block_id:
            w = _np(sd[base + "0.weight"]).reshape(blk.cmid, blk.cin)
            s, sh = _fold(sd, base + "1", cfg.bn_eps)
            recs.append(_rec(T_PW, (blk.cin, blk.cmid, 1, 0, 1), (w * s[:, None], sh)))
            first = False
            j = 3
        j += 1  # Dropout2d (identity at inference, no parameters)
        w = _np(sd[base + f"{j}.weight"]).reshape(blk.cmid, 9)
        s, sh = _fold(sd, base + f"{j + 1}", cfg.bn_eps)
        recs.append(_rec(T_DW, (blk.cmid, blk.stride, 0, 0, 1 if first else 0), (w * s[:, None], sh)))
        j += 3
        w = _np(sd[base + f"{j}.weight"]).reshape(blk.cout, blk.cmid)
        s, sh = _fold(sd, base + f"{j + 1}", cfg.bn_eps)
        recs.append(_rec(T_PW, (blk.cmid, blk.cout, 0, 1 if blk.residual else 0, 0), (w * s[:, None], sh)))
    # ---- head: pn_block (1x1 stride 2, bias) -> mean -> Projection -> (L2 in the kernel)
    cin = blocks[-1].cout
    pn_w = _np(sd["phinet.pn_block.weight"]).reshape(cfg.trunk_dim, cin)
    pn_b = _np(sd["phinet.pn_block.bias"])
    l1 = _np(sd["projection_head.linear1.weight"])
    l2 = _np(sd["projection_head.linear2.weight"])
    g, b = _np(sd["projection_head.layer_norm.weight"]), _np(sd["projection_head.layer_norm.bias"])
    eps_bits = struct.unpack("<i", struct.pack("<f", cfg.ln_eps))[0]
    recs.append(_rec(T_HEAD, (cin, cfg.trunk_dim, cfg.embedding_dim, 2, eps_bits), (pn_w, pn_b, l1, l2, g, b)))
    head = b"AMW1" + struct.pack("<4I", 1, cfg.n_mels, cfg.embedding_dim, len(recs))
    return head + b"".join(recs)


def random_state_dict(seed: int = 0, cfg: StudentConfig = StudentConfig()) -> Dict[str, np.ndarray]:
    """Seeded random-init weights with StudentCLAPAudio's state_dict names (no checkpoint is
    available offline).  He-scaled convolutions, BatchNorm statistics chosen so that
    activations stay O(1): bn0 standardises log-mel dB (mean -30, var 400)."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}

    def bn(prefix, c, mean=0.0, var=1.0):
        sd[prefix + ".weight"] = rng.uniform(0.8, 1.2, c).astype(np.float32)
        sd[prefix + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[prefix + ".running_mean"] = (mean + 0.05 * rng.standard_normal(c)).astype(np.float32)
        sd[prefix + ".running_var"] = (var * rng.uniform(0.8, 1.2, c)).astype(np.float32)

    def conv(name, cout, cin_g, k, gain=2.0):
        std = np.sqrt(gain / (cin_g * k * k))
        sd[name] = (std * rng.standard_normal((cout, cin_g, k, k))).astype(np.float32)

    c0, blocks = block_plan(cfg)
    bn("phinet.bn0", cfg.n_mels, mean=-30.0, var=400.0)
    conv("phinet._layers.1._layers.0.weight", 1, 1, 3, gain=1.0)
    conv("phinet._layers.1._layers.1.weight", c0, 1, 1, gain=2.0)
    bn("phinet._layers.1._layers.2", c0)
    for blk in blocks:
        base = f"phinet._layers.{blk.index}._layers."
        j = 0
        if blk.block_id:
            conv(base + "0.weight", blk.cmid, blk.cin, 1)
            bn(base + "1", blk.cmid)
            j = 3
        j += 1
        conv(base + f"{j}.weight", blk.cmid, 1, 3)
        bn(base + f"{j + 1}", blk.cmid)
        j += 3
        conv(base + f"{j}.weight", blk.cout, blk.cmid, 1, gain=1.0)
        bn(base + f"{j + 1}", blk.cout)
    cin = blocks[-1].cout
    conv("phinet.pn_block.weight", cfg.trunk_dim, cin, 1, gain=1.0)
    sd["phinet.pn_block.bias"] = (0.1 * rng.standard_normal(cfg.trunk_dim)).astype(np.float32)
    sd["projection_head.linear1.weight"] = (rng.standard_normal((cfg.embedding_dim, cfg.trunk_dim)) /
                                             np.sqrt(cfg.trunk_dim)).astype(np.float32)
    sd["projection_head.linear2.weight"] = (rng.standard_normal((cfg.embedding_dim, cfg.embedding_dim)) /
                                             np.sqrt(cfg.embedding_dim)).astype(np.float32)
    sd["projection_head.layer_norm.weight"] = rng.uniform(0.8, 1.2, cfg.embedding_dim).astype(np.float32)
    sd["projection_head.layer_norm.bias"] = (0.05 * rng.standard_normal(cfg.embedding_dim)).astype(np.float32)
    return sd


def count_macs(cfg: StudentConfig = StudentConfig(), T: int = 1001) -> Dict[str, int]:
    """Multiply-accumulates per 10 s window by layer kind (DESIGN.md roofline accounting)."""
    c0, blocks = block_plan(cfg)
    pl, pr, pt, pb = correct_pad((1,) + tuple(cfg.input_hw))
    H = (T + pt + pb - 3) // 2 + 1
    W = (cfg.n_mels + pl + pr - 3) // 2 + 1
    macs = {"stem": H * W * (9 + c0), "pointwise": 0, "depthwise": 0, "head": 0}
    for blk in blocks:
        if blk.block_id:
            macs["pointwise"] += H * W * blk.cin * blk.cmid
        H, W = (H + 2 - 3) // blk.stride + 1, (W + 2 - 3) // blk.stride + 1
        macs["depthwise"] += H * W * blk.cmid * 9
        macs["pointwise"] += H * W * blk.cmid * blk.cout
    cin = blocks[-1].cout
    macs["head"] = cin * cfg.trunk_dim + cfg.trunk_dim * cfg.embedding_dim + cfg.embedding_dim ** 2
    return macs
