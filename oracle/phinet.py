"""Oracle: student CLAP audio encoder in PyTorch-CPU fp32.  TEST INFRASTRUCTURE ONLY.

Restates ``student_clap/models/student_onnx_model.py``:
  * :29-59   PhiNet subclass: bn0 over the mel axis, micromind PhiNet trunk, 1x1 stride-2
             ``pn_block`` conv to 2048 channels, spatial mean;
  * :61-74   Projection: e1 = linear1(x); e2 = linear2(gelu(e1)); LayerNorm(e1 + e2);
  * :166-287 StudentCLAPAudio: (B,1,n_mels,T) -> squeeze/transposes -> PhiNet ->
             projection_head -> F.normalize(p=2);
with the configuration of ``student_clap/config.yaml:15-24`` (alpha=3.0, beta=0.75,
t0=6, N=8, compatibility=True => ReLU6 activations, no squeeze-excite).

The trunk ``micromind.networks.PhiNet`` is a third-party dependency (unpinned in
student_clap/requirements.txt:36) that is absent from /root/reference; its published
architecture (Paissan et al., "PhiNets", 2022; micromind/networks/phinet.py) is restated
here: ZeroPad2d(correct_pad) -> SeparableConv2d(3x3 dw stride 2, 1x1 pw, BN eps=1e-3,
act) -> PhiNetConvBlock x (N+1) with expansion factor
t0*beta*id/N + t0*(N-id)/N, filters 24a/24a/24a/48a then 48a doubling at blocks 5 and 7,
stride 2 at blocks 1, 3, 5, 7, residual when stride 1 and in==out.  Module and
parameter names follow micromind's, so ``state_dict()`` has the keys a real
StudentCLAPAudio checkpoint has.

PARITY UNPINNED: neither micromind nor the shipped ``model_epoch_36.onnx`` is
available here, so the restatement cannot be checked against the reference's own
outputs; weights are seeded-random with calibrated BatchNorm statistics.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass(frozen=True)
class StudentConfig:
    n_mels: int = 128
    alpha: float = 3.0
    beta: float = 0.75
    t_zero: float = 6.0
    num_layers: int = 8
    trunk_dim: int = 2048       # pn_block output channels ("embedding_dim=2048")
    embedding_dim: int = 512
    input_hw: tuple = (640, 128)  # input_shape=(1, 640, n_mels) at build time (:237)


def correct_pad(input_shape, kernel_size):
    """micromind.utils.correct_pad (Keras-style 'same' fix-up for stride-2 convs).
    ``input_shape`` is the (C, H, W) tuple PhiNet was built with."""
    if isinstance(kernel_size, int):
        kernel_size = (kernel_size, kernel_size)
    adjust = (1 - input_shape[0] % 2, 1 - input_shape[1] % 2)
    correct = (kernel_size[0] // 2, kernel_size[1] // 2)
    return (
        int(correct[1] - adjust[1]),
        int(correct[1]),
        int(correct[0] - adjust[0]),
        int(correct[0]),
    )


def xpansion_factor(t_zero, beta, block_id, num_blocks):
    return (t_zero * beta) * block_id / num_blocks + t_zero * (num_blocks - block_id) / num_blocks


class ReLUMax(nn.Module):
    def __init__(self, max_value):
        super().__init__()
        self.max = max_value

    def forward(self, x):
        return torch.clamp(x, min=0, max=self.max)


class SeparableConv2d(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self._layers = nn.ModuleList([
            nn.Conv2d(cin, cin, 3, stride=stride, padding=0, groups=cin, bias=False),
            nn.Conv2d(cin, cout, 1, bias=False),
            nn.BatchNorm2d(cout, eps=1e-3, momentum=0.999),
            ReLUMax(6),
        ])

    def forward(self, x):
        for l in self._layers:
            x = l(x)
        return x


class PhiNetConvBlock(nn.Module):
    def __init__(self, cin, expansion, stride, filters, block_id):
        super().__init__()
        self.skip_conn = False
        layers = []
        cmid = cin
        if block_id:
            cmid = int(expansion * cin)
            layers += [nn.Conv2d(cin, cmid, 1, bias=False),
                       nn.BatchNorm2d(cmid, eps=1e-3, momentum=0.999), ReLUMax(6)]
        layers += [nn.Dropout2d(0.05)]
        layers += [nn.Conv2d(cmid, cmid, 3, stride=stride, padding=1, groups=cmid, bias=False),
                   nn.BatchNorm2d(cmid, eps=1e-3, momentum=0.999), ReLUMax(6)]
        layers += [nn.Conv2d(cmid, filters, 1, bias=False),
                   nn.BatchNorm2d(filters, eps=1e-3, momentum=0.999)]
        self._layers = nn.ModuleList(layers)
        if cin == filters and stride == 1:
            self.skip_conn = True

    def forward(self, x):
        inp = x
        for l in self._layers:
            x = l(x)
        return x + inp if self.skip_conn else x


class PhiNet(nn.Module):
    """micromind PhiNet(compatibility=True) + the reference subclass's bn0 / pn_block."""

    def __init__(self, cfg: StudentConfig):
        super().__init__()
        a = cfg.alpha
        first_conv_filters, b1_filters, b2_filters = 48, 24, 48
        c0 = int(first_conv_filters * a)
        layers = [nn.ZeroPad2d(correct_pad((1,) + tuple(cfg.input_hw), 3)),
                  SeparableConv2d(1, c0, stride=2)]
        N = cfg.num_layers
        cin = c0
        plan = [
            (None, 1, 1, int(b1_filters * a)),
            (1, xpansion_factor(cfg.t_zero, cfg.beta, 1, N), 2, int(b1_filters * a)),
            (2, xpansion_factor(cfg.t_zero, cfg.beta, 2, N), 1, int(b1_filters * a)),
            (3, xpansion_factor(cfg.t_zero, cfg.beta, 3, N), 2, int(b2_filters * a)),
        ]
        block_filters = b2_filters
        for bid in range(4, N + 1):
            ds = bid in (5, 7)
            if ds:
                block_filters *= 2
            plan.append((bid, xpansion_factor(cfg.t_zero, cfg.beta, bid, N), 2 if ds else 1,
                         int(block_filters * a)))
        for bid, exp, stride, filt in plan:
            layers.append(PhiNetConvBlock(cin, exp, stride, filt, bid))
            cin = filt
        self._layers = nn.ModuleList(layers)
        self.bn0 = nn.BatchNorm2d(cfg.n_mels)
        self.pn_block = nn.Conv2d(cin, cfg.trunk_dim, kernel_size=1, stride=2)

    def forward(self, x):
        if x.dim() == 3:
            x = x[:, None]
        x = x.transpose(1, 3)
        x = self.bn0(x)
        x = x.transpose(1, 3)
        for l in self._layers:
            x = l(x)
        x = self.pn_block(x)
        return x.mean((-1, -2))


class Projection(nn.Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.linear1 = nn.Linear(d_in, d_out, bias=False)
        self.linear2 = nn.Linear(d_out, d_out, bias=False)
        self.layer_norm = nn.LayerNorm(d_out)

    def forward(self, x):
        e1 = self.linear1(x)
        e2 = self.linear2(F.gelu(e1))
        return self.layer_norm(e1 + e2)


class StudentCLAPAudio(nn.Module):
    def __init__(self, cfg: StudentConfig = StudentConfig()):
        super().__init__()
        self.cfg = cfg
        self.phinet = PhiNet(cfg)
        self.projection_head = Projection(cfg.trunk_dim, cfg.embedding_dim)

    def forward(self, mel_spec):
        """(B,1,n_mels,T) f32 -> (B,512) L2-normalised (student_onnx_model.py:257-287)."""
        if mel_spec.dim() == 4:
            mel_spec = mel_spec.squeeze(1)
        mel_spec = mel_spec.transpose(1, 2)
        feats = self.phinet(mel_spec)
        emb = self.projection_head(feats)
        return F.normalize(emb, p=2, dim=1)


def synthetic_mel(batch, n_mels, T, seed):
    """Plausible log-mel input for calibration: smooth dB surface in [-80, 20]."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(batch, 1, n_mels, T, generator=g) * 8.0
    tilt = torch.linspace(10.0, -50.0, n_mels).view(1, 1, n_mels, 1)
    slow = torch.randn(batch, 1, 1, T, generator=g).cumsum(-1) * 0.5
    return (base + tilt + slow - 20.0).clamp_(-100.0, 30.0)


def make_random_student(seed=0, cfg: StudentConfig = StudentConfig(), calib_T=201, calib_batch=2):
    """Seeded random student with calibrated BatchNorm running statistics (so activations
    stay O(1) through the stack like a trained network) and non-trivial BN affines."""
    torch.manual_seed(seed)
    m = StudentCLAPAudio(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            with torch.no_grad():
                mod.weight.copy_(0.5 + torch.rand(mod.weight.shape, generator=g))
                mod.bias.copy_(0.2 * torch.randn(mod.bias.shape, generator=g))
                mod.momentum = 1.0
        if isinstance(mod, nn.LayerNorm):
            with torch.no_grad():
                mod.weight.copy_(0.5 + torch.rand(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
    m.train()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout2d):
            mod.eval()
    with torch.no_grad():
        m(synthetic_mel(calib_batch, cfg.n_mels, calib_T, seed + 2))
    m.eval()
    return m


@torch.no_grad()
def embed_segments(model, mels, batch1=True):
    """mels f32[S,1,n_mels,T] -> f32[S,512].  batch1 mirrors the reference loop, which
    feeds one segment per session.run (clap_analyzer.py:526-535)."""
    mels = torch.as_tensor(np.asarray(mels, dtype=np.float32))
    if batch1:
        outs = [model(mels[i : i + 1]) for i in range(mels.shape[0])]
        return torch.cat(outs, 0).numpy() if outs else np.zeros((0, model.cfg.embedding_dim), np.float32)
    return model(mels).numpy()


def count_macs(cfg: StudentConfig, T=1001):
    """Multiply-accumulates per segment, per layer kind (for DESIGN.md / roofline)."""
    m = StudentCLAPAudio(cfg)
    macs = {"pointwise": 0, "depthwise": 0, "head": 0}
    hooks = []

    def hook(mod, inp, out):
        if isinstance(mod, nn.Conv2d):
            k = mod.kernel_size[0] * mod.kernel_size[1]
            n = out.numel() // out.shape[0] * (mod.in_channels // mod.groups) * k
            macs["depthwise" if mod.groups > 1 else "pointwise"] += n
        elif isinstance(mod, nn.Linear):
            macs["head"] += mod.in_features * mod.out_features

    for mod in m.modules():
        if isinstance(mod, (nn.Conv2d, nn.Linear)):
            hooks.append(mod.register_forward_hook(hook))
    m.eval()
    with torch.no_grad():
        m(torch.zeros(1, 1, cfg.n_mels, T))
    for h in hooks:
        h.remove()
    return macs
