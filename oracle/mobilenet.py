"""Oracle: an EfficientAT-style (MobileNetV3) audio encoder in PyTorch-CPU fp32.  TEST INFRASTRUCTURE ONLY.

``config.py:375,382`` and ``tasks/clap_analyzer.py:461`` call the deployed student ("model_epoch_36.onnx")
an *EfficientAT* model, while the only definition in the tree (``student_clap/models/student_onnx_model.py``)
builds a PhiNet.  The file itself is not available offline, so the engine must not assume either: this
module restates the OTHER family -- EfficientAT's ``mn`` networks are torchvision-style MobileNetV3
(Howard et al. 2019; Schmid et al., "Efficient Large-scale Audio Tagging", 2023): 3x3 stride-2 stem with
hardswish, inverted-residual blocks with 3x3 / 5x5 depthwise kernels, ReLU or hardswish, squeeze-excite
(avg-pool -> 1x1 -> ReLU -> 1x1 -> hardsigmoid -> scale), a 1x1 "last conv" with hardswish, global average
pooling, and a Linear -> Hardswish -> Linear head -- wrapped like ``StudentCLAPAudio`` (input
``(B,1,n_mels,T)`` taken as a plain NCHW image: H = mel, W = time; output L2-normalised 512-d).

It exists so the tests can export a second, structurally different graph with the reference's exporter
arguments (``student_onnx_model.py:611-626``) and show the loader is graph-driven.  PARITY UNPINNED
against ``model_epoch_36.onnx`` for the same reason as ``oracle/phinet.py``: weights are seeded-random.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _make_divisible(v: float, divisor: int = 8) -> int:
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


@dataclass(frozen=True)
class MNConfig:
    n_mels: int = 128
    width_mult: float = 1.0
    embedding_dim: int = 512
    head_dim: int = 1280
    # (kernel, expanded, out, squeeze-excite, hardswish, stride): MobileNetV3-large rows (torchvision)
    rows: Tuple[Tuple[int, int, int, bool, bool, int], ...] = (
        (3, 16, 16, False, False, 1),
        (3, 64, 24, False, False, 2),
        (3, 72, 24, False, False, 1),
        (5, 72, 40, True, False, 2),
        (5, 120, 40, True, False, 1),
        (5, 120, 40, True, False, 1),
        (3, 240, 80, False, True, 2),
        (3, 200, 80, False, True, 1),
        (3, 184, 80, False, True, 1),
        (3, 184, 80, False, True, 1),
        (3, 480, 112, True, True, 1),
        (3, 672, 112, True, True, 1),
        (5, 672, 160, True, True, 2),
        (5, 960, 160, True, True, 1),
        (5, 960, 160, True, True, 1),
    )


SMALL_ROWS = (
    (3, 16, 16, True, False, 2),
    (3, 72, 24, False, False, 2),
    (5, 88, 24, False, False, 1),
    (5, 96, 40, True, True, 2),
    (5, 240, 40, True, True, 1),
    (3, 120, 48, True, True, 1),
)


class SqueezeExcite(nn.Module):
    def __init__(self, c: int, squeeze: int):
        super().__init__()
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(c, squeeze, 1)
        self.fc2 = nn.Conv2d(squeeze, c, 1)
        self.activation = nn.ReLU()
        self.scale_activation = nn.Hardsigmoid()

    def forward(self, x):
        s = self.scale_activation(self.fc2(self.activation(self.fc1(self.avgpool(x)))))
        return s * x


def _conv_bn_act(cin, cout, k, stride, groups, act):
    layers = [nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=False),
              nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01)]
    if act is not None:
        layers.append(act())
    return nn.Sequential(*layers)


class InvertedResidual(nn.Module):
    def __init__(self, cin, k, exp, cout, use_se, use_hs, stride):
        super().__init__()
        self.use_res = stride == 1 and cin == cout
        act = nn.Hardswish if use_hs else nn.ReLU
        layers = []
        if exp != cin:
            layers.append(_conv_bn_act(cin, exp, 1, 1, 1, act))
        layers.append(_conv_bn_act(exp, exp, k, stride, exp, act))
        if use_se:
            layers.append(SqueezeExcite(exp, _make_divisible(exp // 4, 8)))
        layers.append(_conv_bn_act(exp, cout, 1, 1, 1, None))
        self.block = nn.Sequential(*layers)

    def forward(self, x):
        y = self.block(x)
        return x + y if self.use_res else y


class MobileNetAudio(nn.Module):
    """(B,1,n_mels,T) -> (B, embedding_dim), L2-normalised."""

    def __init__(self, cfg: MNConfig = MNConfig()):
        super().__init__()
        self.cfg = cfg
        w = cfg.width_mult
        c0 = _make_divisible(16 * w)
        feats = [_conv_bn_act(1, c0, 3, 2, 1, nn.Hardswish)]
        cin = c0
        for k, exp, cout, se, hs, s in cfg.rows:
            exp, cout = _make_divisible(exp * w), _make_divisible(cout * w)
            feats.append(InvertedResidual(cin, k, exp, cout, se, hs, s))
            cin = cout
        last = 6 * cin
        feats.append(_conv_bn_act(cin, last, 1, 1, 1, nn.Hardswish))
        self.features = nn.Sequential(*feats)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Sequential(nn.Linear(last, cfg.head_dim), nn.Hardswish(),
                                        nn.Linear(cfg.head_dim, cfg.embedding_dim))

    def forward(self, mel_spec):
        x = self.features(mel_spec)
        x = torch.flatten(self.avgpool(x), 1)
        return F.normalize(self.classifier(x), p=2, dim=1)


def make_random_mobilenet(seed=0, cfg: MNConfig = MNConfig(), calib_T=201, calib_batch=2):
    """Seeded random network with calibrated BatchNorm statistics (activations stay O(1))."""
    from .phinet import synthetic_mel

    torch.manual_seed(seed)
    m = MobileNetAudio(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            with torch.no_grad():
                mod.weight.copy_(0.5 + torch.rand(mod.weight.shape, generator=g))
                mod.bias.copy_(0.2 * torch.randn(mod.bias.shape, generator=g))
                mod.momentum = 1.0
    m.train()
    with torch.no_grad():
        # log-mel dB values are O(-30 +- 20): standardise inside the stem's BatchNorm via calibration
        m(synthetic_mel(calib_batch, cfg.n_mels, calib_T, seed + 2))
    m.eval()
    return m


@torch.no_grad()
def embed_segments(model, mels):
    import numpy as np
    mels = torch.as_tensor(np.asarray(mels, dtype=np.float32))
    return torch.cat([model(mels[i:i + 1]) for i in range(mels.shape[0])], 0).numpy()
