"""Test helper: write an oracle PyTorch module to an ONNX file with the reference's exporter arguments.

``student_clap/models/student_onnx_model.py:611-626`` calls
``torch.onnx.export(model, dummy(1,1,128,1000), opset_version=17, do_constant_folding=True,
input_names=['mel_spectrogram'], output_names=['embedding'], dynamic_axes={... 3: 'time_frames'})``.
The Python front end of ``torch.onnx.export`` imports the ``onnx`` package (absent here); the TorchScript
exporter's graph construction and the C++ protobuf serialiser it ends in do not, so this helper calls
those two stages directly with the same arguments.  Used by tests only (the file is then read by
``am_clap_load`` and by ``oracle/onnx_ref.py``).
"""
from __future__ import annotations

import os
import warnings

import torch


def export_onnx_bytes(model: torch.nn.Module, n_mels: int = 128, frames: int = 1000) -> bytes:
    from torch.onnx._internal.torchscript_exporter import utils as U

    model.eval()
    x = torch.randn(1, 1, n_mels, frames)
    dyn = {"mel_spectrogram": {3: "time_frames"}}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        graph, params, _ = U._model_to_graph(model, (x,), input_names=["mel_spectrogram"], output_names=["embedding"],
                                             do_constant_folding=True, dynamic_axes=dyn)
        proto = graph._export_onnx(params, 17, dyn, False, torch.onnx.OperatorExportTypes.ONNX, True, True, {}, True,
                                   "", {})[0]
    return bytes(proto)


def export_onnx(model: torch.nn.Module, path: str, n_mels: int = 128, frames: int = 1000, external_data: bool = False) -> str:
    """Writes `path` (and `path + '.data'` with every tensor >= 1 KiB moved out when external_data)."""
    data = export_onnx_bytes(model, n_mels, frames)
    if external_data:
        from . import onnx_rewrite
        data, blob = onnx_rewrite.externalize(data, os.path.basename(path) + ".data")
        with open(path + ".data", "wb") as f:
            f.write(blob)
    with open(path, "wb") as f:
        f.write(data)
    return path
