"""Oracle: ONNX ``ModelProto`` reader + fp32 graph interpreter (PyTorch CPU).  TEST INFRASTRUCTURE ONLY.

The reference runs its audio encoder through onnxruntime
(``tasks/clap_analyzer.py:109-116`` creates the session from ``CLAP_AUDIO_MODEL_PATH``, ``:534`` calls
``session.run(None, {'mel_spectrogram': mel})``); the file is produced by
``torch.onnx.export(opset 17, do_constant_folding=True, input 'mel_spectrogram', dynamic time axis)``
(``student_clap/models/student_onnx_model.py:611-626``).  Neither ``onnx`` nor ``onnxruntime`` is
installable here, so this module restates what they do for the operator set those exports contain:

* the protobuf wire format of onnx.proto3 (ModelProto / GraphProto / NodeProto / AttributeProto /
  TensorProto, incl. the external-data fallback of ``clap_analyzer.py:132-147``), read by hand;
* operator semantics per the ONNX operator specification (opset 17), each evaluated with the
  equivalent fp32 PyTorch CPU call, node by node in file order -- no fusion, no reordering.

It is independent of the C++ loader in ``audiomuse-ai_b200/csrc/onnx_model.cu`` (different language,
no shared code) and is checked against the PyTorch module that produced the file
(``tests/test_onnx_oracle.py``), which is the same check the reference does at export time
(``student_onnx_model.py:640-650``: max |torch - ort| < 1e-5).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# ------------------------------------------------------------------ protobuf wire format


def _varint(b: bytes, i: int):
    r = 0
    s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        if not c & 0x80:
            return r, i
        s += 7


def _fields(b: bytes):
    """Yields (field_number, wire_type, value) of one message; value is int or bytes."""
    i, n = 0, len(b)
    while i < n:
        key, i = _varint(b, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 1:
            v = b[i:i + 8]
            i += 8
        elif wt == 2:
            ln, i = _varint(b, i)
            v = b[i:i + ln]
            i += ln
        elif wt == 5:
            v = b[i:i + 4]
            i += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


def _sint64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(wt, v) -> List[int]:
    if wt == 0:
        return [_sint64(v)]
    out, i = [], 0
    while i < len(v):
        x, i = _varint(v, i)
        out.append(_sint64(x))
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16,
           11: np.float64}


def _tensor(b: bytes, base_dir: Optional[str]) -> "tuple[str, np.ndarray]":
    dims, dtype, name, raw = [], 1, "", None
    f32, i32, i64, f64 = [], [], [], []
    ext: Dict[str, str] = {}
    location = 0
    for fn, wt, v in _fields(b):
        if fn == 1:
            dims += _packed_varints(wt, v)
        elif fn == 2:
            dtype = v
        elif fn == 4:
            f32 += list(struct.unpack(f"<{len(v) // 4}f", v)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fn == 5:
            i32 += _packed_varints(wt, v)
        elif fn == 7:
            i64 += _packed_varints(wt, v)
        elif fn == 8:
            name = v.decode()
        elif fn == 9:
            raw = v
        elif fn == 10:
            f64 += list(struct.unpack(f"<{len(v) // 8}d", v)) if wt == 2 else [struct.unpack("<d", v)[0]]
        elif fn == 13:
            kv = {f: x for f, _, x in _fields(v)}
            ext[kv[1].decode()] = kv[2].decode()
        elif fn == 14:
            location = v
    np_dt = _DTYPES.get(dtype)
    if np_dt is None:
        raise ValueError(f"tensor {name!r}: unsupported data_type {dtype}")
    if location == 1 or ext:
        if base_dir is None:
            raise ValueError(f"tensor {name!r} uses external data but no directory is known")
        path = os.path.join(base_dir, ext["location"])
        off = int(ext.get("offset", "0"))
        with open(path, "rb") as f:
            f.seek(off)
            raw = f.read(int(ext["length"])) if "length" in ext else f.read()
    if raw is not None:
        a = np.frombuffer(raw, dtype=np.dtype(np_dt).newbyteorder("<")).astype(np_dt)
    elif f32:
        a = np.asarray(f32, dtype=np_dt)
    elif i64:
        a = np.asarray(i64, dtype=np_dt)
    elif i32:
        a = np.asarray(i32).astype(np_dt)
    elif f64:
        a = np.asarray(f64, dtype=np_dt)
    else:
        a = np.zeros(0, dtype=np_dt)
    return name, a.reshape(dims)


@dataclass
class Node:
    op: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, object] = field(default_factory=dict)
    name: str = ""


@dataclass
class Graph:
    nodes: List[Node]
    initializers: Dict[str, np.ndarray]
    inputs: List[str]
    outputs: List[str]
    opset: int = 0


def _attr(b: bytes, base_dir):
    name, val = "", None
    floats, ints = [], []
    atype = 0
    for fn, wt, v in _fields(b):
        if fn == 1:
            name = v.decode()
        elif fn == 2:
            val = struct.unpack("<f", v)[0]
        elif fn == 3:
            val = _sint64(v)
        elif fn == 4:
            val = v.decode(errors="replace")
        elif fn == 5:
            val = _tensor(v, base_dir)[1]
        elif fn == 7:
            floats += list(struct.unpack(f"<{len(v) // 4}f", v)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fn == 8:
            ints += _packed_varints(wt, v)
        elif fn == 20:
            atype = v
    if atype == 6 or (val is None and floats):
        val = floats
    elif atype == 7 or (val is None and ints):
        val = ints
    return name, val


def _node(b: bytes, base_dir) -> Node:
    n = Node("", [], [])
    for fn, wt, v in _fields(b):
        if fn == 1:
            n.inputs.append(v.decode())
        elif fn == 2:
            n.outputs.append(v.decode())
        elif fn == 3:
            n.name = v.decode()
        elif fn == 4:
            n.op = v.decode()
        elif fn == 5:
            k, val = _attr(v, base_dir)
            n.attrs[k] = val
    return n


def _value_name(b: bytes) -> str:
    for fn, _, v in _fields(b):
        if fn == 1:
            return v.decode()
    return ""


def load(path_or_bytes) -> Graph:
    """Reads a ModelProto (file path or bytes).  External tensor data is looked up next to the file."""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        data, base = bytes(path_or_bytes), None
    else:
        with open(path_or_bytes, "rb") as f:
            data = f.read()
        base = os.path.dirname(os.path.abspath(path_or_bytes))
    g = Graph([], {}, [], [])
    for fn, _, v in _fields(data):
        if fn == 8:  # opset_import
            for f2, _, x in _fields(v):
                if f2 == 2:
                    g.opset = max(g.opset, x)
        elif fn == 7:  # graph
            for f2, _, x in _fields(v):
                if f2 == 1:
                    g.nodes.append(_node(x, base))
                elif f2 == 5:
                    name, a = _tensor(x, base)
                    g.initializers[name] = a
                elif f2 == 11:
                    g.inputs.append(_value_name(x))
                elif f2 == 12:
                    g.outputs.append(_value_name(x))
    g.inputs = [n for n in g.inputs if n not in g.initializers]
    return g


# ------------------------------------------------------------------ interpreter (fp32, PyTorch CPU)


def run(g: Graph, feeds: Dict[str, np.ndarray]) -> List[np.ndarray]:
    """Evaluates the graph node by node in fp32.  Raises NotImplementedError on an operator outside the
    set the student / MobileNet-style exports use."""
    import torch
    import torch.nn.functional as F

    env: Dict[str, "torch.Tensor"] = {k: torch.as_tensor(np.asarray(v)) for k, v in g.initializers.items()}
    for k, v in feeds.items():
        env[k] = torch.as_tensor(np.ascontiguousarray(v))

    def ints(x):
        return [int(i) for i in (x.tolist() if hasattr(x, "tolist") else x)]

    with torch.no_grad():
        for n in g.nodes:
            a = n.attrs
            x = [env[i] if i else None for i in n.inputs]
            op = n.op
            if op == "Constant":
                y = torch.as_tensor(np.asarray(a["value"]))
            elif op == "Squeeze":
                axes = ints(x[1]) if len(x) > 1 and x[1] is not None else a.get("axes")
                y = x[0]
                for ax in sorted([ax % x[0].dim() for ax in axes], reverse=True):
                    y = y.squeeze(ax)
            elif op == "Unsqueeze":
                axes = ints(x[1]) if len(x) > 1 and x[1] is not None else a.get("axes")
                y = x[0]
                for ax in sorted(axes):
                    y = y.unsqueeze(ax)
            elif op == "Transpose":
                y = x[0].permute(*a["perm"]).contiguous()
            elif op == "Reshape":
                shp = ints(x[1])
                shp = [x[0].shape[i] if s == 0 else s for i, s in enumerate(shp)]
                y = x[0].reshape(shp)
            elif op == "Flatten":
                ax = a.get("axis", 1)
                y = x[0].reshape(int(np.prod(x[0].shape[:ax])), -1)
            elif op == "BatchNormalization":
                y = F.batch_norm(x[0], x[3], x[4], x[1], x[2], False, 0.0, a.get("epsilon", 1e-5))
            elif op == "Pad":
                pads = ints(x[1]) if len(x) > 1 and x[1] is not None else a["pads"]
                val = float(x[2]) if len(x) > 2 and x[2] is not None else float(a.get("value", 0.0))
                if a.get("mode", "constant") != "constant":
                    raise NotImplementedError("Pad mode " + str(a.get("mode")))
                r = x[0].dim()
                tp = []
                for d in range(r - 1, -1, -1):
                    tp += [pads[d], pads[d + r]]
                y = F.pad(x[0], tp, value=val)
            elif op == "Conv":
                p = a.get("pads", [0, 0, 0, 0])
                xi = x[0]
                if p[0] != p[2] or p[1] != p[3]:
                    xi = F.pad(xi, [p[1], p[3], p[0], p[2]])
                    pad = 0
                else:
                    pad = (p[0], p[1])
                y = F.conv2d(xi, x[1], x[2] if len(x) > 2 else None, stride=tuple(a.get("strides", [1, 1])),
                             padding=pad, dilation=tuple(a.get("dilations", [1, 1])), groups=a.get("group", 1))
            elif op == "Clip":
                lo = x[1] if len(x) > 1 and x[1] is not None else a.get("min")
                hi = x[2] if len(x) > 2 and x[2] is not None else a.get("max")
                y = x[0]
                if lo is not None:
                    y = torch.clamp(y, min=float(lo))
                if hi is not None:
                    y = torch.clamp(y, max=float(hi))
            elif op == "Relu":
                y = torch.relu(x[0])
            elif op == "HardSwish":
                y = x[0] * torch.clamp(x[0] / 6.0 + 0.5, 0.0, 1.0)
            elif op == "HardSigmoid":
                y = torch.clamp(a.get("alpha", 0.2) * x[0] + a.get("beta", 0.5), 0.0, 1.0)
            elif op == "Sigmoid":
                y = torch.sigmoid(x[0])
            elif op == "Add":
                y = x[0] + x[1]
            elif op == "Sub":
                y = x[0] - x[1]
            elif op == "Mul":
                y = x[0] * x[1]
            elif op == "Div":
                y = x[0] / x[1]
            elif op == "Pow":
                y = torch.pow(x[0], x[1])
            elif op == "Sqrt":
                y = torch.sqrt(x[0])
            elif op == "Erf":
                y = torch.erf(x[0])
            elif op == "MatMul":
                y = x[0] @ x[1]
            elif op == "Gemm":
                A = x[0].t() if a.get("transA", 0) else x[0]
                Bm = x[1].t() if a.get("transB", 0) else x[1]
                y = a.get("alpha", 1.0) * (A @ Bm)
                if len(x) > 2 and x[2] is not None:
                    y = y + a.get("beta", 1.0) * x[2]
            elif op == "GlobalAveragePool":
                y = x[0].mean((2, 3), keepdim=True)
            elif op in ("ReduceMean", "ReduceL2", "ReduceSum"):
                axes = ints(x[1]) if len(x) > 1 and x[1] is not None else a.get("axes")
                kd = bool(a.get("keepdims", 1))
                if op == "ReduceMean":
                    y = x[0].mean(axes, keepdim=kd)
                elif op == "ReduceSum":
                    y = x[0].sum(axes, keepdim=kd)
                else:
                    y = torch.sqrt((x[0] * x[0]).sum(axes, keepdim=kd))
            elif op == "LayerNormalization":
                ax = a.get("axis", -1)
                y = F.layer_norm(x[0], tuple(x[0].shape[ax:]), x[1], x[2] if len(x) > 2 else None,
                                 a.get("epsilon", 1e-5))
            elif op == "Shape":
                y = torch.tensor(list(x[0].shape), dtype=torch.int64)
            elif op == "Expand":
                y = x[0].expand(*torch.broadcast_shapes(tuple(x[0].shape), tuple(ints(x[1])))).contiguous()
            elif op == "Identity":
                y = x[0]
            else:
                raise NotImplementedError(f"oracle/onnx_ref: operator {op!r} is not supported")
            env[n.outputs[0]] = y
    return [env[o].numpy() for o in g.outputs]


def describe(g: Graph) -> str:
    lines = [f"opset {g.opset}; inputs {g.inputs}; outputs {g.outputs}; {len(g.nodes)} nodes; "
             f"{len(g.initializers)} initializers"]
    for n in g.nodes:
        at = {k: (tuple(v.shape) if isinstance(v, np.ndarray) else v) for k, v in n.attrs.items()}
        lines.append(f"{n.op:20s} {n.inputs} -> {n.outputs} {at}")
    return "\n".join(lines)
