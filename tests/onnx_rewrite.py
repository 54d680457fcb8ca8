"""Test helper: protobuf-level rewrites of an ONNX file, to cover the encodings a real
``torch.onnx.export(opset_version=17)`` file can use that torch's bare C++ serialiser does not emit:

* ``externalize``   moves tensor payloads into a side file (``model.onnx.data``; the reference's
                    external-data fallback, ``tasks/clap_analyzer.py:132-147``);
* ``attrs_to_inputs`` turns Clip(min=, max=) / Pad(pads=, value=) attributes into constant inputs
                    (the opset >= 11 / 13 encodings) and ReduceMean axes inputs into the ``axes`` attribute
                    (opset 17 encoding).
"""
from __future__ import annotations

import struct
from typing import List, Tuple


def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        if not c & 0x80:
            return r, i
        s += 7


def _enc_varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        c = v & 0x7F
        v >>= 7
        if v:
            out.append(c | 0x80)
        else:
            out.append(c)
            return bytes(out)


def parse(b: bytes) -> List[Tuple[int, int, object]]:
    out, i = [], 0
    while i < len(b):
        key, i = _varint(b, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 1:
            v, i = b[i:i + 8], i + 8
        elif wt == 5:
            v, i = b[i:i + 4], i + 4
        elif wt == 2:
            n, i = _varint(b, i)
            v, i = b[i:i + n], i + n
        else:
            raise ValueError(wt)
        out.append((fn, wt, v))
    return out


def ser(fields) -> bytes:
    out = bytearray()
    for fn, wt, v in fields:
        out += _enc_varint(fn << 3 | wt)
        if wt == 0:
            out += _enc_varint(v)
        elif wt == 2:
            out += _enc_varint(len(v)) + v
        else:
            out += v
    return bytes(out)


def _kv(k: str, v: str) -> bytes:
    return ser([(1, 2, k.encode()), (2, 2, v.encode())])


def _map_graph(model: bytes, fn_graph):
    fields = parse(model)
    return ser([(f, w, fn_graph(parse(v)) if f == 7 and w == 2 else v) for f, w, v in fields])


def externalize(model: bytes, location: str, min_bytes: int = 1024):
    """Returns (model bytes with big initializers pointing into `location`, the side file's bytes)."""
    blob = bytearray()

    def graph(gf):
        out = []
        for f, w, v in gf:
            if f == 5 and w == 2:
                tf = parse(v)
                raw = [x for x in tf if x[0] == 9]
                if raw and len(raw[0][2]) >= min_bytes:
                    payload = raw[0][2]
                    while len(blob) % 64:
                        blob.append(0)
                    off = len(blob)
                    blob.extend(payload)
                    tf = [x for x in tf if x[0] != 9]
                    tf += [(13, 2, _kv("location", location)), (13, 2, _kv("offset", str(off))),
                           (13, 2, _kv("length", str(len(payload)))), (14, 0, 1)]
                    v = ser(tf)
            out.append((f, w, v))
        return ser(out)

    return _map_graph(model, graph), bytes(blob)


def _tensor(name: str, dtype: int, dims, raw: bytes) -> bytes:
    return ser([(1, 0, d) for d in dims] + [(2, 0, dtype), (8, 2, name.encode()), (9, 2, raw)])


def attrs_to_inputs(model: bytes) -> bytes:
    def graph(gf):
        out, new_inits = [], []
        consts = {}
        # Constant nodes holding int64 axes (for the ReduceMean rewrite)
        for f, w, v in gf:
            if f == 1 and w == 2:
                nf = parse(v)
                if [x[2] for x in nf if x[0] == 4] == [b"Constant"]:
                    o = [x[2].decode() for x in nf if x[0] == 2][0]
                    for _, _, av in [x for x in nf if x[0] == 5]:
                        af = parse(av)
                        t = [x[2] for x in af if x[0] == 5]
                        if t:
                            tf = parse(t[0])
                            dt = [x[2] for x in tf if x[0] == 2]
                            raw = [x[2] for x in tf if x[0] == 9]
                            if dt == [7] and raw:
                                consts[o] = list(struct.unpack(f"<{len(raw[0]) // 8}q", raw[0]))
        for f, w, v in gf:
            if f == 1 and w == 2:
                nf = parse(v)
                op = [x[2] for x in nf if x[0] == 4][0].decode()
                outs = [x[2].decode() for x in nf if x[0] == 2]
                attrs = {}
                for _, _, av in [x for x in nf if x[0] == 5]:
                    af = parse(av)
                    attrs[[x[2] for x in af if x[0] == 1][0].decode()] = af
                ins = [x[2].decode() for x in nf if x[0] == 1]
                if op == "Clip" and ("min" in attrs or "max" in attrs) and len(ins) == 1:
                    for key in ("min", "max"):
                        if key in attrs:
                            raw = [x[2] for x in attrs[key] if x[0] == 2][0]
                            nm = f"{outs[0]}_{key}"
                            new_inits.append(_tensor(nm, 1, [], raw))
                            ins.append(nm)
                        else:
                            ins.append("")
                    nf = [x for x in nf if x[0] not in (1, 5)]
                    nf = [(1, 2, i.encode()) for i in ins] + nf
                elif op == "Pad" and "pads" in attrs and len(ins) == 1:
                    pads = []
                    for x in attrs["pads"]:
                        if x[0] == 8:
                            if x[1] == 0:
                                pads.append(x[2])
                            else:
                                j = 0
                                while j < len(x[2]):
                                    pv, j = _varint(x[2], j)
                                    pads.append(pv)
                    nm = f"{outs[0]}_pads"
                    new_inits.append(_tensor(nm, 7, [len(pads)], struct.pack(f"<{len(pads)}q", *pads)))
                    ins.append(nm)
                    keep = [(5, 2, ser(attrs["mode"]))] if "mode" in attrs else []
                    nf = [(1, 2, i.encode()) for i in ins] + [x for x in nf if x[0] not in (1, 5)] + keep
                elif op == "ReduceMean" and len(ins) == 2 and ins[1] in consts:
                    axes = consts[ins[1]]
                    a = ser([(1, 2, b"axes"), (20, 0, 7)] + [(8, 0, x & ((1 << 64) - 1)) for x in axes])
                    nf = [(1, 2, ins[0].encode())] + [x for x in nf if x[0] != 1] + [(5, 2, a)]
                v = ser(nf)
            out.append((f, w, v))
        out += [(5, 2, t) for t in new_inits]
        return ser(out)

    return _map_graph(model, graph)
