"""Minimal ONNX reader / writer (protobuf wire format, no `onnx` package needed).

The reference builds its engines from `refiner_hwc.onnx` / `scorer_hwc.onnx` with trtexec
(tools/cvt_onnx2trt.bash:3-15); the B200 path needs only the trained parameters out of those files.
This module reads `ModelProto.graph.initializer` (name, dims, data) and `graph.node`
(op_type, inputs, outputs) — enough for weights.import_onnx to map parameters onto the layer
schedule — and can write a graph of initializers + nodes, which the tests use to build fixtures.

Field numbers follow onnx.proto3: ModelProto.graph = 7; GraphProto.node = 1, .initializer = 5;
NodeProto.input = 1, .output = 2, .name = 3, .op_type = 4; TensorProto.dims = 1, .data_type = 2,
.float_data = 4, .int32_data = 5 (fp16 bit patterns), .name = 8, .raw_data = 9.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np

FLOAT, FLOAT16 = 1, 10
_DTYPES = {FLOAT: np.float32, FLOAT16: np.float16, 11: np.float64, 6: np.int32, 7: np.int64}


# ----------------------------------------------------------------------------- wire format
def _varint(buf: bytes, p: int) -> Tuple[int, int]:
    v, shift = 0, 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, p
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yields (field number, wire type, value); length-delimited values are memoryview slices."""
    p, n = 0, len(buf)
    mv = memoryview(buf)
    while p < n:
        key, p = _varint(buf, p)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _varint(buf, p)
        elif wt == 1:
            v, p = struct.unpack_from("<Q", buf, p)[0], p + 8
        elif wt == 2:
            ln, p = _varint(buf, p)
            if p + ln > n:
                raise ValueError("truncated ONNX file")
            v, p = mv[p:p + ln], p + ln
        elif wt == 5:
            v, p = struct.unpack_from("<I", buf, p)[0], p + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


def _packed_varints(b: bytes) -> List[int]:
    out, p = [], 0
    while p < len(b):
        v, p = _varint(b, p)
        out.append(v)
    return out


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw, floats, ints = FLOAT, "", None, [], []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            dims += _packed_varints(bytes(v)) if wt == 2 else [v]
        elif fn == 2:
            dtype = v
        elif fn == 4:
            floats.append(np.frombuffer(bytes(v), "<f4") if wt == 2 else np.array([struct.unpack("<f", struct.pack("<I", v))[0]], "<f4"))
        elif fn == 5:
            ints += _packed_varints(bytes(v)) if wt == 2 else [v]
        elif fn == 8:
            name = bytes(v).decode()
        elif fn == 9:
            raw = bytes(v)
    if dtype not in _DTYPES:
        raise ValueError(f"initializer '{name}': unsupported ONNX data_type {dtype}")
    if raw is not None:
        arr = np.frombuffer(raw, _DTYPES[dtype])
    elif floats:
        arr = np.concatenate(floats).astype(_DTYPES[dtype])
    elif ints and dtype == FLOAT16:
        arr = np.array(ints, np.uint16).view(np.float16)
    else:
        arr = np.array(ints, _DTYPES[dtype])
    return name, arr.reshape(dims).copy()


def read_onnx(path: str):
    """-> (initializers {name: ndarray}, nodes [(op_type, [inputs], [outputs])]) in graph order."""
    with open(path, "rb") as f:
        buf = f.read()
    graph = None
    for fn, wt, v in _fields(buf):
        if fn == 7 and wt == 2:
            graph = bytes(v)
    if graph is None:
        raise ValueError(f"{path}: no GraphProto (not an ONNX model?)")
    inits: Dict[str, np.ndarray] = {}
    nodes = []
    for fn, wt, v in _fields(graph):
        if fn == 5 and wt == 2:
            name, arr = _tensor(bytes(v))
            inits[name] = arr
        elif fn == 1 and wt == 2:
            op, ins, outs = "", [], []
            for f2, w2, v2 in _fields(bytes(v)):
                if f2 == 1:
                    ins.append(bytes(v2).decode())
                elif f2 == 2:
                    outs.append(bytes(v2).decode())
                elif f2 == 4:
                    op = bytes(v2).decode()
            nodes.append((op, ins, outs))
    return inits, nodes


# ----------------------------------------------------------------------------- writer (fixtures)
def _enc_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(fn: int, payload: bytes) -> bytes:
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def write_onnx(path: str, initializers: Dict[str, np.ndarray], nodes=()) -> None:
    g = bytearray()
    for op, ins, outs in nodes:
        n = b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs) + _ld(4, op.encode())
        g += _ld(1, n)
    for name, arr in initializers.items():
        arr = np.ascontiguousarray(arr)
        dt = FLOAT16 if arr.dtype == np.float16 else FLOAT
        t = b"".join(_enc_varint((1 << 3) | 0) + _enc_varint(int(d)) for d in arr.shape)
        t += _enc_varint((2 << 3) | 0) + _enc_varint(dt) + _ld(8, name.encode())
        t += _ld(9, arr.astype(np.float16 if dt == FLOAT16 else np.float32).tobytes())
        g += _ld(5, t)
    model = _enc_varint((1 << 3) | 0) + _enc_varint(8) + _ld(7, bytes(g))  # ir_version 8
    with open(path, "wb") as f:
        f.write(model)
