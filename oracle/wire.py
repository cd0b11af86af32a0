"""Pure-Python protobuf wire restatement of the messages on the Predict path.

Field numbers follow the reference's generated code:
  proto/tensorflow/serving/predict.pb.go:30-43   PredictRequest{model_spec=1, inputs=2, output_filter=3}
  proto/tensorflow/serving/predict.pb.go:98-100  PredictResponse{outputs=1, model_spec=2}
  proto/tensorflow/serving/model.pb.go:27-91     ModelSpec{name=1, version=2 (Int64Value{value=1}),
                                                  signature_name=3, version_label=4}
  proto/tensorflow/core/framework/tensor.pb.go:25-68       TensorProto
  proto/tensorflow/core/framework/tensor_shape.pb.go:38-96 TensorShapeProto{dim=2{size=1,name=2}}
  proto/tensorflow/core/framework/types.pb.go:30-55        DataType
"""
from __future__ import annotations

import struct

import numpy as np

DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64, DT_HALF, DT_BFLOAT16 = 1, 2, 3, 9, 19, 14
_NP = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_INT64: np.int64}


def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _ld(field: int, payload: bytes) -> bytes:
    return _tag(field, 2) + _varint(len(payload)) + payload


def _read_varint(buf: bytes, pos: int):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _fields(buf: bytes):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield field, wt, v


def encode_model_spec(name: str, version: int | None, signature_name: str = "") -> bytes:
    out = b""
    if name:
        out += _ld(1, name.encode())
    if version is not None:
        inner = (_tag(1, 0) + _varint(version)) if version != 0 else b""
        out += _ld(2, inner)
    if signature_name:
        out += _ld(3, signature_name.encode())
    return out


def decode_model_spec(buf: bytes):
    name, version, sig = "", None, ""
    for f, wt, v in _fields(buf):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:
            version = 0
            for f2, _wt2, v2 in _fields(bytes(v)):
                if f2 == 1:
                    version = v2 - (1 << 64) if v2 >= (1 << 63) else v2
        elif f == 3:
            sig = bytes(v).decode()
    return name, version, sig


def encode_shape(shape) -> bytes:
    out = b""
    for d in shape:
        dim = (_tag(1, 0) + _varint(int(d))) if d != 0 else b""
        out += _ld(2, dim)
    return out


def encode_tensor(arr: np.ndarray, use_content: bool = True) -> bytes:
    dt = {np.dtype(np.float32): DT_FLOAT, np.dtype(np.float64): DT_DOUBLE,
          np.dtype(np.int32): DT_INT32, np.dtype(np.int64): DT_INT64}[arr.dtype]
    out = _tag(1, 0) + _varint(dt) + _ld(2, encode_shape(arr.shape))
    a = np.ascontiguousarray(arr)
    if use_content:
        out += _ld(4, a.tobytes())
    elif dt == DT_FLOAT:
        out += _ld(5, a.astype("<f4").tobytes())
    elif dt == DT_DOUBLE:
        out += _ld(6, a.astype("<f8").tobytes())
    elif dt == DT_INT32:
        out += _ld(7, b"".join(_varint(int(x)) for x in a.ravel()))
    elif dt == DT_INT64:
        out += _ld(10, b"".join(_varint(int(x)) for x in a.ravel()))
    return out


def decode_tensor(buf: bytes) -> np.ndarray:
    dt, shape, content, vals = 0, [], None, []
    for f, wt, v in _fields(buf):
        if f == 1:
            dt = v
        elif f == 2:
            for f2, _w, v2 in _fields(bytes(v)):
                if f2 == 2:
                    size = 0
                    for f3, _w3, v3 in _fields(bytes(v2)):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
                    shape.append(size)
        elif f == 4:
            content = bytes(v)
        elif f == 5:
            vals += list(struct.unpack("<f", bytes(v))) if wt == 5 else list(np.frombuffer(bytes(v), "<f4"))
        elif f == 6:
            vals += list(struct.unpack("<d", bytes(v))) if wt == 1 else list(np.frombuffer(bytes(v), "<f8"))
        elif f in (7, 10):
            if wt == 0:
                vals.append(v)
            else:
                pos, b = 0, bytes(v)
                while pos < len(b):
                    x, pos = _read_varint(b, pos)
                    vals.append(x)
    npdt = _NP[dt]
    n = int(np.prod(shape)) if shape else 1
    if content is not None and len(content):
        return np.frombuffer(content, dtype=npdt).reshape(shape).copy()
    if f in (7, 10) or dt in (DT_INT32, DT_INT64):
        vals = [x - (1 << 64) if x >= (1 << 63) else x for x in vals]
    a = np.array(vals, dtype=npdt)
    if a.size == 1 and n > 1:  # TF semantics: a single value fills the tensor
        a = np.full(n, a[0], dtype=npdt)
    return a.reshape(shape)


def _map_entry(key: str, tensor_bytes: bytes) -> bytes:
    return _ld(1, key.encode()) + _ld(2, tensor_bytes)


def encode_predict_request(name, version, inputs: dict, signature_name="", output_filter=(),
                           use_content=True) -> bytes:
    out = _ld(1, encode_model_spec(name, version, signature_name))
    for k, arr in inputs.items():
        out += _ld(2, _map_entry(k, encode_tensor(arr, use_content)))
    for f in output_filter:
        out += _ld(3, f.encode())
    return out


def decode_predict_request(buf: bytes):
    spec, inputs, filt = ("", None, ""), {}, []
    for f, _wt, v in _fields(buf):
        if f == 1:
            spec = decode_model_spec(bytes(v))
        elif f == 2:
            key, t = "", None
            for f2, _w2, v2 in _fields(bytes(v)):
                if f2 == 1:
                    key = bytes(v2).decode()
                elif f2 == 2:
                    t = decode_tensor(bytes(v2))
            inputs[key] = t
        elif f == 3:
            filt.append(bytes(v).decode())
    return spec, inputs, filt


def encode_predict_response(name, version, outputs: dict, signature_name="") -> bytes:
    out = b""
    for k, arr in outputs.items():
        out += _ld(1, _map_entry(k, encode_tensor(arr, True)))
    out += _ld(2, encode_model_spec(name, version, signature_name))
    return out


def decode_predict_response(buf: bytes):
    spec, outputs = ("", None, ""), {}
    for f, _wt, v in _fields(buf):
        if f == 2:
            spec = decode_model_spec(bytes(v))
        elif f == 1:
            key, t = "", None
            for f2, _w2, v2 in _fields(bytes(v)):
                if f2 == 1:
                    key = bytes(v2).decode()
                elif f2 == 2:
                    t = decode_tensor(bytes(v2))
            outputs[key] = t
    return spec, outputs
