"""SavedModel importer (tfservingcache_b200/savedmodel.py) against SavedModel directories written here byte by
byte in the public formats (protobuf + LevelDB-style table + tensor bundle) -- there is no TensorFlow in the image
and no SavedModel in the reference repo, so this pins self-consistency with the format definitions only."""
import json
import os
import struct

import numpy as np
import pytest

from oracle import models
from tfservingcache_b200 import savedmodel as sm


def _vi(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(f, p):
    return _vi((f << 3) | 2) + _vi(len(p)) + p


def _v(f, v):
    return _vi(f << 3) + _vi(v)


def _block(entries, restart_interval=2):
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(last), len(k)) and last[shared] == k[shared]:
                shared += 1
        out += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        last = k
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path, items, block_items=3):
    items = sorted(items.items())
    data, index = bytearray(), []
    for i in range(0, len(items), block_items):
        chunk = items[i:i + block_items]
        blk = _block(chunk)
        off = len(data)
        data += blk + b"\x00" + struct.pack("<I", sm.mask_crc(sm.crc32c(blk + b"\x00")))
        index.append((chunk[-1][0], _vi(off) + _vi(len(blk))))
    meta = _block([])
    meta_off = len(data)
    data += meta + b"\x00" + struct.pack("<I", sm.mask_crc(sm.crc32c(meta + b"\x00")))
    idx = _block(index, restart_interval=1)
    idx_off = len(data)
    data += idx + b"\x00" + struct.pack("<I", sm.mask_crc(sm.crc32c(idx + b"\x00")))
    footer = _vi(meta_off) + _vi(len(meta)) + _vi(idx_off) + _vi(len(idx))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", sm.TABLE_MAGIC)
    open(path, "wb").write(bytes(data) + footer)


def write_bundle(prefix, tensors):
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    blob, items = bytearray(), {b"": _v(1, 1) + _ld(3, _v(1, 1))}   # num_shards=1, little endian, version{producer=1}
    for name, arr in tensors.items():
        raw = np.ascontiguousarray(arr, np.float32).tobytes()
        shape = b"".join(_ld(2, _v(1, d)) for d in arr.shape)
        entry = _v(1, 1) + _ld(2, shape) + (_v(4, len(blob)) if len(blob) else b"") + _v(5, len(raw)) + \
            _vi((6 << 3) | 5) + struct.pack("<I", sm.mask_crc(sm.crc32c(raw)))
        items[name.encode()] = entry
        blob += raw
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(blob))
    write_table(prefix + ".index", items)


def write_saved_model(path, nodes, signature):
    graph = b"".join(_ld(1, _ld(1, n.encode()) + _ld(2, op.encode()) + b"".join(_ld(3, i.encode()) for i in ins)) for n, op, ins in nodes)
    in_key, in_t, out_key, out_t = signature

    def tinfo(k, t):
        return _ld(1, k.encode()) + _ld(2, _ld(1, t.encode()) + _v(2, 1))
    sig = _ld(1, tinfo(in_key, in_t)) + _ld(2, tinfo(out_key, out_t)) + _ld(3, b"tensorflow/serving/predict")
    meta = _ld(2, graph) + _ld(5, _ld(1, b"serving_default") + _ld(2, sig))
    open(path, "wb").write(_v(1, 1) + _ld(2, meta))


def test_crc32c_known_answer():
    assert sm.crc32c(b"123456789") == 0xE3069283          # CRC-32C (Castagnoli) check value


def test_table_and_bundle_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f"layer{i}/kernel": rng.standard_normal((5 + i, 3)).astype(np.float32) for i in range(7)}
    tensors["a"] = np.float32(0.5).reshape(())
    write_bundle(str(tmp_path / "variables" / "variables"), tensors)
    got = sm.read_bundle(str(tmp_path / "variables" / "variables"))
    assert set(got) == set(tensors)
    for k in tensors:
        assert np.array_equal(got[k], tensors[k]) and got[k].shape == tensors[k].shape
    # a flipped data byte is caught by the per-tensor checksum
    p = str(tmp_path / "variables" / "variables.data-00000-of-00001")
    raw = bytearray(open(p, "rb").read())
    raw[10] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        sm.read_bundle(str(tmp_path / "variables" / "variables"))


def test_half_plus_two_saved_model_converts_to_affine_bundle(tmp_path):
    """Graph shape of TF-Serving's saved_model_half_plus_two: y = Add(Mul(a, x), b) with scalar variables."""
    d = tmp_path / "saved_model_half_plus_two_cpu" / "00000123"
    os.makedirs(d)
    write_bundle(str(d / "variables" / "variables"), {"a": np.array(0.5, np.float32), "b": np.array(2.0, np.float32)})
    write_saved_model(str(d / "saved_model.pb"),
                      [("x", "Placeholder", []), ("a", "VariableV2", []), ("a/read", "Identity", ["a"]), ("b", "VariableV2", []),
                       ("b/read", "Identity", ["b"]), ("Mul", "Mul", ["a/read", "x"]), ("y", "Add", ["Mul", "b/read"])],
                      ("x", "x:0", "y", "y:0"))
    out = tmp_path / "out" / "half_plus_two" / "123"
    man = sm.convert(str(d), str(out))
    assert man["template"] == "affine" and man["signature"] == {"input": "x", "output": "y"}
    oman, blob = models.load_bundle(str(out))
    assert models.forward(oman, blob, np.array([1.0, 2.0, 5.0], np.float32)).tolist() == [2.5, 3.0, 4.5]   # readme.md:40-42


def test_dense_mlp_saved_model_converts(tmp_path):
    rng = np.random.default_rng(1)
    w1, b1 = rng.standard_normal((6, 10)).astype(np.float32), rng.standard_normal(10).astype(np.float32)
    w2, b2 = rng.standard_normal((10, 4)).astype(np.float32), rng.standard_normal(4).astype(np.float32)
    d = tmp_path / "mlp" / "7"
    os.makedirs(d)
    write_bundle(str(d / "variables" / "variables"), {"dense/kernel": w1, "dense/bias": b1, "dense_1/kernel": w2, "dense_1/bias": b2})
    nodes = [("inputs", "Placeholder", []),
             ("dense/kernel", "VarHandleOp", []), ("dense/MatMul/ReadVariableOp", "ReadVariableOp", ["dense/kernel"]),
             ("dense/bias", "VarHandleOp", []), ("dense/BiasAdd/ReadVariableOp", "ReadVariableOp", ["dense/bias"]),
             ("dense/MatMul", "MatMul", ["inputs", "dense/MatMul/ReadVariableOp"]),
             ("dense/BiasAdd", "BiasAdd", ["dense/MatMul", "dense/BiasAdd/ReadVariableOp"]), ("dense/Relu", "Relu", ["dense/BiasAdd"]),
             ("dense_1/kernel", "VarHandleOp", []), ("dense_1/MatMul/ReadVariableOp", "ReadVariableOp", ["dense_1/kernel"]),
             ("dense_1/bias", "VarHandleOp", []), ("dense_1/BiasAdd/ReadVariableOp", "ReadVariableOp", ["dense_1/bias"]),
             ("dense_1/MatMul", "MatMul", ["dense/Relu", "dense_1/MatMul/ReadVariableOp"]),
             ("dense_1/BiasAdd", "BiasAdd", ["dense_1/MatMul", "dense_1/BiasAdd/ReadVariableOp"]),
             ("Identity", "Identity", ["dense_1/BiasAdd"])]
    write_saved_model(str(d / "saved_model.pb"), nodes, ("inputs", "inputs:0", "output_0", "Identity:0"))
    out = tmp_path / "out" / "mlp" / "7"
    man = sm.convert(str(d), str(out))
    assert [(l["in"], l["out"], l["activation"]) for l in man["layers"]] == [(6, 10, "relu"), (10, 4, "linear")]
    oman, blob = models.load_bundle(str(out))
    x = rng.standard_normal((3, 6)).astype(np.float32)
    ref = np.maximum(x.astype(np.float64) @ w1 + b1, 0) @ w2 + b2
    np.testing.assert_allclose(models.forward(oman, blob, x, np.float64), ref, rtol=1e-6, atol=1e-6)


def test_unsupported_graph_is_rejected(tmp_path):
    d = tmp_path / "weird" / "1"
    os.makedirs(d)
    write_bundle(str(d / "variables" / "variables"), {"a": np.array(1.0, np.float32)})
    write_saved_model(str(d / "saved_model.pb"), [("x", "Placeholder", []), ("y", "Softmax", ["x"])], ("x", "x:0", "y", "y:0"))
    with pytest.raises(ValueError):
        sm.convert(str(d), str(tmp_path / "o"))


def test_import_tree_converts_in_place(tmp_path):
    for ver in ("1", "00000002"):
        d = tmp_path / "hp2" / ver
        os.makedirs(d)
        write_bundle(str(d / "variables" / "variables"), {"a": np.array(0.5, np.float32), "b": np.array(float(ver), np.float32)})
        write_saved_model(str(d / "saved_model.pb"),
                          [("x", "Placeholder", []), ("a", "VariableV2", []), ("b", "VariableV2", []), ("Mul", "Mul", ["x", "a"]),
                           ("y", "AddV2", ["b", "Mul"])], ("x", "x:0", "y", "y:0"))
    os.makedirs(tmp_path / "bad" / "1")
    write_bundle(str(tmp_path / "bad" / "1" / "variables" / "variables"), {"a": np.array(1.0, np.float32)})
    write_saved_model(str(tmp_path / "bad" / "1" / "saved_model.pb"), [("x", "Placeholder", []), ("y", "Tanh", ["x"])], ("x", "x:0", "y", "y:0"))
    res = dict(sm.import_tree(str(tmp_path)))
    assert res[str(tmp_path / "hp2" / "1")] == "affine" and res[str(tmp_path / "hp2" / "00000002")] == "affine"
    assert res[str(tmp_path / "bad" / "1")].startswith("error")
    oman, blob = models.load_bundle(str(tmp_path / "hp2" / "00000002"))
    assert models.forward(oman, blob, np.array([2.0], np.float32)).tolist() == [3.0]
    assert sm.import_tree(str(tmp_path)) == [(str(tmp_path / "bad" / "1"), res[str(tmp_path / "bad" / "1")])]   # idempotent
