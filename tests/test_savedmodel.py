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


from savedmodel_fixtures import _mlp_fixture, write_bundle, write_saved_model  # noqa: E402


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


# ---- native importer (csrc/savedmodel.cc through the C ABI): same fixtures, byte-identical bundles -------------------
def _native_convert(src, dst):
    from tfservingcache_b200 import _lib
    os.makedirs(dst, exist_ok=True)
    return _lib.lib.tfsc_savedmodel_convert(str(src).encode(), str(dst).encode())


def test_native_crc32c_matches_python():
    from tfservingcache_b200 import _lib
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 8, 9, 63, 64, 1000, 4097):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert _lib.lib.tfsc_crc32c(data, len(data)) == sm.crc32c(data)
    assert _lib.lib.tfsc_crc32c(b"123456789", 9) == 0xE3069283


def test_native_importer_matches_python_importer(tmp_path):
    rng = np.random.default_rng(11)
    cases = []
    d = tmp_path / "hp2" / "00000123"
    os.makedirs(d)
    write_bundle(str(d / "variables" / "variables"), {"a": np.array(0.5, np.float32), "b": np.array(2.0, np.float32)})
    write_saved_model(str(d / "saved_model.pb"),
                      [("x", "Placeholder", []), ("a", "VariableV2", []), ("a/read", "Identity", ["a"]), ("b", "VariableV2", []),
                       ("b/read", "Identity", ["b"]), ("Mul", "Mul", ["a/read", "x"]), ("y", "Add", ["Mul", "b/read"])],
                      ("x", "x:0", "y", "y:0"))
    cases.append(d)
    for i, (dims, relu_last) in enumerate([((6, 10, 4), False), ((5, 7), True), ((33, 64, 17, 9, 3), False), ((300, 260), False)]):
        d = tmp_path / f"mlp{i}" / "1"
        _mlp_fixture(d, rng, dims, relu_last)
        cases.append(d)
    for d in cases:
        py, nat = tmp_path / "py" / d.parent.name, tmp_path / "nat" / d.parent.name
        sm.convert(str(d), str(py))
        assert _native_convert(d, nat) == 0
        assert open(py / "weights.bin", "rb").read() == open(nat / "weights.bin", "rb").read()
        assert json.load(open(py / "tfsc_model.json")) == json.load(open(nat / "tfsc_model.json"))


def test_native_importer_rejects_bad_inputs(tmp_path):
    from tfservingcache_b200 import _lib
    rng = np.random.default_rng(5)
    assert _native_convert(tmp_path / "missing", tmp_path / "o0") == _lib.E_NOT_FOUND
    d = tmp_path / "weird" / "1"
    os.makedirs(d)
    write_bundle(str(d / "variables" / "variables"), {"a": np.array(1.0, np.float32)})
    write_saved_model(str(d / "saved_model.pb"), [("x", "Placeholder", []), ("y", "Softmax", ["x"])], ("x", "x:0", "y", "y:0"))
    assert _native_convert(d, tmp_path / "o1") == _lib.E_INVALID and b"Softmax" in _lib.lib.tfsc_last_error()
    # corrupted tensor data, corrupted index block, truncated files: an error, never a crash
    d = tmp_path / "m" / "1"
    _mlp_fixture(d, rng)
    p = d / "variables" / "variables.data-00000-of-00001"
    raw = bytearray(open(p, "rb").read())
    raw[5] ^= 0x40
    open(p, "wb").write(bytes(raw))
    assert _native_convert(d, tmp_path / "o2") == _lib.E_INVALID and b"checksum" in _lib.lib.tfsc_last_error()
    raw[5] ^= 0x40
    open(p, "wb").write(bytes(raw))
    assert _native_convert(d, tmp_path / "o3") == 0
    idx = d / "variables" / "variables.index"
    good = open(idx, "rb").read()
    for cut in (0, 10, 47, len(good) - 1):
        open(idx, "wb").write(good[:cut])
        assert _native_convert(d, tmp_path / "o4") == _lib.E_INVALID
    for pos in range(0, len(good), 7):
        bad = bytearray(good)
        bad[pos] ^= 0xFF
        open(idx, "wb").write(bytes(bad))
        assert _native_convert(d, tmp_path / "o5") in (0, _lib.E_INVALID)
    open(idx, "wb").write(good)
    pb = d / "saved_model.pb"
    gpb = open(pb, "rb").read()
    for cut in range(0, len(gpb), 11):
        open(pb, "wb").write(gpb[:cut])
        assert _native_convert(d, tmp_path / "o6") in (0, _lib.E_INVALID)


# ---- f1 pin: fixtures assembled by an independent writer from the reference's own protobuf schema ------------------
def _write_golden_model(golden, name, version_dir):
    import base64
    g = golden("savedmodel_golden.json")["models"][name]
    for rel, b64 in g["files"].items():
        p = os.path.join(version_dir, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "wb") as f:
            f.write(base64.b64decode(b64))
    return g["expect"]


@pytest.mark.parametrize("name", ["half_plus_two", "keras_mlp"])
def test_golden_savedmodel_written_with_reference_schema_converts(name, golden, tmp_path):
    """tests/golden/savedmodel_golden.json was produced by python-protobuf from the FileDescriptorProtos embedded in the
    reference's generated code plus an independent LevelDB-table writer (tests/golden/make_savedmodel_golden.py) -- no code
    shared with the product. Both importers (native C++ through the C ABI, and the Python twin) must read it."""
    import base64
    import ctypes
    import tfservingcache_b200 as t
    src = tmp_path / "src" / name / "1"
    expect = _write_golden_model(golden, name, str(src))
    outs = {}
    for impl in ("native", "python"):
        out = tmp_path / impl / name / "1"
        if impl == "native":
            os.makedirs(out)
            rc = t._lib.lib.tfsc_savedmodel_convert(str(src).encode(), str(out).encode())
            assert rc >= 0, t._lib.lib.tfsc_last_error()
            man = json.load(open(out / "tfsc_model.json"))
        else:
            man = sm.convert(str(src), str(out))
        assert man["template"] == expect["template"] and man["signature"] == {"input": expect["input"], "output": expect["output"]}
        oman, blob = models.load_bundle(str(out))
        x = np.array(expect["x"], np.float32)
        y = models.forward(oman, blob, x, np.float64)
        assert np.max(np.abs(y - np.array(expect["y"]))) <= 1e-5
        outs[impl] = (open(out / "weights.bin", "rb").read(), json.load(open(out / "tfsc_model.json")))
    assert outs["native"][0] == outs["python"][0]          # byte-identical bundles from both importers
    if name == "keras_mlp":
        got = sm.read_bundle(str(src / "variables" / "variables"))
        for k, b64 in expect["tensors"].items():
            assert got[k].astype("<f4").tobytes() == base64.b64decode(b64)
    # the independent writer's checksum (bitwise CRC-32C) and the product's table-driven one agree on the data shard
    data = open(src / "variables" / "variables.data-00000-of-00001", "rb").read()
    assert ctypes.c_uint32(t._lib.lib.tfsc_crc32c(data, len(data))).value == sm.crc32c(data)


@pytest.mark.gpu
def test_golden_half_plus_two_savedmodel_is_served(golden, tmp_path):
    """BASELINE configs[0]: the half_plus_two SavedModel through diskProvider and REST, on the GPU: [1,2,5] -> [2.5,3,4.5]
    (deploy/docker-compose/readme.md:40-42), from the independently written fixture."""
    import torch
    import tfservingcache_b200 as t
    assert torch.cuda.is_available()
    _write_golden_model(golden, "half_plus_two", str(tmp_path / "half_plus_two" / "00000123"))
    expect = _write_golden_model(golden, "keras_mlp", str(tmp_path / "keras_mlp" / "1"))
    cfg = {"modelProvider.type": "diskProvider", "modelProvider.diskProvider.baseDir": str(tmp_path), "gpu.devices": [0],
           "gpu.arenaBytes": 8 << 20, "modelCache.size": 1 << 26, "serving.maxConcurrentModels": 2}
    with t.Server(cfg) as srv:
        st, body = srv.rest_handle("POST", "/v1/models/half_plus_two/versions/123:predict", b'{"instances": [1.0, 2.0, 5.0]}')
        assert st == 200 and json.loads(body) == {"predictions": [2.5, 3.0, 4.5]}
        y = srv.predict("keras_mlp", "1", np.array(expect["x"], np.float32), input_name="inputs")
        assert np.max(np.abs(y - np.array(expect["y"]))) <= 1e-4
