"""Writers for SavedModel fixtures in the public on-disk formats (protobuf + LevelDB-style table + tensor bundle),
used by tests/test_savedmodel.py and the ASan fuzz harness. Test infrastructure only."""
import os
import struct

import numpy as np

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


def _mlp_fixture(d, rng, dims=(6, 10, 4), relu_last=False, tf2_names=True):
    os.makedirs(d)
    tensors, nodes, prev = {}, [("inputs", "Placeholder", [])], "inputs"
    for i in range(len(dims) - 1):
        p = "dense" if i == 0 else f"dense_{i}"
        w, b = rng.standard_normal((dims[i], dims[i + 1])).astype(np.float32), rng.standard_normal(dims[i + 1]).astype(np.float32)
        tensors[p + "/kernel"], tensors[p + "/bias"] = w, b
        nodes += [(p + "/kernel", "VarHandleOp", []), (p + "/MatMul/ReadVariableOp", "ReadVariableOp", [p + "/kernel"]),
                  (p + "/bias", "VarHandleOp", []), (p + "/BiasAdd/ReadVariableOp", "ReadVariableOp", [p + "/bias"]),
                  (p + "/MatMul", "MatMul", [prev, p + "/MatMul/ReadVariableOp"]),
                  (p + "/BiasAdd", "BiasAdd", [p + "/MatMul", p + "/BiasAdd/ReadVariableOp"])]
        prev = p + "/BiasAdd"
        if i < len(dims) - 2 or relu_last:
            nodes.append((p + "/Relu", "Relu", [prev]))
            prev = p + "/Relu"
    nodes.append(("Identity", "Identity", [prev]))
    write_bundle(str(d / "variables" / "variables"), tensors)
    write_saved_model(str(d / "saved_model.pb"), nodes, ("inputs", "inputs:0", "output_0", "Identity:0"))
    return tensors
