"""SavedModel importer: converts a TensorFlow SavedModel directory
(`saved_model.pb` + `variables/variables.index` + `variables/variables.data-*`) of a recognised template into the
native "tfsc-b200-v1" bundle the library pages into HBM -- without TensorFlow.

Formats restated from their public definitions (field numbers as in the reference's generated protos):
  * saved_model.pb: SavedModel{meta_graphs=2{graph_def=2{node=1{name=1,op=2,input=3}}, signature_def=5 map}}
    (proto/tensorflow/core/protobuf/{saved_model,meta_graph}.pb.go, core/framework/{graph,node_def}.pb.go)
  * variables.index: a LevelDB-format table (data blocks with prefix-compressed keys + restart array, 5-byte block
    trailer, 48-byte footer, magic 0xdb4775248b80fb57) mapping "" -> BundleHeaderProto and tensor name ->
    BundleEntryProto{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}
    (proto/tensorflow/core/protobuf/tensor_bundle.pb.go:63-66,121-137)
  * variables.data-SSSSS-of-NNNNN: raw little-endian tensor bytes.
Recognised graphs: y = a*x + b with scalar variables (half_plus_two) and chains of MatMul + BiasAdd/Add (+ Relu)
(dense MLPs). STATUS: validated against bundles produced by this module's own writer (tests/test_savedmodel.py);
no TensorFlow-written SavedModel exists in the reference repo or this image to check against.
"""
from __future__ import annotations

import os
import struct

import numpy as np

from . import modelformat

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

_CRC32C_TABLE = []
for _n in range(256):
    _c = _n
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC32C_TABLE.append(_c)


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = _CRC32C_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def _varint(buf: bytes, pos: int):
    v = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _fields(buf: bytes):
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise ValueError("truncated protobuf")
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield f, wt, v


# ------------------------------------------------------------------------------ table / bundle ----
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        want = struct.unpack("<I", data[offset + size + 1:offset + size + 5])[0]
        if mask_crc(crc32c(data[offset:offset + size + 1])) != want:
            raise ValueError("table block checksum mismatch")
    if ctype != 0:
        raise ValueError("compressed table blocks (snappy) are not supported")
    return block


def _block_entries(block: bytes):
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path: str, verify_checksums: bool = True) -> dict:
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a table file (bad magic)")
    footer = data[-48:]
    _mi_off, p = _varint(footer, 0)
    _mi_size, p = _varint(footer, p)
    idx_off, p = _varint(footer, p)
    idx_size, p = _varint(footer, p)
    out = {}
    for _k, handle in _block_entries(_read_block(data, idx_off, idx_size, verify_checksums)):
        off, q = _varint(handle, 0)
        size, q = _varint(handle, q)
        for k, v in _block_entries(_read_block(data, off, size, verify_checksums)):
            out[bytes(k)] = bytes(v)
    return out


_NP_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def read_bundle(prefix: str, verify_checksums: bool = True) -> dict:
    """variables/variables -> {tensor name: ndarray}"""
    table = read_table(prefix + ".index", verify_checksums)
    num_shards = 1
    for f, wt, v in _fields(table.get(b"", b"")):
        if f == 1 and wt == 0:
            num_shards = v
        elif f == 2 and wt == 0 and v != 0:
            raise ValueError("big-endian tensor bundles are not supported")
    shards, out = {}, {}
    for name, entry in table.items():
        if name == b"":
            continue
        dtype = shard = offset = size = crc = 0
        shape = []
        for f, wt, v in _fields(entry):
            if f == 1 and wt == 0:
                dtype = v
            elif f == 2 and wt == 2:
                for f2, wt2, v2 in _fields(v):
                    if f2 == 2 and wt2 == 2:
                        dim = 0
                        for f3, wt3, v3 in _fields(v2):
                            if f3 == 1 and wt3 == 0:
                                dim = v3
                        shape.append(dim)
            elif f == 3 and wt == 0:
                shard = v
            elif f == 4 and wt == 0:
                offset = v
            elif f == 5 and wt == 0:
                size = v
            elif f == 6 and wt == 5:
                crc = struct.unpack("<I", v)[0]
        if dtype not in _NP_DTYPES:
            continue  # string / resource tensors are irrelevant to the supported templates
        if shard not in shards:
            shards[shard] = open(f"{prefix}.data-{shard:05d}-of-{num_shards:05d}", "rb").read()
        raw = shards[shard][offset:offset + size]
        if verify_checksums and crc and mask_crc(crc32c(raw)) != crc:
            raise ValueError(f"tensor {name!r}: data checksum mismatch")
        out[name.decode()] = np.frombuffer(raw, dtype=_NP_DTYPES[dtype]).reshape(shape).copy()
    return out


# --------------------------------------------------------------------------------- graph side ----
def parse_saved_model(path: str):
    """-> (nodes {name: (op, [inputs])}, signatures {name: (inputs {key: tensor}, outputs {key: tensor}, method)})"""
    buf = open(path, "rb").read()
    for f, wt, v in _fields(buf):
        if f == 2 and wt == 2:  # first MetaGraphDef
            nodes, sigs = {}, {}
            for f2, wt2, v2 in _fields(v):
                if f2 == 2 and wt2 == 2:  # GraphDef
                    for f3, wt3, v3 in _fields(v2):
                        if f3 == 1 and wt3 == 2:
                            name = op = ""
                            inputs = []
                            for f4, wt4, v4 in _fields(v3):
                                if f4 == 1 and wt4 == 2:
                                    name = v4.decode()
                                elif f4 == 2 and wt4 == 2:
                                    op = v4.decode()
                                elif f4 == 3 and wt4 == 2:
                                    inputs.append(v4.decode())
                            nodes[name] = (op, inputs)
                elif f2 == 5 and wt2 == 2:  # signature_def map entry
                    key, sig = "", None
                    for f3, wt3, v3 in _fields(v2):
                        if f3 == 1 and wt3 == 2:
                            key = v3.decode()
                        elif f3 == 2 and wt3 == 2:
                            sig = v3
                    ins, outs, method = {}, {}, ""
                    for f3, wt3, v3 in _fields(sig or b""):
                        if f3 in (1, 2) and wt3 == 2:
                            k2, tname = "", ""
                            for f4, wt4, v4 in _fields(v3):
                                if f4 == 1 and wt4 == 2:
                                    k2 = v4.decode()
                                elif f4 == 2 and wt4 == 2:
                                    for f5, wt5, v5 in _fields(v4):
                                        if f5 == 1 and wt5 == 2:
                                            tname = v5.decode()
                            (ins if f3 == 1 else outs)[k2] = tname
                        elif f3 == 3 and wt3 == 2:
                            method = v3.decode()
                    sigs[key] = (ins, outs, method)
            return nodes, sigs
    raise ValueError("saved_model.pb holds no MetaGraphDef")


def _node(tensor: str) -> str:
    t = tensor.lstrip("^")
    return t.split(":")[0]


def _resolve_variable(nodes, name, variables):
    """Follow Identity / ReadVariableOp chains down to a variable node that exists in the bundle."""
    seen = 0
    while seen < 16:
        op, inputs = nodes[name]
        if op in ("VariableV2", "Variable", "VarHandleOp") and name in variables:
            return name
        if op in ("Identity", "ReadVariableOp") and inputs:
            name = _node(inputs[0])
            seen += 1
            continue
        return None
    return None


def convert(saved_model_dir: str, out_version_dir: str, signature: str = "serving_default") -> dict:
    """SavedModel directory -> tfsc-b200-v1 bundle. Returns the manifest. Raises ValueError for graphs outside the
    recognised templates."""
    nodes, sigs = parse_saved_model(os.path.join(saved_model_dir, "saved_model.pb"))
    variables = read_bundle(os.path.join(saved_model_dir, "variables", "variables"))
    if signature not in sigs:
        predict = [k for k, (_i, _o, m) in sigs.items() if m.endswith("predict")]
        if not predict:
            raise ValueError(f"signature {signature!r} not found and no predict signature present")
        signature = sorted(predict)[0]
    ins, outs, _method = sigs[signature]
    if len(ins) != 1 or len(outs) != 1:
        raise ValueError("only single-input single-output predict signatures are supported")
    (in_key, in_t), (out_key, out_t) = next(iter(ins.items())), next(iter(outs.items()))
    x = _node(in_t)
    # classify / regress signatures are kept; their tf.Example feature is the predict input key (half_plus_two: "x")
    extra = [{"name": k, "method": "classify" if m.endswith("classify") else "regress", "feature": in_key}
             for k, (_i, _o, m) in sorted(sigs.items()) if m.endswith("classify") or m.endswith("regress")]

    def is_x(t):
        n = _node(t)
        while nodes.get(n, ("", []))[0] == "Identity":
            n = _node(nodes[n][1][0])
        return n == x

    cur = _node(out_t)
    while nodes[cur][0] == "Identity":
        cur = _node(nodes[cur][1][0])
    # ---- affine: y = Add(Mul(a, x), b)
    op, inputs = nodes[cur]
    if op in ("Add", "AddV2") and len(inputs) == 2:
        for mul_t, b_t in (inputs, inputs[::-1]):
            mul = _node(mul_t)
            if nodes.get(mul, ("", []))[0] == "Mul":
                m_in = nodes[mul][1]
                for a_t, x_t in (m_in, m_in[::-1]):
                    a_var, b_var = _resolve_variable(nodes, _node(a_t), variables), _resolve_variable(nodes, _node(b_t), variables)
                    if a_var and b_var and is_x(x_t) and variables[a_var].size == 1 and variables[b_var].size == 1:
                        return modelformat.write_affine_bundle(out_version_dir, float(variables[a_var].ravel()[0]),
                                                               float(variables[b_var].ravel()[0]), in_key, out_key, extra)
    # ---- MLP: walk back from the output through [Relu] <- BiasAdd/Add <- MatMul
    layers = []
    while not is_x(cur):
        op, inputs = nodes[cur]
        relu = False
        if op == "Relu":
            relu, cur = True, _node(inputs[0])
            op, inputs = nodes[cur]
        if op not in ("BiasAdd", "Add", "AddV2") or len(inputs) != 2:
            raise ValueError(f"unsupported op {op!r} at node {cur!r}: not an affine or dense-MLP graph")
        mm, bias_var = None, None
        for a_t, b_t in (inputs, inputs[::-1]):
            if nodes.get(_node(a_t), ("", []))[0] == "MatMul":
                mm, bias_var = _node(a_t), _resolve_variable(nodes, _node(b_t), variables)
        if mm is None or bias_var is None:
            raise ValueError(f"node {cur!r}: expected MatMul + bias variable")
        m_in = nodes[mm][1]
        w_var = _resolve_variable(nodes, _node(m_in[1]), variables)
        if w_var is None or variables[w_var].ndim != 2:
            raise ValueError(f"node {mm!r}: MatMul weight is not a 2-D variable")
        layers.append((variables[w_var], variables[bias_var], "relu" if relu else "linear"))
        cur = _node(m_in[0])
        while nodes[cur][0] == "Identity":
            cur = _node(nodes[cur][1][0])
    if not layers:
        raise ValueError("graph is neither y = a*x + b nor a dense MLP")
    layers.reverse()
    return modelformat.write_mlp_bundle(out_version_dir, [l[0] for l in layers], [l[1] for l in layers], [l[2] for l in layers],
                                        in_key, out_key, extra)


def import_tree(base_dir: str, force: bool = False) -> list:
    """Walk <base_dir>/<model>/<version>/ (the disk provider's layout, diskmodelprovider.go:44-76) and write the native
    bundle next to every saved_model.pb that lacks one. Returns [(version_dir, template | error string)]."""
    done = []
    for model in sorted(os.listdir(base_dir)):
        mdir = os.path.join(base_dir, model)
        if not os.path.isdir(mdir):
            continue
        for ver in sorted(os.listdir(mdir)):
            vdir = os.path.join(mdir, ver)
            if not os.path.isfile(os.path.join(vdir, "saved_model.pb")):
                continue
            if os.path.isfile(os.path.join(vdir, "tfsc_model.json")) and not force:
                continue
            try:
                done.append((vdir, convert(vdir, vdir)["template"]))
            except (ValueError, KeyError, OSError) as e:
                done.append((vdir, f"error: {e}"))
    return done


if __name__ == "__main__":
    import sys
    if len(sys.argv) < 2:
        sys.exit("usage: python -m tfservingcache_b200.savedmodel <modelProvider.diskProvider.baseDir> [--force]")
    failed = 0
    for vdir, result in import_tree(sys.argv[1], "--force" in sys.argv[2:]):
        print(f"{vdir}: {result}")
        failed += result.startswith("error")
    sys.exit(1 if failed else 0)
