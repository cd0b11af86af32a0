"""Generates tests/golden/savedmodel_golden.json: two TensorFlow SavedModel directories assembled by an INDEPENDENT writer
(f1 pin). Nothing here imports tfservingcache_b200 or tests/savedmodel_fixtures.py:
  * every protobuf message (SavedModel / MetaGraphDef / GraphDef / NodeDef / SignatureDef / TensorInfo, BundleHeaderProto /
    BundleEntryProto / TensorShapeProto) is built and serialized by python-protobuf from the FileDescriptorProtos EMBEDDED IN
    THE REFERENCE'S OWN generated code (proto/tensorflow/core/protobuf/{saved_model,meta_graph,tensor_bundle}.pb.go,
    core/framework/{graph,node_def,tensor_shape,types}.pb.go) -- so field numbers and encodings are the reference's;
  * the variables.index container is the LevelDB table format, written here from its public description (doc/table_format.md):
    prefix-compressed entries, restart array, 1-byte type + masked CRC-32C trailer per block, metaindex + index blocks,
    48-byte footer with magic 0xdb4775248b80fb57; CRC-32C is computed bit by bit (no table shared with the product).
The half_plus_two graph mirrors TF-Serving's test model (a = 0.5, b = 2, y = a*x + b, with the classify / regress
signatures it also exports). Run in the BUILD container (needs /root/reference):  python tests/golden/make_savedmodel_golden.py
"""
import base64
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (only for build_pool(): the descriptor loader)


# ------------------------------------------------------------------ CRC-32C, bitwise (Castagnoli, reflected 0x82F63B78)
def crc32c_bitwise(data: bytes) -> int:
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def leveldb_mask(crc: int) -> int:   # util/crc32c.h: rotate right by 15 bits and add a constant
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def uvarint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


# ------------------------------------------------------------------ LevelDB table (table/format.cc, table_builder.cc)
def build_block(entries, restart_interval=16):
    buf, restarts, prev = bytearray(), [], b""
    for n, (key, value) in enumerate(entries):
        shared = 0
        if n % restart_interval == 0:
            restarts.append(len(buf))
        else:
            lim = min(len(prev), len(key))
            while shared < lim and prev[shared] == key[shared]:
                shared += 1
        buf += uvarint(shared) + uvarint(len(key) - shared) + uvarint(len(value)) + key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def emit_block(out: bytearray, block: bytes):
    """returns the BlockHandle (offset, size) and appends block + trailer (type 0 = no compression, masked crc32c)"""
    handle = (len(out), len(block))
    trailer_type = b"\x00"
    out += block + trailer_type + struct.pack("<I", leveldb_mask(crc32c_bitwise(block + trailer_type)))
    return handle


def build_table(items: dict, entries_per_block=2) -> bytes:
    keys = sorted(items)
    out, index_entries = bytearray(), []
    for i in range(0, len(keys), entries_per_block):
        chunk = keys[i:i + entries_per_block]
        off, size = emit_block(out, build_block([(k, items[k]) for k in chunk]))
        index_entries.append((chunk[-1], uvarint(off) + uvarint(size)))     # separator = last key of the block
    meta_off, meta_size = emit_block(out, build_block([]))
    idx_off, idx_size = emit_block(out, build_block(index_entries, restart_interval=1))
    footer = uvarint(meta_off) + uvarint(meta_size) + uvarint(idx_off) + uvarint(idx_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    return bytes(out) + footer


# ------------------------------------------------------------------ messages from the reference's descriptors
def messages():
    mg.build_pool()
    get = mg.build_pool.get
    try:
        get("tensorflow.SavedModel")
    except Exception:
        # saved_model.proto is not a dependency of the serving APIs build_pool() loads: add it (and its deps) explicitly
        from google.protobuf import descriptor_pb2, descriptor_pool
        pool = descriptor_pool.Default()
        fds = {}
        for sub in ("core/framework", "core/lib/core", "core/protobuf", "core/example"):
            d = os.path.join(mg.REF, sub)
            for fn in sorted(os.listdir(d)):
                if fn.endswith(".pb.go"):
                    for blob in mg.embedded_descriptors(os.path.join(d, fn)):
                        fd = descriptor_pb2.FileDescriptorProto.FromString(blob)
                        fds[fd.name] = fd
        done = set()

        def add(name):
            if name in done or name not in fds:
                return
            done.add(name)
            for dep in fds[name].dependency:
                add(dep)
            try:
                pool.Add(fds[name])
            except Exception:
                pass
        for want in ("tensorflow/core/protobuf/saved_model.proto", "tensorflow/core/protobuf/tensor_bundle.proto"):
            add(want)
    return {n: get("tensorflow." + n) for n in ("SavedModel", "BundleHeaderProto", "BundleEntryProto")}


def bundle_files(M, tensors: dict):
    """variables.index + variables.data-00000-of-00001 for {name: float32 ndarray}"""
    data = bytearray()
    header = M["BundleHeaderProto"]()
    header.num_shards = 1
    header.endianness = 0       # LITTLE
    header.version.producer = 1
    items = {b"": header.SerializeToString(deterministic=True)}
    for name in sorted(tensors):
        arr = np.ascontiguousarray(tensors[name], dtype="<f4")
        raw = arr.tobytes()
        e = M["BundleEntryProto"]()
        e.dtype = 1             # DT_FLOAT
        for d in arr.shape:
            e.shape.dim.add().size = int(d)
        e.shard_id = 0
        e.offset = len(data)
        e.size = len(raw)
        e.crc32c = leveldb_mask(crc32c_bitwise(raw))
        items[name.encode()] = e.SerializeToString(deterministic=True)
        data += raw
    return build_table(items), bytes(data)


def saved_model_bytes(M, nodes, signatures):
    """nodes: [(name, op, [inputs])]; signatures: {key: (inputs {k: tensor}, outputs {k: tensor}, method)}"""
    sm = M["SavedModel"]()
    sm.saved_model_schema_version = 1
    mgd = sm.meta_graphs.add()
    mgd.meta_info_def.tags.append("serve")
    for name, op, inputs in nodes:
        n = mgd.graph_def.node.add()
        n.name, n.op = name, op
        n.input.extend(inputs)
        if op in ("Placeholder", "VariableV2", "VarHandleOp"):
            n.attr["dtype"].type = 1
    for key, (ins, outs, method) in signatures.items():
        sd = mgd.signature_def[key]
        for k, tname in ins.items():
            sd.inputs[k].name = tname
            sd.inputs[k].dtype = 1
        for k, tname in outs.items():
            sd.outputs[k].name = tname
            sd.outputs[k].dtype = 1
        sd.method_name = method
    return sm.SerializeToString(deterministic=True)


def main():
    M = messages()
    out = {"generator": "tests/golden/make_savedmodel_golden.py (python-protobuf + the reference's embedded descriptors; own LevelDB table writer)",
           "models": {}}
    # ---- half_plus_two as TF-Serving's testdata exports it (TF1 graph: VariableV2 + Identity reads)
    nodes = [("a", "VariableV2", []), ("a/read", "Identity", ["a"]), ("b", "VariableV2", []), ("b/read", "Identity", ["b"]),
             ("x", "Placeholder", []), ("Mul", "Mul", ["a/read", "x"]), ("y", "Add", ["Mul", "b/read"])]
    sigs = {"serving_default": ({"x": "x:0"}, {"y": "y:0"}, "tensorflow/serving/predict"),
            "regress_x_to_y": ({"inputs": "tf_example:0"}, {"outputs": "y:0"}, "tensorflow/serving/regress"),
            "classify_x_to_y": ({"inputs": "tf_example:0"}, {"scores": "y:0"}, "tensorflow/serving/classify")}
    idx, dat = bundle_files(M, {"a": np.float32(0.5).reshape(()), "b": np.float32(2.0).reshape(())})
    out["models"]["half_plus_two"] = {"files": {"saved_model.pb": saved_model_bytes(M, nodes, sigs), "variables/variables.index": idx,
                                                "variables/variables.data-00000-of-00001": dat},
                                      "expect": {"template": "affine", "a": 0.5, "b": 2.0, "input": "x", "output": "y",
                                                 "x": [1.0, 2.0, 5.0], "y": [2.5, 3.0, 4.5]}}
    # ---- a Keras-style dense MLP (TF2 resource variables), 5 tensors spread over several table blocks
    rng = np.random.default_rng(11)
    dims = [12, 20, 16, 8]
    tensors, nodes, prev = {}, [("serving_default_inputs", "Placeholder", [])], "serving_default_inputs"
    for i in range(len(dims) - 1):
        p = "dense" if i == 0 else f"dense_{i}"
        tensors[p + "/kernel"] = rng.standard_normal((dims[i], dims[i + 1])).astype(np.float32)
        tensors[p + "/bias"] = rng.standard_normal(dims[i + 1]).astype(np.float32)
        nodes += [(p + "/kernel", "VarHandleOp", []), (p + "/MatMul/ReadVariableOp", "ReadVariableOp", [p + "/kernel"]),
                  (p + "/bias", "VarHandleOp", []), (p + "/BiasAdd/ReadVariableOp", "ReadVariableOp", [p + "/bias"]),
                  (p + "/MatMul", "MatMul", [prev, p + "/MatMul/ReadVariableOp"]),
                  (p + "/BiasAdd", "BiasAdd", [p + "/MatMul", p + "/BiasAdd/ReadVariableOp"])]
        prev = p + "/BiasAdd"
        if i < len(dims) - 2:
            nodes.append((p + "/Relu", "Relu", [prev]))
            prev = p + "/Relu"
    nodes.append(("StatefulPartitionedCall", "Identity", [prev]))
    idx, dat = bundle_files(M, tensors)
    x = rng.standard_normal((3, dims[0])).astype(np.float32)
    h = x.astype(np.float64)
    for i in range(len(dims) - 1):
        p = "dense" if i == 0 else f"dense_{i}"
        h = h @ tensors[p + "/kernel"].astype(np.float64) + tensors[p + "/bias"].astype(np.float64)
        if i < len(dims) - 2:
            h = np.maximum(h, 0)
    out["models"]["keras_mlp"] = {"files": {"saved_model.pb": saved_model_bytes(M, nodes, {"serving_default": (
        {"inputs": "serving_default_inputs:0"}, {"output_0": "StatefulPartitionedCall:0"}, "tensorflow/serving/predict")}),
        "variables/variables.index": idx, "variables/variables.data-00000-of-00001": dat},
        "expect": {"template": "mlp", "dims": dims, "input": "inputs", "output": "output_0", "x": x.tolist(), "y": h.tolist(),
                   "tensors": {k: base64.b64encode(v.astype("<f4").tobytes()).decode() for k, v in tensors.items()}}}
    for m in out["models"].values():
        m["files"] = {k: base64.b64encode(v).decode() for k, v in m["files"].items()}
    with open(os.path.join(HERE, "savedmodel_golden.json"), "w") as f:
        json.dump(out, f)
    print({k: {fn: len(base64.b64decode(b)) for fn, b in m["files"].items()} for k, m in out["models"].items()})


if __name__ == "__main__":
    main()
