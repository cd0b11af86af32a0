"""Writer for the native model bundle ("tfsc-b200-v1"): <baseDir>/<name>/<version>/
{tfsc_model.json, weights.bin}.  weights.bin is what gets paged verbatim into the HBM arena: per
dense layer W[in,out] row-major fp32 (the TF dense-kernel layout) then b[out], every tensor
256-byte aligned."""
from __future__ import annotations

import json
import os

import numpy as np


def _align256(x: int) -> int:
    return (x + 255) & ~255


def write_mlp_bundle(version_dir: str, weights, biases, activations=None, input_name="x", output_name="y", extra_signatures=None):
    n = len(weights)
    if activations is None:
        activations = ["relu"] * (n - 1) + ["linear"]
    off, layers = 0, []
    for w, b, act in zip(weights, biases, activations):
        fi, fo = w.shape
        w_off = off
        off = _align256(off + fi * fo * 4)
        b_off = off
        off = _align256(off + fo * 4)
        layers.append({"in": int(fi), "out": int(fo), "activation": act, "w_offset": w_off, "b_offset": b_off})
    blob = np.zeros(off // 4, dtype=np.float32)
    for L, w, b in zip(layers, weights, biases):
        blob[L["w_offset"] // 4: L["w_offset"] // 4 + w.size] = np.asarray(w, np.float32).ravel()
        blob[L["b_offset"] // 4: L["b_offset"] // 4 + b.size] = np.asarray(b, np.float32).ravel()
    man = {"format": "tfsc-b200-v1", "template": "mlp", "dtype": "float32",
           "signature": {"input": input_name, "output": output_name}, "layers": layers, "weights_bytes": off}
    if extra_signatures:   # classify / regress signatures: [{"name", "method": "classify"|"regress", "feature"}]
        man["extra_signatures"] = list(extra_signatures)
    _write(version_dir, man, blob)
    return man


def write_affine_bundle(version_dir: str, a: float, b: float, input_name="x", output_name="y", extra_signatures=None):
    blob = np.zeros(128, dtype=np.float32)
    blob[0], blob[64] = a, b
    man = {"format": "tfsc-b200-v1", "template": "affine", "dtype": "float32",
           "signature": {"input": input_name, "output": output_name}, "a_offset": 0, "b_offset": 256,
           "weights_bytes": 512}
    if extra_signatures:
        man["extra_signatures"] = list(extra_signatures)
    _write(version_dir, man, blob)
    return man


def _write(version_dir: str, man: dict, blob: np.ndarray):
    os.makedirs(version_dir, exist_ok=True)
    with open(os.path.join(version_dir, "tfsc_model.json"), "w") as f:
        json.dump(man, f)
    blob.astype("<f4").tofile(os.path.join(version_dir, "weights.bin"))


def _graph_manifest(input_shape, ops, n_buffers, input_name="x", output_name="y", input_dtype="float32"):
    off = 0

    def take(nbytes):
        nonlocal off
        o = off
        off = _align256(off + nbytes)
        return o

    for op in ops:
        if op["op"] in ("conv", "dense"):
            k = op.get("kh", 1) * op.get("kw", 1) * op["c"]
            op["w_offset"] = take(k * op["cout"] * 4)
            op["b_offset"] = take(op["cout"] * 4)
        elif op["op"] in ("layernorm", "embed"):
            op["w_offset"] = take(op["c"] * 4)      # gamma
            op["b_offset"] = take(op["c"] * 4)      # beta
            if op["op"] == "embed":
                op["word_offset"] = take(op["vocab"] * op["c"] * 4)
                op["pos_offset"] = take(op["max_pos"] * op["c"] * 4)
                op["type_offset"] = take(2 * op["c"] * 4)
    return {"format": "tfsc-b200-v1", "template": "graph", "dtype": "float32", "input_dtype": input_dtype,
            "signature": {"input": input_name, "output": output_name}, "input_shape": list(input_shape),
            "n_buffers": n_buffers, "ops": ops, "weights_bytes": off}


def resnet50_manifest(image=224, classes=1000, width=64, blocks=(3, 4, 6, 3)):
    """ResNet-50 v1.5 (torchvision topology: stride on the 3x3 conv) as a graph bundle, NHWC, BatchNorm folded into
    kernel + bias. Buffers: 0 = block input / identity, 1 = 1x1 out, 2 = 3x3 out, 3 = block out, 4 = downsample."""
    ops, h = [], image
    ops.append({"op": "conv", "src": -1, "dst": 0, "h": h, "w": h, "c": 3, "kh": 7, "kw": 7, "stride": 2, "pad": 3,
                "cout": width, "act": "relu"})
    h = (h + 6 - 7) // 2 + 1
    ops.append({"op": "maxpool", "src": 0, "dst": 1, "h": h, "w": h, "c": width, "kh": 3, "kw": 3, "stride": 2, "pad": 1})
    h = (h + 2 - 3) // 2 + 1
    cur, cin = 1, width          # `cur` = buffer holding the block input
    free = [0, 2, 3, 4]
    for li, nb in enumerate(blocks):
        planes = width * (2 ** li)
        for bi in range(nb):
            stride = 2 if (bi == 0 and li > 0) else 1
            a, b, c_, d = [x for x in range(5) if x != cur][:4]
            ho = (h + 2 - 3) // stride + 1
            ops.append({"op": "conv", "src": cur, "dst": a, "h": h, "w": h, "c": cin, "kh": 1, "kw": 1, "stride": 1, "pad": 0,
                        "cout": planes, "act": "relu"})
            ops.append({"op": "conv", "src": a, "dst": b, "h": h, "w": h, "c": planes, "kh": 3, "kw": 3, "stride": stride,
                        "pad": 1, "cout": planes, "act": "relu"})
            ident = cur
            if bi == 0:  # projection shortcut
                ops.append({"op": "conv", "src": cur, "dst": c_, "h": h, "w": h, "c": cin, "kh": 1, "kw": 1, "stride": stride,
                            "pad": 0, "cout": planes * 4, "act": "none"})
                ident = c_
            ops.append({"op": "conv", "src": b, "dst": d, "res": ident, "h": ho, "w": ho, "c": planes, "kh": 1, "kw": 1,
                        "stride": 1, "pad": 0, "cout": planes * 4, "act": "relu"})
            cur, cin, h = d, planes * 4, ho
    a = [x for x in range(5) if x != cur][0]
    ops.append({"op": "avgpool", "src": cur, "dst": a, "h": h, "w": h, "c": cin})
    ops.append({"op": "dense", "src": a, "dst": -2, "h": 1, "w": 1, "c": cin, "cout": classes, "act": "none"})
    return _graph_manifest([image, image, 3], ops, 5)


def write_graph_bundle(version_dir: str, manifest: dict, blob: np.ndarray):
    _write(version_dir, manifest, np.asarray(blob, np.float32))


def bert_manifest(seq=128, hidden=768, layers=12, heads=12, inter=3072, vocab=30522, max_pos=512, labels=2):
    """BERT-base fine-tune variant (Devlin et al. 2018) as a graph bundle: token ids int32 [B, seq] -> logits
    [B, labels]. A sequence is an "image" with h = seq tokens, w = 1, c = width; dense layers are 1x1 convs; the
    attention mask is derived from the ids ([PAD] = 0), token_type is 0.
    Buffers: 0 hidden, 1 qkv / ffn-intermediate, 2 context / post-attention, 3 dense output."""
    ops = [{"op": "embed", "src": -1, "dst": 0, "h": seq, "w": 1, "c": hidden, "vocab": vocab, "max_pos": max_pos, "eps": 1e-12}]

    def dense(src, dst, cin, cout, act="none", res=None):
        o = {"op": "conv", "src": src, "dst": dst, "h": seq, "w": 1, "c": cin, "kh": 1, "kw": 1, "stride": 1, "pad": 0,
             "cout": cout, "act": act}
        if res is not None:
            o["res"] = res
        return o

    for _ in range(layers):
        ops.append(dense(0, 1, hidden, 3 * hidden))                                     # fused Q|K|V projection
        ops.append({"op": "attention", "src": 1, "dst": 2, "h": seq, "w": 1, "c": 3 * hidden, "heads": heads})
        ops.append(dense(2, 3, hidden, hidden))                                         # attention output projection
        ops.append({"op": "layernorm", "src": 3, "res": 0, "dst": 2, "h": seq, "w": 1, "c": hidden, "eps": 1e-12})
        ops.append(dense(2, 1, hidden, inter, act="gelu"))
        ops.append(dense(1, 3, inter, hidden))
        ops.append({"op": "layernorm", "src": 3, "res": 2, "dst": 0, "h": seq, "w": 1, "c": hidden, "eps": 1e-12})
    ops.append({"op": "dense", "src": 0, "dst": 1, "h": 1, "w": 1, "c": hidden, "cout": hidden, "act": "tanh"})   # pooler on [CLS]
    ops.append({"op": "dense", "src": 1, "dst": -2, "h": 1, "w": 1, "c": hidden, "cout": labels, "act": "none"})
    return _graph_manifest([seq], ops, 4, input_name="input_ids", output_name="logits", input_dtype="int32")
