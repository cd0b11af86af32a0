"""Numeric oracle for the executor the reference delegates to (TF-Serving, external and
unpinned: deploy/docker-compose/docker-compose.yaml:22-37) -- PARITY UNPINNED except for the
half_plus_two known answer (deploy/docker-compose/readme.md:40-42).  numpy restatement of the
forward pass of each model template the B200 build executes, reading the same ``weights.bin``
blob + ``tfsc_model.json`` manifest the product pages into HBM (independent parse).

Also holds the seeded synthetic weight generator (integer hash -> uniform fp32), restated
bit-exactly by the product's synthetic provider (csrc/provider.cc) so that 1 GB models never
have to be stored: value(seed, tensor, i) = (u24(mix32(i + k)) * 2^-24 * 2 - 1) * scale.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

M32 = 0xFFFFFFFF


def mix32(x):
    """lowbias32 integer hash on uint32 (numpy array or int)."""
    if isinstance(x, (int, np.integer)):
        x = int(x) & M32
        x ^= x >> 16
        x = (x * 0x7FEB352D) & M32
        x ^= x >> 15
        x = (x * 0x846CA68B) & M32
        x ^= x >> 16
        return x
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(M32)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(M32)
    x ^= x >> np.uint64(16)
    return x


def tensor_key(seed: int, tensor_id: int) -> int:
    return mix32((seed * 0x9E3779B9 + tensor_id * 0x85EBCA6B + 0x165667B1) & M32)


def synth_tensor(seed: int, tensor_id: int, n: int, scale: float, start: int = 0) -> np.ndarray:
    k = tensor_key(seed, tensor_id)
    i = (np.arange(start, start + n, dtype=np.uint64) + np.uint64(k)) & np.uint64(M32)
    h = mix32(i)
    u = (h >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return (u * np.float32(2.0) - np.float32(1.0)) * np.float32(scale)


def weight_scale(fan_in: int) -> float:
    return float(np.float32(math.sqrt(3.0 / float(fan_in))))


BIAS_SCALE = float(np.float32(0.1))


def align256(x: int) -> int:
    return (x + 255) & ~255


def mlp_manifest(dims, activations=None) -> dict:
    """Layout restated from the product format (DESIGN.md 'model bundle'): per layer W[in,out]
    row-major fp32 then b[out], every tensor 256-byte aligned."""
    n_layers = len(dims) - 1
    if activations is None:
        activations = ["relu"] * (n_layers - 1) + ["linear"]
    off, layers = 0, []
    for l in range(n_layers):
        fi, fo = dims[l], dims[l + 1]
        w_off = off
        off = align256(off + fi * fo * 4)
        b_off = off
        off = align256(off + fo * 4)
        layers.append({"in": fi, "out": fo, "activation": activations[l],
                       "w_offset": w_off, "b_offset": b_off})
    return {"format": "tfsc-b200-v1", "template": "mlp", "dtype": "float32",
            "signature": {"input": "x", "output": "y"}, "layers": layers, "weights_bytes": off}


def synth_mlp_blob(dims, seed: int, activations=None):
    man = mlp_manifest(dims, activations)
    blob = np.zeros(man["weights_bytes"] // 4, dtype=np.float32)
    for l, L in enumerate(man["layers"]):
        w = synth_tensor(seed, 2 * l, L["in"] * L["out"], weight_scale(L["in"]))
        b = synth_tensor(seed, 2 * l + 1, L["out"], BIAS_SCALE)
        blob[L["w_offset"] // 4: L["w_offset"] // 4 + w.size] = w
        blob[L["b_offset"] // 4: L["b_offset"] // 4 + b.size] = b
    return man, blob


def affine_manifest() -> dict:
    return {"format": "tfsc-b200-v1", "template": "affine", "dtype": "float32",
            "signature": {"input": "x", "output": "y"}, "a_offset": 0, "b_offset": 256,
            "weights_bytes": 512}


def affine_blob(a: float, b: float):
    man = affine_manifest()
    blob = np.zeros(man["weights_bytes"] // 4, dtype=np.float32)
    blob[0] = a
    blob[64] = b
    return man, blob


def load_bundle(version_dir: str):
    with open(os.path.join(version_dir, "tfsc_model.json")) as f:
        man = json.load(f)
    blob = np.fromfile(os.path.join(version_dir, "weights.bin"), dtype=np.float32)
    return man, blob


def forward(man: dict, blob: np.ndarray, x: np.ndarray, dtype=np.float32) -> np.ndarray:
    """Forward pass in ``dtype`` (float32 = the reference executor's arithmetic type; float64 is
    the arbiter for tolerance questions)."""
    t = man["template"]
    if t == "affine":
        a = dtype(blob[man["a_offset"] // 4])
        b = dtype(blob[man["b_offset"] // 4])
        return (x.astype(dtype) * a + b).astype(dtype)
    if t == "mlp":
        h = x.astype(dtype)
        for L in man["layers"]:
            w = blob[L["w_offset"] // 4: L["w_offset"] // 4 + L["in"] * L["out"]].reshape(L["in"], L["out"])
            b = blob[L["b_offset"] // 4: L["b_offset"] // 4 + L["out"]]
            h = h @ w.astype(dtype) + b.astype(dtype)
            if L["activation"] == "relu":
                h = np.maximum(h, dtype(0))
        return h
    if t == "graph":
        return graph_forward(man, blob, x, dtype)
    raise ValueError(f"unknown template {t}")


def resnet50_ops(image=224, classes=1000):
    """Independent restatement of the ResNet-50 v1.5 topology (He et al. 2015; stride on the 3x3 conv as in
    torchvision / the TF official model), expressed in the bundle's op list."""
    ops = []
    cfg = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]
    size = image
    ops.append(dict(op="conv", src=-1, dst=0, h=size, w=size, c=3, kh=7, kw=7, stride=2, pad=3, cout=64, act="relu"))
    size = (size + 2 * 3 - 7) // 2 + 1
    ops.append(dict(op="maxpool", src=0, dst=1, h=size, w=size, c=64, kh=3, kw=3, stride=2, pad=1))
    size = (size + 2 - 3) // 2 + 1
    cur, cin = 1, 64
    for planes, n_blocks, first_stride in cfg:
        for blk in range(n_blocks):
            stride = first_stride if blk == 0 else 1
            others = [i for i in range(5) if i != cur]
            t1, t2, t3, t4 = others[0], others[1], others[2], others[3]
            out_size = (size + 2 - 3) // stride + 1
            ops.append(dict(op="conv", src=cur, dst=t1, h=size, w=size, c=cin, kh=1, kw=1, stride=1, pad=0, cout=planes, act="relu"))
            ops.append(dict(op="conv", src=t1, dst=t2, h=size, w=size, c=planes, kh=3, kw=3, stride=stride, pad=1, cout=planes, act="relu"))
            shortcut = cur
            if blk == 0:
                ops.append(dict(op="conv", src=cur, dst=t3, h=size, w=size, c=cin, kh=1, kw=1, stride=stride, pad=0, cout=4 * planes, act="none"))
                shortcut = t3
            ops.append(dict(op="conv", src=t2, dst=t4, res=shortcut, h=out_size, w=out_size, c=planes, kh=1, kw=1, stride=1, pad=0,
                            cout=4 * planes, act="relu"))
            cur, cin, size = t4, 4 * planes, out_size
    t1 = [i for i in range(5) if i != cur][0]
    ops.append(dict(op="avgpool", src=cur, dst=t1, h=size, w=size, c=cin))
    ops.append(dict(op="dense", src=t1, dst=-2, h=1, w=1, c=cin, cout=classes, act="none"))
    return ops


def graph_manifest(input_shape, ops, n_buffers=5):
    off = 0
    for o in ops:
        if o["op"] in ("conv", "dense"):
            k = o.get("kh", 1) * o.get("kw", 1) * o["c"]
            o["w_offset"] = off
            off = align256(off + k * o["cout"] * 4)
            o["b_offset"] = off
            off = align256(off + o["cout"] * 4)
    return {"format": "tfsc-b200-v1", "template": "graph", "dtype": "float32", "signature": {"input": "x", "output": "y"},
            "input_shape": list(input_shape), "n_buffers": n_buffers, "ops": ops, "weights_bytes": off}


def synth_graph_blob(man: dict, seed: int) -> np.ndarray:
    blob = np.zeros(man["weights_bytes"] // 4, dtype=np.float32)
    for i, o in enumerate(man["ops"]):
        if o["op"] not in ("conv", "dense"):
            continue
        fan_in = o.get("kh", 1) * o.get("kw", 1) * o["c"]
        w = synth_tensor(seed, 2 * i, fan_in * o["cout"], weight_scale(fan_in))
        b = synth_tensor(seed, 2 * i + 1, o["cout"], BIAS_SCALE)
        blob[o["w_offset"] // 4: o["w_offset"] // 4 + w.size] = w
        blob[o["b_offset"] // 4: o["b_offset"] // 4 + b.size] = b
    return blob


def graph_forward(man: dict, blob: np.ndarray, x: np.ndarray, dtype=np.float64) -> np.ndarray:
    """Interpreter of a graph bundle with torch-CPU functional ops (conv2d / max_pool2d) in `dtype`; NHWC in and
    out, NCHW inside."""
    import torch
    import torch.nn.functional as F
    td = torch.float64 if dtype == np.float64 else torch.float32
    ish = man["input_shape"]
    xb = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).reshape(-1, *ish).to(td)
    bufs = {-1: xb.permute(0, 3, 1, 2).contiguous()}
    for o in man["ops"]:
        src = bufs[o["src"]]
        if o["op"] in ("conv", "dense"):
            kh, kw, c, cout = o.get("kh", 1), o.get("kw", 1), o["c"], o["cout"]
            w = torch.from_numpy(blob[o["w_offset"] // 4: o["w_offset"] // 4 + kh * kw * c * cout].reshape(kh, kw, c, cout)).to(td)
            b = torch.from_numpy(blob[o["b_offset"] // 4: o["b_offset"] // 4 + cout]).to(td)
            if o["op"] == "dense":
                y = src.reshape(src.shape[0], -1) @ w.reshape(c, cout) + b
                y = y.reshape(src.shape[0], cout, 1, 1)
            else:
                y = F.conv2d(src, w.permute(3, 2, 0, 1).contiguous(), b, stride=o["stride"], padding=o["pad"])
            if o.get("res", -100) != -100:
                y = y + bufs[o["res"]]
            if o.get("act", "none") == "relu":
                y = torch.relu(y)
            elif o.get("act") == "gelu":
                y = F.gelu(y)
        elif o["op"] == "maxpool":
            y = F.max_pool2d(src, (o["kh"], o["kw"]), stride=o["stride"], padding=o["pad"])
        elif o["op"] == "avgpool":
            y = src.mean(dim=(2, 3), keepdim=True)
        else:
            raise ValueError(o["op"])
        bufs[o["dst"]] = y
    out = bufs[-2]
    out = out.permute(0, 2, 3, 1).reshape(out.shape[0], -1) if out.shape[2] * out.shape[3] == 1 else out.permute(0, 2, 3, 1)
    return out.numpy().astype(dtype)
