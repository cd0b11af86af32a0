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
    raise ValueError(f"unknown template {t}")
