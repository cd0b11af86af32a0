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


def write_mlp_bundle(version_dir: str, weights, biases, activations=None, input_name="x", output_name="y"):
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
    _write(version_dir, man, blob)
    return man


def write_affine_bundle(version_dir: str, a: float, b: float, input_name="x", output_name="y"):
    blob = np.zeros(128, dtype=np.float32)
    blob[0], blob[64] = a, b
    man = {"format": "tfsc-b200-v1", "template": "affine", "dtype": "float32",
           "signature": {"input": input_name, "output": output_name}, "a_offset": 0, "b_offset": 256,
           "weights_bytes": 512}
    _write(version_dir, man, blob)
    return man


def _write(version_dir: str, man: dict, blob: np.ndarray):
    os.makedirs(version_dir, exist_ok=True)
    with open(os.path.join(version_dir, "tfsc_model.json"), "w") as f:
        json.dump(man, f)
    blob.astype("<f4").tofile(os.path.join(version_dir, "weights.bin"))
