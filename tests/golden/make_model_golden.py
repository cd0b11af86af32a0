"""Generates tests/golden/model_torch_golden.json: fp64 outputs of torchvision's ResNet-50 and transformers' BERT-base
(seeded weights and inputs, tests/torch_export.py) -- the independent numeric pin of the executor the reference delegates to
(TF-Serving is not available; these libraries define the two BASELINE model families). The tests rebuild the same models
live (torchvision / transformers are in the image here and on the GPU box) and additionally require the committed
numbers when the library versions match, so a silent change of seeding or topology is caught.

    python tests/golden/make_model_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch_export as te  # noqa: E402

CASES = {
    "resnet50": dict(seed=5, input_seed=1, batch=2),
    "resnet_small": dict(seed=6, input_seed=3, batch=3, blocks=(1, 1, 1, 1), classes=10, image=64),
    "bert_base": dict(seed=3, input_seed=2, batch=8),
    "bert_small": dict(seed=4, input_seed=5, batch=4, seq=16, hidden=64, layers=2, heads=4, inter=128, vocab=100, max_pos=32, labels=3),
}


def resnet_case(c):
    m = te.torchvision_resnet(c["seed"], c.get("blocks", (3, 4, 6, 3)), c.get("classes", 1000))
    img = c.get("image", 224)
    x = np.random.default_rng(c["input_seed"]).random((c["batch"], img, img, 3)).astype(np.float32)
    return m, x, te.resnet_reference(m, x)


def bert_case(c):
    kw = {k: c[k] for k in ("seq", "hidden", "layers", "heads", "inter", "vocab", "max_pos", "labels") if k in c}
    m = te.hf_bert(c["seed"], **kw)
    seq, vocab = kw.get("seq", 128), kw.get("vocab", 30522)
    ids = np.random.default_rng(c["input_seed"]).integers(1, vocab, (c["batch"], seq)).astype(np.int32)
    ids[-1, seq // 2:] = 0           # [PAD] tail -> masked keys
    if c["batch"] > 3:
        ids[3, seq - seq // 4:] = 0
    return m, ids, te.bert_reference(m, ids)


if __name__ == "__main__":
    import torch
    import torchvision
    import transformers
    out = {"versions": {"torch": torch.__version__, "torchvision": torchvision.__version__, "transformers": transformers.__version__},
           "cases": {}}
    for name, c in CASES.items():
        _m, _x, ref = (resnet_case if name.startswith("resnet") else bert_case)(c)
        out["cases"][name] = {"config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items()},
                              "shape": list(ref.shape), "logits": [float(v) for v in ref.ravel()]}
        print(name, ref.shape, float(np.abs(ref).max()))
    with open(os.path.join(HERE, "model_torch_golden.json"), "w") as f:
        json.dump(out, f)
