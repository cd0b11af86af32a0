"""Device-resident timing of a graph bundle (ResNet-50 = BASELINE configs[1] model, BERT-base = configs[3] model)
through tfsc_predict_device (CUDA events on the launching stream).
Usage: python profiles/time_graph.py resnet50|bert [rows ...]"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tfservingcache_b200 as t
from oracle import models
kind = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
rows_list = [int(v) for v in sys.argv[2:]] or ([1, 8] if kind == "resnet50" else [8])
if kind == "resnet50":
    man = t.modelformat.resnet50_manifest(); oman = models.graph_manifest([224, 224, 3], models.resnet50_ops()); flop = 8.2e9
    mk = lambda r: np.random.default_rng(0).random((r, 224, 224, 3)).astype(np.float32)
else:
    man = t.modelformat.bert_manifest(); oman = models.graph_manifest([128], models.bert_ops(), 4, ("input_ids", "logits"), "int32"); flop = 22.5e9
    mk = lambda r: np.random.default_rng(0).integers(1, 30522, (r, 128)).astype(np.int32)
cfg = {"modelProvider.type": "synthetic", "modelProvider.synthetic.template": "manifest", "modelProvider.synthetic.manifest": man,
       "modelProvider.synthetic.count": 8, "gpu.devices": [0], "gpu.arenaBytes": 4 << 30, "serving.maxConcurrentModels": 8, "modelCache.size": 8 << 30}
srv = t.Server(cfg)
for j in range(4):
    srv.ensure(0, f"m{j}", 1)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
out = []
for rows in rows_list:
    xh = mk(rows)
    x = torch.from_numpy(xh).cuda(); y = torch.empty(rows, man["ops"][-1]["cout"], device="cuda")
    for _ in range(3):
        srv.predict_device(0, "m0", 1, x.data_ptr(), rows, y.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record(st)
    for i in range(iters):
        srv.predict_device(0, f"m{i % 4}", 1, x.data_ptr(), rows, y.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    rec = {"model": kind, "rows": rows, "gpu_ms": round(ms, 3), "gpu_items_per_s": round(rows / ms * 1e3, 1), "gpu_tflops": round(flop * rows / ms / 1e9, 2),
           }
    print(json.dumps(rec), flush=True)
srv.close()
