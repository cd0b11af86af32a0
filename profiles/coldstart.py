"""Cold-start / paging measurement (BASELINE configs[4] 'cold-start storm'): uniform-random
requests over M tenant models (1.02 GB each) with an HBM arena that holds only A of them, host tier warm, so
almost every request is a host->HBM page-in (cachemanager.go:133-143 'reload' branch) through pinned
cudaMemcpyAsync on the copy stream.  Reports load-stall latency percentiles and achieved H2D bandwidth.
Usage: python profiles/coldstart.py [models=48] [arena_models=8] [requests=400] [clients=8]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tfservingcache_b200 as t  # noqa: E402
from tfservingcache_b200 import _lib  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 48
A = int(sys.argv[2]) if len(sys.argv) > 2 else 8
NREQ = int(sys.argv[3]) if len(sys.argv) > 3 else 400
CLIENTS = int(sys.argv[4]) if len(sys.argv) > 4 else 8
DIMS = [9216, 9216, 9216, 9216]
MB = 1019326464
cfg = {"modelProvider.type": "synthetic", "modelProvider.synthetic.dims": DIMS, "modelProvider.synthetic.count": M,
       "modelProvider.synthetic.threads": 32, "gpu.devices": [0], "gpu.arenaBytes": int(A * (MB + 4096)), "gpu.maxBatch": 64,
       "modelCache.size": int((M + 1) * MB), "serving.maxConcurrentModels": 1 << 20, "proxy.replicaPick": "first"}
srv = t.Server(cfg)
t0 = time.time()
for j in range(M):
    srv.ensure(0, f"m{j}", 1)          # cold misses: provider synth -> pinned host -> HBM (LRU keeps the last A resident)
cold_s = time.time() - t0
s0 = srv.stats()
lg = C.CDLL(os.path.join(ROOT, "tools", "libtfsc_loadgen.so"))
lg.tfsc_loadgen_run.restype = C.c_int64
names = b"".join(f"m{j}".encode().ljust(16, b"\0") for j in range(M))
trace = np.random.default_rng(5).integers(0, M, NREQ).astype(np.int32)
x = torch.randn(64, DIMS[0]).pin_memory()
y = torch.empty(CLIENTS, DIMS[-1]).pin_memory()
lat = np.zeros(NREQ, np.float32)
el = C.c_double()
failed = lg.tfsc_loadgen_run(C.cast(_lib.lib.tfsc_predict, C.c_void_p), C.c_void_p(srv._h), names, 16, b"1",
                             trace.ctypes.data_as(C.c_void_p), C.c_int64(NREQ), C.c_void_p(x.data_ptr()), C.c_int64(64), DIMS[0],
                             C.c_void_p(y.data_ptr()), DIMS[-1], CLIENTS, lat.ctypes.data_as(C.c_void_p), C.byref(el))
s1 = srv.stats()
reloads = (s1["cache_total"] - s0["cache_total"]) - (s1["cache_hits_total"] - s0["cache_hits_total"]) - (s1["cache_misses_total"] - s0["cache_misses_total"])
h2d = s1["h2d_weight_bytes"] - s0["h2d_weight_bytes"]
out = {"what": "cold-start storm, uniform requests, host tier warm, HBM arena too small", "models": M, "arena_models": A,
       "requests": NREQ, "clients": CLIENTS, "failed": int(failed), "elapsed_s": round(el.value, 3),
       "qps": round(NREQ / el.value, 1), "reloads": int(reloads), "hits": int(s1["cache_hits_total"] - s0["cache_hits_total"]),
       "misses": int(s1["cache_misses_total"] - s0["cache_misses_total"]), "h2d_weight_gb": round(h2d / 1e9, 1),
       "h2d_gb_per_s": round(h2d / 1e9 / el.value, 1), "evictions_hbm": int(s1["evictions_hbm"] - s0["evictions_hbm"]),
       "load_stall_ms": {"p50": round(float(np.percentile(lat, 50)) / 1e3, 1), "p90": round(float(np.percentile(lat, 90)) / 1e3, 1),
                         "p99": round(float(np.percentile(lat, 99)) / 1e3, 1), "max": round(float(lat.max()) / 1e3, 1)},
       "first_touch": {"models": M, "seconds": round(cold_s, 1), "per_model_ms": round(1e3 * cold_s / M, 1),
                       "note": "provider miss: synthesize 1.02 GB into pinned memory + H2D"},
       "reference_floor": "reference: >= one 500 ms status-poll quantum per load under a global lock (cachemanager.go:175-193)"}
print(json.dumps(out))
srv.close()
