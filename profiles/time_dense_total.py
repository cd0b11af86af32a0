"""Many-launch timing of one tenant-model layer (9216x9216 fp32) through tfsc_k_dense_variant: total CUDA-event time over
N back-to-back launches / N (per-launch event pairs are quantised to ~2 us on these boxes and hide small differences).
Three weight buffers rotate so no launch finds its W in L2. PDL is a process-level switch:
    for p in 0 1; do for v in 1 2 4; do TFSC_PDL=$p python profiles/time_dense_total.py $v; done; done
Usage: python profiles/time_dense_total.py [variant=1] [launches=300] [rows=1,2,4,8]  -> one JSON line per row count
(variant 3 = tensor-core path for every row count, e.g. rows 16,32,48,64)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfservingcache_b200 as t  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 300
row_list = tuple(int(v) for v in sys.argv[3].split(",")) if len(sys.argv) > 3 else (1, 2, 4, 8)
K = N = 9216
lib = t._lib.lib
w3 = [torch.randn(K, N, device="cuda") / 96 for _ in range(3)]
b = torch.randn(N, device="cuda")
for rows in row_list:
    x = torch.randn(rows, K, device="cuda")
    y = torch.empty(rows, N, device="cuda")
    ws_bytes = lib.tfsc_k_dense_workspace(rows, K, N)
    ws = torch.zeros(ws_bytes // 4 + 64, device="cuda")

    def go(n):
        for i in range(n):
            t._lib.check(lib.tfsc_k_dense_variant(variant, x.data_ptr(), w3[i % 3].data_ptr(), b.data_ptr(), y.data_ptr(), rows, K, N, 1,
                                                  ws.data_ptr(), ws_bytes, None))
    go(30)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    go(launches)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / launches
    alg = K * N * 4 + N * 4 + rows * (K + N) * 4
    print(json.dumps({"variant": variant, "pdl": os.environ.get("TFSC_PDL", "0"), "rows": rows, "launches": launches,
                      "us_per_launch": round(us, 2), "GBps": round(alg / us / 1e3, 1)}))
