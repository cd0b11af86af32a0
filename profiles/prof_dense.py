"""Stand-alone driver for ncu captures of the dominant kernel (dense_stream_kernel<R>) through the
C ABI entry tfsc_k_dense, on one tenant-model layer (9216x9216 fp32, 339.7 MB) with only ~1.1 GB
of device memory allocated, so `ncu --set full` kernel replay stays cheap.
Usage: python profiles/prof_dense.py [rows] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfservingcache_b200 as t  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
K = N = 9216
lib = t._lib.lib
ws_bytes = lib.tfsc_k_dense_workspace(rows, K, N)
ws = torch.zeros(ws_bytes // 4 + 64, device="cuda")
ws_ = [torch.randn(K, N, device="cuda") / 96 for _ in range(3)]  # 3 layers: consecutive launches never hit L2
b = torch.randn(N, device="cuda")
x = torch.randn(rows, K, device="cuda")
y = torch.empty(rows, N, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
torch.cuda.synchronize()
for i in range(iters):
    ev[i].record()
    t._lib.check(lib.tfsc_k_dense(x.data_ptr(), ws_[i % 3].data_ptr(), b.data_ptr(), y.data_ptr(), rows, K, N, 1,
                                  ws.data_ptr(), ws_bytes, None))
ev[iters].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
alg = K * N * 4 + N * 4 + rows * (K + N) * 4
print(f"rows={rows} per-launch ms={['%.4f' % m for m in ms]} best GB/s={alg / min(ms) / 1e6:.1f}")
