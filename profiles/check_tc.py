"""Quick on-GPU check + timing of the tensor-core dense path (tfsc_k_dense_tc) against fp64."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfservingcache_b200 as t  # noqa: E402

lib = t._lib.lib
import ctypes as C
lib.tfsc_k_dense_tc.restype = C.c_int
lib.tfsc_k_dense_tc.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p]
shapes = [(16, 128, 256), (16, 512, 256), (9, 512, 512), (32, 1024, 1024), (33, 1000, 512), (48, 2048, 768), (64, 4096, 2048),
          (16, 9216, 9216), (64, 9216, 9216)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for rows, k, n in shapes:
    g = torch.Generator(device="cuda").manual_seed(rows * 131 + k + n)
    x = torch.randn(rows, k, device="cuda", generator=g)
    w = torch.randn(k, n, device="cuda", generator=g) / (k ** 0.5)
    b = torch.randn(n, device="cuda", generator=g)
    y = torch.full((rows, n), float("nan"), device="cuda")
    ws_bytes = lib.tfsc_k_dense_workspace(rows, k, n)
    ws = torch.zeros(ws_bytes // 4 + 64, device="cuda")
    for rep in range(2):
        y.fill_(float("nan"))
        rc = lib.tfsc_k_dense_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, k, n, 1, ws.data_ptr(), ws_bytes, None)
        assert rc == 0, (rc, lib.tfsc_last_error())
        torch.cuda.synchronize()
    ref = torch.relu(x.double() @ w.double() + b.double())
    err = ((y.double() - ref).abs() / ref.abs().clamp(min=1.0)).max().item()
    nan = torch.isnan(y).sum().item()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    iters = 10
    ev[0].record()
    for _ in range(iters):
        lib.tfsc_k_dense_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, k, n, 1, ws.data_ptr(), ws_bytes, None)
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 1e3 / iters
    print(f"rows={rows:3d} k={k:5d} n={n:5d}  max rel err={err:.3e}  nan={nan}  {us:8.1f} us/launch  {k * n * 4 / us / 1e3:8.1f} GB/s(W)", flush=True)
