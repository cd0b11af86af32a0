"""clock64 timeline of CTA 0 of the persistent tcgen05 GEMM (TFSC_GT_TRACE=1) for a few graph-model shapes.
Usage: TFSC_GT_TRACE=1 python profiles/r2/trace_gemm.py"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tfservingcache_b200 as t  # noqa: E402

lib = t._lib.lib
lib.tfsc_debug_gemm_trace.argtypes = [C.POINTER(C.c_longlong)]
names = ["entry", "setup_done", "first_tma", "first_tile_landed", "first_converted", "first_mma", "tile0_commit", "tile0_epi_start",
         "tile0_epi_done", "tile1_commit", "tile1_epi_start", "tile1_epi_done", "exit", "t0c0_tmem_loaded", "t0c0_staged", "t0c0_stored"]
for (m, n, k) in [(25088, 64, 64), (25088, 256, 64), (1024, 3072, 768), (1024, 768, 3072), (6272, 128, 128)]:
    a = torch.randn(m, k, device="cuda"); b = torch.randn(k, n, device="cuda") / k ** 0.5; bias = torch.randn(n, device="cuda")
    c = torch.empty(m, n, device="cuda")
    for _ in range(3):
        t._lib.check(lib.tfsc_k_gemm_tc(a.data_ptr(), b.data_ptr(), bias.data_ptr(), None, c.data_ptr(), m, n, k, k, 1, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.tfsc_k_gemm_tc(a.data_ptr(), b.data_ptr(), bias.data_ptr(), None, c.data_ptr(), m, n, k, k, 1, None)
    e1.record(); torch.cuda.synchronize()
    out = (C.c_longlong * 16)()
    t._lib.check(lib.tfsc_debug_gemm_trace(out))
    st = list(out)[:16]
    rel = {nm: (st[i] - st[0]) for i, nm in enumerate(names) if st[i]}
    print(json.dumps({"M": m, "N": n, "K": k, "us_per_launch": round(e0.elapsed_time(e1) * 1e3 / 20, 2), "clk_from_entry": rel}))
