"""Debug: per-k-block timeline of CTA (0,0) of dense_tc_kernel (needs a -DTFSC_TC_TRACE build)."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfservingcache_b200 as t
lib = t._lib.lib
lib.tfsc_k_dense_tc.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p]
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16
k = n = 9216
x = torch.randn(rows, k, device="cuda"); w = torch.randn(k, n, device="cuda") / 96; b = torch.randn(n, device="cuda")
y = torch.empty(rows, n, device="cuda")
wsb = lib.tfsc_k_dense_workspace(rows, k, n); ws = torch.zeros(wsb // 4 + 64, device="cuda")
for _ in range(3):
    lib.tfsc_k_dense_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, k, n, 1, ws.data_ptr(), wsb, None)
torch.cuda.synchronize()
tr = np.zeros((8, 128), np.int64)
assert lib.tfsc_tc_trace_read(tr.ctypes.data_as(C.c_void_p)) == 0
t0 = tr[0, 0]
names = ["tma_issue", "mma_start", "mma_issued", "conv_begin", "cempty_ok", "full_ok", "conv_done"]
print("kb " + " ".join(f"{n_:>10s}" for n_ in names))
for kb in list(range(0, 14)) + list(range(60, 72)):
    print(f"{kb:2d} " + " ".join(f"{tr[i, kb] - t0:10d}" for i in range(7)))
print("per-kblock period (clk), kb 20..60:", (tr[6, 60] - tr[6, 20]) / 40)
