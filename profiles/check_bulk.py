"""Parity + timing of the bulk-copy ring variant (tfsc_k_dense_variant 2) against the LDG-stream variant (1) and an
fp64 torch reference. Writes gpurun_out/bulk_check.json. Usage: python profiles/check_bulk.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfservingcache_b200 as t  # noqa: E402

lib = t._lib.lib
out = {"parity": [], "timing": []}
torch.manual_seed(0)


def run(variant, x, w, b, relu, ws, ws_bytes):
    rows, K = x.shape
    N = w.shape[1]
    y = torch.full((rows, N), float("nan"), device="cuda")
    t._lib.check(lib.tfsc_k_dense_variant(variant, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, K, N, relu,
                                          ws.data_ptr(), ws_bytes, None))
    torch.cuda.synchronize()
    return y


for (K, N) in [(64, 8), (100, 520), (1000, 512), (777, 1032), (4096, 4096), (9216, 9216), (9216, 8), (20000, 64)]:
    for rows in (1, 2, 3, 4, 5, 8):
        for relu in (0, 1):
            x = torch.randn(rows, K, device="cuda")
            w = torch.randn(K, N, device="cuda") / K ** 0.5
            b = torch.randn(N, device="cuda")
            ws_bytes = lib.tfsc_k_dense_workspace(rows, K, N)
            ws = torch.zeros(ws_bytes // 4 + 64, device="cuda")
            ref = x.double() @ w.double() + b.double()
            if relu:
                ref = ref.clamp_min(0)
            y1 = run(1, x, w, b, relu, ws, ws_bytes)
            e1 = (y1.double() - ref).abs().max().item()
            for variant in (2, 4):
                y2 = run(variant, x, w, b, relu, ws, ws_bytes)
                y2b = run(variant, x, w, b, relu, ws, ws_bytes)     # second launch: self-resetting counters, determinism
                e2 = (y2.double() - ref).abs().max().item()
                out["parity"].append({"K": K, "N": N, "rows": rows, "relu": relu, "variant": variant, "err_ldg": e1, "err_bulk": e2,
                                      "deterministic": bool(torch.equal(y2, y2b)),
                                      "ok": bool(e2 < 2e-4 and torch.equal(y2, y2b))})
            del x, w, b, ws
bad = [p for p in out["parity"] if not p["ok"]]
print(f"parity cases {len(out['parity'])} bad {len(bad)}", bad[:5])

K = N = 9216
ws_l = [torch.randn(K, N, device="cuda") / 96 for _ in range(3)]
b = torch.randn(N, device="cuda")
for rows in (1, 2, 4, 8):
    x = torch.randn(rows, K, device="cuda")
    y = torch.empty(rows, N, device="cuda")
    ws_bytes = lib.tfsc_k_dense_workspace(rows, K, N)
    ws = torch.zeros(ws_bytes // 4 + 64, device="cuda")
    for variant in (1, 2, 4):
        iters = 12
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        torch.cuda.synchronize()
        for i in range(iters):
            ev[i].record()
            t._lib.check(lib.tfsc_k_dense_variant(variant, x.data_ptr(), ws_l[i % 3].data_ptr(), b.data_ptr(), y.data_ptr(), rows, K, N,
                                                  1, ws.data_ptr(), ws_bytes, None))
        ev[iters].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(3, iters))
        alg = K * N * 4 + N * 4 + rows * (K + N) * 4
        rec = {"rows": rows, "variant": variant, "us_median": ms[len(ms) // 2] * 1e3, "us_min": ms[0] * 1e3,
               "GBps_median": alg / ms[len(ms) // 2] / 1e6}
        out["timing"].append(rec)
        print(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bulk_check.json", "w"), indent=1)
sys.exit(1 if bad else 0)
