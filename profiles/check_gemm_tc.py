"""On-GPU check + timing of the tcgen05 GEMM (tfsc_k_gemm_tc) against fp64 and against the CUDA-core GEMM."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfservingcache_b200 as t
lib = t._lib.lib
shapes = [(128, 128, 32, 32), (128, 64, 64, 64), (300, 256, 96, 96), (3136, 64, 147, 148), (1024, 768, 768, 768),
          (1024, 3072, 768, 768), (1024, 768, 3072, 3072), (25088, 64, 64, 64), (6272, 128, 1152, 1152), (392, 2048, 512, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for m, n, k, lda in shapes:
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, lda, device="cuda", generator=g)
    b = torch.randn(k, n, device="cuda", generator=g) / (k ** 0.5)
    bias = torch.randn(n, device="cuda", generator=g)
    r = torch.randn(m, n, device="cuda", generator=g)
    ref = torch.relu(a[:, :k].double() @ b.double() + bias.double() + r.double())
    res = {}
    for name, fn in (("tc", lib.tfsc_k_gemm_tc), ("simt", lib.tfsc_k_gemm)):
        c = torch.full((m, n), float("nan"), device="cuda")
        rc = fn(a.data_ptr(), b.data_ptr(), bias.data_ptr(), r.data_ptr(), c.data_ptr(), m, n, k, lda, 1, None)
        assert rc == 0, (name, rc, lib.tfsc_last_error())
        torch.cuda.synchronize()
        err = ((c.double() - ref).abs() / ref.abs().clamp(min=1.0)).max().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(a.data_ptr(), b.data_ptr(), bias.data_ptr(), r.data_ptr(), c.data_ptr(), m, n, k, lda, 1, None)
        e1.record(); torch.cuda.synchronize()
        res[name] = (err, torch.isnan(c).sum().item(), e0.elapsed_time(e1) / 5 * 1e3)
    tf = 2.0 * m * n * k
    print(f"M={m:6d} N={n:5d} K={k:5d}  tc: err={res['tc'][0]:.2e} nan={res['tc'][1]} {res['tc'][2]:8.1f} us {tf / res['tc'][2] / 1e6:7.1f} TF/s | "
          f"simt: err={res['simt'][0]:.2e} {res['simt'][2]:8.1f} us {tf / res['simt'][2] / 1e6:6.1f} TF/s", flush=True)
