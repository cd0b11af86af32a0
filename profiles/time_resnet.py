"""Device-resident timing of the ResNet-50 graph bundle (BASELINE configs[1] model) through tfsc_predict_device."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfservingcache_b200 as t
man = t.modelformat.resnet50_manifest()
cfg = {"modelProvider.type": "synthetic", "modelProvider.synthetic.template": "manifest", "modelProvider.synthetic.manifest": man,
       "modelProvider.synthetic.count": 8, "gpu.devices": [0], "gpu.arenaBytes": 2 << 30, "serving.maxConcurrentModels": 8,
       "modelCache.size": 4 << 30}
srv = t.Server(cfg)
for j in range(4):
    srv.ensure(0, f"m{j}", 1)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for rows in (1, 2, 4, 8, 16):
    x = torch.rand(rows, 224, 224, 3, device="cuda"); y = torch.empty(rows, 1000, device="cuda")
    for _ in range(3):
        srv.predict_device(0, "m0", 1, x.data_ptr(), rows, y.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    l0 = t._lib.lib.tfsc_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record(st)
    for i in range(iters):
        srv.predict_device(0, f"m{i % 4}", 1, x.data_ptr(), rows, y.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"rows={rows:2d}  {ms:8.3f} ms/batch  {rows / ms * 1e3:9.1f} img/s  {(t._lib.lib.tfsc_kernel_launches() - l0) // iters} launches  {8.2 * rows / ms:7.2f} TFLOP/s", flush=True)
srv.close()
