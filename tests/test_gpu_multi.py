"""-m gpu, needs >= 2 GPUs (skipped otherwise): one process driving two ring members, and the
forward hop a6/X7 as NVLink peer access -- the request tensor lives on GPU i, the owner is GPU j, the owner's
kernels read x and write y in place on GPU i (no host bounce, no copy kernel)."""
import numpy as np
import pytest

import tfservingcache_b200 as t
from oracle import models
from oracle import ring as oring

pytestmark = pytest.mark.gpu
DIMS = [256, 520, 136, 32]


def _need2():
    import os
    import torch
    if os.environ.get("TFSC_REQUIRE_MULTI") == "1":   # a multi-GPU run must not pass by skipping
        assert torch.cuda.is_available() and torch.cuda.device_count() >= 2, "TFSC_REQUIRE_MULTI=1 needs >= 2 GPUs"
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the cross-process forward hop is covered on one GPU by tests/test_gpu_forward.py)")
    return torch


def _cfg(**kw):
    cfg = {"modelProvider.type": "synthetic", "modelProvider.synthetic.dims": DIMS, "modelProvider.synthetic.count": 64,
           "gpu.devices": [0, 1], "gpu.arenaBytes": 16 << 20, "modelCache.size": 1 << 30, "serving.maxConcurrentModels": 8,
           "proxy.replicasPerModel": 1, "proxy.seed": 3, "proxy.replicaPick": "random"}
    cfg.update(kw)
    return cfg


def _ref(j, x):
    man, blob = models.synth_mlp_blob(DIMS, seed=1000 + j)
    return models.forward(man, blob, x, np.float64)


def test_two_members_route_like_the_oracle_and_serve():
    _need2()
    rng = np.random.default_rng(0)
    with t.Server(_cfg()) as srv:
        assert srv.num_nodes == 2
        oc = oring.ClusterConnection(1)
        oc.update([oring.ServingService.from_string(m) for m in ("gpu0:0:0", "gpu1:0:0")])
        seen = set()
        for j in range(24):
            nodes, picked = srv.route(f"m{j}", "1")
            want = int(oc.find_node_for_key(oring.model_key(f"m{j}", "1"))[0].host[3:])
            assert nodes == [want] and picked == 0
            x = rng.standard_normal((3, DIMS[0])).astype(np.float32)
            y = srv.predict(f"m{j}", "1", x)
            assert np.max(np.abs(y - _ref(j, x)) / np.maximum(1, np.abs(_ref(j, x)))) <= 1e-4
            seen.add(want)
        assert seen == {0, 1}
        s0, s1 = srv.stats(0), srv.stats(1)
        assert s0["cache_total"] + s1["cache_total"] == 24 and s0["cache_total"] > 0 and s1["cache_total"] > 0


@pytest.mark.parametrize("rows", [1, 8, 40])
def test_forward_hop_is_peer_access(rows):
    torch = _need2()
    rng = np.random.default_rng(rows)
    with t.Server(_cfg()) as srv:
        # find a model owned by GPU 1, keep its request tensor on GPU 0
        j = next(j for j in range(64) if srv.route(f"m{j}", "1")[0] == [1])
        srv.ensure(1, f"m{j}", 1)
        x = rng.standard_normal((rows, DIMS[0])).astype(np.float32)
        xd = torch.from_numpy(x).to("cuda:0")
        yd = torch.full((rows, DIMS[-1]), float("nan"), device="cuda:0")
        torch.cuda.synchronize(0)
        srv.predict_device(1, f"m{j}", 1, xd.data_ptr(), rows, yd.data_ptr(), 0)   # node 1 computes, tensors on GPU 0
        srv.sync(1)
        ref = _ref(j, x)
        got = yd.cpu().numpy()
        assert not np.isnan(got).any()
        assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))) <= 1e-4
