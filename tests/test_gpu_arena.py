"""-m gpu: the HBM arena under MIXED model sizes. First-fit alone would answer a fragmented arena by
evicting down the LRU; the node instead packs idle resident blocks together with device-to-device copies and keeps them."""
import os

import numpy as np
import pytest

import tfservingcache_b200 as t

pytestmark = pytest.mark.gpu


def _bundle(base, name, fi, fo, seed):
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal((fi, fo)) / np.sqrt(fi)).astype(np.float32)
    b = rng.standard_normal(fo).astype(np.float32)
    man = t.modelformat.write_mlp_bundle(os.path.join(base, name, "1"), [w], [b], ["relu"])
    return w, b, man["weights_bytes"]


def test_fragmented_arena_is_compacted_not_flushed(tmp_path):
    import torch
    assert torch.cuda.is_available()
    base = str(tmp_path)
    models = {}
    for name, fo in (("s1", 512), ("b1", 3072), ("s2", 512), ("b2", 3072), ("m2", 1536)):
        models[name] = _bundle(base, name, 512, fo, sum(map(ord, name)))
    sz = {k: v[2] for k, v in models.items()}
    assert sz["s1"] == 1050624 and sz["b1"] == 6303744 and sz["m2"] == 3151872
    arena = 17 << 20
    cfg = {"modelProvider.type": "diskProvider", "modelProvider.diskProvider.baseDir": base, "gpu.devices": [0],
           "gpu.arenaBytes": arena, "modelCache.size": 1 << 28, "serving.maxConcurrentModels": 4}
    x = np.random.default_rng(1).standard_normal((3, 512)).astype(np.float32)

    def ref(name):
        w, b, _ = models[name]
        return np.maximum(x.astype(np.float64) @ w + b, 0)

    with t.Server(cfg) as srv:
        for name in ("s1", "b1", "s2", "b2"):       # laid out back to back from offset 0; 3.1 MB stay free at the end
            assert np.max(np.abs(srv.predict(name, "1", x) - ref(name))) <= 1e-4
        st0 = srv.stats()
        assert st0["arena_bytes_used"] == sz["s1"] + sz["b1"] + sz["s2"] + sz["b2"] and st0["evictions_hbm"] == 0
        # m2 (3.15 MB) pushes s1 (the LRU) out of the 4-model resident prefix: 1.05 MB at the front + 3.1 MB at the back are
        # free -- enough in total, no single hole large enough
        y = srv.predict("m2", "1", x)
        assert np.max(np.abs(y - ref("m2"))) <= 1e-4
        st1 = srv.stats()
        assert st1["evictions_hbm"] - st0["evictions_hbm"] == 1, "only s1 leaves; first-fit alone would also evict b1"
        assert st1["arena_compactions"] == 1 and st1["arena_compacted_bytes"] == sz["b1"] + sz["s2"] + sz["b2"]
        assert {n for n, *_ in srv.resident(0)} == {"m2", "b2", "s2", "b1"}
        assert st1["h2d_weight_bytes"] - st0["h2d_weight_bytes"] == sz["m2"]      # nothing was paged in again
        for name in ("b1", "s2", "b2", "m2"):         # the moved blocks still hold their weights (hits, no reload)
            assert np.max(np.abs(srv.predict(name, "1", x) - ref(name))) <= 1e-4
        st2 = srv.stats()
        assert st2["h2d_weight_bytes"] == st1["h2d_weight_bytes"] and st2["cache_hits_total"] - st1["cache_hits_total"] == 4
