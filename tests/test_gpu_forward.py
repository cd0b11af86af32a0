"""-m gpu: the forward hop between PROCESSES (a6 / X7, taskhandler.go:95-147) -- one rank per GPU like torchrun, or two
ranks sharing cuda:0 on a 1-GPU box (CUDA IPC works within a device too, so the driver's 1-GPU GPUTEST exercises the same
code: window export / import, control channel, gather / scatter kernels on peer memory). Plus the asynchronous ticket API,
deadlines and the batching window, which share the request path."""
import multiprocessing as mp
import os
import tempfile
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DIMS = [96, 160, 24]
N_MODELS = 24


def _cfg(rank, world, socks, n_gpus, **kw):
    members = [f"gpu{i}:0:0" for i in range(world)]
    cfg = {"modelProvider.type": "synthetic", "modelProvider.synthetic.dims": DIMS, "modelProvider.synthetic.count": N_MODELS,
           "gpu.devices": [rank % n_gpus], "gpu.arenaBytes": 32 << 20, "modelCache.size": 1 << 28, "serving.maxConcurrentModels": 64,
           "gpu.members": members, "gpu.localMembers": [members[rank]], "proxy.replicasPerModel": 1, "proxy.replicaPick": "first",
           "cluster.rank": rank, "cluster.endpoints": socks, "cluster.slotBytes": 1 << 16, "cluster.windowSlots": 16,
           "proxy.grpcTimeout": 20.0}
    cfg.update(kw)
    return cfg


def _rank_main(rank, world, socks, n_gpus, barrier, out):
    try:
        import json

        import torch
        import tfservingcache_b200 as t
        from oracle import wire
        torch.cuda.set_device(rank % n_gpus)
        res = {"rank": rank, "y": {}, "errors": []}
        with t.Server(_cfg(rank, world, socks, n_gpus)) as srv:
            barrier.wait(timeout=120)          # every listener is up
            # every rank dials every other rank at the same moment (what bench.py does before its timed regions): with >= 3
            # ranks an acceptor that waits for its own rank's dial in progress closes a cycle -- the N = 8 hang of round 2
            for p in range(world):
                if p != rank:
                    srv.fwd_peer_window(p)
            rng = np.random.default_rng(100 + rank)
            owned = []
            for j in range(N_MODELS):
                nodes, _picked = srv.route(f"m{j}", "1")
                owned.append(nodes[0] >= 0)
                x = rng.standard_normal((1 + j % 3, DIMS[0])).astype(np.float32)
                y = srv.predict(f"m{j}", "1", x)
                res["y"][j] = (x, y)
            res["owned"] = owned
            # gRPC and REST entry points take the hop too
            j = next(j for j in range(N_MODELS) if not owned[j])
            x = rng.standard_normal((2, DIMS[0])).astype(np.float32)
            _spec, outs = wire.decode_predict_response(srv.grpc_predict(wire.encode_predict_request(f"m{j}", 1, {"x": x})))
            st, body = srv.rest_handle("POST", f"/v1/models/m{j}/versions/1:predict", json.dumps({"instances": x.tolist()}).encode())
            res["wire"] = (j, x, outs["y"], st, np.array(json.loads(body)["predictions"]))
            # errors come back from the owner: unknown model, wrong signature
            bad = next(k for k in range(N_MODELS, N_MODELS + 400) if srv.route(f"m{k}", "1")[0][0] < 0)
            for name, xx in ((f"m{bad}", x), (f"m{j}", x[:, :5])):
                try:
                    srv.predict(name, "1", xx)
                    res["errors"].append(None)
                except t._lib.TfscError as e:
                    res["errors"].append(e.code)
            # the cache tier of an explicit member (no ring lookup): the peer serves one of OUR models when asked directly
            jl = next(k for k in range(N_MODELS) if owned[k])
            res["member"] = (jl, x, srv.predict_member((rank + 1) % world, f"m{jl}", "1", x), srv.predict_member(rank, f"m{j}", "1", x))
            # asynchronous tickets, local and forwarded, all in flight at once
            tickets = [(jj, xx, srv.predict_submit(f"m{jj}", "1", xx)) for jj, (xx, _y) in list(res["y"].items())[:12]]
            res["tickets"] = [(jj, xx, tk.wait(30.0)) for jj, xx, tk in tickets]
            for _jj, _xx, tk in tickets:
                tk.release()
            # device-resident hop: the owner's kernels read x / write y in the ingress rank's window (tfsc_predict_device)
            my_rank, win, _bytes, slot = srv.fwd_window()
            assert my_rank == rank
            peer = (rank + 1) % world
            pwin, pbytes = srv.fwd_peer_window(peer)
            xw = rng.standard_normal((3, DIMS[0])).astype(np.float32)
            # this rank plays INGRESS with slot 15 of its own window: x goes there, the peer is told through the barrier
            t._lib.check(t._lib.lib.tfsc_device_memcpy(win + 15 * slot, xw.ctypes.data, xw.nbytes))
            barrier.wait(timeout=120)          # every rank's slot 15 holds its x
            jo = next(jj for jj in range(N_MODELS) if owned[jj])
            srv.ensure(0, f"m{jo}", 1)
            srv.predict_device(0, f"m{jo}", 1, pwin + 15 * slot, 3, pwin + 15 * slot + slot // 2, 0)   # peer memory in, peer memory out
            srv.sync(0)
            barrier.wait(timeout=120)          # every owner finished writing into its peer's window
            yw = np.empty((3, DIMS[-1]), np.float32)
            t._lib.check(t._lib.lib.tfsc_device_memcpy(yw.ctypes.data, win + 15 * slot + slot // 2, yw.nbytes))
            res["device_hop"] = (xw, yw)   # computed by the PEER with the peer's model
            res["stats"] = srv.stats()
            barrier.wait(timeout=120)          # nobody tears down while a peer still needs it
        out.put(res)
    except BaseException as e:  # noqa: BLE001
        import traceback
        out.put({"rank": rank, "fatal": f"{e!r}\n{traceback.format_exc()}"})
        try:
            barrier.abort()
        except Exception:
            pass


def _ref(j, x):
    from oracle import models
    man, blob = models.synth_mlp_blob(DIMS, seed=1000 + j)
    return models.forward(man, blob, x, np.float64)


@pytest.mark.parametrize("world", [2, 4])
def test_ranks_forward_requests_over_the_window(world):
    import torch
    assert torch.cuda.is_available()
    n_gpus = torch.cuda.device_count()
    if os.environ.get("TFSC_REQUIRE_MULTI") == "1":
        assert n_gpus >= 2, "TFSC_REQUIRE_MULTI=1: this run must exercise two physical GPUs (NVLink)"
    tmp = tempfile.mkdtemp(prefix="tfscfwd")
    socks = [os.path.join(tmp, f"r{r}.sock") for r in range(world)]
    ctx = mp.get_context("spawn")
    barrier, out = ctx.Barrier(world), ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, socks, min(n_gpus, world), barrier, out)) for r in range(world)]
    [p.start() for p in procs]
    results = {}
    deadline = time.time() + 300
    while len(results) < world and time.time() < deadline:
        try:
            r = out.get(timeout=5)
            results[r["rank"]] = r
        except Exception:
            if not any(p.is_alive() for p in procs):
                break
    [p.join(timeout=30) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert len(results) == world, f"ranks reported: {sorted(results)}"
    for r in results.values():
        assert "fatal" not in r, r.get("fatal")
    owners = [[r for r in range(world) if results[r]["owned"][j]] for j in range(N_MODELS)]
    assert all(len(o) == 1 for o in owners), "every model has exactly one owner"
    assert all(any(results[r]["owned"]) for r in range(world))
    for rank, r in results.items():
        for j, (x, y) in r["y"].items():     # local and forwarded answers are the same numbers
            ref = _ref(j, x)
            assert y.shape == ref.shape and np.max(np.abs(y - ref) / np.maximum(1, np.abs(ref))) <= 1e-4, (rank, j)
        j, x, yg, st, yr = r["wire"]
        assert st == 200 and np.max(np.abs(yg - _ref(j, x))) <= 1e-4 and np.max(np.abs(yr - _ref(j, x))) <= 1e-4
        assert r["errors"] == [-5, -3]       # NOT_FOUND and INVALID_ARGUMENT travel back from the owner
        jl, xm, y_peer, y_self = r["member"]
        assert np.max(np.abs(y_peer - _ref(jl, xm))) <= 1e-4 and np.max(np.abs(y_self - _ref(j, xm))) <= 1e-4
        for j, x, y in r["tickets"]:
            assert np.max(np.abs(y - _ref(j, x))) <= 1e-4
        st = r["stats"]
        n_remote = sum(1 for o in r["owned"] if not o)
        assert st["fwd_out_requests"] >= n_remote + 2 and st["fwd_in_requests"] >= 1 and st["fwd_out_failures"] == 2
        assert st["fwd_peer_bytes_read"] > 0 and st["fwd_peer_bytes_written"] > 0
    # device-resident hop: rank r's window slot was processed by rank (r-1) % world with ITS first owned model
    for rank, r in results.items():
        xw, y = r["device_hop"]
        owner = (rank - 1) % world
        jo = next(jj for jj in range(N_MODELS) if results[owner]["owned"][jj])
        assert np.max(np.abs(y - _ref(jo, xw))) <= 1e-4


def test_deadline_and_batching_window():
    import torch
    import tfservingcache_b200 as t
    assert torch.cuda.is_available()
    cfg = {"modelProvider.type": "synthetic", "modelProvider.synthetic.dims": DIMS, "modelProvider.synthetic.count": 4,
           "gpu.devices": [0], "gpu.arenaBytes": 16 << 20, "modelCache.size": 1 << 28, "gpu.tickMicros": 20000, "gpu.maxBatch": 8}
    with t.Server(cfg) as srv:
        x = np.random.default_rng(0).standard_normal((1, DIMS[0])).astype(np.float32)
        srv.predict("m0", "1", x)                       # resident
        st0 = srv.stats()
        # 6 single-row requests submitted within the 20 ms window ride in ONE batch
        tk = [srv.predict_submit("m0", "1", x) for _ in range(6)]
        ys = [k.wait(10.0) for k in tk]
        [k.release() for k in tk]
        st1 = srv.stats()
        assert st1["batches"] - st0["batches"] == 1 and st1["batched_rows"] - st0["batched_rows"] == 6
        assert all(np.array_equal(ys[0], y) for y in ys) and np.max(np.abs(ys[0] - _ref(0, x))) <= 1e-4
        # a deadline that passes while the request waits in the batching window -> DEADLINE_EXCEEDED, never launched
        with pytest.raises(t._lib.TfscError) as e:
            srv.predict_deadline("m0", "1", x, srv.now_ns() + 2_000_000)
        assert e.value.code == t._lib.E_TIMEOUT
        assert srv.stats()["batches"] == st1["batches"]
        # a generous deadline is met
        y = srv.predict_deadline("m0", "1", x, srv.now_ns() + 5_000_000_000)
        assert np.max(np.abs(y - _ref(0, x))) <= 1e-4
        # wait() with a short timeout reports "still in flight" and the ticket stays usable
        k = srv.predict_submit("m1", "1", x)
        try:
            k.wait(0.0005)
        except t._lib.TfscError as e2:
            assert e2.code == t._lib.E_TIMEOUT
        assert np.max(np.abs(k.wait(10.0) - _ref(1, x))) <= 1e-4
        k.release()
