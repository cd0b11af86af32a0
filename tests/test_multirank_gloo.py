"""world_size-2 `gloo` test (CPU) of the N>1 path of bench.py: every rank derives the same global
Zipf trace and ring placement without communication, the ranks' shards partition the requests, and
the max-over-ranks / sum-over-ranks aggregation works."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, zlib
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    import bench
    import tfservingcache_b200 as t
    from oracle import ring as oring

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%d" %% int(os.environ["PORT"]),
                            rank=int(os.environ["RANK"]), world_size=2)
    rank, world = dist.get_rank(), 2
    wl = bench.build_workload(world, 16, 64, 3)
    assert wl["replicas"] == 2 and wl["n_models"] == 32 and len(wl["trace"]) == 64 * 2 * 3
    # identical on every rank (no communication needed on the data path)
    sig = torch.tensor([zlib.crc32(wl["trace"].tobytes()), zlib.crc32(wl["dest"].tobytes())], dtype=torch.int64)
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    assert all(torch.equal(s, sigs[0]) for s in sigs)
    # placement == oracle ring (bit-exact), owner of every request is one of its replicas
    oc = oring.ClusterConnection(2)
    oc.update([oring.ServingService.from_string(m) for m in wl["members"]])
    for j in range(wl["n_models"]):
        want = [int(s.host[3:]) for s in oc.find_node_for_key(oring.model_key("m%%d" %% j, "1"))]
        assert list(wl["owners"][j]) == want
    assert all(d in wl["owners"][m] for m, d in zip(wl["trace"], wl["dest"]))
    # shards partition every step
    for step in range(3):
        mine, groups = bench.step_groups(wl, rank, step, 64 * world)
        cnt = torch.tensor([len(mine)], dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        assert int(cnt) == 64 * world and sum(r for _m, r in groups) == len(mine)
    # timing aggregation: max over ranks
    tt = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    assert float(tt) == 2.0
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29600 + os.getpid() % 300
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in o, o[-2000:]
