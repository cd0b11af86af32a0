"""Host mirror of pkg/taskhandler/cluster.go + taskhandler.go:84-92 over the C ABI (tfsc_ring_*).
The discovery goroutine of the reference is replaced by an explicit update() -- on one box the
membership is the static list of GPUs (discovery services are out of scope)."""
from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass

from ._lib import check, lib


def crc32_ieee(data: bytes) -> int:
    return lib.tfsc_crc32_ieee(data, len(data))


def model_key(model_name: str, version: str) -> str:
    buf = C.create_string_buffer(len(model_name) + len(version) + 8)
    check(lib.tfsc_model_key(model_name.encode(), version.encode(), buf, len(buf)), "model_key")
    return buf.value.decode()


@dataclass(frozen=True)
class ServingService:  # cluster.go:15-19
    host: str
    grpc_port: int
    rest_port: int

    def __str__(self):  # cluster.go:142-144
        return f"{self.host}:{self.rest_port}:{self.grpc_port}"

    @staticmethod
    def from_string(s: str) -> "ServingService":  # cluster.go:146-164
        parts = s.split(":")
        return ServingService(parts[0], rest_port=int(parts[1]), grpc_port=int(parts[2]))


class ClusterConnection:
    def __init__(self, replicas_per_model: float = 0):
        self._h = lib.tfsc_ring_new()
        self.replicas_per_model = replicas_per_model

    def __del__(self):
        if getattr(self, "_h", None):
            lib.tfsc_ring_free(self._h)
            self._h = None

    def update(self, members: list[ServingService]):  # clusterUpdated, cluster.go:104-113
        strs = [str(m).encode() for m in members]
        arr = (C.c_char_p * len(strs))(*strs)
        check(lib.tfsc_ring_set(self._h, arr, len(strs)), "ring_set")

    def find_node_for_key(self, key: str) -> list[ServingService]:  # cluster.go:116-130
        n = int(max(self.replicas_per_model, 1))
        cap = 256 * n + 64
        buf = C.create_string_buffer(cap)
        cnt = check(lib.tfsc_ring_getn(self._h, key.encode(), n, buf, cap), "ring_getn")
        if cnt == 0:
            return []
        return [ServingService.from_string(s) for s in buf.value.decode().split("\n")]

    @property
    def members(self) -> int:
        return lib.tfsc_ring_members(self._h)

    @property
    def points(self) -> int:
        return lib.tfsc_ring_points(self._h)


class ReplicaPicker:
    """Replica choice among the GetN candidates: "random" (reference, taskhandler.go:91), "first",
    "hot-spread" (primary unless the key is hot), "balanced" (hot-spread + least-loaded-replica binding).  Deterministic for a given seed + call sequence."""

    def __init__(self, policy: str = "random", seed: int = 0, hot_fraction: float = 0.5):
        self._h = lib.tfsc_picker_new(policy.encode(), seed, hot_fraction)
        if not self._h:
            raise ValueError(lib.tfsc_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib.tfsc_picker_free(self._h)
            self._h = None

    def pick(self, key: str, n_replicas: int, members: int) -> int:
        return check(lib.tfsc_picker_pick(self._h, key.encode(), n_replicas, members), "picker_pick")

    def pick_ids(self, key: str, member_ids, members: int) -> int:
        """Same with a stable integer id per candidate replica (needed by the "balanced" policy)."""
        arr = (C.c_int * len(member_ids))(*[int(v) for v in member_ids])
        return check(lib.tfsc_picker_pick_ids(self._h, key.encode(), arr, len(member_ids), members), "picker_pick_ids")


class TaskHandler:
    """taskhandler.go:20-92: key = name + "##" + version, uniform random pick among replicas."""

    def __init__(self, cluster: ClusterConnection, seed=None):
        self.cluster = cluster
        self._rand = random.Random(seed)

    def node_for_key(self, model_name: str, version: str) -> ServingService:
        nodes = self.cluster.find_node_for_key(model_key(model_name, version))
        return nodes[self._rand.randrange(len(nodes))]
