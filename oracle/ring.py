"""Restatement of the reference's model->node routing (pure Python, integer-exact).

Follows:
  pkg/taskhandler/taskhandler.go:84-92   nodeForKey: key = name + "##" + version, random pick
  pkg/taskhandler/cluster.go:104-113     clusterUpdated: members -> "host:rest:grpc" -> Set
  pkg/taskhandler/cluster.go:116-130     FindNodeForKey: GetN(key, max(replicasPerModel, 1))
  pkg/taskhandler/cluster.go:142-164     ServingService.String / serviceFromString
and the published algorithm of stathat.com/c/consistent v1.0.0 (go.mod:25; absent from
/root/reference, restated): 20 virtual points per member at crc32_ieee(str(i) + member), a
map hash->member (later insert overwrites, remove deletes the point), ascending sorted hashes,
lookup = first point strictly greater than crc32_ieee(key), wrapping to index 0.
Pinned on the module's own published test vectors, recalled from the public source (see oracle/__init__.py and
tests/test_oracle_pins.py::test_ring_reproduces_upstream_module_test_vectors).
"""
from __future__ import annotations

import bisect
from dataclasses import dataclass

NUMBER_OF_REPLICAS = 20  # consistent.New(): NumberOfReplicas = 20

_CRC_TABLE = []
for _n in range(256):
    _c = _n
    for _ in range(8):
        _c = (_c >> 1) ^ 0xEDB88320 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32_ieee(data: bytes) -> int:
    """Go hash/crc32.ChecksumIEEE (reflected poly 0xEDB88320, init/xorout 0xFFFFFFFF)."""
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


class EmptyCircleError(Exception):
    """consistent.ErrEmptyCircle -> request fails (cluster.go:118-120)."""


class Consistent:
    def __init__(self):
        self.circle: dict[int, str] = {}
        self.members: dict[str, bool] = {}
        self.sorted_hashes: list[int] = []
        self.count = 0

    @staticmethod
    def elt_key(elt: str, idx: int) -> bytes:
        return (str(idx) + elt).encode()

    def _update_sorted(self):
        self.sorted_hashes = sorted(self.circle.keys())

    def add(self, elt: str):
        for i in range(NUMBER_OF_REPLICAS):
            self.circle[crc32_ieee(self.elt_key(elt, i))] = elt
        self.members[elt] = True
        self._update_sorted()
        self.count += 1

    def remove(self, elt: str):
        for i in range(NUMBER_OF_REPLICAS):
            self.circle.pop(crc32_ieee(self.elt_key(elt, i)), None)
        self.members.pop(elt, None)
        self._update_sorted()
        self.count -= 1

    def set(self, elts: list[str]):
        for k in list(self.members.keys()):
            if k not in elts:
                self.remove(k)
        for v in elts:
            if v in self.members:
                continue
            self.add(v)

    def _search(self, key: int) -> int:
        i = bisect.bisect_right(self.sorted_hashes, key)
        return 0 if i >= len(self.sorted_hashes) else i

    def get_n(self, name: str, n: int) -> list[str]:
        if not self.circle:
            raise EmptyCircleError("empty circle")
        if self.count < n:
            n = self.count
        key = crc32_ieee(name.encode())
        i = self._search(key)
        start = i
        res = [self.circle[self.sorted_hashes[i]]]
        if len(res) == n:
            return res
        i = start + 1
        while i != start:
            if i >= len(self.sorted_hashes):
                i = 0
            elem = self.circle[self.sorted_hashes[i]]
            if elem not in res:
                res.append(elem)
            if len(res) == n:
                break
            i += 1
        return res


@dataclass(frozen=True)
class ServingService:
    host: str
    grpc_port: int
    rest_port: int

    def __str__(self) -> str:  # cluster.go:142-144
        return f"{self.host}:{self.rest_port}:{self.grpc_port}"

    @staticmethod
    def from_string(s: str) -> "ServingService":  # cluster.go:146-164
        parts = s.split(":")
        return ServingService(host=parts[0], rest_port=int(parts[1]), grpc_port=int(parts[2]))


class ClusterConnection:
    """cluster.go:44-130 without the discovery goroutine: members are pushed with update()."""

    def __init__(self, replicas_per_model: float = 0):
        self.consistent = Consistent()
        self.replicas_per_model = replicas_per_model

    def update(self, members: list[ServingService]):
        self.consistent.set([str(m) for m in members])

    def find_node_for_key(self, key: str) -> list[ServingService]:
        n = int(max(self.replicas_per_model, 1))
        return [ServingService.from_string(s) for s in self.consistent.get_n(key, n)]


def model_key(model_name: str, version: str) -> str:
    return model_name + "##" + version  # taskhandler.go:85


def node_for_key(cluster: ClusterConnection, model_name: str, version: str, rand_intn) -> ServingService:
    """taskhandler.go:84-92; ``rand_intn(n)`` stands in for math/rand.Intn."""
    nodes = cluster.find_node_for_key(model_key(model_name, version))
    return nodes[rand_intn(len(nodes))]
