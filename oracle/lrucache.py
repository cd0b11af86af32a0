"""Restatement of pkg/cachemanager/lrucache.go:20-105 (byte-capacity LRU; PINNED by
lrucache_test.go:7-115).  File deletion on eviction (lrucache.go:73-78) has no analogue in the
HBM/pinned-host tiers and is reported through ``evicted`` instead."""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass


@dataclass(frozen=True)
class ModelIdentifier:  # cachemanager.go:51-54
    model_name: str
    version: int


@dataclass
class Model:  # cachemanager.go:45-49
    identifier: ModelIdentifier
    path: str
    size_on_disk: int


class LRUCache:
    def __init__(self, base_dir: str, capacity_in_bytes: int):
        self.base_dir = base_dir
        self.capacity = capacity_in_bytes
        self.current_size = 0
        self._od: "OrderedDict[ModelIdentifier, Model]" = OrderedDict()  # last = MRU front
        self.evicted: list[ModelIdentifier] = []

    def get(self, item: ModelIdentifier):  # lrucache.go:43-51
        if item in self._od:
            self._od.move_to_end(item)
            return self._od[item], True
        return None, False

    def put(self, item: ModelIdentifier, model: Model):  # lrucache.go:54-65
        if item not in self._od:
            self.ensure_free_bytes(model.size_on_disk)
            self._od[item] = model
            self.current_size += model.size_on_disk
        else:
            self._od.move_to_end(item)

    def ensure_free_bytes(self, nbytes: int):  # lrucache.go:68-87
        while len(self._od) > 0 and self.capacity - self.current_size < nbytes:
            ident, m = self._od.popitem(last=False)
            self.current_size -= m.size_on_disk
            self.evicted.append(ident)

    def list_models(self) -> list[Model]:  # lrucache.go:89-97, MRU -> LRU
        return list(reversed(self._od.values()))

    def __len__(self):
        return len(self._od)
