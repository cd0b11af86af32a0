"""Host mirror of pkg/cachemanager/lrucache.go (ModelCache :11-18, LRUCache :20-105) over the
C ABI (tfsc_lru_*).  Same method names and semantics as the reference type."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from ._lib import check, lib


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
    def __init__(self, base_dir: str, capacity_in_bytes: int):  # NewLRUCache
        self._h = lib.tfsc_lru_new(base_dir.encode(), capacity_in_bytes)
        self._base_dir = base_dir

    def __del__(self):
        if getattr(self, "_h", None):
            lib.tfsc_lru_free(self._h)
            self._h = None

    def put(self, item: ModelIdentifier, model: Model) -> int:
        return check(lib.tfsc_lru_put(self._h, item.model_name.encode(), item.version, model.path.encode(),
                                      model.size_on_disk), "lru_put")

    def get(self, item: ModelIdentifier):
        size = C.c_int64()
        buf = C.create_string_buffer(4096)
        rc = check(lib.tfsc_lru_get(self._h, item.model_name.encode(), item.version, C.byref(size), buf, 4096), "lru_get")
        if rc == 0:
            return None, False
        return Model(item, buf.value.decode(), size.value), True

    def ensure_free_bytes(self, nbytes: int) -> int:
        return check(lib.tfsc_lru_ensure_free_bytes(self._h, nbytes), "lru_ensure_free_bytes")

    def list_models(self) -> list[Model]:
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            rc = lib.tfsc_lru_list(self._h, buf, cap)
            if rc == -21:
                cap *= 4
                continue
            check(rc, "lru_list")
            break
        out = []
        for line in buf.value.decode().splitlines():
            name, ver, size, path = line.split("\t")
            out.append(Model(ModelIdentifier(name, int(ver)), path, int(size)))
        return out

    @property
    def current_size(self) -> int:
        return lib.tfsc_lru_current_size(self._h)

    @property
    def capacity(self) -> int:
        return lib.tfsc_lru_capacity(self._h)

    def __len__(self):
        return lib.tfsc_lru_len(self._h)

    def base_dir(self) -> str:
        return self._base_dir

    def model_path(self, model: Model) -> str:
        return self._base_dir.rstrip("/") + "/" + model.path
