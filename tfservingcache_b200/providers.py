"""Host mirror of pkg/cachemanager/modelproviders/diskmodelprovider (version-dir matching and
size); loading itself happens inside the library (provider.cc) straight into pinned memory."""
from __future__ import annotations

import ctypes as C

from ._lib import check, lib


class DiskModelProvider:
    def __init__(self, base_dir: str):
        self.base_dir = base_dir

    def find_src_path_for_model(self, model_name: str, version: int) -> str:  # :46-69
        buf = C.create_string_buffer(4096)
        rc = lib.tfsc_disk_find_version_dir(self.base_dir.encode(), model_name.encode(), version, buf, 4096)
        if rc < 0:
            raise FileNotFoundError(lib.tfsc_last_error().decode())
        return buf.value.decode()

    def model_size(self, model_name: str, version: int) -> int:  # :71-83 (fixed: recursive bytes)
        rc = lib.tfsc_disk_model_size(self.base_dir.encode(), model_name.encode(), version)
        if rc < 0:
            raise FileNotFoundError(lib.tfsc_last_error().decode())
        return rc

    def check(self) -> bool:  # :85-88
        return True
