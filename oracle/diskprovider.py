"""Restatement of pkg/cachemanager/modelproviders/diskmodelprovider/diskmodelprovider.go:20-88
(PINNED by diskmodelprovider_test.go:33-87): the version directory is the one whose *name parses
to the requested integer* (``000000042`` == 42); several matches -> the last one in ReadDir
(sorted) order wins; none -> "No matching model found".

Deliberate fix: ModelSize is the recursive byte size of the version
directory, not the directory inode size (diskmodelprovider.go:76-82), because an HBM / pinned-host
byte budget needs real bytes."""
from __future__ import annotations

import os

from .urlmatch import parse_version


def find_src_path_for_model(model_dir: str, model_version: int) -> str:
    match = None
    for name in sorted(os.listdir(model_dir)):  # ioutil.ReadDir sorts by filename
        try:
            v = parse_version(name)
        except ValueError:
            continue
        if v == model_version and os.path.isdir(os.path.join(model_dir, name)):
            match = name
    if match is None:
        raise FileNotFoundError("No matching model found")
    return os.path.join(model_dir, match)


def model_size(base_dir: str, model_name: str, model_version: int) -> int:
    src = find_src_path_for_model(os.path.join(base_dir, model_name), model_version)
    total = 0
    for root, _dirs, files in os.walk(src):
        for f in files:
            total += os.path.getsize(os.path.join(root, f))
    return total
