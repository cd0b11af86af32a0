"""Seeded Zipf(alpha) multi-model request trace: p(rank r) = 1/(r^a H_N),
rank -> model id through a seeded permutation.  Workload generator for bench.py and the tests (not part of the oracle: the product arm of bench.py must not import oracle/)."""
from __future__ import annotations

import numpy as np


def zipf_trace(n_models: int, n_requests: int, alpha: float = 1.0, seed: int = 42) -> np.ndarray:
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, n_models + 1, dtype=np.float64) ** alpha
    p /= p.sum()
    perm = rng.permutation(n_models)
    ranks = rng.choice(n_models, size=n_requests, p=p)
    return perm[ranks].astype(np.int64)


def uniform_trace(n_models: int, n_requests: int, seed: int = 42) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, n_models, size=n_requests).astype(np.int64)
