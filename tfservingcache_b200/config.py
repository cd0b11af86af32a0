"""Configuration surface of cmd/taskhandler/cfg.go:10-66: config.yaml in the CWD, overridden by
TFSC_<KEY with '.' -> '_'> environment variables (viper AutomaticEnv), flattened to the dotted
keys the library reads.  Parsed once, not per request."""
from __future__ import annotations

import json
import os

DEFAULTS = {"healthprobe.modelName": "__TFSERVINGCACHE_PROBE_CHECK__"}  # cfg.go:64-66

KNOWN_KEYS = [
    "proxyRestPort", "proxyGrpcPort", "cacheRestPort", "cacheGrpcPort",
    "metrics.path", "metrics.timeout", "metrics.modelLabels",
    "modelProvider.type", "modelProvider.diskProvider.baseDir", "modelProvider.diskProvider.basePath",
    "modelCache.hostModelPath", "modelCache.size",
    "serving.servingModelPath", "serving.grpcHost", "serving.restHost", "serving.maxConcurrentModels",
    "serving.grpcConfigTimeout", "serving.grpcPredictTimeout", "serving.grpcMaxMsgSize", "serving.metricsPath",
    "serving.modelFetchTimeout",
    "proxy.replicasPerModel", "proxy.grpcTimeout", "proxy.replicaPick", "proxy.seed", "proxy.hotFraction",
    "logging.level", "logging.format", "healthprobe.modelName",
    "gpu.devices", "gpu.arenaBytes", "gpu.maxBatch", "gpu.maxRequestRows", "gpu.stagingSlots",
    "gpu.members", "gpu.localMembers", "gpu.blockingSync",
]


def _flatten(d, prefix, out):
    for k, v in d.items():
        key = f"{prefix}.{k}" if prefix else str(k)
        if isinstance(v, dict) and not key.startswith("serviceDiscovery"):
            _flatten(v, key, out)
        else:
            out[key] = v


def _coerce(v: str):
    try:
        return json.loads(v)
    except (ValueError, TypeError):
        return v


def load_config(path: str | None = "config.yaml", env=None, overrides: dict | None = None) -> dict:
    env = os.environ if env is None else env
    cfg = dict(DEFAULTS)
    if path and os.path.exists(path):
        import yaml
        with open(path) as f:
            _flatten(yaml.safe_load(f) or {}, "", cfg)
    lower = {k.lower(): k for k in set(KNOWN_KEYS) | set(cfg)}
    for ek, ev in env.items():
        if not ek.upper().startswith("TFSC_"):
            continue
        dotted = ek[5:].lower()
        for lk, orig in lower.items():
            if lk.replace(".", "_") == dotted:
                cfg[orig] = _coerce(ev)
    if "TFSC_LOGLEVEL" in env and "logging.level" not in cfg:  # docker-compose.yaml:21 quirk
        cfg["logging.level"] = env["TFSC_LOGLEVEL"]
    if overrides:
        cfg.update(overrides)
    return cfg
