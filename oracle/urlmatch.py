"""Restatement of the protocol shim's request parsing (PINNED by tfservingproxy_test.go:111-234).

  pkg/tfservingproxy/tfservingproxy.go:24      tfServingRestURLMatch
  pkg/tfservingproxy/tfservingproxy.go:93-129  RestProxy.Serve: 404 / 400 bodies
  pkg/tfservingproxy/tfservingproxy.go:246-250 clientForSpec: version = FormatInt(spec.version)
  pkg/cachemanager/cachemanager.go:294-309     handleModelRequest: ParseInt(version, 10, 64)
"""
from __future__ import annotations

import json
import re

_URL_RE = re.compile(r"^/v1/models/(?P<modelName>[^/]+)(/versions/(?P<version>[0-9]+))?", re.IGNORECASE)


def _go_json_line(status: str, message: str) -> str:
    # json.NewEncoder(rw).Encode(struct{Status, Message}) -> compact JSON + "\n"
    return json.dumps({"Status": status, "Message": message}, separators=(",", ":")) + "\n"


NOT_FOUND_BODY = _go_json_line("Error", "Not found")
NO_VERSION_BODY = _go_json_line("Error", "Model version must be provided")


def match_rest_url(url: str):
    """Returns (http_status, model_name, version_string, error_body)."""
    m = _URL_RE.match(url)
    if m is None:
        return 404, "", "", NOT_FOUND_BODY
    version = m.group("version") or ""
    if version == "":
        return 400, m.group("modelName"), "", NO_VERSION_BODY
    return 200, m.group("modelName"), version, ""


def grpc_version_string(version_value) -> str:
    """clientForSpec: a missing Int64Value yields GetValue() == 0 -> "0"."""
    return str(int(version_value or 0))


def parse_version(version: str) -> int:
    """strconv.ParseInt(version, 10, 64): optional sign, decimal digits, int64 range."""
    if not re.fullmatch(r"[+-]?[0-9]+", version):
        raise ValueError(f'strconv.ParseInt: parsing "{version}": invalid syntax')
    v = int(version)
    if not (-(1 << 63) <= v < (1 << 63)):
        raise ValueError(f'strconv.ParseInt: parsing "{version}": value out of range')
    return v
