"""Host mirror of pkg/tfservingproxy/tfservingproxy.go: URL matching, the 404/400 bodies, gRPC
ModelSpec extraction, and RestProxy / GrpcProxy objects whose "director" is the on-GPU server."""
from __future__ import annotations

import ctypes as C

from ._lib import check, lib


def match_rest_url(url: str):
    """tfServingRestURLMatch + Serve status logic (:24, :93-129) -> (status, name, version, body)."""
    name = C.create_string_buffer(1024)
    ver = C.create_string_buffer(64)
    st = check(lib.tfsc_rest_match_url(url.encode(), name, 1024, ver, 64), "rest_match_url")
    body = lib.tfsc_rest_error_body(st).decode() if st != 200 else ""
    return st, name.value.decode(), ver.value.decode(), body


def parse_version(version: str) -> int:
    out = C.c_int64()
    rc = lib.tfsc_parse_version(version.encode(), C.byref(out))
    if rc < 0:
        raise ValueError(lib.tfsc_last_error().decode())
    return out.value


def grpc_model_spec(request_bytes: bytes):
    """clientForSpec (:246-250): (name, version string) from a serialized request."""
    name = C.create_string_buffer(1024)
    ver = C.create_string_buffer(32)
    check(lib.tfsc_grpc_model_spec(request_bytes, len(request_bytes), name, 1024, ver, 32), "grpc_model_spec")
    return name.value.decode(), ver.value.decode()


class RestProxy:
    """NewRestProxy(handler) (:53): `server` plays the director + backend."""

    def __init__(self, server):
        self.server = server

    def serve(self, method: str, url: str, body: bytes = b""):
        return self.server.rest_handle(method, url, body)


class GrpcProxy:
    """NewGrpcProxy (:76): Predict on serialized bytes; MultiInference is unsupported (:215-217)."""

    def __init__(self, server):
        self.server = server

    def predict(self, request_bytes: bytes) -> bytes:
        return self.server.grpc_predict(request_bytes)

    def multi_inference(self, request_bytes: bytes):
        raise NotImplementedError("MultiInference not supported")
