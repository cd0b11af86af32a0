"""Server: the cache tier + proxy tier (cmd/taskhandler/main.go:45-113) for this process's GPUs,
over the C ABI.  predict() = proxyServiceServer.Predict with the forward replaced by on-GPU
execution; grpc_predict()/rest_handle() are the wire-level entry points a front-end binds."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import _lib
from ._lib import TfscStats, TfscTensor, check, lib


class Server:
    def __init__(self, config: dict):
        self._h = lib.tfsc_server_create(json.dumps(config).encode())
        if not self._h:
            code = _lib.E_NO_DEVICE if b"CUDA" in lib.tfsc_last_error() else _lib.E_INVALID
            raise _lib.TfscError(code, "server_create")
        self.config = config

    def close(self):
        if getattr(self, "_h", None):
            lib.tfsc_server_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def num_nodes(self) -> int:
        return lib.tfsc_server_num_nodes(self._h)

    def set_members(self, members: list[str]):
        arr = (C.c_char_p * len(members))(*[m.encode() for m in members])
        check(lib.tfsc_server_set_members(self._h, arr, len(members)), "set_members")

    def route(self, model_name: str, version: str):
        nodes = (C.c_int * 64)()
        picked = C.c_int()
        n = check(lib.tfsc_route(self._h, model_name.encode(), version.encode(), nodes, 64, C.byref(picked)), "route")
        return list(nodes[:n]), picked.value

    def ensure(self, node: int, model_name: str, version: int) -> int:
        return check(lib.tfsc_model_ensure(self._h, node, model_name.encode(), version), "model_ensure")

    def ensure_async(self, node: int, model_name: str, version: int) -> int:
        """fetchModel without waiting for the page-in (launches wait for it on-device)."""
        return check(lib.tfsc_model_ensure_async(self._h, node, model_name.encode(), version), "model_ensure_async")

    def status(self, node: int, model_name: str, version: int) -> int:
        return lib.tfsc_model_status(self._h, node, model_name.encode(), version)

    def _lines(self, fn, node):
        cap = 1 << 20
        buf = C.create_string_buffer(cap)
        check(fn(self._h, node, buf, cap), "list")
        return [l.split("\t") for l in buf.value.decode().splitlines()]

    def resident(self, node: int):
        return [(n, int(v), int(b), int(s)) for n, v, b, s in self._lines(lib.tfsc_resident_list, node)]

    def host_models(self, node: int):
        return [(n, int(v), int(b)) for n, v, b, _p in self._lines(lib.tfsc_host_list, node)]

    def predict(self, model_name: str, version: str, x: np.ndarray, out_capacity_elems: int | None = None,
                input_name: str | None = None) -> np.ndarray:
        is_int = np.issubdtype(np.asarray(x).dtype, np.integer)   # token-id inputs (BERT bundles) travel as DT_INT32
        x = np.ascontiguousarray(x, dtype=np.int32 if is_int else np.float32)
        tin = TfscTensor()
        tin.name = input_name.encode() if input_name else None
        tin.dtype = _lib.DT_INT32 if is_int else _lib.DT_FLOAT
        tin.rank = x.ndim
        for i, d in enumerate(x.shape):
            tin.shape[i] = d
        tin.data = x.ctypes.data
        tin.nbytes = x.nbytes
        cap = out_capacity_elems if out_capacity_elems is not None else max(x.size, 1) * 4 + 65536
        y = np.empty(cap, dtype=np.float32)
        tout = TfscTensor()
        tout.data = y.ctypes.data
        tout.nbytes = y.nbytes
        check(lib.tfsc_predict(self._h, model_name.encode(), version.encode(), C.byref(tin), 1, C.byref(tout), 1),
              "predict")
        shape = tuple(tout.shape[i] for i in range(tout.rank))
        n = int(np.prod(shape)) if shape else 1
        return y[:n].reshape(shape).copy()

    def _tensors(self, x, out_capacity_elems, input_name):
        is_int = np.issubdtype(np.asarray(x).dtype, np.integer)
        x = np.ascontiguousarray(x, dtype=np.int32 if is_int else np.float32)
        tin = TfscTensor()
        tin.name = input_name.encode() if input_name else None
        tin.dtype = _lib.DT_INT32 if is_int else _lib.DT_FLOAT
        tin.rank = x.ndim
        for i, d in enumerate(x.shape):
            tin.shape[i] = d
        tin.data = x.ctypes.data
        tin.nbytes = x.nbytes
        cap = out_capacity_elems if out_capacity_elems is not None else max(x.size, 1) * 4 + 65536
        y = np.empty(cap, dtype=np.float32)
        tout = TfscTensor()
        tout.data = y.ctypes.data
        tout.nbytes = y.nbytes
        return x, tin, y, tout

    @staticmethod
    def _result(y, tout):
        shape = tuple(tout.shape[i] for i in range(tout.rank))
        n = int(np.prod(shape)) if shape else 1
        return y[:n].reshape(shape).copy()

    def predict_deadline(self, model_name: str, version: str, x: np.ndarray, deadline_ns: int, **kw) -> np.ndarray:
        """tfsc_predict_deadline: deadline is absolute on the clock of now_ns() (0 = none)."""
        _x, tin, y, tout = self._tensors(x, kw.get("out_capacity_elems"), kw.get("input_name"))
        check(lib.tfsc_predict_deadline(self._h, model_name.encode(), version.encode(), C.byref(tin), 1, C.byref(tout), 1,
                                        int(deadline_ns)), "predict")
        return self._result(y, tout)

    def predict_member(self, member: int, model_name: str, version: str, x: np.ndarray, deadline_ns: int = 0, **kw) -> np.ndarray:
        """tfsc_predict_member: the cache tier of member `member` (index into gpu.members), no ring lookup."""
        _x, tin, y, tout = self._tensors(x, kw.get("out_capacity_elems"), kw.get("input_name"))
        check(lib.tfsc_predict_member(self._h, member, model_name.encode(), version.encode(), C.byref(tin), 1, C.byref(tout), 1,
                                      int(deadline_ns)), "predict_member")
        return self._result(y, tout)

    @staticmethod
    def now_ns() -> int:
        return lib.tfsc_now_ns()

    def predict_submit(self, model_name: str, version: str, x: np.ndarray, deadline_ns: int = 0, **kw) -> "Ticket":
        """Asynchronous Predict (tfsc_predict_submit): returns a Ticket; .wait() yields the result."""
        xk, tin, y, tout = self._tensors(x, kw.get("out_capacity_elems"), kw.get("input_name"))
        t = C.c_void_p()
        check(lib.tfsc_predict_submit(self._h, model_name.encode(), version.encode(), C.byref(tin), 1, C.byref(tout), 1,
                                      int(deadline_ns), C.byref(t)), "predict_submit")
        return Ticket(t, y, tout)

    def fwd_window(self):
        """(rank, device pointer, bytes, slot bytes) of this rank's forward window (needs cluster.endpoints)."""
        p, n, sb = C.c_void_p(), C.c_size_t(), C.c_size_t()
        rank = check(lib.tfsc_fwd_window(self._h, C.byref(p), C.byref(n), C.byref(sb)), "fwd_window")
        return rank, p.value, n.value, sb.value

    def fwd_peer_window(self, peer_rank: int):
        """(device pointer, bytes) of another rank's window mapped into this process (CUDA IPC)."""
        p, n = C.c_void_p(), C.c_size_t()
        check(lib.tfsc_fwd_peer_window(self._h, peer_rank, C.byref(p), C.byref(n)), "fwd_peer_window")
        return p.value, n.value

    def grpc_predict(self, request_bytes: bytes) -> bytes:
        resp = C.c_void_p()
        n = C.c_size_t()
        check(lib.tfsc_grpc_predict(self._h, request_bytes, len(request_bytes), C.byref(resp), C.byref(n)), "grpc_predict")
        try:
            return C.string_at(resp, n.value)
        finally:
            lib.tfsc_free(resp)

    def _wire_call(self, fn, request_bytes: bytes, where: str) -> bytes:
        resp = C.c_void_p()
        n = C.c_size_t()
        check(fn(self._h, request_bytes, len(request_bytes), C.byref(resp), C.byref(n)), where)
        try:
            return C.string_at(resp, n.value)
        finally:
            lib.tfsc_free(resp)

    def grpc_classify(self, request_bytes: bytes) -> bytes:
        return self._wire_call(lib.tfsc_grpc_classify, request_bytes, "grpc_classify")

    def grpc_regress(self, request_bytes: bytes) -> bytes:
        return self._wire_call(lib.tfsc_grpc_regress, request_bytes, "grpc_regress")

    def grpc_session_run(self, request_bytes: bytes) -> bytes:
        return self._wire_call(lib.tfsc_grpc_session_run, request_bytes, "grpc_session_run")

    def rest_handle(self, method: str, url: str, body: bytes = b""):
        st = C.c_int()
        resp = C.c_void_p()
        n = C.c_size_t()
        check(lib.tfsc_rest_handle(self._h, method.encode(), url.encode(), body, len(body), C.byref(st), C.byref(resp),
                                   C.byref(n)), "rest_handle")
        try:
            return st.value, C.string_at(resp, n.value)
        finally:
            lib.tfsc_free(resp)

    def predict_device(self, node: int, model_name: str, version: int, x_ptr: int, rows: int, y_ptr: int, stream: int = 0):
        check(lib.tfsc_predict_device(self._h, node, model_name.encode(), version, x_ptr, rows, y_ptr, stream),
              "predict_device")

    def set_max_resident(self, node: int, n: int):
        check(lib.tfsc_node_set_max_resident(self._h, node, n), "set_max_resident")

    def sync(self, node: int = 0):
        check(lib.tfsc_node_sync(self._h, node), "node_sync")

    def stats(self, node: int = -1) -> dict:
        st = TfscStats()
        check(lib.tfsc_get_stats(self._h, node, C.byref(st)), "get_stats")
        return st.as_dict()


class Ticket:
    """An in-flight asynchronous Predict (tfsc_ticket). Keeps the output buffer alive until released."""

    def __init__(self, handle, y, tout):
        self._t, self._y, self._tout = handle, y, tout

    def wait(self, timeout_s: float | None = None) -> np.ndarray:
        check(lib.tfsc_predict_wait(self._t, -1 if timeout_s is None else int(timeout_s * 1e9)), "predict_wait")
        return Server._result(self._y, self._tout)

    def release(self):
        if self._t:
            lib.tfsc_predict_release(self._t)
            self._t = None

    __del__ = release
