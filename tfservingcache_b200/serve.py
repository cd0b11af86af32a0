"""Network front-ends of the cache tier (cmd/taskhandler/main.go:45-64): TF-Serving-compatible REST on
`cacheRestPort` and gRPC on `cacheGrpcPort`, both thin shims over the C ABI -- the request bytes go straight to
tfsc_rest_handle / tfsc_grpc_predict (libtfsc_b200.so does parsing, routing, residency, batching, execution).

    python -m tfservingcache_b200.serve [config.yaml]

gRPC is served with generic handlers and identity (de)serializers, so no generated stubs are needed and the
PredictRequest is never re-marshalled (the reference unmarshals + re-marshals it at each of its two tiers).
"""
from __future__ import annotations

import json
import sys
import threading
from concurrent import futures
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

from . import _lib, tfs_wire
from .config import load_config
from .server import Server

_GRPC_CODE = {_lib.E_INVALID: "INVALID_ARGUMENT", _lib.E_TIMEOUT: "DEADLINE_EXCEEDED", _lib.E_NOT_FOUND: "NOT_FOUND",
              _lib.E_EXHAUSTED: "RESOURCE_EXHAUSTED", _lib.E_UNIMPLEMENTED: "UNIMPLEMENTED", _lib.E_INTERNAL: "INTERNAL",
              _lib.E_NO_DEVICE: "UNAVAILABLE", _lib.E_EMPTY_RING: "UNAVAILABLE", _lib.E_BUFFER: "INTERNAL"}


def prometheus_text(srv: Server) -> bytes:
    """Exposition of the reference's metric names (cachemanager.go:24-43, tfservingproxy.go:25-32; labels collapse
    to all_models/-1 as with metrics.modelLabels=false) plus the HBM-side additions of this build."""
    st = srv.stats()
    lab = '{model="all_models",version="-1"}'
    lines = []

    def add(name, kind, help_, samples):
        lines.append(f"# HELP {name} {help_}")
        lines.append(f"# TYPE {name} {kind}")
        lines.extend(samples)

    add("tfservingcache_cache_total", "counter", "The total number of cache misses and hits", [f"tfservingcache_cache_total{lab} {st['cache_total']}"])
    add("tfservingcache_cache_hits_total", "counter", "The total number of cache hits", [f"tfservingcache_cache_hits_total{lab} {st['cache_hits_total']}"])
    add("tfservingcache_cache_misses_total", "counter", "The total number of cache misses", [f"tfservingcache_cache_misses_total{lab} {st['cache_misses_total']}"])
    add("tfservingcache_cache_duration_seconds", "summary", "The duration of cache requests, including hits and misses",
        [f"tfservingcache_cache_duration_seconds_sum{lab} {st['cache_duration_seconds_sum']}", f"tfservingcache_cache_duration_seconds_count{lab} {st['cache_total']}"])
    add("tfservingcache_cache_fetch_duration_seconds", "summary", "The duration of cache fetches (when cache miss)",
        [f"tfservingcache_cache_fetch_duration_seconds_sum{lab} {st['cache_fetch_duration_seconds_sum']}", f"tfservingcache_cache_fetch_duration_seconds_count{lab} {st['cache_misses_total']}"])
    add("tfservingcache_proxy_requests_total", "counter", "The total number of requests",
        [f'tfservingcache_proxy_requests_total{{protocol="rest"}} {st["proxy_requests_rest"]}', f'tfservingcache_proxy_requests_total{{protocol="grpc"}} {st["proxy_requests_grpc"]}'])
    add("tfservingcache_proxy_failures_total", "counter", "The total number of failed requests",
        [f'tfservingcache_proxy_failures_total{{protocol="rest"}} {st["proxy_failures_rest"]}', f'tfservingcache_proxy_failures_total{{protocol="grpc"}} {st["proxy_failures_grpc"]}'])
    ratio = st["cache_hits_total"] / st["cache_total"] if st["cache_total"] else 0.0
    add("tfservingcache_hbm_cache_hit_ratio", "gauge", "cache_hits_total / cache_total (BASELINE metric 'HBM cache hit %')", [f"tfservingcache_hbm_cache_hit_ratio {ratio}"])
    for k in ("arena_bytes_used", "arena_bytes_capacity", "resident_models", "host_models"):
        add(f"tfservingcache_{k}", "gauge", k.replace("_", " "), [f"tfservingcache_{k} {st[k]}"])
    for k in ("h2d_weight_bytes", "h2d_input_bytes", "d2h_output_bytes", "evictions_hbm", "evictions_host", "batches", "batched_rows", "kernel_launches"):
        add(f"tfservingcache_{k}_total", "counter", k.replace("_", " "), [f"tfservingcache_{k}_total {st[k]}"])
    return ("\n".join(lines) + "\n").encode()


def make_rest_server(srv: Server, port: int, host: str = "0.0.0.0", metrics_path: str = "/monitoring/prometheus/metrics") -> ThreadingHTTPServer:
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def _serve(self):
            n = int(self.headers.get("Content-Length") or 0)
            body = self.rfile.read(n) if n else b""
            if self.command == "GET" and self.path.split("?")[0] == metrics_path:  # metrics.path (MetricsHandler, metrics.go:16)
                out = prometheus_text(srv)
                self.send_response(200)
                self.send_header("Content-Type", "text/plain; version=0.0.4")
                self.send_header("Content-Length", str(len(out)))
                self.end_headers()
                self.wfile.write(out)
                return
            status, out = srv.rest_handle(self.command, self.path, body)   # RestProxy.Serve, tfservingproxy.go:93-129
            self.send_response(status)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(out)))
            self.end_headers()
            self.wfile.write(out)

        do_GET = do_POST = _serve

        def log_message(self, *a):
            pass

    httpd = ThreadingHTTPServer((host, port), Handler)
    httpd.daemon_threads = True
    return httpd


def make_grpc_server(srv: Server, port: int, max_msg: int = 16 * 1024 * 1024, workers: int = 64, host: str = "0.0.0.0"):
    import grpc

    ident = lambda b: b  # noqa: E731

    def predict(request: bytes, context):
        try:
            return srv.grpc_predict(request)       # proxyServiceServer.Predict, tfservingproxy.go:201-212
        except _lib.TfscError as e:
            context.abort(getattr(grpc.StatusCode, _GRPC_CODE.get(e.code, "INTERNAL")), str(e))

    def unsupported(name):
        def fn(request: bytes, context):
            # MultiInference is rejected by the reference too (tfservingproxy.go:215-217)
            context.abort(grpc.StatusCode.UNIMPLEMENTED, f"{name} not supported")
        return fn

    # ---- TF-Serving facade: tensorflow.serving.ModelService, the two RPCs the reference's TFServingController
    # issues (servingcontroller.go:88-138). With it the UNMODIFIED reference can use this server as its
    # serving.grpcHost / restHost instead of a tensorflow/serving container.
    def get_model_status(request: bytes, context):
        try:
            name, version = tfs_wire.decode_get_model_status_request(request)
        except Exception as e:  # malformed request
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(e))
        found = []
        for node in range(srv.num_nodes):
            if version is not None:
                st = srv.status(node, name, version)
                if st >= 0:
                    found.append((version, st, 0, ""))
            else:
                known = {v for n, v, _b in srv.host_models(node) if n == name}
                for v in sorted(known):
                    st = srv.status(node, name, v)
                    if st >= 0:
                        found.append((v, st, 0, ""))
        if not found:
            # TF-Serving answers NOT_FOUND; the reference's health probe depends on exactly that code (cachemanager.go:80-84)
            context.abort(grpc.StatusCode.NOT_FOUND, f"Could not find any versions of model {name}")
        best = {}
        for v, st, c, m in found:   # a model spread over several GPUs: report its most advanced replica
            if v not in best or st == _lib.STATE_AVAILABLE:
                best[v] = (v, st, c, m)
        return tfs_wire.encode_get_model_status_response([best[v] for v in sorted(best)])

    def handle_reload_config(request: bytes, context):
        try:
            models = tfs_wire.decode_reload_config_request(request)
        except Exception as e:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(e))
        # the reference lists the models to keep loaded MRU-first (cachemanager.go:167-170): touch them in reverse so
        # the first ends up most recently used; the HBM tier is an LRU, unlisted models simply age out
        try:
            for name, _base, _platform, versions in reversed(models):
                for v in reversed(versions):
                    nodes, picked = srv.route(name, str(v))
                    if nodes[picked] >= 0:
                        srv.ensure(nodes[picked], name, v)
        except _lib.TfscError as e:
            return tfs_wire.encode_reload_config_response(-e.code if -17 < e.code < 0 else 13, str(e))
        return tfs_wire.encode_reload_config_response()

    def get_model_metadata(request: bytes, context):
        # PredictionService.GetModelMetadata (tfservingproxy.go:220-231): same fetchModel path as Predict, answered
        # from the model manifest; the only metadata field TF-Serving knows is "signature_def"
        try:
            name, version, fields = tfs_wire.decode_get_model_metadata_request(request)
        except Exception as e:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(e))
        if not fields or any(f != "signature_def" for f in fields):
            context.abort(grpc.StatusCode.INVALID_ARGUMENT,
                          "Metadata field " + (next((f for f in fields if f != "signature_def"), "") or "<none>") + " is not supported"
                          if fields else "GetModelMetadataRequest must specify at least one metadata_field")
        v = 0 if version is None else version      # clientForSpec: a missing version is "0" (tfservingproxy.go:246-250)
        status, body = srv.rest_handle("GET", f"/v1/models/{name}/versions/{v}/metadata", b"")
        if status != 200:
            msg = body.decode(errors="replace")
            try:
                msg = json.loads(msg).get("error", msg)
            except ValueError:
                pass
            context.abort(grpc.StatusCode.NOT_FOUND if status == 404 else grpc.StatusCode.INTERNAL, msg)
        return tfs_wire.encode_get_model_metadata_response(name, version, tfs_wire.signatures_from_rest_metadata(json.loads(body)))

    health_status = {"serving": True}

    def health_check(request: bytes, context):
        return b"\x08\x01" if health_status["serving"] else b"\x08\x02"  # HealthCheckResponse{status}

    def wire_call(fn_name):
        # Classify / Regress / SessionRun (tfservingproxy.go:173-198,233-244): serialized request in, response out
        def fn(request: bytes, context):
            try:
                return getattr(srv, fn_name)(request)
            except _lib.TfscError as e:
                context.abort(getattr(grpc.StatusCode, _GRPC_CODE.get(e.code, "INTERNAL")), str(e))
        return fn

    methods = {"Predict": grpc.unary_unary_rpc_method_handler(predict, ident, ident),
               "GetModelMetadata": grpc.unary_unary_rpc_method_handler(get_model_metadata, ident, ident),
               "Classify": grpc.unary_unary_rpc_method_handler(wire_call("grpc_classify"), ident, ident),
               "Regress": grpc.unary_unary_rpc_method_handler(wire_call("grpc_regress"), ident, ident),
               "MultiInference": grpc.unary_unary_rpc_method_handler(unsupported("MultiInference"), ident, ident)}
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=workers),
                         options=[("grpc.max_receive_message_length", max_msg), ("grpc.max_send_message_length", max_msg)])
    server.add_generic_rpc_handlers((
        grpc.method_handlers_generic_handler("tensorflow.serving.PredictionService", methods),
        grpc.method_handlers_generic_handler("tensorflow.serving.SessionService",
                                             {"SessionRun": grpc.unary_unary_rpc_method_handler(wire_call("grpc_session_run"), ident, ident)}),
        grpc.method_handlers_generic_handler("tensorflow.serving.ModelService", {
            "GetModelStatus": grpc.unary_unary_rpc_method_handler(get_model_status, ident, ident),
            "HandleReloadConfigRequest": grpc.unary_unary_rpc_method_handler(handle_reload_config, ident, ident)}),
        grpc.method_handlers_generic_handler("grpc.health.v1.Health",
                                             {"Check": grpc.unary_unary_rpc_method_handler(health_check, ident, ident)}),
    ))
    bound = server.add_insecure_port(f"{host}:{port}")
    server.set_health = lambda ok: health_status.__setitem__("serving", bool(ok))  # GrpcProxy.SetHealth, :151-157
    server.bound_port = bound
    return server


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    cfg = load_config(argv[0] if argv else "config.yaml")
    srv = Server(cfg)
    rest = make_rest_server(srv, int(cfg.get("cacheRestPort", 8094)), metrics_path=str(cfg.get("metrics.path", "/monitoring/prometheus/metrics")))
    max_msg = int(cfg.get("serving.grpcMaxMsgSize") or 16 * 1024 * 1024)
    grpc_srv = make_grpc_server(srv, int(cfg.get("cacheGrpcPort", 8095)), max_msg)
    grpc_srv.start()
    threading.Thread(target=rest.serve_forever, daemon=True).start()
    print(f"tfservingcache_b200: REST :{rest.server_port}  gRPC :{grpc_srv.bound_port}  nodes={srv.num_nodes}", flush=True)
    try:
        grpc_srv.wait_for_termination()
    except KeyboardInterrupt:
        pass
    finally:
        rest.shutdown()
        srv.close()


if __name__ == "__main__":
    main()
