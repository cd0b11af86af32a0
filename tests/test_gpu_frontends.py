"""-m gpu: the REST + gRPC front-ends over real sockets (the reference's tfservingproxy_test.go scenarios with a
real executor behind them instead of mocks)."""
import json
import threading
import urllib.error
import urllib.request

import numpy as np
import pytest

import tfservingcache_b200 as t
from oracle import models, wire
from tfservingcache_b200 import serve

pytestmark = pytest.mark.gpu
DIMS = [64, 96, 8]


@pytest.fixture(scope="module")
def endpoints():
    import torch
    assert torch.cuda.is_available()
    cfg = {"modelProvider.type": "synthetic", "modelProvider.synthetic.dims": DIMS, "modelProvider.synthetic.count": 16,
           "gpu.devices": [0], "gpu.arenaBytes": 8 << 20, "serving.maxConcurrentModels": 4, "modelCache.size": 1 << 28}
    srv = t.Server(cfg)
    rest = serve.make_rest_server(srv, 0, "127.0.0.1")
    g = serve.make_grpc_server(srv, 0, host="127.0.0.1")
    g.start()
    th = threading.Thread(target=rest.serve_forever, daemon=True)
    th.start()
    yield srv, f"http://127.0.0.1:{rest.server_port}", f"127.0.0.1:{g.bound_port}"
    rest.shutdown()
    g.stop(0)
    srv.close()


def _ref(j, x):
    man, blob = models.synth_mlp_blob(DIMS, seed=1000 + j)
    return models.forward(man, blob, np.asarray(x, np.float32), np.float64)


def _http(url, body=None):
    req = urllib.request.Request(url, data=body, method="POST" if body is not None else "GET")
    try:
        with urllib.request.urlopen(req, timeout=30) as r:
            return r.status, r.read()
    except urllib.error.HTTPError as e:
        return e.code, e.read()


def test_http_proxy_parses_url_and_predicts(endpoints):
    _srv, base, _ = endpoints
    x = np.random.default_rng(0).standard_normal((2, DIMS[0])).astype(np.float32)
    st, body = _http(f"{base}/v1/models/m3/versions/1:predict", json.dumps({"instances": x.tolist()}).encode())
    assert st == 200
    y = np.array(json.loads(body)["predictions"])
    assert y.shape == (2, DIMS[-1]) and np.max(np.abs(y - _ref(3, x))) <= 1e-4
    st, body = _http(f"{base}/v1/models/m3/versions/1")
    assert st == 200 and json.loads(body)["model_version_status"][0]["state"] == "AVAILABLE"


def test_http_proxy_invalid_url_causes_404_and_missing_version_400(endpoints):
    _srv, base, _ = endpoints
    assert _http(f"{base}/v1/thisisabadrequest/foobar/versions/42") == (404, b'{"Status":"Error","Message":"Not found"}\n')
    assert _http(f"{base}/v1/models/foobar") == (400, b'{"Status":"Error","Message":"Model version must be provided"}\n')
    assert _http(f"{base}/v1/models/unknown/versions/1:predict", b'{"instances": [[1.0]]}')[0] == 404


def test_grpc_proxy_predict_and_errors(endpoints):
    import grpc
    srv, _, target = endpoints
    ch = grpc.insecure_channel(target)
    predict = ch.unary_unary("/tensorflow.serving.PredictionService/Predict", request_serializer=lambda b: b,
                             response_deserializer=lambda b: b)
    x = np.random.default_rng(1).standard_normal((5, DIMS[0])).astype(np.float32)
    spec, outs = wire.decode_predict_response(predict(wire.encode_predict_request("m7", 1, {"x": x})))
    assert spec[:2] == ("m7", 1) and np.max(np.abs(outs["y"] - _ref(7, x))) <= 1e-4
    with pytest.raises(grpc.RpcError) as e:
        predict(wire.encode_predict_request("nope", 1, {"x": x}))
    assert e.value.code() == grpc.StatusCode.NOT_FOUND
    multi = ch.unary_unary("/tensorflow.serving.PredictionService/MultiInference", request_serializer=lambda b: b,
                           response_deserializer=lambda b: b)
    with pytest.raises(grpc.RpcError) as e:
        multi(b"")
    assert e.value.code() == grpc.StatusCode.UNIMPLEMENTED and "MultiInference not supported" in e.value.details()
    health = ch.unary_unary("/grpc.health.v1.Health/Check", request_serializer=lambda b: b, response_deserializer=lambda b: b)
    assert health(b"") == b"\x08\x01"
    st = srv.stats()
    assert st["proxy_requests_grpc"] >= 2 and st["proxy_failures_grpc"] >= 1
    ch.close()


def test_metrics_endpoint_keeps_reference_metric_names(endpoints):
    _srv, base, _ = endpoints
    st, body = _http(f"{base}/monitoring/prometheus/metrics")
    text = body.decode()
    assert st == 200
    for name in ("tfservingcache_cache_total", "tfservingcache_cache_hits_total", "tfservingcache_cache_misses_total",
                 "tfservingcache_cache_duration_seconds", "tfservingcache_cache_fetch_duration_seconds",
                 "tfservingcache_proxy_requests_total", "tfservingcache_proxy_failures_total", "tfservingcache_hbm_cache_hit_ratio"):
        assert f"# TYPE {name} " in text
    assert 'tfservingcache_cache_total{model="all_models",version="-1"}' in text


def test_tfserving_facade_model_service(endpoints, golden):
    """The two ModelService RPCs the reference's TFServingController issues (servingcontroller.go:88-138), so the
    unmodified reference can point serving.grpcHost at this server."""
    import base64
    import grpc
    from tfservingcache_b200 import tfs_wire
    srv, _, target = endpoints
    ch = grpc.insecure_channel(target)
    raw = dict(request_serializer=lambda b: b, response_deserializer=lambda b: b)
    status = ch.unary_unary("/tensorflow.serving.ModelService/GetModelStatus", **raw)
    reload_cfg = ch.unary_unary("/tensorflow.serving.ModelService/HandleReloadConfigRequest", **raw)
    # health probe of the reference: a model that does not exist must answer NOT_FOUND (code 5), cachemanager.go:76-89
    probe = base64.b64decode(golden("modelservice_golden.json")["probe_request"]["b64"])
    with pytest.raises(grpc.RpcError) as e:
        status(probe)
    assert e.value.code() == grpc.StatusCode.NOT_FOUND
    # ReloadConfig as createModelConfig builds it, then poll status like reloadServingConfig does
    req = tfs_wire.encode_reload_config_request([("m9", "/models/m9", "tensorflow", [1]), ("m10", "/models/m10", "tensorflow", [1])])
    assert tfs_wire.decode_reload_config_response(reload_cfg(req)) == (0, "")
    for name in ("m9", "m10"):
        got = tfs_wire.decode_get_model_status_response(status(tfs_wire.encode_get_model_status_request(name, 1)))
        assert got == [(1, 30, 0, "")]          # AVAILABLE
    assert [n for n, *_ in srv.resident(0)][:2] == ["m9", "m10"]   # first listed = most recently used
    # unknown model in the config -> error status in the response (TF-Serving reports it the same way)
    code, msg = tfs_wire.decode_reload_config_response(reload_cfg(tfs_wire.encode_reload_config_request([("zzz", "/models/zzz", "tensorflow", [1])])))
    assert code == 5 and "No matching model" in msg
    ch.close()


def test_grpc_get_model_metadata(endpoints):
    """PredictionService.GetModelMetadata (forwarded by tfservingproxy.go:220-231): signature_def map of the model."""
    import grpc
    from tfservingcache_b200 import tfs_wire
    _srv, _, target = endpoints
    ch = grpc.insecure_channel(target)
    meta = ch.unary_unary("/tensorflow.serving.PredictionService/GetModelMetadata", request_serializer=lambda b: b,
                          response_deserializer=lambda b: b)
    want = tfs_wire.encode_get_model_metadata_response("m5", 1, {"serving_default": {
        "inputs": {"x": ("x:0", 1, [-1, DIMS[0]])}, "outputs": {"y": ("y:0", 1, [-1, DIMS[-1]])},
        "method_name": "tensorflow/serving/predict"}})
    assert meta(tfs_wire.encode_get_model_metadata_request("m5", 1)) == want
    with pytest.raises(grpc.RpcError) as e:
        meta(tfs_wire.encode_get_model_metadata_request("nope", 1))
    assert e.value.code() == grpc.StatusCode.NOT_FOUND
    with pytest.raises(grpc.RpcError) as e:
        meta(tfs_wire.encode_get_model_metadata_request("m5", 1, fields=("bogus",)))
    assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT
    ch.close()


def _varint(v):
    out = b""
    while True:
        b7 = v & 0x7F
        v >>= 7
        out += bytes([b7 | (0x80 if v else 0)])
        if not v:
            return out


def test_request_shape_cannot_undersize_the_response(endpoints):
    """rows are derived from the element count, the response shape from the client's tensor_shape;
    a [1, 2*in_dim] request is two rows and must be rejected, not written into a one-row buffer."""
    import ctypes as C
    srv, base, _ = endpoints
    two_rows_as_one = np.random.default_rng(5).standard_normal((1, 2 * DIMS[0])).astype(np.float32)
    st, body = _http(f"{base}/v1/models/m3/versions/1:predict", json.dumps({"instances": two_rows_as_one.tolist()}).encode())
    assert st == 400 and b"does not match the model signature" in body
    with pytest.raises(t._lib.TfscError) as e:
        srv.grpc_predict(wire.encode_predict_request("m3", 1, {"x": two_rows_as_one}))
    assert e.value.code == t._lib.E_INVALID
    # C ABI with an output buffer sized from the (wrong) shape the client claims: must fail, not overflow
    with pytest.raises(t._lib.TfscError) as e:
        srv.predict("m3", "1", two_rows_as_one, out_capacity_elems=DIMS[-1])
    assert e.value.code == t._lib.E_INVALID
    # a flat vector of 2*in_dim elements is two rows: accepted, shape [2, out]
    flat = two_rows_as_one.reshape(-1)
    y = srv.predict("m3", "1", flat)
    assert y.shape == (2, DIMS[-1]) and np.max(np.abs(y - _ref(3, flat.reshape(2, -1)))) <= 1e-4
    # more than one input tensor is an error (the templates have one input), not silently ignored
    tin = (t._lib.TfscTensor * 2)()
    x = np.zeros((1, DIMS[0]), np.float32)
    for i in range(2):
        tin[i].dtype, tin[i].rank, tin[i].data, tin[i].nbytes = t._lib.DT_FLOAT, 2, x.ctypes.data, x.nbytes
        tin[i].shape[0], tin[i].shape[1] = 1, DIMS[0]
    yb = np.empty(DIMS[-1], np.float32)
    tout = t._lib.TfscTensor()
    tout.data, tout.nbytes = yb.ctypes.data, yb.nbytes
    assert t._lib.lib.tfsc_predict(srv._h, b"m3", b"1", tin, 2, C.byref(tout), 1) == t._lib.E_INVALID


def test_scalar_broadcast_is_bounded(endpoints):
    """a ~30-byte request whose tensor_shape says 2^33 elements must be refused before anything is
    allocated (and nothing may be thrown through the C ABI)."""
    srv, _, _ = endpoints

    def req(dim):
        shape = b"\x12" + _varint(len(b"\x08" + _varint(dim))) + b"\x08" + _varint(dim)     # TensorShapeProto{dim{size}}
        tensor = b"\x08\x01" + b"\x12" + _varint(len(shape)) + shape + b"\x2a\x04" + np.float32(1.5).tobytes()
        entry = b"\x0a\x01x" + b"\x12" + _varint(len(tensor)) + tensor
        spec = b"\x0a\x02m3" + b"\x12\x02\x08\x01"
        return b"\x0a" + _varint(len(spec)) + spec + b"\x12" + _varint(len(entry)) + entry

    for dim in (1 << 33, 1 << 62, (1 << 24) + 1):
        with pytest.raises(t._lib.TfscError) as e:
            srv.grpc_predict(req(dim))
        assert e.value.code == t._lib.E_INVALID
    # the legitimate use still works: one value fills a [2, in] tensor
    ok = wire.encode_predict_request("m3", 1, {"x": np.full((2, DIMS[0]), 1.5, np.float32)})
    _spec, outs = wire.decode_predict_response(srv.grpc_predict(ok))
    want = outs["y"]
    small = req(DIMS[0])   # shape [in], one float_val -> broadcast
    _spec, outs = wire.decode_predict_response(srv.grpc_predict(small))
    assert np.max(np.abs(outs["y"].reshape(-1) - want[0])) <= 1e-6
