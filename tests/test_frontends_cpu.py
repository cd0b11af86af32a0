"""Front-end plumbing that needs no GPU: the gRPC GetModelMetadata handler and the REST handler of serve.py over real
sockets with a stub in place of the native server (the -m gpu twin is tests/test_gpu_frontends.py)."""
import json
import threading
import urllib.request

import pytest

from tfservingcache_b200 import serve, tfs_wire

META = {"model_spec": {"name": "m5", "signature_name": "", "version": "1"}, "metadata": {"signature_def": {"signature_def": {"serving_default": {
    "inputs": {"x": {"dtype": "DT_FLOAT", "tensor_shape": {"dim": [{"size": "-1", "name": ""}, {"size": "64", "name": ""}], "unknown_rank": False},
                     "name": "x:0"}},
    "outputs": {"y": {"dtype": "DT_FLOAT", "tensor_shape": {"dim": [{"size": "-1", "name": ""}, {"size": "8", "name": ""}], "unknown_rank": False},
                      "name": "y:0"}},
    "method_name": "tensorflow/serving/predict"}}}}}


class StubServer:
    num_nodes = 1

    def __init__(self):
        self.calls = []

    def rest_handle(self, method, url, body=b""):
        self.calls.append((method, url, body))
        if "nope" in url:
            return 404, b'{ "error": "No matching model found" }'
        return 200, json.dumps(META).encode()


def test_grpc_get_model_metadata_handler():
    import grpc
    stub = StubServer()
    g = serve.make_grpc_server(stub, 0, host="127.0.0.1")
    g.start()
    try:
        ch = grpc.insecure_channel(f"127.0.0.1:{g.bound_port}")
        meta = ch.unary_unary("/tensorflow.serving.PredictionService/GetModelMetadata", request_serializer=lambda b: b,
                              response_deserializer=lambda b: b)
        want = tfs_wire.encode_get_model_metadata_response("m5", 1, {"serving_default": {
            "inputs": {"x": ("x:0", 1, [-1, 64])}, "outputs": {"y": ("y:0", 1, [-1, 8])}, "method_name": "tensorflow/serving/predict"}})
        assert meta(tfs_wire.encode_get_model_metadata_request("m5", 1)) == want
        assert stub.calls[-1][:2] == ("GET", "/v1/models/m5/versions/1/metadata")
        meta(tfs_wire.encode_get_model_metadata_request("m5", None))
        assert stub.calls[-1][1] == "/v1/models/m5/versions/0/metadata"   # clientForSpec: missing version -> "0"
        for req, code in ((tfs_wire.encode_get_model_metadata_request("nope", 1), grpc.StatusCode.NOT_FOUND),
                          (tfs_wire.encode_get_model_metadata_request("m5", 1, fields=("bogus",)), grpc.StatusCode.INVALID_ARGUMENT),
                          (tfs_wire.encode_get_model_metadata_request("m5", 1, fields=()), grpc.StatusCode.INVALID_ARGUMENT)):
            with pytest.raises(grpc.RpcError) as e:
                meta(req)
            assert e.value.code() == code
        ch.close()
    finally:
        g.stop(0)


def test_rest_handler_passes_method_path_and_body_through():
    stub = StubServer()
    rest = serve.make_rest_server(stub, 0, "127.0.0.1")
    th = threading.Thread(target=rest.serve_forever, daemon=True)
    th.start()
    try:
        base = f"http://127.0.0.1:{rest.server_port}"
        with urllib.request.urlopen(urllib.request.Request(base + "/v1/models/m5/versions/1:predict", data=b'{"instances": [[1]]}',
                                                           method="POST"), timeout=10) as r:
            assert r.status == 200 and json.loads(r.read()) == META and r.headers["Content-Type"] == "application/json"
        assert stub.calls[-1] == ("POST", "/v1/models/m5/versions/1:predict", b'{"instances": [[1]]}')
    finally:
        rest.shutdown()
