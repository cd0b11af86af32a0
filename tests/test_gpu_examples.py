"""-m gpu: Classify / Regress / SessionRun (tfservingproxy.go:173-198,233-244) against request / response bytes serialized
from the reference's own protobuf schema (tests/golden/examples_golden.json), on TF-Serving's half_plus_two model imported
from the independently written SavedModel fixture (tests/golden/savedmodel_golden.json): [1,2,5] -> [2.5,3,4.5]."""
import base64
import json
import os

import numpy as np
import pytest

import tfservingcache_b200 as t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def srv(tmp_path_factory):
    import torch
    assert torch.cuda.is_available()
    from conftest import load_golden
    base = tmp_path_factory.mktemp("repo")
    g = load_golden("savedmodel_golden.json")["models"]["half_plus_two"]
    for rel, b64 in g["files"].items():
        p = os.path.join(base, "half_plus_two", "00000123", rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        open(p, "wb").write(base64.b64decode(b64))
    # a native bundle with hand-declared signatures: 3 scores per example (classify) and a model that cannot regress
    rng = np.random.default_rng(0)
    w, b = rng.standard_normal((4, 3)).astype(np.float32), rng.standard_normal(3).astype(np.float32)
    t.modelformat.write_mlp_bundle(os.path.join(base, "scorer", "1"), [w], [b], ["linear"], "features", "scores",
                                   [{"name": "classify", "method": "classify", "feature": "features"},
                                    {"name": "regress", "method": "regress", "feature": "features"}])
    cfg = {"modelProvider.type": "diskProvider", "modelProvider.diskProvider.baseDir": str(base), "gpu.devices": [0],
           "gpu.arenaBytes": 8 << 20, "modelCache.size": 1 << 26, "serving.maxConcurrentModels": 4}
    s = t.Server(cfg)
    s._scorer = (w, b)
    yield s
    s.close()


def _g(golden, key, which):
    return base64.b64decode(golden("examples_golden.json")[key][which])


def test_regress_and_classify_match_reference_schema_bytes(srv, golden):
    assert srv.grpc_regress(_g(golden, "regress", "request_b64")) == _g(golden, "regress", "response_b64")
    assert srv.grpc_classify(_g(golden, "classify", "request_b64")) == _g(golden, "classify", "response_b64")
    # ExampleListWithContext: the context's features belong to every example
    assert srv.grpc_regress(_g(golden, "regress_with_context", "request_b64")) == _g(golden, "regress_with_context", "response_b64")


def test_signature_errors_are_tf_servings(srv, golden):
    with pytest.raises(t._lib.TfscError) as e:     # serving_default of half_plus_two is a predict signature
        srv.grpc_classify(_g(golden, "classify_on_predict_signature", "request_b64"))
    assert e.value.code == t._lib.E_INVALID and "tensorflow/serving/classify. Was: tensorflow/serving/predict" in str(e.value)
    with pytest.raises(t._lib.TfscError) as e:     # a classify signature asked to regress
        srv.grpc_regress(_g(golden, "classify", "request_b64"))
    assert e.value.code == t._lib.E_INVALID and "Was: tensorflow/serving/classify" in str(e.value)
    st, body = srv.rest_handle("POST", "/v1/models/scorer/versions/1:regress",
                               json.dumps({"signature_name": "regress", "examples": [{"features": [1, 2, 3, 4]}]}).encode())
    assert st == 400 and b"[batch_size, 1]" in body   # 3 outputs per example cannot be a regression


def test_rest_classify_regress(srv):
    st, body = srv.rest_handle("POST", "/v1/models/half_plus_two/versions/123:regress",
                               json.dumps({"signature_name": "regress_x_to_y", "examples": [{"x": 1.0}, {"x": [2.0]}, {"x": 5}]}).encode())
    assert st == 200 and json.loads(body) == {"results": [2.5, 3.0, 4.5]}
    w, b = srv._scorer
    xs = np.random.default_rng(1).standard_normal((2, 4)).astype(np.float32)
    st, body = srv.rest_handle("POST", "/v1/models/scorer/versions/1:classify",
                               json.dumps({"signature_name": "classify", "examples": [{"features": r.tolist()} for r in xs]}).encode())
    assert st == 200
    res = json.loads(body)["results"]
    got = np.array([[c[1] for c in ex] for ex in res])
    assert all(c[0] == "" for ex in res for c in ex) and np.max(np.abs(got - (xs.astype(np.float64) @ w + b))) <= 1e-4
    st, body = srv.rest_handle("POST", "/v1/models/half_plus_two/versions/123:classify", b'{"examples": [{"y": 1.0}]}')
    assert st == 400


def test_session_run_matches_reference_schema_bytes(srv, golden):
    assert srv.grpc_session_run(_g(golden, "session_run", "request_b64")) == _g(golden, "session_run", "response_b64")
    from oracle import wire
    bad = _g(golden, "session_run", "request_b64").replace(b"y:0", b"z:0")
    with pytest.raises(t._lib.TfscError) as e:
        srv.grpc_session_run(bad)
    assert e.value.code == t._lib.E_INVALID
    assert wire is not None
