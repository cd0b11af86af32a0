"""ModelService wire codec (tfservingcache_b200/tfs_wire.py) against bytes serialized by python-protobuf from the
reference's own embedded schema (tests/golden/modelservice_golden.json, made by tests/golden/make_golden.py)."""
import base64

from tfservingcache_b200 import tfs_wire as w


def test_modelservice_codec_matches_reference_schema(golden):
    g = golden("modelservice_golden.json")
    b = base64.b64decode
    assert w.decode_get_model_status_request(b(g["status_request"]["b64"])) == ("foo", 42)
    assert w.encode_get_model_status_request("foo", 42) == b(g["status_request"]["b64"])
    assert w.decode_get_model_status_request(b(g["probe_request"]["b64"])) == ("__TFSERVINGCACHE_PROBE_CHECK__", 1)
    assert w.decode_get_model_status_request(w.encode_get_model_status_request("x", None)) == ("x", None)
    assert w.decode_get_model_status_response(b(g["status_response"]["b64"])) == [tuple(s) for s in g["status_response"]["statuses"]]
    assert w.encode_get_model_status_response([(123, 30, 0, "")]) == b(g["status_response"]["b64"])
    models = [tuple(m[:3]) + (m[3],) for m in g["reload_request"]["models"]]
    assert w.decode_reload_config_request(b(g["reload_request"]["b64"])) == [(n, p, pl, v) for n, p, pl, v in models]
    assert w.encode_reload_config_request(models) == b(g["reload_request"]["b64"])
    assert w.encode_reload_config_response() == b(g["reload_response_ok"]["b64"])
    assert w.decode_reload_config_response(w.encode_reload_config_response(5, "nope")) == (5, "nope")


def test_get_model_metadata_codec_matches_reference_schema(golden):
    g = golden("metadata_golden.json")
    b = base64.b64decode
    r = g["request"]
    assert w.decode_get_model_metadata_request(b(r["b64"])) == (r["name"], r["version"], r["fields"])
    assert w.encode_get_model_metadata_request(r["name"], r["version"], r["fields"]) == b(r["b64"])
    r = g["request_no_version"]
    assert w.decode_get_model_metadata_request(b(r["b64"])) == (r["name"], None, r["fields"])
    for c in g["cases"]:
        (ik, idt, idims), (ok_, odt, odims) = c["input"], c["output"]
        sigs = {"serving_default": {"inputs": {ik: (ik + ":0", idt, idims)}, "outputs": {ok_: (ok_ + ":0", odt, odims)},
                                    "method_name": "tensorflow/serving/predict"}}
        assert w.encode_get_model_metadata_response(c["name"], c["version"], sigs) == b(c["b64"]), c["name"]


def test_signatures_from_rest_metadata_json():
    doc = {"model_spec": {"name": "m", "signature_name": "", "version": "1"},
           "metadata": {"signature_def": {"signature_def": {"serving_default": {
               "inputs": {"input_ids": {"dtype": "DT_INT32", "tensor_shape": {"dim": [{"size": "-1", "name": ""}, {"size": "128", "name": ""}],
                                                                              "unknown_rank": False}, "name": "input_ids:0"}},
               "outputs": {"y": {"dtype": "DT_FLOAT", "tensor_shape": {"dim": [{"size": "-1", "name": ""}], "unknown_rank": False}, "name": "y:0"}},
               "method_name": "tensorflow/serving/predict"}}}}}
    assert w.signatures_from_rest_metadata(doc) == {"serving_default": {
        "inputs": {"input_ids": ("input_ids:0", 3, [-1, 128])}, "outputs": {"y": ("y:0", 1, [-1])},
        "method_name": "tensorflow/serving/predict"}}
