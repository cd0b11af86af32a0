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
