"""Generates the committed golden fixtures.  Run in the BUILD container only (needs
/root/reference); the fixtures travel, the reference does not.

  wire_golden.json   PredictRequest / PredictResponse bytes serialized by python-protobuf from the
                     reference's OWN schema: the gzipped FileDescriptorProtos embedded in
                     proto/tensorflow/**/*.pb.go are extracted and loaded into a DescriptorPool.
  ring_golden.json   Placements computed by the oracle restatement (oracle/ring.py) -- NOT by the
                     Go module (absent): they pin C++ == Python restatement, not Go parity.
  trace_golden.json  Residency-machine outcome traces from oracle/cachemanager.py.
"""
import base64
import gzip
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/proto/tensorflow"


def embedded_descriptors(path):
    src = open(path).read()
    out = []
    for m in re.finditer(r"var fileDescriptor_\w+ = \[\]byte\{(.*?)\n\}", src, re.S):
        raw = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1)))
        out.append(gzip.decompress(raw))
    return out


def build_pool():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from google.protobuf import wrappers_pb2, any_pb2  # noqa: F401  (well-known deps)
    pool = descriptor_pool.Default()
    fds = {}
    for sub in ("core/framework", "core/lib/core", "core/protobuf", "serving"):
        d = os.path.join(REF, sub)
        for fn in sorted(os.listdir(d)):
            if fn.endswith(".pb.go"):
                for blob in embedded_descriptors(os.path.join(d, fn)):
                    fd = descriptor_pb2.FileDescriptorProto.FromString(blob)
                    fds[fd.name] = fd
    done = set()

    def add(name):
        if name in done or name not in fds:
            return
        for dep in fds[name].dependency:
            add(dep)
        try:
            pool.Add(fds[name])
        except Exception:
            pass
        done.add(name)

    for want in ("tensorflow_serving/apis/predict.proto", "tensorflow_serving/apis/get_model_status.proto",
                 "tensorflow_serving/apis/model_management.proto", "tensorflow_serving/apis/get_model_metadata.proto"):
        add(want)
    get = lambda full: message_factory.GetMessageClass(pool.FindMessageTypeByName(full))
    build_pool.get = get
    return get("tensorflow.serving.PredictRequest"), get("tensorflow.serving.PredictResponse")


def wire_golden():
    PredictRequest, PredictResponse = build_pool()
    rng = np.random.default_rng(7)
    cases = []

    def add_case(name, version, sig, arr, mode, key="x"):
        req = PredictRequest()
        req.model_spec.name = name
        if version is not None:
            req.model_spec.version.value = version
        if sig:
            req.model_spec.signature_name = sig
        t = req.inputs[key]
        t.dtype = 1
        for d in arr.shape:
            t.tensor_shape.dim.add().size = d
        if mode == "content":
            t.tensor_content = arr.astype("<f4").tobytes()
        else:
            t.float_val.extend(float(v) for v in arr.ravel())
        cases.append({"name": name, "version": version, "signature": sig, "key": key, "mode": mode,
                      "shape": list(arr.shape), "values_b64": base64.b64encode(arr.astype("<f4").tobytes()).decode(),
                      "request_b64": base64.b64encode(req.SerializeToString(deterministic=True)).decode()})

    add_case("foobar", 42, "", rng.standard_normal((2, 8)).astype(np.float32), "content")
    add_case("half_plus_two", 123, "serving_default", np.array([1.0, 2.0, 5.0], np.float32), "float_val")
    add_case("m0001", None, "", rng.standard_normal((1, 16)).astype(np.float32), "content")
    add_case("m0002", 1, "", rng.standard_normal((3, 4)).astype(np.float32), "float_val", key="inputs")
    add_case("big", 2 ** 40 + 5, "", rng.standard_normal((4, 300)).astype(np.float32), "content")
    # a response as the reference schema serializes it (float_val, TF-Serving's default AsProtoField)
    resp = PredictResponse()
    y = np.array([2.5, 3.0, 4.5], np.float32)
    t = resp.outputs["y"]
    t.dtype = 1
    for d in y.shape:
        t.tensor_shape.dim.add().size = d
    t.float_val.extend(float(v) for v in y.ravel())
    resp.model_spec.name = "half_plus_two"
    resp.model_spec.version.value = 123
    resp.model_spec.signature_name = "serving_default"
    return {"requests": cases,
            "response": {"name": "half_plus_two", "version": 123, "signature": "serving_default", "key": "y",
                         "shape": [3], "values": [2.5, 3.0, 4.5],
                         "response_b64": base64.b64encode(resp.SerializeToString(deterministic=True)).decode()}}


def modelservice_golden():
    """GetModelStatus / ReloadConfig messages as the reference's TFServingController builds them
    (servingcontroller.go:88-138,159-187), serialized by python-protobuf from the reference's embedded schema."""
    build_pool()
    get = build_pool.get
    Req, Resp = get("tensorflow.serving.GetModelStatusRequest"), get("tensorflow.serving.GetModelStatusResponse")
    RReq, RResp = get("tensorflow.serving.ReloadConfigRequest"), get("tensorflow.serving.ReloadConfigResponse")
    out = {}
    r = Req(); r.model_spec.name = "foo"; r.model_spec.version.value = 42
    out["status_request"] = {"name": "foo", "version": 42, "b64": base64.b64encode(r.SerializeToString(deterministic=True)).decode()}
    r2 = Req(); r2.model_spec.name = "__TFSERVINGCACHE_PROBE_CHECK__"; r2.model_spec.version.value = 1
    out["probe_request"] = {"name": "__TFSERVINGCACHE_PROBE_CHECK__", "version": 1, "b64": base64.b64encode(r2.SerializeToString(deterministic=True)).decode()}
    resp = Resp(); m = resp.model_version_status.add(); m.version = 123; m.state = 30; m.status.error_code = 0; m.status.error_message = ""
    m.status.SetInParent()
    out["status_response"] = {"statuses": [[123, 30, 0, ""]], "b64": base64.b64encode(resp.SerializeToString(deterministic=True)).decode()}
    rr = RReq()
    models = [("a", "/models/a", "tensorflow", [1, 2]), ("b", "/models/b", "tensorflow", [7])]
    for name, base, plat, vers in models:
        c = rr.config.model_config_list.config.add(); c.name = name; c.base_path = base; c.model_platform = plat
        c.model_version_policy.specific.versions.extend(vers)
    out["reload_request"] = {"models": [[n, b, p, v] for n, b, p, v in models], "b64": base64.b64encode(rr.SerializeToString(deterministic=True)).decode()}
    ok = RResp(); ok.status.SetInParent()
    out["reload_response_ok"] = {"b64": base64.b64encode(ok.SerializeToString(deterministic=True)).decode()}
    return out


def metadata_golden():
    """GetModelMetadata request / response (tfservingproxy.go:220-231 forwards them untouched) serialized from the
    reference's embedded schema: SignatureDefMap packed into google.protobuf.Any under metadata["signature_def"]."""
    build_pool()
    get = build_pool.get
    Req, Resp = get("tensorflow.serving.GetModelMetadataRequest"), get("tensorflow.serving.GetModelMetadataResponse")
    SigMap = get("tensorflow.serving.SignatureDefMap")
    out = {"cases": []}
    r = Req(); r.model_spec.name = "half_plus_two"; r.model_spec.version.value = 123; r.metadata_field.append("signature_def")
    out["request"] = {"name": "half_plus_two", "version": 123, "fields": ["signature_def"],
                      "b64": base64.b64encode(r.SerializeToString(deterministic=True)).decode()}
    r2 = Req(); r2.model_spec.name = "m0001"; r2.metadata_field.append("signature_def")
    out["request_no_version"] = {"name": "m0001", "version": None, "fields": ["signature_def"],
                                 "b64": base64.b64encode(r2.SerializeToString(deterministic=True)).decode()}
    for name, version, in_key, in_dtype, in_dims, out_key, out_dims in [
            ("half_plus_two", 123, "x", 1, [-1], "y", [-1]),
            ("m0001", 1, "x", 1, [-1, 9216], "y", [-1, 9216]),
            ("bert", 7, "input_ids", 3, [-1, 128], "logits", [-1, 2])]:
        sm = SigMap()
        sd = sm.signature_def["serving_default"]
        for key, dtype, dims, target in ((in_key, in_dtype, in_dims, sd.inputs), (out_key, 1, out_dims, sd.outputs)):
            ti = target[key]
            ti.name = key + ":0"
            ti.dtype = dtype
            for d in dims:
                ti.tensor_shape.dim.add().size = d
        sd.method_name = "tensorflow/serving/predict"
        resp = Resp()
        resp.model_spec.name = name
        resp.model_spec.version.value = version
        resp.metadata["signature_def"].type_url = "type.googleapis.com/tensorflow.serving.SignatureDefMap"
        resp.metadata["signature_def"].value = sm.SerializeToString(deterministic=True)
        out["cases"].append({"name": name, "version": version, "input": [in_key, in_dtype, in_dims], "output": [out_key, 1, out_dims],
                             "b64": base64.b64encode(resp.SerializeToString(deterministic=True)).decode()})
    return out


def ring_golden():
    from oracle import ring
    out = {"crc": {k: ring.crc32_ieee(k.encode()) for k in
                   ["", "a", "123456789", "FoobarA", "half_plus_two##123", "half_plus_two##00000123", "x" * 100]},
           "cases": []}
    for n_members, fmt, k in [(5, "testhost_{i}:{r}:{g}", 3), (100, "testhost_{i}:{r}:{g}", 3),
                              (8, "gpu{i}:0:0", 2), (1, "testhost_{i}:{r}:{g}", 3), (200, "testhost_{i}:{r}:{g}", 1)]:
        members = [fmt.format(i=i, r=8000 + i, g=2000 + i) for i in range(n_members)]
        c = ring.Consistent()
        c.set(members)
        keys = ["FoobarA", "FoobarB", "FoobarC", "FoobarD", "FoobarE", "FoobarF"] + \
               [f"m{j:04d}##1" for j in range(0, 1000, 37)] + ["half_plus_two##123", "half_plus_two##00000123"]
        out["cases"].append({"members": members, "n": k, "points": len(c.sorted_hashes),
                             "placements": {key: c.get_n(key, k) for key in keys}})
    return out


def trace_golden():
    from oracle import cachemanager as ocm
    from oracle.lrucache import Model, ModelIdentifier
    from tools.traces import zipf_trace

    class Prov:
        def __init__(self, sizes):
            self.sizes = sizes

        def model_size(self, name, ver):
            return self.sizes[name]

        def load_model(self, name, ver):
            return Model(ModelIdentifier(name, ver), f"{name}/{ver}", self.sizes[name])

    out = []
    for seed, n_models, cache_models, max_conc in [(1, 12, 6, 3), (2, 30, 10, 4), (3, 8, 8, 8), (4, 20, 5, 5)]:
        sizes = {f"m{j:04d}": 1280 for j in range(n_models)}
        cm = ocm.CacheManager(Prov(sizes), cache_models * 1280, max_conc)
        trace = zipf_trace(n_models, 300, 1.0, seed)
        outcomes = [cm.fetch_model(ModelIdentifier(f"m{j:04d}", 1)) for j in trace]
        out.append({"seed": seed, "n_models": n_models, "cache_models": cache_models, "max_concurrent": max_conc,
                    "trace": [int(j) for j in trace], "outcomes": outcomes,
                    "hits": cm.hits, "misses": cm.misses, "total": cm.total,
                    "final_host": [m.identifier.model_name for m in cm.local_cache.list_models()],
                    "final_resident": [m.identifier.model_name for m in cm.resident_prefix()]})
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    json.dump(wire_golden(), open(os.path.join(here, "wire_golden.json"), "w"), indent=1)
    json.dump(ring_golden(), open(os.path.join(here, "ring_golden.json"), "w"), indent=1)
    json.dump(trace_golden(), open(os.path.join(here, "trace_golden.json"), "w"), indent=1)
    json.dump(modelservice_golden(), open(os.path.join(here, "modelservice_golden.json"), "w"), indent=1)
    json.dump(metadata_golden(), open(os.path.join(here, "metadata_golden.json"), "w"), indent=1)
    print("golden fixtures written")
