"""Generates tests/golden/examples_golden.json: Classify / Regress / SessionRun requests and the responses TF-Serving's
half_plus_two model gives ([1,2,5] -> [2.5,3,4.5]), serialized by python-protobuf from the FileDescriptorProtos embedded in the
reference's generated code (proto/tensorflow/serving/{classification,regression,input,session_service}.pb.go,
core/example/{example,feature}.pb.go, core/protobuf/named_tensor.pb.go). The native codec (csrc/wire.cc) must decode the
requests and emit the responses byte for byte. Run in the BUILD container:  python tests/golden/make_examples_golden.py"""
import base64
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def get_messages(names):
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from google.protobuf import any_pb2, wrappers_pb2  # noqa: F401
    pool = descriptor_pool.Default()
    fds = {}
    for sub in ("core/framework", "core/lib/core", "core/protobuf", "core/example", "serving"):
        d = os.path.join(mg.REF, sub)
        for fn in sorted(os.listdir(d)):
            if fn.endswith(".pb.go"):
                for blob in mg.embedded_descriptors(os.path.join(d, fn)):
                    fd = descriptor_pb2.FileDescriptorProto.FromString(blob)
                    fds[fd.name] = fd
    done = set()

    def add(name):
        if name in done or name not in fds:
            return
        done.add(name)
        for dep in fds[name].dependency:
            add(dep)
        try:
            pool.Add(fds[name])
        except Exception:
            pass
    for want in ("tensorflow_serving/apis/classification.proto", "tensorflow_serving/apis/regression.proto",
                 "tensorflow_serving/apis/session_service.proto"):
        add(want)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName(n)) for n in names}


def b64(m):
    return base64.b64encode(m.SerializeToString(deterministic=True)).decode()


def main():
    M = get_messages(["tensorflow.serving.ClassificationRequest", "tensorflow.serving.ClassificationResponse",
                      "tensorflow.serving.RegressionRequest", "tensorflow.serving.RegressionResponse",
                      "tensorflow.serving.SessionRunRequest", "tensorflow.serving.SessionRunResponse"])
    xs, ys = [1.0, 2.0, 5.0], [2.5, 3.0, 4.5]
    out = {"model": "half_plus_two", "version": 123, "x": xs, "y": ys}

    def fill_examples(req, with_context=False):
        req.model_spec.name = "half_plus_two"
        req.model_spec.version.value = 123
        lst = req.input.example_list_with_context if with_context else req.input.example_list
        for v in xs:
            ex = lst.examples.add()
            if not with_context:
                ex.features.feature["x"].float_list.value.append(v)
            ex.features.feature["ignored_bytes"].bytes_list.value.append(b"abc")
        if with_context:
            lst.context.features.feature["x"].float_list.value.append(2.0)   # every example inherits x = 2

    rq = M["tensorflow.serving.RegressionRequest"]()
    fill_examples(rq)
    rq.model_spec.signature_name = "regress_x_to_y"
    rs = M["tensorflow.serving.RegressionResponse"]()
    for v in ys:
        rs.result.regressions.add().value = v
    rs.model_spec.name, rs.model_spec.signature_name = "half_plus_two", "regress_x_to_y"
    rs.model_spec.version.value = 123
    out["regress"] = {"request_b64": b64(rq), "response_b64": b64(rs)}

    cq = M["tensorflow.serving.ClassificationRequest"]()
    fill_examples(cq)
    cq.model_spec.signature_name = "classify_x_to_y"
    cs = M["tensorflow.serving.ClassificationResponse"]()
    for v in ys:
        cs.result.classifications.add().classes.add().score = v
    cs.model_spec.name, cs.model_spec.signature_name = "half_plus_two", "classify_x_to_y"
    cs.model_spec.version.value = 123
    out["classify"] = {"request_b64": b64(cq), "response_b64": b64(cs)}

    cq2 = M["tensorflow.serving.RegressionRequest"]()
    fill_examples(cq2, with_context=True)
    cq2.model_spec.signature_name = "regress_x_to_y"
    rs2 = M["tensorflow.serving.RegressionResponse"]()
    for _ in xs:
        rs2.result.regressions.add().value = 3.0
    rs2.model_spec.name, rs2.model_spec.signature_name = "half_plus_two", "regress_x_to_y"
    rs2.model_spec.version.value = 123
    out["regress_with_context"] = {"request_b64": b64(cq2), "response_b64": b64(rs2)}

    # default signature of half_plus_two is a predict signature: Classify on it is an error in TF-Serving
    cq3 = M["tensorflow.serving.ClassificationRequest"]()
    fill_examples(cq3)
    out["classify_on_predict_signature"] = {"request_b64": b64(cq3)}

    sq = M["tensorflow.serving.SessionRunRequest"]()
    sq.model_spec.name = "half_plus_two"
    sq.model_spec.version.value = 123
    f = sq.feed.add()
    f.name = "x:0"
    f.tensor.dtype = 1
    f.tensor.tensor_shape.dim.add().size = 3
    f.tensor.float_val.extend(xs)
    sq.fetch.append("y:0")
    ss = M["tensorflow.serving.SessionRunResponse"]()
    t = ss.tensor.add()
    t.name = "y:0"
    t.tensor.dtype = 1
    t.tensor.tensor_shape.dim.add().size = 3
    t.tensor.float_val.extend(ys)
    ss.model_spec.name = "half_plus_two"
    ss.model_spec.version.value = 123
    out["session_run"] = {"request_b64": b64(sq), "response_b64": b64(ss)}
    with open(os.path.join(HERE, "examples_golden.json"), "w") as fjs:
        json.dump(out, fjs, indent=1)
    print({k: (len(v["request_b64"]) if isinstance(v, dict) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
