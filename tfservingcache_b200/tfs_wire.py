"""Wire codec (hand-rolled protobuf) for the two TF-Serving ModelService RPCs the reference's TFServingController
issues (pkg/cachemanager/servingcontroller.go:88-138): HandleReloadConfigRequest and GetModelStatus. Field numbers:
proto/tensorflow/serving/{get_model_status,model_management,model_server_config,file_system_storage_path_source,
status}.pb.go. Used by serve.py's TF-Serving facade so the UNMODIFIED reference can point serving.grpcHost at this
server instead of a tensorflow/serving container."""
from __future__ import annotations


def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _vi(field: int, v: int) -> bytes:
    return _varint(field << 3) + _varint(v)


def _fields(buf: bytes):
    pos, n = 0, len(buf)
    while pos < n:
        key = shift = 0
        while True:
            b = buf[pos]
            pos += 1
            key |= (b & 0x7F) << shift
            if not b & 0x80:
                break
            shift += 7
        f, wt = key >> 3, key & 7
        if wt == 0:
            v = shift = 0
            while True:
                b = buf[pos]
                pos += 1
                v |= (b & 0x7F) << shift
                if not b & 0x80:
                    break
                shift += 7
        elif wt == 2:
            ln = shift = 0
            while True:
                b = buf[pos]
                pos += 1
                ln |= (b & 0x7F) << shift
                if not b & 0x80:
                    break
                shift += 7
            if pos + ln > n:
                raise ValueError("truncated message")
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield f, wt, v


def _i64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def decode_get_model_status_request(buf: bytes):
    """-> (name, version or None)"""
    name, version = "", None
    for f, wt, v in _fields(buf):
        if f == 1 and wt == 2:  # ModelSpec
            for f2, wt2, v2 in _fields(v):
                if f2 == 1 and wt2 == 2:
                    name = v2.decode()
                elif f2 == 2 and wt2 == 2:
                    version = 0
                    for f3, wt3, v3 in _fields(v2):
                        if f3 == 1 and wt3 == 0:
                            version = _i64(v3)
    return name, version


def encode_get_model_status_request(name: str, version: int | None) -> bytes:
    spec = _ld(1, name.encode()) if name else b""
    if version is not None:
        spec += _ld(2, _vi(1, version) if version else b"")
    return _ld(1, spec)


def encode_get_model_status_response(statuses) -> bytes:
    """statuses: [(version, state, error_code, error_message)]"""
    out = b""
    for version, state, code, msg in statuses:
        st = (_vi(1, code) if code else b"") + (_ld(2, msg.encode()) if msg else b"")
        m = (_vi(1, version) if version else b"") + (_vi(2, state) if state else b"") + _ld(3, st)
        out += _ld(1, m)
    return out


def decode_get_model_status_response(buf: bytes):
    out = []
    for f, wt, v in _fields(buf):
        if f == 1 and wt == 2:
            version = state = code = 0
            msg = ""
            for f2, wt2, v2 in _fields(v):
                if f2 == 1 and wt2 == 0:
                    version = _i64(v2)
                elif f2 == 2 and wt2 == 0:
                    state = v2
                elif f2 == 3 and wt2 == 2:
                    for f3, wt3, v3 in _fields(v2):
                        if f3 == 1 and wt3 == 0:
                            code = v3
                        elif f3 == 2 and wt3 == 2:
                            msg = v3.decode()
            out.append((version, state, code, msg))
    return out


def decode_reload_config_request(buf: bytes):
    """-> [(name, base_path, model_platform, [versions])] in list order (createModelConfig, servingcontroller.go:159-187)"""
    models = []
    for f, wt, v in _fields(buf):
        if not (f == 1 and wt == 2):      # ReloadConfigRequest.config
            continue
        for f2, wt2, v2 in _fields(v):
            if not (f2 == 1 and wt2 == 2):  # ModelServerConfig.model_config_list
                continue
            for f3, wt3, v3 in _fields(v2):
                if not (f3 == 1 and wt3 == 2):  # ModelConfigList.config
                    continue
                name = base = platform = ""
                versions = []
                for f4, wt4, v4 in _fields(v3):
                    if f4 == 1 and wt4 == 2:
                        name = v4.decode()
                    elif f4 == 2 and wt4 == 2:
                        base = v4.decode()
                    elif f4 == 4 and wt4 == 2:
                        platform = v4.decode()
                    elif f4 == 7 and wt4 == 2:   # ServableVersionPolicy
                        for f5, wt5, v5 in _fields(v4):
                            if f5 == 102 and wt5 == 2:  # Specific
                                for f6, wt6, v6 in _fields(v5):
                                    if f6 == 1 and wt6 == 0:
                                        versions.append(_i64(v6))
                                    elif f6 == 1 and wt6 == 2:  # packed
                                        versions += [_i64(x) for _f, _w, x in _fields(b"".join(b"\x08" + _varint(y) for y in _unpack(v6)))]
                models.append((name, base, platform, versions))
    return models


def _unpack(buf: bytes):
    pos, n = 0, len(buf)
    while pos < n:
        v = shift = 0
        while True:
            b = buf[pos]
            pos += 1
            v |= (b & 0x7F) << shift
            if not b & 0x80:
                break
            shift += 7
        yield v


def encode_reload_config_request(models) -> bytes:
    cfgs = b""
    for name, base, platform, versions in models:
        spec = _ld(1, b"".join(_varint(v) for v in versions))
        m = _ld(1, name.encode()) + _ld(2, base.encode()) + _ld(4, platform.encode()) + _ld(7, _ld(102, spec))
        cfgs += _ld(1, m)
    return _ld(1, _ld(1, cfgs))


def encode_reload_config_response(code: int = 0, message: str = "") -> bytes:
    st = (_vi(1, code) if code else b"") + (_ld(2, message.encode()) if message else b"")
    return _ld(1, st)


def decode_reload_config_response(buf: bytes):
    code, msg = 0, ""
    for f, wt, v in _fields(buf):
        if f == 1 and wt == 2:
            for f2, wt2, v2 in _fields(v):
                if f2 == 1 and wt2 == 0:
                    code = v2
                elif f2 == 2 and wt2 == 2:
                    msg = v2.decode()
    return code, msg


# ---- PredictionService.GetModelMetadata (tfservingproxy.go:220-231 forwards it by model_spec) ------------------
# GetModelMetadataRequest{model_spec=1, metadata_field=2 rep string}; GetModelMetadataResponse{model_spec=1,
# metadata=2 map<string, google.protobuf.Any{type_url=1, value=2}>}; SignatureDefMap{signature_def=1 map<string,
# SignatureDef{inputs=1 map<string,TensorInfo>, outputs=2 map, method_name=3}>}; TensorInfo{name=1, dtype=2,
# tensor_shape=3 TensorShapeProto{dim=2{size=1}}} (proto/tensorflow/serving/get_model_metadata.pb.go:27,68-70,117-121;
# proto/tensorflow/core/protobuf/meta_graph.pb.go:658-662,698,937-948)
SIGNATURE_DEF_TYPE_URL = "type.googleapis.com/tensorflow.serving.SignatureDefMap"
DT_BY_NAME = {"DT_FLOAT": 1, "DT_INT32": 3, "DT_INT64": 9}


def decode_get_model_metadata_request(buf: bytes):
    """-> (name, version or None, [metadata_field])"""
    spec, fields = b"", []
    for f, wt, v in _fields(buf):
        if f == 1 and wt == 2:
            spec = v
        elif f == 2 and wt == 2:
            fields.append(v.decode())
    name, version = decode_get_model_status_request(_ld(1, spec))
    return name, version, fields


def encode_get_model_metadata_request(name: str, version: int | None, fields=("signature_def",)) -> bytes:
    return encode_get_model_status_request(name, version) + b"".join(_ld(2, f.encode()) for f in fields)


def _tensor_info(name: str, dtype: int, dims) -> bytes:
    shape = b"".join(_ld(2, _vi(1, d) if d else b"") for d in dims)
    return (_ld(1, name.encode()) if name else b"") + (_vi(2, dtype) if dtype else b"") + _ld(3, shape)


def _map_entry(field: int, key: str, value: bytes) -> bytes:
    return _ld(field, (_ld(1, key.encode()) if key else b"") + _ld(2, value))


def encode_signature_def_map(signatures: dict) -> bytes:
    """signatures: {sig_name: {"inputs": {key: (tensor_name, dtype, dims)}, "outputs": {...}, "method_name": str}}.
    Map entries are written in key order (what protobuf's deterministic serialization does)."""
    out = b""
    for sig in sorted(signatures):
        sd = signatures[sig]
        body = b"".join(_map_entry(1, k, _tensor_info(*sd["inputs"][k])) for k in sorted(sd["inputs"]))
        body += b"".join(_map_entry(2, k, _tensor_info(*sd["outputs"][k])) for k in sorted(sd["outputs"]))
        if sd.get("method_name"):
            body += _ld(3, sd["method_name"].encode())
        out += _map_entry(1, sig, body)
    return out


def encode_get_model_metadata_response(name: str, version: int | None, signatures: dict) -> bytes:
    spec = _ld(1, name.encode()) if name else b""
    if version is not None:
        spec += _ld(2, _vi(1, version) if version else b"")
    any_msg = _ld(1, SIGNATURE_DEF_TYPE_URL.encode()) + _ld(2, encode_signature_def_map(signatures))
    return _ld(1, spec) + _map_entry(2, "signature_def", any_msg)


def signatures_from_rest_metadata(doc: dict) -> dict:
    """The REST /metadata JSON (server.cu, TF-Serving's layout) -> the structure encode_signature_def_map takes."""
    out = {}
    for sig, sd in doc["metadata"]["signature_def"]["signature_def"].items():
        conv = lambda m: {k: (v.get("name", ""), DT_BY_NAME.get(v.get("dtype", ""), 0),
                              [int(d["size"]) for d in v.get("tensor_shape", {}).get("dim", [])]) for k, v in m.items()}  # noqa: E731
        out[sig] = {"inputs": conv(sd.get("inputs", {})), "outputs": conv(sd.get("outputs", {})), "method_name": sd.get("method_name", "")}
    return out
