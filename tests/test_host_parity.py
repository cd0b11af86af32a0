"""Host-side logic of libtfsc_b200.so through the C ABI, compared with the oracle and with the
reference's own test vectors.  No GPU needed (no compute entry point is called)."""
import os
import random
import re

import numpy as np
import pytest

import tfservingcache_b200 as t
from oracle import diskprovider as odisk
from oracle import ring as oring
from oracle import urlmatch as ourl
from oracle.lrucache import LRUCache as OLRU
from oracle.lrucache import Model as OModel
from oracle.lrucache import ModelIdentifier as OId

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- ABI surface ------------------------------------------------------------------------------
def test_abi_exports_every_declared_symbol():
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "tfsc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(tfsc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 40
    lib = ctypes.CDLL(t._lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.tfsc_abi_version() == int(re.search(r"#define TFSC_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "tfsc_b200.h")).read()).group(1)) == 2


def test_no_device_fails_loudly():
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(t._lib.TfscError) as e:
        t.Server({"modelProvider.type": "synthetic"})
    assert e.value.code == t._lib.E_NO_DEVICE and "no CPU fallback" in str(e.value)


# ---- LRU: lrucache_test.go ported 1:1 onto the product ----------------------------------------
def _put(cache, v, size=10, name="foo"):
    ident = t.ModelIdentifier(name, v)
    return cache.put(ident, t.Model(ident, "/some/path", size))


def test_cache_add_get():
    cache = t.LRUCache("./cache", 1024)
    _put(cache, 42)
    m, avail = cache.get(t.ModelIdentifier("foo", 42))
    assert avail and m.path == "/some/path" and m.identifier == t.ModelIdentifier("foo", 42) and m.size_on_disk == 10


def test_cache_get_not_present():
    assert t.LRUCache("./cache", 1024).get(t.ModelIdentifier("foo", 42)) == (None, False)


def test_cache_removes_lru_seq_access():
    cache = t.LRUCache("./cache", 95)
    for i in range(1, 11):
        _put(cache, i)
    assert cache.get(t.ModelIdentifier("foo", 1))[1] is False
    assert cache.get(t.ModelIdentifier("foo", 2))[1] is True
    assert cache.current_size == 90


def test_cache_removes_lru_non_seq_access():
    cache = t.LRUCache("./cache", 100)
    for i in range(1, 11):
        _put(cache, i)
    cache.get(t.ModelIdentifier("foo", 1))
    _put(cache, 11)
    assert cache.get(t.ModelIdentifier("foo", 1))[1] is True
    assert cache.get(t.ModelIdentifier("foo", 2))[1] is False


def test_cache_removes_lru_var_sizes():
    cache = t.LRUCache("./cache", 100)
    for i in range(4, 0, -1):
        _put(cache, i, 10 * i)
    _put(cache, 5, 20)
    assert cache.get(t.ModelIdentifier("foo", 4))[1] is False
    assert cache.current_size == 80 and len(cache.list_models()) == 4
    _put(cache, 6, 20)
    assert len(cache.list_models()) == 5


@pytest.mark.parametrize("seed", range(8))
def test_lru_random_trace_matches_oracle(seed):
    rng = random.Random(seed)
    cap = rng.choice([50, 100, 1000, 7])
    prod, orc = t.LRUCache("d", cap), OLRU("d", cap)
    for _ in range(600):
        op = rng.random()
        name, ver = rng.choice("abc"), rng.randrange(12)
        if op < 0.5:
            size = rng.choice([1, 5, 10, 33, 120])
            prod.put(t.ModelIdentifier(name, ver), t.Model(t.ModelIdentifier(name, ver), f"{name}/{ver}", size))
            orc.put(OId(name, ver), OModel(OId(name, ver), f"{name}/{ver}", size))
        elif op < 0.9:
            pm, pa = prod.get(t.ModelIdentifier(name, ver))
            om, oa = orc.get(OId(name, ver))
            assert pa == oa and (not pa or pm.size_on_disk == om.size_on_disk)
        else:
            n = rng.randrange(0, cap + 10)
            prod.ensure_free_bytes(n)
            orc.ensure_free_bytes(n)
        assert prod.current_size == orc.current_size
        assert [(m.identifier.model_name, m.identifier.version, m.size_on_disk) for m in prod.list_models()] == \
               [(m.identifier.model_name, m.identifier.version, m.size_on_disk) for m in orc.list_models()]


# ---- ring: cluster_test.go ported + bit-exact parity with the oracle --------------------------
NODE_NAMES = ["FoobarA", "FoobarB", "FoobarC", "FoobarD", "FoobarE", "FoobarF"]


def _members(n):
    return [t.ServingService(f"testhost_{i}", 2000 + i, 8000 + i) for i in range(n)]


def test_consistent_hashing_for_nodes():
    c = t.ClusterConnection(3)
    c.update(_members(100))
    first = {n: c.find_node_for_key(n) for n in NODE_NAMES}
    for _ in range(300):
        for n in NODE_NAMES:
            assert c.find_node_for_key(n) == first[n]
    assert len(first) == len(NODE_NAMES)


def test_membership_with_one_node():
    c = t.ClusterConnection(3)
    c.update(_members(1))
    for n in NODE_NAMES:
        nodes = c.find_node_for_key(n)
        assert len(nodes) == 1 and nodes[0].host == "testhost_0"


def test_consistent_hashing_during_membership_change():
    c = t.ClusterConnection(3)
    c.update(_members(5))
    first = {n: c.find_node_for_key(n) for n in NODE_NAMES}
    c.update(_members(200))
    assert any(c.find_node_for_key(n) != first[n] for n in NODE_NAMES)
    c.update(_members(5))
    assert all(c.find_node_for_key(n) == first[n] for n in NODE_NAMES)


def test_empty_ring_is_an_error():
    with pytest.raises(t._lib.TfscError) as e:
        t.ClusterConnection(2).find_node_for_key("x")
    assert e.value.code == t._lib.E_EMPTY_RING


def test_crc32_matches_oracle():
    rng = random.Random(0)
    for n in [0, 1, 7, 8, 9, 63, 64, 65, 1000]:
        data = bytes(rng.randrange(256) for _ in range(n))
        assert t.crc32_ieee(data) == oring.crc32_ieee(data)
    assert t.crc32_ieee(b"123456789") == 0xCBF43926


def test_ring_matches_committed_golden(golden):
    for case in golden("ring_golden.json")["cases"]:
        c = t.ClusterConnection(case["n"])
        c.update([t.ServingService.from_string(m) for m in case["members"]])
        assert c.points == case["points"] and c.members == len(case["members"])
        for key, want in case["placements"].items():
            assert [str(s) for s in c.find_node_for_key(key)] == want


@pytest.mark.parametrize("seed", range(6))
def test_ring_random_membership_matches_oracle(seed):
    rng = random.Random(seed)
    prod, orc = t.ClusterConnection(rng.randrange(1, 5)), None
    orc = oring.ClusterConnection(prod.replicas_per_model)
    for _round in range(6):
        n = rng.randrange(1, 40)
        ids = rng.sample(range(60), n)
        prod.update([t.ServingService(f"h{i}", 2000 + i, 8000 + i) for i in ids])
        orc.update([oring.ServingService(f"h{i}", 2000 + i, 8000 + i) for i in ids])
        for j in range(200):
            key = t.model_key(f"model_{rng.randrange(5000)}", str(rng.randrange(1, 4)))
            assert [str(s) for s in prod.find_node_for_key(key)] == [str(s) for s in orc.find_node_for_key(key)]


def test_rest_and_grpc_keys_route_differently_quirk():
    # appendix B: ring key uses the verbatim version string
    assert t.model_key("half_plus_two", "00000123") == "half_plus_two##00000123"
    c = t.ClusterConnection(3)
    c.update(_members(5))
    a = [s.host for s in c.find_node_for_key(t.model_key("half_plus_two", "123"))]
    b = [s.host for s in c.find_node_for_key(t.model_key("half_plus_two", "00000123"))]
    assert a == ["testhost_2", "testhost_4", "testhost_1"] and b == ["testhost_2", "testhost_0", "testhost_1"]


def test_task_handler_picks_among_replicas():
    c = t.ClusterConnection(3)
    c.update(_members(10))
    th = t.TaskHandler(c, seed=1)
    want = {str(s) for s in c.find_node_for_key(t.model_key("m", "1"))}
    seen = {str(th.node_for_key("m", "1")) for _ in range(200)}
    assert seen == want


# ---- request parsing: tfservingproxy_test.go ported + oracle parity ---------------------------
URLS = ["/v1/models/foobar/versions/42", "/v1/thisisabadrequest/foobar/versions/42", "/v1/models/foobar",
        "/V1/MODELS/foobar/VERSIONS/7:predict", "/v1/models/foobar/versions/", "/v1/models/foobar/versions/x1",
        "/v1/models//versions/3", "/v1/models/a/versions/00000123:predict?x=1", "/v1/models/a/versions/12/metadata",
        "/v1/models/foo:predict", "", "/", "/v1/models/", "/v1/models/a/version/3", "prefix/v1/models/a/versions/3",
        "/v1/models/a%20b/versions/9", "/v1/models/a/labels/stable"]


@pytest.mark.parametrize("url", URLS)
def test_rest_url_matches_oracle(url):
    assert t.match_rest_url(url) == ourl.match_rest_url(url)


def test_http_proxy_vectors():
    assert t.match_rest_url("/v1/models/foobar/versions/42")[:3] == (200, "foobar", "42")
    assert t.match_rest_url("/v1/thisisabadrequest/foobar/versions/42")[0] == 404
    assert t.match_rest_url("/v1/models/foobar")[0] == 400


@pytest.mark.parametrize("v", ["42", "00000123", "0", "-5", "+7", "9223372036854775807", "-9223372036854775808",
                               "9223372036854775808", "", "1x", " 1", "1.0", "--1"])
def test_parse_version_matches_oracle(v):
    try:
        want = ourl.parse_version(v)
    except ValueError:
        with pytest.raises(ValueError):
            t.parse_version(v)
    else:
        assert t.parse_version(v) == want


def test_grpc_proxy_parses_request(golden):
    from oracle import wire
    req = wire.encode_predict_request("foobar", 42, {"x": np.zeros((1, 2), np.float32)})
    assert t.grpc_model_spec(req) == ("foobar", "42")
    assert t.grpc_model_spec(wire.encode_predict_request("foobar", None, {})) == ("foobar", "0")
    import base64
    for c in golden("wire_golden.json")["requests"]:
        name, ver = t.grpc_model_spec(base64.b64decode(c["request_b64"]))
        assert name == c["name"] and ver == ourl.grpc_version_string(c["version"])


# ---- disk provider: diskmodelprovider_test.go ported ------------------------------------------
def _dummy(repo, name, version):
    d = os.path.join(repo, name, version)
    os.makedirs(os.path.join(d, "assets"))
    os.makedirs(os.path.join(d, "variables"))
    with open(os.path.join(d, "saved_model.pb"), "w") as f:
        f.write("x" * 10)


def test_disk_provider_loads_correct_model(tmp_path):
    repo = str(tmp_path)
    for v in ("42", "43", "4", "2", "0"):
        _dummy(repo, "myModel", v)
    _dummy(repo, "someDifferentModel", "22")
    _dummy(repo, "someDifferentModel", "42")
    p = t.DiskModelProvider(repo)
    assert p.find_src_path_for_model("myModel", 42) == os.path.join(repo, "myModel", "42")
    assert p.find_src_path_for_model("myModel", 42) == odisk.find_src_path_for_model(os.path.join(repo, "myModel"), 42)
    assert p.model_size("myModel", 42) == odisk.model_size(repo, "myModel", 42) == 10


def test_disk_provider_matches_prefix_zeros(tmp_path):
    repo = str(tmp_path)
    for v in ("000000042", "000000043", "41"):
        _dummy(repo, "myModel", v)
    p = t.DiskModelProvider(repo)
    assert p.find_src_path_for_model("myModel", 42).endswith("000000042")
    with pytest.raises(FileNotFoundError):
        p.find_src_path_for_model("myModel", 44)
    with pytest.raises(FileNotFoundError):
        p.find_src_path_for_model("nope", 1)


def test_disk_provider_refuses_names_that_leave_base_dir(tmp_path):
    """gRPC ModelSpec.name is not constrained by the REST regex: '..' / 'a/b' must not reach the file system."""
    repo = tmp_path / "repo"
    outside = tmp_path / "outside"
    _dummy(str(repo), "inside", "1")
    _dummy(str(tmp_path), "outside", "1")
    p = t.DiskModelProvider(str(repo))
    assert p.find_src_path_for_model("inside", 1).endswith("inside/1")
    for bad in ("../outside", "..", ".", "inside/../inside", "", "a/b"):
        with pytest.raises(FileNotFoundError):
            p.find_src_path_for_model(bad, 1)
        with pytest.raises(FileNotFoundError):
            p.model_size(bad, 1)
    assert outside.exists()


# ---- config surface (cfg.go:10-66) ------------------------------------------------------------
def test_config_yaml_and_env(tmp_path):
    cfgp = tmp_path / "config.yaml"
    cfgp.write_text("proxyRestPort: 8093\nmodelProvider:\n  type: diskProvider\n  diskProvider:\n    baseDir: ./model_repo\n"
                    "modelCache:\n  size: 30000\nserving:\n  maxConcurrentModels: 2\nproxy:\n  replicasPerModel: 3\n")
    cfg = t.load_config(str(cfgp), env={"TFSC_SERVING_MAXCONCURRENTMODELS": "32", "TFSC_PROXY_REPLICASPERMODEL": "2",
                                        "TFSC_LOGLEVEL": "debug", "OTHER": "x"})
    assert cfg["modelProvider.type"] == "diskProvider" and cfg["modelProvider.diskProvider.baseDir"] == "./model_repo"
    assert cfg["serving.maxConcurrentModels"] == 32 and cfg["proxy.replicasPerModel"] == 2
    assert cfg["modelCache.size"] == 30000 and cfg["logging.level"] == "debug"
    assert cfg["healthprobe.modelName"] == "__TFSERVINGCACHE_PROBE_CHECK__"


# ---- replica pick policies (taskhandler.go:91 is "random") --------------------------------------
def test_replica_picker_policies():
    import collections
    first = t.ReplicaPicker("first", 1)
    assert {first.pick("k", 3, 8) for _ in range(50)} == {0}
    rnd = t.ReplicaPicker("random", 1)
    c = collections.Counter(rnd.pick("k", 2, 8) for _ in range(2000))
    assert 800 < c[0] < 1200 and set(c) == {0, 1}
    # same seed + same call sequence -> same decisions (ranks agree without communicating)
    a, b = t.ReplicaPicker("hot-spread", 42), t.ReplicaPicker("hot-spread", 42)
    from tools.traces import zipf_trace
    tr = zipf_trace(1000, 30000, 1.0, 42).tolist()
    pa = [a.pick(f"m{m}##1", 2, 8) for m in tr]
    assert pa == [b.pick(f"m{m}##1", 2, 8) for m in tr]
    spread = {m for m, p in zip(tr, pa) if p == 1}
    top = [m for m, _ in collections.Counter(tr).most_common(2)]
    assert spread == set(top)          # only the models above 0.5/8 of the traffic leave their primary
    with pytest.raises(ValueError):
        t.ReplicaPicker("nope", 1)
    assert t.ReplicaPicker("hot-spread", 3).pick("k", 1, 8) == 0


def test_balanced_picker_evens_out_keys_per_member():
    """The 20-vnode ring alone spreads 1000 keys over 8 members within about +-16 %; binding each cold key to the
    less loaded of its two replicas (sticky power-of-two-choices) brings that to a few keys."""
    import collections
    c = t.ClusterConnection(2)
    members = [t.ServingService(f"gpu{i}", 0, 0) for i in range(8)]
    c.update(members)
    pk, pk2 = t.ReplicaPicker("balanced", 7, 0.25), t.ReplicaPicker("balanced", 7, 0.25)
    bound, first = {}, collections.Counter()
    for j in range(1000):
        key = t.model_key(f"m{j}", "1")
        ids = [int(s.host[3:]) for s in c.find_node_for_key(key)]
        i = pk.pick_ids(key, ids, 8)
        assert i == pk2.pick_ids(key, ids, 8)            # deterministic
        bound[key] = ids[i]
        first[ids[0]] += 1
    load = collections.Counter(bound.values())
    assert max(first.values()) - min(first.values()) > 15    # the ring alone is uneven
    assert max(load.values()) - min(load.values()) <= 4      # balanced binding is not
    for key, gpu in list(bound.items())[:200]:              # sticky: a bound key keeps its member
        ids = [int(s.host[3:]) for s in c.find_node_for_key(key)]
        assert ids[pk.pick_ids(key, ids, 8)] == gpu


def test_native_ring_reproduces_upstream_module_test_vectors():
    """The C++ ring (csrc/ring.cc) against the known answers of stathat.com/c/consistent's own tests
    (see tests/test_oracle_pins.py for their provenance)."""
    import ctypes as C
    from test_oracle_pins import UPSTREAM_GET, UPSTREAM_GET_AFTER_REMOVING_HIJKLMN, UPSTREAM_GETN, UPSTREAM_MEMBERS
    lib = t._lib.lib
    h = lib.tfsc_ring_new()

    def set_members(ms):
        arr = (C.c_char_p * len(ms))(*[m.encode() for m in ms])
        assert lib.tfsc_ring_set(h, arr, len(ms)) >= 0

    def get_n(key, n):
        buf = C.create_string_buffer(1024)
        cnt = lib.tfsc_ring_getn(h, key.encode(), n, buf, 1024)
        assert cnt >= 0
        return buf.value.decode().split("\n") if cnt else []

    try:
        set_members(UPSTREAM_MEMBERS)
        assert lib.tfsc_ring_members(h) == 3 and lib.tfsc_ring_points(h) == 60
        for key, want in UPSTREAM_GET:
            assert get_n(key, 1) == [want]
        for key, n, want in UPSTREAM_GETN:
            assert get_n(key, n) == want
        set_members(["abcdefg", "opqrstu"])
        for key, want in UPSTREAM_GET_AFTER_REMOVING_HIJKLMN:
            assert get_n(key, 1) == [want]
    finally:
        lib.tfsc_ring_free(h)
