"""Pins the CPU oracle against every vector the reference's own tests hold for the path
 before anything is compared against the oracle."""
import os
import zlib

import numpy as np
import pytest

from oracle import cachemanager as ocm
from oracle import diskprovider, models, ring, urlmatch
from oracle.lrucache import LRUCache, Model, ModelIdentifier


# ---- lrucache_test.go:7-115 -----------------------------------------------------------------
def _put(cache, v, size=10, name="foo"):
    ident = ModelIdentifier(name, v)
    cache.put(ident, Model(ident, "/some/path", size))


def test_cache_add_get():
    cache = LRUCache("./cache", 1024)
    _put(cache, 42)
    m, avail = cache.get(ModelIdentifier("foo", 42))
    assert avail and m.path != "" and m.identifier == ModelIdentifier("foo", 42) and m.size_on_disk == 10


def test_cache_get_not_present():
    assert LRUCache("./cache", 1024).get(ModelIdentifier("foo", 42))[1] is False


def test_cache_removes_lru_seq_access():
    cache = LRUCache("./cache", 95)
    for i in range(1, 11):
        _put(cache, i)
    assert cache.get(ModelIdentifier("foo", 1))[1] is False
    assert cache.get(ModelIdentifier("foo", 2))[1] is True
    assert cache.current_size == 90


def test_cache_removes_lru_non_seq_access():
    cache = LRUCache("./cache", 100)
    for i in range(1, 11):
        _put(cache, i)
    cache.get(ModelIdentifier("foo", 1))
    _put(cache, 11)
    assert cache.get(ModelIdentifier("foo", 1))[1] is True
    assert cache.get(ModelIdentifier("foo", 2))[1] is False


def test_cache_removes_lru_var_sizes():
    cache = LRUCache("./cache", 100)
    for i in range(4, 0, -1):
        _put(cache, i, 10 * i)
    _put(cache, 5, 20)
    assert cache.get(ModelIdentifier("foo", 4))[1] is False
    assert cache.current_size == 80 and len(cache.list_models()) == 4
    _put(cache, 6, 20)
    assert len(cache.list_models()) == 5


# ---- tfservingproxy_test.go:111-234 ---------------------------------------------------------
def test_http_proxy_parses_url():
    assert urlmatch.match_rest_url("/v1/models/foobar/versions/42")[:3] == (200, "foobar", "42")


def test_http_proxy_invalid_url_404():
    st, _, _, body = urlmatch.match_rest_url("/v1/thisisabadrequest/foobar/versions/42")
    assert st == 404 and body == '{"Status":"Error","Message":"Not found"}\n'


def test_http_proxy_no_version_400():
    st, name, _, body = urlmatch.match_rest_url("/v1/models/foobar")
    assert st == 400 and name == "foobar" and body == '{"Status":"Error","Message":"Model version must be provided"}\n'


def test_grpc_proxy_version_string():
    assert urlmatch.grpc_version_string(42) == "42"
    assert urlmatch.grpc_version_string(None) == "0"  # missing Int64Value -> "0" (tfservingproxy.go:248)


def test_leading_zero_version_kept_verbatim_then_normalised():
    st, name, ver, _ = urlmatch.match_rest_url("/v1/models/saved_model_half_plus_two_cpu/versions/00000123:predict")
    assert (st, name, ver) == (200, "saved_model_half_plus_two_cpu", "00000123")
    assert urlmatch.parse_version(ver) == 123


# ---- diskmodelprovider_test.go:33-87 --------------------------------------------------------
def _dummy(repo, name, version):
    d = os.path.join(repo, name, version)
    os.makedirs(os.path.join(d, "assets"))
    os.makedirs(os.path.join(d, "variables"))
    open(os.path.join(d, "saved_model.pb"), "w").close()


def test_disk_provider_loads_correct_model(tmp_path):
    repo = str(tmp_path)
    for v in ("42", "43", "4", "2", "0"):
        _dummy(repo, "myModel", v)
    _dummy(repo, "someDifferentModel", "22")
    _dummy(repo, "someDifferentModel", "42")
    assert diskprovider.find_src_path_for_model(os.path.join(repo, "myModel"), 42).endswith("myModel/42")


def test_disk_provider_matches_prefix_zeros(tmp_path):
    repo = str(tmp_path)
    for v in ("000000042", "000000043", "41"):
        _dummy(repo, "myModel", v)
    assert diskprovider.find_src_path_for_model(os.path.join(repo, "myModel"), 42).endswith("000000042")
    with pytest.raises(FileNotFoundError):
        diskprovider.find_src_path_for_model(os.path.join(repo, "myModel"), 44)


# ---- cluster_test.go:51-227 (properties) + CRC known answers --------------------------------
def _members(n):
    return [ring.ServingService(f"testhost_{i}", 2000 + i, 8000 + i) for i in range(n)]


NODE_NAMES = ["FoobarA", "FoobarB", "FoobarC", "FoobarD", "FoobarE", "FoobarF"]


def test_crc32_known_answers():
    assert ring.crc32_ieee(b"123456789") == 0xCBF43926  # CRC-32/IEEE check value
    for s in (b"", b"a", b"FoobarA", b"half_plus_two##123", os.urandom(300)):
        assert ring.crc32_ieee(s) == zlib.crc32(s)


def test_consistent_hashing_for_nodes():
    c = ring.ClusterConnection(3)
    c.update(_members(100))
    first = {n: c.find_node_for_key(n) for n in NODE_NAMES}
    for _ in range(200):
        for n in NODE_NAMES:
            assert c.find_node_for_key(n) == first[n]
    assert all(len(v) == 3 and len(set(v)) == 3 for v in first.values())


def test_membership_with_one_node():
    c = ring.ClusterConnection(3)
    c.update(_members(1))
    for n in NODE_NAMES:
        nodes = c.find_node_for_key(n)
        assert len(nodes) == 1 and nodes[0].host == "testhost_0"


def test_consistent_hashing_during_membership_change():
    c = ring.ClusterConnection(3)
    c.update(_members(5))
    first = {n: c.find_node_for_key(n) for n in NODE_NAMES}
    c.update(_members(200))
    assert any(c.find_node_for_key(n) != first[n] for n in NODE_NAMES)
    c.update(_members(5))
    assert all(c.find_node_for_key(n) == first[n] for n in NODE_NAMES)


def test_empty_ring_errors():
    with pytest.raises(ring.EmptyCircleError):
        ring.ClusterConnection(1).find_node_for_key("x")


def test_ring_matches_committed_golden(golden):
    g = golden("ring_golden.json")
    for k, v in g["crc"].items():
        assert ring.crc32_ieee(k.encode()) == v
    for case in g["cases"]:
        c = ring.Consistent()
        c.set(case["members"])
        assert len(c.sorted_hashes) == case["points"]
        for key, want in case["placements"].items():
            assert c.get_n(key, case["n"]) == want


# ---- half_plus_two known answer (deploy/docker-compose/readme.md:40-42) ---------------------
def test_half_plus_two_known_answer():
    man, blob = models.affine_blob(0.5, 2.0)
    y = models.forward(man, blob, np.array([1.0, 2.0, 5.0], np.float32))
    assert y.tolist() == [2.5, 3.0, 4.5]


# ---- numeric oracle self-consistency: numpy fp32 vs C restatement vs fp64 -------------------
def test_mlp_oracle_c_restatement_agrees():
    import ctypes as C
    so = os.path.join(os.path.dirname(models.__file__), "liboracle_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/liboracle_ref.so not built (make -C oracle)")
    lib = C.CDLL(so)
    dims = [96, 160, 64, 24]
    man, blob = models.synth_mlp_blob(dims, seed=1003)
    # the C generator is bit-identical to the numpy one
    buf = np.empty(dims[0] * dims[1], np.float32)
    lib.oracle_synth_fill(buf.ctypes.data_as(C.c_void_p), C.c_uint32(1003), C.c_uint32(0), C.c_uint64(0),
                          C.c_uint64(buf.size), C.c_float(models.weight_scale(dims[0])))
    assert np.array_equal(buf, blob[:buf.size])
    x = np.random.default_rng(0).standard_normal((5, dims[0])).astype(np.float32)
    y32 = models.forward(man, blob, x, np.float32)
    y64 = models.forward(man, blob, x, np.float64)
    I64 = C.c_int64 * 4
    I3 = C.c_int64 * 3
    relu = (C.c_int * 3)(*[1 if L["activation"] == "relu" else 0 for L in man["layers"]])
    for acc64, ref, tol in ((0, y32, 1e-5), (1, y64, 1e-6)):
        out = np.empty((5, dims[-1]), np.float32)
        rc = lib.oracle_mlp_forward(blob.ctypes.data_as(C.c_void_p), 3, I64(*dims),
                                    I3(*[L["w_offset"] for L in man["layers"]]),
                                    I3(*[L["b_offset"] for L in man["layers"]]), relu,
                                    x.ctypes.data_as(C.c_void_p), C.c_int64(5), out.ctypes.data_as(C.c_void_p), acc64)
        assert rc == 0
        np.testing.assert_allclose(out, ref, rtol=tol, atol=tol)
    assert lib.oracle_crc32_ieee(b"123456789", 9) & 0xFFFFFFFF == 0xCBF43926


# ---- residency machine: reference branches (cachemanager.go:103-150) ------------------------
class _Prov:
    def __init__(self, size=10):
        self.size = size
        self.loads = []

    def model_size(self, name, ver):
        if name.startswith("missing"):
            raise FileNotFoundError("No matching model found")
        return self.size

    def load_model(self, name, ver):
        self.loads.append((name, ver))
        return Model(ModelIdentifier(name, ver), f"{name}/{ver}", self.size)


def test_residency_hit_reload_miss_branches():
    cm = ocm.CacheManager(_Prov(), cache_bytes=40, max_concurrent_models=2)
    A, B, C_, D, E = (ModelIdentifier(n, 1) for n in "ABCDE")
    assert [cm.fetch_model(m) for m in (A, B, A)] == ["miss", "miss", "hit"]
    assert cm.fetch_model(C_) == "miss"              # resident := [C, A]; B falls out of the top-2
    assert cm.serving.status(B) == ocm.END
    assert cm.fetch_model(B) == "reload"             # cached on host, not resident -> reload only
    assert [m.identifier.model_name for m in cm.resident_prefix()] == ["B", "C"]
    assert cm.fetch_model(D) == "miss" and cm.fetch_model(E) == "miss"   # host tier (4 x 10 bytes) evicts A
    assert cm.fetch_model(A) == "miss"
    assert (cm.total, cm.hits, cm.misses) == (8, 1, 6)
    with pytest.raises(FileNotFoundError):
        cm.fetch_model(ModelIdentifier("missing", 1))


def test_residency_matches_committed_golden(golden):
    for case in golden("trace_golden.json"):
        sizes = 1280

        class P(_Prov):
            pass
        cm = ocm.CacheManager(P(sizes), case["cache_models"] * sizes, case["max_concurrent"])
        out = [cm.fetch_model(ModelIdentifier(f"m{j:04d}", 1)) for j in case["trace"]]
        assert out == case["outcomes"]
        assert (cm.hits, cm.misses, cm.total) == (case["hits"], case["misses"], case["total"])


def test_create_model_config_groups_by_name():
    ms = [Model(ModelIdentifier("a", 1), "", 1), Model(ModelIdentifier("b", 7), "", 1), Model(ModelIdentifier("a", 2), "", 1)]
    cfg = ocm.create_model_config(ms, "/models")
    assert [(c["name"], c["versions"], c["base_path"]) for c in cfg] == [("a", [1, 2], "/models/a"), ("b", [7], "/models/b")]
    assert all(c["model_platform"] == "tensorflow" for c in cfg)


# ---- stathat.com/c/consistent v1.0.0 known answers --------------------------------------------------------------
# The module is not vendored in the reference (go.mod:25) and cannot be fetched here. These are the expectations of its
# own test-suite (consistent_test.go: TestGetMultiple, TestGetMultipleRemove, TestGetTwo, TestGetN, TestGetNLess,
# TestGetNMore), quoted from memory of the public source -- not read from disk. The restatement reproduces every one of
# them, which pins the vnode key format (strconv.Itoa(i) + member), the 20 replicas, CRC-32/IEEE and the clockwise walk.
UPSTREAM_MEMBERS = ["abcdefg", "hijklmn", "opqrstu"]
UPSTREAM_GET = [("ggg", "abcdefg"), ("hhh", "opqrstu"), ("iiiii", "hijklmn")]
UPSTREAM_GET_AFTER_REMOVING_HIJKLMN = [("ggg", "abcdefg"), ("hhh", "opqrstu"), ("iiiii", "opqrstu")]
UPSTREAM_GETN = [("99999999", 2, ["abcdefg", "hijklmn"]),               # TestGetTwo / TestGetNLess
                 ("9999999", 3, ["opqrstu", "abcdefg", "hijklmn"]),     # TestGetN
                 ("9999999", 5, ["opqrstu", "abcdefg", "hijklmn"])]     # TestGetNMore: n clamps to the member count


def test_ring_reproduces_upstream_module_test_vectors():
    c = ring.Consistent()
    c.set(UPSTREAM_MEMBERS)
    for key, want in UPSTREAM_GET:
        assert c.get_n(key, 1) == [want]
    for key, n, want in UPSTREAM_GETN:
        assert c.get_n(key, n) == want
    c.set(["abcdefg", "opqrstu"])
    for key, want in UPSTREAM_GET_AFTER_REMOVING_HIJKLMN:
        assert c.get_n(key, 1) == [want]
