"""Host-side parsers of the library under AddressSanitizer + UBSan with random / mutated inputs (no GPU, no CUDA):
malformed PredictRequest bytes, URLs, JSON bodies and manifests must be rejected without out-of-bounds access."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tfservingcache_b200", "csrc")


def test_host_parsers_survive_fuzzing_under_asan_ubsan(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "fuzz_host")
    srcs = [os.path.join(ROOT, "tests", "native", "fuzz_host.cc")] + [os.path.join(CSRC, f) for f in
            ("common.cc", "ring.cc", "lru.cc", "parse.cc", "wire.cc", "model.cc", "savedmodel.cc", "provider.cc")]
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe] + srcs + ["-lpthread"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in (build.stderr or "").lower() and "cannot find" in build.stderr.lower():
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-3000:]
    # a well-formed SavedModel fixture for the importer leg (written by the helpers of tests/test_savedmodel.py)
    import numpy as np
    from savedmodel_fixtures import _mlp_fixture
    fixture, scratch = tmp_path / "sm_fixture", tmp_path / "sm_scratch"
    _mlp_fixture(fixture, np.random.default_rng(0), (12, 20, 5))
    os.makedirs(scratch / "variables")
    os.makedirs(tmp_path / "base" / "m")
    os.symlink(fixture, tmp_path / "base" / "m" / "00000007")
    # Classify / Regress / SessionRun request + response bytes serialized from the reference's schema (tests/golden/examples_golden.json):
    # the native codec is checked against them here, on the CPU, and then fuzzed
    import base64
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "examples_golden.json")))
    exdir = tmp_path / "examples"
    os.makedirs(exdir)
    for key in ("regress", "classify", "regress_with_context", "session_run"):
        for part in ("request", "response"):
            if part + "_b64" in g[key]:
                open(exdir / f"{key}_{part}.bin", "wb").write(base64.b64decode(g[key][part + "_b64"]))
    run = subprocess.run([exe, "20000", str(fixture), str(scratch), str(tmp_path / "base"), str(exdir)], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=1", UBSAN_OPTIONS="print_stacktrace=1"))
    assert run.returncode == 0, (run.stdout + run.stderr)[-4000:]
    assert "fuzz ok" in run.stdout and "examples: " in run.stdout
