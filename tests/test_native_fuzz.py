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
            ("common.cc", "ring.cc", "lru.cc", "parse.cc", "wire.cc", "model.cc")]
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe] + srcs
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in (build.stderr or "").lower() and "cannot find" in build.stderr.lower():
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=1", UBSAN_OPTIONS="print_stacktrace=1"))
    assert run.returncode == 0, (run.stdout + run.stderr)[-4000:]
    assert "fuzz ok" in run.stdout
