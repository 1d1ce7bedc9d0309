"""The C oracle under AddressSanitizer + UndefinedBehaviorSanitizer (the reference runs no race or
sanitizer tooling at all — Makefile:86 `go test` without -race)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_selftest_under_asan_ubsan():
    odir = os.path.join(ROOT, "oracle")
    build = subprocess.run(["make", "-s", "-C", odir, "selftest_san"], capture_output=True, text=True)
    if build.returncode != 0 and ("-lasan" in build.stderr or "No such file" in build.stderr):
        pytest.skip("no sanitizer runtime for this compiler: " + build.stderr.strip().splitlines()[-1])
    assert build.returncode == 0, build.stderr
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1")
    out = subprocess.run([os.path.join(odir, "selftest_san")], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "oracle selftest ok" in out.stdout
