"""The C-ABI library without a GPU: it loads, exports every symbol include/lig.h declares, its
pure-host helpers work, and the scheduling entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from llm_instance_gateway_b200 import _native as N
from llm_instance_gateway_b200.packer import PICK_DTYPE, REQ_DTYPE, RangeError, pack_columns, pack_pod_metrics
from llm_instance_gateway_b200.backend import Metrics, Pod, PodMetrics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lig.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lig_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"liblig.so does not export {name}"
    assert sorted(N.EXPORTED_SYMBOLS) == names
    assert lib.lig_abi_version() == N.LIG_ABI_VERSION
    assert b"sm_100a" in lib.lig_version()


def test_record_sizes_match_header():
    assert C.sizeof(N.LigReq) == 16 and REQ_DTYPE.itemsize == 16
    assert C.sizeof(N.LigPick) == 8 and PICK_DTYPE.itemsize == 8
    assert C.sizeof(N.LigThresholds) == 24


def test_snapshot_bytes_and_layout():
    lib = N.load()
    for P, A in [(0, 0), (1, 0), (1, 1), (33, 5), (64, 32), (512, 256), (4096, 1024)]:
        W = (P + 31) // 32
        Ppad = 32 * W
        want = max(16, (16 * Ppad + 4 * A * W + 15) // 16 * 16)
        assert lib.lig_snapshot_bytes(P, A) == want
    # layout: kv | q | n_active | max_active | bitmap, padding bits cleared
    P, A = 33, 2
    kv = np.arange(P, dtype=np.float64) / 100
    q = np.arange(P, dtype=np.int64) - 3
    na = np.full(P, 2, dtype=np.int64)
    ma = np.full(P, 4, dtype=np.int64)
    bitmap = np.full((A, 2), 0xFFFFFFFF, dtype=np.uint32)
    blob = pack_columns(kv, q, na, ma, bitmap).blob()
    Ppad = 64
    assert np.array_equal(blob[: P * 8].view(np.float64), kv)
    assert np.array_equal(blob[Ppad * 8: Ppad * 8 + P * 4].view(np.int32), q.astype(np.int32))
    off = Ppad * 12
    assert np.array_equal(blob[off: off + P * 2].view(np.uint16), na.astype(np.uint16))
    off = Ppad * 14
    assert np.array_equal(blob[off: off + P * 2].view(np.uint16), ma.astype(np.uint16))
    bm = blob[Ppad * 16: Ppad * 16 + A * 2 * 4].view(np.uint32).reshape(A, 2)
    assert np.array_equal(bm, np.array([[0xFFFFFFFF, 1], [0xFFFFFFFF, 1]], dtype=np.uint32))


def c_pack(lib, kv, q64, na64, ma64, bitmap):
    """The C helpers (lig_pack_pods + lig_pack_snapshot), the way a non-Python host uses them."""
    P = len(kv)
    kv = np.ascontiguousarray(kv, dtype=np.float64)
    q64, na64, ma64 = (np.ascontiguousarray(a, dtype=np.int64) for a in (q64, na64, ma64))
    q, na, ma = np.zeros(P, np.int32), np.zeros(P, np.uint16), np.zeros(P, np.uint16)
    p = lambda a: a.ctypes.data if a.size else None
    rc = lib.lig_pack_pods(P, p(q64), p(na64), p(ma64), p(q), p(na), p(ma))
    if rc != 0:
        return rc, None
    bitmap = np.ascontiguousarray(bitmap, dtype=np.uint32)
    A = bitmap.shape[0]
    out = np.zeros(lib.lig_snapshot_bytes(P, A), dtype=np.uint8)
    N.check(lib.lig_pack_snapshot(out.ctypes.data, P, A, p(kv), p(q), p(na), p(ma), p(bitmap)))
    return 0, out


def test_pack_pods_range_checks_and_saturation():
    """The numpy packer and the C helpers narrow identically: same bytes, same refusals."""
    lib = N.load()
    empty = np.zeros((0, 1), dtype=np.uint32)
    ok = pack_columns([0.1, 0.2, 0.3], [2**31 - 1, -(2**31), 0], [0, 1, 65534], [-7, 70000, 2**40], empty)
    assert ok.q.tolist() == [2**31 - 1, -(2**31), 0]
    assert ok.max_active.tolist() == [0, 65535, 65535]      # saturated, predicate-preserving
    rc, blob = c_pack(lib, [0.1, 0.2, 0.3], [2**31 - 1, -(2**31), 0], [0, 1, 65534], [-7, 70000, 2**40], empty)
    assert rc == 0 and np.array_equal(blob, ok.blob())
    for bad_q in (2**31, -(2**31) - 1):
        with pytest.raises(RangeError) as ei:
            pack_columns([0.0], [bad_q], [0], [0], empty)
        assert "int32" in str(ei.value)
        assert c_pack(lib, [0.0], [bad_q], [0], [0], empty)[0] == N.LIG_ERR_RANGE
        assert b"int32" in lib.lig_last_error()
    with pytest.raises(RangeError):
        pack_columns([0.0], [0], [65535], [0], empty)
    assert c_pack(lib, [0.0], [0], [65535], [0], empty)[0] == N.LIG_ERR_RANGE
    assert lib.lig_pack_pods(1, None, None, None, None, None, None) == N.LIG_ERR_INVALID
    # random pools, ragged P: identical blobs
    rng = np.random.default_rng(3)
    for P, A in [(1, 0), (31, 3), (32, 1), (33, 5), (100, 17), (257, 40)]:
        W = (P + 31) // 32
        kv = rng.random(P)
        q = rng.integers(-5, 300, P)
        na = rng.integers(0, 20, P)
        ma = rng.integers(-3, 70000, P)
        bm = rng.integers(0, 1 << 32, (A, W), dtype=np.uint64).astype(np.uint32)
        rc, blob = c_pack(lib, kv, q, na, ma, bm)
        assert rc == 0 and np.array_equal(blob, pack_columns(kv, q, na, ma, bm).blob()), (P, A)


def test_the_workload_generator_does_not_load_the_cuda_library():
    """bench.py --impl reference builds its inputs with workload.py / packer.py: pure numpy."""
    import subprocess
    import sys
    code = ("import sys; from llm_instance_gateway_b200 import workload as WL; "
            "s = WL.make_snapshot(64, 8); r = WL.make_requests(100, 8); s.packed.blob(); "
            "import ctypes; "
            "assert not any('liblig.so' in l for l in open('/proc/self/maps')), 'liblig.so was loaded'; "
            "assert 'llm_instance_gateway_b200._native' not in sys.modules")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_pack_pod_metrics_interning_and_bitmap():
    pods = [PodMetrics(Pod("p0", "a0"), Metrics(ActiveModels={"foo": 1, "bar": 1}, MaxActiveModels=2,
                                                 WaitingQueueSize=3, KVCacheUsagePercent=0.25)),
            PodMetrics(Pod("p1", "a1"), Metrics(ActiveModels={"bar": 1, "baz": 1}))]
    s = pack_pod_metrics(pods)
    assert (s.P, s.A, s.W) == (2, 3, 1)
    assert s.adapter_ids == {"foo": 0, "bar": 1, "baz": 2}
    assert s.bitmap[:, 0].tolist() == [0b01, 0b11, 0b10]
    assert s.adapter_id("nope") == s.A
    assert s.n_active.tolist() == [2, 2] and s.max_active.tolist() == [2, 0]
    assert s.pods == [Pod("p0", "a0"), Pod("p1", "a1")]
    assert s.algorithmic_snapshot_bytes() == 16 * 2 + 4 * 3 * 1


@pytest.mark.skipif(N.load().lig_device_count() > 0, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    lib = N.load()
    ctx = C.c_void_p()
    rc = lib.lig_create(C.byref(ctx), 0, 64, 32, 1024)
    assert rc == N.LIG_ERR_CUDA and not ctx.value
    assert b"no CPU path" in lib.lig_last_error()
    from llm_instance_gateway_b200.engine import Engine
    with pytest.raises(N.LigError):
        Engine(0, 64, 32, 1024)
    from llm_instance_gateway_b200.scheduling import NewScheduler

    class P:
        def AllPodMetrics(self):
            return []
    with pytest.raises(N.LigError):
        NewScheduler(P())


def test_create_argument_validation():
    lib = N.load()
    ctx = C.c_void_p()
    assert lib.lig_create(None, 0, 64, 32, 1024) == N.LIG_ERR_INVALID
    assert lib.lig_create(C.byref(ctx), 0, 0, 32, 1024) == N.LIG_ERR_INVALID
    assert lib.lig_create(C.byref(ctx), 0, N.LIG_MAX_PODS + 1, 32, 1024) == N.LIG_ERR_INVALID
    assert lib.lig_create(C.byref(ctx), 0, 64, N.LIG_MAX_ADAPTERS + 1, 1024) == N.LIG_ERR_INVALID
    assert lib.lig_create(C.byref(ctx), 0, 64, 32, 0) == N.LIG_ERR_INVALID
    assert b"max_batch" in lib.lig_last_error()


def test_product_does_not_touch_the_oracle():
    """The product tree must not import, link or reference anything under oracle/."""
    pkg = os.path.join(ROOT, "llm_instance_gateway_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".go", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_header_is_plain_c(tmp_path):
    """include/lig.h (the boundary a cgo shim includes) must compile as strict C99 on its own, and
    the record sizes the kernels assume must hold for the C compiler too."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "lig.h"\n'
                   '#include "lig_host_c.h"\n'
                   'typedef char req_is_16[(sizeof(lig_req) == 16) ? 1 : -1];\n'
                   'typedef char pick_is_8[(sizeof(lig_pick) == 8) ? 1 : -1];\n'
                   'typedef char thr_is_24[(sizeof(lig_thresholds) == 24) ? 1 : -1];\n'
                   'int main(void) { return LIG_OK + LIG_ERR_INVALID + (int)LIG_REQ_CRITICAL + LIG_ABI_VERSION; }\n')
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    out = subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                          "-I", os.path.join(ROOT, "include"),
                          "-I", os.path.join(ROOT, "llm_instance_gateway_b200", "csrc", "host"), str(src)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
