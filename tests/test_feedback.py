"""In-batch load feedback (opt-in): oracle/feedback.py is its sequential
definition; the CPU test checks the two inner schedulers of the oracle agree, the GPU tests check
lig_schedule_batch_feedback_device against it bit for bit and that the default path is untouched."""
import numpy as np
import pytest

from llm_instance_gateway_b200 import workload as WL
from oracle import feedback as FB


def test_oracle_feedback_classtab_equals_port(oracle):
    c = WL.CONFIGS["C2"]
    snap = WL.make_snapshot(c["P"], c["A"], seed=9)
    p = snap.packed
    reqs = WL.make_requests(3000, c["A"], seed=10)
    for shards in (None, [(0, 1000), (1000, 3000)]):
        a, ta, _ = FB.schedule_batch_feedback(p.P, p.A, p.kv, snap.q64, p.n_active, p.max_active, p.bitmap, reqs, 5, 256, shards)
        b, tb = FB.schedule_batch_feedback_port(snap.pod_records(), snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, 5, 256, shards)
        assert np.array_equal(a, b) and np.array_equal(ta, tb)
    # it changes the outcome (that is the point) and conserves the picks
    plain, _ = oracle.Pool(snap.pod_records()).schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, 5)
    assert not np.array_equal(plain["pod_idx"], a["pod_idx"])
    assert ta.sum() == (a["pod_idx"] >= 0).sum()
    assert ta.max() < np.bincount(plain["pod_idx"][plain["pod_idx"] >= 0], minlength=p.P).max()   # the herd is spread


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,R,S", [("C2", 5000, 512), ("C3", 70_001, 4096), ("C5", 20_000, 1000)])
def test_gpu_feedback_matches_the_sequential_definition(cfg, R, S, oracle):
    import torch
    from llm_instance_gateway_b200.engine import Engine
    from llm_instance_gateway_b200.packer import PICK_DTYPE
    c = WL.CONFIGS[cfg]
    snap = WL.make_snapshot(c["P"], c["A"], seed=21)
    p = snap.packed
    reqs = WL.make_requests(R, c["A"], seed=22)
    want, want_total, _ = FB.schedule_batch_feedback(p.P, p.A, p.kv, snap.q64, p.n_active, p.max_active, p.bitmap, reqs, 77, S)
    with Engine(0, max_pods=c["P"], max_adapters=c["A"], max_batch=R) as e:
        e.upload_snapshot(1, p)
        plain_before = e.schedule_batch(1, 77, reqs)
        d_reqs = torch.from_numpy(reqs.view(np.uint8).reshape(-1)).cuda()
        d_out = torch.zeros(R * 8, dtype=torch.uint8, device="cuda")
        d_hist = torch.full((c["P"],), -1, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()   # the fills / copy ran on torch's default stream
        stream = torch.cuda.Stream()
        for rep in range(2):                      # the scratch copy is rebuilt from the resident epoch every call
            with torch.cuda.stream(stream):
                e.schedule_batch_feedback_device(1, 77, d_reqs.data_ptr(), R, d_out.data_ptr(), S, 0, d_hist.data_ptr(),
                                                 stream.cuda_stream)
            stream.synchronize()
            got = d_out.cpu().numpy().view(PICK_DTYPE)
            for f in ("status", "n_survivors", "pod_idx"):
                bad = np.nonzero(got[f] != want[f])[0]
                assert bad.size == 0, (rep, f, bad[:5], got[bad[:5]], want[bad[:5]])
            assert np.array_equal(d_hist.cpu().numpy().astype(np.int64), want_total)
        # more windows than needed (trailing empty windows, the multi-rank case) change nothing
        with torch.cuda.stream(stream):
            e.schedule_batch_feedback_device(1, 77, d_reqs.data_ptr(), R, d_out.data_ptr(), S, (R + S - 1) // S + 3, 0,
                                             stream.cuda_stream)
        stream.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(PICK_DTYPE), want)
        # the resident epoch was not modified: the default path still gives the reference's answer
        assert np.array_equal(e.schedule_batch(1, 77, reqs), plain_before)
        ref, _ = oracle.Pool(snap.pod_records()).schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, 77, False,
                                                                oracle.hardware_threads())
        assert np.array_equal(plain_before, ref)
