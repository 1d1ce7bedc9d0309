"""N > 1 on real GPUs (skipped unless the box has >= 2), driven through the C ABI ALONE — no
torch.distributed anywhere:

  * lig_group_*: ONE process owns every GPU (the reference's wiring: one scheduler per ext-proc
    process, pkg/ext-proc/main.go:137).  The snapshot is packed once, copied to member 0 and
    replicated by one in-library ncclBroadcast straight into every member's resident slot; the
    request batch shards contiguously; the picks come back in request order and must equal the
    oracle's result for the whole batch bit for bit.
  * lig_comm_*: one process per GPU (the torchrun shape bench.py uses); rank 0's unique id travels
    through a file; the snapshot broadcast happens inside the library on the rank's own stream.
"""
import os
import time

import numpy as np
import pytest
import torch

from llm_instance_gateway_b200 import _native as N
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.engine import Engine, EngineGroup
from llm_instance_gateway_b200.packer import PICK_DTYPE

pytestmark = pytest.mark.gpu

P, A, R, SEED = 700, 48, 200_003, 321
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")


def _oracle_whole_batch(oracle, snap, reqs, seed):
    want, _ = oracle.Pool(snap.pod_records()).schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, seed,
                                                             False, oracle.hardware_threads())
    return want


@needs2
def test_group_one_process_all_gpus_matches_oracle(oracle):
    world = min(torch.cuda.device_count(), 8)
    snap = WL.make_snapshot(P, A, seed=61)
    snap2 = WL.make_snapshot(P, A, seed=63)
    reqs = WL.make_requests(R, A, seed=62)
    want = _oracle_whole_batch(oracle, snap, reqs, SEED)
    with EngineGroup(list(range(world)), max_pods=1024, max_adapters=64, max_batch=R) as grp:
        assert grp.size == world
        grp.upload_snapshot(5, snap.packed)
        got = grp.schedule_batch(5, SEED, reqs)                     # pageable buffers: bounce path
        assert np.array_equal(got, want)
        # page-locked caller buffers: every member reads its shard over PCIe in place
        lib = N.load()
        h_in = lib.lig_host_alloc(R * 16)
        h_out = lib.lig_host_alloc(R * 8)
        try:
            np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (R * 16)).from_address(h_in))[:] = reqs.view(np.uint8).reshape(-1)
            grp.schedule_batch_ptr(5, SEED, h_in, R, h_out)
            got2 = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (R * 8)).from_address(h_out)).view(PICK_DTYPE).copy()
            assert np.array_equal(got2, want)
        finally:
            lib.lig_host_free(h_in)
            lib.lig_host_free(h_out)
        # every member holds the same tables (the broadcast reached it): same single-device result
        for i in range(world):
            m = grp.member(i)
            assert np.array_equal(m.schedule_batch(5, SEED, np.ascontiguousarray(reqs[:5000])), want[:5000]), i
        # a second epoch, thresholds through the group, ragged tiny batches
        grp.upload_snapshot(6, snap2.packed)
        assert np.array_equal(grp.schedule_batch(6, 9, reqs), _oracle_whole_batch(oracle, snap2, reqs, 9))
        assert np.array_equal(grp.schedule_batch(5, SEED, reqs), want)          # older epoch still resident
        for n in (1, world - 1, world, world + 1, 1000):
            sub = np.ascontiguousarray(reqs[:n])
            assert np.array_equal(grp.schedule_batch(6, 3, sub), _oracle_whole_batch(oracle, snap2, sub, 3)), n
        grp.set_thresholds(0.5, 2, 10)
        oracle.set_thresholds(0.5, 2, 10)
        try:
            assert np.array_equal(grp.schedule_batch(6, 4, reqs), _oracle_whole_batch(oracle, snap2, reqs, 4))
        finally:
            oracle.set_thresholds()


def _comm_worker(rank, world, id_path, out_dir):
    torch.cuda.set_device(rank)
    eng = Engine(rank, max_pods=1024, max_adapters=64, max_batch=R)
    if rank == 0:
        uid = eng.comm_unique_id()
        with open(id_path + ".tmp", "wb") as fh:
            fh.write(uid)
        os.replace(id_path + ".tmp", id_path)
    else:
        t0 = time.time()
        while not os.path.exists(id_path):
            if time.time() - t0 > 120:
                raise RuntimeError("no unique id from rank 0")
            time.sleep(0.01)
        uid = open(id_path, "rb").read()
    eng.comm_init_rank(world, rank, uid)
    snap = WL.make_snapshot(P, A, seed=61) if rank == 0 else None
    # blocking host-array form: the arrays exist on the root only
    eng.comm_upload_snapshot(5, P, A, snap.packed if snap else None, root=0)
    reqs = WL.make_requests(R, A, seed=62)
    lo, hi = R * rank // world, R * (rank + 1) // world
    got = eng.schedule_batch(5, SEED, np.ascontiguousarray(reqs[lo:hi]))
    # device form on the caller's stream: root passes a device blob, nothing is synchronised
    stream = torch.cuda.Stream()
    snap2 = WL.make_snapshot(P, A, seed=63) if rank == 0 else None
    blob = torch.from_numpy(snap2.packed.blob()).cuda() if rank == 0 else None
    d_reqs = torch.from_numpy(np.ascontiguousarray(reqs[lo:hi]).view(np.uint8).reshape(-1)).cuda()
    d_out = torch.zeros((hi - lo) * 8, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        eng.comm_upload_snapshot_device(6, P, A, blob.data_ptr() if rank == 0 else 0, 0, stream.cuda_stream)
        eng.schedule_batch_device(6, 9, d_reqs.data_ptr(), hi - lo, d_out.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    # the per-pod histogram all-reduce of the load-feedback mode
    hist = torch.full((P,), rank + 1, dtype=torch.int32, device="cuda")
    eng.comm_allreduce_i32(hist.data_ptr(), P, 0)
    torch.cuda.synchronize()
    assert int(hist[0]) == world * (world + 1) // 2
    # load feedback across ranks: every window's picks are all-reduced inside the library
    S_FB = 8192
    n_windows = max((R * (r + 1) // world - R * r // world + S_FB - 1) // S_FB for r in range(world))
    d_fb = torch.zeros((hi - lo) * 8, dtype=torch.uint8, device="cuda")
    d_hist = torch.zeros(P, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        eng.schedule_batch_feedback_device(5, SEED, d_reqs.data_ptr(), hi - lo, d_fb.data_ptr(), S_FB, n_windows,
                                           d_hist.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    np.save(os.path.join(out_dir, f"fb_{rank}.npy"), d_fb.cpu().numpy().view(PICK_DTYPE))
    np.save(os.path.join(out_dir, f"fbhist_{rank}.npy"), d_hist.cpu().numpy())
    np.save(os.path.join(out_dir, f"picks5_{rank}.npy"), got)
    np.save(os.path.join(out_dir, f"picks6_{rank}.npy"), d_out.cpu().numpy().view(PICK_DTYPE))
    eng.close()


@needs2
def test_comm_one_process_per_gpu_matches_oracle(tmp_path, oracle):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_comm_worker, args=(world, str(tmp_path / "nccl_id"), str(tmp_path)), nprocs=world, join=True)
    reqs = WL.make_requests(R, A, seed=62)
    for epoch, snap_seed, seed in ((5, 61, SEED), (6, 63, 9)):
        snap = WL.make_snapshot(P, A, seed=snap_seed)
        got = np.concatenate([np.load(os.path.join(str(tmp_path), f"picks{epoch}_{r}.npy")) for r in range(world)])
        assert np.array_equal(got, _oracle_whole_batch(oracle, snap, reqs, seed)), epoch
    # the feedback mode over the ranks == its sequential definition with the same shards
    from oracle import feedback as FB
    snap = WL.make_snapshot(P, A, seed=61)
    pk = snap.packed
    shards = [(R * r // world, R * (r + 1) // world) for r in range(world)]
    want, want_total, _ = FB.schedule_batch_feedback(pk.P, pk.A, pk.kv, snap.q64, pk.n_active, pk.max_active, pk.bitmap,
                                                     reqs, SEED, 8192, shards)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), f"fb_{r}.npy")) for r in range(world)])
    assert np.array_equal(got, want)
    for r in range(world):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), f"fbhist_{r}.npy")).astype(np.int64), want_total), r


def test_group_of_one_needs_no_nccl(oracle):
    snap = WL.make_snapshot(300, 16, seed=7)
    reqs = WL.make_requests(10_000, 16, seed=8)
    with EngineGroup([0], max_pods=512, max_adapters=16, max_batch=10_000) as grp:
        grp.upload_snapshot(1, snap.packed)
        assert np.array_equal(grp.schedule_batch(1, 2, reqs), _oracle_whole_batch(oracle, snap, reqs, 2))
