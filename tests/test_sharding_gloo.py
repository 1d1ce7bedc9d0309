"""World-size-2 gloo tests of the N>1 path on CPU.

The GPU kernels cannot run here, so each rank schedules its shard with the CPU oracle; what is
under test is the multi-rank LOGIC the library and bench.py implement on GPUs: contiguous request
shards, one broadcast of the packed snapshot, counter-based keys of the model-request path
(first_index = shard offset), and the windowed load feedback with one all-reduce of the per-pod
pick histogram per window.  Concatenated per-rank results must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.workload import shard_bounds

P, A, R, S_FB = 100, 9, 3001, 256


def test_shard_bounds_cover_and_order():
    for total in (0, 1, 7, 1024, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as oracle
    # --- the only exchange step of the default path: the packed snapshot, root -> everyone ---
    ref_blob = WL.make_snapshot(P, A, seed=21).packed.blob()
    blob = torch.from_numpy(ref_blob.copy()) if rank == 0 else torch.zeros(len(ref_blob), dtype=torch.uint8)
    dist.broadcast(blob, src=0)
    assert np.array_equal(blob.numpy(), ref_blob)
    snap = WL.make_snapshot(P, A, seed=21)
    pk = snap.packed
    reqs = WL.make_requests(R, A, seed=22)
    lo, hi = shard_bounds(R, rank, world)
    pool = oracle.Pool(snap.pod_records())
    picks, _ = pool.schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, np.ascontiguousarray(reqs[lo:hi]), 99)
    np.save(os.path.join(out_dir, f"picks_{rank}.npy"), picks)
    # --- model requests: rand_key = first_index + i, the shard passes its offset ---
    models = WL.make_models(A)
    ids = WL.make_model_requests(R, A, seed=23)
    mp_ = oracle.Models(WL.model_records(models)).schedule_batch(pool, np.ascontiguousarray(ids[lo:hi]), 7, first_index=lo)
    np.save(os.path.join(out_dir, f"mpicks_{rank}.npy"), mp_)
    # --- load feedback: per window, schedule the local slice, all-reduce the pick histogram, fold it
    #     into the queue sizes (lig_schedule_batch_feedback_device + ncclAllReduce on GPUs) ---
    q = snap.q64.copy()
    out = np.zeros(hi - lo, dtype=oracle.PICK_DTYPE)
    n_windows = max((b - a + S_FB - 1) // S_FB for a, b in (shard_bounds(R, r, world) for r in range(world)))
    for w in range(n_windows):
        tab = oracle.ClassTable(pk.P, pk.A, pk.kv, q.astype(np.int32), pk.n_active, pk.max_active, pk.bitmap)
        a, b = lo + w * S_FB, min(lo + (w + 1) * S_FB, hi)
        hist = torch.zeros(P, dtype=torch.int64)
        if a < b:
            pw = tab.schedule_batch(np.ascontiguousarray(reqs[a:b]), 99)
            out[a - lo:b - lo] = pw
            hist += torch.from_numpy(np.bincount(pw["pod_idx"][pw["pod_idx"] >= 0], minlength=P))
        dist.all_reduce(hist)                              # every rank takes part in every window
        q = np.minimum(q + hist.numpy(), (1 << 31) - 1)
    np.save(os.path.join(out_dir, f"fb_{rank}.npy"), out)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)               # timings are reported as the max over ranks
    assert float(t) == float(world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_equals_whole(tmp_path, oracle):
    from oracle import feedback as FB
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    snap = WL.make_snapshot(P, A, seed=21)
    pk = snap.packed
    reqs = WL.make_requests(R, A, seed=22)
    pool = oracle.Pool(snap.pod_records())
    whole, _ = pool.schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, 99)
    cat = lambda stem: np.concatenate([np.load(os.path.join(str(tmp_path), f"{stem}_{r}.npy")) for r in range(world)])
    assert np.array_equal(cat("picks"), whole)
    ids = WL.make_model_requests(R, A, seed=23)
    assert np.array_equal(cat("mpicks"), oracle.Models(WL.model_records(WL.make_models(A))).schedule_batch(pool, ids, 7, 0))
    shards = [shard_bounds(R, r, world) for r in range(world)]
    want_fb, _, _ = FB.schedule_batch_feedback(pk.P, pk.A, pk.kv, snap.q64, pk.n_active, pk.max_active, pk.bitmap, reqs, 99,
                                               S_FB, shards)
    assert np.array_equal(cat("fb"), want_fb)
