"""World-size-2 gloo test of the N>1 path on CPU: snapshot broadcast + request sharding.

The GPU kernels cannot run here, so each rank schedules its shard with the CPU oracle; what is
under test is the multi-rank plumbing bench.py uses (shard bounds, snapshot replication,
max-over-ranks reduction): concatenated per-rank picks must equal the unsharded result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.sharding import broadcast_snapshot, max_over_ranks, shard_bounds


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
    P, A, R = 100, 9, 3001
    nbytes = None
    if rank == 0:
        snap = WL.make_snapshot(P, A, seed=21)
        blob = torch.from_numpy(snap.packed.blob())
    else:
        from llm_instance_gateway_b200 import _native as N
        blob = torch.zeros(N.load().lig_snapshot_bytes(P, A), dtype=torch.uint8)
    broadcast_snapshot(blob, src=0)
    # every rank must now hold rank 0's snapshot bit for bit
    ref = torch.from_numpy(WL.make_snapshot(P, A, seed=21).packed.blob())
    assert torch.equal(blob, ref)
    # rebuild the pool from the replicated blob's source of truth and schedule this rank's shard
    snap = WL.make_snapshot(P, A, seed=21)
    reqs = WL.make_requests(R, A, seed=22)
    lo, hi = shard_bounds(R, rank, world)
    picks, _ = oracle.Pool(snap.pod_records()).schedule_batch(
        snap.adapter_names(), WL.UNKNOWN_MODEL, np.ascontiguousarray(reqs[lo:hi]), 99)
    np.save(os.path.join(out_dir, f"picks_{rank}.npy"), picks)
    t = max_over_ranks(float(rank + 1), torch.device("cpu"))
    assert t == float(world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_equals_whole(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    P, A, R = 100, 9, 3001
    snap = WL.make_snapshot(P, A, seed=21)
    reqs = WL.make_requests(R, A, seed=22)
    whole, _ = oracle.Pool(snap.pod_records()).schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, 99)
    parts = [np.load(os.path.join(str(tmp_path), f"picks_{r}.npy")) for r in range(world)]
    assert np.array_equal(np.concatenate(parts), whole)
