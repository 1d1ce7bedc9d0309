"""N > 1 on real GPUs (skipped unless the box has >= 2): one process per GPU over NCCL, the snapshot
replicated by one broadcast of the packed blob and consumed in place by
lig_upload_snapshot_device, each rank scheduling its contiguous shard through the C ABI; the
concatenation must equal the oracle's result for the whole batch (same check as the gloo test,
with the CUDA path doing the scheduling)."""
import os
import socket

import numpy as np
import pytest
import torch

from llm_instance_gateway_b200 import workload as WL

pytestmark = pytest.mark.gpu

P, A, R, SEED = 700, 48, 200_003, 321


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from llm_instance_gateway_b200 import _native as N
    from llm_instance_gateway_b200.engine import Engine
    from llm_instance_gateway_b200.packer import PICK_DTYPE
    from llm_instance_gateway_b200.sharding import broadcast_snapshot, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    nbytes = N.load().lig_snapshot_bytes(P, A)
    if rank == 0:
        blob = torch.from_numpy(WL.make_snapshot(P, A, seed=61).packed.blob()).to(dev)
    else:
        blob = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    broadcast_snapshot(blob, src=0)                       # the only exchange step
    eng = Engine(rank, max_pods=1024, max_adapters=64, max_batch=R)
    stream = torch.cuda.current_stream()
    eng.upload_snapshot_device(5, P, A, blob.data_ptr(), stream.cuda_stream)
    reqs = WL.make_requests(R, A, seed=62)
    lo, hi = shard_bounds(R, rank, world)
    d_reqs = torch.from_numpy(np.ascontiguousarray(reqs[lo:hi]).view(np.uint8).reshape(-1)).to(dev)
    d_out = torch.zeros((hi - lo) * 8, dtype=torch.uint8, device=dev)
    eng.schedule_batch_device(5, SEED, d_reqs.data_ptr(), hi - lo, d_out.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"picks_{rank}.npy"), d_out.cpu().numpy().view(PICK_DTYPE))
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_request_sharded_nccl_matches_whole_batch(tmp_path, oracle):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    snap = WL.make_snapshot(P, A, seed=61)
    reqs = WL.make_requests(R, A, seed=62)
    want, _ = oracle.Pool(snap.pod_records()).schedule_batch(snap.adapter_names(), WL.UNKNOWN_MODEL, reqs, SEED,
                                                             False, oracle.hardware_threads())
    got = np.concatenate([np.load(os.path.join(str(tmp_path), f"picks_{r}.npy")) for r in range(world)])
    assert np.array_equal(got, want)
