"""Randomised shapes against the oracle on a B200 (run once per round under gpurun; not a pytest).

N random snapshots with P in [1, 4096] and A in [0, 1024] (shapes drawn to hit both class-build
kernels' word counts, ragged tails and the smem-table limit), adversarial metric values (NaN, inf,
signed zero, threshold boundaries, int32 extremes), random thresholds; every (critical, adapter)
class is requested several times.  Each snapshot is scheduled through the host-buffer call, the
device queue kernel and — after a random delta — lig_update_snapshot, and every pick is compared
with the class-table oracle (which the CPU tests pin to the structure-preserving port).
    python scripts/fuzz_gpu.py [n_snapshots] [seed]
"""
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from llm_instance_gateway_b200 import workload as WL
from llm_instance_gateway_b200.engine import Engine
from llm_instance_gateway_b200.packer import REQ_DTYPE, PICK_DTYPE, pack_columns
from llm_instance_gateway_b200.backend import Pod
from oracle import binding as oracle

KV_SPECIAL = [0.0, -0.0, 0.8, 0.8000000000000002, 0.7999999999999999, 1.0, math.inf, -math.inf, math.nan,
              1e-300, 5e-324, -0.25, 1.0 / 3.0]
Q_SPECIAL = [0, 5, 6, 49, 50, -1, 2**31 - 1, -(2**31)]


def random_snapshot(rng, P, A):
    q = rng.integers(0, 70, P).astype(np.int64)
    sp = rng.random(P) < 0.1
    q[sp] = rng.choice(Q_SPECIAL, size=int(sp.sum()))
    kv = np.round(rng.random(P), int(rng.integers(1, 4)))
    sp = rng.random(P) < 0.1
    kv[sp] = rng.choice(KV_SPECIAL, size=int(sp.sum()))
    ma = rng.integers(0, 6, P).astype(np.int64)
    W = (P + 31) // 32
    bitmap = np.zeros((A, W), dtype=np.uint32)
    na = np.zeros(P, dtype=np.int64)
    if A > 0:
        density = rng.choice([0.0, 0.002, 0.02, 0.3])
        m = rng.random((A, P)) < density
        hot = rng.integers(0, A, size=min(A, 3))        # a few adapters present on most pods
        m[hot] = rng.random((len(hot), P)) < 0.9
        na = m.sum(axis=0).astype(np.int64)
        pad = np.zeros((A, W * 32), dtype=bool)
        pad[:, :P] = m
        bitmap = np.packbits(pad.reshape(A, W, 32), axis=2, bitorder="little").view(np.uint32).reshape(A, W)
    na = np.minimum(na, 65535)
    ids = {WL.adapter_name(a): a for a in range(A)}
    return pack_columns(kv, q, na, ma, np.ascontiguousarray(bitmap), ids, [Pod(f"p{i}", f"a{i}") for i in range(P)])


def all_class_requests(rng, A, reps):
    ids = np.tile(np.arange(-1, A + 2, dtype=np.int32), 2 * reps)
    reqs = np.zeros(len(ids), dtype=REQ_DTYPE)
    reqs["adapter_id"] = ids
    reqs["flags"] = np.repeat(np.arange(2 * reps) % 2, A + 3).astype(np.uint32)
    reqs["rand_key"] = rng.integers(0, 2**63, len(ids), dtype=np.uint64)
    pad = (-len(reqs)) % 4
    if pad:
        reqs = np.concatenate([reqs, reqs[:pad]])
    return np.ascontiguousarray(reqs)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    rng = np.random.default_rng(seed)
    eng = Engine(0, 4096, 1024, 1 << 16)
    stream = torch.cuda.Stream()
    bad = 0
    t0 = time.time()
    shapes = []
    for it in range(n):
        P = int(rng.choice([rng.integers(1, 70), rng.integers(70, 1100), rng.integers(1100, 4097), 4096, 1024, 33]))
        A = int(rng.choice([0, 1, rng.integers(2, 40), rng.integers(40, 1025), 1024]))
        thr = (float(rng.choice([0.8, 0.5, 0.0, 1.0])), int(rng.choice([5, 0, 1, 50])), int(rng.choice([50, 5, 0, 200])))
        shapes.append((P, A))
        pk = random_snapshot(rng, P, A)
        eng.set_thresholds(*thr)
        ep = 10 + 2 * it
        eng.upload_snapshot(ep, pk)
        reqs = all_class_requests(rng, A, reps=int(rng.integers(1, 4)))
        tab = oracle.ClassTable(pk.P, pk.A, pk.kv, pk.q, pk.n_active, pk.max_active, pk.bitmap, thresholds=thr)
        want = tab.schedule_batch(reqs, it)
        got_host = eng.schedule_batch(ep, it, reqs)
        d_req = torch.from_numpy(reqs.view(np.uint8).reshape(-1)).cuda()
        d_out = [torch.zeros(len(reqs) * 8, dtype=torch.uint8, device="cuda") for _ in range(3)]
        torch.cuda.synchronize()               # the fills above ran on torch's default stream
        with torch.cuda.stream(stream):
            eng.schedule_batches_device(ep, it, [d_req.data_ptr()] * 3, len(reqs), [o.data_ptr() for o in d_out],
                                        stream.cuda_stream)
        stream.synchronize()
        ok = np.array_equal(got_host, want)
        why = [] if ok else [f"host path: {int((got_host != want).sum())}/{len(want)} differ"]
        for b, o in enumerate(d_out):          # batch b of a queue draws with seed + b
            wb = want if b == 0 else tab.schedule_batch(reqs, it + b)
            gb = o.cpu().numpy().view(PICK_DTYPE)
            if not np.array_equal(gb, wb):
                ok = False
                i = int(np.nonzero(gb != wb)[0][0])
                why.append(f"queue batch {b}: {int((gb != wb).sum())}/{len(wb)} differ, first at {i}: got {gb[i]} want {wb[i]} req {reqs[i]}")
        # a random delta on top: dirty pods get fresh rows and adapter sets
        nd = int(rng.integers(1, max(2, P // 8)))
        dirty = np.sort(rng.choice(P, size=min(nd, P), replace=False)).astype(np.int32)
        pk2 = random_snapshot(rng, P, A)
        kv, q, na, ma = pk.kv.copy(), pk.q.copy(), pk.n_active.copy(), pk.max_active.copy()
        bm = pk.bitmap.copy()
        kv[dirty], q[dirty], na[dirty], ma[dirty] = pk2.kv[dirty], pk2.q[dirty], pk2.n_active[dirty], pk2.max_active[dirty]
        for p in dirty:
            w, bit = p >> 5, np.uint32(1 << (p & 31))
            if A:
                bm[:, w] = (bm[:, w] & ~bit) | (pk2.bitmap[:, w] & bit)
        rows = np.ascontiguousarray(pk2.bitmap[:, :].T) if A else np.zeros((0,), np.uint32)
        # per dirty pod: the list of adapters it now serves
        offs, ads = [0], []
        for p in dirty:
            mine = np.nonzero((pk2.bitmap[:, p >> 5] >> np.uint32(p & 31)) & 1)[0] if A else np.zeros(0, np.int64)
            ads.extend(int(a) for a in mine)
            offs.append(len(ads))
        eng.update_snapshot(ep + 1, ep, dirty, kv[dirty], q[dirty], na[dirty], ma[dirty],
                            np.asarray(offs, dtype=np.int32), np.asarray(ads, dtype=np.int32))
        tab2 = oracle.ClassTable(P, A, kv, q, na, ma, np.ascontiguousarray(bm), thresholds=thr)
        ok2 = np.array_equal(eng.schedule_batch(ep + 1, it, reqs), tab2.schedule_batch(reqs, it))
        if not (ok and ok2):
            bad += 1
            print(f"MISMATCH #{it}: P={P} A={A} thr={thr} R={len(reqs)} full={ok} delta={ok2} {why} kernel={eng.pick_kernel_info(ep)}", flush=True)
    eng.close()
    print(f"fuzz: {n} snapshots (P x A from {min(shapes)} to {max(shapes)}), {bad} mismatches, {time.time() - t0:.1f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
