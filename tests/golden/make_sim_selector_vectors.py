#!/usr/bin/env python3
"""Golden vectors for oracle/sim_selectors.py from the REFERENCE'S OWN method bodies.

The simulator cannot be imported here (simpy is absent), so the methods on the candidate-set path
are lifted out of the reference source with `ast` and compiled into stub classes:

    simulations/llm_ig_simulation/src/loadbalancer.py   LoadBalancer.get_pending_tokens_perc,
        get_lora_affinity, find_target_pod_based_on_min_pending, find_target_pod_based_on_min_kv_cache
    simulations/llm_ig_simulation/src/llmactor.py       LLMActor.get_num_tokens, get_num_tokens_in_decode,
        get_queue_size, get_min_expected_num_tokens_in_kvcache_after_prefill

`random.choice` is replaced by a recorder that returns the first candidate and keeps the list —
the candidate list is the deterministic part (the reference's choice itself is unseeded).
Run in the build container only (needs /root/reference); writes tests/golden/sim_selector_vectors.json.

    python tests/golden/make_sim_selector_vectors.py
"""
import ast
import json
import os
import types

import numpy as np

REF = "/root/reference/simulations/llm_ig_simulation/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift(path, cls, names, extra_globals):
    tree = ast.parse(open(path).read())
    cdef = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    funcs = [n for n in cdef.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {f.name for f in funcs} == set(names), (cls, names)
    new = ast.ClassDef(name=cls, bases=[], keywords=[], body=funcs, decorator_list=[])
    mod = ast.Module(body=[new], type_ignores=[])
    ast.fix_missing_locations(mod)
    g = dict(extra_globals)
    exec(compile(mod, path, "exec"), g)
    return g[cls]


class Store:
    def __init__(self, items):
        self.items = list(items)


class Wrapped:            # recompute_store holds PriorityItem(priority, item)
    def __init__(self, item):
        self.item = item


class Item:
    def __init__(self, i, o, r):
        self.input_size, self.output_size, self.output_size_remaining = i, o, r


class Recorder:
    def __init__(self):
        self.last = None

    def choice(self, seq):
        self.last = list(seq)
        return self.last[0]


def main():
    consts = {}
    exec(open(os.path.join(REF, "constants.py")).read(), consts)
    rec = Recorder()
    fake_random = types.SimpleNamespace(choice=rec.choice)
    Actor = lift(os.path.join(REF, "llmactor.py"), "LLMActor",
                 ["get_num_tokens", "get_num_tokens_in_decode", "get_queue_size",
                  "get_min_expected_num_tokens_in_kvcache_after_prefill"], {"np": np})
    LB = lift(os.path.join(REF, "loadbalancer.py"), "LoadBalancer",
              ["get_pending_tokens_perc", "get_lora_affinity", "find_target_pod_based_on_min_pending",
               "find_target_pod_based_on_min_kv_cache"],
              {"np": np, "random": fake_random, "List": list, "LLMActor": Actor,
               "MAX_GPU_MEMORY_PERC_BEFORE_RECOMPUTE": consts["MAX_GPU_MEMORY_PERC_BEFORE_RECOMPUTE"]})
    loras = list(consts["LORA_DICT"])
    rng = np.random.default_rng(20240921)
    cases = []
    for case in range(200):
        n_pods = 8 if case < 150 else int(rng.integers(1, 9))
        pods_json, actors = [], []
        for p in range(n_pods):
            def items(n):
                out = []
                for _ in range(n):
                    i, o = int(rng.integers(16, 2048)), int(rng.integers(1, 1024))
                    out.append([i, o, int(rng.integers(0, o + 1))])
                return out
            # ties on purpose: a third of the pods are idle, some share identical queues
            busy = rng.random() > 0.35
            dec = items(int(rng.integers(0, 6))) if busy else []
            pre = items(int(rng.integers(0, 4))) if busy else []
            rcp = items(int(rng.integers(0, 2))) if busy and rng.random() < 0.2 else []
            if case % 7 == 0 and p > 0 and rng.random() < 0.5:      # clone the previous pod's load
                dec, pre, rcp = [list(x) for x in pods_json[-1]["decode"]], [list(x) for x in pods_json[-1]["prefill"]], []
            lo = sorted(rng.choice(loras, size=int(rng.integers(0, 4)), replace=False).tolist())
            mx = consts["MAX_NUM_TOKENS_ALLOWED"] - sum(consts["LORA_DICT"][l] for l in lo)   # continous_batching.py:94-97
            pods_json.append({"lora_loaded": lo, "max_num_tokens_allowed": int(mx), "decode": dec, "prefill": pre, "recompute": rcp})
            a = Actor.__new__(Actor)
            a.lora_loaded = set(lo)
            a.max_num_tokens_allowed = int(mx)
            a.decode_store = Store(Item(*x) for x in dec)
            a.prefill_store = Store(Item(*x) for x in pre)
            a.recompute_store = Store(Wrapped(Item(*x)) for x in rcp)
            actors.append(a)
        lb = LB.__new__(LB)
        lb.list_of_llmactors = actors
        out = {"pods": pods_json, "pending_perc": [float(lb.get_pending_tokens_perc(a)) for a in actors],
               "expected_kv": [int(a.get_min_expected_num_tokens_in_kvcache_after_prefill()) for a in actors],
               "lora_affinity": {}, "min_pending": {}, "min_kv": {}}
        for lora in [""] + loras:
            aff = lb.get_lora_affinity(lora)
            idx = [actors.index(a) for a in aff]
            out["lora_affinity"][lora] = idx
            for safe in (False, True):
                rec.last = None
                got = lb.find_target_pod_based_on_min_pending(aff, safe)
                out["min_pending"][f"{lora}|{int(safe)}"] = {"candidates": rec.last if got is not None else [],
                                                           "returned_actor": actors.index(got) if got is not None else None}
            rec.last = None
            got = lb.find_target_pod_based_on_min_kv_cache(aff)
            out["min_kv"][lora] = {"candidates": rec.last if got is not None else [],
                                   "returned_actor": actors.index(got) if got is not None else None}
        cases.append(out)
    path = os.path.join(HERE, "sim_selector_vectors.json")
    with open(path, "w") as fh:
        json.dump({"source": "simulations/llm_ig_simulation/src/{loadbalancer,llmactor,constants}.py, methods lifted by ast",
                   "loras": loras, "cases": cases}, fh, separators=(",", ":"))
    print(path, len(cases), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
