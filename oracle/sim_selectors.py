"""Restatement of the simulator's candidate-set selectors.  TEST INFRASTRUCTURE ONLY.

north_star asks for picks "bit-exact against the reference Go scheduler and
simulations/llm_ig_simulation".  The simulator is a different algorithm family from the Go filter
tree, has no tests, and every routing function ends in an UNSEEDED
random.choice / random.randint — its picks cannot be pinned.  What is deterministic, and restated
here as pure functions over a per-pod snapshot, are the CANDIDATE SETS the random choice is taken
from (paths relative to simulations/llm_ig_simulation/src/):

    lora_affinity          loadbalancer.py:130-139   get_lora_affinity
    min_pending_candidates loadbalancer.py:236-268   find_target_pod_based_on_min_pending
    min_kv_candidates      loadbalancer.py:271-295   find_target_pod_based_on_min_kv_cache
    pending_tokens_perc    loadbalancer.py:104-110   get_pending_tokens_perc
    expected_kv_after_prefill  llmactor.py:63-73     get_min_expected_num_tokens_in_kvcache_after_prefill

Pinned against the reference's own method bodies: tests/golden/make_sim_selector_vectors.py
extracts those methods from the reference source with `ast` (the module itself cannot be imported:
simpy is absent) and runs them on seeded C1-shaped states (8 pods, the 4 LoRA adapters of
constants.py:21); tests/test_sim_selectors.py compares this restatement with the committed
vectors.  "Bit-exact" here means: identical candidate index lists (order included) and identical
float64 ratios.

A quirk kept as is: both find_target_pod_* functions draw an index INTO THE `pods` ARGUMENT but
return self.list_of_llmactors[index] (loadbalancer.py:265-266, :292-293) — when `pods` is a
filtered subset (the LoRA-affinity pods) the returned actor is the one at that position of the
FULL list.  `resolve_quirk` reproduces the mapping.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Set

MAX_NUM_TOKENS_ALLOWED = 2810 * 16 - 512          # constants.py:11-17
MAX_GPU_MEMORY_PERC_BEFORE_RECOMPUTE = 0.9        # constants.py:18
LORA_DICT = {"tweet": 1600, "sql": 1600, "dummy-1": 0, "dummy-2": 0}   # constants.py:21


@dataclass
class Item:                      # request.py: the fields the selectors read
    input_size: int
    output_size: int
    output_size_remaining: int


@dataclass
class PodState:                  # llmactor.py:8-19, a frozen view of one LLMActor
    lora_loaded: Set[str] = field(default_factory=set)
    max_num_tokens_allowed: int = MAX_NUM_TOKENS_ALLOWED
    prefill: List[Item] = field(default_factory=list)       # prefill_store.items
    decode: List[Item] = field(default_factory=list)        # decode_store.items
    recompute: List[Item] = field(default_factory=list)     # recompute_store.items[*].item


def num_tokens_in_decode(p: PodState) -> int:                               # llmactor.py:21-29
    return sum(x.input_size + x.output_size - x.output_size_remaining for x in p.decode)


def expected_kv_after_prefill(p: PodState) -> int:                          # llmactor.py:63-73
    n = num_tokens_in_decode(p)
    if p.recompute:
        it = p.recompute[0]
        return n + it.input_size + it.output_size - it.output_size_remaining
    if p.prefill:
        it = p.prefill[0]
        return n + it.input_size + it.output_size - it.output_size_remaining
    return n


def pending_tokens_perc(p: PodState) -> float:                              # loadbalancer.py:104-110
    pending = sum(x.output_size + x.input_size for x in p.decode) + sum(x.output_size + x.input_size for x in p.prefill)
    return pending / p.max_num_tokens_allowed


def lora_affinity(pods: Sequence[PodState], lora_requested: str) -> List[int]:     # loadbalancer.py:130-139
    if not lora_requested:
        return list(range(len(pods)))
    have = [i for i, p in enumerate(pods) if lora_requested in p.lora_loaded]
    if have:
        return have
    fewest = min(len(p.lora_loaded) for p in pods)
    return [i for i, p in enumerate(pods) if len(p.lora_loaded) == fewest]


def min_pending_candidates(pods: Sequence[PodState], eviction_safe: bool = False,
                           max_kv_perc: float = MAX_GPU_MEMORY_PERC_BEFORE_RECOMPUTE) -> List[int]:
    """Indices INTO `pods`.  loadbalancer.py:236-261 — note the elif of the eviction_safe branch
    appends ties without re-checking the KV condition, as the reference does."""
    cand: List[int] = []
    best = float("inf")
    for i, p in enumerate(pods):
        pend = pending_tokens_perc(p)
        kv = expected_kv_after_prefill(p) / (p.max_num_tokens_allowed + 0.0)
        if eviction_safe:
            if pend < best and kv < max_kv_perc:
                cand, best = [i], pend
            elif pend == best:
                cand.append(i)
        else:
            if pend < best:
                cand, best = [i], pend
            elif pend == best:
                cand.append(i)
    return cand


def min_kv_candidates(pods: Sequence[PodState]) -> List[int]:               # loadbalancer.py:271-288
    cand: List[int] = []
    best = float("inf")
    for i, p in enumerate(pods):
        kv = expected_kv_after_prefill(p) / (p.max_num_tokens_allowed + 0.0)
        if kv < best:
            best, cand = kv, [i]
        elif kv == best:
            cand.append(i)
    return cand


def resolve_quirk(index_into_pods: int) -> int:
    """The actor the reference returns for a drawn candidate index: list_of_llmactors[index]
    (loadbalancer.py:265-266, :292-293), i.e. the index is NOT mapped back through `pods`."""
    return index_into_pods
