"""CPU check of the division-free Int31n the pick kernel uses (lig_device.cuh: make_entry /
int31n_magic): for every survivor count n the device can see (2..32768) the magic quotient
umulhi(v, M) >> shift equals floor(v / n) for all v < 2^31, and the rejection test
q >= floor((2^31-1)/n) + pow2 equals Go's v > 2^31-1-(2^31 % n)."""
import numpy as np


def entry(n):
    l = int(n - 1).bit_length()          # ceil(log2 n)
    shift = l - 1
    magic = ((1 << (32 + shift)) + n - 1) // n
    assert magic < (1 << 32)
    pow2 = int(n & (n - 1) == 0)
    return shift, magic, pow2


def test_magic_division_exact_for_all_n():
    rng = np.random.default_rng(0)
    base = np.concatenate([rng.integers(0, 1 << 31, 4000, dtype=np.uint64),
                           np.array([0, 1, 2, (1 << 31) - 1, (1 << 31) - 2, 1 << 30, (1 << 30) + 1], dtype=np.uint64)])
    for n in range(2, 32769):
        shift, magic, pow2 = entry(n)
        # values around multiples of n near both ends of the range, plus the random base
        top = ((1 << 31) // n) * n
        edge = np.array([n - 1, n, n + 1, top - 1, top, min(top + 1, (1 << 31) - 1), top - n, top - n - 1],
                        dtype=np.uint64)
        v = np.concatenate([base, edge]) if n % 97 == 0 or n < 300 or n > 32700 else edge
        v = v[v < np.uint64(1 << 31)]          # Int31() yields 31-bit values
        q = ((v * np.uint64(magic)) >> np.uint64(32)) >> np.uint64(shift)
        assert np.array_equal(q, v // np.uint64(n)), n
        q_limit = ((((1 << 31) - 1) * magic) >> 32 >> shift) + pow2
        assert q_limit == (1 << 31) // n, n
        go_reject = v > np.uint64((1 << 31) - 1 - ((1 << 31) % n))
        if pow2:
            assert not go_reject.any()
        assert np.array_equal(q >= np.uint64(q_limit), go_reject), n


def test_single_survivor_entry_never_rejects():
    # make_entry(n = 1): shift 0, magic 0, pow2 bit set -> q = 0 < q_limit = 1 for every draw
    shift, magic, pow2 = 0, 0, 1
    q_limit = ((((1 << 31) - 1) * magic) >> 32 >> shift) + pow2
    for v in (0, 1, (1 << 31) - 1):
        q = (v * magic) >> 32 >> shift
        assert q < q_limit


def test_uint64_products_do_not_overflow():
    # v < 2^31 and M < 2^32, so v * M < 2^63: the numpy check above is exact
    assert ((1 << 31) - 1) * ((1 << 32) - 1) < (1 << 63)
