"""Host-side logic of the window-partitioned multi-GPU MSM (no GPU): every window
is owned by exactly one rank, and the signed-digit decomposition the kernels use
reconstructs the scalar for every window width."""
import random

from oracle import bls12381 as b


def test_partition_covers_every_window_once(pkg):
    for n in (1 << 10, 1 << 16, 1 << 20, 1 << 22):
        for world in (1, 2, 4, 8):
            owned = []
            W = None
            for rank in range(world):
                first, step, nwin, mine = pkg.comm_partition(n, world, rank)
                W = nwin
                assert first == rank and step == world
                owned.extend(range(first, nwin, step))
                assert mine == len(range(first, nwin, step))
            assert sorted(owned) == list(range(W))


def signed_digits(k, c):
    """Restatement of k_big_digits: W = ceil(128 / c) digits in [-2^(c-1), 2^(c-1)] of a GLV half < 2^127."""
    W = (128 + c - 1) // c
    M = 1 << (c - 1)
    out, carry = [], 0
    for w in range(W):
        t = ((k >> (w * c)) & ((1 << c) - 1)) + carry
        carry = 1 if t > M else 0
        out.append(t - 2 * M if carry else t)
    assert carry == 0
    return out


def test_signed_digit_decomposition_reconstructs():
    random.seed(3)
    ks = [0, 1, 2**127 - 1, 2**126, 2**127 - 2**64] + [random.randrange(2**127) for _ in range(200)]
    for c in range(2, 19):
        M = 1 << (c - 1)
        for k in ks:
            d = signed_digits(k, c)
            assert all(-M <= x <= M for x in d)
            assert sum(x << (c * w) for w, x in enumerate(d)) == k


def test_partition_rejects_bad_rank_and_world(pkg):
    """cdl_comm_partition with world < 1 or rank outside [0, world) reports zero windows instead of
    dividing by zero."""
    for world, rank in ((0, 0), (-1, 0), (4, 4), (4, -1)):
        assert pkg.comm_partition(1 << 16, world, rank) == (0, 0, 0, 0)


def test_sorted_entry_count_bound():
    """The large-MSM path indexes its sorted (point, sign) entries with 32 bits: 2 * n * windows-per-rank
    must stay below 2^32 (big_msm_on_device returns CDL_ERR_TOO_LARGE otherwise).  At c = 16 a single
    rank owns 8 windows, so the bound bites above 2^28 terms, well inside the accepted n < 2^30."""
    c, W = 16, 8
    assert 2 * (1 << 28) * W == 1 << 32
    assert 2 * ((1 << 28) - 1) * W < 1 << 32
