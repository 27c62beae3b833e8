"""NumPy emulation of the warp selection networks of csrc/nlk_search.cuh (warp_sort128,
warp_top32_of_128): element r*32 + lane lives in register r of lane `lane`, a stage with stride
< 32 is a shuffle-xor compare-exchange inside each register row, larger strides pair registers.
Checks that the truncated network (75 row stages) returns exactly the first 32 keys of the
full sort -- the property the k-NN parity of the GPU tests rests on when a launch keeps <= 32
candidates.  Keys are unique ((distance bits << 32) | scan index), as in the kernel."""
import numpy as np
import pytest

LANES = np.arange(32)


def row_stage(row, stride, asc_row):
    """warp_row_stage: asc_row may be a per-lane boolean array"""
    other = row[LANES ^ stride]
    lower = (LANES & stride) == 0
    take_min = lower == asc_row
    less = row < other
    return np.where(less == take_min, row, other)


def sort128(keys):
    key = keys.reshape(4, 32).copy()
    size = 2
    while size <= 128:
        stride = size >> 1
        while stride > 0:
            if stride >= 32:
                rs = stride >> 5
                for r in range(4):
                    if (r & rs) == 0:
                        asc = ((r << 5) & size) == 0
                        a, b = key[r].copy(), key[r + rs].copy()
                        sw = (a > b) == asc
                        key[r], key[r + rs] = np.where(sw, b, a), np.where(sw, a, b)
            else:
                for r in range(4):
                    asc = (((r << 5) | LANES) & size) == 0
                    key[r] = row_stage(key[r], stride, asc)
            stride >>= 1
        size <<= 1
    return key.reshape(-1)


def top32_of_128(keys):
    key = keys.reshape(4, 32).copy()
    size = 2
    while size <= 32:
        stride = size >> 1
        while stride > 0:
            for r in range(4):
                up = np.full(32, (r & 1) == 0) if size == 32 else (LANES & size) == 0
                key[r] = row_stage(key[r], stride, up)
            stride >>= 1
        size <<= 1
    key[0] = np.minimum(key[0], key[1])
    key[2] = np.minimum(key[2], key[3])
    for stride in (16, 8, 4, 2, 1):
        key[0] = row_stage(key[0], stride, np.full(32, True))
        key[2] = row_stage(key[2], stride, np.full(32, False))
    key[0] = np.minimum(key[0], key[2])
    for stride in (16, 8, 4, 2, 1):
        key[0] = row_stage(key[0], stride, np.full(32, True))
    return key[0]


def make_keys(rng, n, ties):
    d = rng.uniform(0, 1000, n).astype(np.float32)
    if ties:
        d = np.round(d / 250).astype(np.float32)      # a handful of distinct distances: many ties
    keys = (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    pad = np.full(128 - n, np.uint64(0xFFFFFFFFFFFFFFFF))
    return np.concatenate([keys, pad])


@pytest.mark.parametrize("n", [121, 128, 66, 33, 32, 7])
@pytest.mark.parametrize("ties", [False, True])
def test_networks_select_the_sorted_prefix(n, ties):
    rng = np.random.default_rng(100 * n + ties)
    for _ in range(20):
        keys = make_keys(rng, n, ties)
        want = np.sort(keys)
        assert np.array_equal(sort128(keys), want)
        assert np.array_equal(top32_of_128(keys), want[:32])
        # stable order: equal distances come out by scan index
        idx = (want[:min(n, 32)] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        dist = (want[:min(n, 32)] >> np.uint64(32)).astype(np.uint32).view(np.float32)
        same = dist[1:] == dist[:-1]
        assert np.all(idx[1:][same] > idx[:-1][same])
