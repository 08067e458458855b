"""Index arithmetic of the one-step plan kernel (csrc/train.cu plan_small_kernel), restated in numpy and checked against a
stable argsort.  The kernel itself is compared bit-for-bit with the multi-kernel sort on the GPU
(test_gpu_train.py::test_single_kernel_plan_equals_multi_kernel_plan); this model covers the geometries no GPU test
reaches — every digit width (5..8 bits), slice sizes 1,024 / 2,048 / 4,096, n at the capacity limit and ragged tails."""
import numpy as np
import pytest

CTAS, WARPS, ITEMS = 8, 32, 4                     # PS_CTAS, PS_WARPS, PS_ITEMS
MAX_N = CTAS * WARPS * 32 * ITEMS


def bits_for(n):
    b = 1
    while (1 << b) < n:
        b += 1
    return b


def model_sort(keys, key_space):
    """Two-pass LSD sort exactly as the kernel lays it out: warp g = cta * 32 + w owns entries [g*chunk, (g+1)*chunk),
    rows of 32 lanes, warp-private counter columns, bases = digit-major, then CTA, then warp order."""
    n = len(keys)
    kb = bits_for(key_space)
    assert n <= MAX_N and kb <= 16
    ib, db = bits_for(n), max(5, (kb + 1) // 2)
    ndig, dmask = 1 << db, (1 << db) - 1
    chunk = 32
    while chunk * CTAS * WARPS < n:
        chunk *= 2
    assert chunk <= 32 * ITEMS
    sl = chunk * WARPS                            # slice: positions owned by one CTA
    assert sl & (sl - 1) == 0
    x = (keys.astype(np.int64) << ib) | np.arange(n)
    assert x.max() < 0xffffffff
    src = x
    for shift in (ib, ib + db):
        dig = (src >> shift) & dmask
        cnt = np.zeros((CTAS, WARPS, ndig), np.int64)
        for g in range(CTAS * WARPS):
            lo, hi = g * chunk, min(n, (g + 1) * chunk)
            if lo < hi:
                np.add.at(cnt[g // WARPS, g % WARPS], dig[lo:hi], 1)
        tot = cnt.sum(axis=1)                                            # [cta][digit]
        wpre = np.cumsum(cnt, axis=1) - cnt                              # entries of the digit in lower warps of the CTA
        total = tot.sum(axis=0)
        excl = np.cumsum(total) - total                                  # block scan over digits
        lower = np.cumsum(tot, axis=0) - tot                             # same digit in lower CTAs
        dst = np.full(CTAS * sl, -1, np.int64)                           # the cluster-distributed array
        for g in range(CTAS * WARPS):
            c, w = g // WARPS, g % WARPS
            cur = wpre[c, w].copy()
            for row in range(g * chunk, min(n, (g + 1) * chunk), 32):    # one row of 32 lanes at a time
                for i in range(row, min(n, row + 32, (g + 1) * chunk)):  # lane order inside the row = rank order
                    d = dig[i]
                    pos = excl[d] + lower[c, d] + cur[d]
                    cur[d] += 1
                    assert dst[(pos >> bits_for(sl)) * sl + (pos & (sl - 1))] == -1
                    dst[pos] = src[i]
        assert (dst[:n] >= 0).all() and (dst[n:] == -1).all()
        src = dst[:n]
    return (src >> ib).astype(np.int64), (src & ((1 << ib) - 1)).astype(np.int64)


@pytest.mark.parametrize("n,key_space", [(1, 3), (31, 68), (33, 68), (1600, 68), (4000, 524), (8192, 2000), (8193, 4801),
                                         (10668, 4801), (16385, 16297), (28986, 16297), (5656, 40962), (24577, 30000),
                                         (MAX_N, 65536), (MAX_N - 1, 1025)])
def test_plan_small_layout_is_a_stable_sort(n, key_space):
    rng = np.random.default_rng(n + key_space)
    keys = rng.integers(0, key_space, n)
    if n > 100:                                   # a hub row (one relation hit by a third of the batch) and the sentinel key
        keys[rng.integers(0, n, n // 3)] = key_space - 2
        keys[rng.integers(0, n, n // 50)] = key_space - 1
    skeys, perm = model_sort(keys, key_space)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(perm, order)
    assert np.array_equal(skeys, keys[order])
