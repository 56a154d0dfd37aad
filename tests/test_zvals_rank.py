"""Host-side restatement of k_zvals' rank placement (csrc/ray_kernels.cuh; Renderer.cpp:119 sorts the concatenation of the 32 stratified
and the 16 near-surface z values): both runs ascend, so the stable rank of a value is its index plus the number of values of the other
run that precede it, found by a 6-step binary search over 32 lanes.  Checked against numpy's stable argsort, ties included."""
import numpy as np


def _count(run, n, x, pred):
    """the kernel's search: largest c with pred(run[c - 1], x), for an ascending run of n <= 32 values held one per lane"""
    c, step = 0, 32
    while step > 0:
        lane = (c + step - 1) & 31                      # the shuffle's source lane (wraps; gated by the bound below)
        if c + step <= n and pred(run[lane], x):
            c += step
        step >>= 1
    return c


def test_rank_placement_equals_stable_sort():
    rng = np.random.default_rng(0)
    for trial in range(400):
        ns = 16 if trial % 4 == 0 else int(rng.integers(1, 17))
        hi = 6 if trial % 3 == 0 else 1000              # small range: many ties inside and between the runs
        v0 = np.sort(rng.integers(0, hi, 32)).astype(np.float32)
        v1 = np.zeros(32, np.float32)
        v1[:ns] = np.sort(rng.integers(0, hi, ns)).astype(np.float32)
        cat = np.concatenate([v0, v1[:ns]])
        order = np.argsort(cat, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(cat.size)
        out = np.full(cat.size, np.nan, np.float32)
        for l in range(32):
            r0 = l + _count(v1, ns, v0[l], lambda e, x: e < x)
            assert r0 == rank[l]
            out[r0] = v0[l]
            if l < ns:
                r1 = l + _count(v0, 32, v1[l], lambda e, x: e <= x)
                assert r1 == rank[32 + l]
                out[r1] = v1[l]
        assert np.array_equal(out, np.sort(cat, kind="stable"))
