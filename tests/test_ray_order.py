"""Host logic of the multi-GPU ray order (api.cu map_order / RayOrder): every rank renders a contiguous block of the
rank-major order; the blocks of all ranks together cover the reference's frame-major batch exactly once, and every rank
gets the same number of rays of every frame.  Runs on CPU (the rule is exported through the C ABI for this purpose)."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("world,frames,pix", [(1, 5, 1000), (2, 5, 2000), (4, 5, 4000), (8, 5, 8000), (8, 3, 8), (2, 3, 7)])
def test_rank_major_order_is_a_balanced_permutation(nsb, world, frames, pix):
    L = nsb.load_library()
    n = frames * pix
    src = np.empty(n, np.int64); frm = np.empty(n, np.int64)
    f = C.c_int(0)
    for i in range(n):
        src[i] = L.nsb_ray_order_source(world, n, pix, i, C.byref(f)); frm[i] = f.value
    assert np.array_equal(np.sort(src), np.arange(n))                       # a permutation of the reference batch
    assert np.array_equal(frm, src // pix)                                  # the frame the kernel uses is the frame of that element
    if world > 1 and pix % world == 0:
        per = n // world
        for r in range(world):                                              # every rank: pix / world rays of every frame
            counts = np.bincount(frm[r * per:(r + 1) * per], minlength=frames)
            assert np.all(counts == pix // world)
    else:                                                                   # not divisible: contiguous slices of the reference order
        assert np.array_equal(src, np.arange(n))
