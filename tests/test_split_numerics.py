"""CPU check of the precision claim behind the decoder kernels (common.cuh): an fp32 operand carried as
hi = fp16(x), lo = fp16(x - hi) and a product evaluated as a_lo.b_hi + a_hi.b_lo + a_hi.b_hi with wide accumulation is
fp32-grade (~2^-22 per product), while a single fp16 product (precision = 1) is not.  numpy emulation; no GPU needed."""
import numpy as np


def split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def mm3(a, w):
    ah, al = split(a); wh, wl = split(w)
    return al @ wh.T + ah @ wl.T + ah @ wh.T


def test_three_product_fp16_split_is_fp32_grade():
    rs = np.random.RandomState(0)
    for scale_a, scale_w, K in ((1.0, 0.3, 96), (0.01, 0.2, 32), (30.0, 0.4, 64)):
        a = (rs.standard_normal((512, K)) * scale_a).astype(np.float32)
        w = (rs.uniform(-1, 1, (32, K)) * scale_w).astype(np.float32)
        ref = a.astype(np.float64) @ w.astype(np.float64).T
        norm = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64).T        # sum |a_k w_k|: the natural error scale
        err3 = np.abs(mm3(a, w) - ref).max() / norm.max()
        err1 = np.abs(a.astype(np.float16).astype(np.float64) @ w.astype(np.float16).astype(np.float64).T - ref).max() / norm.max()
        assert err3 < 2e-6, (scale_a, scale_w, K, err3)       # 2^-22 = 2.4e-7 per product, plus the 3e-8 floor of the lo part (below)
        assert err1 > 20 * err3                               # one fp16 product: ~2^-11 per operand
    # operands far below the fp16 normal range (the fine grid is initialised at 1e-4): the error is the ABSOLUTE floor, i.e. tiny
    # against the O(0.1) decoder outputs such terms are added to, although large relative to the terms themselves
    a = (rs.standard_normal((512, 32)) * 1e-4).astype(np.float32)
    w = (rs.uniform(-1, 1, (32, 32)) * 0.3).astype(np.float32)
    assert np.abs(mm3(a, w) - a.astype(np.float64) @ w.astype(np.float64).T).max() < 32 * 0.3 * 2.0 ** -25


def test_split_error_bound():
    """|hi + lo - x| <= max(2^-22 |x|, 2^-25): 22 significant bits while lo is a normal fp16 (|x| >= 0.125), and the absolute
    half-spacing of the fp16 subnormal grid (2^-25 = 3e-8) below that."""
    rs = np.random.RandomState(1)
    x = np.concatenate([rs.uniform(-8, 8, 20000), rs.uniform(-0.125, 0.125, 20000), rs.uniform(-6e-5, 6e-5, 10000)]).astype(np.float32)
    hi, lo = split(x)
    err = np.abs(hi + lo - x.astype(np.float64))
    assert np.all(err <= np.maximum(2.0 ** -22 * np.abs(x), 2.0 ** -25) * 1.0001)
