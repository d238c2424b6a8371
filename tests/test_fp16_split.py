"""Numerics of the tensor-core operand representation (pde_surrogate_b200/csrc/conv_tc.cuh), restated in
numpy: x * 2^s = h1 + h2 with h1 = fp16(x * 2^s), h2 = fp16(x * 2^s - h1), and a product a*w evaluated as
a1*w1 + a1*w2 + a2*w1.  No GPU: this pins the error model DESIGN.md quotes for the kernels."""
import numpy as np


def split2(x, s):
    y = (np.asarray(x, np.float32) * np.float32(2.0 ** s)).astype(np.float32)
    h1 = y.astype(np.float16)
    h2 = (y - h1.astype(np.float32)).astype(np.float16)
    return h1, h2


def test_two_piece_representation_error():
    rs = np.random.RandomState(0)
    for s, mag in ((4, 1.0), (4, 1e-3), (8, 0.05), (8, 1e-4), (0, 100.0)):
        x = (rs.standard_normal(200000) * mag).astype(np.float32)
        h1, h2 = split2(x, s)
        rec = (h1.astype(np.float64) + h2.astype(np.float64)) / 2.0 ** s
        err = np.abs(rec - x.astype(np.float64))
        bound = np.maximum(2.0 ** -22 * np.abs(x), 2.0 ** -24 / 2.0 ** s)  # 2x the quoted model: fp32 rounding of y
        assert np.all(err <= bound), (s, mag, float((err / bound).max()))
        # the leading piece alone is an fp16 rounding: 2^-11 relative in the normal range
        big = np.abs(x) * 2.0 ** s > 2.0 ** -13
        lead = np.abs(h1.astype(np.float64) / 2.0 ** s - x)[big]
        assert np.all(lead <= 2.0 ** -11 * np.abs(x[big]) * (1 + 1e-6))


def test_three_products_match_fp32_dot():
    """K = 9 * 128 products (a 3x3 dense layer): the three-product evaluation with fp32 accumulation of each
    class is as accurate as an fp32 dot product (both ~1e-7 relative to the fp64 value)."""
    rs = np.random.RandomState(1)
    K, n = 9 * 128, 400
    a = np.maximum(rs.standard_normal((n, K)), 0).astype(np.float32)            # post-ReLU activations
    w = (rs.standard_normal((n, K)) / np.sqrt(K)).astype(np.float32)            # filter row
    a1, a2 = split2(a, 4)
    w1, w2 = split2(w, 8)
    f = lambda t: t.astype(np.float64)
    lead = (f(a1) * f(w1)).sum(1)
    cross = (f(a1) * f(w2)).sum(1) + (f(a2) * f(w1)).sum(1)
    got = (lead.astype(np.float32) + cross.astype(np.float32)).astype(np.float64) / 2.0 ** 12
    ref = (f(a) * f(w)).sum(1)
    fp32 = np.einsum("ij,ij->i", a, w).astype(np.float64)
    scale = np.abs(f(a) * f(w)).sum(1)
    e_split = np.abs(got - ref) / scale
    e_fp32 = np.abs(fp32 - ref) / scale
    assert e_split.max() < 3e-7, e_split.max()
    assert np.median(e_split) < 4 * max(np.median(e_fp32), 2e-8)
    # dropping the cross terms is what a single fp16 product would give: three orders worse
    e_lead = np.abs(lead / 2.0 ** 12 - ref) / scale
    assert np.median(e_lead) > 100 * np.median(e_split)


def test_dynamic_gradient_scale_rule():
    """Gradients: the power of two that brings the running |G| maximum to [2^10, 2^11) keeps every element
    below the fp16 overflow threshold with 5 binades of headroom and resolves elements 2^14 below the
    maximum to fp32-level absolute accuracy."""
    for gmax in (3e-7, 1e-3, 0.7, 5e4):
        bits = np.float32(gmax).view(np.uint32)
        e = 10 - (int((bits >> 23) & 0xFF) - 127)
        scaled_max = gmax * 2.0 ** e
        assert 2.0 ** 10 <= scaled_max < 2.0 ** 11
        assert scaled_max * 32 < 65504
        x = np.float32(gmax * 2.0 ** -14 * 0.731)
        h1, h2 = split2(x, e)
        rec = (float(h1) + float(h2)) / 2.0 ** e
        assert abs(rec - float(x)) <= 2.0 ** -24 * gmax  # fp32-level relative to the tensor's scale
