"""The schedule of the large MSM (csrc/msm_large.cu) through the C-ABI — host logic, no GPU: window rule of the reference
(src/msm/mod.rs:112-116, 322-325), signed-digit windows that always hold the biased scalar, bucket-reduction levels."""
import ctypes as C

import pytest


def plan(n, bits):
    import lasso_b200 as lb

    out = (C.c_int * 16)()
    assert lb.lib().lasso_msm_plan_info(C.c_size_t(n), C.c_uint(bits), out) == 0
    return dict(c=out[0], nw=out[1], nbits=out[2], NB=out[3], S=out[4], nlev=out[5], L=[out[6 + k] for k in range(out[5])])


@pytest.mark.parametrize("log_n", [14, 16, 18, 20, 22, 24, 26])
@pytest.mark.parametrize("bits", [1, 16, 20, 60, 128, 252, 253])
def test_plan_invariants(log_n, bits):
    p = plan(1 << log_n, bits)
    ref_c = int(log_n * 0.69) + 2  # ln_without_floats(n) + 2 = floor(log2(n) * 69 / 100) + 2 for n >= 32
    assert p["c"] == min(max(min(ref_c, 17), 8), max(2, bits + 1))
    assert p["c"] * p["nw"] >= bits + 2          # s + bias < 2^(c * nw): the top field never overflows
    assert p["c"] * (p["nw"] - 1) < bits + 2     # and no window is wasted
    assert p["c"] * p["nw"] <= 287               # the biased scalar fits 9 limbs
    assert p["NB"] == 1 << (p["c"] - 1)
    prod = 1
    for L in p["L"]:
        assert 1 < L <= 16
        prod *= L
    assert prod == p["NB"]                       # the levels reduce the buckets of a window to one group
    assert p["S"] >= 64 and p["S"] >= 4 * ((1 << log_n) // p["NB"])


def test_signed_digits_reconstruct_the_scalar():
    """the offset trick of the kernels, restated: field w of s + sum_w 2^(c-1) 2^(cw), minus 2^(c-1)"""
    import random

    L = 2**252 + 27742317777372353535851937790883648493
    rnd = random.Random(5)
    for log_n, bits in [(16, 253), (22, 253), (26, 253), (22, 16), (20, 20), (18, 60)]:
        p = plan(1 << log_n, bits)
        c, nw = p["c"], p["nw"]
        bias = sum(1 << (w * c + c - 1) for w in range(nw))
        for _ in range(300):
            s = rnd.randrange(0, min(L, 1 << bits))
            v = s + bias
            assert v < 1 << (c * nw)
            digits = [((v >> (w * c)) & ((1 << c) - 1)) - (1 << (c - 1)) for w in range(nw)]
            assert all(-(1 << (c - 1)) <= d < (1 << (c - 1)) for d in digits)
            assert sum(d << (w * c) for w, d in enumerate(digits)) == s
