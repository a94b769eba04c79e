"""Every known-answer test the reference holds for the hot path,
restated against the oracle.  file:line = /root/reference/src/..."""
import numpy as np
import pytest

import oracle_lib as ol
import pyref
from oracle_lib import L_FR, P, fr_array, fr_ints, lib, sz


def s(v):  # small signed ints -> residues
    return [x % L_FR for x in v]


def test_eq_evals_and_evaluate_kat():
    # poly/dense_mlpoly.rs:436-458: Z=[1,2,1,4], r=(4,3) -> evaluate = 28 ; Appendix D1
    r = fr_array([4, 3])
    ev = np.zeros((4, 4), dtype=np.uint64)
    lib().orc_eq_evals(P(r), sz(2), P(ev))
    assert fr_ints(ev) == s([6, -9, -8, 12])
    Z = fr_array([1, 2, 1, 4])
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_evaluate(P(Z), sz(4), P(r), P(out))
    assert fr_ints(out) == [28]
    # constant-8 poly evaluates to 8 at (3,4)  (dense_mlpoly.rs:628-648)
    Z8 = fr_array([8, 8, 8, 8])
    lib().orc_evaluate(P(Z8), sz(4), P(fr_array([3, 4])), P(out))
    assert fr_ints(out) == [8]


def test_bind_top_bot_kat():
    # Appendix D1
    Z = fr_array([1, 2, 1, 4])
    lib().orc_bind(1, P(Z), sz(4), P(fr_array([4])))
    assert fr_ints(Z[:2]) == [1, 10]
    z2 = Z[:2].copy()
    lib().orc_bind(1, P(z2), sz(2), P(fr_array([3])))
    assert fr_ints(z2[:1]) == [28]
    Z = fr_array([1, 2, 1, 4])
    lib().orc_bind(0, P(Z), sz(4), P(fr_array([3])))
    assert fr_ints(Z[:2]) == [4, 10]
    z2 = Z[:2].copy()
    lib().orc_bind(0, P(z2), sz(2), P(fr_array([4])))
    assert fr_ints(z2[:1]) == [28]


@pytest.mark.parametrize("ell", [1, 2, 5, 8])
def test_eq_evals_vs_naive_bit_order(ell):
    # dense_mlpoly.rs:529-583: evals() == naive bitwise product, r[0] <-> MSB; factored == outer product
    rng = np.random.default_rng(ell)
    r_int = [int.from_bytes(rng.bytes(40), "little") % L_FR for _ in range(ell)]
    ev = np.zeros((1 << ell, 4), dtype=np.uint64)
    lib().orc_eq_evals(P(fr_array(r_int)), sz(ell), P(ev))
    naive = pyref.eq_evals(r_int)
    assert fr_ints(ev) == naive
    left = ell // 2
    Lv, Rv = pyref.eq_evals(r_int[:left]), pyref.eq_evals(r_int[left:])
    assert [a * b % L_FR for a in Lv for b in Rv] == naive


def test_unipoly_kats():
    # unipoly.rs:129-157: evals [1,6,15] -> coeffs [1,3,2], eval(3) = 28
    c = np.zeros((3, 4), dtype=np.uint64)
    lib().orc_unipoly_from_evals(P(fr_array([1, 6, 15])), sz(3), P(c))
    assert fr_ints(c) == [1, 3, 2]
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_unipoly_evaluate(P(c), sz(3), P(fr_array([3])), P(out))
    assert fr_ints(out) == [28]
    # unipoly.rs:160-189: evals [1,7,23,55] -> [1,3,2,1], eval(4) = 109
    c = np.zeros((4, 4), dtype=np.uint64)
    lib().orc_unipoly_from_evals(P(fr_array([1, 7, 23, 55])), sz(4), P(c))
    assert fr_ints(c) == [1, 3, 2, 1]
    lib().orc_unipoly_evaluate(P(c), sz(4), P(fr_array([4])), P(out))
    assert fr_ints(out) == [109]
    # random degree-9 (the LT C=8 round polynomial shape) against independent Lagrange interpolation
    rng = np.random.default_rng(1)
    ev = [int.from_bytes(rng.bytes(40), "little") % L_FR for _ in range(10)]
    c = np.zeros((10, 4), dtype=np.uint64)
    lib().orc_unipoly_from_evals(P(fr_array(ev)), sz(10), P(c))
    assert fr_ints(c) == pyref.interpolate(ev)


def test_gaussian_elimination_kat():
    # utils/gaussian_elimination.rs:73-81: the reference's own augmented matrix -> [2, 12, 3]
    aug = fr_array([1, 0, 0, 2, 1, 1, 1, 17, 1, 2, 4, 38])
    out = np.zeros((3, 4), dtype=np.uint64)
    lib().orc_gaussian_elimination(P(aug), sz(3), P(out))
    assert fr_ints(out) == [2, 12, 3]


def test_sumcheck_prove_arbitrary_kat():
    # sumcheck.rs:459-513: A=B=C=[8..15], g = product, scripted challenges [3,1,3]; Appendix D2
    vals = list(range(8, 16))
    polys = fr_array(vals * 3)
    revals = np.zeros((3, 4, 4), dtype=np.uint64)
    comp = np.zeros((3, 3, 4), dtype=np.uint64)
    fin = np.zeros((3, 4), dtype=np.uint64)
    rc = lib().orc_sumcheck_product_kat(P(polys), sz(3), sz(8), P(fr_array([3, 1, 3])), P(revals), P(comp), P(fin), None)
    assert rc == 0  # verifier's e == A(r)B(r)C(r)
    assert fr_ints(revals.reshape(-1, 4)) == [3572, 10044, 21700, 40076, 17261, 22815, 29449, 37259,
                                              10648, 12167, 13824, 15625]
    assert fr_ints(comp.reshape(-1, 4)) == [3572, 1824, 256, 17261, 492, 16, 10648, 66, 1]
    assert fr_ints(fin) == [25, 25, 25]
    assert sum(v**3 for v in vals) == 13616 == 3572 + 10044


def test_grand_product_kat():
    # grand_product.rs:270-283: tree of [1,2,3,4] = 24, GP argument prove -> verify ; Appendix D3
    out = np.zeros(4, dtype=np.uint64)
    assert lib().orc_grand_product_kat(P(fr_array([1, 2, 3, 4])), sz(4), P(out)) == 0
    assert fr_ints(out) == [24]
    rng = np.random.default_rng(2)
    v = [int.from_bytes(rng.bytes(40), "little") % L_FR for _ in range(64)]
    assert lib().orc_grand_product_kat(P(fr_array(v)), sz(64), P(out)) == 0
    prod = 1
    for x in v:
        prod = prod * x % L_FR
    assert fr_ints(out) == [prod]


def materialize(kind, C, log_m, log_r=0):
    nsub = lib().orc_num_subtables(kind, sz(C), sz(log_m), sz(log_r))
    out = np.zeros((nsub, 1 << log_m, 4), dtype=np.uint64)
    lib().orc_materialize_subtables(kind, sz(C), sz(log_m), sz(log_r), P(out))
    return [fr_ints(out[k]) for k in range(nsub)]


def test_subtable_materialization_kats():
    # and.rs:69-92, or.rs, xor.rs:69-93 (M=16, first 11 entries)
    assert materialize(0, 4, 4)[0][:11] == [0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 2]
    assert materialize(1, 4, 4)[0][:11] == [0, 1, 2, 3, 1, 1, 3, 3, 2, 3, 2]
    assert materialize(2, 4, 4)[0][:11] == [0, 1, 2, 3, 1, 0, 3, 2, 2, 3, 0]
    # lt.rs:113-139 (M=16): LT and EQ tables
    lt, eq = materialize(3, 4, 4)
    for idx in range(16):
        assert lt[idx] == int((idx >> 2) < (idx & 3)) and eq[idx] == int((idx >> 2) == (idx & 3))
    # range_check.rs:101-128 (M=2^16, LOG_R=40): full = identity, remainder cut off at 2^8, zeros
    full, rem, zeros = materialize(4, 3, 16, 40)
    assert full[:4] == [0, 1, 2, 3] and full[65535] == 65535
    assert rem[255] == 255 and rem[256] == 0 and rem[65535] == 0 and zeros == [0] * 65536


def test_combine_lookups_kats():
    out = np.zeros(4, dtype=np.uint64)
    vals = fr_array([100, 200, 300, 400])
    expected = 100 + (1 << 8) * 200 + (1 << 16) * 300 + (1 << 24) * 400  # and.rs:94-110 (same in or/xor)
    for kind in (0, 1, 2):
        lib().orc_combine_lookups(kind, sz(4), sz(16), sz(0), P(vals), P(out))
        assert fr_ints(out) == [expected]
    # lt.rs:86-111: T = LT0 + LT1*EQ0 + LT2*EQ0*EQ1 + ...
    lib().orc_combine_lookups(3, sz(4), sz(4), sz(0), P(fr_array([10, 1, 20, 0, 30, 1, 40, 1])), P(out))
    assert fr_ints(out) == [30]  # the reference's vector: 10 + 20*1 + 30*1*0 + 40*1*0*1
    v = [3, 5, 7, 11, 13, 17, 19, 23]  # LT0,EQ0,LT1,EQ1,...
    lib().orc_combine_lookups(3, sz(4), sz(16), sz(0), P(fr_array(v)), P(out))
    assert fr_ints(out) == [3 + 7 * 5 + 13 * 5 * 11 + 19 * 5 * 11 * 17]
    # range_check.rs:78-86: weights 2^(i*log M)
    lib().orc_combine_lookups(4, sz(3), sz(8), sz(20), P(fr_array([1, 2, 3])), P(out))
    assert fr_ints(out) == [1 + (2 << 8) + (3 << 16)]


@pytest.mark.parametrize("kind,C,log_m,log_r", [(0, 1, 4, 0), (1, 2, 4, 0), (2, 2, 4, 0), (3, 2, 4, 0), (4, 3, 4, 10)])
def test_materialization_mle_parity(kind, C, log_m, log_r):
    # subtables/test.rs:15-39 materialization_mle_parity_test!: table == MLE on the whole hypercube
    tabs = materialize(kind, C, log_m, log_r)
    out = np.zeros(4, dtype=np.uint64)
    for k, tab in enumerate(tabs):
        for idx in range(1 << log_m):
            point = fr_array([(idx >> (log_m - 1 - b)) & 1 for b in range(log_m)])
            lib().orc_evaluate_subtable_mle(kind, sz(C), sz(log_m), sz(log_r), sz(k), P(point), sz(log_m), P(out))
            assert fr_ints(out) == [tab[idx]]


def test_valid_merged_poly():
    # and.rs:112-137 / xor.rs / or.rs: merged lookup polys evaluate to the table entries at Boolean points
    nz = np.array([[0, 2], [5, 9]], dtype=np.uint64)  # C=2 dims x s=2
    for kind, exp in ((0, [0, 0, 1, 0]), (1, [0, 2, 1, 3]), (2, [0, 2, 0, 3])):
        E = np.zeros((2, 2, 4), dtype=np.uint64)
        lib().orc_lookup_polys(kind, sz(2), sz(4), sz(0), P(nz), sz(2), P(E))
        assert fr_ints(E.reshape(-1, 4)) == exp


def test_split_bits_and_densify_fixture():
    # memory_checking.rs:794-831 fixture + Appendix D4: accesses [1,2,1,5], m=8
    idx = np.array([[1], [2], [1], [5]], dtype=np.uint64)
    dim = np.zeros(4, dtype=np.uint64)
    rd = np.zeros(4, dtype=np.uint64)
    fin = np.zeros(8, dtype=np.uint64)
    lib().orc_densify(P(idx), sz(4), sz(1), sz(3), P(dim), P(rd), P(fin))
    assert dim.tolist() == [1, 2, 1, 5] and rd.tolist() == [0, 0, 1, 0] and fin.tolist() == [0, 2, 1, 0, 0, 1, 0, 0]
    table = fr_array(list(range(10, 18)))
    out = np.zeros((2 * 8 + 2 * 4, 4), dtype=np.uint64)
    lib().orc_gp_fingerprints(P(table), sz(8), P(dim), P(rd), P(fin), sz(4), P(fr_array([100])), P(fr_array([200])), P(out))
    f = fr_ints(out)
    assert f[:8] == [800, 901, 1002, 1103, 1204, 1305, 1406, 1507]
    assert f[8:16] == [800, 20901, 11002, 1103, 1204, 11305, 1406, 1507]
    assert f[16:20] == [901, 1002, 10901, 1305] and f[20:24] == [10901, 11002, 20901, 11305]
    prod = lambda v: __import__("functools").reduce(lambda a, b: a * b % L_FR, v, 1)
    assert prod(f[:8]) * prod(f[20:24]) % L_FR == prod(f[16:20]) * prod(f[8:16]) % L_FR
    # densify pads with address 0 (densified.rs:37): 3 lookups -> s=4, pad hits address 0
    idx = np.array([[1], [2], [1]], dtype=np.uint64)
    lib().orc_densify(P(idx), sz(3), sz(1), sz(3), P(dim), P(rd), P(fin))
    assert dim.tolist() == [1, 2, 1, 0] and rd.tolist() == [0, 0, 1, 0] and fin.tolist() == [1, 2, 1, 0, 0, 0, 0, 0]


def test_polynomial_commit_roundtrip_kat():
    """poly/dense_mlpoly.rs:586-624 check_polynomial_commit: Z = [1, 2, 1, 4], r = [4, 3] -> eval 28; commit with
    PolyCommitmentGens::new(2, b"test-two"), PolyEvalProof::prove, verify accepts — and rejects eval + 1.  The same
    roundtrip at 2^6 elements walks three rounds of the Bulletproofs reduction (dot_product.rs:350-384)."""
    lib = ol.lib()
    Z = ol.fr_array([1, 2, 1, 4])
    r = ol.fr_array([4, 3])
    seed = ol.fr_array([12345])
    out = np.zeros(4, dtype=np.uint64)
    for tamper in (0, 1):
        rc = lib.orc_polyeval_roundtrip(ol.P(Z), ol.sz(4), ol.P(r), ol.sz(2), b"test-two", ol.P(seed), tamper, ol.P(out))
        assert rc == 0
        assert ol.fr_ints(out.reshape(1, 4)) == [28]
    rng = np.random.default_rng(4)
    Z = ol.rand_fr(rng, 64)
    r = ol.rand_fr(rng, 6)
    for tamper in (0, 1):
        assert lib.orc_polyeval_roundtrip(ol.P(Z), ol.sz(64), ol.P(r), ol.sz(6), b"test-two", ol.P(seed), tamper, ol.P(out)) == 0
