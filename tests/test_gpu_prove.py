"""GPU parity on the whole path: Densify -> commit -> prove through the C-ABI must produce the same
commitment bytes, the same Fiat-Shamir challenges and the same proof bytes as the CPU oracle, whose
verifier must accept them (e2e_test.rs:64-99 shapes + the bench shapes)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

CASES = [  # name, kind, C, log_m, log_r, lookups, same_index
    ("prove_4d_lt", 3, 4, 4, 0, 16, True),
    ("prove_4d_lt_big_s", 3, 4, 4, 0, 128, False),
    ("prove_4d_and", 0, 4, 4, 0, 16, True),
    ("prove_3d_range", 4, 3, 8, 40, 16, False),
    ("and_c1_bench_shape", 0, 1, 16, 0, 1 << 10, True),
    ("xor_c4", 2, 4, 16, 0, 1 << 12, True),
    ("xor_c4_indep", 2, 4, 16, 0, 1 << 11, False),
    ("or_c2_ragged", 1, 2, 8, 0, 700, False),
    ("lt_c8", 3, 8, 8, 0, 1 << 9, False),
    ("range_c4", 4, 4, 16, 40, 1 << 10, False),
    # edge shapes: non-power-of-two C (zero-padded merges), the smallest sizes, a single dimension of LT
    ("xor_c3", 2, 3, 8, 0, 64, False),
    ("and_s2", 0, 2, 4, 0, 2, False),
    ("or_s4_ragged", 1, 4, 4, 0, 3, False),
    ("lt_c1", 3, 1, 4, 0, 8, False),
    ("range_c2_small_r", 4, 2, 8, 12, 16, False),
    ("xor_c16_m16", 2, 16, 4, 0, 16, False),
]


@pytest.fixture(scope="module")
def ctx():
    import lasso_b200 as lb

    c = lb.Context(0)
    yield c
    c.close()


def make_inputs(C, log_m, n, seed, same):
    rng = np.random.default_rng(seed)
    col = rng.integers(0, 1 << log_m, size=(n, 1), dtype=np.uint64)
    idx = np.repeat(col, C, axis=1) if same else rng.integers(0, 1 << log_m, size=(n, C), dtype=np.uint64)
    s = 1 << max(0, (n - 1).bit_length())
    r = ol.rand_fr(rng, max(1, s.bit_length() - 1))
    return np.ascontiguousarray(idx), r, ol.rand_fr(rng, 1)[0], s


@pytest.mark.parametrize("name,kind,C,log_m,log_r,n,same", CASES)
def test_prove_matches_oracle(ctx, name, kind, C, log_m, log_r, n, same):
    import lasso_b200 as lb

    idx, r, seed, s = make_inputs(C, log_m, n, len(name), same)
    S = lb.Strategy(kind, C, log_m, log_r)
    need = lb.gens_points_needed(C, s, S.num_memories, log_m)
    stream = np.ascontiguousarray(ol.generators(max(need, 300))[:need])
    gens = lb.SparsePolyCommitmentGens.new(ctx, b"gens_sparse_poly", C, s, S.num_memories, log_m, stream=stream)
    dense = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, log_m)
    assert dense.s == s
    commitment = dense.commit(gens)
    proof = lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r, gens, tape_seed=seed)
    ref = ol.prove(kind, C, log_m, log_r, idx, r, stream, seed, flags=1)
    assert ref["rc"] == 0
    assert commitment == ref["commitment"]
    nch = min(len(proof.challenges), len(ref["challenges"]))
    assert (proof.challenges[:nch] == ref["challenges"][:nch]).all(), "Fiat-Shamir challenges diverge"
    assert len(proof.challenges) == len(ref["challenges"])
    assert proof.bytes == ref["proof"]


@pytest.mark.parametrize("name,kind,C,log_m,log_r,n,same", [CASES[0], CASES[3]])
def test_prove_without_multiples_table(ctx, monkeypatch, name, kind, C, log_m, log_r, n, same):
    """LASSO_B200_NO_MULTIPLES=1: the openings fall back to the bucket MSM + per-step kernels (the path a sharded
    proof and memory-constrained setups use); bytes must not change."""
    import lasso_b200 as lb

    monkeypatch.setenv("LASSO_B200_NO_MULTIPLES", "1")
    idx, r, seed, s = make_inputs(C, log_m, n, len(name), same)
    S = lb.Strategy(kind, C, log_m, log_r)
    need = lb.gens_points_needed(C, s, S.num_memories, log_m)
    stream = np.ascontiguousarray(ol.generators(max(need, 300))[:need])
    gens = lb.SparsePolyCommitmentGens.new(ctx, b"gens_sparse_poly", C, s, S.num_memories, log_m, stream=stream)
    dense = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, log_m)
    commitment = dense.commit(gens)
    proof = lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r, gens, tape_seed=seed)
    ref = ol.prove(kind, C, log_m, log_r, idx, r, stream, seed, flags=1)
    assert commitment == ref["commitment"] and proof.bytes == ref["proof"]


def test_headline_config_full_size(ctx):
    """BASELINE configs[1] at its FULL size — XOR, C=4, M=2^16, 2^20 lookups: commitment and proof bytes of the GPU
    path equal the oracle's (which its own verifier accepts).  The oracle takes ~10-20 s on the box's host cores."""
    import lasso_b200 as lb

    C, log_m, n = 4, 16, 1 << 20
    idx, r, seed, s = make_inputs(C, log_m, n, 2024, True)
    S = lb.Strategy(lb.XOR, C, log_m)
    need = lb.gens_points_needed(C, s, S.num_memories, log_m)
    stream = np.ascontiguousarray(ol.generators(need))
    gens = lb.SparsePolyCommitmentGens.new(ctx, b"gens_sparse_poly", C, s, S.num_memories, log_m, stream=stream)
    dense = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, log_m)
    commitment = dense.commit(gens)
    proof = lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r, gens, tape_seed=seed)
    # size-independent sanity first: deterministic, and a different tape seed changes only the opening proofs
    proof2 = lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r, gens, tape_seed=seed)
    assert proof2.bytes == proof.bytes
    ref = ol.prove(lb.XOR, C, log_m, 0, idx, r, stream, seed, flags=1)
    assert ref["rc"] == 0
    assert commitment == ref["commitment"]
    assert proof.bytes == ref["proof"]


def test_densified_fields(ctx):
    import lasso_b200 as lb

    # memory_checking.rs:794-831 fixture through the product path: accesses [1,2,1,5], m = 8
    idx = np.array([[1], [2], [1], [5]], dtype=np.uint64)
    d = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, 3)
    assert d.dim_usize.tolist() == [[1, 2, 1, 5]]
    assert ol.fr_ints(d.read[0]) == [0, 0, 1, 0]
    assert ol.fr_ints(d.final[0]) == [0, 2, 1, 0, 0, 1, 0, 0]
    # padding with address 0 (densified.rs:37)
    d = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx[:3], 3)
    assert d.s == 4 and d.dim_usize.tolist() == [[1, 2, 1, 0]] and ol.fr_ints(d.final[0])[0] == 1
    with pytest.raises(lb.LassoError) as e:
        lb.DensifiedRepresentation.from_lookup_indices(ctx, np.array([[9]], dtype=np.uint64), 3)
    assert e.value.code == 3


def test_gpu_densify_matches_host_scan(monkeypatch):
    """densify_kernels.cu (stable LSD radix sort by address, read[k] = position - start[address]) against the host
    timestamp scan, through the public fields of DensifiedRepresentation (densified.rs:8-18); skewed addresses included."""
    import lasso_b200 as lb

    rng = np.random.default_rng(77)
    n, C, log_m = 5000, 3, 8
    idx = rng.integers(0, 1 << log_m, size=(n, C), dtype=np.uint64)
    idx[:, 1] = rng.integers(0, 3, size=n)          # heavy collisions: timestamps up to ~n/3
    idx[100:400, 2] = 17                             # a long run of one address
    fields = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("LASSO_B200_GPU_DENSIFY", mode)
        monkeypatch.setenv("LASSO_B200_HOST_DENSIFY", "0" if mode == "1" else "1")
        c = lb.Context(0)
        d = lb.DensifiedRepresentation.from_lookup_indices(c, idx, log_m)
        fields[mode] = (d.dim_usize.copy(), d.read.copy(), d.final.copy())
        del d
        c.close()
    for a, b in zip(fields["1"], fields["0"]):
        assert (a == b).all()
    # and against a plain Python restatement of densified.rs:44-51
    s = 8192
    for i in range(C):
        fin = [0] * (1 << log_m)
        rd = []
        for k in range(s):
            a = int(idx[k, i]) if k < n else 0
            rd.append(fin[a])
            fin[a] += 1
        assert ol.fr_ints(fields["1"][1][i][:64]) == rd[:64]
        assert ol.fr_ints(fields["1"][2][i][:32]) == fin[:32]
    with pytest.raises(lb.LassoError):
        monkeypatch.setenv("LASSO_B200_GPU_DENSIFY", "1")
        monkeypatch.setenv("LASSO_B200_HOST_DENSIFY", "0")
        c = lb.Context(0)
        bad = idx.copy()
        bad[7, 0] = 1 << log_m
        lb.DensifiedRepresentation.from_lookup_indices(c, bad, log_m)


def test_prove_rejects_wrong_r_length(ctx):
    import lasso_b200 as lb

    idx, r, seed, s = make_inputs(2, 4, 16, 1, True)
    S = lb.Strategy(lb.XOR, 2, 4)
    stream = np.ascontiguousarray(ol.generators(300)[: lb.gens_points_needed(2, s, 2, 4)])
    gens = lb.SparsePolyCommitmentGens.new(ctx, b"g", 2, s, 2, 4, stream=stream)
    dense = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, 4)
    with pytest.raises(lb.LassoError) as e:
        lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r[:-1], gens, tape_seed=seed)
    assert e.value.code == 1  # assert_eq!(r.len(), log2(s)) surge.rs:131
