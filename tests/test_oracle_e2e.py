"""The reference's four end-to-end accept tests (e2e_test.rs:64-99) restated: oracle prover ->
oracle verifier accepts; a single flipped proof element rejects."""
import numpy as np
import pytest

import oracle_lib as ol

CASES = [  # (name, kind, C, log_m, log_r, sparsity)   e2e_test.rs:64-99 + the bench's AND C=1 (bench.rs:158-232)
    ("prove_4d_lt", 3, 4, 4, 0, 16),
    ("prove_4d_lt_big_s", 3, 4, 4, 0, 128),
    ("prove_4d_and", 0, 4, 4, 0, 16),
    ("prove_3d_range", 4, 3, 8, 40, 16),
    ("xor_c4", 2, 4, 4, 0, 64),
    ("or_c2_ragged", 1, 2, 4, 0, 50),  # non-power-of-two lookups: padded with address 0
    ("and_c1", 0, 1, 4, 0, 32),
]


def inputs(kind, C, log_m, n, seed=0, same_index=True):
    rng = np.random.default_rng(seed)
    col = rng.integers(0, 1 << log_m, size=(n, 1), dtype=np.uint64)
    idx = np.repeat(col, C, axis=1) if same_index else rng.integers(0, 1 << log_m, size=(n, C), dtype=np.uint64)
    s = 1 << (n - 1).bit_length()
    r = ol.rand_fr(rng, max(1, s.bit_length() - 1))
    seed_fr = ol.rand_fr(rng, 1)[0]
    return np.ascontiguousarray(idx), r, seed_fr


@pytest.mark.parametrize("name,kind,C,log_m,log_r,n", CASES)
def test_e2e_accept_and_reject(name, kind, C, log_m, log_r, n):
    gens = ol.generators(300)
    for same in (True, False):
        idx, r, seed = inputs(kind, C, log_m, n, seed=len(name), same_index=same)
        res = ol.prove(kind, C, log_m, log_r, idx, r, gens, seed, flags=1)
        assert res["rc"] == 0, "oracle verifier rejected an honest proof"
        assert len(res["proof"]) > 0 and len(res["challenges"]) > 0
        # determinism: same inputs -> same bytes
        res2 = ol.prove(kind, C, log_m, log_r, idx, r, gens, seed, flags=1)
        assert res2["proof"] == res["proof"] and res2["commitment"] == res["commitment"]
    assert ol.prove(kind, C, log_m, log_r, idx, r, gens, seed, flags=1 | 2)["rc"] == 1
    assert ol.prove(kind, C, log_m, log_r, idx, r, gens, seed, flags=1 | 4)["rc"] == 1


def test_thread_count_does_not_change_bytes():
    gens = ol.generators(300)
    idx, r, seed = inputs(2, 4, 4, 64, seed=9)
    a = ol.prove(2, 4, 4, 0, idx, r, gens, seed, flags=0, nthreads=1)
    b = ol.prove(2, 4, 4, 0, idx, r, gens, seed, flags=0, nthreads=4)
    assert a["proof"] == b["proof"]
    ol.lib().orc_set_num_threads(ol.lib().orc_num_threads())
