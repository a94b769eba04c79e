"""Committed golden vectors (tests/golden/proofs.json, made by tests/golden/make_golden.py): the oracle must keep
reproducing them on the CPU, and the CUDA path must reproduce them on the GPU."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "proofs.json")))
NGENS = 600


def _inputs(case):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden import inputs

    return inputs(case["C"], case["log_m"], case["lookups"], case["seed"], case["same_index"])


def _check(case, commitment, proof, challenges):
    assert len(commitment) == case["commitment_len"] and len(proof) == case["proof_len"]
    assert hashlib.sha256(commitment).hexdigest() == case["commitment_sha256"]
    assert hashlib.sha256(proof).hexdigest() == case["proof_sha256"]
    assert proof[:96].hex() == case["proof_head_hex"]
    assert len(challenges) == case["n_challenges"]
    assert challenges[0].tobytes().hex() == case["first_challenge_hex"]


def test_generator_stream_is_pinned():
    g = ol.generators(NGENS)
    assert hashlib.sha256(np.ascontiguousarray(g[:8]).tobytes()).hexdigest() == GOLD["generators_head_sha256"]


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_reproduces_golden(case):
    idx, r, seed, s = _inputs(case)
    res = ol.prove(case["kind"], case["C"], case["log_m"], case["log_r"], idx, r, ol.generators(NGENS), seed, flags=1)
    assert res["rc"] == 0
    _check(case, res["commitment"], res["proof"], res["challenges"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_gpu_reproduces_golden(case):
    import lasso_b200 as lb

    idx, r, seed, s = _inputs(case)
    S = lb.Strategy(case["kind"], case["C"], case["log_m"], case["log_r"])
    ctx = lb.Context(0)
    need = lb.gens_points_needed(case["C"], s, S.num_memories, case["log_m"])
    stream = np.ascontiguousarray(ol.generators(NGENS)[:need])
    gens = lb.SparsePolyCommitmentGens.new(ctx, b"gens_sparse_poly", case["C"], s, S.num_memories, case["log_m"], stream=stream)
    dense = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, case["log_m"])
    com = dense.commit(gens)
    proof = lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r, gens, tape_seed=seed)
    _check(case, com, proof.bytes, proof.challenges)
    ctx.close()
