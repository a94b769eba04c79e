"""tests/golden/big_proofs.json: the reduced-size entries are re-derived here with the oracle (seconds); the at-size
entries (2^20 / 2^22 / 2^24 lookups: minutes to tens of minutes of CPU, tens of GB) are produced offline by
tests/golden/make_golden_big.py and compared with the GPU bytes by tests/test_gpu_big_configs.py and bench.py."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
import workloads as wl

HERE = os.path.dirname(os.path.abspath(__file__))
DOC = json.load(open(os.path.join(HERE, "golden", "big_proofs.json")))


@pytest.mark.parametrize("name", ["and_c1_s10", "xor_c4_s14", "lt_c8_s14", "rc40_c4_s14"])
def test_oracle_reproduces_reduced_golden(name):
    g = DOC["cases"][name]
    kind, C, log_m, log_r, log_s, idx, r, tape_seed = wl.config_inputs(name)
    assert hashlib.sha256(idx.tobytes()).hexdigest() == g["indices_sha256"]
    need = wl.gens_needed(C, log_s, wl.num_memories(kind, C), log_m)
    assert need == g["n_generators"]
    gens = np.ascontiguousarray(ol.generators(need))
    assert hashlib.sha256(gens.tobytes()).hexdigest() == g["generators_sha256"]
    res = ol.prove(kind, C, log_m, log_r, idx, r, gens, tape_seed, flags=1)
    assert res["rc"] == 0
    assert hashlib.sha256(res["commitment"]).hexdigest() == g["commitment_sha256"]
    assert hashlib.sha256(res["proof"]).hexdigest() == g["proof_sha256"]


def test_at_size_entries_present():
    for name in ("xor_c4_s20", "lt_c8_s22", "rc40_c4_s24"):
        g = DOC["cases"][name]
        kind, C, log_m, log_r, log_s, seed = wl.CONFIGS[name]
        assert (g["kind"], g["C"], g["log_m"], g["log_r"], g["log_s"], g["seed"]) == (kind, C, log_m, log_r, log_s, seed)
        assert g["oracle_verifier"] == "accepted" and len(g["proof_sha256"]) == 64


def test_gens_needed_matches_the_library_formula():
    # surge.rs:32-58: the widest of the three PolyCommitmentGens (+ Q, h)
    assert wl.gens_needed(4, 20, 4, 16) == 4098
    assert wl.gens_needed(8, 22, 16, 16) == 8194
    assert wl.gens_needed(4, 24, 4, 16) == 16386
    assert wl.gens_needed(1, 10, 1, 16) == 258
