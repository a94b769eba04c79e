"""Generate tests/golden/*.json: seeded inputs -> commitment / proof bytes of the CPU oracle (whose verifier must accept).

The reference is Rust and cannot run in this environment (no cargo), and its own tests store no bytes, so these
vectors are produced by the ORACLE restatement (oracle/), which is itself pinned against every known-answer test the
reference holds (tests/test_oracle_kats.py) — they are regression pins that let the GPU path be checked against
committed bytes.  Regenerate with:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = [  # name, kind, C, log_m, log_r, lookups, same index in every dimension, seed
    ("and_c1_s1024", 0, 1, 16, 0, 1024, True, 11),      # BASELINE configs[0] shape
    ("xor_c4_s256", 2, 4, 16, 0, 256, True, 12),         # the headline strategy, small
    ("lt_c4_s128", 3, 4, 4, 0, 128, False, 13),          # e2e_test.rs prove_4d_lt_big_s shape
    ("range40_c3_s16", 4, 3, 8, 40, 16, False, 14),      # e2e_test.rs prove_3d_range shape
    ("or_c2_s700", 1, 2, 8, 0, 700, False, 15),          # ragged: padded with address 0
]


def inputs(C, log_m, n, seed, same):
    rng = np.random.default_rng(seed)
    col = rng.integers(0, 1 << log_m, size=(n, 1), dtype=np.uint64)
    idx = np.repeat(col, C, axis=1) if same else rng.integers(0, 1 << log_m, size=(n, C), dtype=np.uint64)
    s = 1 << max(0, (n - 1).bit_length())
    r = ol.rand_fr(rng, max(1, s.bit_length() - 1))
    return np.ascontiguousarray(idx), r, ol.rand_fr(rng, 1)[0], s


def main():
    out = {"generator_label": "gens_sparse_poly", "cases": []}
    gens = ol.generators(600)
    for name, kind, C, log_m, log_r, n, same, seed in CASES:
        idx, r, tape_seed, s = inputs(C, log_m, n, seed, same)
        res = ol.prove(kind, C, log_m, log_r, idx, r, gens, tape_seed, flags=1)
        assert res["rc"] == 0, name
        out["cases"].append({
            "name": name, "kind": kind, "C": C, "log_m": log_m, "log_r": log_r, "lookups": n, "same_index": same,
            "seed": seed, "n_challenges": int(len(res["challenges"])),
            "commitment_sha256": hashlib.sha256(res["commitment"]).hexdigest(), "commitment_len": len(res["commitment"]),
            "proof_sha256": hashlib.sha256(res["proof"]).hexdigest(), "proof_len": len(res["proof"]),
            "proof_head_hex": res["proof"][:96].hex(), "first_challenge_hex": res["challenges"][0].tobytes().hex(),
        })
    # the generator stream itself (first 8 points) so a change in the sampling is caught too
    out["generators_head_sha256"] = hashlib.sha256(np.ascontiguousarray(gens[:8]).tobytes()).hexdigest()
    with open(os.path.join(HERE, "proofs.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
