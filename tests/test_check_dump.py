"""integration/check_dump.py (the consumer of the Rust dumper's vectors, integration/rust/README.md) — exercised here
with a directory written in the dumper's exact format from the ORACLE's outputs: this checks the file format and the
checker's plumbing (no Rust toolchain in this image); the parity statement itself needs the files from a cargo run."""
import json
import os
import subprocess
import sys

import numpy as np

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_check_dump_accepts_oracle_written_vectors(tmp_path):
    kind, C, log_m, log_r, n = 2, 2, 8, 0, 200  # XOR, ragged lookup count
    rng = np.random.default_rng(7)
    idx = np.ascontiguousarray(rng.integers(0, 1 << log_m, size=(n, C), dtype=np.uint64))
    r = ol.rand_fr(rng, 8)
    seed = ol.rand_fr(rng, 1)[0]
    need = 66
    gens = np.zeros((need, 8), dtype=np.uint64)
    ol.lib().orc_sample_generators(ol.sz(need), b"gens_sparse_poly", ol.P(gens))
    res = ol.prove(kind, C, log_m, log_r, idx, r, gens, seed, flags=1)
    assert res["rc"] == 0
    d = tmp_path / "xor_c2_s200"
    d.mkdir()
    idx.tofile(d / "indices.u64")
    np.ascontiguousarray(r).tofile(d / "r.fr")
    gens.tofile(d / "gens.aff")
    np.ascontiguousarray(seed).tofile(d / "tape_seed.fr")
    (d / "commitment.bin").write_bytes(res["commitment"])
    (d / "proof.bin").write_bytes(res["proof"])
    (d / "manifest.json").write_text(json.dumps({
        "case": "xor_c2_s200", "kind": kind, "C": C, "log_m": log_m, "log_r": log_r, "lookups": n, "num_memories": C,
        "n_generators": need, "generator_label": "gens_sparse_poly", "transcript_label": "example", "tape_label": "proof",
        "deterministic_test_rng": True}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "integration", "check_dump.py"), str(tmp_path)],
                         capture_output=True, text=True, timeout=300)
    assert "CHECK_DUMP PASS" in out.stdout, out.stdout + out.stderr
    # and a flipped proof byte is reported
    b = bytearray(res["proof"])
    b[100] ^= 1
    (d / "proof.bin").write_bytes(bytes(b))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "integration", "check_dump.py"), str(tmp_path)],
                         capture_output=True, text=True, timeout=300)
    assert "CHECK_DUMP FAIL" in out.stdout and "DIFFER@100" in out.stdout
