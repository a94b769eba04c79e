"""Synthetic workloads shared by bench.py, the golden-vector scripts and the GPU tests: the BASELINE.json
configurations with explicit, recorded seeds (the reference draws its inputs from ark_std::test_rng,
src/benches/bench.rs:13-34; here the seed is part of the contract so that a hash of the proof can be pinned).

Nothing in here touches the oracle's code: only its Python big-int -> Montgomery helpers."""
import numpy as np

BENCH_SEED = 0x4C4153534F  # "LASSO"

# name -> (strategy kind, C, log_m, log_r, log_s, seed)        kinds: 0 AND, 1 OR, 2 XOR, 3 LT, 4 RANGE_CHECK
CONFIGS = {
    "and_c1_s10": (0, 1, 16, 0, 10, BENCH_SEED + 1),    # BASELINE configs[0]
    "xor_c4_s20": (2, 4, 16, 0, 20, BENCH_SEED),        # BASELINE configs[1] — the headline (bench.py, rank 0)
    "lt_c8_s22": (3, 8, 16, 0, 22, BENCH_SEED + 3),     # BASELINE configs[2]
    "rc40_c4_s24": (4, 4, 16, 40, 24, BENCH_SEED + 4),  # BASELINE configs[3] (LOG_R = 40: range_check.rs:103-136)
    # reduced sizes of the same shapes (oracle finishes in seconds; used by the CPU-side golden checks)
    "xor_c4_s14": (2, 4, 16, 0, 14, BENCH_SEED + 10),
    "lt_c8_s14": (3, 8, 16, 0, 14, BENCH_SEED + 11),
    "rc40_c4_s14": (4, 4, 16, 40, 14, BENCH_SEED + 12),
}


def num_memories(kind, C):
    return 2 * C if kind == 3 else C


def make_inputs(log_s, C, log_m, seed):
    """Synthetic lookups mirroring src/benches/bench.rs:13-34: one uniform index per lookup, repeated in all
    C dimensions ([x; C]); r = log2(s) uniform field elements; tape seed = one more."""
    import oracle_lib as ol  # big-int -> Montgomery helpers only

    rng = np.random.default_rng(seed)
    n = 1 << log_s
    col = rng.integers(0, 1 << log_m, size=(n, 1), dtype=np.uint64)
    idx = np.ascontiguousarray(np.repeat(col, C, axis=1))
    r = ol.rand_fr(rng, log_s)
    tape_seed = ol.rand_fr(rng, 1)[0]
    return idx, r, tape_seed


def config_inputs(name):
    kind, C, log_m, log_r, log_s, seed = CONFIGS[name]
    idx, r, tape_seed = make_inputs(log_s, C, log_m, seed)
    return kind, C, log_m, log_r, log_s, idx, r, tape_seed


def gens_needed(C, log_s, alpha, log_m):
    """lasso_gens_points_needed (surge.rs:32-58) in Python: the widest of the three PolyCommitmentGens + 2."""
    def nv(x):
        p = 1
        while p < x:
            p <<= 1
        return p.bit_length() - 1
    s = 1 << log_s
    mx = max(nv(2 * C * s), nv(C) + log_m, nv(alpha * s))
    return (1 << (mx - mx // 2)) + 2
