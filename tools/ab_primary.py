"""A/B of the primary sumcheck inside ONE process: the bind fused into the next round's evaluation launch against
LASSO_B200_UNFUSED_PRIMARY=1 (two launches per round), interleaved proof by proof on the same context and inputs.
Prints the median commit+prove time and the library's Sumcheck.prove span of both arms, and checks the bytes agree."""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import lasso_b200 as lb

log_s = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 30
C, log_m = 4, 16
idx, r, seed = bench.make_inputs(log_s, C, log_m, bench.wl.BENCH_SEED)  # s = 2^20: the bench workload
spans_on = os.environ.get("LASSO_B200_SPANS") == "1"  # the library then synchronises around every span
ctx = lb.Context(0)
S = lb.Strategy(lb.XOR, C, log_m)
need = lb.gens_points_needed(C, 1 << log_s, 4, log_m)
gens = lb.SparsePolyCommitmentGens.new(ctx, b"gens_sparse_poly", C, 1 << log_s, 4, log_m, stream=bench.generator_stream(lb, need))
res = {"fused": [], "unfused": []}
span = {"fused": [], "unfused": []}
digest = {}
for it in range(3 + pairs):
    for arm in (("fused", "unfused") if it % 2 == 0 else ("unfused", "fused")):
        if arm == "unfused":
            os.environ["LASSO_B200_UNFUSED_PRIMARY"] = "1"
        else:
            os.environ.pop("LASSO_B200_UNFUSED_PRIMARY", None)
        dense = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, log_m)
        dense.commit(gens)  # also drains the asynchronous densify
        ctx.spans()
        t0 = time.perf_counter()
        dense.commit(gens)
        p = lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r, gens, tape_seed=seed)
        dt = 1e3 * (time.perf_counter() - t0)
        sp = ctx.spans()
        del dense
        if it < 3:
            continue
        res[arm].append(dt)
        span[arm].append(sp.get("Sumcheck.prove", float("nan")))
        digest.setdefault(arm, hashlib.sha256(p.bytes).hexdigest())
for arm in ("fused", "unfused"):
    a = np.array(res[arm]); b = np.array(span[arm])
    print("%-8s commit+prove median %.3f ms  min %.3f  p90 %.3f | Sumcheck.prove median %.3f ms  (%d proofs)"
          % (arm, np.median(a), a.min(), np.percentile(a, 90), np.median(b), len(a)))
print("same bytes:", digest["fused"] == digest["unfused"])
if log_s == 20:  # the bench workload: its proof hash is pinned by the CPU oracle
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "big_proofs.json")))
    want = gold["cases"]["xor_c4_s20"]["proof_sha256"]
    print("golden match:", digest["fused"] == want)
