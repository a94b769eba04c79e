"""CUDA-synchronised spans of one proof of a BASELINE configuration (LASSO_B200_SPANS=1: the analogue of the
reference's tracing spans).  usage: python tools/spans_config.py xor_c4_s20|lt_c8_s22|rc40_c4_s24 [reps]"""
import os, sys, time
os.environ["LASSO_B200_SPANS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lasso_b200 as lb
import workloads as wl
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "xor_c4_s20"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind, C, log_m, log_r, log_s, idx, r, seed = wl.config_inputs(name)
ctx = lb.Context(0)
S = lb.Strategy(kind, C, log_m, log_r)
s = 1 << log_s
need = lb.gens_points_needed(C, s, S.num_memories, log_m)
gens = lb.SparsePolyCommitmentGens.new(ctx, b"gens_sparse_poly", C, s, S.num_memories, log_m, stream=bench.generator_stream(lb, need))
for it in range(reps):
    t = time.time(); dense = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, log_m); t1 = time.time()
    com = dense.commit(gens); t2 = time.time()
    p = lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r, gens, tape_seed=seed); t3 = time.time()
    print("%s: densify %.1f ms commit %.1f ms prove %.1f ms, %d launches" % (name, (t1 - t) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, ctx.launches), flush=True)
    sp = ctx.spans()
    print("  " + " · ".join("%s %.2f" % (k, v) for k, v in sorted(sp.items(), key=lambda kv: -kv[1])), flush=True)
    del dense
