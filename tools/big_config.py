"""Run one BASELINE config end to end on the GPU(s) and (optionally) check the bytes against the CPU oracle.
usage: python tools/big_config.py KIND C LOG_M LOG_R LOG_S [check] [reps]     (torchrun + env SHARDED=1 for one sharded proof)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lasso_b200 as lb
import oracle_lib as ol

kind, C, log_m, log_r, log_s = [int(x) for x in sys.argv[1:6]]
check = len(sys.argv) > 6 and sys.argv[6] == "check"
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
sharded = os.environ.get("SHARDED") == "1" and world > 1
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << log_s
rng = np.random.default_rng(5)
col = rng.integers(0, 1 << log_m, size=(n, 1), dtype=np.uint64)
idx = np.ascontiguousarray(np.repeat(col, C, axis=1))
r = ol.rand_fr(rng, log_s); seed = ol.rand_fr(rng, 1)[0]
S = lb.Strategy(kind, C, log_m, log_r)
ctx = lb.Context(local)
if sharded:
    ctx.init_comm(rank, world)
need = lb.gens_points_needed(C, n, S.num_memories, log_m)
t = time.time(); stream = np.ascontiguousarray(ol.generators(need)); tg = time.time() - t
t = time.time(); gens = lb.SparsePolyCommitmentGens.new(ctx, b"g", C, n, S.num_memories, log_m, stream=stream); tt = time.time() - t
if rank == 0:
    print("generators: %d points, sample %.1f s, table %.2f s" % (need, tg, tt), flush=True)
for it in range(reps):
    t0 = time.time(); dense = lb.DensifiedRepresentation.from_lookup_indices(ctx, idx, log_m); t1 = time.time()
    com = dense.commit(gens); t2 = time.time()
    proof = lb.SparsePolynomialEvaluationProof.prove(ctx, S, dense, r, gens, tape_seed=seed); t3 = time.time()
    if rank == 0:
        print("kind=%d C=%d M=2^%d s=2^%d world=%d%s: densify %.1f ms, commit %.1f ms, prove %.1f ms -> %.3g lookups/s (proof %d B)" % (
            kind, C, log_m, log_s, world, " sharded" if sharded else "", (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3,
            n / (t3 - t0), len(proof.bytes)), flush=True)
    del dense
if check and rank == 0:
    t = time.time()
    ref = ol.prove(kind, C, log_m, log_r, idx, r, stream, seed, flags=1)
    print("oracle: rc=%d in %.1f s; commitment equal: %s; proof equal: %s" % (ref["rc"], time.time() - t, com == ref["commitment"],
                                                                             proof.bytes == ref["proof"]), flush=True)
    assert ref["rc"] == 0 and com == ref["commitment"] and proof.bytes == ref["proof"]
if world > 1:
    dist.barrier(); dist.destroy_process_group()
