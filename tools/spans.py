import os, sys, time, json
os.environ["LASSO_B200_SPANS"]="1"
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, lasso_b200 as lb, bench
C,log_m,log_s=4,16,int(sys.argv[1]) if len(sys.argv)>1 else 20
s=1<<log_s
idx,r,seed=bench.make_inputs(log_s,C,log_m,1)
ctx=lb.Context(0); S=lb.Strategy(lb.XOR,C,log_m)
need=lb.gens_points_needed(C,s,4,log_m)
t=time.time(); stream=lb.sample_generators(b"gens_sparse_poly",need); print("sample gens",time.time()-t)
t=time.time(); gens=lb.SparsePolyCommitmentGens.new(ctx,b"g",C,s,4,log_m,stream=stream); print("gens create",time.time()-t)
for it in range(3):
    t=time.time(); dense=lb.DensifiedRepresentation.from_lookup_indices(ctx,idx,log_m); t1=time.time()
    com=dense.commit(gens); t2=time.time()
    p=lb.SparsePolynomialEvaluationProof.prove(ctx,S,dense,r,gens,tape_seed=seed); t3=time.time()
    print("densify %.1f ms commit %.1f ms prove %.1f ms"%((t1-t)*1e3,(t2-t1)*1e3,(t3-t2)*1e3), len(p.bytes), ctx.launches)
    print({k:round(v,2) for k,v in ctx.spans().items()})
