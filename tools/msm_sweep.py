"""BASELINE config 5: VariableBaseMSM-only sweep.  lasso_msm (host buffers in, one point out: includes the upload of
bases + scalars and the affine->niels conversion) vs the restated CPU msm_bigint_wnaf (single MSM = serial in the
reference, msm/mod.rs:125-147), uniform full-width scalars and "Lasso-shaped" small scalars (< 2^16).
usage: python tools/msm_sweep.py [max_log_n] [cpu_max_log_n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lasso_b200 as lb
import oracle_lib as ol
from oracle_lib import P, sz

max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 22
cpu_max = int(sys.argv[2]) if len(sys.argv) > 2 else 18
ctx = lb.Context(0)
pool = np.ascontiguousarray(ol.generators(8194))
rng = np.random.default_rng(3)
print("log_n scalars     gpu_ms   terms/s(gpu)   cpu_ms(1 thread)  same_point")
for log_n in range(16, max_log + 1, 2):
    n = 1 << log_n
    bases = np.ascontiguousarray(np.tile(pool[:8192], (n // 8192, 1)))
    for name, bits in (("full-253", 253), ("small-16", 16)):
        if bits <= 16:  # Montgomery form of small integers (F::from(u64))
            small = rng.integers(0, 1 << bits, size=n, dtype=np.uint64)
            sc = np.zeros((n, 4), dtype=np.uint64)
            ol.lib().orc_fr_from_u64_batch(P(small), sz(n), P(sc))
        else:  # limbs < 2^251 < l read as Montgomery residues: uniform full-width field elements
            raw = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64)
            raw[:, 3] &= (1 << 59) - 1
            sc = np.ascontiguousarray(raw)
        lb.msm(ctx, bases[:1024], sc[:1024])  # warm-up
        t = time.time(); got = lb.msm(ctx, bases, sc); dt = time.time() - t
        cpu_ms, same = float("nan"), "-"
        if log_n <= cpu_max:
            ref = np.zeros(16, dtype=np.uint64)
            t = time.time(); ol.lib().orc_msm(P(bases), P(sc), sz(n), 1, P(ref)); cpu_ms = (time.time() - t) * 1e3
            same = bool(ol.lib().orc_point_eq(P(got), P(ref)) == 1)
        print("%5d %-9s %9.1f %14.3g %18.1f  %s" % (log_n, name, dt * 1e3, n / dt, cpu_ms, same), flush=True)
