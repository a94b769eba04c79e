"""Runs only the bind kernel (K1) on 5 x 2^22 elements — the ncu --set full target for the roofline kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lasso_b200 as lb
ctx = lb.Context(0)
print("bind ms", ctx.bench_bind(1 << 22, 5, 5))
