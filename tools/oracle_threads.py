"""How the CPU oracle port scales with host threads / malloc arenas on this box (picks the fair CPU baseline setting)."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench, oracle_lib as ol
    t = int(sys.argv[2]); log_s = int(sys.argv[3])
    olib, idx, r, seed, gens, cores = bench.cpu_sample(log_s, 4, 16, threads=t)
    ol.prove(2, 4, 16, 0, idx, r, gens, seed, flags=0)
    t0 = time.time(); ol.prove(2, 4, 16, 0, idx, r, gens, seed, flags=0); dt = time.time() - t0
    print("threads=%d arena=%s log_s=%d: %.2f s" % (t, os.environ.get("ORACLE_ARENA_MAX", "1"), log_s, dt), flush=True)
else:
    n = os.cpu_count()
    for arena in ("1", "0"):
        for t in sorted({n, n // 2, n // 4, 32, 16}):
            if t < 1: continue
            env = dict(os.environ, ORACLE_ARENA_MAX=arena)
            subprocess.run([sys.executable, __file__, "child", str(t), "18"], env=env)
