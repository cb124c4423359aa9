"""Times the on-device m x m solve (blocked Cholesky + triangular solves) in isolation: python tools/time_solve.py [m]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200"))
import numpy as np
import scs_b200 as S

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
B = rng.standard_normal((m + 64, m))
M = B.T @ B / m + np.eye(m)
b = rng.standard_normal(m)
ctx = S.default_context()
ctx.set_profiling(True)
for it in range(int(os.environ.get("SOLVE_ITERS", "4"))):
    ctx.stage_ms(reset=True)
    t0 = time.perf_counter()
    d, fb = ctx.linear_solve(M, b)
    t1 = time.perf_counter()
    st = ctx.stage_ms()
    print(f"iter {it}: wall {1e3*(t1-t0):.2f} ms (includes H2D of M), device solve stage {st['solve'][0]:.3f} ms, fallback={fb}")
print("rel err", np.linalg.norm(d - np.linalg.solve(M, b)) / np.linalg.norm(d))
