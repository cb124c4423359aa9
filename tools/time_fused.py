"""Times the streaming passes on a synthetic shard: python tools/time_fused.py n m [reps]
Prints per-launch ms and GB/s (8*n*m algorithmic bytes per pass) for k_forward, k_adjoint and k_fused_grad.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import scs_b200 as S
from oracle import synth

n, m = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
p = S.Problem.synthetic(n, m, S.LogisticLoss(1 / n, "consistent"), 1e-3)
ctx = p.ctx
x = synth.make_x0(m) * 0.3
by = 8.0 * n * m
res = {}
for mode in ("two_pass", "fused"):
    p.set_stream_mode(mode)
    for k in range(2):
        p.loss_eval((0.5 + 0.1 * k) * x, weights="ggn")
    ctx.set_profiling(True)
    ctx.stage_ms(reset=True)
    for k in range(reps):
        f, g, *_ = p.loss_eval((1.0 + 0.01 * k) * x, weights="ggn")
    st = ctx.stage_ms(reset=True)
    ctx.set_profiling(False)
    res[mode] = (f, g)
    for nm in ("forward", "adjoint", "fused"):
        ms, calls = st[nm]
        if calls:
            print(f"{mode:9s} {nm:8s}: {ms / calls:8.3f} ms/launch  {by / (ms / calls * 1e-3) / 1e9:8.1f} GB/s  ({calls} launches)")
f0, g0 = res["two_pass"]
f1, g1 = res["fused"]
print("rel diff f", abs(f1 - f0) / abs(f0), "g", np.linalg.norm(g1 - g0) / np.linalg.norm(g0))
p.close()
