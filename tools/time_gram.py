"""Times the emulated-fp64 Gram stages on a synthetic shard: python tools/time_gram.py n m [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import scs_b200 as S
from oracle import synth

n, m = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
p = S.Problem.synthetic(n, m, S.LogisticLoss(1 / n, "consistent"), 1e-3)
p.set_gram_mode("i8")
ctx = p.ctx
x = synth.make_x0(m) * 0.3
G0 = p.gram(x, weights="ggn")
ctx.set_profiling(True)
ctx.stage_ms(reset=True)
for k in range(reps):
    G = p.gram((1.0 + 0.01 * k) * x, weights="ggn")
st = ctx.stage_ms(reset=True)
nmod, bits = p.gram_info()
for nm in ("residues", "gram", "gram_finalize"):
    ms, calls = st[nm]
    if calls:
        extra = f"  {nmod * n * m * (m + 1) / (ms / calls * 1e-3) / 1e12:8.1f} TOP/s" if nm == "gram" else ""
        print(f"{nm:14s}: {ms / calls:8.3f} ms/launch{extra}")
print("moduli", nmod, "bits", bits, "2cta", os.environ.get("SCS_I8_2CTA", "1"))
p.close()
