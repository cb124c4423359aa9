"""int8 tensor-pipe diagnostics: peak (operands resident) and the k_i8syrk main loop on an L2-resident operand."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200"))
import scs_b200 as S
ctx = S.default_context()
print("peak (burst, sustained) TOP/s:", ctx.measure_i8_peak(1.0))
for mode, what in ((0, "ring, no TMA (commit per stage)"), (1, "ring, A slab via TMA (16 KB/stage)"), (2, "ring, A+B via TMA (48 KB/stage)")):
    print(f"pipe probe mode {mode} [{what}]: {ctx.i8_pipe_probe(mode):.0f} TOP/s")
