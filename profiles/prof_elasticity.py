"""ncu driver: BASELINE config 4 (3-D elasticity p=2, 2x2x2 patches of 75^3 elements), pattern + 2 assemblies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gismo_b200 as g
progs = [g.expr_compile(t) for t in ("x", "y*z", "1")]
pb = g.host.multipatch_grid_problem(3, 2, [2, 2, 2], int(os.environ.get("NELEM", "75")), rhs_programs=progs, form=g.capi.FORM_ELASTICITY, coef=(2.0, 1.5))
A = g.DeviceAssembler(pb); A.buildPattern()
for _ in range(2):
    A.assemble()
t = A.timings(); print("ms:", list(t.sweep_ms), t.total_ms)
