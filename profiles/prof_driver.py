"""Small driver for ncu captures: config-2 problem (or --nelem), pattern + N assemblies."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gismo_b200 as g
ap = argparse.ArgumentParser(); ap.add_argument("--nelem", type=int, default=125); ap.add_argument("--degree", type=int, default=3)
ap.add_argument("--reps", type=int, default=2); ap.add_argument("--dim", type=int, default=3)
a = ap.parse_args()
prog = g.expr_compile("3*pi^2*sin(pi*x)*sin(pi*y)*sin(pi*z)" if a.dim == 3 else "2*pi^2*sin(pi*x)*sin(pi*y)")
pb = g.host.poisson_box_problem(a.dim, a.degree, a.nelem, prog)
A = g.DeviceAssembler(pb); A.buildPattern()
for _ in range(a.reps):
    A.assemble()
t = A.timings(); print("ms:", t.geometry_ms, list(t.sweep_ms), t.rhs_ms, t.total_ms)
if os.environ.get("PROF_CG"):
    import time, numpy as np
    b = A.rhs()[:, 0]
    n = int(os.environ["PROF_CG"])
    t0 = time.time(); u, it, res = A.cg(b, max_iter=n, tol=1e-30); dt = time.time() - t0
    print(f"cg: {it} iterations in {dt*1e3:.1f} ms = {dt/it*1e3:.3f} ms/iter, rel.res {res:.3e}")
