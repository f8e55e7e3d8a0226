import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, goldenutil as G, refutil as R, gismo_b200 as g
from gismo_b200 import capi
lib = capi.load_library()
for name in G.names("full") + G.names("fingerprint"):
    pb, z = G.load(name, g.expr_compile)
    os.environ.pop("GSB200_NO_TMA", None)
    a = R.lib_assemble(lib, pb)
    os.environ["GSB200_NO_TMA"] = "1"
    b = R.lib_assemble(lib, pb)
    dv = np.abs(a[2] - b[2]); dr = np.abs(a[3] - b[3]).max()
    bad = np.nonzero(dv > 1e-12 * np.abs(b[2]).max())[0]
    print(f"{name:28s} nnz {len(a[2]):8d} bad {len(bad):8d} max {dv.max():.3e} rhs {dr:.2e}", end="")
    if len(bad):
        cols = np.searchsorted(a[0], bad, side="right") - 1
        print("  cols", np.unique(cols)[:12], "n", len(np.unique(cols)), " a", a[2][bad[:3]], " b", b[2][bad[:3]], end="")
    print()
