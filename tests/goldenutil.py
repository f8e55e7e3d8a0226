"""Load tests/golden/*.npz (made by make_golden.py from the reference) into C-ABI problems."""
import glob
import os

import numpy as np

from gismo_b200.capi import PatchData, Problem
from gismo_b200 import host

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def names(kind):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        with np.load(f) as z:
            if str(z["kind"]) == kind:
                out.append(os.path.splitext(os.path.basename(f))[0])
    return out


def load(name, compile_fn, with_rhs=True):
    """-> (Problem, dict of expected arrays)."""
    z = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    dim, ncomp, npatches = int(z["dim"]), int(z["ncomp"]), int(z["npatches"])
    patches = []
    for k in range(npatches):
        sdeg = [int(v) for v in z[f"p{k}_sdeg"]]
        sk = [z[f"p{k}_sk{i}"] for i in range(dim)]
        if f"p{k}_dofmap" in z:
            dofmap = z[f"p{k}_dofmap"]
        else:  # fingerprint cases: single patch, all sides eliminated -> rebuilt by the host builder
            nfun = [len(sk[i]) - sdeg[i] - 1 for i in range(dim)]
            m = host.DofMapper([int(np.prod(nfun))])
            m.eliminate(0, host.boundary_indices(nfun))
            m.finalize()
            dofmap = m.patch_map(0)
        patches.append(PatchData(sdeg, sk, [int(v) for v in z[f"p{k}_gdeg"]], [z[f"p{k}_gk{i}"] for i in range(dim)],
                                 z[f"p{k}_coefs"], dofmap, z.get(f"p{k}_weights")))
    progs = [compile_fn(str(t)) for t in z["rhs_text"]] if with_rhs else None
    nfixed = int(z["nfixed"])
    neumann = []
    if "neumann_sides" in z and len(z["neumann_sides"]):
        neu = [compile_fn(str(t)) for t in z["neu_text"]]
        neumann = [(int(p), int(s), neu) for p, s in z["neumann_sides"]]
    nrhs = int(z["nrhs"]) if "nrhs" in z else 1
    pb = Problem(patches, int(z["nfree"]), nfixed, form=int(z["form"]), ncomp=ncomp,
                 fixed=z["fixed"] if nfixed and z["fixed"].size else None, nrhs=nrhs, coef=tuple(z["coef"]), quA=float(z["quA"]),
                 quB=int(z["quB"]), rhs_programs=progs, neumann=neumann)
    return pb, z


def probe_vector(n):
    return np.cos(0.37 * np.arange(n) + 0.11)


def check_against(result, z, tol):
    """result = (outer, inner, values, rhs) from any implementation; z = golden dict."""
    import scipy.sparse as sp
    outer, inner, values, rhs = result[:4]
    if str(z["kind"]) == "full":
        assert np.array_equal(outer, z["outer"]), "outer index array differs from the reference"
        assert np.array_equal(inner, z["inner"]), "inner index array differs from the reference"
        scale = np.abs(z["values"]).max()
        ev = np.abs(values - z["values"]).max() / scale
        rs = max(np.abs(z["rhs"]).max(), 1e-300)
        er = np.abs(rhs - z["rhs"]).max() / rs
        assert ev <= tol, f"matrix values differ: {ev:.3e}"
        assert er <= tol, f"rhs differs: {er:.3e}"
        return ev, er
    n = int(z["nfree"])
    assert len(values) == int(z["nnz"])
    if str(z["kind"]) == "sampled":          # million-DOF cases: samples (every stride-th entry) and norms of the fingerprint vectors
        st = int(z["stride"])
        assert int(outer.astype(np.int64).sum()) == int(z["outer_checksum"]) and int(inner.astype(np.int64).sum()) == int(z["inner_checksum"])
        K = sp.csc_matrix((values, inner, outer), shape=(n, n))
        scale = float(z["maxK"])
        Kx, diag = K @ probe_vector(n), K.diagonal()
        e1 = max(np.abs(Kx[::st] - z["Kx_s"]).max() / (scale * 50), abs(np.linalg.norm(Kx) - float(z["Kx_norm"])) / float(z["Kx_norm"]))
        e2 = max(np.abs(diag[::st] - z["diag_s"]).max() / scale, abs(np.linalg.norm(diag) - float(z["diag_norm"])) / float(z["diag_norm"]))
        e3 = abs(values.sum() - float(z["sumK"])) / (scale * np.sqrt(len(values)))
        er = max(np.abs(rhs[::st, 0] - z["rhs_s"]).max() / float(z["rhs_max"]), abs(np.linalg.norm(rhs) - float(z["rhs_norm"])) / float(z["rhs_norm"]))
        assert max(e1, e2, e3) <= tol, f"fingerprints differ: Kx {e1:.2e} diag {e2:.2e} sum {e3:.2e}"
        assert er <= tol, f"rhs differs: {er:.3e}"
        return max(e1, e2, e3), er
    assert np.array_equal(outer, z["outer"])
    assert int(inner.astype(np.int64).sum()) == int(z["inner_checksum"])
    K = sp.csc_matrix((values, inner, outer), shape=(n, n))
    scale = float(z["maxK"])
    e1 = np.abs(K @ probe_vector(n) - z["Kx"]).max() / (scale * 50)
    e2 = np.abs(K.diagonal() - z["diag"]).max() / scale
    e3 = abs(values.sum() - float(z["sumK"])) / (scale * np.sqrt(len(values)))
    er = np.abs(rhs - z["rhs"]).max() / np.abs(z["rhs"]).max()
    assert max(e1, e2, e3) <= tol, f"fingerprints differ: Kx {e1:.2e} diag {e2:.2e} sum {e3:.2e}"
    assert er <= tol, f"rhs differs: {er:.3e}"
    return max(e1, e2, e3), er


# Error-norm fixtures (make_golden.py "norms_*"): exact solution and its analytic gradient.  The reference differentiates the exact
# solution numerically (gsFunctionExpr::deriv_into: exprtk::derivative with h = 1e-5, gsFunctionExpr.hpp:582), so its H1 error integral
# carries ~1e-10 of finite-difference error: that one entry is compared at 1e-8, the other three at 1e-12.
NORM_CASES = {
    "norms_sq_p2_curved": ("sin(pi*x)*sin(pi*y)", ["pi*cos(pi*x)*sin(pi*y)", "pi*sin(pi*x)*cos(pi*y)"]),
    "norms_cube_p3_curved": ("sin(pi*x)*sin(pi*y)*sin(pi*z)", ["pi*cos(pi*x)*sin(pi*y)*sin(pi*z)", "pi*sin(pi*x)*cos(pi*y)*sin(pi*z)",
                                                               "pi*sin(pi*x)*sin(pi*y)*cos(pi*z)"]),
    "norms_annulus_nurbs_p3": ("x^2*y+y^2*x", ["2*x*y+y^2", "x^2+2*y*x"]),
    "norms_grid2x2_p2": ("sin(pi*x)*sin(pi*y)", ["pi*cos(pi*x)*sin(pi*y)", "pi*sin(pi*x)*cos(pi*y)"]),
}


def check_norms(got, z, tol=1e-12):
    """got = [int (u_h-u_ex)^2, int |grad(u_h-u_ex)|^2, int u_h^2, int |grad u_h|^2] against the reference's gsExprEvaluator integrals."""
    ref = z["norms"]
    assert abs(got[2] - ref[2]) <= tol * abs(ref[2]), f"int u_h^2 differs: {got[2]} vs {ref[2]}"
    assert abs(got[3] - ref[3]) <= tol * abs(ref[3]), f"int |grad u_h|^2 differs: {got[3]} vs {ref[3]}"
    assert abs(got[0] - ref[0]) <= tol * abs(ref[2]) and abs(got[0] - ref[0]) <= 1e-9 * abs(ref[0]), f"L2 error differs: {got[0]} vs {ref[0]}"
    assert abs(got[1] - ref[1]) <= 1e-8 * abs(ref[1]), f"H1 error differs: {got[1]} vs {ref[1]}"


# Fixtures whose eliminated values the reference computed by L2-projection (DirichletValues = 102): Dirichlet data and the sides that
# carry it (single patches: sides 1..2d minus the Neumann ones; the 2-D grid: read off the DOF map, host.multipatch_topology_2d)
L2PROJ_CASES = {
    "sq_p2_m8_expr_l2proj": ("x+y", [(0, s) for s in (1, 2, 3, 4)]),
    "norms_annulus_nurbs_p3": ("x^2*y+y^2*x", [(0, s) for s in (1, 2, 3, 4)]),
    "annulus_p3_neumann_expr": ("x+y", [(0, 2), (0, 4)]),
    "cube_p2_curved_l2proj": ("x+y*z", [(0, s) for s in range(1, 7)]),
    "cube_p3_visitor_l2proj": ("x*y+z", [(0, s) for s in range(1, 7)]),
    "grid2x2_p3_l2proj": ("sin(x)+y", None),
    "elasticity_sq_p2_l2proj": (["0.01*y", "0.01*x*y"], [(0, s) for s in (1, 2, 3, 4)]),      # vector-valued: one datum per component
}


def l2proj_sides(name, pb):
    text, sides = L2PROJ_CASES[name]
    if sides is None:
        _, sides = host.multipatch_topology_2d(pb)
    return text, list(sides)
