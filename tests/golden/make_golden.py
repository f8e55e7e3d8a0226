"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libgsref.so, built
from /root/reference by oracle/Makefile).  Run here (build container) only:

    python tests/golden/make_golden.py

Each fixture stores the flattened inputs (what gsB200Flatten.h extracts) and the reference's own
assembled CSC matrix / right-hand side, so that the GPU box — which has no /root/reference —
can check both the C restatement and the CUDA path against the real thing.  Large cases store
only fingerprints (nnz, sum K_ij, ||rhs||_2, probes K*x for a fixed x).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refutil as R  # noqa: E402

PI2 = "2*pi^2*sin(pi*x)*sin(pi*y)"
PI3 = "3*pi^2*sin(pi*x)*sin(pi*y)*sin(pi*z)"

FULL = {
    # name: reference configuration (see oracle/ref_driver.cpp)
    "sq_p2_m8_visitor": dict(dim=2, degree=2, nelem=8, geometry=0, rhs=[PI2], dir_values=100),
    "sq_p2_m8_expr_l2proj": dict(dim=2, degree=2, nelem=8, geometry=0, path=1, rhs=[PI2], dirichlet=["x+y"], dir_values=102),
    "cube_p3_curved_m4": dict(dim=3, degree=3, nelem=4, geometry=1, rhs=[PI3], dirichlet=["x+y*z"]),
    "cube_p2_m5": dict(dim=3, degree=2, nelem=5, geometry=0, rhs=[PI3], dir_values=100),
    "cube_p1_m5": dict(dim=3, degree=1, nelem=5, geometry=0, rhs=[PI3], dirichlet=["x+y*z"]),
    "cube_p4_curved_m2": dict(dim=3, degree=4, nelem=2, geometry=1, rhs=[PI3], dirichlet=["x*y"]),
    "annulus_nurbs_p3_m6": dict(dim=2, degree=3, nelem=6, geometry=2, rhs=["sin(x)*y"], dirichlet=["x*y"]),
    "grid2x2_p2_m4": dict(dim=2, degree=2, nelem=4, geometry=3, grid=(2, 2, 1), rhs=[PI2], dirichlet=["x*y"]),
    "grid2x2x2_p2_m3": dict(dim=3, degree=2, nelem=3, geometry=3, grid=(2, 2, 2), rhs=[PI3], dirichlet=["x"]),
    "yeti_mp2_p2_m2": dict(dim=2, degree=2, nelem=2, geometry=4, xml="domain2d/yeti_mp2.xml", rhs=["1"]),
    "elasticity_2cubes_p2": dict(dim=3, degree=2, nelem=2, geometry=3, grid=(2, 1, 1), path=1, form=1, lam=2.0, mu=1.5,
                                 rhs=["x", "y*z", "1"], dirichlet=["0.1*x", "0", "y"], dir_values=101),
    # Neumann sides: gsVisitorNeumann (scalar flux) and assembleBdr(u*g_N.tr()*nv(G)) (vector data)
    "sq_p2_neumann_visitor": dict(dim=2, degree=2, nelem=5, geometry=0, rhs=[PI2], dirichlet=["x*y"], neumann_mask=(1 << 2) | (1 << 4), neu=["1+x*y"]),
    "annulus_p3_neumann_expr": dict(dim=2, degree=3, nelem=4, geometry=2, path=1, rhs=["sin(x)*y"], dirichlet=["x+y"], dir_values=102,
                                    neumann_mask=(1 << 1) | (1 << 3), neu=["x-y", "cos(x)"]),
    "cube_p2_curved_neumann": dict(dim=3, degree=2, nelem=3, geometry=1, rhs=["x*y*z"], dirichlet=["x"], neumann_mask=(1 << 2) | (1 << 3) | (1 << 5), neu=["1+x*z"]),
    "cubes2_p2_neumann_expr": dict(dim=3, degree=2, nelem=2, geometry=3, grid=(2, 1, 1), path=1, rhs=["1"], dirichlet=["z"],
                                   neumann_mask=(1 << 6) | (1 << 1), neu=["x", "y*z", "1"]),
    # the stock input of examples/poisson2_example.cpp (BASELINE config 1): 2-patch NURBS quarter annulus,
    # mixed Dirichlet (L2-projected) / Neumann, read by the reference from its own XML file
    "poisson2d_bvp_stock_r2": dict(dim=2, degree=0, nelem=2, geometry=5, xml="pde/poisson2d_bvp.xml", path=1, dir_values=102),
    "elasticity_sq_p2": dict(dim=2, degree=2, nelem=4, geometry=1, path=1, form=1, lam=80000.0, mu=80000.0,
                             rhs=["1", "x"], dirichlet=["0", "0.01*x"], dir_values=101),
    # mass form: gsGenericAssembler::assembleMass (gsVisitorMass.h:30-157) + assembleMoments, and u*u.tr()*meas(G) with
    # interpolated Dirichlet data through the expression path (elimination term -M g on the right-hand side)
    "mass_sq_p2_visitor": dict(dim=2, degree=2, nelem=6, geometry=1, path=0, form=2, rhs=["1+x*y"]),
    "mass_cube_p3_curved_expr": dict(dim=3, degree=3, nelem=2, geometry=1, path=1, form=2, rhs=["x+z"], dirichlet=["x*y"], dir_values=101),
    "mass_grid2x2_p2_expr": dict(dim=2, degree=2, nelem=3, geometry=3, grid=(2, 2, 1), path=1, form=2, rhs=["x"], dirichlet=["1+y"], dir_values=101),
    # solved Poisson problems with the reference's error-norm integrals (gsExprEvaluator::integral, poisson2_example.cpp:174-177):
    # the fixture carries the solution coefficients and int (u_ex-u_h)^2, int |grad(u_ex-u_h)|^2, int u_h^2, int |grad u_h|^2
    "norms_sq_p2_curved": dict(dim=2, degree=2, nelem=6, geometry=1, path=1, rhs=[PI2], dirichlet=["sin(pi*x)*sin(pi*y)"], dir_values=101,
                               exact="sin(pi*x)*sin(pi*y)"),
    "norms_cube_p3_curved": dict(dim=3, degree=3, nelem=2, geometry=1, path=1, rhs=[PI3], dirichlet=["sin(pi*x)*sin(pi*y)*sin(pi*z)"], dir_values=101,
                                 exact="sin(pi*x)*sin(pi*y)*sin(pi*z)"),
    "norms_annulus_nurbs_p3": dict(dim=2, degree=3, nelem=5, geometry=2, path=1, rhs=["-2*x-2*y"], dirichlet=["x^2*y+y^2*x"], dir_values=102,
                                   exact="x^2*y+y^2*x"),
    "norms_grid2x2_p2": dict(dim=2, degree=2, nelem=4, geometry=3, grid=(2, 2, 1), path=1, rhs=[PI2], dirichlet=["sin(pi*x)*sin(pi*y)"], dir_values=101,
                             exact="sin(pi*x)*sin(pi*y)"),
    # Dirichlet values by L2-projection (dirichlet::l2Projection = 102, gsDirichletValues.h:257-435) in 3-D and across patches
    "cube_p2_curved_l2proj": dict(dim=3, degree=2, nelem=3, geometry=1, path=1, rhs=[PI3], dirichlet=["x+y*z"], dir_values=102),
    "grid2x2_p3_l2proj": dict(dim=2, degree=3, nelem=3, geometry=3, grid=(2, 2, 1), path=1, rhs=[PI2], dirichlet=["sin(x)+y"], dir_values=102),
    "elasticity_sq_p2_l2proj": dict(dim=2, degree=2, nelem=4, geometry=1, path=1, form=1, lam=80000.0, mu=80000.0,
                                    rhs=["1", "x"], dirichlet=["0.01*y", "0.01*x*y"], dir_values=102),
    "cube_p3_visitor_l2proj": dict(dim=3, degree=3, nelem=2, geometry=1, path=0, rhs=[PI3], dirichlet=["x*y+z"], dir_values=102),
    # mixed degrees per direction, several right-hand sides
    "cube_p232_curved_m3": dict(dim=3, degree=2, nelem=3, geometry=1, rhs=[PI3], dirichlet=["x+y*z"], degree_dir=[2, 3, 2]),
    "sq_p31_m5": dict(dim=2, degree=1, nelem=5, geometry=1, rhs=[PI2], dirichlet=["x*y"], degree_dir=[3, 1]),
    "sq_p3_nrhs2": dict(dim=2, degree=3, nelem=5, geometry=1, rhs=["x*y", "1+x"], nrhs=2, dir_values=100),
    "cube_p2_nrhs3_interp": dict(dim=3, degree=2, nelem=2, geometry=1, rhs=["x", "y*z", "1"], dirichlet=["x", "y", "z*x"], nrhs=3, dir_values=101),
}
FINGERPRINT = {
    "cube_p3_m16": dict(dim=3, degree=3, nelem=16, geometry=0, rhs=[PI3], dir_values=100, threads=8),
    "sq_p2_m64": dict(dim=2, degree=2, nelem=64, geometry=0, rhs=[PI2], dir_values=100, threads=8),
    "cube_p4_m8": dict(dim=3, degree=4, nelem=8, geometry=0, rhs=[PI3], dir_values=100, threads=8),
    "cube_p2_m16_expr": dict(dim=3, degree=2, nelem=16, geometry=0, path=1, rhs=[PI3], dirichlet=["x*y*z"], dir_values=101, threads=1),
    # BASELINE config 1 at its stated size: the stock input of poisson2_example with 6 uniform refinements (-r 6), degree 2
    "poisson2d_bvp_stock_r6": dict(dim=2, degree=0, nelem=6, geometry=5, xml="pde/poisson2d_bvp.xml", path=1, dir_values=102, threads=8),
    # BASELINE config 3 shape: the 21-patch yeti footprint, p=2, 16x16 elements per patch, glued interfaces
    "yeti_mp2_p2_m8": dict(dim=2, degree=2, nelem=8, geometry=4, xml="domain2d/yeti_mp2.xml", rhs=["1+x"], dirichlet=["0.1*y"], threads=8),
    # BASELINE config 4 shape: 2x2x2 patches, p=2, vector-valued linear elasticity through the expression path
    "elasticity_8cubes_p2_m5": dict(dim=3, degree=2, nelem=5, geometry=3, grid=(2, 2, 2), path=1, form=1, lam=80000.0, mu=80000.0,
                                    rhs=["0", "0", "-1000"], dirichlet=["0", "0", "0.001*x"], dir_values=101, threads=1),
}
# above a million DOFs only samples and norms of the fingerprint vectors are kept (SAMPLE_STRIDE), so the fixture stays small
SAMPLED = {
    # curved geometry, non-uniform Jacobian, interpolated Dirichlet data at 1.09 M DOFs (3-D p=3, 100^3 elements)
    "cube_p3_curved_m50": dict(dim=3, degree=3, nelem=50, geometry=1, rhs=[PI3], dirichlet=["x+y*z"], threads=8),
}
SAMPLE_STRIDE = 97
KEEP_DOFMAP = {"yeti_mp2_p2_m8", "elasticity_8cubes_p2_m5", "poisson2d_bvp_stock_r6"}


def pack_inputs(ref):
    d = {"nfree": ref.nfree, "nfixed": ref.nfixed, "ncomp": ref.ncomp, "dim": ref.dim, "form": ref.form, "nrhs": getattr(ref, "nrhs", 1),
         "coef": np.asarray(ref.coef), "quA": ref.quA, "quB": ref.quB, "npatches": len(ref.patches),
         "rhs_text": np.asarray(ref.rhs_text), "fixed": ref.fixed,
         "neumann_sides": np.asarray(ref.neumann_sides, dtype=np.int32).reshape(-1, 2), "neu_text": np.asarray(ref.neu_text)}
    for k, p in enumerate(ref.patches):
        d[f"p{k}_sdeg"] = np.asarray(p.space_degree)
        d[f"p{k}_gdeg"] = np.asarray(p.geo_degree)
        for i in range(p.dim):
            d[f"p{k}_sk{i}"] = p.space_knots[i]
            d[f"p{k}_gk{i}"] = p.geo_knots[i]
        d[f"p{k}_coefs"] = p.geo_coefs
        d[f"p{k}_dofmap"] = p.dofmap
        if p.geo_weights is not None:
            d[f"p{k}_weights"] = p.geo_weights
    return d


def probe_vector(n):
    return np.cos(0.37 * np.arange(n) + 0.11)


def main():
    only = set(sys.argv[1:])
    import scipy.sparse as sp
    for name, cfg in SAMPLED.items():
        if only and name not in only:
            continue
        ref = R.ref_run(**cfg)
        d = pack_inputs(ref)
        K = sp.csc_matrix((ref.values, ref.inner, ref.outer), shape=(ref.nfree, ref.nfree))
        Kx = K @ probe_vector(ref.nfree)
        diag = K.diagonal()
        for k in list(d):
            if k.endswith("_dofmap"):
                del d[k]
        d.update(kind="sampled", config=repr(cfg), stride=SAMPLE_STRIDE, nnz=len(ref.values), sumK=ref.values.sum(), maxK=np.abs(ref.values).max(),
                 inner_checksum=np.int64(ref.inner.astype(np.int64).sum()), outer_checksum=np.int64(ref.outer.astype(np.int64).sum()),
                 Kx_s=Kx[::SAMPLE_STRIDE], diag_s=diag[::SAMPLE_STRIDE], rhs_s=ref.rhs[::SAMPLE_STRIDE, 0],
                 Kx_norm=np.linalg.norm(Kx), diag_norm=np.linalg.norm(diag), rhs_norm=np.linalg.norm(ref.rhs), rhs_max=np.abs(ref.rhs).max(),
                 seconds=ref.seconds)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "N", ref.nfree, "nnz", len(ref.values), "reference seconds", ref.seconds)
    for name, cfg in FULL.items():
        if only and name not in only:
            continue
        ref = R.ref_run(**cfg)
        d = pack_inputs(ref)
        d.update(outer=ref.outer, inner=ref.inner, values=ref.values, rhs=ref.rhs, kind="full", config=repr(cfg))
        if cfg.get("exact"):
            d.update(solution=ref.solution, norms=ref.norms, exact_text=cfg["exact"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "N", ref.nfree, "nnz", len(ref.values))
    for name, cfg in FINGERPRINT.items():
        if only and name not in only:
            continue
        ref = R.ref_run(**cfg)
        d = pack_inputs(ref)
        import scipy.sparse as sp
        K = sp.csc_matrix((ref.values, ref.inner, ref.outer), shape=(ref.nfree, ref.nfree))
        x = probe_vector(ref.nfree)
        d.update(kind="fingerprint", config=repr(cfg), nnz=len(ref.values), sumK=ref.values.sum(),
                 rhs_norm=np.linalg.norm(ref.rhs), Kx=K @ x, maxK=np.abs(ref.values).max(), rhs=ref.rhs,
                 outer=ref.outer, inner_checksum=np.int64(ref.inner.astype(np.int64).sum()),
                 diag=K.diagonal())
        # dof maps of fingerprint cases are large; they are rebuilt by gismo_b200.host in the tests
        for k in list(d):
            if k.endswith("_dofmap") and name not in KEEP_DOFMAP:
                del d[k]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "N", ref.nfree, "nnz", len(ref.values), repr(ref.values.sum()), repr(np.linalg.norm(ref.rhs)))


if __name__ == "__main__":
    main()
