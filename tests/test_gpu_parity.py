"""Parity tests proper: the CUDA path, called through the C ABI (libgsb200.so), against
 (1) the reference's own results (tests/golden, made by the unmodified reference),
 (2) the C oracle on seeded random inputs,
 (3) size-independent properties at BASELINE config-2 size (3D p=3, 125^3 elements, 2.0 M DOFs):
     Kronecker structure of the stiffness matrix on an affine patch, analytic nnz, symmetry,
     separable load vector, determinism.
Bars: sparsity pattern and DOF mapping bit-exact; values/rhs within 1e-12 of max|K| / max|rhs|
(BASELINE.json north_star).  All tests need a CUDA device.
"""
import ctypes as C

import os

import numpy as np
import pytest

import goldenutil as G
import refutil as R
import gismo_b200 as g
from gismo_b200 import capi, host

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def lib():
    lib = capi.load_library()
    n = C.c_int(0)
    lib.gsb200_device_count(C.byref(n))
    assert n.value > 0, "no CUDA device: the gpu-marked tests must run on the B200 box"
    return lib


@pytest.mark.parametrize("name", G.names("full") + G.names("fingerprint"))
def test_cuda_matches_reference_fixture(lib, name):
    pb, z = G.load(name, g.expr_compile)
    G.check_against(R.lib_assemble(lib, pb), z, TOL)


@pytest.mark.parametrize("name", ["cube_p3_curved_m4", "grid2x2_p2_m4", "elasticity_2cubes_p2"])
def test_one_shot_host_entry_point(lib, name):
    pb, z = G.load(name, g.expr_compile)
    G.check_against(g.assemble_host(pb), z, TOL)


@pytest.mark.parametrize("dim,p,m,seed", [(3, 3, 6, 12345), (3, 2, 7, 7), (2, 4, 9, 99), (3, 4, 4, 5), (2, 1, 11, 3), (3, 1, 6, 8)])
def test_cuda_matches_oracle_on_seeded_random_inputs(lib, dim, p, m, seed):
    rng = np.random.default_rng(seed)
    prog = g.expr_compile("exp(-x)*cos(3*y)+z^2-0.3")
    pb0 = host.poisson_box_problem(dim, p, [m, m + 1, m + 2][:dim], prog)
    pa = pb0.patches[0]
    # curved geometry: perturbed control points, non-uniform knots, random Dirichlet data
    coefs = pa.geo_coefs + rng.uniform(-0.08, 0.08, pa.geo_coefs.shape)
    knots = []
    for k in range(dim):
        kn = pa.space_knots[k].copy()
        inner = slice(p + 1, len(kn) - p - 1)
        h = 1.0 / len(kn)
        kn[inner] = np.sort(kn[inner] + rng.uniform(-0.3 * h, 0.3 * h, kn[inner].shape))
        knots.append(kn)
    patch = capi.PatchData(pa.space_degree, knots, pa.geo_degree, pa.geo_knots, coefs, pa.dofmap)
    fixed = rng.uniform(-1, 1, (pb0.nfixed, 1))
    pb = capi.Problem([patch], pb0.nfree, pb0.nfixed, fixed=fixed, rhs_programs=[prog])
    ok, msg = R.compare_csc(R.lib_assemble(lib, pb), R.oracle_assemble(pb), TOL)
    assert ok, msg


def test_chunked_equals_unchunked_bitwise_and_deterministic(lib):
    pb, z = G.load("cube_p3_m16", g.expr_compile)
    a = R.lib_assemble(lib, pb)
    b = R.lib_assemble(lib, pb, workspace_limit=30_000_000)
    c = R.lib_assemble(lib, pb)
    assert b[4].nchunks > 1
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], c[2]) and np.array_equal(a[3], c[3])
    G.check_against(b, z, TOL)


@pytest.mark.parametrize("name,nranks", [("cube_p3_curved_m4", 2), ("cube_p3_m16", 4)])
def test_rank_slabs_cover_the_matrix(lib, name, nranks):
    pb0, z = G.load(name, g.expr_compile)
    n = pb0.nfree
    nnz = 0; rhs = np.zeros((n, 1)); values = []; inner = []; lens = np.zeros(n, np.int64)
    for r in range(nranks):
        pb, _ = G.load(name, g.expr_compile)
        pb.struct.rank, pb.struct.nranks = r, nranks
        o, i, v, b, _ = R.lib_assemble(lib, pb)
        ln = np.diff(o)
        assert not np.any((ln > 0) & (lens > 0)), "a column is owned by two ranks"
        lens += ln; values.append(v); inner.append(i); rhs += b
    outer = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    G.check_against((outer, np.concatenate(inner), np.concatenate(values), rhs), z, TOL)


def _gauss_1d_matrices(p, m):
    """1-D mass/stiffness matrices and load factors of the uniform open knot vector, computed
    with scipy's B-splines and the same q=p+1 Gauss rule — independent of the oracle and of the kernels."""
    from scipy.interpolate import BSpline
    kv = host.KnotVector.open_uniform(0, 1, 0, 2); kv.setDegree(p); kv.uniformRefine(m - 1)
    t = kv.knots; n = kv.size
    xg, wg = np.polynomial.legendre.leggauss(p + 1)
    br = np.unique(t)
    pts = np.concatenate([(b - a) / 2 * (xg + 1) + a for a, b in zip(br[:-1], br[1:])])
    wts = np.concatenate([(b - a) / 2 * wg for a, b in zip(br[:-1], br[1:])])
    B = BSpline.design_matrix(pts, t, p).toarray()
    dB = np.zeros_like(B)
    for i in range(n):
        c = np.zeros(n); c[i] = 1.0
        dB[:, i] = BSpline(t, c, p)(pts, nu=1)
    M = B.T @ (wts[:, None] * B); K1 = dB.T @ (wts[:, None] * dB)
    load = B.T @ (wts * np.sin(np.pi * (pts - 0.5)))      # geometry is the cube centred at 0: x = xi - 0.5
    return M, K1, load, n


def test_full_size_config2_properties(lib):
    """BASELINE config 2: 3-D unit cube, p=3, 125^3 elements -> 2 000 376 DOFs, nnz = 870^3."""
    p, m = 3, 125
    prog = g.expr_compile("3*pi^2*sin(pi*x)*sin(pi*y)*sin(pi*z)")
    pb = host.poisson_box_problem(3, p, m, prog)
    A = g.DeviceAssembler(pb)
    nnz = A.buildPattern()
    assert pb.nfree == 126 ** 3 and nnz == 870 ** 3          # SURVEY 8(a) sparsity facts
    A.assemble()
    outer, inner, values = A.matrix()
    rhs = A.rhs()[:, 0]
    A.assemble()                                             # determinism: second pass is bit-identical
    o2, i2, v2 = A.matrix()
    assert np.array_equal(values, v2)
    del o2, i2, v2
    A.close()
    M, K1, load, n1 = _gauss_1d_matrices(p, m)
    nf = n1 - 2                                              # free functions per direction
    assert outer[-1] == nnz and np.all(np.diff(outer) <= (2 * p + 1) ** 3)
    # separable load vector: rhs_i = 3 pi^2 prod_k load[i_k]
    li = load[1:-1]
    exp_rhs = 3 * np.pi ** 2 * np.einsum("k,j,i->kji", li, li, li).ravel()
    assert np.abs(rhs - exp_rhs).max() <= TOL * np.abs(exp_rhs).max()
    # Kronecker structure on sampled columns (direction 0 fastest in the numbering)
    rng = np.random.default_rng(2024)
    scale = np.abs(values[:: 997]).max()
    cols = np.concatenate([[0, 1, nf, nf * nf + 3, pb.nfree - 1], rng.integers(0, pb.nfree, 400)])
    for c in cols:
        c0, c1, c2 = c % nf + 1, (c // nf) % nf + 1, c // (nf * nf) + 1
        rows = inner[outer[c]:outer[c + 1]]
        assert np.all(np.diff(rows) > 0)
        r0, r1, r2 = rows % nf + 1, (rows // nf) % nf + 1, rows // (nf * nf) + 1
        exp = K1[r0, c0] * M[r1, c1] * M[r2, c2] + M[r0, c0] * K1[r1, c1] * M[r2, c2] + M[r0, c0] * M[r1, c1] * K1[r2, c2]
        assert np.abs(values[outer[c]:outer[c + 1]] - exp).max() <= TOL * scale
        want = sum(1 for a in range(max(1, c0 - p), min(nf, c0 + p) + 1)) * \
            sum(1 for a in range(max(1, c1 - p), min(nf, c1 + p) + 1)) * sum(1 for a in range(max(1, c2 - p), min(nf, c2 + p) + 1))
        assert len(rows) == want
    # symmetry of sampled entries: K[r,c] == K[c,r] to round-off
    for c in cols[:50]:
        for k in range(outer[c], outer[c + 1], 37):
            r = inner[k]
            seg = inner[outer[r]:outer[r + 1]]
            pos = outer[r] + np.searchsorted(seg, c)
            assert inner[pos] == c and abs(values[pos] - values[k]) <= TOL * scale


def test_consumer_spmv_and_cg(lib):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    pb, z = G.load("cube_p3_m16", g.expr_compile)
    A = g.DeviceAssembler(pb)
    A.assemble()
    K = A.scipy_matrix()
    x = G.probe_vector(pb.nfree)
    y = A.spmv(x)
    assert np.abs(y - K @ x).max() <= 1e-12 * np.abs(y).max()
    b = A.rhs()[:, 0]
    u, iters, res = A.cg(b, max_iter=2000, tol=1e-12)
    uref = spl.spsolve(K.tocsc(), b)
    assert res <= 1e-11 and iters < 2000
    assert np.abs(u - uref).max() <= 1e-8 * np.abs(uref).max()
    # manufactured solution u = sin(pi x) sin(pi y) sin(pi z) on the cube centred at 0 -> cos products; just sanity
    assert np.all(np.isfinite(u))
    nreg, ntab = A.spmv_info()
    nf = round(pb.nfree ** (1 / 3))
    assert ntab == 1 and nreg == (nf - 6) ** 3                # interior columns of the cube take the 8-byte path
    v = A.device_view()
    assert (v.col_begin, v.col_end) == (0, pb.nfree)
    A.close()


@pytest.mark.parametrize("name", ["grid2x2_p2_m4", "elasticity_2cubes_p2", "yeti_mp2_p2_m2", "annulus_nurbs_p3_m6"])
def test_spmv_regular_and_general_columns(lib, name):
    """y = K x against scipy on multi-patch / vector-valued / NURBS fixtures: columns matching a reference stencil read no
    row indices, all others do; same product."""
    pb, z = G.load(name, g.expr_compile)
    A = g.DeviceAssembler(pb)
    A.assemble()
    K = A.scipy_matrix()
    x = G.probe_vector(pb.nfree)
    y = A.spmv(x)
    assert np.abs(y - K @ x).max() <= 1e-12 * np.abs(y).max()
    nreg, ntab = A.spmv_info()
    assert 0 <= nreg <= pb.nfree and ntab >= 1
    b = A.rhs()[:, 0]
    if pb.nfixed > 0 and np.linalg.norm(b) > 0:          # (a pure Neumann problem is singular)
        u, iters, res = A.cg_solve(b, max_iter=5000, tol=1e-11, check_every=7)
        assert res <= 1e-11
        assert np.linalg.norm(K @ u - b) <= 1e-9 * np.linalg.norm(b)
    A.close()


def _torchrun(nproc, script, *args, timeout=600):
    import subprocess, sys, socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script, *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("nproc", [2, 4])
def test_nccl_exchange_and_distributed_cg(lib, nproc):
    """N ranks, one per GPU, the library's own NCCL communicator: patch-wise ownership with the coupled-column exchange
    (grid2x2, yeti_mp2, 8 cubes, elasticity) and column slabs with the neighbour-halo CG (single patches) against the reference
    fixtures.  Needs N GPUs on the box (tests/nccl_worker.py)."""
    n = C.c_int(0); lib.gsb200_device_count(C.byref(n))
    if n.value < nproc:
        pytest.skip(f"{nproc} GPUs needed, {n.value} visible")
    here = os.path.dirname(os.path.abspath(__file__))
    r = _torchrun(nproc, os.path.join(here, "nccl_worker.py"), "grid2x2_p2_m4", "yeti_mp2_p2_m2", "grid2x2x2_p2_m3", "elasticity_2cubes_p2",
                  "cube_p3_m16", "cube_p3_curved_m4")
    assert r.returncode == 0 and "NCCLWORKER ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_compiled_source_term_equals_interpreter(lib):
    """Repeated assemblies switch the geometry kernel to the NVRTC-compiled source term (csrc/jit.cuh).  The compiled term uses
    the interpreter's operations one by one (__dmul_rn, ... : nothing is contracted), so the load vector agrees to the rounding
    of its atomic accumulation order; the matrix does not depend on the source term and stays bit-identical."""
    text = "3*pi^2*sin(pi*x)*sin(pi*y)*sin(pi*z)+0.125*x*y-exp(-z*z)/7"        # a program no other test uses: its first launches interpret
    pb = host.poisson_box_problem(3, 3, 9, g.expr_compile(text))
    A = g.DeviceAssembler(pb)
    A.assemble()
    first = A.matrix() + (A.rhs(),)
    used_first = A.jit_launches()
    for _ in range(3):
        A.assemble()
    last = A.matrix() + (A.rhs(),)
    used_last = A.jit_launches()
    A.close()
    ok, msg = R.compare_csc(last, R.oracle_assemble(pb), TOL)
    assert ok, msg
    if os.environ.get("GSB200_JIT", "1") != "1" or used_last == 0:
        pytest.skip("NVRTC path not active (GSB200_JIT set or libnvrtc missing)")
    assert used_first == 0 and used_last > 0
    assert np.array_equal(first[2], last[2])
    assert np.abs(first[3] - last[3]).max() <= 1e-14 * np.abs(first[3]).max()


@pytest.mark.parametrize("dim,p,m,quA,quB", [(3, 2, 6, 1.0, 2), (3, 3, 5, 1.0, 2), (2, 2, 9, 2.0, 1), (3, 1, 7, 1.0, 2)])
def test_other_quadrature_sizes(lib, dim, p, m, quA, quB):
    """q != p+1 Gauss points per direction (quA/quB options): the generic / ring kernels on the layouts chosen for the window kernels."""
    pb = host.poisson_box_problem(dim, p, m, g.expr_compile("1+x*y" if dim == 2 else "1+x*y-z"))
    pb.struct.quA, pb.struct.quB = quA, quB
    ok, msg = R.compare_csc(R.lib_assemble(lib, pb), R.oracle_assemble(pb), TOL)
    assert ok, msg


def test_measured_peaks_are_plausible(lib):
    pk = g.measure_peaks(0)
    assert 5 < pk["fp64_tflops"] < 100 and 500 < pk["hbm_gbs"] < 10000


def test_gismo_shim_dropin(lib):
    """The C++ drop-in shims (gismo_b200/host/*.h) on real gismo objects next to the unmodified reference
    assemblers: tests/shim/shim_test.cpp, built here against /root/reference, executed on the GPU box."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim", "_build", "shim_test")
    if not os.path.exists(exe):
        pytest.skip("tests/shim/_build/shim_test not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0 and "SHIM RESULT PASS" in out.stdout


@pytest.mark.parametrize("name", ["cube_p2_m5", "cube_p3_m16", "cube_p3_curved_m4", "elasticity_8cubes_p2_m5"])
def test_fused_second_and_last_sweep_opt_in(lib, name, monkeypatch):
    """GSB200_S23=1 (opt-in, profiles/r02_s23_experiment.txt): same matrix as the default path and the reference."""
    monkeypatch.setenv("GSB200_S23", "1")
    pb, z = G.load(name, g.expr_compile)
    G.check_against(R.lib_assemble(lib, pb), z, TOL)


@pytest.mark.parametrize("name,chunks", [("cube_p3_m16", 3), ("cube_p3_curved_m4", 8), ("sq_p2_m64", 4), ("grid2x2_p2_m4", 4)])
def test_values_only_reassembly_with_streamed_column_ranges(lib, name, chunks, monkeypatch):
    """gsb200_assemble_to_host, then gsb200_set_fixed + gsb200_assemble_values_to_host on the kept pattern (what
    gsPoissonAssemblerB200::setKeepPattern drives): finished column ranges travel while later chunks integrate.  Same values
    bit for bit as a one-piece assembly of the changed problem; multi-patch problems fall back to one delivery."""
    monkeypatch.setenv("GSB200_DELIVER_CHUNKS", str(chunks))
    monkeypatch.setenv("GSB200_DELIVER_MIN_NNZ", "0")
    pb, z = G.load(name, g.expr_compile)
    rng = np.random.default_rng(5)
    fixed2 = rng.uniform(-1, 1, (pb.nfixed, pb.nrhs))
    first, again = R.lib_reassemble(lib, pb, fixed2)
    G.check_against(first, z, TOL)
    assert first[4].nchunks > 1 or len(pb.patches) > 1         # (nchunks counts one per patch at least)
    ref = R.lib_assemble(lib, pb.with_fixed(fixed2))
    assert np.array_equal(again[2], ref[2])
    ok, msg = R.compare_csc(again, ref, 1e-13)
    assert ok, msg


def test_curved_geometry_at_a_million_dofs(lib):
    """3-D p=3 on the curved cube, 100^3 elements, 1.09 M DOFs, interpolated Dirichlet data: the reference ran this here
    (tests/golden/make_golden.py, SAMPLED) and left samples + norms of K x, diag K, rhs and checksums of the index arrays."""
    if "cube_p3_curved_m50" not in G.names("sampled"):
        pytest.skip("sampled fixture not generated")
    pb, z = G.load("cube_p3_curved_m50", g.expr_compile)
    assert pb.nfree == 102 ** 3 and int(z["nnz"]) == 700 ** 3
    G.check_against(R.lib_assemble(lib, pb), z, TOL)


def test_determinism_with_inhomogeneous_dirichlet_data(lib):
    """Repeated assemblies with non-zero eliminated values: every matrix entry has exactly one owner thread, so the values are
    bit-identical; the elimination terms -K g are accumulated on the owner's right-hand-side entry with atomicAdd, whose order is
    free: the right-hand side may move by a few ulp of its largest term (bound asserted: 1e-14 of max|rhs|)."""
    pb, z = G.load("cube_p3_curved_m4", g.expr_compile)
    runs = [R.lib_assemble(lib, pb) for _ in range(4)]
    for r in runs[1:]:
        assert np.array_equal(r[2], runs[0][2]) and np.array_equal(r[1], runs[0][1])
        assert np.abs(r[3] - runs[0][3]).max() <= 1e-14 * np.abs(runs[0][3]).max()
    G.check_against(runs[-1], z, TOL)
    # multi-patch: coupled interface columns are summed with atomicAdd as well
    pb, z = G.load("grid2x2x2_p2_m3", g.expr_compile)
    a, b = R.lib_assemble(lib, pb), R.lib_assemble(lib, pb)
    assert np.abs(a[2] - b[2]).max() <= 1e-14 * np.abs(a[2]).max() and np.abs(a[3] - b[3]).max() <= 1e-14 * np.abs(a[3]).max()


@pytest.mark.parametrize("mode", ["1", "3"])
def test_first_sweep_output_full_and_half_rows(lib, mode, monkeypatch):
    """GSB200_A1BLK=1: the first sweep stores all 2p+1 deltas per function; 3 (opt-in, 3-D degree 3; measured slower, profiles/r02_a1_half_experiment.txt): delta >= 0 only and the second
    sweep reads the rest at the mirrored pair (43 % less HBM traffic for A1).  Same matrix either way, also chunked."""
    monkeypatch.setenv("GSB200_A1BLK", mode)
    pb, z = G.load("cube_p3_curved_m4", g.expr_compile)
    full = R.lib_assemble(lib, pb)
    G.check_against(full, z, TOL)
    capped = R.lib_assemble(lib, pb, workspace_limit=8_000_000)
    assert capped[4].nchunks > 1 and np.array_equal(capped[2], full[2])
    pb16, z16 = G.load("cube_p3_m16", g.expr_compile)
    G.check_against(R.lib_assemble(lib, pb16), z16, TOL)


@pytest.mark.parametrize("name", sorted(G.NORM_CASES))
def test_field_norms_match_the_reference(lib, name):
    """f4: gsb200_field_norms on the reference's own solution against its gsExprEvaluator integrals (L2 / H1 error, poisson2_example.cpp:174-177),
    and against the C oracle on a random field."""
    pb, z = G.load(name, g.expr_compile)
    ex, grads = G.NORM_CASES[name]
    exp, gp = g.expr_compile(ex), [g.expr_compile(t) for t in grads]
    A = g.DeviceAssembler(pb)
    G.check_norms(A.field_norms(z["solution"], exp, gp), z)
    u = np.random.default_rng(3).uniform(-1, 1, pb.nfree)
    a, b = A.field_norms(u, exp, gp), R.oracle_field_norms(pb, u, exp, gp)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    # the device solve reproduces the reference's solution and therefore its error norms
    A.assemble()
    x, it, res = A.cg_solve(A.rhs()[:, 0], max_iter=5000, tol=1e-13, check_every=5)
    assert np.abs(x - z["solution"]).max() <= 1e-9 * np.abs(z["solution"]).max()
    A.close()


@pytest.mark.parametrize("name", sorted(G.L2PROJ_CASES))
def test_dirichlet_l2_projection_matches_the_reference(lib, name):
    """f2: eliminated-DOF values by L2-projection on the device (gsb200_project_dirichlet) against gsDirichletValuesByL2Projection /
    computeDirichletDofsL2Proj of the reference (2-D, 3-D, NURBS, with Neumann sides, across patches); then the assembly with them."""
    pb, z = G.load(name, g.expr_compile)
    text, sides = G.l2proj_sides(name, pb)
    A = g.DeviceAssembler(pb.with_fixed(None))
    data = [g.expr_compile(t) for t in text] if isinstance(text, list) else g.expr_compile(text)      # vector-valued: one datum per component
    fx, it, res = A.project_dirichlet([(p_, s_, data) for p_, s_ in sides])
    ref = z["fixed"][:, 0]
    assert res <= 1e-12 and np.abs(fx - ref).max() <= 1e-9 * np.abs(ref).max()
    A.assemble()
    G.check_against(A.matrix() + (A.rhs(),), z, 1e-9)
    A.close()
