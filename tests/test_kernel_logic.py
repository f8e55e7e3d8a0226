"""Index logic of the CUDA kernels, validated without a GPU.

tests/emul builds the very kernel bodies of gismo_b200/csrc with -DGSB200_EMULATE (a single-
threaded interpreter of the launch grid; test harness only, never loaded by the package) and
the same host orchestration.  These tests pin: pattern build, slot tables, sweep segments,
chunking under a workspace cap, slab partitioning across ranks, all against the reference's
fixtures.  The same cases run on the real device in test_gpu_parity.py.
"""
import ctypes as C

import numpy as np
import pytest

import goldenutil as G
import refutil as R
from gismo_b200 import capi

TOL = 1e-12


@pytest.fixture(scope="module")
def emul():
    return R.emul_lib()


@pytest.mark.parametrize("name", G.names("full") + ["sq_p2_m64", "cube_p2_m16_expr", "yeti_mp2_p2_m8", "elasticity_8cubes_p2_m5", "poisson2d_bvp_stock_r6"])
def test_interpreted_kernels_match_reference(emul, name):
    pb, z = G.load(name, R.emul_compile)
    G.check_against(R.lib_assemble(emul, pb), z, TOL)


@pytest.mark.parametrize("name,limit", [("cube_p3_curved_m4", 7_000_000), ("sq_p2_m64", 300_000), ("grid2x2_p2_m4", 11_000)])
def test_chunked_assembly_under_workspace_cap(emul, name, limit):
    pb, z = G.load(name, R.emul_compile)
    res = R.lib_assemble(emul, pb, workspace_limit=limit)
    assert res[4].nchunks > len(pb.patches), "cap did not force chunking"
    G.check_against(res, z, TOL)


def test_workspace_cap_too_small_is_an_error(emul):
    pb, _ = G.load("cube_p2_m5", R.emul_compile)
    with pytest.raises(RuntimeError, match="workspace limit"):
        R.lib_assemble(emul, pb, workspace_limit=1000)


@pytest.mark.parametrize("name,nranks", [("cube_p3_curved_m4", 2), ("cube_p2_m5", 3), ("sq_p2_m8_visitor", 4)])
def test_slab_partition_over_ranks(emul, name, nranks):
    """Each rank owns a slab of CSC columns; the union is the reference matrix, no overlap."""
    pb0, z = G.load(name, R.emul_compile)
    n = pb0.nfree
    outer = np.zeros(n + 1, np.int64); cols = [None] * n; rhs = np.zeros((n, 1))
    for r in range(nranks):
        pb, _ = G.load(name, R.emul_compile)
        pb.struct.rank, pb.struct.nranks = r, nranks
        o, i, v, b, _ = R.lib_assemble(emul, pb)
        for c in range(n):
            if o[c + 1] > o[c]:
                assert cols[c] is None, "column owned twice"
                cols[c] = (i[o[c]:o[c + 1]], v[o[c]:o[c + 1]])
        rhs += b
    assert all(c is not None for c in cols)
    inner = np.concatenate([c[0] for c in cols]); values = np.concatenate([c[1] for c in cols])
    outer[1:] = np.cumsum([len(c[0]) for c in cols])
    G.check_against((outer.astype(np.int32), inner, values, rhs), z, TOL)


def test_bad_inputs_are_rejected(emul):
    pb, _ = G.load("cube_p2_m5", R.emul_compile)
    h = C.c_void_p()
    pb.struct.abi_version = 99
    assert emul.gsb200_create(C.byref(pb.struct), 0, C.byref(h)) == -1
    pb.struct.abi_version = capi.ABI_VERSION
    pb.struct.form = 7
    assert emul.gsb200_create(C.byref(pb.struct), 0, C.byref(h)) == -2
    pb.struct.form = 0
    pb.patches[0].dofmap[3] = 10 ** 6
    assert emul.gsb200_create(C.byref(pb.struct), 0, C.byref(h)) == -1
    assert b"out of range" in emul.gsb200_last_error()


def test_assemble_before_pattern_is_a_state_error(emul):
    pb, _ = G.load("cube_p2_m5", R.emul_compile)
    h = C.c_void_p()
    assert emul.gsb200_create(C.byref(pb.struct), 0, C.byref(h)) == 0
    assert emul.gsb200_assemble(h) == -7
    emul.gsb200_destroy(h)


@pytest.mark.parametrize("dim,p,m,quA,quB", [(3, 2, 4, 1.0, 2), (3, 3, 3, 1.0, 2), (2, 2, 6, 2.0, 1), (3, 1, 5, 1.0, 2)])
def test_other_quadrature_sizes_take_the_generic_kernels(emul, dim, p, m, quA, quB):
    """q != p+1 points per direction (gsQuadrature.h:152-171 with quA/quB options) leaves the window kernels' fast path: the generic
    sweep must still honour the layouts chosen for them.  Checked against the C oracle (element loop)."""
    from gismo_b200 import host
    pb = host.poisson_box_problem(dim, p, m, R.emul_compile("1+x*y" if dim == 2 else "1+x*y-z"))
    pb.struct.quA, pb.struct.quB = quA, quB
    ok, msg = R.compare_csc(R.lib_assemble(emul, pb), R.oracle_assemble(pb), TOL)
    assert ok, msg


@pytest.mark.parametrize("name", ["cube_p1_m5", "cube_p2_m5", "cube_p3_curved_m4", "grid2x2x2_p2_m3", "elasticity_2cubes_p2"])
def test_fused_second_and_last_sweep(emul, name, monkeypatch):
    """GSB200_S23=1: directions 1 and 2 contracted in one kernel (fused23.cuh), rows of A2 through shared memory."""
    monkeypatch.setenv("GSB200_S23", "1")
    pb, z = G.load(name, R.emul_compile)
    G.check_against(R.lib_assemble(emul, pb), z, TOL)


@pytest.mark.parametrize("name,chunks", [("cube_p3_m16", 3), ("cube_p2_m5", 8), ("sq_p2_m64", 4)])
def test_values_only_reassembly_with_streamed_column_ranges(emul, name, chunks, monkeypatch):
    """gsb200_assemble_to_host, then gsb200_set_fixed + gsb200_assemble_values_to_host on the kept pattern: the last direction
    is cut into delivery chunks (finished column ranges travel first); same numbers as one-piece assemblies."""
    monkeypatch.setenv("GSB200_DELIVER_CHUNKS", str(chunks))
    monkeypatch.setenv("GSB200_DELIVER_MIN_NNZ", "0")
    pb, z = G.load(name, R.emul_compile)
    rng = np.random.default_rng(5)
    fixed2 = rng.uniform(-1, 1, (pb.nfixed, pb.nrhs))
    first, again = R.lib_reassemble(emul, pb, fixed2)
    G.check_against(first, z, TOL)
    assert first[4].nchunks > 1
    pb2, _ = G.load(name, R.emul_compile)
    pb2 = pb2.with_fixed(fixed2)
    ref = R.lib_assemble(emul, pb2)
    assert np.array_equal(again[2], ref[2])          # chunking never changes a bit (every entry has one owner thread)
    ok, msg = R.compare_csc(again, ref, 1e-14)       # rhs: the -K g terms are accumulated atomically
    assert ok, msg


@pytest.mark.parametrize("mode", ["1", "3"])
def test_first_sweep_output_full_and_half_rows(emul, mode, monkeypatch):
    """GSB200_A1BLK=1: A1 keeps all 2p+1 deltas per function; 3 (opt-in, 3-D p=3): delta >= 0 only, the second sweep reads the
    others at the mirrored pair (terms.cuh T3SymS2U).  Same matrix either way; also under a workspace cap and split over ranks."""
    monkeypatch.setenv("GSB200_A1BLK", mode)
    pb, z = G.load("cube_p3_curved_m4", R.emul_compile)
    full = R.lib_assemble(emul, pb)
    G.check_against(full, z, TOL)
    capped = R.lib_assemble(emul, pb, workspace_limit=8_000_000)
    assert capped[4].nchunks > 1 and np.array_equal(capped[2], full[2])
    pb16, z16 = G.load("cube_p3_m16", R.emul_compile)
    G.check_against(R.lib_assemble(emul, pb16), z16, TOL)


@pytest.mark.parametrize("name", sorted(G.NORM_CASES))
def test_field_norm_kernel_matches_the_reference(emul, name):
    """gsb200_field_norms (k_field_norms, interpreted) against the reference's gsExprEvaluator integrals and the C oracle."""
    import gismo_b200 as g
    pb, z = G.load(name, R.emul_compile)
    ex, grads = G.NORM_CASES[name]
    exp, gp = R.emul_compile(ex), [R.emul_compile(t) for t in grads]
    A = g.DeviceAssembler(pb, lib=emul)
    got = A.field_norms(z["solution"], exp, gp)
    G.check_norms(got, z)
    u = np.random.default_rng(3).uniform(-1, 1, pb.nfree)
    a, b = A.field_norms(u, exp, gp), R.oracle_field_norms(pb, u, exp, gp)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    A.close()


@pytest.mark.parametrize("name", sorted(G.L2PROJ_CASES))
def test_dirichlet_l2_projection_matches_the_reference(emul, name):
    """f2: gsb200_project_dirichlet (matrix-free boundary mass + Jacobi-CG, interpreted) against the eliminated values the reference
    computed with gsDirichletValuesByL2Projection / computeDirichletDofsL2Proj; then the assembly with those values."""
    import gismo_b200 as g
    pb, z = G.load(name, R.emul_compile)
    text, sides = G.l2proj_sides(name, pb)
    A = g.DeviceAssembler(pb.with_fixed(None), lib=emul)          # no eliminated values given: the device computes them
    data = [R.emul_compile(t) for t in text] if isinstance(text, list) else R.emul_compile(text)      # vector-valued: one datum per component
    fx, it, res = A.project_dirichlet([(p_, s_, data) for p_, s_ in sides])
    ref = z["fixed"][:, 0]
    assert res <= 1e-12 and np.abs(fx - ref).max() <= 1e-9 * np.abs(ref).max()
    A.assemble()
    G.check_against(A.matrix() + (A.rhs(),), z, 1e-9)            # the rhs carries -K g with the projected g
    A.close()


def test_patches_are_balanced_over_ranks_longest_first(emul):
    """gsb200_create distributes whole patches over the ranks: longest processing time first onto the least loaded rank (SURVEY 8e:
    the 21 yeti patches over 8 GPUs).  The library's choice (which columns a rank stores) must be the one of distributed.patch_owners,
    and the ranks' pieces together give the reference matrix."""
    from gismo_b200 import distributed as D
    pb0, z = G.load("yeti_mp2_p2_m2", R.emul_compile)
    nranks = 8
    costs = []
    for pa in pb0.patches:
        c = 1
        for k in range(pb0.dim):
            c *= (len(np.unique(pa.space_knots[k])) - 1) * (pa.space_degree[k] + 1)
        costs.append(c)
    owner = D.patch_owners(costs, nranks)
    assert max(np.bincount(owner, minlength=nranks)) - min(np.bincount(owner, minlength=nranks)) <= 1      # equal patches: 3/3/3/3/3/2/2/2
    counts = np.zeros(pb0.nfree + 1, np.int32)
    for pa in pb0.patches:
        np.add.at(counts, pa.dofmap[pa.dofmap < pb0.nfree], 1)
    pieces = []
    for r in range(nranks):
        pb = pb0.with_fixed(pb0.fixed, rank=r, nranks=nranks)
        o, i, v, b, _ = R.lib_assemble(emul, pb)
        stored = np.diff(o) > 0
        mine = np.zeros(pb0.nfree, bool)
        for ip, pa in enumerate(pb0.patches):
            if owner[ip] == r:
                g_ = pa.dofmap[pa.dofmap < pb0.nfree]
                mine[g_] = True
        single = counts[:pb0.nfree] == 1
        assert np.array_equal(stored & single, mine & single), f"rank {r} stores other patches than the balanced assignment gives it"
        assert np.all(stored[~single & (counts[:pb0.nfree] > 1)]), "coupled columns are patterned on every rank"
        pieces.append((o, i, v, b))
    # sum the coupled columns by hand (what gsb200_exchange does) and compare with the reference
    lens = np.max([np.diff(p[0]) for p in pieces], axis=0)
    outer = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    values = np.zeros(outer[-1]); inner = np.zeros(outer[-1], np.int32); rhs = np.zeros_like(pieces[0][3])
    for o, i, v, b in pieces:
        rhs += b
        for c in np.nonzero(np.diff(o) > 0)[0]:
            values[outer[c]:outer[c + 1]] += v[o[c]:o[c + 1]]
            inner[outer[c]:outer[c + 1]] = i[o[c]:o[c + 1]]
    G.check_against((outer, inner, values, rhs), z, TOL)


def test_consumer_entry_points_reject_bad_states(emul):
    """Error behaviour of the round-2 entry points (status codes, not crashes): exchange / CG before an assembly, field norms of a
    vector-valued space, malformed Dirichlet sides, a multi-rank exchange without a communicator."""
    import ctypes as C
    import gismo_b200 as g
    from gismo_b200 import capi
    pb, _ = G.load("cube_p2_m5", R.emul_compile)
    A = g.DeviceAssembler(pb, lib=emul)
    assert emul.gsb200_exchange(A._h) == -7                       # GSB200_ESTATE: nothing assembled yet
    it, res = C.c_int(0), C.c_double(0)
    assert emul.gsb200_cg_solve(A._h, None, None, 10, 1e-8, 5, C.byref(it), C.byref(res)) == -7
    A.assemble()
    A.exchange()                                                  # one rank: a no-op
    assert A.comm_stats() == (0, 0) or A.comm_stats()[0] == 0
    bad = (capi.Neumann * 1)()
    bad[0].patch, bad[0].side, bad[0].ndata = 0, 9, 1             # side out of range
    out = np.zeros(pb.nfixed)
    assert emul.gsb200_project_dirichlet(A._h, bad, 1, 10, 1e-8, out.ctypes.data_as(C.POINTER(C.c_double)), None, None) == -1
    A.close()
    pe, _ = G.load("elasticity_sq_p2", R.emul_compile)
    E = g.DeviceAssembler(pe, lib=emul)
    with pytest.raises(capi.Gsb200Error):
        E.field_norms(np.zeros(pe.nfree))                         # scalar spaces only
    E.close()
    p2 = pb.with_fixed(pb.fixed, rank=0, nranks=2)
    B = g.DeviceAssembler(p2, lib=emul)
    B.assemble()
    with pytest.raises(capi.Gsb200Error):
        B.exchange()                                              # two ranks, no communicator / callback: refused, not skipped
    B.close()
