"""Host side: C-ABI library surface, expression compiler, problem builders (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import goldenutil as G
import refutil as R
import gismo_b200 as g
from gismo_b200 import capi, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gsb200.h")).read()
    declared = set(re.findall(r"\b(gsb200_[a-z0-9_]+)\s*\(", hdr))
    lib = capi.load_library()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/gsb200.h but not exported"
    assert lib.gsb200_abi_version() == capi.ABI_VERSION


def test_struct_layout_matches_header():
    # sizes computed by the C compiler for the same header
    src = '#include "gsb200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",sizeof(gsb200_basis),sizeof(gsb200_patch),sizeof(gsb200_program),sizeof(gsb200_problem),sizeof(gsb200_device_view),sizeof(gsb200_timings));}'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(capi.Basis), C.sizeof(capi.Patch), C.sizeof(capi.Program), C.sizeof(capi.ProblemStruct),
                     C.sizeof(capi.DeviceView), C.sizeof(capi.Timings)]


def test_no_cpu_fallback_without_device():
    lib = capi.load_library()
    n = C.c_int(0)
    lib.gsb200_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a device is present")
    pb = host.poisson_box_problem(2, 2, 4)
    with pytest.raises(g.Gsb200Error, match="no CUDA device"):
        g.DeviceAssembler(pb)
    with pytest.raises(g.Gsb200Error, match="no CUDA device"):
        g.assemble_host(pb)


@pytest.mark.parametrize("text,pt,val", [
    ("2*pi^2*sin(pi*x)*sin(pi*y)", (0.3, 0.7, 0.0), 2 * np.pi ** 2 * np.sin(np.pi * 0.3) * np.sin(np.pi * 0.7)),
    ("-x^2 + 3*(y - z)/2", (1.5, 2.0, 0.5), -2.25 + 2.25),
    ("exp(-x)*cos(y)+sqrt(abs(z))", (0.2, 1.0, -4.0), np.exp(-0.2) * np.cos(1.0) + 2.0),
    ("1e-3*x + .5", (2.0, 0, 0), 0.502),
    ("tanh(x)-sinh(y)*cosh(z)+log(2)+tan(0.1)", (0.3, 0.2, 0.1), np.tanh(0.3) - np.sinh(0.2) * np.cosh(0.1) + np.log(2) + np.tan(0.1)),
])
def test_expression_compiler(text, pt, val):
    prog = g.expr_compile(text)
    assert abs(capi.expr_eval(prog, *pt) - val) < 1e-14 * max(1, abs(val))


@pytest.mark.parametrize("bad", ["foo(x)", "x +", "(x", "x y", "w"])
def test_expression_compiler_rejects(bad):
    with pytest.raises(g.Gsb200Error):
        g.expr_compile(bad)


def test_uniform_refine_matches_reference_expression():
    kv = host.KnotVector.open_uniform(0, 1, 0, 2)
    kv.setDegree(3)
    kv.uniformRefine(124)
    assert kv.size == 128 and len(np.unique(kv.knots)) == 126
    z = dict(np.load(G.GOLDEN + "/cube_p3_m16.npz"))
    kv = host.KnotVector.open_uniform(0, 1, 0, 2); kv.setDegree(3); kv.uniformRefine(15)
    assert np.array_equal(kv.knots, z["p0_sk0"])      # bit-identical to gsKnotVector::uniformRefine
    if R.have_ref():
        lib = R.ref_lib()
        base = np.array([0, 0, 0, 0.3, 0.3, 1, 1, 1.0])
        out = np.zeros(64); n = C.c_int(0)
        lib.gsref_uniform_refine(base.ctypes.data_as(C.POINTER(C.c_double)), len(base), 2, 3, out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n))
        kv = host.KnotVector(2, base); kv.uniformRefine(3)
        assert np.array_equal(kv.knots, out[:n.value])


@pytest.mark.parametrize("name,dim,p,m", [("cube_p2_m5", 3, 2, 5), ("sq_p2_m8_visitor", 2, 2, 8), ("cube_p1_m5", 3, 1, 5)])
def test_box_builder_reproduces_reference_flattening(name, dim, p, m):
    z = dict(np.load(G.GOLDEN + f"/{name}.npz"))
    pb = host.poisson_box_problem(dim, p, m)
    assert pb.nfree == int(z["nfree"]) and pb.nfixed == int(z["nfixed"])
    assert np.array_equal(pb.patches[0].dofmap, z["p0_dofmap"])
    for i in range(dim):
        assert np.array_equal(pb.patches[0].space_knots[i], z[f"p0_sk{i}"])
    assert np.array_equal(pb.patches[0].geo_coefs, z["p0_coefs"])


def test_dof_mapper_coupling_order():
    # two 1-D "patches" of 4 functions glued end to start, ends eliminated (gsDofMapper.cpp:281-323)
    m = host.DofMapper([4, 4])
    m.match(0, 3, 1, 0)
    m.eliminate(0, [0]); m.eliminate(1, [3])
    m.finalize()
    assert m.nfree == 5 and m.nfixed == 2
    assert list(m.patch_map(0)) == [5, 0, 1, 4] and list(m.patch_map(1)) == [4, 2, 3, 6]


def test_embedded_geometry_source_is_in_sync():
    """csrc/geometry_src.inc (what NVRTC compiles) must be the text of csrc/geometry.cuh."""
    import os
    csrc = os.path.join(R.ROOT, "gismo_b200", "csrc")
    src = open(os.path.join(csrc, "geometry.cuh")).read()
    inc = open(os.path.join(csrc, "geometry_src.inc")).read()
    assert src in inc, "run __graft_entry__.build() to regenerate geometry_src.inc"


@pytest.mark.parametrize("text,dim,pgl,rational,fspec", [("3*pi^2*sin(pi*x)*sin(pi*y)*sin(pi*z)", 3, 2, 0, 1), ("2*pi^2*sin(pi*x)*sin(pi*y)", 2, 0, 1, 0),
                                                        ("x^2+exp(-y)/(1+z*z)-sqrt(abs(x))+cos(x)^3", 3, 3, 0, 0)])
def test_source_term_translates_and_compiles_with_nvrtc(text, dim, pgl, rational, fspec):
    """The NVRTC path of repeated assemblies (csrc/jit.cuh): program -> straight-line CUDA + geometry.cuh -> sm_100a cubin.
    Needs libnvrtc only (no device)."""
    import ctypes as C
    import gismo_b200 as g
    from gismo_b200 import capi
    lib = capi.load_library()
    pr = g.expr_compile(text)
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    P = (capi.Program * 1)(capi.Program(len(pr.ops), pr.ops.ctypes.data_as(ip), len(pr.consts), pr.consts.ctypes.data_as(dp)))
    log = C.create_string_buffer(8000)
    rc = lib.gsb200_jit_compile_check(P, 1, dim, pgl, rational, fspec, log, 8000)
    if rc and b"libnvrtc not available" in lib.gsb200_last_error():
        pytest.skip("libnvrtc not found")
    assert rc == 0, lib.gsb200_last_error().decode() + log.value.decode()


@pytest.mark.parametrize("name,dim,p,grid,m,form", [("grid2x2x2_p2_m3", 3, 2, (2, 2, 2), 3, 0), ("grid2x2_p2_m4", 2, 2, (2, 2), 4, 0),
                                                    ("elasticity_8cubes_p2_m5", 3, 2, (2, 2, 2), 5, 1)])
def test_multipatch_grid_builder_reproduces_the_reference_numbering(name, dim, p, grid, m, form):
    """host.multipatch_grid_problem (array operations on the glued lattice) against the flattened inputs the reference produced
    for the same grids: gsDofMapper numbering, knots after setDegree/uniformRefine, control points, patch order."""
    import os
    from gismo_b200 import host
    z = np.load(os.path.join(R.ROOT, "tests", "golden", name + ".npz"), allow_pickle=True)
    pb = host.multipatch_grid_problem(dim, p, grid, m, form=form)
    assert (pb.nfree, pb.nfixed) == (int(z["nfree"]), int(z["nfixed"]))
    for ip, pa in enumerate(pb.patches):
        assert np.array_equal(pa.dofmap, z[f"p{ip}_dofmap"])
        assert np.allclose(np.asarray(pa.geo_coefs), np.asarray(z[f"p{ip}_coefs"]))
        for k in range(dim):
            assert np.array_equal(pa.space_knots[k], z[f"p{ip}_sk{k}"])


def test_coupled_column_ranges_scalar_and_vector():
    from gismo_b200 import host, distributed as D
    pb = host.multipatch_grid_problem(2, 2, (2, 2), 4)
    assert D.coupled_column_ranges(pb) == [(D.first_coupled_column(pb), pb.nfree)]
    pe = host.multipatch_grid_problem(3, 2, (2, 2, 2), 3, form=1)
    runs = D.coupled_column_ranges(pe)
    n1 = pe.nfree // 3
    assert len(runs) == 3 and all(b == (c + 1) * n1 for c, (a, b) in enumerate(runs)) and len({b - a for a, b in runs}) == 1


def test_yeti_topology_refines_like_the_reference():
    """BASELINE config 3 (filedata/domain2d/yeti_mp2.xml, 21 patches): the topology is read off the coarse reference fixture
    and the discretisation rebuilt at another refinement by the host builder; at nelem = 8 it must reproduce, bit for bit,
    the DOF map the reference itself produced (tests/golden/yeti_mp2_p2_m8.npz)."""
    import goldenutil as G
    small, _ = G.load("yeti_mp2_p2_m2", lambda t: None, with_rhs=False)
    big, z = G.load("yeti_mp2_p2_m8", lambda t: None, with_rhs=False)
    interfaces, dirichlet = host.multipatch_topology_2d(small)
    assert len(small.patches) == 21 and len(interfaces) >= 20 and len(dirichlet) >= 1
    pb = host.refine_multipatch_2d(small, 2, 8)
    assert (pb.nfree, pb.nfixed) == (big.nfree, big.nfixed)
    for a, b in zip(pb.patches, big.patches):
        assert np.array_equal(a.dofmap, b.dofmap)
        assert all(np.allclose(x, y, rtol=0, atol=1e-15) for x, y in zip(a.space_knots, b.space_knots))
    same = host.refine_multipatch_2d(small, 2, 2)
    assert all(np.array_equal(a.dofmap, b.dofmap) for a, b in zip(same.patches, small.patches))
