"""The C restatement (oracle/gsb_oracle.c) against the reference's own results.

Pins the oracle: every golden fixture (made by the unmodified reference, tests/golden/
make_golden.py) must be reproduced with a bit-exact pattern and values/rhs within 1e-13 of
max|K| / max|rhs|; plus the reference's fingerprints recorded in SURVEY.md 8(c) and the
identities of SURVEY 8(c)-4.
"""
import numpy as np
import pytest

import goldenutil as G
import refutil as R

TOL = 1e-13


def _compile(text):
    return R.emul_compile(text)


@pytest.mark.parametrize("name", G.names("full"))
def test_oracle_matches_reference_fixture(name):
    pb, z = G.load(name, _compile)
    G.check_against(R.oracle_assemble(pb), z, TOL)


@pytest.mark.parametrize("name", ["sq_p2_m64", "cube_p4_m8", "cube_p2_m16_expr", "yeti_mp2_p2_m8", "elasticity_8cubes_p2_m5"])
def test_oracle_matches_reference_fingerprint(name):
    pb, z = G.load(name, _compile)
    G.check_against(R.oracle_assemble(pb), z, TOL)


def test_survey_fingerprints_are_the_fixture():
    # 3D p=3 16^3 -> nnz 1 225 043, sum K = 157.7158163265144, ||rhs|| = 0.12136673377211117
    z = dict(np.load(G.GOLDEN + "/cube_p3_m16.npz"))
    assert int(z["nnz"]) == 1225043
    assert abs(float(z["sumK"]) - 157.7158163265144) < 1e-10
    assert abs(float(z["rhs_norm"]) - 0.12136673377211117) < 1e-15
    z = dict(np.load(G.GOLDEN + "/sq_p2_m64.npz"))
    assert int(z["nnz"]) == 98596
    assert abs(float(z["sumK"]) - 336.355555555571) < 1e-10
    assert abs(float(z["rhs_norm"]) - 0.15411969832894112) < 1e-15


def test_gauss_nodes_closed_forms():
    # reference tables gsGaussRule.hpp:218-300 hold these well-known values
    x, w = R.oracle_gauss(2)
    assert np.allclose(x, [-1 / np.sqrt(3), 1 / np.sqrt(3)], rtol=0, atol=2e-16) and np.allclose(w, [1, 1], atol=2e-16)
    x, w = R.oracle_gauss(3)
    assert np.allclose(x, [-np.sqrt(0.6), 0, np.sqrt(0.6)], atol=2e-16) and np.allclose(w, [5 / 9, 8 / 9, 5 / 9], atol=2e-16)
    for n in range(1, 9):   # exact on monomials up to degree 2n-1 (unittests/gsQuadratureRules_test.cpp)
        x, w = R.oracle_gauss(n)
        for k in range(2 * n):
            exact = 0.0 if k % 2 else 2.0 / (k + 1)
            assert abs(np.dot(w, x ** k) - exact) < 5e-15


def test_basis_partition_of_unity_and_derivative_sum():
    import ctypes as C
    lib = R.oracle_lib()
    rng = np.random.default_rng(12345)
    for p in (1, 2, 3, 4):
        knots = np.concatenate([[0.0] * (p + 1), np.sort(rng.uniform(0.05, 0.95, 7)), [1.0] * (p + 1)])
        if p > 1:
            knots[p + 3] = knots[p + 2]   # a double knot
        for u in rng.uniform(0, 1, 25):
            val = np.zeros(p + 1); der = np.zeros(p + 1); first = C.c_int(0)
            lib.gsbo_basis_eval(knots.ctypes.data_as(C.POINTER(C.c_double)), len(knots), p, float(u), C.byref(first),
                                val.ctypes.data_as(C.POINTER(C.c_double)), der.ctypes.data_as(C.POINTER(C.c_double)))
            assert abs(val.sum() - 1.0) < 1e-14 and abs(der.sum()) < 1e-11
            from scipy.interpolate import BSpline
            for a in range(p + 1):
                c = np.zeros(len(knots) - p - 1); c[first.value + a] = 1.0
                b = BSpline(knots, c, p)
                assert abs(b(u) - val[a]) < 1e-13 and abs(b(u, nu=1) - der[a]) < 1e-10


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref/libgsref.so not built here")
def test_live_reference_vs_oracle_and_tables():
    import ctypes as C
    ref = R.ref_run(dim=3, degree=2, nelem=3, geometry=1, rhs=["x*y+z"], dirichlet=["sin(x)"])
    pb = ref.problem(compile_fn=_compile)
    ok, msg = R.compare_csc(R.oracle_assemble(pb), (ref.outer, ref.inner, ref.values, ref.rhs), TOL)
    assert ok, msg
    lib = R.ref_lib()
    for n in range(1, 8):   # oracle's Newton nodes == the reference's 30-digit tables, to the last bit or ulp
        x = np.zeros(n); w = np.zeros(n)
        lib.gsref_gauss(n, x.ctypes.data_as(C.POINTER(C.c_double)), w.ctypes.data_as(C.POINTER(C.c_double)))
        xo, wo = R.oracle_gauss(n)
        assert np.abs(x - xo).max() <= 2.3e-16 and np.abs(w - wo).max() <= 2.3e-16


@pytest.mark.parametrize("name", sorted(G.NORM_CASES))
def test_oracle_norm_integrals_match_the_reference(name):
    """gsbo_field_norms against ev.integral((u_ex - u_sol).sqNorm() * meas(G)) etc. of the reference (gsExprEvaluator) on the
    reference's own solution coefficients."""
    pb, z = G.load(name, R.emul_compile)
    ex, grads = G.NORM_CASES[name]
    got = R.oracle_field_norms(pb, z["solution"], R.emul_compile(ex), [R.emul_compile(t) for t in grads])
    G.check_norms(got, z)
    only_field = R.oracle_field_norms(pb, z["solution"])
    assert only_field[0] == only_field[2] and only_field[1] == only_field[3]
