"""Test-only helpers: the checkers (oracle/_ref/*.so) and the kernel interpreter (tests/emul).

Nothing here is imported by the package.  `ref_*` functions need oracle/_ref/libgsref.so,
which exists only where it was built from /root/reference (the build container) or travelled
with the snapshot; tests that need it skip when it is absent.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import List, Optional, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gismo_b200 import capi  # noqa: E402
from gismo_b200.capi import PatchData, Problem, ProblemStruct  # noqa: E402

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "libgsb_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgsref.so")
EMUL_SO = os.path.join(ROOT, "tests", "emul", "_build", "libgsb200_emul.so")


def build_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/libgsb_oracle.so"])


def build_emul():
    subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "tests", "emul")])


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.gsbo_assemble.argtypes = [C.POINTER(ProblemStruct), C.POINTER(C.c_int64), _ip, _ip, _dp, _dp]
        lib.gsbo_gauss.argtypes = [C.c_int, _dp, _dp]
        lib.gsbo_basis_eval.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), _dp, _dp]
        lib.gsbo_last_error.restype = C.c_char_p
        lib.gsbo_field_norms.argtypes = [C.POINTER(ProblemStruct), _dp, C.POINTER(capi.Program), C.POINTER(capi.Program), _dp]
        _oracle = lib
    return _oracle


def oracle_assemble(pb: Problem):
    """CPU restatement: returns (outer, inner, values, rhs)."""
    lib = oracle_lib()
    nnz = C.c_int64(0)
    rc = lib.gsbo_assemble(C.byref(pb.struct), C.byref(nnz), None, None, None, None)
    if rc:
        raise RuntimeError(lib.gsbo_last_error().decode())
    outer = np.zeros(pb.nfree + 1, np.int32)
    inner = np.zeros(nnz.value, np.int32)
    values = np.zeros(nnz.value, np.float64)
    rhs = np.zeros((pb.nfree, pb.nrhs), np.float64, order="F")
    rc = lib.gsbo_assemble(C.byref(pb.struct), C.byref(nnz), outer.ctypes.data_as(_ip), inner.ctypes.data_as(_ip),
                           values.ctypes.data_as(_dp), rhs.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError(lib.gsbo_last_error().decode())
    return outer, inner, values, rhs


def oracle_field_norms(pb, u, exact=None, exact_grad=None) -> np.ndarray:
    """oracle/gsb_oracle.c::gsbo_field_norms (checker only)."""
    lib = oracle_lib()
    def prog(cp):
        return capi.Program(len(cp.ops), cp.ops.ctypes.data_as(_ip), len(cp.consts), cp.consts.ctypes.data_as(_dp))
    u = np.ascontiguousarray(u, dtype=np.float64).ravel()
    ex = C.byref(prog(exact)) if exact is not None else None
    eg = (capi.Program * len(exact_grad))(*[prog(cp) for cp in exact_grad]) if exact_grad is not None else None
    out = np.zeros(4)
    if lib.gsbo_field_norms(C.byref(pb.struct), u.ctypes.data_as(_dp), ex, eg, out.ctypes.data_as(_dp)):
        raise RuntimeError(lib.gsbo_last_error().decode())
    return out


def oracle_gauss(n: int):
    x = np.zeros(n); w = np.zeros(n)
    oracle_lib().gsbo_gauss(n, x.ctypes.data_as(_dp), w.ctypes.data_as(_dp))
    return x, w


# --------------------------------------------------------------------------- reference
class RefConfig(C.Structure):
    _fields_ = [("dim", C.c_int32), ("degree", C.c_int32), ("nelem", C.c_int32), ("geometry", C.c_int32),
                ("grid", C.c_int32 * 3), ("path", C.c_int32), ("form", C.c_int32), ("dir_values", C.c_int32),
                ("threads", C.c_int32), ("degree_elevate", C.c_int32), ("rhs", C.c_char_p * 3),
                ("dir", C.c_char_p * 3), ("xml", C.c_char_p), ("lambda_", C.c_double), ("mu", C.c_double),
                ("neumann_mask", C.c_int32), ("neu_n", C.c_int32), ("neu", C.c_char_p * 3),
                ("degree_dir", C.c_int32 * 3), ("nrhs", C.c_int32), ("exact", C.c_char_p)]


_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.gsref_run.restype = C.c_void_p
        lib.gsref_run.argtypes = [C.POINTER(RefConfig)]
        lib.gsref_free.argtypes = [C.c_void_p]
        lib.gsref_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), _dp, _dp, _ip]
        lib.gsref_solution.argtypes = [C.c_void_p, _dp, _dp]
        lib.gsref_csc.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _dp]
        lib.gsref_patch_info.argtypes = [C.c_void_p, C.c_int, _ip]
        lib.gsref_patch_data.argtypes = [C.c_void_p, C.c_int] + [_dp] * 8 + [_ip]
        lib.gsref_last_error.restype = C.c_char_p
        lib.gsref_gauss.argtypes = [C.c_int, _dp, _dp]
        lib.gsref_uniform_refine.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, C.POINTER(C.c_int)]
        lib.gsref_neumann.argtypes = [C.c_void_p, _ip, C.c_int32]
        lib.gsref_text.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        _ref = lib
    return _ref


class RefResult:
    """Output of the reference's own assembler plus the flattened inputs it ran on."""

    def __init__(self):
        self.outer = self.inner = self.values = self.rhs = self.fixed = None
        self.patches: List[PatchData] = []
        self.nfree = self.nfixed = self.ncomp = self.dim = 0
        self.seconds = 0.0
        self.elements = self.qpoints = 0
        self.quA, self.quB = 1.0, 1
        self.form = 0
        self.coef = (0.0, 0.0)
        self.rhs_text: List[str] = []
        self.neumann_sides: List[Tuple[int, int]] = []
        self.neu_text: List[str] = []

    def problem(self, with_rhs: bool = True, compile_fn=None) -> Problem:
        progs = None
        if with_rhs and self.rhs_text:
            compile_fn = compile_fn or capi.expr_compile
            progs = [compile_fn(t) for t in self.rhs_text]
        compile_fn = compile_fn or capi.expr_compile
        neumann = [(p, s, [compile_fn(t) for t in self.neu_text]) for p, s in self.neumann_sides]
        return Problem(self.patches, self.nfree, self.nfixed, form=self.form, ncomp=self.ncomp,
                       fixed=self.fixed if self.nfixed else None, nrhs=1, coef=self.coef, quA=self.quA,
                       quB=self.quB, rhs_programs=progs, neumann=neumann)


def ref_run(dim=3, degree=2, nelem=4, geometry=0, grid=(1, 1, 1), path=0, form=0, dir_values=101,
            threads=1, rhs=None, dirichlet=None, xml=None, lam=0.0, mu=0.0, degree_elevate=0,
            neumann_mask=0, neu=None, degree_dir=None, nrhs=1, exact=None) -> RefResult:
    lib = ref_lib()
    cfg = RefConfig()
    cfg.dim, cfg.degree, cfg.nelem, cfg.geometry = dim, degree, nelem, geometry
    for k in range(3):
        cfg.grid[k] = grid[k] if k < len(grid) else 1
    cfg.path, cfg.form, cfg.dir_values, cfg.threads, cfg.degree_elevate = path, form, dir_values, threads, degree_elevate
    ncomp = dim if form == 1 else 1
    if form == 0 and path == 0 and nrhs > 1:
        ncomp = nrhs                      # components of the source / Dirichlet functions = right-hand-side columns
    cfg.nrhs = nrhs
    cfg.exact = exact.encode() if exact else None
    for k, v in enumerate(degree_dir or []):
        cfg.degree_dir[k] = v
    rhs = list(rhs or ["0"] * ncomp)
    dirichlet = list(dirichlet or ["0"] * ncomp)
    for k in range(ncomp):
        cfg.rhs[k] = rhs[k].encode()
        cfg.dir[k] = dirichlet[k].encode()
    cfg.xml = xml.encode() if xml else None
    cfg.lambda_, cfg.mu = lam, mu
    neu = list(neu or [])
    cfg.neumann_mask, cfg.neu_n = neumann_mask, len(neu)
    for k, t in enumerate(neu):
        cfg.neu[k] = t.encode()
    h = lib.gsref_run(C.byref(cfg))
    if not h:
        raise RuntimeError("reference failed: " + lib.gsref_last_error().decode())
    try:
        sizes = (C.c_int64 * 8)()
        sec, quA, quB = C.c_double(0), C.c_double(0), C.c_int32(0)
        lib.gsref_sizes(h, sizes, C.byref(sec), C.byref(quA), C.byref(quB))
        R = RefResult()
        R.nfree, R.nfixed, nnz, npatches, R.ncomp, R.dim, R.elements, R.qpoints = [int(v) for v in sizes]
        R.seconds, R.quA, R.quB = sec.value, quA.value, quB.value
        R.form, R.coef, R.rhs_text = form, (lam, mu), rhs
        pairs = (C.c_int32 * 128)()
        nn = lib.gsref_neumann(h, pairs, 128)
        R.neumann_sides = [(pairs[2 * i], pairs[2 * i + 1]) for i in range(nn)]
        def texts(which):
            out = []
            buf = C.create_string_buffer(1024)
            k = 0
            while lib.gsref_text(h, which, k, buf, 1024) > 0:
                out.append(buf.value.decode()); k += 1
            return out
        R.rhs_text, R.neu_text = texts(0), texts(1)
        R.outer = np.zeros(R.nfree + 1, np.int32)
        R.inner = np.zeros(nnz, np.int32)
        R.values = np.zeros(nnz)
        R.nrhs = nrhs if (form == 0 and path == 0) else 1
        R.rhs = np.zeros((R.nfree, R.nrhs), order="F")
        R.fixed = np.zeros((max(R.nfixed, 1), R.nrhs), order="F")
        lib.gsref_csc(h, R.outer.ctypes.data_as(_ip), R.inner.ctypes.data_as(_ip), R.values.ctypes.data_as(_dp),
                      R.rhs.ctypes.data_as(_dp), R.fixed.ctypes.data_as(_dp))
        R.fixed = R.fixed[:R.nfixed]
        R.solution, R.norms = None, None
        if exact:
            sol, nrm = np.zeros(R.nfree), np.zeros(4)
            if lib.gsref_solution(h, sol.ctypes.data_as(_dp), nrm.ctypes.data_as(_dp)) > 0:
                R.solution, R.norms = sol, nrm
        for k in range(npatches):
            info = (C.c_int32 * 15)()
            lib.gsref_patch_info(h, k, info)
            d = R.dim
            sk = [np.zeros(max(info[3 + i], 1)) for i in range(3)]
            gk = [np.zeros(max(info[9 + i], 1)) for i in range(3)]
            coefs = np.zeros((info[13], d), order="F")
            weights = np.zeros(info[13])
            dofmap = np.zeros(info[12] * R.ncomp, np.int32)
            lib.gsref_patch_data(h, k, *[a.ctypes.data_as(_dp) for a in sk], *[a.ctypes.data_as(_dp) for a in gk],
                                 coefs.ctypes.data_as(_dp), weights.ctypes.data_as(_dp), dofmap.ctypes.data_as(_ip))
            R.patches.append(PatchData([info[i] for i in range(d)], sk[:d], [info[6 + i] for i in range(d)], gk[:d],
                                       coefs, dofmap, weights if info[14] else None))
        return R
    finally:
        lib.gsref_free(h)


# --------------------------------------------------------------------------- interpreter
_emul = None


def emul_lib():
    global _emul
    if _emul is None:
        build_emul()
        lib = C.CDLL(EMUL_SO)
        capi.declare(lib)
        _emul = lib
    return _emul


def emul_compile(text: str) -> capi.CompiledProgram:
    lib = emul_lib()
    ops = np.zeros(256, np.int32); consts = np.zeros(256)
    nops, ncst = C.c_int32(0), C.c_int32(0)
    rc = lib.gsb200_expr_compile(text.encode(), ops.ctypes.data_as(_ip), 256, C.byref(nops),
                                 consts.ctypes.data_as(_dp), 256, C.byref(ncst))
    if rc:
        raise RuntimeError(lib.gsb200_last_error().decode())
    return capi.CompiledProgram(ops[:nops.value].copy(), consts[:ncst.value].copy(), text)


def lib_assemble(lib, pb: Problem, device: int = 0, workspace_limit: int = 0):
    """Drive any library exporting the C ABI (the product .so or the interpreter)."""
    h = C.c_void_p()
    def chk(rc):
        if rc:
            raise RuntimeError(f"gsb200 error {rc}: {lib.gsb200_last_error().decode()}")
    chk(lib.gsb200_create(C.byref(pb.struct), device, C.byref(h)))
    try:
        if workspace_limit:
            chk(lib.gsb200_set_workspace_limit(h, workspace_limit))
        chk(lib.gsb200_build_pattern(h))
        chk(lib.gsb200_assemble(h))
        nnz = C.c_int64(0)
        chk(lib.gsb200_nnz(h, C.byref(nnz)))
        outer = np.zeros(pb.nfree + 1, np.int32)
        inner = np.zeros(nnz.value, np.int32)
        values = np.zeros(nnz.value)
        rhs = np.zeros((pb.nfree, pb.nrhs), order="F")
        chk(lib.gsb200_download_csc(h, outer.ctypes.data_as(_ip), inner.ctypes.data_as(_ip), values.ctypes.data_as(_dp)))
        chk(lib.gsb200_download_rhs(h, rhs.ctypes.data_as(_dp)))
        tm = capi.Timings()
        lib.gsb200_timings_get(h, C.byref(tm))
        return outer, inner, values, rhs, tm
    finally:
        lib.gsb200_destroy(h)


def lib_reassemble(lib, pb: Problem, fixed2: np.ndarray, device: int = 0):
    """First delivery (gsb200_assemble_to_host), then new eliminated-DOF values (gsb200_set_fixed) and a values-only
    re-assembly on the kept pattern (gsb200_assemble_values_to_host).  Returns both results."""
    h = C.c_void_p()
    def chk(rc):
        if rc:
            raise RuntimeError(f"gsb200 error {rc}: {lib.gsb200_last_error().decode()}")
    chk(lib.gsb200_create(C.byref(pb.struct), device, C.byref(h)))
    try:
        chk(lib.gsb200_build_pattern(h))
        nnz = C.c_int64(0)
        chk(lib.gsb200_nnz(h, C.byref(nnz)))
        outer = np.zeros(pb.nfree + 1, np.int32)
        inner = np.zeros(nnz.value, np.int32)
        values = np.zeros(nnz.value)
        rhs = np.zeros((pb.nfree, pb.nrhs), order="F")
        chk(lib.gsb200_assemble_to_host(h, outer.ctypes.data_as(_ip), inner.ctypes.data_as(_ip), values.ctypes.data_as(_dp), rhs.ctypes.data_as(_dp)))
        tm = capi.Timings()
        lib.gsb200_timings_get(h, C.byref(tm))
        first = (outer, inner, values.copy(), rhs.copy(), tm)
        f2 = np.asfortranarray(fixed2, dtype=np.float64)
        chk(lib.gsb200_set_fixed(h, f2.ctypes.data_as(_dp)))
        values[:] = np.nan; rhs[:] = np.nan
        chk(lib.gsb200_assemble_values_to_host(h, values.ctypes.data_as(_dp), rhs.ctypes.data_as(_dp)))
        return first, (outer, inner, values, rhs, tm)
    finally:
        lib.gsb200_destroy(h)


def compare_csc(a, b, tol=1e-12) -> Tuple[bool, str]:
    """pattern bit-exact; values / rhs relative to max|K| / max|rhs|."""
    (o1, i1, v1, r1), (o2, i2, v2, r2) = a[:4], b[:4]
    if not np.array_equal(o1, o2):
        return False, "outer index arrays differ"
    if not np.array_equal(i1, i2):
        return False, "inner index arrays differ"
    ev = np.max(np.abs(v1 - v2)) / max(np.max(np.abs(v2)), 1e-300) if len(v1) else 0.0
    er = np.max(np.abs(r1 - r2)) / max(np.max(np.abs(r2)), 1e-300) if r2.size and np.max(np.abs(r2)) > 0 else float(np.max(np.abs(r1 - r2))) if r2.size else 0.0
    ok = ev <= tol and er <= tol
    return ok, f"max|dK|/max|K| = {ev:.3e}, max|drhs|/max|rhs| = {er:.3e}"
