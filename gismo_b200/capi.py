"""ctypes view of the C ABI declared in include/gsb200.h.

The structures below mirror the header field by field; `Problem` owns the numpy buffers a
`gsb200_problem` points into.  The product library (gismo_b200/csrc/libgsb200.so) is the
only thing this module loads — there is no CPU fallback: if the CUDA extension is missing
or no device is present, calls raise `Gsb200Error`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

MAX_DIM = 3
ABI_VERSION = 2
FORM_POISSON, FORM_ELASTICITY, FORM_MASS = 0, 1, 2
RHS_NONE, RHS_PROGRAM, RHS_SAMPLES = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class Basis(C.Structure):
    _fields_ = [("dim", C.c_int32), ("degree", C.c_int32 * MAX_DIM), ("nknots", C.c_int32 * MAX_DIM),
                ("knots", _dp * MAX_DIM)]


class Patch(C.Structure):
    _fields_ = [("space", Basis), ("geo", Basis), ("geo_coefs", _dp), ("geo_weights", _dp), ("dofmap", _ip)]


class Program(C.Structure):
    _fields_ = [("nops", C.c_int32), ("ops", _ip), ("nconsts", C.c_int32), ("consts", _dp)]


class Neumann(C.Structure):
    _fields_ = [("patch", C.c_int32), ("side", C.c_int32), ("ndata", C.c_int32), ("data", Program * MAX_DIM)]


class ProblemStruct(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("form", C.c_int32), ("npatches", C.c_int32),
                ("patches", C.POINTER(Patch)), ("ncomp", C.c_int32), ("nfree", C.c_int32),
                ("nfixed", C.c_int32), ("fixed", _dp), ("nrhs", C.c_int32), ("coef", C.c_double * 4),
                ("quA", C.c_double), ("quB", C.c_int32), ("rhs_kind", C.c_int32),
                ("rhs_programs", C.POINTER(Program)), ("rhs_samples", C.POINTER(_dp)),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("nneumann", C.c_int32), ("neumann", C.POINTER(Neumann))]


class DeviceView(C.Structure):
    _fields_ = [("nnz", C.c_int64), ("ncols", C.c_int32), ("col_begin", C.c_int32), ("col_end", C.c_int32),
                ("outer", C.c_void_p), ("inner", C.c_void_p), ("values", C.c_void_p), ("rhs", C.c_void_p)]


class Timings(C.Structure):
    _fields_ = [("geometry_ms", C.c_float), ("sweep_ms", C.c_float * MAX_DIM), ("rhs_ms", C.c_float),
                ("pattern_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_int32),
                ("sweep_bytes", C.c_int64 * MAX_DIM), ("sweep_flops", C.c_int64 * MAX_DIM),
                ("nchunks", C.c_int32)]


COMM_ID_BYTES = 128
# int fn(void *ctx, double *buf, int64_t count, void *stream): in-place sum of doubles over the ranks
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.c_void_p)


class Gsb200Error(RuntimeError):
    pass


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class PatchData:
    """Plain description of one patch (what gsB200Flatten.h extracts from gismo objects)."""

    def __init__(self, space_degree: Sequence[int], space_knots: Sequence[np.ndarray],
                 geo_degree: Sequence[int], geo_knots: Sequence[np.ndarray], geo_coefs: np.ndarray,
                 dofmap: np.ndarray, geo_weights: Optional[np.ndarray] = None):
        self.dim = len(space_degree)
        self.space_degree = [int(p) for p in space_degree]
        self.space_knots = [_as_f64(k) for k in space_knots]
        self.geo_degree = [int(p) for p in geo_degree]
        self.geo_knots = [_as_f64(k) for k in geo_knots]
        # N_geo x dim, column-major (Eigen) == Fortran order
        self.geo_coefs = np.asfortranarray(np.asarray(geo_coefs, dtype=np.float64))
        self.geo_weights = None if geo_weights is None else _as_f64(geo_weights).ravel()
        self.dofmap = np.ascontiguousarray(np.asarray(dofmap, dtype=np.int32).ravel())

    @property
    def nfun(self) -> List[int]:
        return [len(k) - p - 1 for k, p in zip(self.space_knots, self.space_degree)]

    @property
    def nbasis(self) -> int:
        return int(np.prod(self.nfun))


class Problem:
    """Owns every buffer of a gsb200_problem; `.struct` is what the C ABI receives."""

    def __init__(self, patches: List[PatchData], nfree: int, nfixed: int, form: int = FORM_POISSON,
                 ncomp: int = 1, fixed: Optional[np.ndarray] = None, nrhs: int = 1,
                 coef: Sequence[float] = (0.0, 0.0), quA: float = 1.0, quB: int = 1,
                 rhs_programs: Optional[List["CompiledProgram"]] = None, rank: int = 0, nranks: int = 1,
                 neumann: Optional[List[tuple]] = None):
        self.patches = patches
        self.nfree, self.nfixed, self.form, self.ncomp, self.nrhs = int(nfree), int(nfixed), form, ncomp, nrhs
        self._ctor = dict(form=form, ncomp=ncomp, nrhs=nrhs, coef=tuple(coef), quA=quA, quB=quB, rhs_programs=rhs_programs,
                          rank=rank, nranks=nranks, neumann=neumann)
        self.fixed = None if fixed is None else np.asfortranarray(np.asarray(fixed, dtype=np.float64).reshape(nfixed, -1))
        self.rhs_programs = rhs_programs or []
        self._keep = []
        pa = (Patch * len(patches))()
        for k, p in enumerate(patches):
            for b, deg, kn in ((pa[k].space, p.space_degree, p.space_knots), (pa[k].geo, p.geo_degree, p.geo_knots)):
                b.dim = p.dim
                for i in range(p.dim):
                    b.degree[i] = deg[i]
                    b.nknots[i] = len(kn[i])
                    b.knots[i] = kn[i].ctypes.data_as(_dp)
            pa[k].geo_coefs = p.geo_coefs.ctypes.data_as(_dp)
            pa[k].geo_weights = p.geo_weights.ctypes.data_as(_dp) if p.geo_weights is not None else None
            pa[k].dofmap = p.dofmap.ctypes.data_as(_ip)
        self._pa = pa
        s = ProblemStruct()
        s.abi_version = ABI_VERSION
        s.form = form
        s.npatches = len(patches)
        s.patches = pa
        s.ncomp, s.nfree, s.nfixed, s.nrhs = ncomp, self.nfree, self.nfixed, nrhs
        s.fixed = self.fixed.ctypes.data_as(_dp) if self.fixed is not None else None
        for i, c in enumerate(coef):
            s.coef[i] = float(c)
        s.quA, s.quB = float(quA), int(quB)
        if self.rhs_programs:
            pr = (Program * len(self.rhs_programs))()
            for i, cp in enumerate(self.rhs_programs):
                pr[i].nops = len(cp.ops)
                pr[i].ops = cp.ops.ctypes.data_as(_ip)
                pr[i].nconsts = len(cp.consts)
                pr[i].consts = cp.consts.ctypes.data_as(_dp)
            self._pr = pr
            s.rhs_kind = RHS_PROGRAM
            s.rhs_programs = pr
        else:
            s.rhs_kind = RHS_NONE
        s.rank, s.nranks = rank, nranks
        # neumann: list of (patch, side, [CompiledProgram, ...])
        self.neumann = neumann or []
        if self.neumann:
            nm = (Neumann * len(self.neumann))()
            for i, (patch, side, progs) in enumerate(self.neumann):
                nm[i].patch, nm[i].side, nm[i].ndata = int(patch), int(side), len(progs)
                for k, cp in enumerate(progs):
                    nm[i].data[k].nops = len(cp.ops)
                    nm[i].data[k].ops = cp.ops.ctypes.data_as(_ip)
                    nm[i].data[k].nconsts = len(cp.consts)
                    nm[i].data[k].consts = cp.consts.ctypes.data_as(_dp)
            self._nm = nm
            s.nneumann = len(self.neumann)
            s.neumann = nm
        self.struct = s

    @property
    def dim(self) -> int:
        return self.patches[0].dim

    def with_fixed(self, fixed: Optional[np.ndarray], **changes) -> "Problem":
        """The same problem with other eliminated-DOF values (and any other constructor argument in `changes`)."""
        kw = dict(self._ctor)
        kw.update(changes)
        return Problem(self.patches, self.nfree, self.nfixed, fixed=fixed, **kw)


class CompiledProgram:
    def __init__(self, ops: np.ndarray, consts: np.ndarray, text: str = ""):
        self.ops = np.ascontiguousarray(ops, dtype=np.int32)
        self.consts = np.ascontiguousarray(consts, dtype=np.float64)
        self.text = text


_LIB = None


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libgsb200.so")


def load_library():
    """Load the CUDA extension; raise loudly if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise Gsb200Error(f"CUDA extension {path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    _LIB = declare(C.CDLL(path))
    return _LIB


def declare(lib, optional=()):
    """ctypes signatures of every entry point of include/gsb200.h (also used for the interpreter build of the tests);
    names in `optional` may be absent from `lib`."""
    lib.gsb200_last_error.restype = C.c_char_p
    lib.gsb200_create.argtypes = [C.POINTER(ProblemStruct), C.c_int, C.POINTER(C.c_void_p)]
    lib.gsb200_destroy.argtypes = [C.c_void_p]
    lib.gsb200_destroy.restype = None
    lib.gsb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.gsb200_set_workspace_limit.argtypes = [C.c_void_p, C.c_int64]
    lib.gsb200_build_pattern.argtypes = [C.c_void_p]
    lib.gsb200_assemble.argtypes = [C.c_void_p]
    lib.gsb200_synchronize.argtypes = [C.c_void_p]
    lib.gsb200_nnz.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.gsb200_device_view_get.argtypes = [C.c_void_p, C.POINTER(DeviceView)]
    lib.gsb200_timings_get.argtypes = [C.c_void_p, C.POINTER(Timings)]
    lib.gsb200_download_csc.argtypes = [C.c_void_p, _ip, _ip, _dp]
    lib.gsb200_download_rhs.argtypes = [C.c_void_p, _dp]
    lib.gsb200_assemble_to_host.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp]
    lib.gsb200_assemble_host.argtypes = [C.POINTER(ProblemStruct), C.c_int, C.POINTER(C.c_int64), _ip, _ip, _dp, _dp]
    lib.gsb200_download_pattern.argtypes = [C.c_void_p, _ip, _ip]
    lib.gsb200_set_fixed.argtypes = [C.c_void_p, _dp]
    lib.gsb200_assemble_values_to_host.argtypes = [C.c_void_p, _dp, _dp]
    lib.gsb200_host_pin.argtypes = [C.c_void_p, C.c_int64]
    lib.gsb200_host_unpin.argtypes = [C.c_void_p]
    lib.gsb200_spmv_host.argtypes = [C.c_void_p, _dp, _dp]
    lib.gsb200_spmv_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gsb200_diag_device.argtypes = [C.c_void_p, C.c_void_p]
    lib.gsb200_diag_host.argtypes = [C.c_void_p, _dp]
    lib.gsb200_cg_host.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double, C.POINTER(C.c_int), _dp]
    lib.gsb200_expr_compile.argtypes = [C.c_char_p, _ip, C.c_int32, _ip, _dp, C.c_int32, _ip]
    lib.gsb200_jit_launches.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.gsb200_jit_compile_check.argtypes = [C.POINTER(Program), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.gsb200_expr_eval_host.argtypes = [C.POINTER(Program), C.c_double, C.c_double, C.c_double, _dp]
    lib.gsb200_measure_peaks.argtypes = [C.c_int, _dp, _dp, _dp]
    lib.gsb200_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.gsb200_trim.argtypes = [C.c_int]
    lib.gsb200_comm_unique_id.argtypes = [C.c_void_p]
    lib.gsb200_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    lib.gsb200_set_comm.argtypes = [C.c_void_p, C.c_void_p]
    lib.gsb200_set_allreduce.argtypes = [C.c_void_p, ALLREDUCE_FN, C.c_void_p]
    lib.gsb200_exchange.argtypes = [C.c_void_p]
    lib.gsb200_comm_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    lib.gsb200_cg_solve.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), _dp]
    lib.gsb200_field_norms.argtypes = [C.c_void_p, _dp, C.POINTER(Program), C.POINTER(Program), _dp]
    lib.gsb200_project_dirichlet.argtypes = [C.c_void_p, C.POINTER(Neumann), C.c_int, C.c_int, C.c_double, _dp, C.POINTER(C.c_int), _dp]
    lib.gsb200_cg_info.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int32)]
    lib.gsb200_cg_solution_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.gsb200_spmv_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    return lib


def check(status: int):
    if status != 0:
        raise Gsb200Error(f"gsb200 error {status}: {load_library().gsb200_last_error().decode()}")


def expr_compile(text: str) -> CompiledProgram:
    """gsb200_expr_compile: exprtk-style string -> reverse-polish program."""
    lib = load_library()
    ops = np.zeros(256, dtype=np.int32)
    consts = np.zeros(256, dtype=np.float64)
    nops, ncst = C.c_int32(0), C.c_int32(0)
    check(lib.gsb200_expr_compile(text.encode(), ops.ctypes.data_as(_ip), 256, C.byref(nops),
                                  consts.ctypes.data_as(_dp), 256, C.byref(ncst)))
    return CompiledProgram(ops[:nops.value].copy(), consts[:ncst.value].copy(), text)


def expr_eval(prog: CompiledProgram, x: float, y: float = 0.0, z: float = 0.0) -> float:
    lib = load_library()
    p = Program(len(prog.ops), prog.ops.ctypes.data_as(_ip), len(prog.consts), prog.consts.ctypes.data_as(_dp))
    out = C.c_double(0)
    check(lib.gsb200_expr_eval_host(C.byref(p), x, y, z, C.byref(out)))
    return out.value
