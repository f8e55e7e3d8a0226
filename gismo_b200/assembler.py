"""Python face of the device assembler: same verbs as the reference's assembler objects
(gsAssembler.h:415,614,618 — assemble(), matrix(), rhs(), numDofs()), driving the C ABI.

Every call goes through gismo_b200/csrc/libgsb200.so; if that extension or a CUDA device is
missing the calls raise (there is no CPU path in the package).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import capi
from .capi import Problem, Timings, DeviceView, check

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class DeviceAssembler:
    """Handle on a device-resident problem (gsb200_create ... gsb200_destroy)."""

    def __init__(self, problem: Problem, device: int = 0, stream: Optional[int] = None,
                 workspace_limit: int = 0, lib=None):
        """lib: another build of the same C ABI (the tests drive the kernel interpreter through this class); default:
        the CUDA extension, which must exist (no fallback)."""
        self.lib = lib if lib is not None else capi.load_library()
        self.problem = problem
        self._h = C.c_void_p()
        self._check(self.lib.gsb200_create(C.byref(problem.struct), device, C.byref(self._h)))
        if stream is not None:
            self._check(self.lib.gsb200_set_stream(self._h, C.c_void_p(stream)))
        if workspace_limit:
            self._check(self.lib.gsb200_set_workspace_limit(self._h, workspace_limit))
        self._pattern = False

    def _check(self, status: int) -> None:
        if status != 0:
            raise capi.Gsb200Error(f"gsb200 error {status}: {self.lib.gsb200_last_error().decode()}")

    def close(self):
        if self._h:
            self.lib.gsb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- reference-facing verbs -------------------------------------------------------
    def numDofs(self) -> int:
        return self.problem.nfree

    def buildPattern(self) -> int:
        self._check(self.lib.gsb200_build_pattern(self._h))
        self._pattern = True
        return self.nnz()

    def assemble(self, sync: bool = True) -> None:
        if not self._pattern:
            self.buildPattern()
        self._check(self.lib.gsb200_assemble(self._h))
        if sync:
            self.synchronize()

    def synchronize(self) -> None:
        self._check(self.lib.gsb200_synchronize(self._h))

    def nnz(self) -> int:
        n = C.c_int64(0)
        self._check(self.lib.gsb200_nnz(self._h, C.byref(n)))
        return n.value

    def matrix(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(outer, inner, values) exactly as Eigen's compressed gsSparseMatrix stores them."""
        nnz = self.nnz()
        outer = np.zeros(self.problem.nfree + 1, np.int32)
        inner = np.zeros(nnz, np.int32)
        values = np.zeros(nnz, np.float64)
        self._check(self.lib.gsb200_download_csc(self._h, outer.ctypes.data_as(_ip), inner.ctypes.data_as(_ip),
                                           values.ctypes.data_as(_dp)))
        return outer, inner, values

    def matrix_into(self, outer: np.ndarray, inner: np.ndarray, values: np.ndarray) -> None:
        self._check(self.lib.gsb200_download_csc(self._h, outer.ctypes.data_as(_ip), inner.ctypes.data_as(_ip),
                                           values.ctypes.data_as(_dp)))

    def assemble_into(self, outer: np.ndarray, inner: np.ndarray, values: np.ndarray, rhs: Optional[np.ndarray] = None) -> None:
        """gsb200_assemble_to_host: assemble and deliver into caller buffers (pattern arrays travel while
        the values are integrated).  What gsPoissonAssemblerB200::assemble() does after sizing the matrix."""
        if not self._pattern:
            self.buildPattern()
        self._check(self.lib.gsb200_assemble_to_host(self._h, outer.ctypes.data_as(_ip), inner.ctypes.data_as(_ip), values.ctypes.data_as(_dp),
                                               rhs.ctypes.data_as(_dp) if rhs is not None else None))

    def pattern_into(self, outer: np.ndarray, inner: np.ndarray) -> None:
        """gsb200_download_pattern: the index arrays, once per mesh."""
        if not self._pattern:
            self.buildPattern()
        self._check(self.lib.gsb200_download_pattern(self._h, outer.ctypes.data_as(_ip), inner.ctypes.data_as(_ip)))

    def set_fixed(self, fixed: Optional[np.ndarray]) -> None:
        """gsb200_set_fixed: new eliminated-DOF values (nfixed x nrhs, column-major) on the kept pattern."""
        if fixed is None:
            self._check(self.lib.gsb200_set_fixed(self._h, None))
        else:
            f = np.asfortranarray(fixed, dtype=np.float64)
            self._check(self.lib.gsb200_set_fixed(self._h, f.ctypes.data_as(_dp)))

    def assemble_values_into(self, values: np.ndarray, rhs: Optional[np.ndarray] = None) -> None:
        """gsb200_assemble_values_to_host: re-assembly on the kept pattern, values (+ rhs) only travel."""
        if not self._pattern:
            self.buildPattern()
        self._check(self.lib.gsb200_assemble_values_to_host(self._h, values.ctypes.data_as(_dp),
                                                      rhs.ctypes.data_as(_dp) if rhs is not None else None))

    def rhs(self) -> np.ndarray:
        r = np.zeros((self.problem.nfree, self.problem.nrhs), np.float64, order="F")
        self._check(self.lib.gsb200_download_rhs(self._h, r.ctypes.data_as(_dp)))
        return r

    def rhs_into(self, rhs: np.ndarray) -> None:
        self._check(self.lib.gsb200_download_rhs(self._h, rhs.ctypes.data_as(_dp)))

    def scipy_matrix(self):
        import scipy.sparse as sp
        o, i, v = self.matrix()
        n = self.problem.nfree
        return sp.csc_matrix((v, i, o), shape=(n, n))

    # --- device-side access -----------------------------------------------------------
    def device_view(self) -> DeviceView:
        v = DeviceView()
        self._check(self.lib.gsb200_device_view_get(self._h, C.byref(v)))
        return v

    def timings(self) -> Timings:
        t = Timings()
        self._check(self.lib.gsb200_timings_get(self._h, C.byref(t)))
        return t

    def jit_launches(self) -> int:
        """Geometry launches of the last assemble() that ran the NVRTC-compiled source term."""
        n = C.c_int(0)
        self._check(self.lib.gsb200_jit_launches(self._h, C.byref(n)))
        return n.value

    # --- consumer -------------------------------------------------------
    def spmv(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self._check(self.lib.gsb200_spmv_host(self._h, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp)))
        return y

    def spmv_info(self):
        """(columns on the regular-stencil path, offset tables) of the SpMV."""
        n, t = C.c_int64(0), C.c_int32(0)
        self._check(self.lib.gsb200_spmv_info(self._h, C.byref(n), C.byref(t)))
        return n.value, t.value

    # --- multi-GPU: the library's own NCCL communicator, or a caller-supplied reduction ------------
    def comm_init(self, unique_id: bytes) -> None:
        """gsb200_comm_init (collective): `unique_id` = the 128 bytes rank 0 got from comm_unique_id()."""
        buf = C.create_string_buffer(bytes(unique_id), capi.COMM_ID_BYTES)
        self._check(self.lib.gsb200_comm_init(self._h, buf))

    def set_allreduce(self, fn) -> None:
        """gsb200_set_allreduce: fn(buffer_address, count, stream) sums `count` doubles in place over the ranks."""
        def trampoline(ctx, buf, count, stream):
            try:
                fn(C.addressof(buf.contents), int(count), stream)
                return 0
            except Exception:       # noqa: the C side turns the status into an error
                return 1
        self._ar_cb = capi.ALLREDUCE_FN(trampoline)        # keep the callback object alive
        self._check(self.lib.gsb200_set_allreduce(self._h, self._ar_cb, None))

    def exchange(self) -> None:
        """gsb200_exchange: coupled columns + right-hand side summed over the ranks (after assemble)."""
        self._check(self.lib.gsb200_exchange(self._h))

    def comm_stats(self):
        b, n = C.c_int64(0), C.c_int32(0)
        self._check(self.lib.gsb200_comm_stats(self._h, C.byref(b), C.byref(n)))
        return b.value, n.value

    def cg_solve(self, b: Optional[np.ndarray] = None, max_iter: int = 1000, tol: float = 1e-10, check_every: int = 10, want_x: bool = True):
        """gsb200_cg_solve: Jacobi-CG across the ranks; b None = the assembled (exchanged) right-hand side."""
        bp = None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.float64).ravel()
            bp = b.ctypes.data_as(_dp)
        x = np.zeros(self.problem.nfree) if want_x else None
        it, res = C.c_int(0), C.c_double(0)
        self._check(self.lib.gsb200_cg_solve(self._h, bp, x.ctypes.data_as(_dp) if want_x else None, max_iter, tol, check_every,
                                       C.byref(it), C.byref(res)))
        return x, it.value, res.value

    def field_norms(self, u: np.ndarray, exact=None, exact_grad=None) -> np.ndarray:
        """gsb200_field_norms: [int (u_h-u_ex)^2, int |grad(u_h-u_ex)|^2, int u_h^2, int |grad u_h|^2] of the discrete field with free
        coefficients u (exact / exact_grad: CompiledProgram / list of dim CompiledPrograms, or None)."""
        u = np.ascontiguousarray(u, dtype=np.float64).ravel()
        def prog(cp):
            return capi.Program(len(cp.ops), cp.ops.ctypes.data_as(_ip), len(cp.consts), cp.consts.ctypes.data_as(_dp))
        ex = C.byref(prog(exact)) if exact is not None else None
        eg = None
        if exact_grad is not None:
            arr = (capi.Program * len(exact_grad))(*[prog(cp) for cp in exact_grad])
            eg = arr
        out = np.zeros(4)
        self._check(self.lib.gsb200_field_norms(self._h, u.ctypes.data_as(_dp), ex, eg, out.ctypes.data_as(_dp)))
        return out

    def project_dirichlet(self, sides, max_iter: int = 2000, tol: float = 1e-13):
        """gsb200_project_dirichlet: sides = [(patch, side, CompiledProgram g), ...]; returns (fixed values, iterations, rel. residual)."""
        nm = (capi.Neumann * max(len(sides), 1))()
        for i, (patch, side, cps) in enumerate(sides):
            cps = list(cps) if isinstance(cps, (list, tuple)) else [cps]          # one program per component of the space
            nm[i].patch, nm[i].side, nm[i].ndata = int(patch), int(side), len(cps)
            for c, cp in enumerate(cps):
                nm[i].data[c].nops = len(cp.ops); nm[i].data[c].ops = cp.ops.ctypes.data_as(_ip)
                nm[i].data[c].nconsts = len(cp.consts); nm[i].data[c].consts = cp.consts.ctypes.data_as(_dp)
        self._keep_sides = sides
        out = np.zeros(max(self.problem.nfixed, 1))
        it, res = C.c_int(0), C.c_double(0)
        self._check(self.lib.gsb200_project_dirichlet(self._h, nm, len(sides), max_iter, tol, out.ctypes.data_as(_dp), C.byref(it), C.byref(res)))
        return out[:self.problem.nfixed], it.value, res.value

    def cg_info(self):
        """(device ms of the last solve's iteration loop, halo-exchange mode?)"""
        ms, h = C.c_double(0), C.c_int32(0)
        self._check(self.lib.gsb200_cg_info(self._h, C.byref(ms), C.byref(h)))
        return ms.value, bool(h.value)

    def cg(self, b: np.ndarray, max_iter: int = 1000, tol: float = 1e-10):
        b = np.ascontiguousarray(b, dtype=np.float64).ravel()
        x = np.zeros_like(b)
        it, res = C.c_int(0), C.c_double(0)
        self._check(self.lib.gsb200_cg_host(self._h, b.ctypes.data_as(_dp), x.ctypes.data_as(_dp), max_iter, tol,
                                      C.byref(it), C.byref(res)))
        return x, it.value, res.value


def assemble_host(problem: Problem, device: int = 0):
    """gsb200_assemble_host: host buffers in, host buffers out — what the C++ shim
    gsPoissonAssemblerB200::assemble() calls."""
    lib = capi.load_library()
    nnz = C.c_int64(0)
    check(lib.gsb200_assemble_host(C.byref(problem.struct), device, C.byref(nnz), None, None, None, None))
    outer = np.zeros(problem.nfree + 1, np.int32)
    inner = np.zeros(nnz.value, np.int32)
    values = np.zeros(nnz.value)
    rhs = np.zeros((problem.nfree, problem.nrhs), order="F")
    check(lib.gsb200_assemble_host(C.byref(problem.struct), device, C.byref(nnz), outer.ctypes.data_as(_ip),
                                   inner.ctypes.data_as(_ip), values.ctypes.data_as(_dp), rhs.ctypes.data_as(_dp)))
    return outer, inner, values, rhs


def measure_peaks(device: int = 0):
    lib = capi.load_library()
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    check(lib.gsb200_measure_peaks(device, C.byref(a), C.byref(b), C.byref(c)))
    return {"fp64_tflops": a.value, "dmma_tflops": b.value, "hbm_gbs": c.value}
