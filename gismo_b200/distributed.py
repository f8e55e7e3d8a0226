"""Multi-GPU plumbing: one process per GPU, torch.distributed for the little that
has to be exchanged.

Row/column ownership means the data path needs no collective for a single patch (each rank
assembles and keeps the CSC columns of its slab).  For multi-patch problems the columns of the
coupled interface DOFs — the contiguous tail block of the numbering (gsDofMapper.cpp:281-323) —
receive contributions from every rank that owns an adjacent patch: they are patterned identically
on all ranks and their value block (and the rhs) is summed with one all_reduce.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .capi import Problem


def first_coupled_column(problem: Problem) -> int:
    """Global index of the first free DOF shared by more than one (patch, local) pre-image."""
    counts = np.zeros(problem.nfree, dtype=np.int32)
    for p in problem.patches:
        g = p.dofmap[p.dofmap < problem.nfree]
        np.add.at(counts, g, 1)
    multi = np.nonzero(counts > 1)[0]
    return int(multi[0]) if len(multi) else problem.nfree


class _CudaView:
    """Minimal __cuda_array_interface__ wrapper so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, n: int, dtype: torch.dtype, device: int) -> torch.Tensor:
    typestr = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_CudaView(ptr, n, typestr), device=torch.device("cuda", device))


def coupled_column_ranges(problem: Problem):
    """Contiguous runs [a, b) of global columns shared by more than one (patch, local) pre-image: one tail run for a scalar
    space, one per component block for a vector-valued one (component-major numbering, gsDofMapper.cpp:255-265)."""
    counts = np.zeros(problem.nfree + 1, dtype=np.int32)
    for p in problem.patches:
        g = p.dofmap[p.dofmap < problem.nfree]
        np.add.at(counts, g, 1)
    multi = np.concatenate([[False], counts[:problem.nfree] > 1, [False]])
    edges = np.flatnonzero(multi[1:] != multi[:-1])
    return [(int(edges[k]), int(edges[k + 1])) for k in range(0, len(edges), 2)]


def reduce_coupled_columns(values: torch.Tensor, rhs: torch.Tensor, outer: np.ndarray, c0,
                           group: Optional[dist.ProcessGroup] = None) -> None:
    """In place: sum the value blocks of the coupled columns (c0 = first coupled column of a scalar space, or the list of
    runs from coupled_column_ranges) and the whole rhs over all ranks."""
    n = len(outer) - 1
    runs = [(int(c0), n)] if np.isscalar(c0) else list(c0)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for a, b in runs:
            if a < b and int(outer[b]) > int(outer[a]):
                block = values[int(outer[a]):int(outer[b])]
                dist.all_reduce(block, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(rhs, op=dist.ReduceOp.SUM, group=group)


def merge_rank_matrices(parts, c0: int, n: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Verification helper: glue per-rank CSC pieces (owned columns + reduced coupled block) into one CSC."""
    lens = np.zeros(n, dtype=np.int64)
    for outer, _, _ in parts:
        ln = np.diff(outer.astype(np.int64))
        take = (lens == 0)
        lens[take] = ln[take]
    new_outer = np.concatenate([[0], np.cumsum(lens)])
    inner = np.zeros(int(new_outer[-1]), dtype=np.int32)
    values = np.zeros(int(new_outer[-1]), dtype=np.float64)
    done = np.zeros(n, dtype=bool)
    for outer, inn, val in parts:
        ln = np.diff(outer.astype(np.int64))
        for c in np.nonzero((ln > 0) & ~done)[0]:
            a, b = int(outer[c]), int(outer[c + 1])
            inner[new_outer[c]:new_outer[c + 1]] = inn[a:b]
            values[new_outer[c]:new_outer[c + 1]] = val[a:b]
            done[c] = True
    return new_outer.astype(np.int32), inner, values


class DistributedCG:
    """Jacobi-preconditioned CG on the device-resident, column-partitioned matrix (mirrors
    gsSparseSolver<>::CGDiagonal, gsSparseSolver.h:71-72).  Every rank keeps full-length vectors; the matrix product is the
    library's warp-per-row SpMV over the columns the rank owns (gsb200_spmv_device) followed by ONE all_reduce of y, dot products
    are local (replicated vectors).  world_size 1 works without a process group."""

    def __init__(self, assembler, device: int, group: Optional[dist.ProcessGroup] = None, reduced: bool = True):
        """reduced: the coupled columns (multi-patch) already hold the all-reduced values on every rank that stores them
        (reduce_coupled_columns was called); irrelevant for single-patch slabs."""
        self.A, self.device, self.group = assembler, device, group
        self.n = assembler.problem.nfree
        self.multi = dist.is_initialized() and dist.get_world_size(group) > 1
        from .capi import check
        v = assembler.device_view()
        outer = device_tensor(v.outer, self.n + 1, torch.int64, device)
        stored = (outer[1:] - outer[:-1]) > 0
        # diagonal of the columns this rank stores (library kernel: no nnz-sized temporaries), summed over the ranks; coupled
        # columns hold the full sums on every rank that patterns them after the exchange: divide by the number of holders
        diag = torch.empty(self.n, dtype=torch.float64, device=outer.device)
        check(assembler.lib.gsb200_diag_device(assembler._h, diag.data_ptr()))
        diag = torch.where(stored, diag, torch.zeros_like(diag))
        holders = stored.to(torch.float64)
        if self.multi:
            dist.all_reduce(diag, group=group); dist.all_reduce(holders, group=group)
        self.holders = torch.clamp(holders, min=1.0) if reduced else torch.ones_like(holders)
        self.diag = diag / self.holders
        self.diag = torch.where(self.diag == 0, torch.ones_like(self.diag), self.diag)

    def matvec(self, x: torch.Tensor) -> torch.Tensor:
        from .capi import check
        y = torch.empty_like(x)
        check(self.A.lib.gsb200_spmv_device(self.A._h, x.data_ptr(), y.data_ptr()))
        if self.multi:
            y /= self.holders                     # columns patterned on several ranks carry the same (already reduced) values
            dist.all_reduce(y, group=self.group)
        return y

    def solve(self, b: torch.Tensor, max_iter: int = 1000, tol: float = 1e-10):
        x = torch.zeros_like(b); r = b.clone(); z = r / self.diag; p = z.clone()
        rz = torch.dot(r, z); bb = torch.dot(b, b); it = 0
        while it < max_iter:
            q = self.matvec(p)
            alpha = rz / torch.dot(p, q)
            x += alpha * p; r -= alpha * q
            rr = torch.dot(r, r)
            it += 1
            if float(rr) <= tol * tol * float(bb):
                break
            z = r / self.diag
            rz2 = torch.dot(r, z)
            p = z + (rz2 / rz) * p; rz = rz2
        return x, it, float(torch.sqrt(torch.dot(r, r) / bb))
