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
