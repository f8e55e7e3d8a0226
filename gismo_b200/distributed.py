"""Multi-GPU plumbing: one process per GPU.  The data path is in the library (include/gsb200.h,
csrc/consumer.cuh): it owns the NCCL communicator, sums the coupled interface columns and the right-hand side
(gsb200_exchange, K4) and runs the CG consumer across the ranks (gsb200_cg_solve).  What is left here is the
bootstrap a host transport has to do - ship the 128-byte NCCL id from rank 0 to the others - done with
torch.distributed, and helpers for the tests (a gloo all-reduce for the GPU-less container, gluing the ranks'
pieces together for verification).

Row/column ownership means the data path needs no collective for a single patch (each rank assembles and keeps
the CSC columns of its slab).  For multi-patch problems the columns of the coupled interface DOFs — the tail
block(s) of the numbering (gsDofMapper.cpp:281-323) — receive contributions from every rank that owns an
adjacent patch: they are patterned identically on all ranks and their value blocks are summed.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .capi import Problem, check


def init_comm(assembler, group: Optional[dist.ProcessGroup] = None) -> None:
    """Give `assembler` (created with problem.rank / problem.nranks = this process's rank / world size) its NCCL
    communicator: rank 0 draws the id (gsb200_comm_unique_id), torch.distributed ships it, every rank joins
    (gsb200_comm_init).  Collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return
    box = [None]
    if dist.get_rank(group) == 0:
        buf = C.create_string_buffer(capi.COMM_ID_BYTES)
        check(assembler.lib.gsb200_comm_unique_id(buf))
        box[0] = bytes(buf.raw)
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    assembler.comm_init(box[0])


def use_torch_allreduce(assembler, group: Optional[dist.ProcessGroup] = None, device: Optional[int] = None) -> None:
    """Route the library's reductions through torch.distributed (gsb200_set_allreduce): the buffer the library hands over
    is aliased as a tensor - host memory for the interpreter build of the tests (gloo), device memory otherwise."""
    def allreduce(addr: int, count: int, stream) -> None:
        if device is None:
            t = torch.from_numpy(np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_double)), shape=(count,)))
        else:                                   # the library's kernels run on its own stream: order the two by hand
            torch.cuda.synchronize(device)
            t = device_tensor(addr, count, torch.float64, device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if device is not None:
            torch.cuda.synchronize(device)
    assembler.set_allreduce(allreduce)


class _CudaView:
    """Minimal __cuda_array_interface__ wrapper so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, n: int, dtype: torch.dtype, device: int) -> torch.Tensor:
    typestr = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_CudaView(ptr, n, typestr), device=torch.device("cuda", device))


def first_coupled_column(problem: Problem) -> int:
    """Global index of the first free DOF shared by more than one (patch, local) pre-image."""
    runs = coupled_column_ranges(problem)
    return runs[0][0] if runs else problem.nfree


def coupled_column_ranges(problem: Problem):
    """Contiguous runs [a, b) of global columns shared by more than one (patch, local) pre-image: one tail run for a scalar
    space, one per component block for a vector-valued one (component-major numbering, gsDofMapper.cpp:255-265).  The
    library derives the same runs for gsb200_exchange; this copy serves the verification helpers."""
    counts = np.zeros(problem.nfree + 1, dtype=np.int32)
    for p in problem.patches:
        g = p.dofmap[p.dofmap < problem.nfree]
        np.add.at(counts, g, 1)
    multi = np.concatenate([[False], counts[:problem.nfree] > 1, [False]])
    edges = np.flatnonzero(multi[1:] != multi[:-1])
    return [(int(edges[k]), int(edges[k + 1])) for k in range(0, len(edges), 2)]


def patch_owners(costs, nranks: int):
    """The library's patch -> rank assignment (gsb200_create): longest processing time first onto the least loaded rank."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])        # stable, like std::stable_sort
    load = [0] * nranks
    owner = [0] * len(costs)
    for i in order:
        r = min(range(nranks), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return owner


def merge_rank_matrices(parts, n: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Verification helper: glue per-rank CSC pieces (owned columns + exchanged coupled columns, which every rank holds)
    into one CSC: each column is taken from the first rank that stores it."""
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
