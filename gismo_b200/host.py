"""Host-side problem builders mirroring the reference's set-up objects.

Only what the synthetic configs of SURVEY 8(d) need so that benchmarks and tests can run on
a machine without G+Smo: knot vectors with the reference's uniform refinement, the unit
square/cube B-spline geometries of gsNurbsCreator, and the DOF numbering of gsDofMapper.
In a G+Smo application these come from the real objects through
gismo_b200/host/gsB200Flatten.h instead.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .capi import PatchData, Problem, FORM_POISSON, FORM_ELASTICITY, CompiledProgram


class KnotVector:
    """gsKnotVector<T> (gsKnotVector.h:79): sorted knots with repetitions + degree."""

    def __init__(self, degree: int, knots: Sequence[float]):
        self.degree = int(degree)
        self.knots = np.asarray(knots, dtype=np.float64).copy()

    @classmethod
    def open_uniform(cls, first: float, last: float, interior: int, mult_ends: int, degree: Optional[int] = None):
        """gsKnotVector(first,last,interior,mult_ends) as used by gsNurbsCreator (KV(0,1,0,2))."""
        deg = mult_ends - 1 if degree is None else degree
        h = (last - first) / (interior + 1)
        mid = [first + i * h for i in range(1, interior + 1)]
        return cls(deg, [first] * mult_ends + mid + [last] * mult_ends)

    def unique(self) -> np.ndarray:
        return np.unique(self.knots)

    def uniformRefine(self, numKnots: int = 1, mult: int = 1) -> None:
        """gsKnotVector::uniformRefine -> getUniformRefinementKnots (gsKnotVector.hpp:1048-1063):
        new knots are prev + i*((next-prev)/(numKnots+1)), evaluated in exactly that order."""
        u = self.unique()
        new = []
        prev = self.knots[0]
        for nxt in u[1:]:
            step = (nxt - prev) / float(numKnots + 1)
            for i in range(1, numKnots + 1):
                new.extend([prev + float(i) * step] * mult)
            prev = nxt
        self.knots = np.sort(np.concatenate([self.knots, np.asarray(new, dtype=np.float64)]), kind="stable")

    def degreeElevate(self, i: int = 1) -> None:
        """raise the degree keeping smoothness: every distinct knot gains i repetitions."""
        u = self.unique()
        self.knots = np.sort(np.concatenate([self.knots, np.repeat(u, i)]), kind="stable")
        self.degree += i

    def setDegree(self, p: int) -> None:
        """gsBasis::setDegree on a B-spline basis = elevate/reduce to degree p (elevation only here)."""
        if p < self.degree:
            raise ValueError("degree reduction is not needed by the synthetic configs")
        if p > self.degree:
            self.degreeElevate(p - self.degree)

    @property
    def size(self) -> int:
        return len(self.knots) - self.degree - 1

    def copy(self) -> "KnotVector":
        return KnotVector(self.degree, self.knots)


def bspline_box(dim: int, r: float = 1.0, origin: Sequence[float] = (0.0, 0.0, 0.0)):
    """gsNurbsCreator::BSplineSquare(r,x,y) / BSplineCube(r,x,y,z) (gsNurbsCreator.hpp:731-746):
    degree-1 single-element patch.  NOTE the reference's cube is centred at (x,y,z) while its
    square has its lower-left corner there."""
    kv = KnotVector.open_uniform(0.0, 1.0, 0, 2)
    n = 2 ** dim
    C = np.zeros((n, dim))
    for idx in range(n):
        for k in range(dim):
            bit = (idx >> k) & 1
            C[idx, k] = (bit - 0.5) * r + origin[k] if dim == 3 else bit * r + origin[k]
    return [kv.copy() for _ in range(dim)], C


class DofMapper:
    """Numbering of gsDofMapper (gsDofMapper.cpp:240-344) for one scalar component, or
    component-major blocks for several (gsDofMapper.cpp:255-265): free DOFs first (patch-then-
    local order, coupled ones after all standard ones in first-appearance order), eliminated last."""

    def __init__(self, patch_sizes: Sequence[int], ncomp: int = 1):
        self.sizes = [int(s) for s in patch_sizes]
        self.offset = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.ncomp = ncomp
        n = int(self.offset[-1])
        # 0 = free, >0 coupling id, <0 eliminated id (same encoding as the reference, per component)
        self.tag = np.zeros((ncomp, n), dtype=np.int64)
        self._next_cpl = [1] * ncomp
        self._next_elim = -1

    def eliminate(self, patch: int, local: Iterable[int], comp: Optional[int] = None) -> None:
        comps = range(self.ncomp) if comp is None else [comp]
        local = np.asarray(list(local) if not isinstance(local, np.ndarray) else local, dtype=np.int64)
        ks = self.offset[patch] + local
        if all(np.all(self.tag[c, ks] == 0) for c in comps):   # common case: all still free
            for c in comps:
                self.tag[c, ks] = self._next_elim - np.arange(len(ks))
                self._next_elim -= len(ks)
            return
        for i in local:
            for c in comps:
                k = self.offset[patch] + i
                old = self.tag[c, k]
                if old == 0:
                    self.tag[c, k] = self._next_elim
                    self._next_elim -= 1
                elif old > 0:
                    self.tag[c][self.tag[c] == old] = self._next_elim
                    self._next_elim -= 1

    def match(self, p1: int, i1: int, p2: int, i2: int) -> None:
        """matchDof (gsDofMapper.cpp:99-157) for all components."""
        for c in range(self.ncomp):
            t = self.tag[c]
            k1, k2 = self.offset[p1] + i1, self.offset[p2] + i2
            d1, d2 = t[k1], t[k2]
            if d1 > d2:
                d1, d2, k1, k2 = d2, d1, k2, k1
            if d1 < 0:
                if d2 < 0:
                    t[t == d2] = d1
                elif d2 == 0:
                    t[k2] = d1
                else:
                    t[t == d2] = d1
            elif d1 == 0:
                if d2 == 0:
                    t[k1] = t[k2] = self._next_cpl[c]
                    self._next_cpl[c] += 1
                else:
                    t[k1] = d2
            else:
                if d1 != d2:
                    t[t == d2] = d1

    def finalize(self) -> None:
        n = int(self.offset[-1])
        self.index = np.zeros((self.ncomp, n), dtype=np.int64)
        free_counts, elim_counts = [], []
        staged = []
        for c in range(self.ncomp):
            t = self.tag[c]
            nstd = int(np.sum(t == 0))
            loc = np.zeros(n, dtype=np.int64)
            kind = (t < 0).astype(np.int8)      # 0 free (standard/coupled), 1 eliminated
            loc[t == 0] = np.arange(nstd)
            cpl_ids, elim_ids = {}, {}
            for sel, base, store in ((t > 0, nstd, cpl_ids), (t < 0, 0, elim_ids)):
                if np.any(sel):
                    ids, first_pos, inv = np.unique(t[sel], return_index=True, return_inverse=True)
                    order = np.argsort(first_pos, kind="stable")      # first-appearance order
                    rank = np.empty(len(ids), dtype=np.int64)
                    rank[order] = np.arange(len(ids))
                    loc[sel] = base + rank[inv]
                    store.update({int(i): 0 for i in ids})
            free_counts.append(nstd + len(cpl_ids))
            elim_counts.append(len(elim_ids))
            staged.append((loc, kind))
        self.nfree = int(sum(free_counts))
        self.nfixed = int(sum(elim_counts))
        fo = np.concatenate([[0], np.cumsum(free_counts)])
        eo = np.concatenate([[0], np.cumsum(elim_counts)])
        for c, (loc, kind) in enumerate(staged):
            self.index[c] = np.where(kind == 0, loc + fo[c], loc + self.nfree + eo[c])

    def patch_map(self, patch: int) -> np.ndarray:
        """ncomp blocks of the patch's global indices (the `dofmap` field of the C ABI)."""
        a, b = self.offset[patch], self.offset[patch + 1]
        return np.ascontiguousarray(self.index[:, a:b].astype(np.int32)).ravel()


def boundary_indices(nfun: Sequence[int]) -> np.ndarray:
    """All basis functions on the boundary of a tensor patch, in the order
    gsBasis::allBoundary() returns them (sorted ascending local index)."""
    dim = len(nfun)
    grids = np.meshgrid(*[np.arange(n) for n in nfun], indexing="ij")
    on_b = np.zeros(nfun, dtype=bool)
    for k in range(dim):
        on_b |= (grids[k] == 0) | (grids[k] == nfun[k] - 1)
    # local index with direction 0 fastest
    idx = np.zeros(nfun, dtype=np.int64)
    stride = 1
    for k in range(dim):
        idx += grids[k] * stride
        stride *= nfun[k]
    return np.sort(idx[on_b])


def poisson_box_problem(dim: int, degree: int, nelem: Sequence[int] | int, rhs_program: Optional[CompiledProgram] = None,
                        rank: int = 0, nranks: int = 1, form: int = FORM_POISSON) -> Problem:
    """BSplineSquare/BSplineCube, setDegree(p), uniformRefine(m-1),
    homogeneous Dirichlet on every side, dirichlet::elimination."""
    if isinstance(nelem, int):
        nelem = [nelem] * dim
    gkv, C = bspline_box(dim)
    skv = []
    for k in range(dim):
        kv = gkv[k].copy()
        kv.setDegree(degree)
        if nelem[k] > 1:
            kv.uniformRefine(nelem[k] - 1)
        skv.append(kv)
    nfun = [kv.size for kv in skv]
    mapper = DofMapper([int(np.prod(nfun))])
    # gsDofMapper eliminates boundary DOFs side by side (markBoundary), ids in first-hit order;
    # finalize renumbers them by first appearance in local order, so the order of calls is irrelevant
    mapper.eliminate(0, boundary_indices(nfun))
    mapper.finalize()
    patch = PatchData([degree] * dim, [kv.knots for kv in skv], [kv.degree for kv in gkv], [kv.knots for kv in gkv], C,
                      mapper.patch_map(0))
    return Problem([patch], mapper.nfree, mapper.nfixed, form=form, ncomp=1,
                   rhs_programs=[rhs_program] if rhs_program is not None else None, rank=rank, nranks=nranks)


def multipatch_grid_problem(dim: int, degree: int, grid: Sequence[int], nelem: int, rhs_programs=None, form: int = FORM_POISSON,
                            coef=(0.0, 0.0), rank: int = 0, nranks: int = 1) -> Problem:
    """SURVEY 8(d) configs 3/4 in synthetic form: a conforming grid of unit boxes (gsNurbsCreator::BSplineSquareGrid /
    BSplineCubeGrid patch order: LAST direction fastest), every patch with its own degree-p basis of `nelem` elements per
    direction, C0 gluing of all interfaces (iFace::glue), homogeneous Dirichlet on the outer boundary (elimination).
    The numbering is gsDofMapper's (gsDofMapper.cpp:240-344): standard free DOFs in patch-then-local order, then the
    coupled interface DOFs in first-appearance order, then the eliminated ones; vector-valued forms get component-major
    blocks.  Built with array operations on the global lattice of glued functions (DofMapper.match is per DOF)."""
    grid = [int(g) for g in grid][:dim]
    ncomp = dim if form == FORM_ELASTICITY else 1
    nf1 = nelem + degree                                   # functions per direction and patch
    lat = [g * (nf1 - 1) + 1 for g in grid]                 # glued lattice per direction
    patches_abc = []
    for a in np.ndindex(*grid):                             # last index fastest = reference order
        patches_abc.append(a)
    npatch = len(patches_abc)
    nloc = nf1 ** dim
    # node id of every (patch, local function); local index has direction 0 fastest
    loc = np.arange(nloc)
    lk = [(loc // (nf1 ** k)) % nf1 for k in range(dim)]
    node = np.empty((npatch, nloc), dtype=np.int64)
    outer_b = np.zeros((npatch, nloc), dtype=bool)
    for ip, abc in enumerate(patches_abc):
        nid = np.zeros(nloc, dtype=np.int64); stride = 1; ob = np.zeros(nloc, dtype=bool)
        for k in range(dim):
            gk = abc[k] * (nf1 - 1) + lk[k]
            nid += gk * stride; stride *= lat[k]
            ob |= (gk == 0) | (gk == lat[k] - 1)
        node[ip] = nid; outer_b[ip] = ob
    flat_node = node.ravel(); flat_ob = outer_b.ravel()
    counts = np.bincount(flat_node, minlength=int(np.prod(lat)))
    multi = counts[flat_node] > 1
    std = ~flat_ob & ~multi
    cpl = ~flat_ob & multi
    nstd = int(std.sum())
    index1 = np.zeros(npatch * nloc, dtype=np.int64)
    index1[std] = np.arange(nstd)

    def first_appearance_rank(sel):
        ids, first_pos, inv = np.unique(flat_node[sel], return_index=True, return_inverse=True)
        order = np.argsort(first_pos, kind="stable")
        rk = np.empty(len(ids), dtype=np.int64); rk[order] = np.arange(len(ids))
        return rk[inv], len(ids)
    ncpl = nelim = 0
    if cpl.any():
        r, ncpl = first_appearance_rank(cpl); index1[cpl] = nstd + r
    if flat_ob.any():
        r, nelim = first_appearance_rank(flat_ob); index1[flat_ob] = r       # offset added below
    nfree1 = nstd + ncpl
    nfree, nfixed = nfree1 * ncomp, nelim * ncomp
    patches = []
    for ip, abc in enumerate(patches_abc):
        gkv, C = bspline_box(dim, 1.0, [float(v) + (0.5 if dim == 3 else 0.0) for v in abc] + [0.0] * (3 - dim))
        skv = []
        for k in range(dim):
            kv = gkv[k].copy(); kv.setDegree(degree)
            if nelem > 1:
                kv.uniformRefine(nelem - 1)
            skv.append(kv)
        i1 = index1[ip * nloc:(ip + 1) * nloc]; ob = flat_ob[ip * nloc:(ip + 1) * nloc]
        dm = np.concatenate([np.where(ob, i1 + nfree + c * nelim, i1 + c * nfree1) for c in range(ncomp)]).astype(np.int32)
        patches.append(PatchData([degree] * dim, [kv.knots for kv in skv], [kv.degree for kv in gkv], [kv.knots for kv in gkv], C, dm))
    return Problem(patches, nfree, nfixed, form=form, ncomp=ncomp, coef=tuple(coef) + (0.0,) * (4 - len(coef)),
                   rhs_programs=rhs_programs, rank=rank, nranks=nranks)


def _side_locals(nfun: Sequence[int], side: int) -> np.ndarray:
    """Local indices of the functions on side 1..4 of a 2-D patch (gsBoundary.h:58-60: west, east, south, north),
    ordered by the running index of the other direction."""
    n0, n1 = int(nfun[0]), int(nfun[1])
    if side == 1:
        return np.arange(n1) * n0
    if side == 2:
        return np.arange(n1) * n0 + (n0 - 1)
    if side == 3:
        return np.arange(n0)
    return np.arange(n0) + (n1 - 1) * n0


def multipatch_topology_2d(problem: Problem):
    """Interfaces and Dirichlet sides of a flattened 2-D multi-patch discretisation, read off its DOF map: two patch sides
    that carry the same global indices are glued (gsBoxTopology interface; reversed order = flipped orientation), a side
    whose functions are all eliminated and glued to nobody is a Dirichlet side.  This is how the topology of a reference
    geometry (filedata/domain2d/yeti_mp2.xml, read by the reference when the fixture was made) reaches the GPU box, where
    the reference's XML reader does not exist."""
    sides = {}
    for ip, pa in enumerate(problem.patches):
        nfun = [len(pa.space_knots[k]) - pa.space_degree[k] - 1 for k in range(2)]
        for s in (1, 2, 3, 4):
            sides[(ip, s)] = pa.dofmap[_side_locals(nfun, s)]
    interfaces, glued = [], set()
    keys = sorted(sides)
    for a in range(len(keys)):
        for b in range(a + 1, len(keys)):
            ka, kb = keys[a], keys[b]
            if ka[0] == kb[0] or len(sides[ka]) != len(sides[kb]):
                continue
            if np.array_equal(sides[ka], sides[kb]):
                interfaces.append((ka[0], ka[1], kb[0], kb[1], False)); glued.update((ka, kb))
            elif np.array_equal(sides[ka], sides[kb][::-1]):
                interfaces.append((ka[0], ka[1], kb[0], kb[1], True)); glued.update((ka, kb))
    dirichlet = [k for k in keys if k not in glued and np.all(sides[k] >= problem.nfree)]
    return interfaces, dirichlet


def refine_multipatch_2d(problem: Problem, degree: int, nelem: int, rhs_programs=None, rank: int = 0, nranks: int = 1) -> Problem:
    """The discretisation the reference builds from the same geometry with setDegree(degree) + uniformRefine(nelem - 1)
    (oracle/ref_driver.cpp geometry 4), at any refinement: knots from the patches' geometry bases, interfaces and Dirichlet
    sides from multipatch_topology_2d(problem), numbering of gsDofMapper (gsDofMapper.cpp:240-344: standard free DOFs in
    patch-then-local order, coupled ones by first appearance, eliminated ones last).  Scalar spaces."""
    assert problem.dim == 2 and problem.ncomp == 1
    interfaces, dirichlet = multipatch_topology_2d(problem)
    skvs, nfuns = [], []
    for pa in problem.patches:
        kvs = []
        for k in range(2):
            kv = KnotVector(pa.geo_degree[k], pa.geo_knots[k])
            kv.setDegree(degree)
            if nelem > 1:
                kv.uniformRefine(nelem - 1)
            kvs.append(kv)
        skvs.append(kvs); nfuns.append([kv.size for kv in kvs])
    sizes = [n[0] * n[1] for n in nfuns]
    offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    total = int(offset[-1])
    parent = np.arange(total, dtype=np.int64)

    def find(i):
        r = i
        while parent[r] != r:
            r = parent[r]
        while parent[i] != r:
            parent[i], i = r, parent[i]
        return r
    for pa_, sa, pb_, sb, flip in interfaces:
        la = offset[pa_] + _side_locals(nfuns[pa_], sa)
        lb = offset[pb_] + _side_locals(nfuns[pb_], sb)
        assert len(la) == len(lb), "non-conforming interface"
        if flip:
            lb = lb[::-1]
        for x, y in zip(la.tolist(), lb.tolist()):
            rx, ry = find(x), find(y)
            if rx != ry:
                parent[max(rx, ry)] = min(rx, ry)
    node = np.arange(total, dtype=np.int64)            # only functions on glued sides can have another root
    for pa_, sa, pb_, sb, _ in interfaces:
        for ip, s in ((pa_, sa), (pb_, sb)):
            for x in (offset[ip] + _side_locals(nfuns[ip], s)).tolist():
                node[x] = find(x)
    elim_node = np.zeros(total, dtype=bool)
    for ip, s in dirichlet:
        elim_node[node[offset[ip] + _side_locals(nfuns[ip], s)]] = True
    elim = elim_node[node]
    multi = np.bincount(node, minlength=total)[node] > 1
    std, cpl = ~elim & ~multi, ~elim & multi
    index = np.zeros(total, dtype=np.int64)
    nstd = int(std.sum())
    index[std] = np.arange(nstd)

    def first_appearance_rank(sel):
        ids, first_pos, inv = np.unique(node[sel], return_index=True, return_inverse=True)
        order = np.argsort(first_pos, kind="stable")
        rk = np.empty(len(ids), dtype=np.int64); rk[order] = np.arange(len(ids))
        return rk[inv], len(ids)
    ncpl = nelim = 0
    if cpl.any():
        r, ncpl = first_appearance_rank(cpl); index[cpl] = nstd + r
    nfree = nstd + ncpl
    if elim.any():
        r, nelim = first_appearance_rank(elim); index[elim] = nfree + r
    patches = []
    for ip, pa in enumerate(problem.patches):
        patches.append(PatchData([degree] * 2, [kv.knots for kv in skvs[ip]], pa.geo_degree, pa.geo_knots, pa.geo_coefs,
                                 index[offset[ip]:offset[ip + 1]].astype(np.int32), pa.geo_weights))
    return Problem(patches, nfree, nelim, form=problem.form, ncomp=1, rhs_programs=rhs_programs if rhs_programs is not None else problem.rhs_programs,
                   rank=rank, nranks=nranks)
