/** @file gsB200Flatten.h

    Host-side extraction layer of the B200 assembly back-end: flattens the G+Smo
    objects an assembler holds (gsMultiPatch, gsMultiBasis, gsDofMapper, Dirichlet
    values, options, source term) into the plain-old-data problem description of the
    C ABI (include/gsb200.h).  Header-only C++ over gismo + Eigen, as the reference's
    host side is; it performs no arithmetic of the hot path.

    Reference conventions relied upon (file:line in gismo v24.08.0):
      - knots with repetitions:      gsKnotVector::data()/size()   gsKnotVector.h:242,285
      - per-direction component:     gsTensorBSplineBasis::knots(i) gsTensorBSplineBasis.h:192
      - control points, col-major:   gsGeometry::coefs()            gsGeometry.h:343
      - NURBS weights / source:      gsRationalBasis::weights()     gsRationalBasis.h:290
      - global numbering:            gsDofMapper::index(i,k,c)      gsDofMapper.h:325-329
      - free / eliminated counts:    gsDofMapper::freeSize()/boundarySize()
      - source-term strings:         gsFunctionExpr::expression(i)  gsFunctionExpr.h:169
*/
#pragma once
#include <limits>

#include <gismo.h>
#include <gsb200.h>
#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

namespace gismo
{
namespace b200
{

/// Owns every buffer a gsb200_problem points into.
struct gsB200Problem
{
    gsb200_problem pb;
    std::vector<gsb200_patch> patches;
    std::deque<std::vector<double> >  dbl;   // knots, coefs, weights
    std::deque<std::vector<int32_t> > idx;   // dof maps, program opcodes
    std::vector<double> fixed;
    std::vector<gsb200_program> programs;
    std::vector<gsb200_neumann> neumann;
    std::vector<gsb200_neumann> dirichlet;    // Dirichlet sides with their data, when the values are to be L2-projected on the device

    gsB200Problem() { std::memset(&pb, 0, sizeof(pb)); pb.abi_version = GSB200_ABI_VERSION; pb.nranks = 1; }
private:
    gsB200Problem(const gsB200Problem &);
    gsB200Problem & operator=(const gsB200Problem &);
};

namespace internal
{
template <short_t d, class T>
bool flattenTensorBasis(const gsBasis<T> & b, gsb200_basis & out, gsB200Problem & st)
{
    const gsTensorBSplineBasis<d,T> * tb = dynamic_cast<const gsTensorBSplineBasis<d,T>*>(&b);
    if (!tb) return false;
    out.dim = d;
    for (short_t i = 0; i != d; ++i)
    {
        const gsKnotVector<T> & kv = tb->knots(i);
        st.dbl.push_back(std::vector<double>(kv.data(), kv.data() + kv.size()));
        out.degree[i] = kv.degree();
        out.nknots[i] = static_cast<int32_t>(kv.size());
        out.knots[i]  = st.dbl.back().data();
    }
    return true;
}

template <class T>
void flattenBasis(const gsBasis<T> & b, gsb200_basis & out, gsB200Problem & st)
{
    std::memset(&out, 0, sizeof(out));
    if (flattenTensorBasis<2,T>(b, out, st) || flattenTensorBasis<3,T>(b, out, st)) return;
    GISMO_ERROR("gsB200: only 2D/3D tensor-product B-spline bases are supported by the device path");
}

/// geometry basis + optional weights (gsTensorBSpline / gsTensorNurbs)
template <class T>
void flattenGeometry(const gsGeometry<T> & g, gsb200_patch & out, gsB200Problem & st)
{
    const gsBasis<T> & gb = g.basis();
    out.geo_weights = NULL;
    if (gb.isRational())
    {
        flattenBasis(gb.source(), out.geo, st);
        const gsMatrix<T> & w = gb.weights();
        st.dbl.push_back(std::vector<double>(w.data(), w.data() + w.size()));
        out.geo_weights = st.dbl.back().data();
    }
    else
        flattenBasis(gb, out.geo, st);
    const gsMatrix<T> & c = g.coefs();
    GISMO_ENSURE(c.cols() == out.geo.dim, "gsB200: geometry must map R^d -> R^d");
    st.dbl.push_back(std::vector<double>(c.data(), c.data() + c.size()));
    out.geo_coefs = st.dbl.back().data();
}
} // namespace internal

/// Compile the components of a gsFunctionExpr into device programs.
template <class T>
void flattenSource(const gsFunction<T> & f, index_t ncompExpected, gsB200Problem & st)
{
    const gsFunctionExpr<T> * fe = dynamic_cast<const gsFunctionExpr<T>*>(&f);
    GISMO_ENSURE(fe, "gsB200: the source term must be a gsFunctionExpr (or pass samples)");
    GISMO_ENSURE(fe->targetDim() == ncompExpected, "gsB200: source term has wrong target dimension");
    st.programs.resize(fe->targetDim());
    for (short_t c = 0; c != fe->targetDim(); ++c)
    {
        std::vector<int32_t> ops(GSB200_PROGRAM_MAX_OPS);
        std::vector<double>  cst(GSB200_PROGRAM_MAX_OPS);
        int32_t nops = 0, ncst = 0;
        if (gsb200_expr_compile(fe->expression(c).c_str(), ops.data(), (int32_t)ops.size(), &nops,
                                cst.data(), (int32_t)cst.size(), &ncst) != GSB200_OK)
            GISMO_ERROR("gsB200: cannot compile source term '" << fe->expression(c) << "': " << gsb200_last_error());
        ops.resize(nops); cst.resize(ncst);
        st.idx.push_back(ops); st.dbl.push_back(cst);
        st.programs[c].nops = nops;       st.programs[c].ops = st.idx.back().data();
        st.programs[c].nconsts = ncst;    st.programs[c].consts = st.dbl.back().data();
    }
    st.pb.rhs_kind = GSB200_RHS_PROGRAM;
    st.pb.rhs_programs = st.programs.data();
}

namespace internal
{
template <class T>
gsb200_program compileExpr(const std::string & text, gsB200Problem & st)
{
    std::vector<int32_t> ops(GSB200_PROGRAM_MAX_OPS);
    std::vector<double>  cst(GSB200_PROGRAM_MAX_OPS);
    int32_t nops = 0, ncst = 0;
    if (gsb200_expr_compile(text.c_str(), ops.data(), (int32_t)ops.size(), &nops,
                            cst.data(), (int32_t)cst.size(), &ncst) != GSB200_OK)
        GISMO_ERROR("gsB200: cannot compile '" << text << "': " << gsb200_last_error());
    ops.resize(nops); cst.resize(ncst);
    st.idx.push_back(ops); st.dbl.push_back(cst);
    gsb200_program pr;
    pr.nops = nops; pr.ops = st.idx.back().data(); pr.nconsts = ncst; pr.consts = st.dbl.back().data();
    return pr;
}
} // namespace internal

/// Neumann sides of a gsBoundaryConditions (bc.neumannSides(), gsBoundaryConditions.h:439): one entry per
/// (patch, side) with its gsFunctionExpr data: 1 component = scalar flux (gsVisitorNeumann), dim components =
/// vector dotted with the outer normal (u*g_N.tr()*nv(G), poisson2_example.cpp:153).
template <class T>
void flattenNeumann(const gsBoundaryConditions<T> & bc, short_t dim, gsB200Problem & st)
{
    st.neumann.clear();
    for (typename gsBoundaryConditions<T>::const_iterator it = bc.neumannSides().begin(); it != bc.neumannSides().end(); ++it)
    {
        const gsFunctionExpr<T> * fe = dynamic_cast<const gsFunctionExpr<T>*>(it->function().get());
        GISMO_ENSURE(fe, "gsB200: Neumann data must be a gsFunctionExpr");
        GISMO_ENSURE(fe->targetDim() == 1 || fe->targetDim() == dim, "gsB200: Neumann data must have 1 or dim components");
        gsb200_neumann nm;
        std::memset(&nm, 0, sizeof(nm));
        nm.patch = it->patch(); nm.side = it->side().index(); nm.ndata = fe->targetDim();
        for (short_t c = 0; c != fe->targetDim(); ++c) nm.data[c] = internal::compileExpr<T>(fe->expression(c), st);
        st.neumann.push_back(nm);
    }
    st.pb.nneumann = static_cast<int32_t>(st.neumann.size());
    st.pb.neumann = st.neumann.empty() ? NULL : st.neumann.data();
}

/// Dirichlet sides of a gsBoundaryConditions (bc.dirichletSides(), gsBoundaryConditions.h:433) with scalar gsFunctionExpr data
/// given in physical coordinates, for gsb200_project_dirichlet (the device-side gsDirichletValuesByL2Projection,
/// gsDirichletValues.h:257-435).  Returns false (and leaves \a st untouched) if a condition is outside that form.
template <class T>
bool flattenDirichlet(const gsBoundaryConditions<T> & bc, gsB200Problem & st, index_t ncomp = 1)
{
    std::vector<gsb200_neumann> sides;
    for (typename gsBoundaryConditions<T>::const_iterator it = bc.dirichletSides().begin(); it != bc.dirichletSides().end(); ++it)
    {
        // all components of the space at once (unkComponent -1, or 0 for a scalar space), data in physical coordinates
        if (it->unknown() != 0 || it->parametric() || ncomp > GSB200_MAX_DIM) return false;
        if (!(it->unkComponent() == -1 || (ncomp == 1 && it->unkComponent() == 0))) return false;
        gsb200_neumann sd;
        std::memset(&sd, 0, sizeof(sd));
        sd.patch = it->patch(); sd.side = it->side().index(); sd.ndata = static_cast<int32_t>(ncomp);
        const gsFunctionExpr<T> * fe = it->isHomogeneous() ? NULL : dynamic_cast<const gsFunctionExpr<T>*>(it->function().get());
        if (!it->isHomogeneous() && (!fe || fe->targetDim() != ncomp)) return false;
        for (index_t c = 0; c != ncomp; ++c)
            sd.data[c] = internal::compileExpr<T>(fe ? fe->expression(c) : std::string("0"), st);
        sides.push_back(sd);
    }
    st.dirichlet.swap(sides);
    return true;
}

/** Flatten a whole (multi-patch) discretisation.
    \param mp      geometry patches        \param mb   solution bases (one per patch)
    \param mapper  finalized DOF mapper with \a ncomp components
    \param fixed   eliminated-DOF values, boundarySize() x nrhs (may be empty = homogeneous) */
template <class T>
void flatten(const gsMultiPatch<T> & mp, const gsMultiBasis<T> & mb, const gsDofMapper & mapper,
             index_t ncomp, const gsMatrix<T> & fixed, const gsOptionList & opt, int form,
             gsB200Problem & st)
{
    GISMO_ENSURE(mp.nPatches() == mb.nBases(), "gsB200: patches/bases mismatch");
    st.patches.resize(mp.nPatches());
    for (size_t k = 0; k != mp.nPatches(); ++k)
    {
        gsb200_patch & P = st.patches[k];
        std::memset(&P, 0, sizeof(P));
        internal::flattenBasis(mb.basis(k), P.space, st);
        internal::flattenGeometry(mp.patch(k), P, st);
        GISMO_ENSURE(P.space.dim == P.geo.dim, "gsB200: basis/geometry dimension mismatch");
        const index_t sz = mb.basis(k).size();
        std::vector<int32_t> dm(static_cast<size_t>(sz) * ncomp);
        for (index_t c = 0; c != ncomp; ++c)
            for (index_t i = 0; i != sz; ++i)
                dm[static_cast<size_t>(c) * sz + i] = mapper.index(i, k, c);
        st.idx.push_back(dm);
        P.dofmap = st.idx.back().data();
    }
    gsb200_problem & pb = st.pb;
    pb.form = form;
    pb.npatches = static_cast<int32_t>(st.patches.size());
    pb.patches = st.patches.data();
    pb.ncomp = ncomp;
    pb.nfree = mapper.freeSize();
    pb.nfixed = mapper.boundarySize();
    pb.nrhs = 1;
    if (fixed.size() != 0)
    {
        GISMO_ENSURE(fixed.rows() == mapper.boundarySize(), "gsB200: fixedDofs has wrong size");
        st.fixed.assign(fixed.data(), fixed.data() + fixed.size());
        pb.fixed = st.fixed.data();
        pb.nrhs = static_cast<int32_t>(fixed.cols());
    }
    pb.quA = opt.askReal("quA", 1.0);
    pb.quB = opt.askInt("quB", 1);
    GISMO_ENSURE(opt.askInt("quRule", 1) == 1, "gsB200: only Gauss-Legendre quadrature (quRule=1) is supported");
    GISMO_ENSURE(!opt.askSwitch("overInt", false), "gsB200: overInt (boundary over-integration, quAb/quBb) is not supported");
}

/// RAII handle of a device assembler (gsb200_create ... gsb200_destroy): never leaks a device context when an
/// exception unwinds between the calls.  Also remembers the host arrays it page-locked for repeated deliveries.
struct gsB200Handle
{
    gsb200_assembler * h;
    void * pinned[2];
    gsB200Handle() : h(NULL) { pinned[0] = pinned[1] = NULL; }
    ~gsB200Handle() { reset(); }
    void unpin() { for (int k = 0; k != 2; ++k) { if (pinned[k]) gsb200_host_unpin(pinned[k]); pinned[k] = NULL; } }
    void reset() { unpin(); if (h) gsb200_destroy(h); h = NULL; }
private:
    gsB200Handle(const gsB200Handle &); gsB200Handle & operator=(const gsB200Handle &);
};

inline void check(int rc) { if (rc != GSB200_OK) GISMO_ERROR("gsB200: " << gsb200_last_error()); }

/** First assembly on a mesh: pattern + values + right-hand side straight into a gsSparseMatrix / gsMatrix.  The matrix
    is sized like Eigen's compressed form (SparseMatrix.h:150-177,626,649) and the library writes outerIndexPtr /
    innerIndexPtr / valuePtr / rhs.data() itself (pinned staging ring drained by host threads, no intermediate
    std::vector).  \a handle keeps the device context for reassembleInto(). */
template <class T>
void assembleInto(gsB200Problem & st, int device, gsB200Handle & handle, gsSparseMatrix<T> & m, gsMatrix<T> & rhs,
                  gsMatrix<T> * projected = NULL)
{
    const index_t n = st.pb.nfree;
    handle.reset();
    check(gsb200_create(&st.pb, device, &handle.h));
    if (projected && !st.dirichlet.empty())
    {   // eliminated values by L2-projection of the Dirichlet data on the device (they also become the assembly's values)
        projected->setZero(st.pb.nfixed, 1);
        int iters = 0; double res = 0;
        check(gsb200_project_dirichlet(handle.h, st.dirichlet.data(), static_cast<int>(st.dirichlet.size()), 4000, 1e-14,
                                       projected->data(), &iters, &res));
    }
    check(gsb200_build_pattern(handle.h));
    int64_t nnz = 0;
    check(gsb200_nnz(handle.h, &nnz));
    GISMO_ENSURE(nnz <= static_cast<int64_t>(std::numeric_limits<index_t>::max()),
                 "gsB200: nnz exceeds index_t; use the device view (gsb200_device_view_get)");
    rhs.setZero(n, st.pb.nrhs);
    m.resize(n, n);
    m.resizeNonZeros(static_cast<index_t>(nnz));
    check(gsb200_assemble_to_host(handle.h, m.outerIndexPtr(), m.innerIndexPtr(), m.valuePtr(), rhs.data()));
}

/// True if \a handle holds the pattern \a m was filled from (same size, still compressed).
template <class T>
bool canReassemble(const gsB200Handle & handle, const gsSparseMatrix<T> & m, const gsMatrix<T> & rhs)
{
    int64_t nnz = 0;
    return handle.h && m.isCompressed() && gsb200_nnz(handle.h, &nnz) == GSB200_OK && nnz == m.nonZeros()
        && rhs.rows() == m.rows() && m.rows() > 0;
}

/** Re-assembly on the kept device pattern (same mesh, geometry, source term; new eliminated-DOF values \a fixed, may be
    empty): values and right-hand side only are recomputed and transferred.  The value array and the right-hand side are
    page-locked once (cudaHostRegister through gsb200_host_pin; released by the handle), so that the copy engine writes
    them directly at the PCIe rate while the later chunks of the matrix are still being integrated. */
template <class T>
void reassembleInto(gsB200Handle & handle, const gsMatrix<T> & fixed, gsSparseMatrix<T> & m, gsMatrix<T> & rhs)
{
    if (handle.pinned[0] != static_cast<void*>(m.valuePtr()) || handle.pinned[1] != static_cast<void*>(rhs.data()))
    {
        handle.unpin();
        if (gsb200_host_pin(m.valuePtr(), static_cast<int64_t>(m.nonZeros()) * sizeof(T)) == GSB200_OK) handle.pinned[0] = m.valuePtr();
        if (gsb200_host_pin(rhs.data(), static_cast<int64_t>(rhs.size()) * sizeof(T)) == GSB200_OK) handle.pinned[1] = rhs.data();
        // a failed registration is not an error: the staging ring serves pageable memory
    }
    if (fixed.size() != 0) check(gsb200_set_fixed(handle.h, fixed.data()));
    check(gsb200_assemble_values_to_host(handle.h, m.valuePtr(), rhs.data()));
}

} // namespace b200
} // namespace gismo
