/** @file gsPoissonAssemblerB200.h

    Drop-in replacement for gsPoissonAssembler<T> (gsPoissonAssembler.h:35-142) whose
    assemble() runs on a B200 through the C ABI in include/gsb200.h.  Constructors,
    refresh(), matrix(), rhs(), system(), numDofs(), fixedDofs(), constructSolution()
    are inherited unchanged; only the virtual assemble() (gsAssembler.h:415,
    gsPoissonAssembler.hpp:41-78) is overridden, so callers such as
    unittests/gsPoissonSolver_test.cpp:109-121 switch by changing the type name.

    Also provided: gsExprAssemblerB200<T>, a sibling of gsExprAssembler<T>
    (gsExprAssembler.h:29-634) exposing the same set-up surface and named entry points
    for the two bilinear forms of the BASELINE configs (the reference's
    assemble(expr...) is a variadic template, not a virtual, so it cannot be overridden).

    No CPU fallback: unsupported options raise GISMO_ERROR, as SURVEY 8(b) requires.
*/
#pragma once

#include <gismo.h>
#include <gsAssembler/gsPoissonAssembler.h>
#include "gsB200Flatten.h"
#include "gsB200LinearOperator.h"

namespace gismo
{

template <class T = real_t>
class gsPoissonAssemblerB200 : public gsPoissonAssembler<T>
{
public:
    typedef gsPoissonAssembler<T> Base;

    gsPoissonAssemblerB200(const gsPoissonPde<T> & pde, const gsMultiBasis<T> & bases)
    : Base(pde, bases), m_device(0), m_keepPattern(false), m_rank(0), m_nranks(1), m_deviceDirichlet(false) { }

    gsPoissonAssemblerB200(const gsPoissonPde<T> & pde, const gsMultiBasis<T> & bases,
                           dirichlet::strategy dirStrategy, iFace::strategy intStrategy = iFace::glue)
    : Base(pde, bases, dirStrategy, intStrategy), m_device(0), m_keepPattern(false), m_rank(0), m_nranks(1), m_deviceDirichlet(false) { }

    gsPoissonAssemblerB200(gsMultiPatch<T> const & patches, gsMultiBasis<T> const & basis,
                           gsBoundaryConditions<T> const & bconditions, const gsFunction<T> & rhs,
                           dirichlet::strategy dirStrategy = dirichlet::elimination,
                           iFace::strategy intStrategy = iFace::glue)
    : Base(patches, basis, bconditions, rhs, dirStrategy, intStrategy), m_device(0), m_keepPattern(false), m_rank(0), m_nranks(1), m_deviceDirichlet(false) { }

    void setDevice(int device) { m_device = device; }
    /// Re-use the sparsity pattern of the previous assemble() (same mesh and boundary conditions; new data): values and
    /// right-hand side only are recomputed and transferred.  refresh() resets it.
    void setKeepPattern(bool keep) { m_keepPattern = keep; }
    virtual void refresh() { Base::refresh(); m_handle.reset(); }

    /// The device-resident matrix of the last assemble() as a gsLinearOperator (and its Jacobi preconditioner) for the
    /// reference's iterative solvers; valid until the next refresh() / destruction of this assembler.
    typename gsB200LinearOperator<T>::Ptr deviceOperator() const
    { return gsB200LinearOperator<T>::make(m_handle.h, m_system.matrix().rows()); }
    typename gsB200JacobiOp<T>::Ptr devicePreconditioner() const
    { return gsB200JacobiOp<T>::make(m_handle.h, m_system.matrix().rows()); }
    /// Multi-GPU: this process integrates share \a rank of \a nranks (one process per GPU; include/gsb200.h "Multi-GPU").
    /// After assemble() the caller joins the ranks with gsb200_comm_init(deviceHandle(), id) and gsb200_exchange(deviceHandle()).
    void setRanks(int rank, int nranks) { m_rank = rank; m_nranks = nranks; }
    /// With DirichletValues = l2Projection: compute the eliminated values on the device (gsb200_project_dirichlet) instead of
    /// Base::computeDirichletDofs(); conditions outside its form (non-expression data, parametric data) keep the host code.
    void setDeviceDirichlet(bool on) { m_deviceDirichlet = on; }
    gsb200_assembler * deviceHandle() const { return m_handle.h; }

    /// Main assembly routine: same contract as gsPoissonAssembler<T>::assemble().
    virtual void assemble()
    {
        GISMO_ASSERT(m_system.initialized(),
                     "Sparse system is not initialized, call initialize() or refresh()");
        GISMO_ENSURE(m_options.getInt("DirichletStrategy") == dirichlet::elimination,
                     "gsB200: only DirichletStrategy=elimination (11) is supported");
        GISMO_ENSURE(m_options.getInt("InterfaceStrategy") == iFace::glue,
                     "gsB200: only InterfaceStrategy=glue (1) is supported");

        // Dirichlet values: the reference's host code (inputs of the path), or - L2-projection, on request - the device
        b200::gsB200Problem st;
        const bool projectOnDevice = m_deviceDirichlet && m_options.getInt("DirichletValues") == dirichlet::l2Projection
            && !(m_keepPattern && b200::canReassemble(m_handle, m_system.matrix(), m_system.rhs()))
            && b200::flattenDirichlet(m_pde_ptr->bc(), st);
        if (!projectOnDevice) Base::computeDirichletDofs();
        else
        {
            if (m_ddof.size() == 0) m_ddof.resize(m_system.numUnknowns());      // as gsAssembler::computeDirichletDofs does (gsAssembler.hpp:243-244)
            m_ddof[0].resize(0, 0);
        }

        // repeated assemble() on the same mesh (setKeepPattern): new Dirichlet values go up, values and rhs come back
        if (m_keepPattern && b200::canReassemble(m_handle, m_system.matrix(), m_system.rhs()))
        {
            b200::reassembleInto(m_handle, m_ddof[0], m_system.matrix(), m_system.rhs());
            return;
        }

        b200::flatten(m_pde_ptr->domain(), m_bases[0], m_system.colMapper(0), 1, m_ddof[0],
                      m_options, GSB200_FORM_POISSON, st);
        const gsPoissonPde<T> & ppde = static_cast<const gsPoissonPde<T>&>(*m_pde_ptr);
        st.pb.nrhs = ppde.numRhs();          // not inferred from the Dirichlet matrix (it is empty for a pure Neumann problem)
        st.pb.rank = m_rank; st.pb.nranks = m_nranks;
        b200::flattenSource(*ppde.rhs(), st.pb.nrhs, st);
        b200::flattenNeumann(m_pde_ptr->bc(), m_pde_ptr->domain().parDim(), st);   // gsVisitorNeumann on the device

        // explicit device handle: pattern + values straight into m_system (no static state, no staging vectors)
        b200::assembleInto(st, m_device, m_handle, m_system.matrix(), m_system.rhs(), projectOnDevice ? &m_ddof[0] : NULL);
    }

protected:
    using Base::m_system;
    using Base::m_options;
    using Base::m_pde_ptr;
    using Base::m_bases;
    using Base::m_ddof;
    int m_device;
    bool m_keepPattern;
    int m_rank, m_nranks;
    bool m_deviceDirichlet;
    b200::gsB200Handle m_handle;
};

/** Sibling of gsExprAssembler<T> for the forms of the BASELINE configs.  Usage mirrors
    poisson2_example.cpp:90-149 / linear_elasticity_example.cpp:108-194:

        gsExprAssemblerB200<> A;  A.setIntegrationElements(mb);
        A.setGeometry(mp);  A.setSpace(mb, dim);  A.setup(bc, dirichlet::l2Projection);
        A.assemblePoisson(f);            // igrad(u,G)*igrad(u,G).tr()*meas(G) , u*ff*meas(G)
        A.assembleElasticity(lambda, mu, f);
        A.matrix(); A.rhs();
*/
template <class T = real_t>
class gsExprAssemblerB200
{
public:
    gsExprAssemblerB200() : m_ref(1,1), m_mp(NULL), m_mb(NULL), m_bc(NULL), m_dim(1), m_device(0), m_keepPattern(false), m_lastForm(-1), m_deviceDirichlet(false), m_projBc(NULL) { }

    void setOptions(const gsOptionList & o) { m_ref.setOptions(o); }
    gsOptionList & options() { return m_ref.options(); }
    void setIntegrationElements(const gsMultiBasis<T> & mb) { m_ref.setIntegrationElements(mb); m_mb = &mb; }
    void setGeometry(const gsMultiPatch<T> & mp) { m_mp = &mp; }
    void setDevice(int device) { m_device = device; }
    void setKeepPattern(bool keep) { m_keepPattern = keep; }
    /// With dirValues = l2Projection in setup(): leave the projection of the Dirichlet data to the device (gsb200_project_dirichlet,
    /// the device-side gsDirichletValuesByL2Projection); call before setup().  gsFunctionExpr data for all components of the space.
    void setDeviceDirichlet(bool on) { m_deviceDirichlet = on; }

    /// getSpace + space::setup (gsExprAssembler.h:166, gsExpressions.h:1091): the DOF
    /// mapper and the Dirichlet values are computed by the reference's own host code.
    void setup(const gsBoundaryConditions<T> & bc, index_t dim = 1,
               dirichlet::values dirValues = dirichlet::l2Projection)
    {
        GISMO_ENSURE(m_mp && m_mb, "gsB200: set geometry and integration elements first");
        m_dim = dim;
        m_ref.getMap(*m_mp);
        typename gsExprAssembler<T>::space u = m_ref.getSpace(*m_mb, dim);
        m_projBc = NULL;
        if (m_deviceDirichlet && dirValues == dirichlet::l2Projection)
        {
            b200::gsB200Problem probe;
            if (b200::flattenDirichlet(bc, probe, dim)) m_projBc = &bc;     // the conditions fit the device path
        }
        u.setup(bc, m_projBc ? dirichlet::homogeneous : dirValues, 0);      // the mapper comes from the reference either way
        m_ref.initSystem();
        m_mapper = u.mapper();
        m_fixed = u.fixedPart();
        m_handle.reset();
    }

    index_t numDofs() const { return m_mapper.freeSize(); }
    const gsSparseMatrix<T> & matrix() const { return m_matrix; }
    const gsMatrix<T> & rhs() const { return m_rhs; }
    const gsDofMapper & mapper() const { return m_mapper; }
    const gsMatrix<T> & fixedPart() const { return m_fixed; }

    void assemblePoisson(const gsFunction<T> & f) { run(GSB200_FORM_POISSON, 0, 0, f); }
    /// Poisson + assembleBdr(bc.get("Neumann"), u * g_N.tr() * nv(G))  (poisson2_example.cpp:145-153)
    void assemblePoisson(const gsFunction<T> & f, const gsBoundaryConditions<T> & bc) { m_bc = &bc; run(GSB200_FORM_POISSON, 0, 0, f); m_bc = NULL; }
    void assembleElasticity(T lambda, T mu, const gsFunction<T> & f) { run(GSB200_FORM_ELASTICITY, lambda, mu, f); }

private:
    void run(int form, T c0, T c1, const gsFunction<T> & f)
    {
        if (m_keepPattern && form == m_lastForm && b200::canReassemble(m_handle, m_matrix, m_rhs))
        {
            b200::reassembleInto(m_handle, m_fixed, m_matrix, m_rhs);
            return;
        }
        b200::gsB200Problem st;
        gsMatrix<T> nofixed;
        const bool project = m_projBc && b200::flattenDirichlet(*m_projBc, st, m_dim);
        b200::flatten(*m_mp, *m_mb, m_mapper, m_dim, project ? nofixed : m_fixed, m_ref.options(), form, st);
        st.pb.coef[0] = c0; st.pb.coef[1] = c1;
        st.pb.nrhs = 1;
        b200::flattenSource(f, form == GSB200_FORM_ELASTICITY ? m_dim : 1, st);
        if (m_bc) b200::flattenNeumann(*m_bc, m_mp->parDim(), st);
        b200::assembleInto(st, m_device, m_handle, m_matrix, m_rhs, project ? &m_fixed : NULL);
        m_lastForm = form;
    }

    gsExprAssembler<T> m_ref;      // used for set-up only (mapper, Dirichlet values)
    const gsMultiPatch<T> * m_mp;
    const gsMultiBasis<T> * m_mb;
    const gsBoundaryConditions<T> * m_bc;
    index_t m_dim;
    int m_device;
    bool m_keepPattern;
    int m_lastForm;
    bool m_deviceDirichlet;
    const gsBoundaryConditions<T> * m_projBc;
    b200::gsB200Handle m_handle;
    gsDofMapper m_mapper;
    gsMatrix<T> m_fixed;
    gsSparseMatrix<T> m_matrix;
    gsMatrix<T> m_rhs;
};

} // namespace gismo
