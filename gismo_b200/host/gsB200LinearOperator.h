/** @file gsB200LinearOperator.h

    The device-resident system matrix as a gsLinearOperator<T> (gsLinearOperator.h:48-54), so that the reference's
    own iterative solvers (gsConjugateGradient.h:29-56, gsIterativeSolver.h:41-60) run on the matrix the B200 path
    assembled WITHOUT it ever being copied to the host: apply() ships the vector to the device, multiplies there
    (gsb200_spmv_host: regular-stencil SpMV, csrc/consumer.cuh) and ships the product back.  gsB200JacobiOp is the
    matching diagonal preconditioner (what gsSparseSolver<>::CGDiagonal uses, gsSparseSolver.h:71-72).

    The operators do not own the device handle: they are valid as long as the assembler shim that made them
    (gsPoissonAssemblerB200::deviceOperator()) is alive and has not been refreshed.
*/
#pragma once

#include <gismo.h>
#include <gsb200.h>

namespace gismo
{

template <class T = real_t>
class gsB200LinearOperator : public gsLinearOperator<T>
{
public:
    typedef memory::shared_ptr<gsB200LinearOperator> Ptr;
    typedef memory::unique_ptr<gsB200LinearOperator> uPtr;

    gsB200LinearOperator(gsb200_assembler * handle, index_t n) : m_handle(handle), m_n(n)
    { GISMO_ENSURE(handle, "gsB200LinearOperator: no device matrix (assemble first)"); }

    static Ptr make(gsb200_assembler * handle, index_t n) { return Ptr(new gsB200LinearOperator(handle, n)); }

    void apply(const gsMatrix<T> & input, gsMatrix<T> & x) const
    {
        GISMO_ASSERT(input.rows() == m_n, "gsB200LinearOperator: dimension mismatch");
        x.resize(m_n, input.cols());
        for (index_t j = 0; j != input.cols(); ++j)
            if (gsb200_spmv_host(m_handle, input.col(j).data(), x.col(j).data()) != GSB200_OK)
                GISMO_ERROR("gsB200: " << gsb200_last_error());
    }

    index_t rows() const { return m_n; }
    index_t cols() const { return m_n; }

private:
    gsb200_assembler * m_handle;
    index_t m_n;
};

/// x = D^{-1} input with the diagonal of the device matrix (fetched once).
template <class T = real_t>
class gsB200JacobiOp : public gsLinearOperator<T>
{
public:
    typedef memory::shared_ptr<gsB200JacobiOp> Ptr;

    gsB200JacobiOp(gsb200_assembler * handle, index_t n) : m_diag(n)
    {
        GISMO_ENSURE(handle, "gsB200JacobiOp: no device matrix (assemble first)");
        if (gsb200_diag_host(handle, m_diag.data()) != GSB200_OK) GISMO_ERROR("gsB200: " << gsb200_last_error());
    }
    static Ptr make(gsb200_assembler * handle, index_t n) { return Ptr(new gsB200JacobiOp(handle, n)); }

    void apply(const gsMatrix<T> & input, gsMatrix<T> & x) const
    { x = input.array().colwise() / m_diag.array(); }

    index_t rows() const { return m_diag.rows(); }
    index_t cols() const { return m_diag.rows(); }

private:
    gsVector<T> m_diag;
};

} // namespace gismo
