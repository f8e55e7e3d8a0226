/* gsb200.h — C ABI of the B200-native isogeometric system-assembly path.
 *
 * This is the drop-in boundary behind G+Smo's assembler interface
 * (gsPoissonAssembler<T>::assemble()/matrix()/rhs(), gsExprAssembler<T>::assemble(...)).
 * Everything crossing it is plain-old-data: host pointers, sizes, status codes.
 * No torch / Eigen / gismo types appear here.  All citations are file:line in the
 * reference tree (gismo/gismo v24.08.0).
 *
 * What the caller flattens (reference object -> POD field):
 *   gsTensorBSplineBasis<d>::knots(i)  (gsTensorBSplineBasis.h:192,
 *       gsKnotVector::data()/size()  gsKnotVector.h:285,242)      -> gsb200_basis.knots[i]
 *   gsGeometry::coefs()      (gsGeometry.h:343, Eigen col-major)  -> gsb200_patch.geo_coefs
 *   gsRationalBasis weights  (gsRationalBasis.h)                  -> gsb200_patch.geo_weights
 *   gsDofMapper::asVector(c) (gsDofMapper.cpp:80-87)              -> gsb200_patch.dofmap
 *   gsDofMapper::freeSize()/boundarySize() (gsDofMapper.h:436-461)-> nfree / nfixed
 *   gsAssembler::fixedDofs() / m_ddof (gsAssembler.h:276-295)     -> gsb200_problem.fixed
 *   options quA/quB (gsAssembler.hpp:30-42, gsQuadrature.h:152)   -> quA / quB
 *   gsFunctionExpr source term (gsFunctionExpr.hpp:513-533)       -> rhs program (gsb200_expr_compile)
 *
 * What comes back is exactly Eigen's compressed column-major triple of
 * gsSparseMatrix<T,0,index_t> (gsSparseMatrix.h:139; SparseMatrix.h:150-172):
 * outer[cols+1], inner[nnz] ascending per column, values[nnz]; plus the dense
 * column-major right-hand side (rows x nrhs).
 *
 * Error convention: every entry point returns 0 on success and a
 * negative GSB200_E* code otherwise; gsb200_last_error() gives the message the
 * C++ shim turns into GISMO_ERROR / std::runtime_error (gsDebug.h:89-132).
 * There is no CPU fallback: without a usable CUDA device every compute entry
 * point fails with GSB200_ENODEVICE.
 */
#ifndef GSB200_H
#define GSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB200_ABI_VERSION 2
#define GSB200_MAX_DIM 3

enum {
    GSB200_OK = 0,
    GSB200_EINVAL = -1,      /* malformed problem description              */
    GSB200_EUNSUPPORTED = -2,/* valid for the reference, outside this path */
    GSB200_ENODEVICE = -3,   /* no CUDA device / driver                    */
    GSB200_ECUDA = -4,       /* CUDA runtime error (message has details)   */
    GSB200_ENOMEM = -5,
    GSB200_ERANGE = -6,      /* nnz does not fit the 32-bit index_t of the reference */
    GSB200_ESTATE = -7       /* call order violated (e.g. download before assemble)  */
};

/* Bilinear form selector.  POISSON = gsVisitorPoisson.h:89-106 /
 * igrad(u,G)*igrad(u,G).tr()*meas(G) (poisson2_example.cpp:145);
 * ELASTICITY = linear_elasticity_example.cpp:183-190 (lambda, mu in coef[0..1]);
 * MASS = gsVisitorMass.h (u*u.tr()*meas(G)). */
enum { GSB200_FORM_POISSON = 0, GSB200_FORM_ELASTICITY = 1, GSB200_FORM_MASS = 2 };

/* Source-term description (gsFunctionExpr replacement). */
enum { GSB200_RHS_NONE = 0, GSB200_RHS_PROGRAM = 1, GSB200_RHS_SAMPLES = 2 };

/* Opcodes of the reverse-polish source-term program (stack machine over doubles).
 * GSB200_OP_CONST is followed by one extra word: the index into the literal pool. */
enum {
    GSB200_OP_CONST = 0, GSB200_OP_X = 1, GSB200_OP_Y = 2, GSB200_OP_Z = 3,
    GSB200_OP_ADD = 4, GSB200_OP_SUB = 5, GSB200_OP_MUL = 6, GSB200_OP_DIV = 7,
    GSB200_OP_POW = 8, GSB200_OP_NEG = 9, GSB200_OP_SIN = 10, GSB200_OP_COS = 11,
    GSB200_OP_TAN = 12, GSB200_OP_EXP = 13, GSB200_OP_LOG = 14, GSB200_OP_SQRT = 15,
    GSB200_OP_ABS = 16, GSB200_OP_TANH = 17, GSB200_OP_SINH = 18, GSB200_OP_COSH = 19,
    GSB200_OP_SQR = 20,   /* x*x            (emitted for x^2)        */
    GSB200_OP_SINPI = 21, /* sin(pi*x)      (emitted for sin(pi*E))  */
    GSB200_OP_COSPI = 22  /* cos(pi*x)                                 */
};
#define GSB200_PROGRAM_MAX_OPS 256
#define GSB200_PROGRAM_MAX_STACK 32

/* Tensor-product B-spline basis of one patch (gsTensorBSplineBasis<d>). */
typedef struct gsb200_basis {
    int32_t dim;                          /* parametric dimension d (2 or 3)           */
    int32_t degree[GSB200_MAX_DIM];       /* degree per direction                      */
    int32_t nknots[GSB200_MAX_DIM];       /* knots incl. repetitions per direction     */
    const double *knots[GSB200_MAX_DIM];  /* non-decreasing, open or not               */
} gsb200_basis;

/* One patch: discretisation basis + geometry map + local->global DOF map. */
typedef struct gsb200_patch {
    gsb200_basis space;        /* solution space basis (always polynomial: gsMultiBasis.hpp:37-41) */
    gsb200_basis geo;          /* the geometry's own (coarse) basis (gsGeometry.hpp:539-597)       */
    const double *geo_coefs;   /* N_geo x dim, column-major                                        */
    const double *geo_weights; /* NULL (B-spline) or N_geo NURBS weights                           */
    const int32_t *dofmap;     /* ncomp blocks of n_basis global indices: index(i,patch,comp);
                                  index >= nfree means eliminated, row (index-nfree) of `fixed`    */
} gsb200_patch;

/* Compiled source term: reverse-polish program over x,y,z (see gsb200_expr_compile). */
typedef struct gsb200_program {
    int32_t nops;
    const int32_t *ops;      /* opcode stream                         */
    int32_t nconsts;
    const double *consts;    /* literal pool                          */
} gsb200_program;

/* Neumann boundary load on one patch side (gsVisitorNeumann.h:83-136 / assembleBdr
 * gsExprAssembler.h:835-895).  side: 1 west, 2 east, 3 south, 4 north, 5 front, 6 back
 * (gsBoundary.h:58-60).  ndata == 1: rhs_i += int N_i g |n|  (visitor convention, scalar flux);
 * ndata == dim: rhs_i += int N_i (g . n)  with n the unnormalised outer normal
 * (expression u*g_N.tr()*nv(G), poisson2_example.cpp:153). */
typedef struct gsb200_neumann {
    int32_t patch, side, ndata;
    gsb200_program data[GSB200_MAX_DIM];
} gsb200_neumann;

typedef struct gsb200_problem {
    int32_t abi_version;       /* must be GSB200_ABI_VERSION                                  */
    int32_t form;              /* GSB200_FORM_*                                               */
    int32_t npatches;
    const gsb200_patch *patches;
    int32_t ncomp;             /* unknown components: 1 (Poisson/mass) or dim (elasticity)    */
    int32_t nfree;             /* gsDofMapper::freeSize()  = matrix rows = cols               */
    int32_t nfixed;            /* gsDofMapper::boundarySize()                                 */
    const double *fixed;       /* nfixed x nrhs column-major eliminated-DOF values, or NULL (=0) */
    int32_t nrhs;              /* right-hand-side columns (1 for all stock configs)           */
    double coef[4];            /* form coefficients: elasticity lambda, mu                    */
    double quA;                /* Gauss points per direction = floor(quA*p + quB + 0.5)       */
    int32_t quB;
    int32_t rhs_kind;          /* GSB200_RHS_*                                                */
    const gsb200_program *rhs_programs; /* nrhs*ncomp... one program per rhs row component:
                                  Poisson: nrhs programs; elasticity: ncomp programs (nrhs=1) */
    const double *const *rhs_samples;   /* per patch: values at all quadrature points, tensor
                                  order (direction 0 fastest), one block per component       */
    /* Partition of the work over ranks.  Rank r integrates and owns the
       matrix columns/rows of its share; nranks==1 means everything. */
    int32_t rank, nranks;
    int32_t nneumann;                   /* Neumann sides (scalar problems)                     */
    const gsb200_neumann *neumann;
} gsb200_problem;

typedef struct gsb200_assembler gsb200_assembler; /* opaque device-side state */

/* Device-resident result (64-bit offsets; required beyond int32 nnz). */
typedef struct gsb200_device_view {
    int64_t nnz;
    int32_t ncols;           /* = nfree                                       */
    int32_t col_begin, col_end; /* smallest range holding the columns this rank stores (its slab; with several
                                   patches: from its first to its last stored column, others in between are empty) */
    const int64_t *outer;    /* device, ncols+1                               */
    const int32_t *inner;    /* device, nnz                                   */
    const double *values;    /* device, nnz                                   */
    const double *rhs;       /* device, nfree x nrhs                          */
} gsb200_device_view;

/* Per-stage device times of the last gsb200_assemble(), CUDA events, milliseconds. */
typedef struct gsb200_timings {
    float geometry_ms;       /* K0: map data + coefficient tensor at quadrature points   */
    float sweep_ms[GSB200_MAX_DIM]; /* K2: one sum-factorisation sweep per direction      */
    float rhs_ms;            /* K3: load vector + elimination                            */
    float pattern_ms;        /* K1 (last gsb200_build_pattern)                           */
    float total_ms;          /* whole gsb200_assemble()                                  */
    int32_t launches;        /* kernels launched by the last gsb200_assemble()           */
    int64_t sweep_bytes[GSB200_MAX_DIM]; /* algorithmic bytes read+written per sweep     */
    int64_t sweep_flops[GSB200_MAX_DIM]; /* FP64 flops executed per sweep                */
    int32_t nchunks;
} gsb200_timings;

/* Threading contract.  A gsb200_assembler is driven by one host thread at a time (like the reference's assemblers,
   SURVEY 8b "threading"); different assemblers - e.g. one per GPU - may be driven by different threads concurrently:
   the error string, the launch planner and the launch counter are thread-local, the kernel-attribute and NVRTC caches are
   keyed per device and mutex-protected.  No entry point requires the caller to hold a CUDA context; each selects the
   assembler's device itself. */
const char *gsb200_last_error(void);
int gsb200_abi_version(void);
/* Number of visible CUDA devices (0 and GSB200_OK when none). */
int gsb200_device_count(int *count);

/* Upload the problem to `device`, build 1-D basis/quadrature tables there. */
int gsb200_create(const gsb200_problem *problem, int device, gsb200_assembler **out);
void gsb200_destroy(gsb200_assembler *a);
/* Device memory of destroyed assemblers stays in the library's private stream-ordered pool (a later assembler reuses the
   multi-GB buffers without going to the driver; the automatic workspace budget counts it as available).  gsb200_trim gives
   it back to the driver, e.g. before another library needs the device memory. */
int gsb200_trim(int device);
/* Use an existing cudaStream_t (e.g. torch's current stream); NULL = default stream. */
int gsb200_set_stream(gsb200_assembler *a, void *cuda_stream);
/* Cap for intermediate sum-factorisation storage in bytes (0 = automatic). */
int gsb200_set_workspace_limit(gsb200_assembler *a, int64_t bytes);

/* K1: sparsity pattern of the sparse system on the device
   (replaces gsSparseSystem::reserve + sorted insertion of Eigen coeffRef,
   gsSparseSystem.h:327-370, SparseMatrix.h:208-225,1366-1396). */
int gsb200_build_pattern(gsb200_assembler *a);
/* K0+K2+K3: assemble values and right-hand side on the device, asynchronously on the
   assembler's stream (replaces gsAssembler::apply<gsVisitorPoisson> + push,
   gsAssembler.h:668-722, gsVisitorPoisson.h:62-118, gsSparseSystem.h:972-1010, and the
   gsExprAssembler equivalents gsExprAssembler.h:754-833). */
int gsb200_assemble(gsb200_assembler *a);
int gsb200_synchronize(gsb200_assembler *a);

int gsb200_nnz(const gsb200_assembler *a, int64_t *nnz);
int gsb200_device_view_get(const gsb200_assembler *a, gsb200_device_view *view);
int gsb200_timings_get(const gsb200_assembler *a, gsb200_timings *t);

/* Copy the result into caller-allocated host buffers laid out like Eigen's compressed
   SparseMatrix (outer: nfree+1, inner/values: nnz) and gsMatrix rhs (nfree x nrhs). */
int gsb200_download_csc(gsb200_assembler *a, int32_t *outer, int32_t *inner, double *values);
int gsb200_download_rhs(gsb200_assembler *a, double *rhs);

/* Assemble and deliver in one pipelined step (pattern must be built): the column pointers and row
   indices travel to the host on a copy stream WHILE the values are being integrated, the values and
   the right-hand side follow as soon as the last sweep is done.  Buffers as for gsb200_download_*
   (pinned host memory gives the full PCIe rate).  This is the data path of
   gsPoissonAssemblerB200::assemble() after m_system.matrix().resizeNonZeros(nnz)
   (SparseMatrix.h:626,649; valuePtr/innerIndexPtr/outerIndexPtr :150-172). */
int gsb200_assemble_to_host(gsb200_assembler *a, int32_t *outer, int32_t *inner, double *values, double *rhs);

/* Re-assembly on a kept handle (same mesh, new coefficients / Dirichlet values / a Newton or time step): the index arrays
   were delivered once (gsb200_download_pattern or a first gsb200_assemble_to_host), only values and right-hand side travel:
   5.3 GB instead of 7.9 GB at config 2.  gsb200_set_fixed replaces the eliminated-DOF values (gsAssembler::m_ddof,
   gsAssembler.hpp:232-297) without touching the pattern. */
int gsb200_download_pattern(gsb200_assembler *a, int32_t *outer, int32_t *inner);
int gsb200_set_fixed(gsb200_assembler *a, const double *fixed);
int gsb200_assemble_values_to_host(gsb200_assembler *a, double *values, double *rhs);
/* Page-lock / release a caller buffer (cudaHostRegister) so that deliveries go straight to it at the PCIe rate.  Optional:
   pageable destinations are served through an internal pinned ring drained by host threads. */
int gsb200_host_pin(void *p, int64_t bytes);
int gsb200_host_unpin(void *p);

/* One call, host buffers in, host buffers out; stateless.  Pass outer/inner/values/rhs = NULL to query *nnz (the pattern is
   built and dropped); with buffers, *nnz on entry is checked against the problem (0 = unchecked).  Callers that assemble
   more than once keep an explicit handle instead (gsb200_create ... gsb200_destroy), as gsPoissonAssemblerB200 does. */
int gsb200_assemble_host(const gsb200_problem *problem, int device, int64_t *nnz,
                         int32_t *outer, int32_t *inner, double *values, double *rhs);

/* ---- Multi-GPU: one rank per GPU, problem.rank / problem.nranks say which share this assembler integrates.
   A single patch is cut into slabs of matrix columns along the last direction (no exchange needed for the matrix); several
   patches are distributed whole (longest first onto the least loaded rank), and the columns of DOFs shared by patches of
   different ranks - patterned identically on every rank - are summed by gsb200_exchange together with the right-hand side.
   The reference has no distributed assembly (gsExprAssembler.h:661 "mpi assemly. ???"); its transport wrappers are
   gsMpiComm.h.  The library speaks NCCL itself (loaded at run time, no link dependency):
     rank 0:      gsb200_comm_unique_id(id)  -> ship the 128 bytes with the host transport at hand (MPI_Bcast, a file, ...)
     every rank:  gsb200_comm_init(a, id)       (collective: ncclCommInitRank on the assembler's device)
   or adopts a communicator the application already has (gsb200_set_comm, ncclComm_t; not owned), or - CUDA-aware MPI, tests -
   calls back for an in-place sum of doubles over the ranks (gsb200_set_allreduce; `buf` is device memory, `stream` the
   cudaStream_t the data is ordered on). */
#define GSB200_COMM_ID_BYTES 128
typedef int (*gsb200_allreduce_fn)(void *ctx, double *buf, int64_t count, void *stream);
int gsb200_comm_unique_id(void *id128);
int gsb200_comm_init(gsb200_assembler *a, const void *id128);
int gsb200_set_comm(gsb200_assembler *a, void *nccl_comm);
int gsb200_set_allreduce(gsb200_assembler *a, gsb200_allreduce_fn fn, void *ctx);
/* K4, after gsb200_assemble: one grouped collective on the assembler's stream - an all-reduce per run of coupled columns,
   straight on the value array the final sweep wrote, plus the right-hand side (afterwards every rank holds the full rhs and
   the full sums of the coupled columns).  A no-op for nranks == 1. */
int gsb200_exchange(gsb200_assembler *a);
/* Bytes this rank contributed to the last gsb200_exchange / gsb200_cg_solve, and the number of exchanges so far. */
int gsb200_comm_stats(const gsb200_assembler *a, int64_t *bytes_last, int32_t *calls);

/* Consumer: y = A x on the device-resident matrix, and a Jacobi-
   preconditioned CG mirroring gsSparseSolver<>::CGDiagonal (gsSparseSolver.h:71-72).
   x/y/b are host pointers of length nfree. */
int gsb200_spmv_host(gsb200_assembler *a, const double *x, double *y);
/* Same product on DEVICE vectors of length nfree, asynchronously on the assembler's stream: y[c] = sum over the stored entries
   of column c (= row c, the forms are symmetric), 0 for columns this rank does not own.  With one rank per GPU the full
   product is the sum of the ranks' y (one all_reduce, gismo_b200/distributed.py::DistributedCG; SURVEY 8e "CG consumer"). */
int gsb200_spmv_device(gsb200_assembler *a, const double *x_dev, double *y_dev);
/* Diagonal of the stored columns into a DEVICE vector of length nfree (1.0 where the rank stores no diagonal entry):
   the Jacobi preconditioner of the CG consumer. */
int gsb200_diag_device(gsb200_assembler *a, double *d_dev);
int gsb200_diag_host(gsb200_assembler *a, double *d);      /* the same into a host vector (gsB200JacobiOp) */
int gsb200_cg_host(gsb200_assembler *a, const double *b, double *x, int max_iter,
                   double tol, int *iters, double *rel_residual);
/* The same solver across the ranks of the communicator, all scalars on the device (the host looks at the residual every
   `check_every` iterations).  b_host == NULL: the assembled right-hand side (column 0; exchanged, for nranks > 1);
   x_host may be NULL (gsb200_cg_solution_device).  Column slabs of one patch: every rank iterates on its own slab and
   trades only the rows its columns reach into the neighbours' slabs (ncclSend/ncclRecv), two scalar reductions per
   iteration; patch-wise ownership: full-length reduction of the product.  Every rank ends with the whole solution. */
int gsb200_cg_solve(gsb200_assembler *a, const double *b_host, double *x_host, int max_iter, double tol, int check_every,
                    int *iters, double *rel_residual);
int gsb200_cg_solution_device(gsb200_assembler *a, const double **x_dev);
/* Device time of the iteration loop of the last gsb200_cg_solve (CUDA events; set-up such as NCCL's lazy connections and the
   gathering of the solution excluded) and whether the ranks traded halos (1) or reduced full-length products (0). */
int gsb200_cg_info(const gsb200_assembler *a, double *loop_ms, int32_t *halo_exchange);
/* Columns whose row set is a translate of a reference stencil (the SpMV reads 8 instead of 12 bytes per entry there). */
int gsb200_spmv_info(gsb200_assembler *a, int64_t *regular_columns, int32_t *tables);

/* f4: integrals of a discrete scalar field u_h (free coefficients `u_free`, eliminated ones from the problem's
   `fixed` / gsb200_set_fixed) over the whole domain, with the quadrature rule of the assembly:
     out4[0] = int (u_h - u_ex)^2      out4[1] = int |grad(u_h - u_ex)|^2      out4[2] = int u_h^2      out4[3] = int |grad u_h|^2
   = ev.integral((u_ex - u_sol).sqNorm() * meas(G)), ev.integral((igrad(u_ex) - igrad(u_sol, G)).sqNorm() * meas(G)), ...
   (gsExprEvaluator.h:152-230, examples/poisson2_example.cpp:174-177).  `exact` may be NULL (out4[0] = out4[2]); `exact_grad`
   = dim programs for the components of grad u_ex, or NULL (out4[1] = -1: not available). */
int gsb200_field_norms(gsb200_assembler *a, const double *u_free, const gsb200_program *exact,
                       const gsb200_program *exact_grad, double *out4);

/* f2: eliminated-DOF values by L2-projection of the Dirichlet data onto the boundary trace space
   (gsDirichletValuesByL2Projection, gsDirichletValues.h:257-435; option DirichletValues = l2Projection (102)): boundary mass
   matrix over the eliminated functions of the listed sides, right-hand side int g N_i m (m = the measure the reference uses
   there: |det J| at the boundary points), Jacobi-CG like the reference's
   CGDiagonal - matrix-free on the device.  `sides`: (patch, side 1..2d, ndata = ncomp, data[c] = g_c) per Dirichlet side.  The result
   (nfixed values, numbering of gsDofMapper::global_to_bindex) is written to fixed_out (may be NULL) and becomes the
   assembler's eliminated values (as gsb200_set_fixed).  Scalar and vector-valued spaces (component by component), one right-hand side. */
int gsb200_project_dirichlet(gsb200_assembler *a, const gsb200_neumann *sides, int nsides, int max_iter, double tol,
                             double *fixed_out, int *iters, double *rel_residual);

/* Compile an exprtk-style source term ("2*pi^2*sin(pi*x)*sin(pi*y)") into a
   reverse-polish program.  Buffers are caller-allocated; on success *nops and *nconsts
   hold the used lengths.  Supported: + - * / ^, unary -, parentheses, x y z, pi,
   numeric literals, sin cos tan exp log sqrt abs tanh sinh cosh. */
int gsb200_expr_compile(const char *expr, int32_t *ops, int32_t ops_cap, int32_t *nops,
                        double *consts, int32_t consts_cap, int32_t *nconsts);
/* Geometry-kernel launches of the last gsb200_assemble() that ran the NVRTC-compiled source term (csrc/jit.cuh;
   policy GSB200_JIT=0|1|2: never, from the third use of a program in the process (default), at first use). */
int gsb200_jit_launches(const gsb200_assembler *a, int *count);
/* Diagnostic: translate `nprogs` source-term programs to CUDA and compile them with NVRTC together with the
   geometry kernel <dim, pgl, rational, fspec> exactly as repeated assemblies do (csrc/jit.cuh; needs libnvrtc,
   no device).  The compiler log is copied to `log` (may be NULL). */
int gsb200_jit_compile_check(const gsb200_program *progs, int nprogs, int dim, int pgl, int rational, int fspec,
                             char *log, int log_cap);
/* Host evaluation of a compiled program (used by tests to pin the device VM). */
int gsb200_expr_eval_host(const gsb200_program *prog, double x, double y, double z, double *out);

/* Measure the device's FP64 FMA and copy-bandwidth peaks (roofline denominators
   MEASURED_PEAKS.json does not carry). */
int gsb200_measure_peaks(int device, double *fp64_tflops, double *dmma_tflops, double *hbm_gbs);

#ifdef __cplusplus
}
#endif
#endif /* GSB200_H */
