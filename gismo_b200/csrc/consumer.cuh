// consumer.cuh — what follows the assembly on the device (included at the end of gsb200.cu):
//   K4  exchange of the coupled (patch-interface) columns and of the right-hand side between the ranks of a multi-GPU run,
//       NCCL collectives issued straight on the final sweep's buffers (the reference has no distributed assembly:
//       gsExprAssembler.h:661 "// mpi assemly. ???"; its MPI wrappers are gsMpiComm.h);
//   SpMV on the CSC arrays read as CSR of the symmetric matrix; columns whose row set is a translate of a reference stencil
//       ("regular": interior columns of a tensor patch) read 8 B per entry instead of 12;
//   Jacobi-preconditioned CG (gsSparseSolver<>::CGDiagonal, gsSparseSolver.h:71-72; gsConjugateGradient.hpp) with all scalars on
//       the device; across ranks either a neighbour halo exchange of the search direction (contiguous column slabs) or, for
//       patch-wise ownership, a full-length reduction of the product.
// NCCL is loaded at run time (dlopen; the copy a host application already loaded - e.g. torch's - is shared), so the library has no
// link-time dependency on it and a single-GPU caller never touches it.

namespace gsb {

#ifndef GSB200_EMULATE
// ------------------------------------------------------------------ NCCL, loaded on demand
struct NcclApi {
    void *handle = 0; bool tried = false;
    int (*GetUniqueId)(void *) = 0;
    int (*CommInitRank)(void **, int, ncclUniqueIdPod, int) = 0;
    int (*CommDestroy)(void *) = 0;
    int (*CommCount)(void *, int *) = 0;
    int (*CommUserRank)(void *, int *) = 0;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = 0;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = 0;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = 0;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = 0;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = 0;
    int (*GroupStart)() = 0;
    int (*GroupEnd)() = 0;
    const char *(*GetErrorString)(int) = 0;
};
enum { NCCL_INT32 = 2, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };     // nccl.h: ncclInt32, ncclFloat64, ncclSum
static NcclApi &nccl_api()
{
    static NcclApi api; static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (api.tried) return api;
    api.tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (api.handle) break; }   // already in the process?
    if (!api.handle) if (const char *e = getenv("GSB200_NCCL_LIB")) api.handle = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    for (const char *n : names) { if (api.handle) break; api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
    if (!api.handle) return api;
#define GSB_NCCL_SYM(field, name) *(void **)(&api.field) = dlsym(api.handle, name)
    GSB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); GSB_NCCL_SYM(CommInitRank, "ncclCommInitRank"); GSB_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    GSB_NCCL_SYM(CommCount, "ncclCommCount"); GSB_NCCL_SYM(CommUserRank, "ncclCommUserRank");
    GSB_NCCL_SYM(AllReduce, "ncclAllReduce"); GSB_NCCL_SYM(AllGather, "ncclAllGather"); GSB_NCCL_SYM(Broadcast, "ncclBroadcast");
    GSB_NCCL_SYM(Send, "ncclSend"); GSB_NCCL_SYM(Recv, "ncclRecv"); GSB_NCCL_SYM(GroupStart, "ncclGroupStart"); GSB_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    GSB_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef GSB_NCCL_SYM
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd || !api.Broadcast || !api.AllGather) { api.handle = 0; }
    return api;
}
static int nccl_check(int rc, const char *what)
{
    if (rc == 0) return 0;
    NcclApi &N = nccl_api();
    set_error("%s: NCCL error %d (%s)", what, rc, N.GetErrorString ? N.GetErrorString(rc) : "?");
    return GSB200_ECUDA;
}
void destroy_comm(void *comm) { NcclApi &N = nccl_api(); if (N.handle && N.CommDestroy && comm) N.CommDestroy(comm); }
static int need_nccl(NcclApi **out)
{
    NcclApi &N = nccl_api();
    if (!N.handle) { set_error("libnccl.so.2 could not be loaded (set GSB200_NCCL_LIB): multi-GPU exchange needs NCCL, there is no fallback"); return GSB200_EUNSUPPORTED; }
    *out = &N; return 0;
}
#endif

// in-place sum of `count` doubles over the ranks, ordered on the assembler's stream
static int comm_allreduce(gsb200_assembler *a, double *buf, i64 count)
{
    if (a->nranks == 1 || count <= 0) return 0;
    if (a->ar_fn) {
        const int rc = a->ar_fn(a->ar_ctx, buf, count, (void *)(size_t)a->stream);
        if (rc) { set_error("all-reduce callback failed with %d", rc); return GSB200_ECUDA; }
        return 0;
    }
#ifndef GSB200_EMULATE
    if (a->comm) { NcclApi *N; GSB_TRY(need_nccl(&N)); return nccl_check(N->AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, a->comm, a->stream), "ncclAllReduce"); }
#endif
    set_error("nranks = %d but no communicator: call gsb200_comm_init / gsb200_set_comm / gsb200_set_allreduce first", a->nranks);
    return GSB200_ESTATE;
}
static int comm_group(gsb200_assembler *a, bool begin)
{
#ifndef GSB200_EMULATE
    if (a->comm && !a->ar_fn) { NcclApi *N; GSB_TRY(need_nccl(&N)); return nccl_check(begin ? N->GroupStart() : N->GroupEnd(), begin ? "ncclGroupStart" : "ncclGroupEnd"); }
#endif
    (void)a; (void)begin; return 0;
}

// ------------------------------------------------------------------ kernels
// column c is "regular" with table t if it has the table's length and inner[k] - c equals the table's offsets
GSB_GLOBAL void k_spmv_classify(int n, const i64 *ptr, const int *idx, int ntab, const int *tlen, const int *toff, int tstride, unsigned char *reg)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const i64 b = ptr[c]; const int len = (int)(ptr[c + 1] - b);
    unsigned char r = 0;
    for (int t = 0; t < ntab && !r; ++t) {
        if (tlen[t] != len || len == 0) continue;
        const int *o = toff + (i64)t * tstride;
        bool same = idx[b] - c == o[0] && idx[b + len - 1] - c == o[len - 1];
        for (int k = 1; same && k + 1 < len; ++k) same = idx[b + k] - c == o[k];
        if (same) r = (unsigned char)(t + 1);
    }
    reg[c] = r;
}
#ifndef GSB200_EMULATE
// One warp per stored column (= row, the forms are symmetric).  Regular columns take their row indices from the offset table
// (L1-resident), so only the values stream from HBM: 8 B per entry; the others read 12 B.  x is gathered through L1/L2.
// dot != 0: adds sum_c x[c] * y[c] over the processed columns (CG's p.Ap) with one atomic per warp.
GSB_GLOBAL void __launch_bounds__(256) k_spmv_reg(int c0, int c1, const i64 *ptr, const int *idx, const double *val, const unsigned char *reg,
                                                  const int *tlen, const int *toff, int tstride, const double *x, double *y, const double *scale, double *dot)
{
    const int lane = threadIdx.x & 31;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((i64)gridDim.x * blockDim.x) >> 5;
    double dsum = 0.0;
    for (i64 r = c0 + warp0; r < c1; r += nwarp) {
        const i64 b = ptr[r];
        const int t = reg[r];
        double s = 0.0;
        // tiles of 8 x 32 entries: all value (and index) loads of a tile are issued before the first gather of x depends on them
        if (t) {
            const int len = tlen[t - 1]; const int *o = toff + (i64)(t - 1) * tstride;
            const double *v = val + b, *xr = x + r;
            for (int base = 0; base < len; base += 256) {
                double vv[8]; int oo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { const int k = base + j * 32 + lane; const bool ok = k < len; vv[j] = ok ? __ldcs(v + k) : 0.0; oo[j] = ok ? __ldg(o + k) : 0; }
#pragma unroll
                for (int j = 0; j < 8; ++j) s = fma(vv[j], __ldg(xr + oo[j]), s);
            }
        } else {
            const int len = (int)(ptr[r + 1] - b);
            const double *v = val + b; const int *ix = idx + b;
            for (int base = 0; base < len; base += 256) {
                double vv[8]; int oo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { const int k = base + j * 32 + lane; const bool ok = k < len; vv[j] = ok ? __ldcs(v + k) : 0.0; oo[j] = ok ? __ldcs(ix + k) : (int)r; }
#pragma unroll
                for (int j = 0; j < 8; ++j) s = fma(vv[j], __ldg(x + oo[j]), s);
            }
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
        if (lane == 0) {
            if (scale) s *= scale[r];
            y[r] = s;
            if (dot) dsum = fma(x[r], s, dsum);
        }
    }
    if (dot && lane == 0 && dsum != 0.0) atomicAdd(dot, dsum);
}
#endif
GSB_GLOBAL void k_spmv_range(int c0, int c1, const i64 *ptr, const int *idx, const double *val, const double *x, double *y, const double *scale, double *dot)
{
    const int r = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= c1) return;
    double s = 0.0;
    for (i64 k = ptr[r]; k < ptr[r + 1]; ++k) s = fma(val[k], x[idx[k]], s);
    if (scale) s *= scale[r];
    y[r] = s;
    if (dot) atomic_add(dot, x[r] * s);
}
// stored[c] = 1.0 if the rank stores column c; diag[c] = its diagonal entry (0 where not stored)
GSB_GLOBAL void k_cg_diag(int n, const i64 *ptr, const int *idx, const double *val, double *diag, double *stored)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double s = 0.0;
    for (i64 k = ptr[r]; k < ptr[r + 1]; ++k) if (idx[k] == r) s = val[k];
    diag[r] = s; stored[r] = ptr[r + 1] > ptr[r] ? 1.0 : 0.0;
}
// d = holders > 0 ? diag / holders : 1 (a column nobody stores: an isolated DOF), 1 where the diagonal vanishes; h = 1 / max(holders, 1)
GSB_GLOBAL void k_cg_prep(int n, double *diag, double *holders)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double h = holders[i] > 0.0 ? holders[i] : 1.0;
    const double d = diag[i] / h;
    diag[i] = d == 0.0 ? 1.0 : d; holders[i] = 1.0 / h;
}
GSB_DEVICE void cg_block_add(double s, double *out)
{
#ifndef GSB200_EMULATE
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(out, s);
#else
    if (s != 0.0) *out += s;
#endif
}
// scalars S (device): [0] rz  [1] pq  [2] rz_new  [3] rr  [4] bb
// r = b, z = r / d, p = z on [c0, c1); rz += r.z, bb += b.b
// (the dot products take the entries [d0, d1) only: ranks that keep whole vectors split the index range between them, so that the
// reduced scalars - and with them every rank's iterates - are bitwise the same everywhere).  Fixed-size grids, grid-stride loops:
// one atomic per warp and scalar.
GSB_GLOBAL void k_cg_start(int c0, int c1, int d0, int d1, const double *b, const double *d, double *x, double *r, double *z, double *p, double *S)
{
    double s0 = 0.0, s1 = 0.0;
    for (i64 i = c0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < c1; i += (i64)gridDim.x * blockDim.x) {
        const double bi = b[i]; const double zi = bi / d[i];
        x[i] = 0.0; r[i] = bi; z[i] = zi; p[i] = zi;
        if (i >= d0 && i < d1) { s0 = fma(bi, zi, s0); s1 = fma(bi, bi, s1); }
    }
    cg_block_add(s0, S + 0); cg_block_add(s1, S + 4);
}
GSB_GLOBAL void k_cg_dot(int c0, int c1, const double *u, const double *v, double *out)
{
    double s0 = 0.0;
    for (i64 i = c0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < c1; i += (i64)gridDim.x * blockDim.x) s0 = fma(u[i], v[i], s0);
    cg_block_add(s0, out);
}
// alpha = rz / pq;  x += alpha p;  r -= alpha q;  z = r / d;  rz_new += r.z;  rr += r.r
GSB_GLOBAL void k_cg_step1(int c0, int c1, int d0, int d1, const double *p, const double *q, const double *d, double *x, double *r, double *z, double *S)
{
    const double alpha = S[0] / S[1];
    double s0 = 0.0, s1 = 0.0;
    for (i64 i = c0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < c1; i += (i64)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, q[i], r[i]); const double zi = ri / d[i];
        r[i] = ri; z[i] = zi;
        if (i >= d0 && i < d1) { s0 = fma(ri, zi, s0); s1 = fma(ri, ri, s1); }
    }
    cg_block_add(s0, S + 2); cg_block_add(s1, S + 3);
}
// beta = rz_new / rz;  p = z + beta p
GSB_GLOBAL void k_cg_step2(int c0, int c1, const double *z, double *p, const double *S)
{
    const double beta = S[2] / S[0];
    for (i64 i = c0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < c1; i += (i64)gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}
// rotate the scalars for the next iteration: rz = rz_new, last_rr = rr; accumulators cleared
GSB_GLOBAL void k_cg_rotate(double *S)
{
    if (blockIdx.x || threadIdx.x) return;
    S[0] = S[2]; S[5] = S[3]; S[1] = 0.0; S[2] = 0.0; S[3] = 0.0;
}
// first / last stored column and smallest / largest row index they reference: out = {c_first, c_last + 1, r_min, r_max + 1, stored count}
GSB_GLOBAL void k_col_extent(int n, const i64 *ptr, const int *idx, int *out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n || ptr[c + 1] == ptr[c]) return;
#ifndef GSB200_EMULATE
    atomicMin(out + 0, c); atomicMax(out + 1, c + 1); atomicMin(out + 2, idx[ptr[c]]); atomicMax(out + 3, idx[ptr[c + 1] - 1] + 1); atomicAdd(out + 4, 1);
#else
    out[0] = std::min(out[0], c); out[1] = std::max(out[1], c + 1); out[2] = std::min(out[2], idx[ptr[c]]); out[3] = std::max(out[3], idx[ptr[c + 1] - 1] + 1); out[4] += 1;
#endif
}

// ------------------------------------------------------------------ f4: norms of a discrete field (gsExprEvaluator::integral)
// One thread per quadrature point of a patch: geometry (point, Jacobian) by the tensor sum over the geometry's own basis, the discrete
// field and its parametric gradient over the (p+1)^d active functions (free coefficients from u, eliminated ones from `fixed`),
// physical gradient through the inverse Jacobian; out[0..3] += w |det J| { (u_h - u_ex)^2, |grad(u_h - u_ex)|^2, u_h^2, |grad u_h|^2 }.
// Replaces ev.integral((u_ex - u_sol).sqNorm() * meas(G)) and ev.integral((igrad(u_ex) - igrad(u_sol, G)).sqNorm() * meas(G))
// (examples/poisson2_example.cpp:174-177, gsExprEvaluator.h:152-230, 461-518).
struct NormArgs {
    int qn[3];
    const double2 *gtab[3]; const int *gfirst[3]; int pg1[3], ngeo[3];
    const double *hpt[3]; const double *gwp[3];
    const double *coefs; const double *weights; i64 ngeo_total;
    const double2 *tabl[3]; const int *first[3]; int p1[3], q[3], nfun[3];
    const int *dofmap; const double *u; const double *fixed; int nfree;
    int has_exact, has_grad; DevProgram ex, exg[3];
    double *out;
};
template <int DIM>
GSB_GLOBAL void k_field_norms(const NormArgs A)
{
    i64 total = 1; for (int k = 0; k < DIM; ++k) total *= A.qn[k];
    const i64 id = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (id < total) {
        int ql[3] = {0, 0, 0}; { i64 r = id; for (int k = 0; k < DIM; ++k) { ql[k] = (int)(r % A.qn[k]); r /= A.qn[k]; } }
        // ---- geometry
        double W = 0.0, dW[3] = {0, 0, 0}, xn[3] = {0, 0, 0}, dxn[3][3];
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) dxn[a][c] = 0.0;
        int cnt[3] = {0, 0, 0}, gf[3] = {0, 0, 0};
        for (int k = 0; k < DIM; ++k) gf[k] = A.gfirst[k][ql[k]];
        for (;;) {
            i64 idx = 0; double v = 1.0, dv[3] = {0, 0, 0}; double2 b[3];
            for (int k = DIM - 1; k >= 0; --k) idx = idx * A.ngeo[k] + (gf[k] + cnt[k]);
            for (int k = 0; k < DIM; ++k) { b[k] = A.gtab[k][(i64)ql[k] * A.pg1[k] + cnt[k]]; v *= b[k].x; }
            for (int k = 0; k < DIM; ++k) { dv[k] = b[k].y; for (int i = 0; i < DIM; ++i) if (i != k) dv[k] *= b[i].x; }
            const double wt = A.weights ? A.weights[idx] : 1.0;
            W += wt * v;
            for (int k = 0; k < DIM; ++k) dW[k] += wt * dv[k];
            for (int c = 0; c < DIM; ++c) {
                const double C = A.coefs[(i64)c * A.ngeo_total + idx];
                xn[c] += wt * v * C;
                for (int k = 0; k < DIM; ++k) dxn[k][c] += wt * dv[k] * C;
            }
            int k = 0;
            while (k < DIM && ++cnt[k] >= A.pg1[k]) { cnt[k] = 0; ++k; }
            if (k == DIM) break;
        }
        double x[3] = {0, 0, 0}, Jt[3][3];           // Jt[a][c] = d x_c / d xi_a
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) Jt[a][c] = a == c ? 1.0 : 0.0;
        for (int c = 0; c < DIM; ++c) { x[c] = xn[c] / W; for (int a = 0; a < DIM; ++a) Jt[a][c] = (dxn[a][c] * W - xn[c] * dW[a]) / (W * W); }
        // ---- discrete field and its parametric gradient
        double uh = 0.0, du[3] = {0, 0, 0};
        int e[3] = {0, 0, 0}, t[3] = {0, 0, 0}, f0[3] = {0, 0, 0};
        for (int k = 0; k < DIM; ++k) { e[k] = ql[k] / A.q[k]; t[k] = ql[k] - e[k] * A.q[k]; f0[k] = A.first[k][e[k]]; }
        for (int k = 0; k < 3; ++k) cnt[k] = 0;
        for (;;) {
            i64 li = 0; double v = 1.0, dv[3] = {0, 0, 0}; double2 b[3];
            for (int k = DIM - 1; k >= 0; --k) li = li * A.nfun[k] + (f0[k] + cnt[k]);
            for (int k = 0; k < DIM; ++k) { b[k] = A.tabl[k][(i64)ql[k] * A.p1[k] + cnt[k]]; v *= b[k].x; }
            for (int k = 0; k < DIM; ++k) { dv[k] = b[k].y; for (int i = 0; i < DIM; ++i) if (i != k) dv[k] *= b[i].x; }
            const int g = A.dofmap[li];
            const double cf = g < A.nfree ? A.u[g] : (A.fixed ? A.fixed[g - A.nfree] : 0.0);
            uh = fma(cf, v, uh);
            for (int k = 0; k < DIM; ++k) du[k] = fma(cf, dv[k], du[k]);
            int k = 0;
            while (k < DIM && ++cnt[k] >= A.p1[k]) { cnt[k] = 0; ++k; }
            if (k == DIM) break;
        }
        // ---- physical gradient: Jt g = du  (cofactor inverse)
        double det, gr[3] = {0, 0, 0};
        if (DIM == 2) {
            det = Jt[0][0] * Jt[1][1] - Jt[0][1] * Jt[1][0];
            gr[0] = (Jt[1][1] * du[0] - Jt[0][1] * du[1]) / det;
            gr[1] = (-Jt[1][0] * du[0] + Jt[0][0] * du[1]) / det;
        } else {
            const double c00 = Jt[1][1] * Jt[2][2] - Jt[1][2] * Jt[2][1], c01 = Jt[1][2] * Jt[2][0] - Jt[1][0] * Jt[2][2], c02 = Jt[1][0] * Jt[2][1] - Jt[1][1] * Jt[2][0];
            det = Jt[0][0] * c00 + Jt[0][1] * c01 + Jt[0][2] * c02;
            const double c10 = Jt[0][2] * Jt[2][1] - Jt[0][1] * Jt[2][2], c11 = Jt[0][0] * Jt[2][2] - Jt[0][2] * Jt[2][0], c12 = Jt[0][1] * Jt[2][0] - Jt[0][0] * Jt[2][1];
            const double c20 = Jt[0][1] * Jt[1][2] - Jt[0][2] * Jt[1][1], c21 = Jt[0][2] * Jt[1][0] - Jt[0][0] * Jt[1][2], c22 = Jt[0][0] * Jt[1][1] - Jt[0][1] * Jt[1][0];
            // inverse(Jt)[r][c] = cofactor[c][r] / det
            gr[0] = (c00 * du[0] + c10 * du[1] + c20 * du[2]) / det;
            gr[1] = (c01 * du[0] + c11 * du[1] + c21 * du[2]) / det;
            gr[2] = (c02 * du[0] + c12 * du[1] + c22 * du[2]) / det;
        }
        double w = fabs(det);
        for (int k = 0; k < DIM; ++k) w *= A.hpt[k][ql[k]] * A.gwp[k][ql[k]];
        const double ue = A.has_exact ? program_eval(A.ex, x[0], x[1], x[2]) : 0.0;
        double ge2 = 0.0, g2 = 0.0;
        for (int c = 0; c < DIM; ++c) {
            const double gx = A.has_grad ? program_eval(A.exg[c], x[0], x[1], x[2]) : 0.0;
            ge2 = fma(gr[c] - gx, gr[c] - gx, ge2); g2 = fma(gr[c], gr[c], g2);
        }
        s0 = w * (uh - ue) * (uh - ue); s1 = w * ge2; s2 = w * uh * uh; s3 = w * g2;
    }
    cg_block_add(s0, A.out + 0); cg_block_add(s1, A.out + 1); cg_block_add(s2, A.out + 2); cg_block_add(s3, A.out + 3);
}

// ------------------------------------------------------------------ f2: Dirichlet values by L2-projection onto the boundary trace space
// gsDirichletValuesByL2Projection (gsDirichletValues.h:257-435; visitor path: gsAssembler.hpp computeDirichletDofsL2Proj): boundary mass
// matrix M_ij = sum over Dirichlet sides int N_i N_j m over the ELIMINATED functions, right-hand side int g N_i m (m = |det J|, see below), solved with
// Jacobi-CG (the reference: gsSparseSolver<>::CGDiagonal).  Here M is never formed: M x = sum_sides B^T (w|n| .* (B x)) with the face
// kernels (k_face_eval, k_face_load); the weights w m and the data w m g come from k_face_geometry.
struct ProjSide { FaceArgs F; FaceLoadArgs L; FaceEvalArgs E; i64 npt, nfn; double *Wb, *Gb, *Tb; };

static void proj_side_setup(gsb200_assembler *a, int patch, int side, int comp, ProjSide &S)
{
    const PatchDev &P = a->patches[patch];
    const int dim = a->dim, dir = (side - 1) / 2, upper = (side - 1) % 2;
    memset(&S.F, 0, sizeof S.F); memset(&S.L, 0, sizeof S.L); memset(&S.E, 0, sizeof S.E);
    S.F.dim = dim; S.F.dir = dir; S.F.upper = upper; S.L.dim = dim; S.L.dir = dir; S.E.dim = dim; S.E.dir = dir;
    S.npt = 1; S.nfn = 1;
    for (int k = 0; k < dim; ++k) {
        const Dir1D &d = P.dir[k];
        S.F.qn[k] = d.Q; S.F.gtab[k] = d.d_gtab; S.F.gfirst[k] = d.d_gfirst; S.F.pg1[k] = d.pg1; S.F.ngeo[k] = d.ngeo; S.F.hpt[k] = d.d_hpt; S.F.gwp[k] = d.d_gwp;
        S.L.nfun[k] = d.nfun; S.L.p1[k] = d.p + 1; S.L.q[k] = d.q; S.L.Q[k] = d.Q; S.L.ffirst[k] = d.d_ffirst; S.L.flast[k] = d.d_flast; S.L.tab[k] = d.d_tab;
        S.E.qn[k] = d.Q; S.E.nfun[k] = d.nfun; S.E.p1[k] = d.p + 1; S.E.q[k] = d.q; S.E.tabl[k] = d.d_tabl; S.E.first[k] = d.d_first;
        if (k != dir) { S.npt *= d.Q; S.nfn *= d.nfun; }
    }
    const Dir1D &dd = P.dir[dir];
    for (int k2 = 0; k2 < dd.pg1; ++k2) S.F.bgeo[k2] = dd.bgeo[upper][k2];
    S.F.bgfirst = dd.bgfirst[upper];
    for (int k2 = 0; k2 <= dd.p; ++k2) { S.L.bval[k2] = dd.bval[upper][k2]; S.E.bval[k2] = dd.bval[upper][k2]; }
    S.L.bfirst = dd.bfirst[upper]; S.L.nb1 = dd.p + 1; S.E.bfirst = S.L.bfirst; S.E.nb1 = S.L.nb1;
    S.F.coefs = P.d_coefs; S.F.weights = P.d_weights; S.F.ngeo_total = P.ngeo_total;
    S.L.dofmap = P.d_dofmap + (i64)comp * P.nb; S.L.nfree = a->nfree; S.L.to_fixed = 1;      // component-major blocks of the DOF map
    S.E.dofmap = P.d_dofmap + (i64)comp * P.nb; S.E.nfree = a->nfree;
}
template <class K2, class K3, class ARGS>
static void face_launch(int dim, K2 k2, K3 k3, i64 n, stream_t s, const ARGS &A)
{
    if (dim == 2) GSB_LAUNCH(k2, dim3((unsigned)((n + 127) / 128)), dim3(128), s, A);
    else GSB_LAUNCH(k3, dim3((unsigned)((n + 127) / 128)), dim3(128), s, A);
}

static int project_dirichlet(gsb200_assembler *a, const gsb200_neumann *sides, int nsides, int max_iter, double tol, double *fixed_out, int *iters, double *relres)
{
    const int nb = a->nfixed; stream_t s = a->stream;
    if (a->nrhs != 1) { set_error("Dirichlet L2-projection: one right-hand side"); return GSB200_EUNSUPPORTED; }
    if (nb == 0) { if (iters) *iters = 0; if (relres) *relres = 0.0; return 0; }
    // one entry per (side, component): a vector-valued space is projected component by component (the boundary mass matrix is block
    // diagonal), the data of a side has one program per component
    const int nsides_in = nsides;
    for (int i = 0; i < nsides_in; ++i)
        if (sides[i].ndata != a->ncomp) { set_error("Dirichlet side %d carries %d data components, the space has %d", i, sides[i].ndata, a->ncomp); return GSB200_EINVAL; }
    nsides = nsides_in * a->ncomp;
    std::vector<ProjSide> S((size_t)nsides);
    std::vector<void *> bufs;
    auto alloc = [&](size_t n) -> double * { void *p = 0; if (dev_malloc(&p, sizeof(double) * std::max<size_t>(n, 1))) return (double *)0; bufs.push_back(p); return (double *)p; };
    const size_t nprog0 = a->prog_bufs.size();
    int rc = 0;
    gsb200_program one; int one_ops[2] = {GSB200_OP_CONST, 0}; double one_c[1] = {1.0};
    one.nops = 2; one.ops = one_ops; one.nconsts = 1; one.consts = one_c;
    double *vec[7];       // x r z p q d + scalars
    for (int k = 0; k < 7; ++k) { vec[k] = alloc((size_t)nb + 8); if (!vec[k]) rc = GSB200_ENOMEM; }
    double *X = vec[0], *R = vec[1], *Z = vec[2], *Pv = vec[3], *Q = vec[4], *Dg = vec[5], *Sc = vec[6];
    for (int i = 0; i < nsides && !rc; ++i) {
        const gsb200_neumann &sd = sides[i / a->ncomp];
        const int comp = i % a->ncomp;
        if (sd.patch < 0 || sd.patch >= (int)a->patches.size() || sd.side < 1 || sd.side > 2 * a->dim) { set_error("Dirichlet side %d malformed", i / a->ncomp); rc = GSB200_EINVAL; break; }
        ProjSide &P = S[i];
        proj_side_setup(a, sd.patch, sd.side, comp, P);
        P.Wb = alloc((size_t)P.npt); P.Gb = alloc((size_t)P.npt); P.Tb = alloc((size_t)P.npt);
        if (!P.Wb || !P.Gb || !P.Tb) { rc = GSB200_ENOMEM; break; }
        // the reference integrates with md.measure of a gsMapData whose side is not set (gsDirichletValues.h:273 / gsAssembler.hpp:412):
        // the VOLUME measure |det J| at the boundary points, not the surface measure; restated as it is
        P.F.ndata = 1; P.F.vol_measure = 1;
        if ((rc = upload_device_program(a, one, &P.F.prog[0]))) break;
        P.F.Fb = P.Wb; face_launch(a->dim, k_face_geometry<2>, k_face_geometry<3>, P.npt, s, P.F);          // w |n|
        if ((rc = upload_device_program(a, sd.data[comp], &P.F.prog[0]))) break;
        P.F.Fb = P.Gb; face_launch(a->dim, k_face_geometry<2>, k_face_geometry<3>, P.npt, s, P.F);          // w |n| g
    }
    auto apply = [&](const double *x, double *y, bool diag) {      // y = M x  (diag: y = diag M)
        dev_memset(y, 0, sizeof(double) * (size_t)nb, s);
        for (int i = 0; i < nsides; ++i) {
            ProjSide &P = S[i];
            FaceLoadArgs L = P.L; L.rhs = y;
            if (diag) { L.Fb = P.Wb; L.square = 1; }
            else {
                FaceEvalArgs E = P.E; E.x = x; E.Wb = P.Wb; E.out = P.Tb;
                face_launch(a->dim, k_face_eval<2>, k_face_eval<3>, P.npt, s, E);
                L.Fb = P.Tb;
            }
            face_launch(a->dim, k_face_load<2>, k_face_load<3>, P.nfn, s, L);
        }
    };
    int it = 0; double rr = 0.0, bb = 0.0;
    if (!rc) {
#ifndef GSB200_EMULATE
        const dim3 gO(148 * 2), tv(256);
#else
        const dim3 gO(1), tv(1);
#endif
        // right-hand side b_i = int g N_i |n| into R; diagonal into Dg (an eliminated DOF no Dirichlet side touches keeps value 0)
        dev_memset(R, 0, sizeof(double) * (size_t)nb, s);
        for (int i = 0; i < nsides; ++i) { FaceLoadArgs L = S[i].L; L.rhs = R; L.Fb = S[i].Gb; face_launch(a->dim, k_face_load<2>, k_face_load<3>, S[i].nfn, s, L); }
        apply(0, Dg, true);
        dev_memset(Q, 0, sizeof(double) * (size_t)nb, s);
        GSB_LAUNCH(k_cg_prep, dim3((nb + 127) / 128), dim3(128), s, nb, Dg, Q);      // zero diagonal -> 1 (Q is scratch for the holders slot)
        rc = dev_d2d(Q, R, sizeof(double) * (size_t)nb, s);                           // b
        if (!rc) rc = dev_memset(Sc, 0, 8 * sizeof(double), s);
        if (!rc) {
            GSB_LAUNCH(k_cg_start, gO, tv, s, 0, nb, 0, nb, Q, Dg, X, R, Z, Pv, Sc);
            double hs[8]; rc = dev_d2h(hs, Sc, sizeof hs, s);
            bb = hs[4]; rr = bb;
            const double thr = tol * tol * bb;
            while (!rc && it < max_iter && rr > thr) {
                apply(Pv, Q, false);
                GSB_LAUNCH(k_cg_dot, gO, tv, s, 0, nb, Pv, Q, Sc + 1);
                GSB_LAUNCH(k_cg_step1, gO, tv, s, 0, nb, 0, nb, Pv, Q, Dg, X, R, Z, Sc);
                GSB_LAUNCH(k_cg_step2, gO, tv, s, 0, nb, Z, Pv, Sc);
                GSB_LAUNCH(k_cg_rotate, dim3(1), dim3(32), s, Sc);
                ++it;
                if (it % 5 == 0 || it == max_iter) { rc = dev_d2h(hs, Sc, sizeof hs, s); rr = hs[5]; }
            }
            if (!rc) { rc = dev_d2h(hs, Sc, sizeof hs, s); if (it) rr = hs[5]; }
        }
        if (!rc && fixed_out) rc = dev_d2h(fixed_out, X, sizeof(double) * (size_t)nb, s);
        if (!rc) {      // the projected values become the assembler's eliminated-DOF values
            if (!a->d_fixed) rc = dev_malloc((void **)&a->d_fixed, sizeof(double) * (size_t)nb);
            if (!rc) rc = dev_d2d(a->d_fixed, X, sizeof(double) * (size_t)nb, s);
        }
    }
    if (!rc) rc = dev_sync(s); else dev_sync(s);
    for (void *p : bufs) dev_free(p);
    while (a->prog_bufs.size() > nprog0) { dev_free(a->prog_bufs.back()); a->prog_bufs.pop_back(); }
    if (iters) *iters = it;
    if (relres) *relres = bb > 0 ? sqrt(rr / bb) : 0.0;
    return rc;
}

// ------------------------------------------------------------------ SpMV set-up: offset tables of the regular columns
static int spmv_prepare(gsb200_assembler *a)
{
    if (a->spmv_ready) return 0;
    const int N = a->nfree; stream_t s = a->stream;
    // candidate reference columns: the middle function of every (patch, component)
    std::vector<int> cand;
    for (size_t ip = 0; ip < a->patches.size(); ++ip) for (int c = 0; c < a->ncomp; ++c) if (a->mid_dof[ip * a->ncomp + c] >= 0) cand.push_back(a->mid_dof[ip * a->ncomp + c]);
    std::vector<std::vector<int>> tabs;
    for (int g : cand) {
        if ((int)tabs.size() >= 32) break;
        i64 be[2]; GSB_TRY(dev_d2h(be, a->d_colptr + g, 2 * sizeof(i64), s));
        const int len = (int)(be[1] - be[0]);
        if (len <= 0 || len > 4096) continue;
        std::vector<int> o((size_t)len); GSB_TRY(dev_d2h(o.data(), a->d_inner + be[0], sizeof(int) * (size_t)len, s));
        for (int &v : o) v -= g;
        bool dup = false; for (auto &t : tabs) if (t == o) dup = true;
        if (!dup) tabs.push_back(o);
    }
    a->reg_ntab = (int)tabs.size(); a->reg_stride = 1;
    for (auto &t : tabs) a->reg_stride = std::max(a->reg_stride, (int)t.size());
    std::vector<int> tlen(std::max(1, a->reg_ntab)), toff((size_t)std::max(1, a->reg_ntab) * a->reg_stride, 0);
    for (int t = 0; t < a->reg_ntab; ++t) { tlen[t] = (int)tabs[t].size(); std::copy(tabs[t].begin(), tabs[t].end(), toff.begin() + (size_t)t * a->reg_stride); }
    dev_free(a->d_reg); dev_free(a->d_regoff); dev_free(a->d_reglen); a->d_reg = 0; a->d_regoff = 0; a->d_reglen = 0;
    GSB_TRY(upload(&a->d_reglen, tlen, s)); GSB_TRY(upload(&a->d_regoff, toff, s));
    GSB_TRY(dev_malloc((void **)&a->d_reg, (size_t)N + 1));
    GSB_LAUNCH(k_spmv_classify, dim3((N + 127) / 128), dim3(128), s, N, a->d_colptr, a->d_inner, a->reg_ntab, a->d_reglen, a->d_regoff, a->reg_stride, a->d_reg);
    // extent of the stored columns (ownership for the distributed CG, gsb200_device_view)
    int ext[5] = {N, 0, N, 0, 0}; int *d_ext = 0;
    GSB_TRY(dev_malloc((void **)&d_ext, sizeof ext)); GSB_TRY(dev_h2d(d_ext, ext, sizeof ext, s));
    GSB_LAUNCH(k_col_extent, dim3((N + 127) / 128), dim3(128), s, N, a->d_colptr, a->d_inner, d_ext);
    GSB_TRY(dev_d2h(ext, d_ext, sizeof ext, s)); dev_free(d_ext);
    if (ext[4] == 0) { ext[0] = ext[1] = ext[2] = ext[3] = 0; }
    a->own_c0 = ext[0]; a->own_c1 = ext[1]; a->need_lo = std::min(ext[2], ext[0]); a->need_hi = std::max(ext[3], ext[1]); a->own_contig = ext[4] == ext[1] - ext[0];
    a->spmv_ready = true;
    return 0;
}

// y[c] = (A x)[c] * scale[c] for the columns [c0, c1) (the others are left alone)
static int spmv_range(gsb200_assembler *a, int c0, int c1, const double *x, double *y, const double *scale, double *dot)
{
    if (c1 <= c0) return 0;
    GSB_TRY(spmv_prepare(a));
#ifndef GSB200_EMULATE
    if (!dry_run()) {
        k_spmv_reg<<<148 * 8, 256, 0, a->stream>>>(c0, c1, a->d_colptr, a->d_inner, a->d_values, a->d_reg, a->d_reglen, a->d_regoff, a->reg_stride, x, y, scale, dot);
        note_launch();
    }
#else
    GSB_LAUNCH(k_spmv_range, dim3((c1 - c0 + 127) / 128), dim3(128), a->stream, c0, c1, a->d_colptr, a->d_inner, a->d_values, x, y, scale, dot);
#endif
    return 0;
}

// ------------------------------------------------------------------ K4: coupled columns + right-hand side across the ranks
static int exchange(gsb200_assembler *a)
{
    a->xchg_bytes = 0;
    if (a->nranks == 1) return 0;
    GSB_TRY(comm_group(a, true));
    int rc = 0;
    for (size_t k = 0; k + 1 < a->coupled_off.size() && !rc; k += 2) {
        const i64 b = a->coupled_off[k], e = a->coupled_off[k + 1];
        if (e > b) { rc = comm_allreduce(a, a->d_values + b, e - b); a->xchg_bytes += 8 * (e - b); }
    }
    if (!rc) { rc = comm_allreduce(a, a->d_rhs, (i64)a->nfree * a->nrhs); a->xchg_bytes += 8 * (i64)a->nfree * a->nrhs; }
    const int rc2 = comm_group(a, false);
    ++a->xchg_calls;
    return rc ? rc : rc2;
}

// ------------------------------------------------------------------ CG
struct CgPlan { bool halo; int c0, c1; std::vector<int> ext; };    // ext: per rank {c0, c1, need_lo, need_hi}

static int cg_plan(gsb200_assembler *a, CgPlan &P)
{
    const int N = a->nfree;
    GSB_TRY(spmv_prepare(a));
    P.halo = false; P.c0 = 0; P.c1 = N;
    if (a->nranks == 1) return 0;
#ifndef GSB200_EMULATE
    if (a->comm && !a->ar_fn && a->coupled_runs.empty() && a->patches.size() == 1) {
        // contiguous column slabs: every rank learns everybody's owned range and row reach
        NcclApi *NC; GSB_TRY(need_nccl(&NC));
        int mine[4] = {a->own_c0, a->own_c1, a->need_lo, a->need_hi}; int *d_all = 0;
        if (!a->own_contig) mine[0] = mine[1] = -1;
        GSB_TRY(dev_malloc((void **)&d_all, sizeof(int) * 4 * (size_t)(a->nranks + 1)));
        GSB_TRY(dev_h2d(d_all + 4 * a->nranks, mine, sizeof mine, a->stream));
        GSB_TRY(nccl_check(NC->AllGather(d_all + 4 * a->nranks, d_all, 4, NCCL_INT32, a->comm, a->stream), "ncclAllGather"));
        P.ext.assign(4 * (size_t)a->nranks, 0);
        GSB_TRY(dev_d2h(P.ext.data(), d_all, sizeof(int) * 4 * (size_t)a->nranks, a->stream));
        dev_free(d_all);
        bool ok = true; int covered = 0;
        for (int r = 0; r < a->nranks; ++r) { if (P.ext[4 * r] < 0) ok = false; covered += P.ext[4 * r + 1] - P.ext[4 * r]; }
        if (ok && covered == N) { P.halo = true; P.c0 = a->own_c0; P.c1 = a->own_c1; }
    }
#endif
    return 0;
}

#ifndef GSB200_EMULATE
// the pieces of `v` this rank's rows reach into other ranks' slabs, and vice versa
static int cg_halo(gsb200_assembler *a, const CgPlan &P, double *v)
{
    NcclApi *NC; GSB_TRY(need_nccl(&NC));
    const int me = a->rank;
    GSB_TRY(nccl_check(NC->GroupStart(), "ncclGroupStart"));
    int rc = 0;
    for (int r = 0; r < a->nranks && !rc; ++r) {
        if (r == me) continue;
        // what I need from r: [need_lo, need_hi) of mine  intersected with r's slab
        const int lo = std::max(P.ext[4 * me + 2], P.ext[4 * r]), hi = std::min(P.ext[4 * me + 3], P.ext[4 * r + 1]);
        if (hi > lo) { rc = nccl_check(NC->Recv(v + lo, (size_t)(hi - lo), NCCL_FLOAT64, r, a->comm, a->stream), "ncclRecv"); a->xchg_bytes += 8 * (i64)(hi - lo); }
        const int slo = std::max(P.ext[4 * r + 2], P.ext[4 * me]), shi = std::min(P.ext[4 * r + 3], P.ext[4 * me + 1]);
        if (!rc && shi > slo) rc = nccl_check(NC->Send(v + slo, (size_t)(shi - slo), NCCL_FLOAT64, r, a->comm, a->stream), "ncclSend");
    }
    const int rc2 = nccl_check(NC->GroupEnd(), "ncclGroupEnd");
    return rc ? rc : rc2;
}
#endif

static int cg_solve(gsb200_assembler *a, const double *b_dev, int max_iter, double tol, int check_every, int *iters, double *rel_residual)
{
    const int N = a->nfree; stream_t s = a->stream;
    for (int k = 0; k < 8; ++k) if (!a->cgv[k]) GSB_TRY(dev_malloc((void **)&a->cgv[k], sizeof(double) * (size_t)(N + 8)));
    double *X = a->cgv[0], *R = a->cgv[1], *Z = a->cgv[2], *Pv = a->cgv[3], *Q = a->cgv[4], *Dg = a->cgv[5], *H = a->cgv[6], *S = a->cgv[7];
    CgPlan P; GSB_TRY(cg_plan(a, P));
    a->xchg_bytes = 0;
    const dim3 gN((N + 127) / 128), t(128);
    const int c0 = P.c0, c1 = P.c1;
#ifndef GSB200_EMULATE
    const dim3 gO(148 * 4), tv(256);        // grid-stride vector kernels
#else
    const dim3 gO(1), tv(1);
#endif
    const bool multi = a->nranks > 1;
    const bool rep = multi && !P.halo;          // whole vectors on every rank: the dot products are split by index range and reduced
    const int d0 = rep ? (int)((i64)N * a->rank / a->nranks) : c0, d1 = rep ? (int)((i64)N * (a->rank + 1) / a->nranks) : c1;
    const dim3 gD = gO;
    // preconditioner: diagonal of the stored columns; columns stored by several ranks (coupled, already exchanged) count once
    GSB_LAUNCH(k_cg_diag, gN, t, s, N, a->d_colptr, a->d_inner, a->d_values, Dg, H);
    if (multi && !P.halo) { GSB_TRY(comm_allreduce(a, Dg, N)); GSB_TRY(comm_allreduce(a, H, N)); }
    GSB_LAUNCH(k_cg_prep, gN, t, s, N, Dg, H);
    GSB_TRY(dev_memset(S, 0, 8 * sizeof(double), s));
    if (multi && !P.halo) { GSB_TRY(dev_memset(X, 0, sizeof(double) * (size_t)N, s)); }
    GSB_LAUNCH(k_cg_start, gO, tv, s, c0, c1, d0, d1, b_dev, Dg, X, R, Z, Pv, S);
    if (multi) GSB_TRY(comm_allreduce(a, S, 8));
    double hs[8]; GSB_TRY(dev_d2h(hs, S, sizeof hs, s));
    const double bb = hs[4]; double rr = bb;
    const double thr = tol * tol * bb;
    int it = 0;
    if (check_every < 1) check_every = 1;
#ifndef GSB200_EMULATE
    // GSB200_CG_PROFILE=1: device time of the phases of the second burst's iterations (events on the stream), printed to stderr
    const bool prof = getenv("GSB200_CG_PROFILE") != 0;
    cudaEvent_t pe[6]; float pacc[5] = {0, 0, 0, 0, 0}; int pn = 0;
    if (prof) for (auto &e : pe) cudaEventCreate(&e);
#define GSB_CG_MARK(i) do { if (prof && it >= check_every) cudaEventRecord(pe[i], s); } while (0)
#else
#define GSB_CG_MARK(i) do { } while (0)
#endif
#ifndef GSB200_EMULATE
    // NCCL connects its peer-to-peer and ring channels at the first use of each pattern (hundreds of milliseconds): one untimed
    // round of the loop's collectives (the halo exchange is idempotent, the scalars are scratch) before the clock starts
    if (P.halo) { GSB_TRY(cg_halo(a, P, Pv)); GSB_TRY(comm_allreduce(a, S + 6, 1)); GSB_TRY(comm_allreduce(a, S + 6, 2)); }
    else if (multi) { GSB_TRY(comm_allreduce(a, S + 6, 1)); GSB_TRY(comm_allreduce(a, S + 6, 2)); }
    a->xchg_bytes = 0;
    cudaEvent_t le0, le1; cudaEventCreate(&le0); cudaEventCreate(&le1); cudaEventRecord(le0, s);
#endif
    while (it < max_iter && rr > thr) {
        const int burst = std::min(check_every, max_iter - it);
        for (int k = 0; k < burst; ++k) {
            GSB_CG_MARK(0);
#ifndef GSB200_EMULATE
            if (P.halo) GSB_TRY(cg_halo(a, P, Pv));
#endif
            GSB_CG_MARK(1);
            if (P.halo || !multi) {
                GSB_TRY(spmv_range(a, c0, c1, Pv, Q, 0, S + 1));
                GSB_CG_MARK(2);
                if (P.halo) GSB_TRY(comm_allreduce(a, S + 1, 1));
            } else {
                // patch-wise ownership: every rank multiplies the columns it stores (shared ones weighted 1/holders), the products add up
                GSB_TRY(spmv_range(a, 0, N, Pv, Q, H, 0));
                GSB_TRY(comm_allreduce(a, Q, N)); a->xchg_bytes += 8 * (i64)N;
                GSB_LAUNCH(k_cg_dot, gD, tv, s, d0, d1, Pv, Q, S + 1);
                GSB_TRY(comm_allreduce(a, S + 1, 1));
            }
            GSB_CG_MARK(3);
            GSB_LAUNCH(k_cg_step1, gO, tv, s, c0, c1, d0, d1, Pv, Q, Dg, X, R, Z, S);
            GSB_CG_MARK(4);
            if (multi) GSB_TRY(comm_allreduce(a, S + 2, 2));
            GSB_LAUNCH(k_cg_step2, gO, tv, s, c0, c1, Z, Pv, S);
            GSB_LAUNCH(k_cg_rotate, dim3(1), dim3(32), s, S);
            GSB_CG_MARK(5);
#ifndef GSB200_EMULATE
            if (prof && it >= check_every && it < 2 * check_every) {
                cudaEventSynchronize(pe[5]);
                for (int i = 0; i < 5; ++i) { float ms = 0; cudaEventElapsedTime(&ms, pe[i], pe[i + 1]); pacc[i] += ms; }
                ++pn;
            }
#endif
            ++it;
        }
        GSB_TRY(dev_d2h(hs, S, sizeof hs, s));
        rr = hs[5];
        if (!(rr == rr)) { set_error("CG broke down (NaN residual) at iteration %d", it); return GSB200_ECUDA; }
    }
#ifndef GSB200_EMULATE
    cudaEventRecord(le1, s); cudaEventSynchronize(le1); { float ms = 0; cudaEventElapsedTime(&ms, le0, le1); a->cg_loop_ms = ms; }
    cudaEventDestroy(le0); cudaEventDestroy(le1);
    if (P.halo) {       // every rank gets the whole solution: each slab is broadcast by its owner
        NcclApi *NC; GSB_TRY(need_nccl(&NC));
        GSB_TRY(nccl_check(NC->GroupStart(), "ncclGroupStart"));
        int rc = 0;
        for (int r = 0; r < a->nranks && !rc; ++r) {
            const int lo = P.ext[4 * r], hi = P.ext[4 * r + 1];
            if (hi > lo) rc = nccl_check(NC->Broadcast(X + lo, X + lo, (size_t)(hi - lo), NCCL_FLOAT64, r, a->comm, s), "ncclBroadcast");
        }
        const int rc2 = nccl_check(NC->GroupEnd(), "ncclGroupEnd");
        if (rc || rc2) return rc ? rc : rc2;
    }
#endif
    GSB_TRY(dev_sync(s));
#ifndef GSB200_EMULATE
    if (prof) {
        if (pn) fprintf(stderr, "[gsb200 cg] rank %d: per iteration (ms, %d samples): halo %.3f  spmv %.3f  reduce(p.Ap) %.3f  update x,r,z %.3f  reduce(r.z, r.r) + update p %.3f\n",
                        a->rank, pn, pacc[0] / pn, pacc[1] / pn, pacc[2] / pn, pacc[3] / pn, pacc[4] / pn);
        for (auto &e : pe) cudaEventDestroy(e);
    }
#endif
#undef GSB_CG_MARK
    a->cg_halo_mode = P.halo;
    if (iters) *iters = it;
    if (rel_residual) *rel_residual = bb > 0 ? sqrt(rr / bb) : 0.0;
    return 0;
}

} // namespace gsb

// ====================================================================== C ABI (consumer side)
extern "C" {

int gsb200_comm_unique_id(void *id128)
{
    if (!id128) return GSB200_EINVAL;
#ifndef GSB200_EMULATE
    NcclApi *N; GSB_TRY(need_nccl(&N));
    return nccl_check(N->GetUniqueId(id128), "ncclGetUniqueId");
#else
    memset(id128, 0, GSB200_COMM_ID_BYTES); return GSB200_OK;
#endif
}

int gsb200_comm_init(gsb200_assembler *a, const void *id128)
{
    if (!a || !id128) return GSB200_EINVAL;
#ifndef GSB200_EMULATE
    GSB_TRY(select_device(a->device));
    NcclApi *N; GSB_TRY(need_nccl(&N));
    if (a->comm && a->comm_owned) N->CommDestroy(a->comm);
    a->comm = 0; a->comm_owned = false;
    ncclUniqueIdPod id; memcpy(&id, id128, sizeof id);
    GSB_TRY(nccl_check(N->CommInitRank(&a->comm, a->nranks, id, a->rank), "ncclCommInitRank"));
    a->comm_owned = true;
    return GSB200_OK;
#else
    set_error("no NCCL in the interpreter build: use gsb200_set_allreduce"); return GSB200_EUNSUPPORTED;
#endif
}

int gsb200_set_comm(gsb200_assembler *a, void *nccl_comm)
{
    if (!a) return GSB200_EINVAL;
#ifndef GSB200_EMULATE
    NcclApi *N; GSB_TRY(need_nccl(&N));
    if (nccl_comm && N->CommCount && N->CommUserRank) {
        int cnt = 0, rk = -1; N->CommCount(nccl_comm, &cnt); N->CommUserRank(nccl_comm, &rk);
        if (cnt != a->nranks || rk != a->rank) { set_error("communicator has rank %d of %d, the problem says rank %d of %d", rk, cnt, a->rank, a->nranks); return GSB200_EINVAL; }
    }
    if (a->comm && a->comm_owned) N->CommDestroy(a->comm);
    a->comm = nccl_comm; a->comm_owned = false;
    return GSB200_OK;
#else
    (void)nccl_comm; set_error("no NCCL in the interpreter build: use gsb200_set_allreduce"); return GSB200_EUNSUPPORTED;
#endif
}

int gsb200_set_allreduce(gsb200_assembler *a, gsb200_allreduce_fn fn, void *ctx)
{
    if (!a) return GSB200_EINVAL;
    a->ar_fn = fn; a->ar_ctx = ctx; return GSB200_OK;
}

int gsb200_exchange(gsb200_assembler *a)
{
    if (!a) { set_error("null assembler"); return GSB200_EINVAL; }
    if (!a->assembled) { set_error("gsb200_exchange before gsb200_assemble"); return GSB200_ESTATE; }
    GSB_TRY(select_device(a->device));
    return exchange(a);
}

int gsb200_comm_stats(const gsb200_assembler *a, int64_t *bytes_last, int32_t *calls)
{
    if (!a) return GSB200_EINVAL;
    if (bytes_last) *bytes_last = a->xchg_bytes;
    if (calls) *calls = a->xchg_calls;
    return GSB200_OK;
}

int gsb200_spmv_device(gsb200_assembler *a, const double *x_dev, double *y_dev)
{
    if (!a || !x_dev || !y_dev) return GSB200_EINVAL;
    if (!a->assembled) { set_error("spmv before assemble"); return GSB200_ESTATE; }
    GSB_TRY(select_device(a->device));
    return spmv_range(a, 0, a->nfree, x_dev, y_dev, 0, 0);
}

int gsb200_spmv_info(gsb200_assembler *a, int64_t *regular_columns, int32_t *tables)
{
    if (!a) return GSB200_EINVAL;
    if (!a->pattern_built) { set_error("pattern not built"); return GSB200_ESTATE; }
    GSB_TRY(select_device(a->device));
    GSB_TRY(spmv_prepare(a));
    if (regular_columns) {
        std::vector<unsigned char> reg((size_t)a->nfree + 1);
        GSB_TRY(dev_d2h(reg.data(), a->d_reg, (size_t)a->nfree, a->stream));
        int64_t n = 0; for (int c = 0; c < a->nfree; ++c) if (reg[c]) ++n;
        *regular_columns = n;
    }
    if (tables) *tables = a->reg_ntab;
    return GSB200_OK;
}

int gsb200_diag_device(gsb200_assembler *a, double *d_dev)
{
    if (!a || !d_dev) return GSB200_EINVAL;
    if (!a->assembled) { set_error("diag before assemble"); return GSB200_ESTATE; }
    GSB_TRY(select_device(a->device));
    const int n = a->nfree;
    GSB_LAUNCH(k_diag, dim3((n + 127) / 128), dim3(128), a->stream, n, a->d_colptr, a->d_inner, a->d_values, d_dev);
    return GSB200_OK;
}

int gsb200_diag_host(gsb200_assembler *a, double *d)
{
    if (!a || !d) return GSB200_EINVAL;
    if (!a->assembled) { set_error("diag before assemble"); return GSB200_ESTATE; }
    GSB_TRY(select_device(a->device));
    const int n = a->nfree;
    if (!a->cgv[1]) GSB_TRY(dev_malloc((void **)&a->cgv[1], sizeof(double) * (size_t)(n + 8)));
    GSB_LAUNCH(k_diag, dim3((n + 127) / 128), dim3(128), a->stream, n, a->d_colptr, a->d_inner, a->d_values, a->cgv[1]);
    return dev_d2h(d, a->cgv[1], sizeof(double) * (size_t)n, a->stream);
}

int gsb200_spmv_host(gsb200_assembler *a, const double *x, double *y)
{
    if (!a || !x || !y) return GSB200_EINVAL;
    if (!a->assembled) { set_error("spmv before assemble"); return GSB200_ESTATE; }
    GSB_TRY(select_device(a->device));
    const int n = a->nfree;
    for (int k = 0; k < 2; ++k) if (!a->cgv[k]) GSB_TRY(dev_malloc((void **)&a->cgv[k], sizeof(double) * (size_t)(n + 8)));
    GSB_TRY(dev_h2d(a->cgv[0], x, sizeof(double) * (size_t)n, a->stream));
    GSB_TRY(dev_memset(a->cgv[1], 0, sizeof(double) * (size_t)n, a->stream));
    GSB_TRY(spmv_range(a, 0, n, a->cgv[0], a->cgv[1], 0, 0));
    return dev_d2h(y, a->cgv[1], sizeof(double) * (size_t)n, a->stream);
}

int gsb200_cg_solve(gsb200_assembler *a, const double *b_host, double *x_host, int max_iter, double tol, int check_every,
                    int *iters, double *rel_residual)
{
    if (!a) return GSB200_EINVAL;
    if (!a->assembled) { set_error("cg before assemble"); return GSB200_ESTATE; }
    GSB_TRY(select_device(a->device));
    const int n = a->nfree;
    const double *b = a->d_rhs;
    if (b_host) {
        if (!a->cg_b) GSB_TRY(dev_malloc((void **)&a->cg_b, sizeof(double) * (size_t)(n + 1)));
        GSB_TRY(dev_h2d(a->cg_b, b_host, sizeof(double) * (size_t)n, a->stream));
        b = a->cg_b;
    }
    GSB_TRY(cg_solve(a, b, max_iter, tol, check_every, iters, rel_residual));
    if (x_host) GSB_TRY(dev_d2h(x_host, a->cgv[0], sizeof(double) * (size_t)n, a->stream));
    return GSB200_OK;
}

int gsb200_cg_host(gsb200_assembler *a, const double *b, double *x, int max_iter, double tol, int *iters, double *rel_residual)
{
    if (!b || !x) return GSB200_EINVAL;
    return gsb200_cg_solve(a, b, x, max_iter, tol, 1, iters, rel_residual);
}

int gsb200_field_norms(gsb200_assembler *a, const double *u_free, const gsb200_program *exact, const gsb200_program *exact_grad, double *out4)
{
    if (!a || !u_free || !out4) { set_error("field_norms: null argument"); return GSB200_EINVAL; }
    if (a->ncomp != 1) { set_error("field norms: scalar spaces only"); return GSB200_EUNSUPPORTED; }
    GSB_TRY(select_device(a->device));
    const int n = a->nfree; stream_t s = a->stream;
    double *d_u = 0, *d_out = 0;
    GSB_TRY(dev_malloc((void **)&d_u, sizeof(double) * (size_t)(n + 1)));
    GSB_TRY(dev_malloc((void **)&d_out, 4 * sizeof(double)));
    int rc = dev_h2d(d_u, u_free, sizeof(double) * (size_t)n, s);
    if (!rc) rc = dev_memset(d_out, 0, 4 * sizeof(double), s);
    NormArgs A; memset(&A, 0, sizeof A);
    const size_t nbuf0 = a->prog_bufs.size();
    if (!rc && exact) { rc = upload_device_program(a, *exact, &A.ex); A.has_exact = 1; }
    if (!rc && exact_grad) { for (int c = 0; c < a->dim && !rc; ++c) rc = upload_device_program(a, exact_grad[c], &A.exg[c]); A.has_grad = 1; }
    for (size_t ip = 0; ip < a->patches.size() && !rc; ++ip) {
        const PatchDev &P = a->patches[ip];
        i64 npt = 1;
        for (int k = 0; k < a->dim; ++k) {
            const Dir1D &d = P.dir[k];
            A.qn[k] = d.Q; A.gtab[k] = d.d_gtab; A.gfirst[k] = d.d_gfirst; A.pg1[k] = d.pg1; A.ngeo[k] = d.ngeo; A.hpt[k] = d.d_hpt; A.gwp[k] = d.d_gwp;
            A.tabl[k] = d.d_tabl; A.first[k] = d.d_first; A.p1[k] = d.p + 1; A.q[k] = d.q; A.nfun[k] = d.nfun;
            npt *= d.Q;
        }
        A.coefs = P.d_coefs; A.weights = P.d_weights; A.ngeo_total = P.ngeo_total;
        A.dofmap = P.d_dofmap; A.u = d_u; A.fixed = a->d_fixed; A.nfree = n; A.out = d_out;
        if (a->dim == 2) GSB_LAUNCH(k_field_norms<2>, dim3((unsigned)((npt + 127) / 128)), dim3(128), s, A);
        else GSB_LAUNCH(k_field_norms<3>, dim3((unsigned)((npt + 127) / 128)), dim3(128), s, A);
    }
    if (!rc) rc = dev_d2h(out4, d_out, 4 * sizeof(double), s);
    if (!rc && !exact) out4[0] = out4[2];
    if (!rc && !exact_grad) out4[1] = exact ? -1.0 : out4[3];
    while (a->prog_bufs.size() > nbuf0) { dev_free(a->prog_bufs.back()); a->prog_bufs.pop_back(); }
    dev_free(d_u); dev_free(d_out);
    return rc;
}

int gsb200_project_dirichlet(gsb200_assembler *a, const gsb200_neumann *sides, int nsides, int max_iter, double tol,
                             double *fixed_out, int *iters, double *rel_residual)
{
    if (!a || (nsides > 0 && !sides) || nsides < 0) { set_error("project_dirichlet: bad arguments"); return GSB200_EINVAL; }
    GSB_TRY(select_device(a->device));
    return project_dirichlet(a, sides, nsides, max_iter, tol, fixed_out, iters, rel_residual);
}

int gsb200_cg_info(const gsb200_assembler *a, double *loop_ms, int32_t *halo_exchange)
{
    if (!a) return GSB200_EINVAL;
    if (loop_ms) *loop_ms = a->cg_loop_ms;
    if (halo_exchange) *halo_exchange = a->cg_halo_mode ? 1 : 0;
    return GSB200_OK;
}

int gsb200_cg_solution_device(gsb200_assembler *a, const double **x_dev)
{
    if (!a || !x_dev) return GSB200_EINVAL;
    if (!a->cgv[0]) { set_error("no CG solution yet"); return GSB200_ESTATE; }
    *x_dev = a->cgv[0]; return GSB200_OK;
}

} // extern "C"
