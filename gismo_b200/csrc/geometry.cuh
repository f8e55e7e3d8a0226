// geometry.cuh — K0 of the assembly path: source-term machine, geometry map data and coefficient tensor at the
// quadrature points.  Included by kernels.cuh (inside namespace gsb) AND embedded verbatim as the text that
// NVRTC compiles together with the problem's source term (jit.cuh, GSB_JIT_SOURCE): keep it self-contained —
// it may only use i64, the GSB_* function macros, GSB_MAXP, the GSB200_* constants, st_stream and CUDA builtins.
// ------------------------------------------------------------------------------------
// Source-term stack machine (exprtk replacement).
// Short programs travel inside the kernel arguments (constant bank: no dependent global loads
// in the interpreter loop); long ones stay in global memory.
#define GSB_INLINE_OPS 48
#define GSB_INLINE_CONSTS 12
struct DevProgram {
    const int *ops; const double *consts; int nops;
    int inl;                                   // 1: use the inline copies below
    signed char iops[GSB_INLINE_OPS]; double iconsts[GSB_INLINE_CONSTS];
};
#ifndef GSB_JIT_SOURCE
GSB_HD double program_eval(const DevProgram &pr, double x, double y, double z)
{
    // stack machine with the top of stack cached in a register (t); st[] holds the rest
    double st[GSB200_PROGRAM_MAX_STACK];
    double t = 0.0;
    int sp = 0;
    for (int i = 0; i < pr.nops; ++i) {
        const int op = pr.inl ? (int)pr.iops[i] : pr.ops[i];
        switch (op) {
        case GSB200_OP_CONST: { ++i; const int ci = pr.inl ? (int)pr.iops[i] : pr.ops[i]; st[sp++] = t; t = pr.inl ? pr.iconsts[ci] : pr.consts[ci]; break; }
        case GSB200_OP_X: st[sp++] = t; t = x; break;
        case GSB200_OP_Y: st[sp++] = t; t = y; break;
        case GSB200_OP_Z: st[sp++] = t; t = z; break;
        case GSB200_OP_ADD: t = st[--sp] + t; break;
        case GSB200_OP_SUB: t = st[--sp] - t; break;
        case GSB200_OP_MUL: t = st[--sp] * t; break;
        case GSB200_OP_DIV: t = st[--sp] / t; break;
        case GSB200_OP_POW: t = pow(st[--sp], t); break;
        case GSB200_OP_NEG: t = -t; break;
        case GSB200_OP_SIN: t = sin(t); break;
        case GSB200_OP_COS: t = cos(t); break;
        case GSB200_OP_TAN: t = tan(t); break;
        case GSB200_OP_EXP: t = exp(t); break;
        case GSB200_OP_LOG: t = log(t); break;
        case GSB200_OP_SQRT: t = sqrt(t); break;
        case GSB200_OP_ABS: t = fabs(t); break;
        case GSB200_OP_TANH: t = tanh(t); break;
        case GSB200_OP_SINH: t = sinh(t); break;
        case GSB200_OP_COSH: t = cosh(t); break;
        case GSB200_OP_SQR: t = t * t; break;
#ifdef GSB200_EMULATE
        case GSB200_OP_SINPI: t = sin(3.14159265358979323846 * t); break;
        case GSB200_OP_COSPI: t = cos(3.14159265358979323846 * t); break;
#else
        case GSB200_OP_SINPI: t = sinpi(t); break;
        case GSB200_OP_COSPI: t = cospi(t); break;
#endif
        default: return NAN;
        }
    }
    return t;
}

#endif

// ------------------------------------------------------------------------------------
// K0: geometry map data at every quadrature point of the (chunk of the) patch:
// x, Jacobian (gsGeometry.hpp:557-564, rational quotient rule gsRationalBasis.h:481-520),
// measure and inverse (gsFunction.hpp:702-751), quadrature weight (gsQuadRule.h:190-200)
// folded into the form's coefficient tensor; and the load density w|J|f(x).
// Output layout: comp-major, then q0, q1, (q2) with the LAST direction fastest.
struct GeoArgs {
    int dim;
    int qn[3], qoff[3];            // window of 1-D points handled (count, first)
    const double2 *gtab[3]; const int *gfirst[3]; int pg1[3], ngeo[3];
    const double *hpt[3]; const double *gwp[3];              // per 1-D point: half element width, reference Gauss weight
    const double *coefs; const double *weights; i64 ngeo_total;
    int form, brow, bcol; double lambda, mu;
    int symD;                      // 1: write the 3/6 unique components, 0: all dim*dim
    double *D; i64 dstride;        // may be NULL (load only)
    double *F; i64 fstride; int nf; DevProgram prog[3];
};
// Common tail of the geometry kernels: inverse/measure of the Jacobian, quadrature weight, coefficient tensor of
// the form and load density at one point.
// SM = false: D / F live in global memory (streaming stores); SM = true: they are a shared-memory tile of the fused
// geometry + first-sweep kernel (fused.cuh), plain stores.
template <bool SM> GSB_DEVICE void geo_store(double *p, double v) { if (SM) *p = v; else st_stream(p, v); }
GSB_DEVICE void st_out(double *p, double v, int wb) { if (wb) *p = v; else st_stream(p, v); }
template <int DIM, int FSPEC, bool SM>      // FSPEC 1: Poisson with the symmetric coefficient tensor only (no run-time form dispatch)
GSB_DEVICE void geo_finish_to(const GeoArgs &A, double *Dp, i64 dstride, double *Fp, i64 fstride, i64 id, const int (&ql)[DIM], const double (&x)[3], const double (&J)[DIM][DIM])
{
    double Ji[DIM][DIM], det;   // Ji[a][c] = (J^-1)[a][c]
    if (DIM == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
    } else {
        const int X = DIM == 3 ? 2 : 0;  // keeps the 2-D instantiation in bounds
        const double c00 = J[1][1] * J[X][X] - J[1][X] * J[X][1], c01 = J[1][X] * J[X][0] - J[1][0] * J[X][X],
                     c02 = J[1][0] * J[X][1] - J[1][1] * J[X][0];
        det = J[0][0] * c00 + J[0][1] * c01 + J[0][X] * c02;
        const double id_ = 1.0 / det;
        Ji[0][0] = c00 * id_; Ji[0][1] = (J[0][X] * J[X][1] - J[0][1] * J[X][X]) * id_; Ji[0][X] = (J[0][1] * J[1][X] - J[0][X] * J[1][1]) * id_;
        Ji[1][0] = c01 * id_; Ji[1][1] = (J[0][0] * J[X][X] - J[0][X] * J[X][0]) * id_; Ji[1][X] = (J[0][X] * J[1][0] - J[0][0] * J[1][X]) * id_;
        Ji[X][0] = c02 * id_; Ji[X][1] = (J[0][1] * J[X][0] - J[0][0] * J[X][1]) * id_; Ji[X][X] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id_;
    }
    // quadrature weight: hprod * (w_0 w_1 w_2), same association as the reference
    double hprod = 1.0, wp = 1.0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        const double h = A.hpt[k][ql[k]];
        hprod *= (h == 0.0 ? 0.5 : h);
        const double g = A.gwp[k][ql[k]];
        wp = (k == 0) ? g : wp * g;
    }
    const double weight = hprod * wp * fabs(det);
#ifdef GSB_JIT_SOURCE       // the source term compiled for this problem (jit.cuh) instead of the stack machine
    if (Fp) for (int c = 0; c < A.nf; ++c) geo_store<SM>(Fp + c * fstride + id, weight * gsb_jit_src(c, x[0], x[1], x[2]));
#else
    if (Fp) for (int c = 0; c < A.nf; ++c) geo_store<SM>(Fp + c * fstride + id, weight * program_eval(A.prog[c], x[0], x[1], x[2]));
#endif
    if (!Dp) return;
    if (FSPEC != 1 && A.form == GSB200_FORM_MASS) { Dp[id] = weight; return; }
    double G[DIM][DIM];   // (J^-1 J^-T)_ab
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
        for (int b = a; b < DIM; ++b) {       // symmetric: the products commute, so the mirrored entry is the same number
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) s += Ji[a][c] * Ji[b][c];
            G[a][b] = s; G[b][a] = s;
        }
    if (FSPEC == 1 || (A.form == GSB200_FORM_POISSON && A.symD)) {
        int c = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a)
#pragma unroll
            for (int b = a; b < DIM; ++b) geo_store<SM>(Dp + (c++) * dstride + id, weight * G[a][b]);
        return;
    }
    if (A.form == GSB200_FORM_POISSON) {
#pragma unroll
        for (int a = 0; a < DIM; ++a)
#pragma unroll
            for (int b = 0; b < DIM; ++b) Dp[(a * DIM + b) * dstride + id] = weight * G[a][b];
        return;
    }
    // elasticity block (row comp r = brow carried by the partner/test function, col comp c = bcol by the owner):
    // E_{a'b'} = w ( lambda Ji[a'][r] Ji[b'][c] + mu ( Ji[a'][c] Ji[b'][r] + delta_rc G[a'][b'] ) ), a' on the row function.
    // The sweeps put the FIRST tensor index on the owner, hence the transpose when storing.
    const int r = A.brow, cc = A.bcol;
    double Jr[DIM], Jc[DIM];          // columns r and cc of J^-1, selected without indexing registers dynamically
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        Jr[k] = Ji[k][0]; Jc[k] = Ji[k][0];
#pragma unroll
        for (int m = 1; m < DIM; ++m) { if (r == m) Jr[k] = Ji[k][m]; if (cc == m) Jc[k] = Ji[k][m]; }
    }
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
        for (int b = 0; b < DIM; ++b) {
            const double E = A.lambda * Jr[b] * Jc[a] + A.mu * (Jc[b] * Jr[a] + (r == cc ? G[b][a] : 0.0));
            Dp[(a * DIM + b) * dstride + id] = weight * E;
        }
}

template <int DIM, int FSPEC>
GSB_DEVICE void geo_finish(const GeoArgs &A, i64 id, const int (&ql)[DIM], const double (&x)[3], const double (&J)[DIM][DIM])
{
    geo_finish_to<DIM, FSPEC, false>(A, A.D, A.dstride, A.F, A.fstride, id, ql, x, J);
}

// K0, line-factorised.  All threads of a block share the quadrature indices of the leading
// directions and differ only in the LAST one, so the tensor-product sum over the geometry's control points is
// split: the block first contracts the leading directions into "line coefficients"
//   E[a_L][field][kind] = sum_{a_0(,a_1)} B^(kind)(q_0(,q_1)) C_field[a_0(,a_1), a_L],  kind = value, d/dxi_0 (, d/dxi_1)
// (shared memory, a few hundred FMAs per block), then every thread finishes with a short 1-D sum over the
// (geometry degree + 1) functions of the last direction.  Fields = the geoDim coordinates (times the weight
// for a rational geometry) and the weight itself (gsGeometry.hpp:557-564, gsRationalBasis.h:481-520,
// gsFunction.hpp:702-751), ~5x fewer FP64 operations per point at degree 1-3 than the per-point tensor sum.
#define GSB_GEO_MAXA 48
#ifndef GSB200_EMULATE
#define GSB_SHARED __shared__
#define GSB_ALIGN16 __align__(16)
#define GSB_SYNCTHREADS() __syncthreads()
#define GSB_COOP_FIRST ((int)threadIdx.x)
#define GSB_COOP_STEP ((int)blockDim.x)
#else
#define GSB_SHARED
#define GSB_ALIGN16
#define GSB_SYNCTHREADS()
#define GSB_COOP_FIRST 0
#define GSB_COOP_STEP 1
#endif
// PGL = (geometry degree + 1) of the last direction (0 = run time), RATIONAL = NURBS weights present,
// FSPEC = 1 for the Poisson form with symmetric coefficient storage (the hot configuration)
template <int DIM, int PGL, bool RATIONAL, int FSPEC>
GSB_DEVICE void geometry_line_body(const GeoArgs &A)
{
    constexpr int L = DIM - 1, NFM = DIM + 1;
    GSB_SHARED double E[GSB_GEO_MAXA][NFM][DIM];
    const int q0blk = blockIdx.x * blockDim.x, qlast = q0blk + threadIdx.x;
    const bool active = qlast < A.qn[L];
    int ql[DIM];
    i64 id;
    ql[L] = (active ? qlast : A.qn[L] - 1) + A.qoff[L];
    if (DIM == 3) { ql[1] = blockIdx.y + A.qoff[1]; ql[0] = blockIdx.z + A.qoff[0]; id = ((i64)blockIdx.z * A.qn[1] + blockIdx.y) * A.qn[L] + qlast; }
    else { ql[0] = blockIdx.y + A.qoff[0]; id = (i64)blockIdx.y * A.qn[L] + qlast; }
    constexpr bool rational = RATIONAL;
    constexpr int nf = rational ? DIM + 1 : DIM;
    // range of last-direction control points touched by the block (gfirst is non-decreasing along the points)
    const int qb_first = q0blk + A.qoff[L];
    const int qb_last = (q0blk + (int)blockDim.x - 1 < A.qn[L] ? q0blk + (int)blockDim.x - 1 : A.qn[L] - 1) + A.qoff[L];
    const int pgL = PGL ? PGL : A.pg1[L];
    const int lo = A.gfirst[L][qb_first], hi = A.gfirst[L][qb_last] + pgL;
    const int gfL = A.gfirst[L][ql[L]];
    int gf[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) gf[k] = A.gfirst[k][ql[k]];
    double val[NFM], dd[NFM][DIM];      // dd[f][k] = d field_f / d xi_k
#pragma unroll
    for (int f = 0; f < NFM; ++f) { val[f] = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dd[f][k] = 0.0; }
    for (int abase = lo; abase < hi; abase += GSB_GEO_MAXA) {
        const int cnt = hi - abase < GSB_GEO_MAXA ? hi - abase : GSB_GEO_MAXA;
        GSB_SYNCTHREADS();
        for (int item = GSB_COOP_FIRST; item < cnt * nf; item += GSB_COOP_STEP) {
            const int aa = item / nf, f = item - aa * nf, aL = abase + aa;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            const int n1 = DIM == 3 ? A.pg1[1] : 1;
            for (int a1 = 0; a1 < n1; ++a1) {
                const double2 b1 = DIM == 3 ? A.gtab[1][(i64)ql[1] * A.pg1[1] + a1] : make_double2(1.0, 0.0);
                const i64 row = DIM == 3 ? ((i64)aL * A.ngeo[1] + (gf[1] + a1)) * A.ngeo[0] + gf[0] : (i64)aL * A.ngeo[0] + gf[0];
                for (int a0 = 0; a0 < A.pg1[0]; ++a0) {
                    const double2 b0 = A.gtab[0][(i64)ql[0] * A.pg1[0] + a0];
                    const i64 idx = row + a0;
                    double C = f < DIM ? A.coefs[(i64)f * A.ngeo_total + idx] : 1.0;
                    if (rational) C *= A.weights[idx];
                    s0 = fma(b0.x * b1.x, C, s0); s1 = fma(b0.y * b1.x, C, s1); s2 = fma(b0.x * b1.y, C, s2);
                }
            }
            E[aa][f][0] = s0; E[aa][f][1] = s1; if (DIM == 3) E[aa][f][DIM - 1] = s2;
        }
        GSB_SYNCTHREADS();
#pragma unroll
        for (int k = 0; k < (PGL ? PGL : GSB_MAXP + 1); ++k) {
            if (k >= pgL) break;
            const int aa = gfL + k - abase;
            if (aa < 0 || aa >= cnt) continue;
            const double2 bL = A.gtab[L][(i64)ql[L] * pgL + k];
#pragma unroll
            for (int f = 0; f < NFM; ++f) {
                if (f >= nf) break;
                val[f] = fma(bL.x, E[aa][f][0], val[f]);
                dd[f][0] = fma(bL.x, E[aa][f][1], dd[f][0]);
                if (DIM == 3) dd[f][1] = fma(bL.x, E[aa][f][DIM - 1], dd[f][1]);
                dd[f][L] = fma(bL.y, E[aa][f][0], dd[f][L]);
            }
        }
    }
    if (!active) return;
    double x[3] = {0.0, 0.0, 0.0}, J[DIM][DIM];   // J[c][a] = d x_c / d xi_a
    if (rational) {
        const double W = val[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            x[c] = val[c] / W;
#pragma unroll
            for (int a = 0; a < DIM; ++a) J[c][a] = (dd[c][a] * W - val[c] * dd[DIM][a]) / (W * W);
        }
    } else {
#pragma unroll
        for (int c = 0; c < DIM; ++c) { x[c] = val[c];
#pragma unroll
            for (int a = 0; a < DIM; ++a) J[c][a] = dd[c][a]; }
    }
    geo_finish<DIM, FSPEC>(A, id, ql, x, J);
}

template <int DIM, int PGL, bool RATIONAL, int FSPEC>
GSB_GLOBAL void k_geometry_line(const GeoArgs A) { geometry_line_body<DIM, PGL, RATIONAL, FSPEC>(A); }

