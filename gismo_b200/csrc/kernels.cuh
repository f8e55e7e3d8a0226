// kernels.cuh — device kernels of the B200 assembly path (sm_100a).
//
// Algorithm: GLOBAL sum factorisation with row (here: CSC-column) ownership.
//   K_{j,i} = sum_q sum_{ab} D_ab(q) d_a N_j(q) d_b N_i(q)  with  N_i = prod_k B_{i_k}(xi_k)
// is contracted one parametric direction at a time over the whole patch:
//   K0  geometry:   D_ab(q) = w_q |det J| (J^-1 J^-T)_ab  (+ load density w|J|f)   [a6,a7,a8,a18]
//   S1  sweep dir 0: A1[o][(i0,d0)][q1,q2] = sum_{q0} B^a_{i0} B^b_{i0+d0} D_c
//   S2  sweep dir 1: A2[g][(i1,d1)][(i0,d0)][q2]
//   S3  sweep dir 2: K[(i0,i1,i2),(d0,d1,d2)] -> written straight into its CSC slot
// Each sweep thread owns one "column" of the remaining index space and marches along the
// swept direction keeping, in registers, the (p+1)^2 partial sums of the function pairs
// alive on the current knot span; a pair is complete when its older function leaves the
// span window, and is then emitted exactly once.  No atomics, no element-local matrices,
// every matrix entry is produced by exactly one thread (its column owner).
// 1-D basis values/derivatives come from de Boor / A2.3 recursion evaluated in registers
// once per 1-D quadrature point (k_basis_table) and are broadcast from L1.
//
// Reference functions replaced: a2 gsBSplineBasis.hpp:863-1041, a3
// gsTensorBSplineBasis.hpp:166-205, a4 gsTensorBasis.hpp:634-728, a6 gsGeometry.hpp:539-597,
// a7 gsFunction.hpp:604-862, a8 gsQuadRule.h:177-201, a10 gsVisitorPoisson.h:62-118,
// a11 gsExpressions.h (igrad/ijac/idiv/meas), a14/a15 gsSparseSystem.h:972-1010 /
// gsExprAssembler.h:553-630, a16 SparseMatrix.h:208-225,1366-1396, a18 gsFunctionExpr.hpp:513.
#pragma once
#include "platform.cuh"
#include "../../include/gsb200.h"

namespace gsb {

#define GSB_MAXP 7            // max degree of the 1-D table kernels
#ifdef GSB200_EMULATE
#define GSB_CX constexpr
#else
#define GSB_CX __host__ __device__ constexpr
#endif

#include "terms.cuh"   // static_for, term tables, output-group helpers (also part of the text NVRTC compiles, jit.cuh)

// ------------------------------------------------------------------------------------
// 1-D B-spline values + first derivatives on knot span s at u: Cox-de Boor triangle kept
// in registers, derivative from the degree p-1 row (NURBS-book A2.3 with n=1).
GSB_HD void bspline_ders(const double *kn, int p, int s, double u, double *val, double *der)
{
    double ndu[(GSB_MAXP + 1) * (GSB_MAXP + 1)], left[GSB_MAXP + 1], right[GSB_MAXP + 1];
    const int p1 = p + 1;
    ndu[0] = 1.0;
    for (int j = 1; j <= p; ++j) {
        left[j] = u - kn[s + 1 - j];
        right[j] = kn[s + j] - u;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            ndu[j * p1 + r] = right[r + 1] + left[j - r];
            const double temp = ndu[r * p1 + j - 1] / ndu[j * p1 + r];
            ndu[r * p1 + j] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        ndu[j * p1 + j] = saved;
    }
    for (int j = 0; j <= p; ++j) val[j] = ndu[j * p1 + p];
    for (int r = 0; r <= p; ++r) {
        double d = 0.0;
        if (r >= 1) d = (1.0 / ndu[p * p1 + r - 1]) * ndu[(r - 1) * p1 + p - 1];
        if (r <= p - 1) d += (-1.0 / ndu[p * p1 + r]) * ndu[r * p1 + p - 1];
        der[r] = d * (double)p;
    }
}

// One thread per 1-D quadrature point (e,t) of the solution basis: mapped Gauss node
// (gsQuadRule.h:177-201), basis values/derivatives stored in SLOT order (slot = function
// index mod (p+1)) so that the sweep kernels index their register accumulators statically.
struct BasisTableArgs {
    const double *knots; const int *span; const double *gnodes; const double *gweights;
    int p, nel, q;
    double2 *tab;      // [nel*q][p+1] slot order
    double2 *tabl;     // [nel*q][p+1] local order (a-th active function of the span)
    double *upt;       // [nel*q] point coordinate
    double *hpt;       // [nel*q] half element width h
    double *gwp;       // [nel*q] reference Gauss weight of the point
};
GSB_GLOBAL void k_basis_table(const BasisTableArgs A)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= A.nel * A.q) return;
    const int e = id / A.q, t = id - e * A.q, s = A.span[e], p1 = A.p + 1;
    const double lower = A.knots[s], h = (A.knots[s + 1] - lower) / 2.0;
    const double u = h * (A.gnodes[t] + 1.0) + lower;
    double val[GSB_MAXP + 1], der[GSB_MAXP + 1];
    bspline_ders(A.knots, A.p, s, u, val, der);
    const int first = s - A.p;
    for (int a = 0; a < p1; ++a) A.tab[(i64)id * p1 + (first + a) % p1] = make_double2(val[a], der[a]);
    for (int a = 0; a < p1; ++a) A.tabl[(i64)id * p1 + a] = make_double2(val[a], der[a]);
    A.upt[id] = u;
    A.hpt[id] = h;
    A.gwp[id] = A.gweights[t];
}

// Geometry basis (its own, usually coarse, knot vector) at the same 1-D points.
struct GeoTableArgs {
    const double *knots; int nknots, p, npts;
    const double *upt;
    double2 *gtab;     // [npts][p+1] in local order
    int *gfirst;       // [npts] first active geometry function
};
GSB_GLOBAL void k_geo_table(const GeoTableArgs A)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= A.npts) return;
    const double u = A.upt[id];
    int lo = A.p, hi = A.nknots - A.p - 1;   // span search: upper_bound - 1 (gsKnotVector.hpp:747-783)
    if (u >= A.knots[hi]) { lo = hi - 1; while (A.knots[lo] == A.knots[lo + 1]) --lo; }
    else while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (A.knots[mid] <= u) lo = mid; else hi = mid; }
    double val[GSB_MAXP + 1], der[GSB_MAXP + 1];
    bspline_ders(A.knots, A.p, lo, u, val, der);
    for (int a = 0; a <= A.p; ++a) A.gtab[(i64)id * (A.p + 1) + a] = make_double2(val[a], der[a]);
    A.gfirst[id] = lo - A.p;
}

#include "geometry.cuh"   // K0: source-term machine, map data, coefficient tensor (also the text NVRTC compiles, jit.cuh)
#include "fused.cuh"      // K0 + first sweep in one kernel (D stays in shared memory); NVRTC text as well

// ------------------------------------------------------------------------------------
// Neumann boundary load (gsVisitorNeumann.h:83-136, gsExprAssembler.h:835-895): flux density at the
// boundary quadrature points of one patch side.  The fixed direction contributes one node at the
// boundary parameter (its basis values come pre-evaluated from the host), the outer normal is built
// from the first minors of the Jacobian (gsFunction.hpp:613-699).
struct FaceArgs {
    int dim, dir, upper;
    int qn[3];                                   // points per face direction (qn[dir] unused)
    const double2 *gtab[3]; const int *gfirst[3]; int pg1[3], ngeo[3];
    const double *hpt[3]; const double *gwp[3];
    double2 bgeo[GSB_MAXP + 1]; int bgfirst;       // geometry basis (value, derivative) of `dir` at the boundary parameter
    const double *coefs; const double *weights; i64 ngeo_total;
    int ndata; DevProgram prog[3];
    double *Fb;                                  // [qa][qb], last face direction fastest
    int vol_measure;                             // scalar data times |det J| instead of |n| (what the reference's Dirichlet L2-projection weighs with)
};
template <int DIM>
GSB_GLOBAL void k_face_geometry(const FaceArgs A)
{
    int fd[2] = {0, 0}, nf = 0;                  // the face's own directions, ascending
    for (int k = 0; k < DIM; ++k) if (k != A.dir) fd[nf++] = k;
    const i64 total = (DIM == 3) ? (i64)A.qn[fd[0]] * A.qn[fd[1]] : (i64)A.qn[fd[0]];
    const i64 id = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    int ql[3] = {0, 0, 0};
    if (DIM == 3) { ql[fd[1]] = (int)(id % A.qn[fd[1]]); ql[fd[0]] = (int)(id / A.qn[fd[1]]); } else ql[fd[0]] = (int)id;
    double W = 0.0, dW[3] = {0, 0, 0}, xn[3] = {0, 0, 0}, dxn[3][3];
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) dxn[a][c] = 0.0;
    int cnt[3] = {0, 0, 0}, gf[3];
    for (int k = 0; k < DIM; ++k) gf[k] = (k == A.dir) ? A.bgfirst : A.gfirst[k][ql[k]];
    for (;;) {
        i64 idx = 0; double v = 1.0, dv[3]; double2 b[3];
        for (int k = DIM - 1; k >= 0; --k) idx = idx * A.ngeo[k] + (gf[k] + cnt[k]);
        for (int k = 0; k < DIM; ++k) { b[k] = (k == A.dir) ? A.bgeo[cnt[k]] : A.gtab[k][(i64)ql[k] * A.pg1[k] + cnt[k]]; v *= b[k].x; }
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
    for (int c = 0; c < DIM; ++c) { x[c] = xn[c] / W; for (int a = 0; a < DIM; ++a) Jt[a][c] = (dxn[a][c] * W - xn[c] * dW[a]) / (W * W); }
    double nrm[3] = {0, 0, 0}, det;
    if (DIM == 2) {
        det = Jt[0][0] * Jt[1][1] - Jt[0][1] * Jt[1][0];
        const int o = 1 - A.dir;
        nrm[0] = Jt[o][1]; nrm[1] = -Jt[o][0];
    } else {
        det = Jt[0][0] * (Jt[1][1] * Jt[2][2] - Jt[1][2] * Jt[2][1]) - Jt[0][1] * (Jt[1][0] * Jt[2][2] - Jt[1][2] * Jt[2][0]) +
              Jt[0][2] * (Jt[1][0] * Jt[2][1] - Jt[1][1] * Jt[2][0]);
        const int r0 = A.dir == 0 ? 1 : 0, r1 = A.dir == 2 ? 1 : 2;
        nrm[0] = Jt[r0][1] * Jt[r1][2] - Jt[r0][2] * Jt[r1][1];
        nrm[1] = -(Jt[r0][0] * Jt[r1][2] - Jt[r0][2] * Jt[r1][0]);
        nrm[2] = Jt[r0][0] * Jt[r1][1] - Jt[r0][1] * Jt[r1][0];
    }
    const int side = 2 * A.dir + A.upper + 1;     // sideOrientation(s), gsBoundary.h:1029-1035
    const double sgn = (((side + (side + 1) / 2) % 2) ? 1.0 : -1.0) * (det < 0 ? -1.0 : 1.0);
    double nn = 0.0;
    for (int c = 0; c < DIM; ++c) { nrm[c] *= sgn; nn += nrm[c] * nrm[c]; }
    double flux;
    if (A.ndata == 1) flux = program_eval(A.prog[0], x[0], x[1], x[2]) * (A.vol_measure ? fabs(det) : sqrt(nn));
    else { flux = 0.0; for (int c = 0; c < DIM; ++c) flux += program_eval(A.prog[c], x[0], x[1], x[2]) * nrm[c]; }
    double w = 1.0;                               // the fixed direction contributes h=0 -> 0.5 and weight 2
    for (int k = 0; k < DIM; ++k) if (k != A.dir) w *= A.hpt[k][ql[k]] * A.gwp[k][ql[k]];
    A.Fb[id] = w * flux;
}

// One thread per face basis function: contracts the flux density with the tensor basis over the
// function's support and adds the result to the rhs rows of the (<= p+1) functions alive at the boundary.
struct FaceLoadArgs {
    int dim, dir;
    int nfun[3], p1[3], q[3], Q[3];
    const int *ffirst[3], *flast[3]; const double2 *tab[3];
    double bval[GSB_MAXP + 1]; int bfirst, nb1;   // solution basis values of `dir` at the boundary parameter
    const double *Fb; const int *dofmap; double *rhs; int nfree;
    // boundary L2-projection (gsDirichletValues.h:257-435): rows are the ELIMINATED DOFs (index - nfree); square: N_i^2 (diagonal)
    int to_fixed, square;
};
template <int DIM>
GSB_GLOBAL void k_face_load(const FaceLoadArgs A)
{
    int fd[2] = {0, 0}, nf = 0;
    for (int k = 0; k < DIM; ++k) if (k != A.dir) fd[nf++] = k;
    const i64 total = (DIM == 3) ? (i64)A.nfun[fd[0]] * A.nfun[fd[1]] : (i64)A.nfun[fd[0]];
    const i64 id = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    int fi[3] = {0, 0, 0};
    if (DIM == 3) { fi[fd[1]] = (int)(id % A.nfun[fd[1]]); fi[fd[0]] = (int)(id / A.nfun[fd[1]]); } else fi[fd[0]] = (int)id;
    const int a = fd[0], b = fd[1];
    double s = 0.0;
    for (int ea = A.ffirst[a][fi[a]]; ea <= A.flast[a][fi[a]]; ++ea)
        for (int ta = 0; ta < A.q[a]; ++ta) {
            const int qa = ea * A.q[a] + ta;
            double va = A.tab[a][(i64)qa * A.p1[a] + fi[a] % A.p1[a]].x;
            if (A.square) va *= va;
            if (DIM == 2) { s = fma(va, A.Fb[qa], s); continue; }
            double sb = 0.0;
            for (int eb = A.ffirst[b][fi[b]]; eb <= A.flast[b][fi[b]]; ++eb)
                for (int tb = 0; tb < A.q[b]; ++tb) {
                    const int qb = eb * A.q[b] + tb;
                    double vb = A.tab[b][(i64)qb * A.p1[b] + fi[b] % A.p1[b]].x;
                    if (A.square) vb *= vb;
                    sb = fma(vb, A.Fb[(i64)qa * A.Q[b] + qb], sb);
                }
            s = fma(va, sb, s);
        }
    for (int k = 0; k < A.nb1; ++k) {
        fi[A.dir] = A.bfirst + k;
        const i64 li = ((i64)(DIM == 3 ? fi[2] : 0) * A.nfun[1] + fi[1]) * A.nfun[0] + fi[0];
        const int g = A.dofmap[li];
        const double v = s * (A.square ? A.bval[k] * A.bval[k] : A.bval[k]);
        if (A.to_fixed) { if (g >= A.nfree && v != 0.0) atomic_add(A.rhs + (g - A.nfree), v); }
        else if (g < A.nfree && v != 0.0) atomic_add(A.rhs + g, v);
    }
}

// Trace of a field of ELIMINATED coefficients x (index - nfree) at the boundary quadrature points of a side, times the point weights
// Wb (= w |n|, k_face_geometry with data 1): out = Wb .* (B x).  With k_face_load(to_fixed) this applies the boundary mass matrix of
// the Dirichlet L2-projection without forming it.
struct FaceEvalArgs {
    int dim, dir;
    int qn[3], nfun[3], p1[3], q[3];
    const double2 *tabl[3]; const int *first[3];
    double bval[GSB_MAXP + 1]; int bfirst, nb1;
    const int *dofmap; const double *x; int nfree;
    const double *Wb; double *out;
};
template <int DIM>
GSB_GLOBAL void k_face_eval(const FaceEvalArgs A)
{
    int fd[2] = {0, 0}, nf = 0;
    for (int k = 0; k < DIM; ++k) if (k != A.dir) fd[nf++] = k;
    const i64 total = (DIM == 3) ? (i64)A.qn[fd[0]] * A.qn[fd[1]] : (i64)A.qn[fd[0]];
    const i64 id = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    int ql[3] = {0, 0, 0};
    if (DIM == 3) { ql[fd[1]] = (int)(id % A.qn[fd[1]]); ql[fd[0]] = (int)(id / A.qn[fd[1]]); } else ql[fd[0]] = (int)id;
    const int a = fd[0], b = fd[1];
    const int fa0 = A.first[a][ql[a] / A.q[a]], fb0 = DIM == 3 ? A.first[b][ql[b] / A.q[b]] : 0;
    double s = 0.0;
    for (int k = 0; k < A.nb1; ++k) {
        if (A.bval[k] == 0.0) continue;
        int fi[3] = {0, 0, 0};
        fi[A.dir] = A.bfirst + k;
        double sk = 0.0;
        for (int ja = 0; ja < A.p1[a]; ++ja) {
            fi[a] = fa0 + ja;
            const double va = A.tabl[a][(i64)ql[a] * A.p1[a] + ja].x;
            for (int jb = 0; jb < (DIM == 3 ? A.p1[b] : 1); ++jb) {
                double v = va;
                if (DIM == 3) { fi[b] = fb0 + jb; v *= A.tabl[b][(i64)ql[b] * A.p1[b] + jb].x; }
                const i64 li = ((i64)(DIM == 3 ? fi[2] : 0) * A.nfun[1] + fi[1]) * A.nfun[0] + fi[0];
                const int g = A.dofmap[li];
                if (g >= A.nfree) sk = fma(v, A.x[g - A.nfree], sk);
            }
        }
        s = fma(A.bval[k], sk, s);
    }
    A.out[id] = A.Wb[id] * s;
}

// ------------------------------------------------------------------------------------
// Final-stage scatter context: where an owner/partner pair lands in the CSC arrays.
struct FinalArgs {
    int dim, L;                    // L = last (swept) direction
    int n[3], p[3];                // functions / degree per direction
    const int *plo[3]; const int *phi[3];   // co-occurring partner range per function, per direction
    const int *dofmap; i64 nb;     // ncomp blocks of nb global indices
    int brow, bcol;
    const unsigned char *colflag;  // per (comp, local fn): 0 skip, 1 canonical, 2 generic, 3 canonical + full stencil
    const i64 *ownrec;             // per (comp, local fn): (colptr << 2) | flag
    const unsigned *st; int nrun;  // canonical slot table: per (row component, fn, run) start | mask<<16 of the standard rows
    const unsigned *st2;           // ... and of the rows coupled with other patches (0: single patch)
    const i64 *colptr; const int *inner; double *values;
    double *rhs; const double *fixed; int nfree, nfixed, nrhs;
};
struct FinalCtx { i64 li_low, dj_low, nlow; int r_low, bit0; };

GSB_DEVICE bool final_init(const FinalArgs &F, i64 outer, i64 inner, FinalCtx &c)
{
    // 2-D: inner = (i0, d0).  3-D: outer = i1, inner = (i0, d1, d0) so that consecutive lanes hit
    // consecutive slots (d1, d0) of one CSC column.
    const int W0 = 2 * F.p[0] + 1;
    if (F.dim == 2) {
        const int i0 = (int)(inner / W0), d0 = (int)(inner % W0) - F.p[0];
        const int j0 = i0 + d0;
        if (j0 < F.plo[0][i0] || j0 > F.phi[0][i0]) return false;
        c.bit0 = d0 + F.p[0];
        c.li_low = i0; c.dj_low = d0; c.nlow = F.n[0]; c.r_low = 0;
        return true;
    }
    const int W1 = 2 * F.p[1] + 1;
    const int i0 = (int)(inner / (W1 * W0)), r = (int)(inner % (W1 * W0));
    const int d1 = r / W0 - F.p[1], d0 = r % W0 - F.p[0];
    const int i1 = (int)outer;
    const int j0 = i0 + d0, j1 = i1 + d1;
    if (j0 < F.plo[0][i0] || j0 > F.phi[0][i0]) return false;
    if (j1 < F.plo[1][i1] || j1 > F.phi[1][i1]) return false;
    c.bit0 = d0 + F.p[0];
    c.li_low = (i64)i1 * F.n[0] + i0; c.dj_low = (i64)d1 * F.n[0] + d0; c.nlow = (i64)F.n[0] * F.n[1];
    c.r_low = d1 + F.p[1];
    return true;
}

// Per-owner data: ONE packed word per (component, local function) prepared after the pattern
// build (k_owner_records): (colptr[gi] << 2) | flag, 0 for eliminated / foreign columns.  The
// sweep fetches it one span before the owner's first emit and keeps it in registers.
struct OwnerCache { int fun; i64 rec; };

GSB_DEVICE void owner_load(const FinalArgs &F, const FinalCtx &c, int iL, OwnerCache &oc)
{
    oc.fun = iL;
    oc.rec = F.ownrec[F.bcol * F.nb + (i64)iL * c.nlow + c.li_low];
}

GSB_GLOBAL void k_owner_records(i64 n, int nfree, const int *dofmap, const unsigned char *colflag, const i64 *colptr, i64 *ownrec)
{
    const i64 id = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const int g = dofmap[id];
    const int flag = g < nfree ? (int)colflag[id] : 0;
    ownrec[id] = flag ? ((colptr[g] << 2) | flag) : 0;
}

// -> position in `values` for the common case (canonical column, free row); slow paths
// (eliminated row, generic column) are completed here and -1 is returned.
GSB_DEVICE i64 final_prepare(const FinalArgs &F, const FinalCtx &c, const OwnerCache &oc, int dL, double val)
{
    const int flag = (int)(oc.rec & 3);
    if (!flag) return -1;
    const i64 base = oc.rec >> 2;
    if (flag == 3) {        // full interior stencil: every partner is free, rank is the lexicographic stencil index (per row component)
        const i64 full = (i64)(2 * F.p[0] + 1) * (2 * F.p[1] + 1) * (F.dim == 3 ? 2 * F.p[2] + 1 : 1);
        return base + F.brow * full + ((F.dim == 2) ? (dL + F.p[1]) : ((dL + F.p[2]) * (2 * F.p[1] + 1) + c.r_low)) * (2 * F.p[0] + 1) + c.bit0;
    }
    const i64 li = (i64)oc.fun * c.nlow + c.li_low;
    const i64 lj = li + (i64)dL * c.nlow + c.dj_low;
    if (flag == 1) {
        const int run = (F.dim == 2) ? (dL + F.p[1]) : ((dL + F.p[2]) * (2 * F.p[1] + 1) + c.r_low);
        const unsigned w = F.st[((i64)F.brow * F.nb + li) * F.nrun + run];
        const unsigned mask = w >> 16;
        if ((mask >> c.bit0) & 1u)                   // partner row is free: its rank in the column is known
            return base + (int)(w & 0xffffu) + popc(mask & ((1u << c.bit0) - 1u));
        if (F.st2) {                                 // ... or a DOF shared with another patch: second sequence of the column
            const unsigned w2 = F.st2[((i64)F.brow * F.nb + li) * F.nrun + run];
            const unsigned mask2 = w2 >> 16;
            if ((mask2 >> c.bit0) & 1u) return base + (int)(w2 & 0xffffu) + popc(mask2 & ((1u << c.bit0) - 1u));
        }
        if (!F.fixed) return -1;
    }
    const int gi = F.dofmap[F.bcol * F.nb + li];
    const int gj = F.dofmap[F.brow * F.nb + lj];
    if (gj >= F.nfree) {                             // eliminated row: by symmetry of the form this is the
        if (F.fixed)                                 // -K(i,j) g_j term of gsSparseSystem.h:1004
            for (int r = 0; r < F.nrhs; ++r) atomic_add(F.rhs + (i64)r * F.nfree + gi, -val * F.fixed[(i64)r * F.nfixed + (gj - F.nfree)]);
        return -1;
    }
    int lo = 0, hi = (int)(F.colptr[gi + 1] - base) - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (F.inner[base + mid] < gj) lo = mid + 1; else hi = mid; }
    atomic_add(F.values + base + lo, val);
    return -1;
}

// ------------------------------------------------------------------------------------
// The sweep kernel (S1/S2/S3).  Thread = one column of the non-swept index space (lanes run
// along the contiguous dimension of both input and output), blockIdx.y = group of owner
// slots handled, blockIdx.z = sweep segment.
struct SweepArgs {
    const int *first, *nexit; const double2 *tab; const double2 *tabl; int q, p;     // swept direction tables (slot / local order)
    const int *seg;                                            // [nseg][4] e_begin,e_end,x_min,x_max
    const double *in; double *out;
    i64 in_cs, in_es, in_ts, in_os, in_is; int e_in0;          // input strides: comp, element, point, outer, inner
    i64 out_cs, out_fs, out_ds, out_os, out_os2, out_od, out_bs, out_is, out_bq;   // output strides: comp, owner fn, delta, outer (split at out_od), block, inner
    // symmetry of the form (D symmetric): the first sweep may emit only delta >= 0 (half_out, delta index offset
    // d_off instead of p) and the second sweep then writes every value to its mirrored slot as well (mirror)
    int half_out, d_off, mirror; i64 out_dshift, out_nprev;
    int pf_dist;                 // window kernel: spans of L2 prefetch distance (0 = none)
    // window kernel: optional 3-level decomposition of the inner column index for the INPUT address
    // (inner = (x, y, z) with z fastest: x * in_bs2 + y * in_bs + z * in_is), in_bq = 0: plain inner * in_is
    i64 in_bq, in_bs, in_bq2, in_bs2;
    // window kernel, mirrored reads: inner = (.., delta slot, t) with mir_q points per slot and mir_w slots; the mirrored pair of
    // (outer, slot) is (outer + slot - mir_p, 2 mir_p - slot), valid while 0 <= outer + slot - mir_p < mir_n
    int mir_q, mir_w, mir_p, mir_n;
    int wb_stores;               // window kernel: plain write-back stores instead of streaming ones (pieces that do not fill 32-byte sectors
                                 // must wait in L2 for their neighbours, else DRAM read-modify-writes them)
    i64 ncol; i64 ninner;
    FinalArgs fin;
};

// Register-resident state of one sweep thread: partial sums of the (owner slot, partner slot)
// pairs alive on the current knot span, for every output component.
template <int P1, class T, int IS, bool FINAL>
struct SweepCore {
    static constexpr int NOUT = T::NOUT, NIN = T::NIN, NT = T::NT;
    double acc[IS][P1][NOUT];
    OwnerCache oc[FINAL ? IS : 1];
    i64 obase_m;                                  // output base of the mirrored column, or -1

    GSB_MEMBER void zero()
    {
#pragma unroll
        for (int is = 0; is < (FINAL ? IS : 1); ++is) { oc[is].fun = -1; oc[is].rec = 0; }
        obase_m = -1;
#pragma unroll
        for (int is = 0; is < IS; ++is)
#pragma unroll
            for (int js = 0; js < P1; ++js)
#pragma unroll
                for (int o = 0; o < NOUT; ++o) acc[is][js][o] = 0.0;
    }

    // one quadrature point: v = the NIN input components, tb = slot-ordered (value, derivative) table row
    template <bool TBS = false>   // TBS: the table row is in shared memory (plain loads) instead of global/L1
    GSB_MEMBER void point(const double (&v)[NIN], const double2 *tb, int grp)
    {
        double2 bj[P1];
#pragma unroll
        for (int js = 0; js < P1; ++js) bj[js] = TBS ? tb[js] : ld_keep2(tb + js);
#pragma unroll
        for (int is = 0; is < IS; ++is) {
            double2 bi;
            if (IS == P1) bi = bj[is]; else bi = TBS ? tb[grp * IS + is] : ld_keep2(tb + grp * IS + is);
            // z[o][b] = sum over the terms with that (o,b) of B^(a)_owner * in_c
            double z[NOUT][2];
            static_for<0, NT>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const double bo = T::a(k) ? bi.y : bi.x;
                if constexpr (T::first(k)) z[T::o(k)][T::b(k)] = bo * v[T::c(k)];
                else z[T::o(k)][T::b(k)] = fma(bo, v[T::c(k)], z[T::o(k)][T::b(k)]);
            });
#pragma unroll
            for (int js = 0; js < P1; ++js)
                static_for<0, NOUT>([&](auto oc) {
                    constexpr int o = decltype(oc)::value;
                    if constexpr (T::has(o, 0)) acc[is][js][o] = fma(bj[js].x, z[o][0], acc[is][js][o]);
                    if constexpr (T::has(o, 1)) acc[is][js][o] = fma(bj[js].y, z[o][1], acc[is][js][o]);
                });
        }
    }

    GSB_MEMBER void emit(const SweepArgs &A, const FinalCtx &fc, i64 obase, int is, int js, int fi, int d)
    {
        if (FINAL) {
            OwnerCache &c = oc[FINAL ? is : 0];
            if (c.fun != fi) owner_load(A.fin, fc, fi, c);      // only at a segment start
            const i64 pos = final_prepare(A.fin, fc, c, d, acc[is][js][0]);
            if (pos >= 0) A.fin.values[pos] = acc[is][js][0];
        } else {
            if (A.half_out && d < 0) return;      // the mirrored pair (partner, -d) carries this value
            const i64 o0 = (i64)fi * A.out_fs + (i64)(d + A.d_off) * A.out_ds + obase;
#pragma unroll
            for (int o = 0; o < NOUT; ++o) A.out[o * A.out_cs + o0] = acc[is][js][o];
            if (obase_m >= 0) {                   // same number, roles of owner and partner exchanged
                const i64 om = (i64)(fi + d) * A.out_fs + (i64)(A.d_off - d) * A.out_ds + obase_m;
                static_for<0, NOUT>([&](auto oc_) {
                    constexpr int o = decltype(oc_)::value;
                    A.out[T::omirror(o) * A.out_cs + om] = acc[is][js][o];
                });
            }
        }
    }

    // functions leaving the span window after element e complete their pairs; a pair is emitted
    // by the segment that owns its OWNER function (x_min <= owner < x_max)
    GSB_MEMBER void exits(const SweepArgs &A, const FinalCtx &fc, i64 obase, int nx, int f0, int grp, int x_min, int x_max)
    {
        const int ph = f0 % P1;
        for (int k = 0; k < nx; ++k) {
            const int x = f0 + k;
            if (x >= x_max) break;                 // later exits only involve owners >= x_max
            const int sx = (ph + k) % P1;
#pragma unroll
            for (int is = 0; is < IS; ++is) {
                const int so = grp * IS + is;
                const int fi = f0 + ((so - ph + P1) % P1);
                const bool wr = (fi >= x_min) && (fi < x_max);
                if (fi == x) {
#pragma unroll
                    for (int js = 0; js < P1; ++js) {
                        const int fj = f0 + ((js - ph + P1) % P1);
                        if (wr && fj >= x) emit(A, fc, obase, is, js, fi, fj - fi);
#pragma unroll
                        for (int o = 0; o < NOUT; ++o) acc[is][js][o] = 0.0;
                    }
                    // the function that takes over this slot is known now: fetch its column data early
                    if (FINAL) { const int fn = x + P1; if (fn >= x_min && fn < x_max) owner_load(A.fin, fc, fn, oc[FINAL ? is : 0]); else oc[FINAL ? is : 0].fun = -1; }
                } else if (fi > x) {
#pragma unroll
                    for (int js = 0; js < P1; ++js)
                        if (js == sx) {
                            if (wr) emit(A, fc, obase, is, js, fi, x - fi);
#pragma unroll
                            for (int o = 0; o < NOUT; ++o) acc[is][js][o] = 0.0;
                        }
                }
            }
        }
    }
};

// Output base of a thread's column; *mirror receives the base of the column with owner and partner exchanged in
// the previous direction ((i0,d0) -> (i0+d0,-d0)) when the sweep has to write mirrored values, else -1.
GSB_DEVICE i64 sweep_obase(const SweepArgs &A, i64 outer, i64 inner, i64 *mirror)
{
    const i64 i0 = outer / A.out_od, d0 = outer % A.out_od;
    const i64 in_part = (inner / A.out_bq) * A.out_bs + (inner % A.out_bq) * A.out_is;
    *mirror = (A.mirror && d0 > 0 && i0 + d0 < A.out_nprev) ? (i0 + d0) * A.out_os + (A.out_dshift - d0) * A.out_os2 + in_part : -1;
    return i0 * A.out_os + (d0 + A.out_dshift) * A.out_os2 + in_part;
}

// Generic variant: inputs straight from global memory (any strides / alignment).
template <int P1, class T, int IS, bool FINAL>
GSB_GLOBAL void k_sweep(const SweepArgs A)
{
    const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= A.ncol) return;
    const int grp = blockIdx.y, sg = blockIdx.z;
    const int e_begin = A.seg[4 * sg + 0], e_end = A.seg[4 * sg + 1], x_min = A.seg[4 * sg + 2], x_max = A.seg[4 * sg + 3];
    const i64 outer = col / A.ninner, inner = col - outer * A.ninner;
    const double *inp = A.in + outer * A.in_os + inner * A.in_is;
    FinalCtx fc;
    i64 obase = 0, obase_mirror = -1;
    if (FINAL) { if (!final_init(A.fin, outer, inner, fc)) return; }
    else obase = sweep_obase(A, outer, inner, &obase_mirror);
    constexpr int NIN = T::NIN;
    SweepCore<P1, T, IS, FINAL> core;
    core.zero();
    core.obase_m = obase_mirror;
    const int q = A.q;
    for (int e = e_begin; e < e_end; ++e) {
        const int f0 = A.first[e];
        const double *ine = inp + (i64)(e - A.e_in0) * A.in_es;
        const double2 *tbe = A.tab + (i64)e * q * P1;
        for (int t = 0; t < q; ++t) {
            double v[NIN];
#pragma unroll
            for (int c = 0; c < NIN; ++c) v[c] = ld_keep(ine + c * A.in_cs + t * A.in_ts);
            core.point(v, tbe + t * P1, grp);
        }
        core.exits(A, fc, obase, A.nexit[e], f0, grp, x_min, x_max);
    }
}

// ------------------------------------------------------------------------------------
// Window variant of the sweep (default).  A thread owns one column of the non-swept index space AND one
// group of output components (OMASK), and keeps the partial sums of ALL (p+1)^2 function pairs alive on
// the current knot span in WINDOW order: acc[a][b] belongs to the pair (f0+a, f0+b), f0 = first function
// of the span.  When f0 leaves the window its 2p+1 pairs are complete, are written once, and the window
// shifts by one (register renaming after unrolling, no slot arithmetic, no cross-thread sharing, no
// barriers).  The inputs of span e+1 are requested while span e is being integrated (register double
// buffer) and lines further ahead are pulled into L2 with prefetch hints, so the HBM latency hides behind
// the FP64 work without a shared-memory ring.  Knot multiplicities > 1 simply exit several functions.
#ifndef GSB200_EMULATE
#define GSB_NOINLINE static __device__ __noinline__
GSB_DEVICE void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
#define GSB_NOINLINE static
static inline void prefetch_l2(const void *) {}
#endif

// boundary-adjacent, coupled and eliminated entries of the final scatter: out of line, the hot loop stays small
GSB_NOINLINE void final_slow(const FinalArgs &F, const FinalCtx &c, int fun, i64 rec, int dL, double val)
{
    OwnerCache oc; oc.fun = fun; oc.rec = rec;
    const i64 pos = final_prepare(F, c, oc, dL, val);
    if (pos >= 0) F.values[pos] = val;
}

// canonical column whose stencil is cut by eliminated functions: slot from the per-run (start, mask) word; an
// eliminated partner goes the slow way (right-hand-side contribution)
GSB_DEVICE void final_canonical(const FinalArgs &F, const FinalCtx &c, int fun, i64 rec, int dL, int run, double val)
{
    const unsigned w = F.st[((i64)F.brow * F.nb + (i64)fun * c.nlow + c.li_low) * F.nrun + run];
    const unsigned mask = w >> 16;
    if ((mask >> c.bit0) & 1u) { st_stream(F.values + (rec >> 2) + (int)(w & 0xffffu) + popc(mask & ((1u << c.bit0) - 1u)), val); return; }
    if (F.st2) {
        const unsigned w2 = F.st2[((i64)F.brow * F.nb + (i64)fun * c.nlow + c.li_low) * F.nrun + run];
        const unsigned mask2 = w2 >> 16;
        if ((mask2 >> c.bit0) & 1u) { st_stream(F.values + (rec >> 2) + (int)(w2 & 0xffffu) + popc(mask2 & ((1u << c.bit0) - 1u)), val); return; }
    }
    if (F.fixed) final_slow(F, c, fun, rec, dL, val);
}


#ifndef GSB200_EMULATE
GSB_DEVICE void cp_async8(double *dst_smem, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
GSB_DEVICE void cp_async16(double2 *dst_smem, const double2 *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
GSB_DEVICE void warp_sync() { __syncwarp(); }
GSB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> GSB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#else
static inline void cp_async8(double *dst, const double *src) { *dst = *src; }
static inline void warp_sync() {}
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
#endif
// shared memory per CTA of 128 threads and resident CTAs per SM the window kernel is built for
template <int P1, class T, unsigned OMASK, int NS> GSB_CX int window_smem() { return NS * P1 * used_count<T>(OMASK) * 128 * 8 + NS * 4 * P1 * P1 * 16; }
template <int P1, class T, unsigned OMASK, int NS> GSB_CX int window_minb() { if (NS == 0) return 3; int n = 220 * 1024 / window_smem<P1, T, OMASK, NS>(); return n > 4 ? 4 : (n < 1 ? 1 : n); }

#ifndef GSB_WINDOW_HOLD
#define GSB_WINDOW_HOLD(P1_) (((P1_) * 8) % 32 != 0)
#endif
template <int P1, class T, unsigned OMASK, bool FINAL, int NS>
GSB_GLOBAL void
#ifndef GSB200_EMULATE
__launch_bounds__(128, (window_minb<P1, T, OMASK, NS>()))
#endif
k_sweepw(const GSB_GRID_CONSTANT SweepArgs A)
{
    constexpr int NIN = T::NIN, NT = T::NT, NOUT = T::NOUT, NG = mask_count(OMASK), NQ = P1;
    // every thread of a warp stays in the loop (the warp shares the staged basis table); threads without a column
    // work on a clamped one and own nothing
    const i64 col_raw = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = col_raw < A.ncol;
    const i64 col = live ? col_raw : A.ncol - 1;
    const int sg = blockIdx.z;
    const int e_begin = A.seg[4 * sg + 0], e_end = A.seg[4 * sg + 1];
    const i64 outer = col / A.ninner, inner = col - outer * A.ninner;
    i64 in_part = inner * A.in_is;
    if (A.in_bq > 0) { const i64 r = inner % A.in_bq2; in_part = (inner / A.in_bq2) * A.in_bs2 + (r / A.in_bq) * A.in_bs + (r % A.in_bq) * A.in_is; }
    const double *inp = A.in + outer * A.in_os + in_part - (i64)A.e_in0 * A.in_es;
    const double *inp_m = inp; bool mir_neg = false;      // same column of the mirrored pair; delta < 0 ?
    if (A.mir_w > 0) {
        const int di = (int)((inner / A.mir_q) % A.mir_w);
        const i64 om = outer + di - A.mir_p;
        mir_neg = di < A.mir_p;
        if (om >= 0 && om < A.mir_n) inp_m = A.in + om * A.in_os + in_part + (i64)(2 * (A.mir_p - di)) * A.mir_q * A.in_is - (i64)A.e_in0 * A.in_es;
        else mir_neg = false;
    }
    FinalCtx fc;
    i64 obase = 0, unused_mirror = -1;
    i64 fin_c0 = 0, fin_ww = 0; int fin_pl = 0, fin_w1 = 1;
    if (FINAL) {
        live = final_init(A.fin, outer, inner, fc) && live;
        const int W0 = 2 * A.fin.p[0] + 1;
        if (A.fin.dim == 2) { fin_ww = W0; fin_c0 = fc.bit0; fin_pl = A.fin.p[1]; fin_w1 = 1; }
        else { fin_ww = (i64)(2 * A.fin.p[1] + 1) * W0; fin_c0 = (i64)fc.r_low * W0 + fc.bit0; fin_pl = A.fin.p[2]; fin_w1 = 2 * A.fin.p[1] + 1; }
        fin_c0 += (i64)A.fin.brow * (2 * fin_pl + 1) * fin_ww;        // rows of a vector-valued space: component-major blocks of the stencil
    } else obase = sweep_obase(A, outer, inner, &unused_mirror);
    const int x_min = live ? A.seg[4 * sg + 2] : 0, x_max = live ? A.seg[4 * sg + 3] : 0;    // empty owner range: nothing is written

    double acc[P1][P1][NG];
#pragma unroll
    for (int a = 0; a < P1; ++a)
#pragma unroll
        for (int b = 0; b < P1; ++b)
#pragma unroll
            for (int g = 0; g < NG; ++g) acc[a][b][g] = 0.0;
    // HOLD: completed pairs with a negative delta wait for their owner's exit (hold[k][j]: window position k, delta -j) so that a
    // whole row of 2p+1 deltas is stored at once.  Needed when a span's q points do not fill 32-byte sectors: pieces stored one
    // span apart outlive the L2 and are read-modify-written in DRAM (profiles/r01b_layout_experiments.txt)
    constexpr bool HOLD = GSB_WINDOW_HOLD(P1) || FINAL || T::whole_rows();      // the final scatter always writes whole columns of deltas: one owner record per exit
    double hold[HOLD ? P1 : 1][HOLD ? P1 : 1][NG];
#pragma unroll
    for (int k = 0; k < (HOLD ? P1 : 1); ++k)
#pragma unroll
        for (int j = 0; j < (HOLD ? P1 : 1); ++j)
#pragma unroll
            for (int g = 0; g < NG; ++g) hold[k][j][g] = 0.0;
    int f0 = A.first[e_begin];
    i64 rec[P1];      // FINAL: packed (colptr << 2 | flag) of the window's functions as owners, 0 = not ours
#pragma unroll
    for (int k = 0; k < P1; ++k) {
        rec[k] = 0;
        if (FINAL) { const int fn = f0 + k; if (fn >= x_min && fn < x_max) rec[k] = A.fin.ownrec[A.fin.bcol * A.fin.nb + (i64)fn * fc.nlow + fc.li_low]; }
    }
    // thread-constant output strides of the two families of completed pairs
    const i64 st_b = A.out_ds, st_a = A.out_fs - A.out_ds, base_d = (i64)A.d_off * A.out_ds + obase;

    // NS >= 2: thread-private ring in shared memory filled by cp.async (LDGSTS): [stage][row = (t, used component)][tid];
    // NS == 0: register double buffer (the loads of span e+1 are in flight while span e is integrated)
    constexpr int NINg = used_count<T>(OMASK), NR = NQ * NINg, NSR = NS > 0 ? NS : 1;
#ifndef GSB200_EMULATE
    extern __shared__ __align__(16) double ring_smem[];
    double *ring = ring_smem + threadIdx.x;
#define GSB_RING(s_, r_) ring[((s_) * NR + (r_)) * 128]
#else
    double ring[NSR * NR];
#define GSB_RING(s_, r_) ring[(s_) * NR + (r_)]
#endif
    // the span's basis table (NQ x P1 value/derivative pairs, local order) travels with the stage: one copy per warp
    constexpr int NTAB = NQ * P1;
#if !defined(GSB200_EMULATE)
    constexpr bool STAGE_TAB = NS > 0 && NTAB <= 32;
    double2 *tabring = reinterpret_cast<double2 *>(ring_smem + (size_t)NSR * NR * 128) + (threadIdx.x >> 5) * NTAB;   // [stage][warp][NTAB]
    const int lane = threadIdx.x & 31;
#else
    constexpr bool STAGE_TAB = false;
#endif
    double vc[NS == 0 ? NR : 1], vn[NS == 0 ? NR : 1];
    const double *pin[NIN];       // per (virtual) input component: where this thread's column starts
    static_for<0, NIN>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        if constexpr (T::uses_c(OMASK, c)) {
            constexpr int mode = T::in_mode(c);
            const bool at_mirror = mode == 2 || (mode == 0 && mir_neg);
            pin[c] = at_mirror ? inp_m + T::in_srcm(c) * A.in_cs : inp + T::in_src(c) * A.in_cs;
        }
    });
    auto issue = [&](int e, int st) {
        const i64 pe = (i64)e * A.in_es;
#if !defined(GSB200_EMULATE)
        if constexpr (STAGE_TAB) { if (lane < NTAB) cp_async16(tabring + (size_t)st * 4 * NTAB + lane, A.tabl + (i64)e * NTAB + lane); }
#endif
#pragma unroll
        for (int t = 0; t < NQ; ++t)
            static_for<0, NIN>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                if constexpr (T::uses_c(OMASK, c)) {
                    constexpr int r = used_rank<T>(OMASK, c);
                    if constexpr (NS == 0) vn[t * NINg + r] = ld_stream(pin[c] + pe + t * A.in_ts);
                    else cp_async8(&GSB_RING(st, t * NINg + r), pin[c] + pe + t * A.in_ts);
                }
            });
    };
    if constexpr (NS == 0) {
        issue(e_begin, 0);
#pragma unroll
        for (int r = 0; r < NR; ++r) vc[r] = vn[r];
    } else {
#pragma unroll
        for (int k = 0; k < NS - 1; ++k) { if (e_begin + k < e_end) issue(e_begin + k, k); cp_async_commit(); }
    }
    int stg = 0;
    const int pf = A.pf_dist;
    for (int e = e_begin; e < e_end; ++e) {
        const int nx = A.nexit[e];
        if constexpr (NS == 0) { if (e + 1 < e_end) issue(e + 1, 0); }
        else {
            const int sn = stg == 0 ? NS - 1 : stg - 1;
            if constexpr (STAGE_TAB) warp_sync();      // every lane is done with the stage about to be refilled
            if (e + NS - 1 < e_end) issue(e + NS - 1, sn);
            cp_async_commit();
            cp_async_wait<(NS > 0 ? NS - 1 : 0)>();
            if constexpr (STAGE_TAB) warp_sync();      // ... and sees the table entries the other lanes copied
        }
        if (pf > 0 && e + pf < e_end) {
            const i64 pe = (i64)(e + pf) * A.in_es;
#pragma unroll
            for (int t = 0; t < NQ; ++t)
                static_for<0, NIN>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    if constexpr (T::uses_c(OMASK, c)) prefetch_l2(pin[c] + pe + t * A.in_ts);
                });
        }
        const double2 *tb = A.tabl + (i64)e * NQ * P1;
#if !defined(GSB200_EMULATE)
        const double2 *tbs = tabring + (size_t)stg * 4 * NTAB;
#endif
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
            double2 bw[P1];
#pragma unroll
            for (int k = 0; k < P1; ++k) {
#if !defined(GSB200_EMULATE)
                if constexpr (STAGE_TAB) bw[k] = tbs[t * P1 + k]; else
#endif
                bw[k] = ld_keep2(tb + t * P1 + k);
            }
            double v[NIN];
            static_for<0, NIN>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                if constexpr (T::uses_c(OMASK, c)) {
                    if constexpr (NS == 0) v[c] = vc[t * NINg + used_rank<T>(OMASK, c)];
                    else v[c] = GSB_RING(stg, t * NINg + used_rank<T>(OMASK, c));
                }
            });
#pragma unroll
            for (int a = 0; a < P1; ++a) {
                double z[NOUT][2];
                static_for<0, NT>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    if constexpr ((OMASK >> T::o(k)) & 1u) {
                        const double bo = T::a(k) ? bw[a].y : bw[a].x;
                        if constexpr (T::first(k)) z[T::o(k)][T::b(k)] = bo * v[T::c(k)];
                        else z[T::o(k)][T::b(k)] = fma(bo, v[T::c(k)], z[T::o(k)][T::b(k)]);
                    }
                });
#pragma unroll
                for (int b = 0; b < P1; ++b)
                    static_for<0, NOUT>([&](auto oc_) {
                        constexpr int o = decltype(oc_)::value;
                        if constexpr ((OMASK >> o) & 1u) {
                            constexpr int g = mask_rank(OMASK, o);
                            if constexpr (T::has(o, 0)) acc[a][b][g] = fma(bw[b].x, z[o][0], acc[a][b][g]);
                            if constexpr (T::has(o, 1)) acc[a][b][g] = fma(bw[b].y, z[o][1], acc[a][b][g]);
                        }
                    });
            }
        }
        // exits: f0 leaves the window, its pairs are complete
        for (int x = 0; x < nx; ++x) {
            if constexpr (HOLD) {
                // whole-row emission: the pairs (f0, f0-j) completed at earlier exits wait in hold[0][j]; together with (f0, f0+b)
                // they are the owner's complete row of 2p+1 deltas, written in one go (contiguous across the warp's lanes)
                if (FINAL) {
                    if (rec[0]) {
                        const int flag = (int)(rec[0] & 3);
                        if (flag == 3) {
#pragma unroll
                            for (int j = 1; j < P1; ++j) st_stream(A.fin.values + (rec[0] >> 2) + (i64)(fin_pl - j) * fin_ww + fin_c0, hold[0][j][0]);
#pragma unroll
                            for (int b = 0; b < P1; ++b) st_stream(A.fin.values + (rec[0] >> 2) + (i64)(b + fin_pl) * fin_ww + fin_c0, acc[0][b][0]);
                        } else {
                            const int dlo = A.fin.plo[A.fin.L][f0] - f0, dhi = A.fin.phi[A.fin.L][f0] - f0;     // co-occurring partners only
#pragma unroll
                            for (int j = 1; j < P1; ++j)
                                if (-j >= dlo) {
                                    if (flag == 1) final_canonical(A.fin, fc, f0, rec[0], -j, (fin_pl - j) * fin_w1 + fc.r_low, hold[0][j][0]);
                                    else final_slow(A.fin, fc, f0, rec[0], -j, hold[0][j][0]);
                                }
#pragma unroll
                            for (int b = 0; b < P1; ++b)
                                if (b <= dhi) {
                                    if (flag == 1) final_canonical(A.fin, fc, f0, rec[0], b, (b + fin_pl) * fin_w1 + fc.r_low, acc[0][b][0]);
                                    else final_slow(A.fin, fc, f0, rec[0], b, acc[0][b][0]);
                                }
                        }
                    }
                } else if (f0 >= x_min && f0 < x_max) {
                    const i64 o0 = (i64)f0 * A.out_fs + base_d;
                    static_for<0, NOUT>([&](auto oc_) {
                        constexpr int o = decltype(oc_)::value;
                        if constexpr ((OMASK >> o) & 1u) {
                            constexpr int g = mask_rank(OMASK, o);
#pragma unroll
                            for (int j = P1 - 1; j >= 1; --j) st_out(A.out + o * A.out_cs + o0 - j * st_b, hold[0][j][g], A.wb_stores);
#pragma unroll
                            for (int b = 0; b < P1; ++b) st_out(A.out + o * A.out_cs + o0 + b * st_b, acc[0][b][g], A.wb_stores);
                        }
                    });
                }
                // (f0+a, f0) is complete now: it waits for its owner; then everything moves one window position down
#pragma unroll
                for (int k = 1; k < P1; ++k)
#pragma unroll
                    for (int g = 0; g < NG; ++g) hold[k][k][g] = acc[k][0][g];
#pragma unroll
                for (int k = 0; k < P1; ++k)
#pragma unroll
                    for (int j = 1; j < P1; ++j)
#pragma unroll
                        for (int g = 0; g < NG; ++g) hold[k][j][g] = (k + 1 < P1) ? hold[(k + 1) % P1][j][g] : 0.0;
            } else
            if (FINAL) {
                if (rec[0]) {          // owner f0, partners f0+b
#pragma unroll
                    for (int b = 0; b < P1; ++b) {
                        if (b + x >= P1) break;        // x-th exit of one span: f0+b entered later, the pair never co-occurs
                        const int flag = (int)(rec[0] & 3);
                        if (flag == 3) st_stream(A.fin.values + (rec[0] >> 2) + (i64)(b + fin_pl) * fin_ww + fin_c0, acc[0][b][0]);
                        else if (flag == 1) final_canonical(A.fin, fc, f0, rec[0], b, (b + fin_pl) * fin_w1 + fc.r_low, acc[0][b][0]);
                        else final_slow(A.fin, fc, f0, rec[0], b, acc[0][b][0]);
                    }
                }
#pragma unroll
                for (int a = 1; a < P1; ++a) {   // owner f0+a, partner f0
                    if (rec[a] && a + x < P1) {
                        const int flag = (int)(rec[a] & 3);
                        if (flag == 3) st_stream(A.fin.values + (rec[a] >> 2) + (i64)(fin_pl - a) * fin_ww + fin_c0, acc[a][0][0]);
                        else if (flag == 1) final_canonical(A.fin, fc, f0 + a, rec[a], -a, (fin_pl - a) * fin_w1 + fc.r_low, acc[a][0][0]);
                        else final_slow(A.fin, fc, f0 + a, rec[a], -a, acc[a][0][0]);
                    }
                }
            } else {
                const i64 o0 = (i64)f0 * A.out_fs + base_d;
                if (f0 >= x_min && f0 < x_max) {
#pragma unroll
                    for (int b = 0; b < P1; ++b)
                        if (b + x < P1)
                            static_for<0, NOUT>([&](auto oc_) {
                                constexpr int o = decltype(oc_)::value;
                                if constexpr ((OMASK >> o) & 1u) st_out(A.out + o * A.out_cs + o0 + b * st_b, acc[0][b][mask_rank(OMASK, o)], A.wb_stores);
                            });
                }
#pragma unroll
                for (int a = 1; a < P1; ++a) {
                    if (f0 + a >= x_min && f0 + a < x_max && a + x < P1)
                        static_for<0, NOUT>([&](auto oc_) {
                            constexpr int o = decltype(oc_)::value;
                            if constexpr (((OMASK >> o) & 1u) && !T::out_sym(o)) st_out(A.out + o * A.out_cs + o0 + a * st_a, acc[a][0][mask_rank(OMASK, o)], A.wb_stores);
                        });
                }
            }
            // shift the window by one function
#pragma unroll
            for (int a = 0; a < P1; ++a)
#pragma unroll
                for (int b = 0; b < P1; ++b)
#pragma unroll
                    for (int g = 0; g < NG; ++g) acc[a][b][g] = (a + 1 < P1 && b + 1 < P1) ? acc[(a + 1) % P1][(b + 1) % P1][g] : 0.0;
            ++f0;
            if (FINAL) {
#pragma unroll
                for (int k = 0; k + 1 < P1; ++k) rec[k] = rec[k + 1];
                const int fn = f0 + P1 - 1;
                rec[P1 - 1] = (fn >= x_min && fn < x_max) ? A.fin.ownrec[A.fin.bcol * A.fin.nb + (i64)fn * fc.nlow + fc.li_low] : 0;
            }
        }
        if constexpr (NS == 0) {
#pragma unroll
            for (int r = 0; r < NR; ++r) vc[r] = vn[r];
        } else stg = (stg + 1 == NS) ? 0 : stg + 1;
    }
#undef GSB_RING
}

#include "fused23.cuh"    // second + last sweep of 3-D forms in one kernel (A2 stays in shared memory)

// ------------------------------------------------------------------------------------
// K3: load vector, one direction at a time: out[i][col] = sum_{q in supp(i)} B_i(q) in[q][col].
struct VSweepArgs {
    const int *ffirst, *flast;    // per function: first/last element of its support
    const int *first; const double2 *tab; int q, p1;
    int x_lo, x_hi;               // functions produced
    int e_in0;                    // first element present in `in`
    const double *in; i64 in_qs, in_os, in_is;    // point stride, column strides
    double *out; i64 out_fs, out_os, out_is;      // function stride, column strides
    i64 ncol, ninner;
    // final: scatter into rhs through the dof map
    int final_; int n0, n1, dimlow; const int *dofmap; double *rhs; int nfree;
};
GSB_GLOBAL void k_vsweep(const VSweepArgs A)
{
    const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= A.ncol) return;
    const int x = A.x_lo + blockIdx.y;
    if (x >= A.x_hi) return;
    const i64 outer = col / A.ninner, inner = col - outer * A.ninner;
    const double *inp = A.in + outer * A.in_os + inner * A.in_is;
    double s = 0.0;
    const int slot = x % A.p1;
    for (int e = A.ffirst[x]; e <= A.flast[x]; ++e)
        for (int t = 0; t < A.q; ++t)
            s = fma(A.tab[((i64)e * A.q + t) * A.p1 + slot].x, inp[(i64)((e - A.e_in0) * A.q + t) * A.in_qs], s);
    if (!A.final_) { A.out[(i64)x * A.out_fs + outer * A.out_os + inner * A.out_is] = s; return; }
    // column = (i1, i0) [3-D] or i0 [2-D]; local index = (x*n1 + i1)*n0 + i0
    const i64 li = (A.dimlow == 2) ? ((i64)x * A.n1 + outer) * A.n0 + inner : (i64)x * A.n0 + inner;
    const int g = A.dofmap[li];
    if (g < A.nfree) atomic_add(A.rhs + g, s);
}

// Window variant of the load-vector sweep: one thread per column marches along the swept direction with the (p+1)
// partial sums of the functions alive on the current span (window order); the function leaving the window is complete.
// Reads the load density once (the kernel above reads every point once per overlapping function).
template <int P1>
GSB_GLOBAL void k_vsweepw(const VSweepArgs A, const double2 *tabl, const int *nexit, const int *seg)
{
    constexpr int NQ = P1;
    const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= A.ncol) return;
    const int sg = blockIdx.y;
    const int e_begin = seg[4 * sg + 0], e_end = seg[4 * sg + 1], x_min = seg[4 * sg + 2], x_max = seg[4 * sg + 3];
    const i64 outer = col / A.ninner, inner = col - outer * A.ninner;
    const double *inp = A.in + outer * A.in_os + inner * A.in_is - (i64)A.e_in0 * NQ * A.in_qs;
    double acc[P1];
#pragma unroll
    for (int k = 0; k < P1; ++k) acc[k] = 0.0;
    int f0 = A.first[e_begin];
    double vc[NQ], vn[NQ];
#pragma unroll
    for (int t = 0; t < NQ; ++t) vc[t] = ld_stream(inp + (i64)(e_begin * NQ + t) * A.in_qs);
    for (int e = e_begin; e < e_end; ++e) {
        const int nx = nexit[e];
        if (e + 1 < e_end) {
#pragma unroll
            for (int t = 0; t < NQ; ++t) vn[t] = ld_stream(inp + (i64)((e + 1) * NQ + t) * A.in_qs);
        }
        const double2 *tb = tabl + (i64)e * NQ * P1;
#pragma unroll
        for (int t = 0; t < NQ; ++t)
#pragma unroll
            for (int k = 0; k < P1; ++k) acc[k] = fma(ld_keep2(tb + t * P1 + k).x, vc[t], acc[k]);
        for (int x = 0; x < nx; ++x) {
            if (f0 >= x_min && f0 < x_max) {
                if (!A.final_) A.out[(i64)f0 * A.out_fs + outer * A.out_os + inner * A.out_is] = acc[0];
                else {
                    const i64 li = (A.dimlow == 2) ? ((i64)f0 * A.n1 + outer) * A.n0 + inner : (i64)f0 * A.n0 + inner;
                    const int g = A.dofmap[li];
                    if (g < A.nfree) atomic_add(A.rhs + g, acc[0]);
                }
            }
#pragma unroll
            for (int k = 0; k + 1 < P1; ++k) acc[k] = acc[k + 1];
            acc[P1 - 1] = 0.0;
            ++f0;
        }
#pragma unroll
        for (int t = 0; t < NQ; ++t) vc[t] = vn[t];
    }
}

// ------------------------------------------------------------------------------------
// K1: sparsity pattern.  Column gi of the CSC matrix = union over the (patch, local fn)
// pre-images of gi of the free members of the tensor-product stencil of that function.
struct PatArgs {
    int dim, ncomp;
    int n[3], p[3];
    const int *plo[3]; const int *phi[3];
    const int *dofmap; i64 nb;
    int nfree;
    int own_lo, own_hi;            // owner range in the last direction (multi-rank slabs)
    const int *npre;               // per global dof: number of pre-images
    unsigned long long *len;       // per global column: entry count (upper bound for coupled columns)
    const i64 *colptr; int *inner; int *cursor;
    unsigned char *colflag; unsigned *st, *st2; int nrun;      // st2: slot words of the coupled rows (0: single patch)
    unsigned char *gneed;          // per global column: 1 sort, 2 sort+unique
};

GSB_DEVICE void pat_decode(const PatArgs &A, i64 li, int *i)
{
    for (int k = 0; k < A.dim; ++k) { i[k] = (int)(li % A.n[k]); li /= A.n[k]; }
}

GSB_GLOBAL void k_pat_preimages(const int *dofmap, i64 n, int nfree, int *npre)
{
    const i64 id = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const int g = dofmap[id];
    if (g < nfree) atomic_add(npre + g, 1);
}

// pass 0: count; pass 1: fill
// Body of the pattern kernels for one (component, local function).  PASS 1 with `stage` != 0 writes the row indices of a column that
// this thread owns alone into stage[0..n) (shared memory; flushed with coalesced stores by k_pattern_staged) and reports where they go.
template <int PASS>
GSB_DEVICE void pattern_column(const PatArgs &A, i64 id, int *stage, i64 *stage_base, int *stage_n)
{
    if (id >= A.nb * A.ncomp) return;
    const int cc = (int)(id / A.nb);
    const i64 li = id - (i64)cc * A.nb;
    const int gi = A.dofmap[id];
    if (PASS == 1) A.colflag[id] = 0;
    if (gi >= A.nfree) return;
    int i[3] = {0, 0, 0};
    pat_decode(A, li, i);
    const bool multi = A.npre[gi] > 1;
    // columns shared by several patches are patterned on every rank (their values are
    // all-reduced afterwards); all others only by the rank that owns the function
    if (!multi && (i[A.dim - 1] < A.own_lo || i[A.dim - 1] >= A.own_hi)) return;
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int k = 0; k < A.dim; ++k) { lo[k] = A.plo[k][i[k]]; hi[k] = A.phi[k][i[k]]; }
    if (PASS == 0) {
        unsigned long long cnt = 0;
        for (int cr = 0; cr < A.ncomp; ++cr)
            for (int j2 = lo[2]; j2 <= hi[2]; ++j2) for (int j1 = lo[1]; j1 <= hi[1]; ++j1) for (int j0 = lo[0]; j0 <= hi[0]; ++j0) {
                const i64 lj = ((i64)j2 * A.n[1] + j1) * A.n[0] + j0;
                if (A.dofmap[(i64)cr * A.nb + lj] < A.nfree) ++cnt;
            }
        if (multi) {
#ifdef GSB200_EMULATE
            A.len[gi] += cnt;
#else
            atomicAdd(A.len + gi, cnt);
#endif
        } else A.len[gi] = cnt;
        return;
    }
    // fill
    const i64 base = A.colptr[gi];
    if (multi) {
        // coupled column: append at an atomic cursor, sorted + deduplicated afterwards
        int cnt = 0;
        for (int cr = 0; cr < A.ncomp; ++cr)
            for (int j2 = lo[2]; j2 <= hi[2]; ++j2) for (int j1 = lo[1]; j1 <= hi[1]; ++j1) for (int j0 = lo[0]; j0 <= hi[0]; ++j0)
                if (A.dofmap[(i64)cr * A.nb + ((i64)j2 * A.n[1] + j1) * A.n[0] + j0] < A.nfree) ++cnt;
        i64 pos = base + atomic_add(A.cursor + gi, cnt);
        for (int cr = 0; cr < A.ncomp; ++cr)
            for (int j2 = lo[2]; j2 <= hi[2]; ++j2) for (int j1 = lo[1]; j1 <= hi[1]; ++j1) for (int j0 = lo[0]; j0 <= hi[0]; ++j0) {
                const int gj = A.dofmap[(i64)cr * A.nb + ((i64)j2 * A.n[1] + j1) * A.n[0] + j0];
                if (gj < A.nfree) A.inner[pos++] = gj;
            }
        A.colflag[id] = 2; A.gneed[gi] = 2;
        return;
    }
    // A column that lives in one patch.  Its rows come in two ascending sequences per row component: the standard DOFs in stencil order,
    // then the DOFs coupled with other patches (numbered after all standard ones, gsDofMapper.cpp:281-323) in stencil order.  They are
    // written in that order - which is the sorted order whenever the checks below hold - and every (row component, run) gets two
    // (start, mask) words, one per sequence, from which the final sweep takes its slots without searching.
    i64 pos = base;
    bool mono = true; int prev_block_last = -1;
    i64 full = 1;
    for (int k = 0; k < A.dim; ++k) full *= 2 * A.p[k] + 1;
    bool all_full = true;            // every row component contributes the whole (2p+1)^d stencil of standard DOFs
    for (int cr = 0; cr < A.ncomp; ++cr) {
        int nstd = 0, ncpl = 0;
        for (int j2 = lo[2]; j2 <= hi[2]; ++j2) for (int j1 = lo[1]; j1 <= hi[1]; ++j1) for (int j0 = lo[0]; j0 <= hi[0]; ++j0) {
            const int gj = A.dofmap[(i64)cr * A.nb + ((i64)j2 * A.n[1] + j1) * A.n[0] + j0];
            if (gj < A.nfree) { if (A.npre[gj] > 1) ++ncpl; else ++nstd; }
        }
        i64 ps = pos, pc = pos + nstd;
        int prev_s = prev_block_last, prev_c = -1;
        for (int j2 = lo[2]; j2 <= hi[2]; ++j2) for (int j1 = lo[1]; j1 <= hi[1]; ++j1) {
            unsigned mask = 0, mask2 = 0; const int start = (int)(ps - base), start2 = (int)(pc - base);
            for (int j0 = lo[0]; j0 <= hi[0]; ++j0) {
                const int gj = A.dofmap[(i64)cr * A.nb + ((i64)j2 * A.n[1] + j1) * A.n[0] + j0];
                if (gj >= A.nfree) continue;
                if (A.npre[gj] > 1) {
                    if (stage) stage[pc - base] = gj; else A.inner[pc] = gj;
                    ++pc;
                    if (gj <= prev_c) mono = false;
                    prev_c = gj;
                    mask2 |= 1u << (j0 - i[0] + A.p[0]);
                } else {
                    if (stage) stage[ps - base] = gj; else A.inner[ps] = gj;
                    ++ps;
                    if (gj <= prev_s) mono = false;
                    prev_s = gj;
                    mask |= 1u << (j0 - i[0] + A.p[0]);
                }
            }
            const int run = (A.dim == 2) ? (j1 - i[1] + A.p[1]) : ((j2 - i[2] + A.p[2]) * (2 * A.p[1] + 1) + (j1 - i[1] + A.p[1]));
            // (the rows of a column do not depend on the column's own component: the columns (cc, li) of a vector-valued space write the same words)
            A.st[((i64)cr * A.nb + li) * A.nrun + run] = (unsigned)start | (mask << 16);
            if (A.st2) A.st2[((i64)cr * A.nb + li) * A.nrun + run] = (unsigned)start2 | (mask2 << 16);
            else if (mask2) mono = false;
        }
        // the first coupled row must follow the last standard one, and this block the previous one
        if (ncpl > 0 && nstd > 0) { const int first_c = stage ? stage[pos + nstd - base] : A.inner[pos + nstd]; if (first_c <= prev_s) mono = false; }
        prev_block_last = ncpl > 0 ? prev_c : prev_s;
        if (nstd != full || ncpl != 0) all_full = false;
        pos += nstd + ncpl;
    }
    // 3: whole (2p+1)^d stencil of standard DOFs (per row component, component-major rows) -> closed-form slots; 1: canonical column
    // (slots from the (start, mask) words); 2: anything else (rows sorted afterwards, binary search in the final sweep)
    if (mono && all_full) A.colflag[id] = 3;
    else if (mono) A.colflag[id] = 1;
    else { A.colflag[id] = 2; A.gneed[gi] = 1; }
    if (stage) { *stage_base = base; *stage_n = (int)(pos - base); }
}

template <int PASS>
GSB_GLOBAL void k_pattern(const PatArgs A)
{
    pattern_column<PASS>(A, (i64)blockIdx.x * blockDim.x + threadIdx.x, 0, 0, 0);
}

#ifndef GSB200_EMULATE
// Fill pass with coalesced stores: one warp per block, every lane builds its column in a private shared-memory row (odd stride:
// conflict-free), then the warp copies the 32 rows to their places in `inner` 128 bytes at a time.  The one-thread-per-column
// version above stores 4 bytes per lane into 32 different lines per instruction (7.6 ms for 2.6 GB at config 2).
GSB_GLOBAL void __launch_bounds__(32) k_pattern_staged(const PatArgs A, const int stride)
{
    extern __shared__ int pat_stage[];
    const int lane = threadIdx.x;
    i64 base = 0; int n = 0;
    pattern_column<1>(A, (i64)blockIdx.x * 32 + lane, pat_stage + lane * stride, &base, &n);
    __syncwarp();
    for (int c = 0; c < 32; ++c) {
        const i64 bc = __shfl_sync(0xffffffffu, base, c);
        const int nc = __shfl_sync(0xffffffffu, n, c);
        for (int k = lane; k < nc; k += 32) A.inner[bc + k] = pat_stage[c * stride + k];
    }
}
#endif

// sort (and for coupled columns deduplicate) the row indices of the flagged columns
GSB_GLOBAL void k_pat_sort(int ncols, const unsigned char *gneed, const i64 *colptr, int *inner, unsigned long long *len)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ncols || !gneed[g]) return;
    int *a = inner + colptr[g];
    const int n = (int)len[g];
    // Shell sort (Ciura gaps): a coupled column of an 8-patch vector-valued problem carries up to 8 x 3 x (2p+1)^3 appended entries,
    // plain insertion sort made the pattern build of config 4 take a second
    const int gaps[9] = {1750, 701, 301, 132, 57, 23, 10, 4, 1};
    for (int gi = 0; gi < 9; ++gi) {
        const int gap = gaps[gi];
        if (gap >= n) continue;
        for (int i = gap; i < n; ++i) { const int v = a[i]; int j = i - gap; while (j >= 0 && a[j] > v) { a[j + gap] = a[j]; j -= gap; } a[j + gap] = v; }
    }
    if (gneed[g] == 2) {
        int m = 0;
        for (int i = 0; i < n; ++i) if (i == 0 || a[i] != a[m - 1]) a[m++] = a[i];
        len[g] = (unsigned long long)m;
    }
}

GSB_GLOBAL void k_pat_compact(int ncols, const i64 *oldptr, const i64 *newptr, const int *oldinner, int *newinner)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ncols) return;
    const i64 n = newptr[g + 1] - newptr[g];
    for (i64 k = 0; k < n; ++k) newinner[newptr[g] + k] = oldinner[oldptr[g] + k];
}

// 64-bit device column pointers -> the 32-bit outerIndexPtr of gsSparseMatrix<T,0,index_t>
GSB_GLOBAL void k_narrow_outer(int n, const i64 *ptr, int *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)ptr[i];
}

// single-thread exclusive scan used by the interpreter build; the product build uses cub
GSB_GLOBAL void k_scan_serial(const unsigned long long *len, i64 *ptr, int n)
{
    if (blockIdx.x || threadIdx.x) return;
    i64 s = 0;
    for (int i = 0; i < n; ++i) { ptr[i] = s; s += (i64)len[i]; }
    ptr[n] = s;
}

// ------------------------------------------------------------------------------------
// Consumer kernels: y = A x on the CSC arrays read as CSR of the (symmetric) matrix.
GSB_GLOBAL void k_spmv(int n, const i64 *ptr, const int *idx, const double *val, const double *x, double *y)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double s = 0.0;
    for (i64 k = ptr[r]; k < ptr[r + 1]; ++k) s = fma(val[k], x[idx[k]], s);
    y[r] = s;
}
#ifndef GSB200_EMULATE
// One warp per row: the lanes stride over the row's entries (values and indices are read as whole 256/128-byte
// segments), partial sums meet in a shuffle tree.  HBM-bound: 12 B per stored entry.
GSB_GLOBAL void __launch_bounds__(256) k_spmv_warp(int n, const i64 *ptr, const int *idx, const double *val, const double *x, double *y)
{
    const int lane = threadIdx.x & 31;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 r = warp0; r < n; r += nwarp) {
        const i64 b = ptr[r], e = ptr[r + 1];
        double s = 0.0;
        for (i64 k = b + lane; k < e; k += 32) s = fma(__ldcs(val + k), __ldg(x + __ldcs(idx + k)), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[r] = s;
    }
}
#endif
GSB_GLOBAL void k_diag(int n, const i64 *ptr, const int *idx, const double *val, double *d)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double s = 1.0;
    for (i64 k = ptr[r]; k < ptr[r + 1]; ++k) if (idx[k] == r) s = val[k];
    d[r] = s;
}
GSB_GLOBAL void k_dot(int n, const double *a, const double *b, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    for (int k = i; k < n; k += gridDim.x * blockDim.x) s = fma(a[k], b[k], s);
#ifndef GSB200_EMULATE
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) != 0) return;
#endif
    if (s != 0.0) atomic_add(out, s);
}
// z = a + alpha * b  (element-wise)  and  z = a / d
GSB_GLOBAL void k_axpy(int n, const double *a, double alpha, const double *b, double *z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = fma(alpha, b[i], a[i]);
}
GSB_GLOBAL void k_div(int n, const double *a, const double *d, double *z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = a[i] / d[i];
}

} // namespace gsb
