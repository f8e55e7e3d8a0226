/* gsb_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference's (gismo/gismo v24.08.0) CPU
 * algorithm for the isogeometric system-assembly hot path, taking the same POD problem
 * description as the product's C ABI (include/gsb200.h).  It exists to CHECK the CUDA
 * path; nothing in gismo_b200/ may link, import or call it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py compares this file against
 * fixtures produced by the reference's own assemblers (oracle/ref_driver.cpp compiled
 * from /root/reference by oracle/Makefile into oracle/_ref/) — pattern bit-exact, values
 * and right-hand side to 1e-13 relative — and against the fingerprints SURVEY.md 8(c)
 * records (sum K_ij, ||rhs||_2).
 *
 * The algorithm is the element loop of the visitor path, deliberately NOT the
 * sum-factorised row-owner algorithm the CUDA kernels use, so the two are independent:
 *   element enumeration      gsTensorDomainIterator.h:69-118 (lexicographic, dir 0 fastest)
 *   Gauss rule + affine map  gsGaussRule.hpp:20-60, gsQuadRule.hpp:87-117, gsQuadRule.h:177-201
 *   1-D basis + derivative   gsBSplineBasis.hpp:863-1041 (NURBS-book A2.3)
 *   tensor products          gsTensorBasis.hpp:634-697
 *   actives                  gsTensorBSplineBasis.hpp:166-205
 *   geometry values/Jacobian gsGeometry.hpp:539-574, gsRationalBasis.h:481-520
 *   measure, inverse         gsFunction.hpp:702-751, gsMatrixAddons.h:56-85
 *   local stiffness / load   gsVisitorPoisson.h:89-106, gsAssembler.h:39-46
 *   elasticity blocks        linear_elasticity_example.cpp:183-190, gsExprAssembler.h:583-625
 *   scatter + elimination    gsSparseSystem.h:972-1010
 *   sorted sparse insertion  SparseMatrix.h:208-225,1366-1396 (result: compressed CSC)
 */
#include "../include/gsb200.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define MAXP 8          /* max degree supported by the oracle */
#define MAXQ 12         /* max Gauss points per direction      */

static char g_err[512];
const char *gsbo_last_error(void) { return g_err; }

/* ------------------------------------------------------------------ */
/* Gauss-Legendre nodes/weights on [-1,1].  The reference ships 30-digit
 * literal tables (gsGaussRule.hpp:218-547); we recompute them by Newton
 * iteration on P_n in long double and round to double (same doubles to
 * within 1 ulp; tests/test_oracle_golden.py pins this against the tables). */
int gsbo_gauss(int n, double *nodes, double *weights)
{
    if (n < 1 || n > 64) return -1;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double x = cosl(pi * (i + 0.75L) / (n + 0.5L)), dp = 0;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1, p1 = x;
            for (int k = 2; k <= n; ++k) {
                long double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
                p0 = p1; p1 = p2;
            }
            if (n == 1) { p0 = 1; p1 = x; }
            dp = n * (x * p1 - p0) / (x * x - 1);
            long double dx = p1 / dp;
            x -= dx;
            if (fabsl(dx) < 1e-19L) break;
        }
        { /* recompute derivative at the converged node */
            long double p0 = 1, p1 = x;
            for (int k = 2; k <= n; ++k) {
                long double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
                p0 = p1; p1 = p2;
            }
            dp = n * (x * p1 - p0) / (x * x - 1);
        }
        long double w = 2 / ((1 - x * x) * dp * dp);
        nodes[i] = (double)(-x); nodes[n - 1 - i] = (double)x;
        weights[i] = weights[n - 1 - i] = (double)w;
    }
    if (n % 2) nodes[n / 2] = 0.0;
    return 0;
}

/* number of Gauss points, gsQuadrature.h:152-171 */
static int num_nodes(double quA, int quB, int p) { return (int)(quA * p + quB + 0.5); }

/* ------------------------------------------------------------------ */
/* Knot-vector helpers: domain = [knots[p], knots[n-p-1]], elements =
 * non-empty spans between distinct knots inside it. */
typedef struct {
    int p, nk, nfun, nel;
    const double *kn;
    int *span;      /* per element: index s with kn[s] < kn[s+1], the knot span   */
} kv1d;

static int kv_init(kv1d *k, const double *kn, int nk, int p)
{
    k->p = p; k->nk = nk; k->kn = kn; k->nfun = nk - p - 1;
    if (k->nfun < 1 || p < 1 || p > MAXP) return -1;
    k->span = (int *)malloc(sizeof(int) * (size_t)nk);
    k->nel = 0;
    for (int s = p; s < nk - p - 1; ++s)
        if (kn[s] < kn[s + 1]) k->span[k->nel++] = s;
    return k->nel > 0 ? 0 : -1;
}

/* span of a point strictly inside the domain: upper_bound - 1
 * (gsKnotVector.hpp:747-783), right end closed. */
static int kv_find(const kv1d *k, double u)
{
    int lo = k->p, hi = k->nk - k->p - 1; /* kn[lo] <= u <= kn[hi] */
    if (u >= k->kn[hi]) { int s = hi - 1; while (k->kn[s] == k->kn[s + 1]) --s; return s; }
    while (hi - lo > 1) { int mid = (lo + hi) / 2; if (k->kn[mid] <= u) lo = mid; else hi = mid; }
    return lo;
}

/* values and first derivatives of the p+1 functions alive on span s at u
 * (A2.3 restricted to n=1; gsBSplineBasis.hpp:945-1040). */
static void bspline_ders(const double *kn, int p, int s, double u, double *val, double *der)
{
    double ndu[(MAXP + 1) * (MAXP + 1)], left[MAXP + 1], right[MAXP + 1];
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
        if (r >= 1) { const double a = 1.0 / ndu[p * p1 + r - 1]; d = a * ndu[(r - 1) * p1 + p - 1]; }
        if (r <= p - 1) { const double a = -1.0 / ndu[p * p1 + r]; d += a * ndu[r * p1 + p - 1]; }
        der[r] = d * (double)p;
    }
}

/* ------------------------------------------------------------------ */
/* growable sorted column (Eigen's uncompressed insertion, SparseMatrix.h:1366-1396) */
typedef struct { int n, cap; int *idx; double *val; } spcol;

static double *col_ref(spcol *c, int row)
{
    int lo = 0, hi = c->n;
    while (lo < hi) { int mid = (lo + hi) / 2; if (c->idx[mid] < row) lo = mid + 1; else hi = mid; }
    if (lo < c->n && c->idx[lo] == row) return &c->val[lo];
    if (c->n == c->cap) {
        c->cap = c->cap ? 2 * c->cap : 32;
        c->idx = (int *)realloc(c->idx, sizeof(int) * (size_t)c->cap);
        c->val = (double *)realloc(c->val, sizeof(double) * (size_t)c->cap);
    }
    memmove(c->idx + lo + 1, c->idx + lo, sizeof(int) * (size_t)(c->n - lo));
    memmove(c->val + lo + 1, c->val + lo, sizeof(double) * (size_t)(c->n - lo));
    c->idx[lo] = row; c->val[lo] = 0.0; c->n++;
    return &c->val[lo];
}

/* ------------------------------------------------------------------ */
/* source-term stack machine (independent of the product's evaluator) */
static double prog_eval(const gsb200_program *pr, const double *x)
{
    double st[GSB200_PROGRAM_MAX_STACK]; int sp = 0;
    for (int i = 0; i < pr->nops; ++i) {
        const int op = pr->ops[i];
        switch (op) {
        case GSB200_OP_CONST: st[sp++] = pr->consts[pr->ops[++i]]; break;
        case GSB200_OP_X: st[sp++] = x[0]; break;
        case GSB200_OP_Y: st[sp++] = x[1]; break;
        case GSB200_OP_Z: st[sp++] = x[2]; break;
        case GSB200_OP_ADD: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
        case GSB200_OP_SUB: sp--; st[sp - 1] = st[sp - 1] - st[sp]; break;
        case GSB200_OP_MUL: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
        case GSB200_OP_DIV: sp--; st[sp - 1] = st[sp - 1] / st[sp]; break;
        case GSB200_OP_POW: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
        case GSB200_OP_NEG: st[sp - 1] = -st[sp - 1]; break;
        case GSB200_OP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
        case GSB200_OP_COS: st[sp - 1] = cos(st[sp - 1]); break;
        case GSB200_OP_TAN: st[sp - 1] = tan(st[sp - 1]); break;
        case GSB200_OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
        case GSB200_OP_LOG: st[sp - 1] = log(st[sp - 1]); break;
        case GSB200_OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
        case GSB200_OP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
        case GSB200_OP_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
        case GSB200_OP_SINH: st[sp - 1] = sinh(st[sp - 1]); break;
        case GSB200_OP_COSH: st[sp - 1] = cosh(st[sp - 1]); break;
        case GSB200_OP_SQR: st[sp - 1] = st[sp - 1] * st[sp - 1]; break;
        case GSB200_OP_SINPI: st[sp - 1] = sin(3.14159265358979323846 * st[sp - 1]); break;
        case GSB200_OP_COSPI: st[sp - 1] = cos(3.14159265358979323846 * st[sp - 1]); break;
        default: return NAN;
        }
    }
    return sp == 1 ? st[0] : NAN;
}

/* d x d inverse by cofactors + determinant (gsMatrixAddons.h:56-85 does the same
 * through Eigen's fixed-size inverse()). M is row-major M[r*d+c]. */
static double inv_det(const double *M, int d, double *Inv)
{
    if (d == 2) {
        const double det = M[0] * M[3] - M[1] * M[2];
        Inv[0] = M[3] / det; Inv[1] = -M[1] / det; Inv[2] = -M[2] / det; Inv[3] = M[0] / det;
        return det;
    }
    const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
    const double id = 1.0 / det;
    Inv[0] = c00 * id; Inv[1] = (M[2] * M[7] - M[1] * M[8]) * id; Inv[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    Inv[3] = c01 * id; Inv[4] = (M[0] * M[8] - M[2] * M[6]) * id; Inv[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    Inv[6] = c02 * id; Inv[7] = (M[1] * M[6] - M[0] * M[7]) * id; Inv[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    return det;
}

/* ------------------------------------------------------------------ */
/* Assemble.  Pass outer/inner/values/rhs = NULL to query *nnz only.
 * Returns 0 on success.  The result is the compressed CSC triple + rhs. */
int gsbo_assemble(const gsb200_problem *pb, int64_t *nnz_out, int32_t *outer, int32_t *inner,
                  double *values, double *rhs)
{
    static spcol *cols = NULL; static int cols_n = 0; static double *rhs_acc = NULL;
    static const gsb200_problem *cached = NULL;
    if (!pb) { snprintf(g_err, sizeof g_err, "null problem"); return -1; }
    /* second call with buffers reuses the assembly done by the size query */
    if (!(cached == pb && cols && outer)) {
        if (cols) { for (int i = 0; i < cols_n; ++i) { free(cols[i].idx); free(cols[i].val); } free(cols); free(rhs_acc); cols = NULL; }
        const int d = pb->patches[0].space.dim, nc = pb->ncomp, N = pb->nfree, nrhs = pb->nrhs;
        if (d < 2 || d > 3) { snprintf(g_err, sizeof g_err, "dim %d unsupported", d); return -1; }
        cols = (spcol *)calloc((size_t)N, sizeof(spcol)); cols_n = N;
        rhs_acc = (double *)calloc((size_t)N * nrhs, sizeof(double));
        for (int ip = 0; ip < pb->npatches; ++ip) {
            const gsb200_patch *pa = &pb->patches[ip];
            kv1d ks[3], kg[3];
            int q[3], nq = 1, nact = 1, nactg = 1, nbasis = 1;
            double gn[3][MAXQ], gw[3][MAXQ];
            for (int k = 0; k < d; ++k) {
                if (kv_init(&ks[k], pa->space.knots[k], pa->space.nknots[k], pa->space.degree[k]) ||
                    kv_init(&kg[k], pa->geo.knots[k], pa->geo.nknots[k], pa->geo.degree[k])) {
                    snprintf(g_err, sizeof g_err, "bad knot vector patch %d dir %d", ip, k); return -1; }
                q[k] = num_nodes(pb->quA, pb->quB, ks[k].p);
                if (q[k] < 1 || q[k] > MAXQ) { snprintf(g_err, sizeof g_err, "bad quadrature size"); return -1; }
                gsbo_gauss(q[k], gn[k], gw[k]);
                nq *= q[k]; nact *= ks[k].p + 1; nactg *= kg[k].p + 1; nbasis *= ks[k].nfun;
            }
            const int nloc = nact * nc;
            double *bv = (double *)malloc(sizeof(double) * (size_t)nact * nq);          /* values n x Q           */
            double *bd = (double *)malloc(sizeof(double) * (size_t)nact * d * nq);      /* derivs (r*d+k) x Q     */
            double *pg = (double *)malloc(sizeof(double) * (size_t)nact * d);           /* physical grads d x n   */
            double *lm = (double *)malloc(sizeof(double) * (size_t)nloc * nloc);
            double *lr = (double *)malloc(sizeof(double) * (size_t)nloc * nrhs);
            int *act = (int *)malloc(sizeof(int) * (size_t)nact);
            double *gv = (double *)malloc(sizeof(double) * (size_t)nactg), *gd = (double *)malloc(sizeof(double) * (size_t)nactg * d);
            int el[3] = {0, 0, 0};
            int64_t qoff = 0; /* running offset into rhs_samples (tensor-ordered per element not used) */
            (void)qoff;
            for (;;) {
                /* ---- quadrature on this element: gsQuadRule.h:177-201 ---- */
                double lower[3], h[3], hprod = 1.0, u1[3][MAXQ];
                double v1[3][MAXQ][MAXP + 1], d1[3][MAXQ][MAXP + 1];
                for (int k = 0; k < d; ++k) {
                    const int s = ks[k].span[el[k]];
                    lower[k] = ks[k].kn[s];
                    h[k] = (ks[k].kn[s + 1] - lower[k]) / 2.0;
                    hprod *= (h[k] == 0.0 ? 0.5 : h[k]);
                    for (int t = 0; t < q[k]; ++t) {
                        u1[k][t] = h[k] * (gn[k][t] + 1.0) + lower[k];
                        bspline_ders(ks[k].kn, ks[k].p, s, u1[k][t], v1[k][t], d1[k][t]);
                    }
                }
                /* ---- actives: gsTensorBSplineBasis.hpp:166-205 ---- */
                {
                    int r = 0, a[3] = {0, 0, 0};
                    for (;;) {
                        int idx = 0;
                        for (int k = d - 1; k >= 0; --k) idx = idx * ks[k].nfun + (ks[k].span[el[k]] - ks[k].p + a[k]);
                        act[r++] = idx;
                        int k = 0;
                        while (k < d && ++a[k] > ks[k].p) { a[k] = 0; ++k; }
                        if (k == d) break;
                    }
                }
                memset(lm, 0, sizeof(double) * (size_t)nloc * nloc);
                memset(lr, 0, sizeof(double) * (size_t)nloc * nrhs);
                /* ---- loop over tensor quadrature points, direction 0 fastest ---- */
                int t[3] = {0, 0, 0};
                for (int kq = 0; kq < nq; ++kq) {
                    double u[3] = {0, 0, 0}, w = hprod;
                    /* weight: product in direction order (gsQuadRule.hpp:104-111), times hprod */
                    { double wp = gw[0][t[0]]; for (int k = 1; k < d; ++k) wp *= gw[k][t[k]]; w = hprod * wp; }
                    for (int k = 0; k < d; ++k) u[k] = u1[k][t[k]];
                    /* tensor basis values / derivatives (gsTensorBasis.hpp:666-696) */
                    {
                        int r = 0, a[3] = {0, 0, 0};
                        for (;;) {
                            double v = 1.0;
                            for (int k = 0; k < d; ++k) v *= v1[k][t[k]][a[k]];
                            bv[r] = v;
                            for (int k = 0; k < d; ++k) {
                                double dv = d1[k][t[k]][a[k]];
                                for (int i = 0; i < d; ++i) if (i != k) dv *= v1[i][t[i]][a[i]];
                                bd[r * d + k] = dv;
                            }
                            ++r;
                            int k = 0;
                            while (k < d && ++a[k] > ks[k].p) { a[k] = 0; ++k; }
                            if (k == d) break;
                        }
                    }
                    /* geometry: x and Jt[a][c] = d x_c / d xi_a (gsGeometry.hpp:557-564) */
                    double x[3] = {0, 0, 0}, Jt[9], J[9], Jinv[9];
                    {
                        int sg[3]; double gv1[3][MAXP + 1], gd1[3][MAXP + 1];
                        int ngeo = 1;
                        for (int k = 0; k < d; ++k) {
                            sg[k] = kv_find(&kg[k], u[k]);
                            bspline_ders(kg[k].kn, kg[k].p, sg[k], u[k], gv1[k], gd1[k]);
                            ngeo *= kg[k].nfun;
                        }
                        for (int i = 0; i < d * d; ++i) Jt[i] = 0.0;
                        double W = 0.0, dW[3] = {0, 0, 0}, xn[3] = {0, 0, 0}, dxn[9] = {0};
                        int a[3] = {0, 0, 0};
                        for (;;) {
                            int idx = 0;
                            for (int k = d - 1; k >= 0; --k) idx = idx * kg[k].nfun + (sg[k] - kg[k].p + a[k]);
                            double v = 1.0, dv[3];
                            for (int k = 0; k < d; ++k) v *= gv1[k][a[k]];
                            for (int k = 0; k < d; ++k) {
                                dv[k] = gd1[k][a[k]];
                                for (int i = 0; i < d; ++i) if (i != k) dv[k] *= gv1[i][a[i]];
                            }
                            const double wt = pa->geo_weights ? pa->geo_weights[idx] : 1.0;
                            W += wt * v;
                            for (int k = 0; k < d; ++k) dW[k] += wt * dv[k];
                            for (int c = 0; c < d; ++c) {
                                const double C = pa->geo_coefs[(size_t)c * ngeo + idx];
                                xn[c] += wt * v * C;
                                for (int k = 0; k < d; ++k) dxn[k * d + c] += wt * dv[k] * C;
                            }
                            int k = 0;
                            while (k < d && ++a[k] > kg[k].p) { a[k] = 0; ++k; }
                            if (k == d) break;
                        }
                        /* quotient rule (gsRationalBasis.h:481-520); W==1, dW==0 for B-splines */
                        for (int c = 0; c < d; ++c) {
                            x[c] = xn[c] / W;
                            for (int k = 0; k < d; ++k) Jt[k * d + c] = (dxn[k * d + c] * W - xn[c] * dW[k]) / (W * W);
                        }
                    }
                    for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) J[r * d + c] = Jt[c * d + r];
                    const double det = inv_det(J, d, Jinv);
                    const double weight = w * fabs(det);
                    /* physical gradients: J^{-T} * grad (gsAssembler.h:39-46) */
                    for (int r = 0; r < nact; ++r)
                        for (int c = 0; c < d; ++c) {
                            double s = 0.0;
                            for (int a = 0; a < d; ++a) s += Jinv[a * d + c] * bd[r * d + a];
                            pg[r * d + c] = s;
                        }
                    /* source term */
                    double fv[3 * 4] = {0};
                    const int nf = (pb->form == GSB200_FORM_ELASTICITY) ? nc : nrhs;
                    if (pb->rhs_kind == GSB200_RHS_PROGRAM)
                        for (int c = 0; c < nf; ++c) fv[c] = prog_eval(&pb->rhs_programs[c], x);
                    /* local load (gsVisitorPoisson.h:101) */
                    if (pb->form == GSB200_FORM_ELASTICITY) {
                        for (int c = 0; c < nc; ++c) for (int i = 0; i < nact; ++i) lr[c * nact + i] += weight * (bv[i] * fv[c]);
                    } else {
                        for (int c = 0; c < nrhs; ++c) for (int i = 0; i < nact; ++i) lr[c * nloc + i] += weight * (bv[i] * fv[c]);
                    }
                    /* local matrix */
                    if (pb->form == GSB200_FORM_POISSON) {
                        for (int j = 0; j < nact; ++j) for (int i = 0; i < nact; ++i) {
                            double s = 0.0;
                            for (int c = 0; c < d; ++c) s += pg[i * d + c] * pg[j * d + c];
                            lm[(size_t)j * nloc + i] += weight * s;
                        }
                    } else if (pb->form == GSB200_FORM_MASS) {
                        for (int j = 0; j < nact; ++j) for (int i = 0; i < nact; ++i)
                            lm[(size_t)j * nloc + i] += weight * (bv[i] * bv[j]);
                    } else { /* elasticity: lambda div div + mu (grad+grad^T):grad */
                        const double la = pb->coef[0], mu = pb->coef[1];
                        for (int cc = 0; cc < nc; ++cc) for (int j = 0; j < nact; ++j)
                            for (int rr = 0; rr < nc; ++rr) for (int i = 0; i < nact; ++i) {
                                double dot = 0.0;
                                if (rr == cc) for (int c = 0; c < d; ++c) dot += pg[i * d + c] * pg[j * d + c];
                                const double s = la * pg[i * d + rr] * pg[j * d + cc] + mu * (pg[i * d + cc] * pg[j * d + rr] + dot);
                                lm[(size_t)(cc * nact + j) * nloc + (rr * nact + i)] += weight * s;
                            }
                    }
                    { int k = 0; while (k < d && ++t[k] >= q[k]) { t[k] = 0; ++k; } }
                }
                /* ---- scatter with elimination: gsSparseSystem.h:972-1010,
                        vector layout gsExprAssembler.h:583-625 ---- */
                for (int rr = 0; rr < nc; ++rr) for (int i = 0; i < nact; ++i) {
                    const int ii = pa->dofmap[(size_t)rr * nbasis + act[i]];
                    if (ii >= N) continue;
                    for (int c = 0; c < nrhs; ++c) rhs_acc[(size_t)c * N + ii] += lr[c * nloc + rr * nact + i];
                    for (int cc = 0; cc < nc; ++cc) for (int j = 0; j < nact; ++j) {
                        const int jj = pa->dofmap[(size_t)cc * nbasis + act[j]];
                        const double v = lm[(size_t)(cc * nact + j) * nloc + (rr * nact + i)];
                        if (jj < N) *col_ref(&cols[jj], ii) += v;
                        else if (pb->fixed)
                            for (int c = 0; c < nrhs; ++c) rhs_acc[(size_t)c * N + ii] -= v * pb->fixed[(size_t)c * pb->nfixed + (jj - N)];
                    }
                }
                int k = 0;
                while (k < d && ++el[k] >= ks[k].nel) { el[k] = 0; ++k; }
                if (k == d) break;
            }
            /* ---- Neumann sides of this patch: gsVisitorNeumann.h:83-136 (scalar data * |n|) or the
                    expression u*g_N.tr()*nv(G) (vector data . outer normal), boundary quadrature with one
                    node in the fixed direction (gsGaussRule.hpp:28-36, gsQuadRule.h:190-197), outer normal
                    from the first minors of the Jacobian (gsFunction.hpp:613-699), pushToRhs
                    (gsSparseSystem.h:884-899) ---- */
            for (int in = 0; in < pb->nneumann; ++in) {
                const gsb200_neumann *nm = &pb->neumann[in];
                if (nm->patch != ip) continue;
                const int dir = (nm->side - 1) / 2, upper = (nm->side - 1) % 2;
                const double ub = upper ? ks[dir].kn[ks[dir].nk - ks[dir].p - 1] : ks[dir].kn[ks[dir].p];
                const int eb = upper ? ks[dir].nel - 1 : 0;
                int el2[3] = {0, 0, 0};
                el2[dir] = eb;
                for (;;) {
                    double lower[3], h[3], hprod = 1.0, u1[3][MAXQ], v1[3][MAXQ][MAXP + 1], d1[3][MAXQ][MAXP + 1];
                    int qq[3];
                    for (int k = 0; k < d; ++k) {
                        const int s = ks[k].span[el2[k]];
                        qq[k] = (k == dir) ? 1 : q[k];
                        if (k == dir) { h[k] = 0.0; hprod *= 0.5; u1[k][0] = ub; bspline_ders(ks[k].kn, ks[k].p, s, ub, v1[k][0], d1[k][0]); continue; }
                        lower[k] = ks[k].kn[s]; h[k] = (ks[k].kn[s + 1] - lower[k]) / 2.0; hprod *= h[k];
                        for (int t = 0; t < q[k]; ++t) { u1[k][t] = h[k] * (gn[k][t] + 1.0) + lower[k]; bspline_ders(ks[k].kn, ks[k].p, s, u1[k][t], v1[k][t], d1[k][t]); }
                    }
                    { int r = 0, a[3] = {0, 0, 0};
                      for (;;) { int idx = 0; for (int k = d - 1; k >= 0; --k) idx = idx * ks[k].nfun + (ks[k].span[el2[k]] - ks[k].p + a[k]); act[r++] = idx;
                                 int k = 0; while (k < d && ++a[k] > ks[k].p) { a[k] = 0; ++k; } if (k == d) break; } }
                    memset(lr, 0, sizeof(double) * (size_t)nloc * nrhs);
                    int t[3] = {0, 0, 0};
                    const int nqb = qq[0] * qq[1] * (d == 3 ? qq[2] : 1);
                    for (int kq = 0; kq < nqb; ++kq) {
                        double u[3] = {0, 0, 0}, wp = 1.0;
                        for (int k = 0; k < d; ++k) { u[k] = u1[k][t[k]]; const double wk = (k == dir) ? 2.0 : gw[k][t[k]]; wp = (k == 0) ? wk : wp * wk; }
                        const double w = hprod * wp;
                        { int r = 0, a[3] = {0, 0, 0};
                          for (;;) { double v = 1.0; for (int k = 0; k < d; ++k) v *= v1[k][t[k]][a[k]]; bv[r++] = v;
                                     int k = 0; while (k < d && ++a[k] > ks[k].p) { a[k] = 0; ++k; } if (k == d) break; } }
                        double x[3] = {0, 0, 0}, Jt[9];
                        { int sg[3]; double gv1[3][MAXP + 1], gd1[3][MAXP + 1]; int ngeo = 1;
                          for (int k = 0; k < d; ++k) { sg[k] = kv_find(&kg[k], u[k]); bspline_ders(kg[k].kn, kg[k].p, sg[k], u[k], gv1[k], gd1[k]); ngeo *= kg[k].nfun; }
                          double W = 0.0, dW[3] = {0, 0, 0}, xn[3] = {0, 0, 0}, dxn[9] = {0};
                          int a[3] = {0, 0, 0};
                          for (;;) {
                              int idx = 0; for (int k = d - 1; k >= 0; --k) idx = idx * kg[k].nfun + (sg[k] - kg[k].p + a[k]);
                              double v = 1.0, dv[3]; for (int k = 0; k < d; ++k) v *= gv1[k][a[k]];
                              for (int k = 0; k < d; ++k) { dv[k] = gd1[k][a[k]]; for (int i = 0; i < d; ++i) if (i != k) dv[k] *= gv1[i][a[i]]; }
                              const double wt = pa->geo_weights ? pa->geo_weights[idx] : 1.0;
                              W += wt * v; for (int k = 0; k < d; ++k) dW[k] += wt * dv[k];
                              for (int c = 0; c < d; ++c) { const double C = pa->geo_coefs[(size_t)c * ngeo + idx]; xn[c] += wt * v * C; for (int k = 0; k < d; ++k) dxn[k * d + c] += wt * dv[k] * C; }
                              int k = 0; while (k < d && ++a[k] > kg[k].p) { a[k] = 0; ++k; } if (k == d) break;
                          }
                          for (int c = 0; c < d; ++c) { x[c] = xn[c] / W; for (int k = 0; k < d; ++k) Jt[k * d + c] = (dxn[k * d + c] * W - xn[c] * dW[k]) / (W * W); } }
                        /* outer normal: n_i = sgn * det_sgn * (-1)^i * det(first minor (dir, i) of Jt) */
                        double nrm[3] = {0, 0, 0}, detJ;
                        if (d == 2) {
                            detJ = Jt[0] * Jt[3] - Jt[1] * Jt[2];
                            const int o = 1 - dir;
                            nrm[0] = Jt[o * 2 + 1]; nrm[1] = -Jt[o * 2 + 0];
                        } else {
                            detJ = Jt[0] * (Jt[4] * Jt[8] - Jt[5] * Jt[7]) - Jt[1] * (Jt[3] * Jt[8] - Jt[5] * Jt[6]) + Jt[2] * (Jt[3] * Jt[7] - Jt[4] * Jt[6]);
                            const int r0 = dir == 0 ? 1 : 0, r1 = dir == 2 ? 1 : 2;
                            const double *A = &Jt[r0 * 3], *B = &Jt[r1 * 3];
                            nrm[0] = A[1] * B[2] - A[2] * B[1]; nrm[1] = -(A[0] * B[2] - A[2] * B[0]); nrm[2] = A[0] * B[1] - A[1] * B[0];
                        }
                        /* sideOrientation(s) = ((s + (s+1)/2) % 2) ? +1 : -1   (gsBoundary.h:1029-1035) */
                        const double sgn = (((nm->side + (nm->side + 1) / 2) % 2) ? 1.0 : -1.0) * (detJ < 0 ? -1.0 : 1.0);
                        double nn = 0.0; for (int c = 0; c < d; ++c) { nrm[c] *= sgn; nn += nrm[c] * nrm[c]; }
                        nn = sqrt(nn);
                        double flux;
                        if (nm->ndata == 1) flux = prog_eval(&nm->data[0], x) * nn;
                        else { flux = 0.0; for (int c = 0; c < d; ++c) flux += prog_eval(&nm->data[c], x) * nrm[c]; }
                        for (int i = 0; i < nact; ++i) lr[i] += w * bv[i] * flux;
                        { int k = 0; while (k < d && ++t[k] >= qq[k]) { t[k] = 0; ++k; } }
                    }
                    for (int i = 0; i < nact; ++i) { const int ii = pa->dofmap[act[i]]; if (ii < N) rhs_acc[ii] += lr[i]; }
                    int k = 0;
                    for (;;) { if (k == dir) { ++k; continue; } if (k >= d) break; if (++el2[k] < ks[k].nel) break; el2[k] = 0; ++k; }
                    if (k >= d) break;
                }
            }
            free(bv); free(bd); free(pg); free(lm); free(lr); free(act); free(gv); free(gd);
            for (int k = 0; k < d; ++k) { free(ks[k].span); free(kg[k].span); }
        }
        cached = pb;
    }
    int64_t nnz = 0;
    for (int i = 0; i < cols_n; ++i) nnz += cols[i].n;
    if (nnz_out) *nnz_out = nnz;
    if (outer && inner && values) {
        int64_t o = 0;
        for (int i = 0; i < cols_n; ++i) {
            outer[i] = (int32_t)o;
            memcpy(inner + o, cols[i].idx, sizeof(int) * (size_t)cols[i].n);
            memcpy(values + o, cols[i].val, sizeof(double) * (size_t)cols[i].n);
            o += cols[i].n;
        }
        outer[cols_n] = (int32_t)o;
        if (rhs) memcpy(rhs, rhs_acc, sizeof(double) * (size_t)cols_n * pb->nrhs);
        for (int i = 0; i < cols_n; ++i) { free(cols[i].idx); free(cols[i].val); }
        free(cols); free(rhs_acc); cols = NULL; rhs_acc = NULL; cached = NULL;
    }
    return 0;
}

/* 1-D evaluation exposed for tests (partition of unity, derivative sums) */
int gsbo_basis_eval(const double *knots, int nknots, int p, double u, int *first, double *val, double *der)
{
    kv1d k;
    if (kv_init(&k, knots, nknots, p)) return -1;
    const int s = kv_find(&k, u);
    bspline_ders(knots, p, s, u, val, der);
    *first = s - p;
    free(k.span);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Norm integrals of a discrete scalar field (test infrastructure, like everything in this file): element loop with the
 * assembly's Gauss rule, restating ev.integral((u_ex - u_sol).sqNorm() * meas(G)) and
 * ev.integral((igrad(u_ex) - igrad(u_sol, G)).sqNorm() * meas(G)) of examples/poisson2_example.cpp:174-177
 * (gsExprEvaluator.h:152-230: the element-wise quadrature loop; gsExpressions.h solution / igrad evaluation).
 * out4 = { int (u_h-u_ex)^2, int |grad(u_h-u_ex)|^2, int u_h^2, int |grad u_h|^2 }; exact / exact_grad may be NULL. */
int gsbo_field_norms(const gsb200_problem *pb, const double *u_free, const gsb200_program *exact,
                     const gsb200_program *exact_grad, double *out4)
{
    if (!pb || !u_free || !out4 || pb->ncomp != 1) { snprintf(g_err, sizeof g_err, "field norms: bad arguments"); return -1; }
    const int d = pb->patches[0].space.dim, N = pb->nfree;
    for (int k = 0; k < 4; ++k) out4[k] = 0.0;
    for (int ip = 0; ip < pb->npatches; ++ip) {
        const gsb200_patch *pa = &pb->patches[ip];
        kv1d ks[3], kg[3];
        int q[3], ngeo = 1;
        double gn[3][MAXQ], gw[3][MAXQ];
        for (int k = 0; k < d; ++k) {
            if (kv_init(&ks[k], pa->space.knots[k], pa->space.nknots[k], pa->space.degree[k]) ||
                kv_init(&kg[k], pa->geo.knots[k], pa->geo.nknots[k], pa->geo.degree[k])) { snprintf(g_err, sizeof g_err, "bad knot vector"); return -1; }
            q[k] = num_nodes(pb->quA, pb->quB, ks[k].p);
            gsbo_gauss(q[k], gn[k], gw[k]);
            ngeo *= kg[k].nfun;
        }
        int el[3] = {0, 0, 0};
        for (;;) {                                   /* elements */
            int t[3] = {0, 0, 0};
            for (;;) {                               /* quadrature points of the element */
                double u[3], w = 1.0, sv[3][MAXP + 1], sd[3][MAXP + 1], gv[3][MAXP + 1], gd[3][MAXP + 1];
                int s[3], sg[3];
                for (int k = 0; k < d; ++k) {
                    s[k] = ks[k].span[el[k]];
                    const double lo = ks[k].kn[s[k]], h = (ks[k].kn[s[k] + 1] - lo) / 2.0;
                    u[k] = h * (gn[k][t[k]] + 1.0) + lo; w *= h * gw[k][t[k]];
                    bspline_ders(ks[k].kn, ks[k].p, s[k], u[k], sv[k], sd[k]);
                    sg[k] = kv_find(&kg[k], u[k]);
                    bspline_ders(kg[k].kn, kg[k].p, sg[k], u[k], gv[k], gd[k]);
                }
                /* geometry */
                double W = 0.0, dW[3] = {0, 0, 0}, xn[3] = {0, 0, 0}, dxn[9] = {0}, x[3] = {0, 0, 0}, Jt[9], J[9], Jinv[9];
                int a[3] = {0, 0, 0};
                for (;;) {
                    int idx = 0;
                    for (int k = d - 1; k >= 0; --k) idx = idx * kg[k].nfun + (sg[k] - kg[k].p + a[k]);
                    double v = 1.0, dv[3];
                    for (int k = 0; k < d; ++k) v *= gv[k][a[k]];
                    for (int k = 0; k < d; ++k) { dv[k] = gd[k][a[k]]; for (int i = 0; i < d; ++i) if (i != k) dv[k] *= gv[i][a[i]]; }
                    const double wt = pa->geo_weights ? pa->geo_weights[idx] : 1.0;
                    W += wt * v;
                    for (int k = 0; k < d; ++k) dW[k] += wt * dv[k];
                    for (int c = 0; c < d; ++c) {
                        const double C = pa->geo_coefs[(size_t)c * ngeo + idx];
                        xn[c] += wt * v * C;
                        for (int k = 0; k < d; ++k) dxn[k * d + c] += wt * dv[k] * C;
                    }
                    int k = 0;
                    while (k < d && ++a[k] > kg[k].p) { a[k] = 0; ++k; }
                    if (k == d) break;
                }
                for (int c = 0; c < d; ++c) { x[c] = xn[c] / W; for (int k = 0; k < d; ++k) Jt[k * d + c] = (dxn[k * d + c] * W - xn[c] * dW[k]) / (W * W); }
                for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) J[r * d + c] = Jt[c * d + r];
                const double det = inv_det(J, d, Jinv);
                /* discrete field */
                double uh = 0.0, du[3] = {0, 0, 0};
                a[0] = a[1] = a[2] = 0;
                for (;;) {
                    size_t li = 0;
                    for (int k = d - 1; k >= 0; --k) li = li * (size_t)ks[k].nfun + (size_t)(s[k] - ks[k].p + a[k]);
                    double v = 1.0, dv[3];
                    for (int k = 0; k < d; ++k) v *= sv[k][a[k]];
                    for (int k = 0; k < d; ++k) { dv[k] = sd[k][a[k]]; for (int i = 0; i < d; ++i) if (i != k) dv[k] *= sv[i][a[i]]; }
                    const int g = pa->dofmap[li];
                    const double cf = g < N ? u_free[g] : (pb->fixed ? pb->fixed[g - N] : 0.0);
                    uh += cf * v;
                    for (int k = 0; k < d; ++k) du[k] += cf * dv[k];
                    int k = 0;
                    while (k < d && ++a[k] > ks[k].p) { a[k] = 0; ++k; }
                    if (k == d) break;
                }
                double gr[3] = {0, 0, 0};                 /* J^{-T} du */
                for (int c = 0; c < d; ++c) for (int k = 0; k < d; ++k) gr[c] += Jinv[k * d + c] * du[k];
                const double weight = w * fabs(det);
                const double ue = exact ? prog_eval(exact, x) : 0.0;
                double ge2 = 0.0, g2 = 0.0;
                for (int c = 0; c < d; ++c) {
                    const double gx = exact_grad ? prog_eval(&exact_grad[c], x) : 0.0;
                    ge2 += (gr[c] - gx) * (gr[c] - gx); g2 += gr[c] * gr[c];
                }
                out4[0] += weight * (uh - ue) * (uh - ue); out4[1] += weight * ge2; out4[2] += weight * uh * uh; out4[3] += weight * g2;
                int k = 0;
                while (k < d && ++t[k] >= q[k]) { t[k] = 0; ++k; }
                if (k == d) break;
            }
            int k = 0;
            while (k < d && ++el[k] >= ks[k].nel) { el[k] = 0; ++k; }
            if (k == d) break;
        }
        for (int k = 0; k < d; ++k) { free(ks[k].span); free(kg[k].span); }
    }
    if (!exact) out4[0] = out4[2];
    if (!exact_grad) out4[1] = exact ? -1.0 : out4[3];
    return 0;
}
