// fused.cuh — K0 + first sweep in ONE kernel: the coefficient tensor D (and the load density F) of a knot span
// is evaluated into shared memory by the CTA and consumed from there, so it never exists in HBM.
// Included by kernels.cuh (inside namespace gsb) AND embedded as text for NVRTC together with terms.cuh and
// geometry.cuh (jit.cuh): self-contained, no standard headers.
//
// Replaces, per patch chunk: k_geometry_line (a6/a7/a8/a18: gsGeometry.hpp:539-597, gsFunction.hpp:702-751,
// gsQuadRule.h:177-201, gsFunctionExpr.hpp:513-533) + the first k_sweepw launches (a10/a11: contraction of direction 0)
// + the first load-vector sweep.  12.9 GB of HBM traffic per assembly at config 2 (D written + re-read, F likewise).
//
// CTA = GSB_FUSE_TC columns of the non-swept index space (consecutive points of the LAST direction at fixed middle-direction
// point), warp-specialised:
//   producer warps (p+1 of them, one per direction-0 point of a span; lane = column): Jacobian, inverse, measure, weight, D
//       components, w|J|f of their point -> shared-memory tile.  The leading directions of the geometry are contracted once per
//       patch into line coefficients (k_line_coefs: they do not depend on the last direction) and read through L1.
//   consumer warps (one per output group, NG outputs each): (p+1)^2 * NG accumulators per thread in SLOT order (slot = function
//       index mod p+1: no window shifting), whole rows of 2p+1 deltas stored at the owner's exit; every first-sweep output has
//       exactly one term (component c, derivative flags a, b), so all groups run the SAME code with the term as data; group 0 also
//       carries the load vector.
// Tiles go round a ring of NB buffers with full/empty mbarriers: no CTA-wide barrier in the loop, producers run ahead.
#define GSB_FUSE_TC 32
#ifndef GSB_FUSE_PREG
#define GSB_FUSE_PREG 104     // registers per producer / consumer thread after the split (setmaxnreg; sum * 128 threads * 2 CTAs = register file)
#define GSB_FUSE_CREG 152
#endif
#define GSB_FUSE_MAXF 3       // load-vector components carried along

struct FusedArgs {
    GeoArgs G;                                          // D / F pointers unused (tiles in shared memory)
    const int *first, *nexit; const double2 *tab;       // direction 0: first function / exits per span, basis table [e][t][slot] (slot = function index mod p+1)
    const double *lc; int lc_nL;                        // line coefficients [q1][q0][a_L][field][kind] of the patch (k_line_coefs), functions of the last direction
    const int *seg;                                     // [gridDim.z][4] e_begin, e_end, x_min, x_max
    int ncolL, nrows;                                   // columns along the last direction (a tile never straddles a row); rows (3-D: Q1, 2-D: 1)
    double *out; i64 out_cs, out_fs, out_bq, out_bs, out_is; int d_off;      // A1 addressing (as SweepArgs); the delta stride is p+1 (blocked layouts)
    i64 out_ds;                                         // ... or, ROWS layout A1[o][i0][d0][q1][q2]: the (runtime) stride of a delta row
    int half_rows;                                      // blocked layouts: store delta >= 0 only (the next sweep reads delta < 0 at the mirrored pair, T3SymS2U)
    double *v1; i64 v1_cs, v1_fs; int nf;               // first load-vector sweep: V1[c][i0][column]
};

#ifndef GSB200_EMULATE
#define GSB_GRID_CONSTANT __grid_constant__
#define GSB_THREADS(tid_) for (int tid_ = (int)threadIdx.x, gsb_once_ = 1; gsb_once_; gsb_once_ = 0)
// split-phase CTA barrier (mbarrier in shared memory): every thread arrives once per phase and waits for the phase later
GSB_DEVICE void fbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
GSB_DEVICE void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
GSB_DEVICE void fbar_arrive(unsigned long long *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory"); }
GSB_DEVICE void fbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok, ns = 32;
    for (;;) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(ns);            // a spinning warp would take issue slots from the working ones
        if (ns < 256) ns *= 2;
    }
}
#else
#define GSB_GRID_CONSTANT
#define GSB_THREADS(tid_) for (int tid_ = 0; tid_ < (int)blockDim.x; ++tid_)
static inline void prefetch_l1(const void *) {}
static inline void fbar_init(unsigned long long *, unsigned) {}
static inline void fbar_arrive(unsigned long long *) {}
static inline void fbar_wait(unsigned long long *, unsigned) {}
#endif

// Line coefficients of a patch: E[q1][q0][a_L][field][kind] = sum_{a_0(,a_1)} B^(kind)(q_0(,q_1)) C_field[a_0(,a_1), a_L], kind = value,
// d/dxi_0 (, d/dxi_1); fields = the geoDim coordinates (times the weight for a rational geometry) and the weight.  One thread per
// (q1, q0, a_L, field).  Same contraction as the first half of geometry_line_body.
#ifndef GSB_JIT_SOURCE
struct LineCoefArgs {
    int dim, rational; int Q0, Q1, nL;
    const double2 *gtab[2]; const int *gfirst[2]; int pg1[2], ngeo[2];
    const double *coefs; const double *weights; i64 ngeo_total;
    double *lc;
};
GSB_GLOBAL void k_line_coefs(const LineCoefArgs A)
{
    const int nfg = A.rational ? A.dim + 1 : A.dim;
    const i64 id = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 total = (i64)A.Q1 * A.Q0 * A.nL * nfg;
    if (id >= total) return;
    const int f = (int)(id % nfg); i64 r = id / nfg;
    const int aL = (int)(r % A.nL); r /= A.nL;
    const int q0 = (int)(r % A.Q0), q1 = (int)(r / A.Q0);
    const int gf0 = A.gfirst[0][q0], gf1 = A.dim == 3 ? A.gfirst[1][q1] : 0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const int n1 = A.dim == 3 ? A.pg1[1] : 1;
    for (int a1 = 0; a1 < n1; ++a1) {
        const double2 b1 = A.dim == 3 ? A.gtab[1][(i64)q1 * A.pg1[1] + a1] : make_double2(1.0, 0.0);
        const i64 rowi = A.dim == 3 ? ((i64)aL * A.ngeo[1] + (gf1 + a1)) * A.ngeo[0] + gf0 : (i64)aL * A.ngeo[0] + gf0;
        for (int a0 = 0; a0 < A.pg1[0]; ++a0) {
            const double2 b0 = A.gtab[0][(i64)q0 * A.pg1[0] + a0];
            const i64 idx = rowi + a0;
            double C = f < A.dim ? A.coefs[(i64)f * A.ngeo_total + idx] : 1.0;
            if (A.rational) C *= A.weights[idx];
            s0 = fma(b0.x * b1.x, C, s0); s1 = fma(b0.y * b1.x, C, s1); s2 = fma(b0.x * b1.y, C, s2);
        }
    }
    double *o = A.lc + id * A.dim;
    o[0] = s0; o[1] = s1; if (A.dim == 3) o[2] = s2;
}
#endif

template <int P1, int NG>
struct FusedThread {             // consumer state
    double acc[P1][P1][NG];      // SLOT order: acc[sa][sb] belongs to the active functions with index = sa, sb (mod p+1); no shifting
    double hold[P1][P1][NG];     // hold[s][j]: completed pair (owner in slot s, delta -j) waiting for the owner's exit
    double *pw[NG];              // where the row of the NEXT exiting function goes, per output (advanced by the function stride)
    int pk[NG];                  // the output's term GSB_PK(o, c, a, b): component and the derivative flags of owner / partner
    int ngv;                     // outputs this thread really has (the last group may be short)
    int live;                    // column inside the patch?
    // producer state
    double2 bL[GSB_MAXP + 1];    // geometry basis of the last direction at the thread's column
    int qL, gfL;
};

// the single term of output o of a first-sweep table: GSB_PK(o, c, a, b)
template <class T> GSB_CX int fused_term_of(int o) { for (int k = 0; k < T::NT; ++k) if (T::o(k) == o) return T::pk(k); return -1; }
template <class T> GSB_CX bool fused_one_term_per_output()
{
    if (T::NT != T::NOUT) return false;
    for (int o = 0; o < T::NOUT; ++o) { int n = 0; for (int k = 0; k < T::NT; ++k) if (T::o(k) == o) ++n; if (n != 1) return false; }
    return true;
}
// shared memory of one instantiation with NB tile buffers; threads per CTA
template <int P1, class T> GSB_CX int fused_smem(int nb)
{
    return 8 * (nb * (T::NIN + GSB_FUSE_MAXF) * P1 * GSB_FUSE_TC + nb * 2 * P1 * P1 + GSB_FUSE_MAXF * P1 * GSB_FUSE_TC) + 16 * nb + 64;
}
template <int P1, class T, int NG> GSB_CX int fused_threads() { return GSB_FUSE_TC * (P1 + (T::NOUT + NG - 1) / NG); }

template <int DIM, int P1, class T, int NG, int PGL, bool RATIONAL, int FSPEC, bool ROWS = false>
GSB_DEVICE void geo_sweep_body(const FusedArgs &A)
{
    static_assert(fused_one_term_per_output<T>(), "first-sweep tables have one term per output");
    constexpr int L = DIM - 1, NFM = DIM + 1, NIN = T::NIN, NOUT = T::NOUT, NGRP = (NOUT + NG - 1) / NG, TC = GSB_FUSE_TC;
    constexpr int NCONS = TC * NGRP, NPROD = TC * P1, NTHR = NCONS + NPROD, NC = NIN + GSB_FUSE_MAXF, NPT = P1 * TC, nfg = RATIONAL ? DIM + 1 : DIM;
    constexpr int NB = fused_smem<P1, T>(4) <= 48 * 1024 ? 4 : (fused_smem<P1, T>(3) <= 48 * 1024 ? 3 : 2);      // ring depth
    constexpr int TILE = NC * NPT, DS = P1;                        // doubles per tile buffer; delta stride of A1
    constexpr bool FULLG = NOUT % NG == 0;                         // every group has NG outputs
    GSB_SHARED double Dt[NB * TILE];                               // [buffer][component][point][column]
    GSB_SHARED GSB_ALIGN16 double tsel[NB * 2 * P1 * P1];                      // the span's basis table: [buffer][values | derivatives][point][slot]
    GSB_SHARED double vacc[GSB_FUSE_MAXF][P1][GSB_FUSE_TC];      // load-vector accumulators of group 0, slot order (kept out of the register budget)
    GSB_SHARED unsigned long long full[NB], empty[NB];
    const GeoArgs &G = A.G;
    const int tiles = (A.ncolL + TC - 1) / TC;
    const int row = (int)(blockIdx.x / tiles), col0 = (int)(blockIdx.x % tiles) * TC;
    const int sg = blockIdx.z;
    const int e_begin = A.seg[4 * sg + 0], e_end = A.seg[4 * sg + 1], seg_xmin = A.seg[4 * sg + 2], seg_xmax = A.seg[4 * sg + 3];
    const int pgL = PGL ? PGL : G.pg1[L];
    const int ql1 = DIM == 3 ? row + G.qoff[1] : 0;
    const int nf = A.nf;
    const i64 lc_pt = (i64)A.lc_nL * (nfg * DIM);                  // doubles per (q1, q0) point of the patch's line coefficients
    const double *lc_row = A.lc + (i64)ql1 * G.qn[0] * lc_pt;
#ifdef GSB200_EMULATE
    static thread_local FusedThread<P1, NG> *gsb_states = 0; static thread_local int gsb_nstates = 0;
    if (gsb_nstates < NTHR) { delete[] gsb_states; gsb_states = new FusedThread<P1, NG>[NTHR]; gsb_nstates = NTHR; }
#define GSB_TH gsb_states[tid]
#else
    FusedThread<P1, NG> gsb_state;
#define GSB_TH gsb_state
    if (threadIdx.x == 0) { for (int b = 0; b < NB; ++b) { fbar_init(&full[b], NPROD); fbar_init(&empty[b], NCONS); } }
#endif

    // ---- per-thread set-up: threads [0, NCONS) are consumers (group = warp), the rest producers (direction-0 point = warp)
    GSB_THREADS(tid) {
        FusedThread<P1, NG> &th = GSB_TH;
        const int lane = tid % TC;
        const int col = col0 + lane;
        const bool live = col < A.ncolL;
        const int colc = live ? col : A.ncolL - 1;
        th.live = live ? 1 : 0;
        if (tid < NCONS) {
            const int grp = tid / TC;
            const i64 inner = (i64)row * A.ncolL + colc;
            const i64 obase = (inner / A.out_bq) * A.out_bs + (inner % A.out_bq) * A.out_is + (i64)A.d_off * (ROWS ? A.out_ds : (i64)DS) + (i64)A.first[e_begin] * A.out_fs;
            th.ngv = 0;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int i = grp * NG + g;
                th.pk[g] = fused_term_of<T>(T::order(i < NOUT ? i : 0));
                th.pw[g] = A.out + (i64)(th.pk[g] >> 8) * A.out_cs + obase;
                if (i < NOUT) th.ngv = g + 1;
            }
#pragma unroll
            for (int a = 0; a < P1; ++a)
#pragma unroll
                for (int b = 0; b < P1; ++b)
#pragma unroll
                    for (int g = 0; g < NG; ++g) { th.acc[a][b][g] = 0.0; th.hold[a][b][g] = 0.0; }
            if (grp == 0)
                for (int c = 0; c < GSB_FUSE_MAXF; ++c)
                    for (int k = 0; k < P1; ++k) vacc[c][k][lane] = 0.0;
        } else {
            th.qL = colc + G.qoff[L];
            th.gfL = G.gfirst[L][th.qL];
#pragma unroll
            for (int k = 0; k < (PGL ? PGL : GSB_MAXP + 1); ++k) if (k < pgL) th.bL[k] = G.gtab[L][(i64)th.qL * pgL + k];
        }
    }
    GSB_SYNCTHREADS();      // barriers initialised, vacc zeroed

    // producer: map data + coefficient tensor (+ load density) of point (e_, t0) x column -> tile tb; the span's basis table rides along
    auto produce = [&](int tid, int e_, int tb) {
        FusedThread<P1, NG> &th = GSB_TH;
        const int pt = tid - NCONS, t0 = pt / TC, lane = pt - t0 * TC;
        double *Db = Dt + tb * TILE, *Tb = tsel + tb * (2 * P1 * P1);
        if (pt < P1 * P1) { const double2 v = ld_keep2(A.tab + (i64)e_ * P1 * P1 + pt); Tb[pt] = v.x; Tb[P1 * P1 + pt] = v.y; }
        const int q0 = e_ * P1 + t0 + G.qoff[0];
        const double *Ep = lc_row + (i64)q0 * lc_pt + (i64)th.gfL * (nfg * DIM);
        if (e_ + 1 < e_end) {       // next span's coefficients: pull them into L1 while this span is computed
            const char *nx_ = (const char *)(Ep + (i64)P1 * lc_pt);
            for (int b = 0; b < pgL * nfg * DIM * 8; b += 128) prefetch_l1(nx_ + b);
            prefetch_l1(nx_ + pgL * nfg * DIM * 8 - 8);
        }
        double val[NFM], dd[NFM][DIM];
#pragma unroll
        for (int f = 0; f < NFM; ++f) { val[f] = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) dd[f][k] = 0.0; }
#pragma unroll
        for (int k = 0; k < (PGL ? PGL : GSB_MAXP + 1); ++k) {
            if (k >= pgL) break;
            const double *Ea = Ep + k * (nfg * DIM);
            const double2 bL = th.bL[k];
#pragma unroll
            for (int f = 0; f < NFM; ++f) {
                if (f >= nfg) break;
                const double e0 = ld_keep(Ea + f * DIM), e1 = ld_keep(Ea + f * DIM + 1), e2 = DIM == 3 ? ld_keep(Ea + f * DIM + DIM - 1) : 0.0;
                val[f] = fma(bL.x, e0, val[f]);
                dd[f][0] = fma(bL.x, e1, dd[f][0]);
                if (DIM == 3) dd[f][1] = fma(bL.x, e2, dd[f][1]);
                dd[f][L] = fma(bL.y, e0, dd[f][L]);
            }
        }
        double x[3] = {0.0, 0.0, 0.0}, J[DIM][DIM];   // J[c][a] = d x_c / d xi_a
        if (RATIONAL) {
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
        int ql[DIM];
        ql[0] = q0; if (DIM == 3) ql[1] = ql1; ql[L] = th.qL;
        geo_finish_to<DIM, FSPEC, true>(G, Db, NPT, nf ? Db + NIN * NPT : (double *)0, NPT, t0 * TC + lane, ql, x, J);
    };
    // consumer: integrate span e_ from tile tb
    auto consume = [&](int tid, int tb) {
        FusedThread<P1, NG> &th = GSB_TH;
        const int grp = tid / TC, lane = tid - grp * TC;
        const double *Db = Dt + tb * TILE, *Tb = tsel + tb * (2 * P1 * P1);
#pragma unroll
        for (int g = 0; g < NG; ++g)
            if (FULLG || g < th.ngv) {
                const int pk = th.pk[g];
                const double *pv = Db + ((pk >> 2) & 63) * NPT + lane, *pa = Tb + ((pk >> 1) & 1) * (P1 * P1), *pb = Tb + (pk & 1) * (P1 * P1);
#pragma unroll
                for (int t = 0; t < P1; ++t) {
                    const double v = pv[t * TC];
                    double bo[P1], bp[P1];
#ifndef GSB200_EMULATE
                    if constexpr (P1 % 2 == 0) {      // table rows are 16-byte aligned: half as many shared-memory instructions
                        const double2 *pa2 = reinterpret_cast<const double2 *>(pa + t * P1), *pb2 = reinterpret_cast<const double2 *>(pb + t * P1);
#pragma unroll
                        for (int k = 0; k < P1 / 2; ++k) { const double2 u = pa2[k], w = pb2[k]; bo[2 * k] = u.x; bo[2 * k + 1] = u.y; bp[2 * k] = w.x; bp[2 * k + 1] = w.y; }
                    } else
#endif
                    {
#pragma unroll
                        for (int k = 0; k < P1; ++k) { bo[k] = pa[t * P1 + k]; bp[k] = pb[t * P1 + k]; }
                    }
#pragma unroll
                    for (int a = 0; a < P1; ++a) {
                        const double z = bo[a] * v;
#pragma unroll
                        for (int b = 0; b < P1; ++b) th.acc[a][b][g] = fma(bp[b], z, th.acc[a][b][g]);
                    }
                }
            }
        if (grp == 0) {
#pragma unroll
            for (int c = 0; c < GSB_FUSE_MAXF; ++c)
                if (c < nf) {
                    double va[P1];
#pragma unroll
                    for (int k = 0; k < P1; ++k) va[k] = vacc[c][k][lane];
#pragma unroll
                    for (int t = 0; t < P1; ++t) {
                        const double fv = Db[(NIN + c) * NPT + t * TC + lane];
#pragma unroll
                        for (int k = 0; k < P1; ++k) va[k] = fma(Tb[t * P1 + k], fv, va[k]);
                    }
#pragma unroll
                    for (int k = 0; k < P1; ++k) vacc[c][k][lane] = va[k];
                }
        }
    };
    // exit of function x in slot S = x mod p+1: its row of 2p+1 deltas is complete (the negative ones waited in hold[S][.]); the pairs
    // (x + a, x) are complete as well and wait in hold for THEIR owner's exit; the slot is cleared for the function that enters next.
    // S is a compile-time value: the accumulators never move between registers.
    auto exitS = [&](auto sc, int tid, int x) {
        constexpr int S = decltype(sc)::value;
        FusedThread<P1, NG> &th = GSB_TH;
        const int grp = tid / TC, lane = tid - grp * TC;
        const bool mine = th.live && x >= seg_xmin && x < seg_xmax;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            if constexpr (ROWS) {
                // every (function, delta) row of A1 is contiguous along the columns: a completed pair is stored at once, by owner
                if (th.live && (FULLG || g < th.ngv)) {
                    double *po = th.pw[g];
                    if (x >= seg_xmin && x < seg_xmax) {
#pragma unroll
                        for (int b = 0; b < P1; ++b) { st_stream(po, th.acc[S][(S + b) % P1][g]); po += A.out_ds; }
                    }
                    const i64 sa = A.out_fs - A.out_ds;        // owner x + a, delta -a
                    po = th.pw[g];
#pragma unroll
                    for (int a = 1; a < P1; ++a) { po += sa; if (x + a >= seg_xmin && x + a < seg_xmax) st_stream(po, th.acc[(S + a) % P1][S][g]); }
                }
                th.pw[g] += A.out_fs;
            } else {
                if (mine && (FULLG || g < th.ngv)) {
                    double *po = th.pw[g];
                    if (!A.half_rows) {
#pragma unroll
                        for (int j = P1 - 1; j >= 1; --j) st_stream(po - j * DS, th.hold[S][j][g]);
                    }
#pragma unroll
                    for (int b = 0; b < P1; ++b) st_stream(po + b * DS, th.acc[S][(S + b) % P1][g]);
                }
                th.pw[g] += A.out_fs;
#pragma unroll
                for (int a = 1; a < P1; ++a) th.hold[(S + a) % P1][a][g] = th.acc[(S + a) % P1][S][g];
            }
#pragma unroll
            for (int k = 0; k < P1; ++k) { th.acc[S][k][g] = 0.0; th.acc[k][S][g] = 0.0; }
        }
        if (grp == 0) {
#pragma unroll
            for (int c = 0; c < GSB_FUSE_MAXF; ++c)
                if (c < nf) {
                    if (mine) A.v1[c * A.v1_cs + (i64)x * A.v1_fs + (i64)row * A.ncolL + (col0 + lane)] = vacc[c][S][lane];
                    vacc[c][S][lane] = 0.0;
                }
        }
    };
    auto exitX = [&](int tid, int x) {
        const int sx = x % P1;
        static_for<0, P1>([&](auto sc) { if (sx == decltype(sc)::value) exitS(sc, tid, x); });
    };

    // ---- ring of NB tiles: producers fill (full[b] completes when all of them arrived), consumers drain (empty[b]).
    // The two roles are separate loops (separate register allocations: the accumulators are not live in the producer loop); the
    // interpreter build walks them in lock step.
#ifndef GSB200_EMULATE
    // register budget per role where both are whole warpgroups (setmaxnreg works on 4 aligned warps): the producers give registers
    // back, the consumers (accumulators + held pairs + operands) take them
    constexpr bool REBALANCE = NCONS % 128 == 0 && NPROD % 128 == 0;
    if (threadIdx.x >= NCONS) {
        if constexpr (REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GSB_FUSE_PREG));
        const int tid = threadIdx.x;
        int tb = 0; unsigned par = 0;       // parity of the ring round
        for (int e = e_begin; e < e_end; ++e) {
            fbar_wait(&empty[tb], par ^ 1u);        // first round: passes at once
            produce(tid, e, tb);
            fbar_arrive(&full[tb]);
            if (++tb == NB) { tb = 0; par ^= 1u; }
        }
    } else {
        if constexpr (REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GSB_FUSE_CREG));
        const int tid = threadIdx.x;
        int f0 = A.first[e_begin], tb = 0; unsigned par = 0;
        int e = e_begin;
        while (e < e_end) {
            // fast path: p+1 spans with one exit each, starting at slot 0 -> every slot index below is a compile-time value
            bool fast = f0 % P1 == 0 && e + P1 <= e_end;
            if (fast) {
#pragma unroll
                for (int i = 0; i < P1; ++i) fast = fast && A.nexit[e + i] == 1;
            }
            if (fast) {
                static_for<0, P1>([&](auto sc) {
                    fbar_wait(&full[tb], par);
                    consume(tid, tb);
                    fbar_arrive(&empty[tb]);
                    exitS(sc, tid, f0 + decltype(sc)::value);
                    if (++tb == NB) { tb = 0; par ^= 1u; }
                });
                e += P1; f0 += P1;
            } else {
                const int nx = A.nexit[e];
                fbar_wait(&full[tb], par);
                consume(tid, tb);
                fbar_arrive(&empty[tb]);
                for (int x = 0; x < nx; ++x) exitX(tid, f0 + x);
                f0 += nx; ++e;
                if (++tb == NB) { tb = 0; par ^= 1u; }
            }
        }
    }
#else
    int f0 = A.first[e_begin], tb = 0;
    for (int e = e_begin; e < e_end; ++e) {
        const int nx = A.nexit[e];
        GSB_THREADS(tid) if (tid >= NCONS) produce(tid, e, tb);
        GSB_THREADS(tid) if (tid < NCONS) { consume(tid, tb); for (int x = 0; x < nx; ++x) exitX(tid, f0 + x); }
        f0 += nx;
        if (++tb == NB) tb = 0;
    }
#endif
#undef GSB_TH
}

template <int DIM, int P1, class T, int NG, int PGL, bool RATIONAL, int FSPEC, bool ROWS = false>
GSB_GLOBAL void
#ifndef GSB200_EMULATE
__launch_bounds__((fused_threads<P1, T, NG>()), (512 / fused_threads<P1, T, NG>() > 0 ? 512 / fused_threads<P1, T, NG>() : 1))
#endif
k_geo_sweep(const GSB_GRID_CONSTANT FusedArgs A) { geo_sweep_body<DIM, P1, T, NG, PGL, RATIONAL, FSPEC, ROWS>(A); }
