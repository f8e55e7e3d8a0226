// fused23.cuh — second AND last sweep of a 3-D form in one kernel: the intermediate A2 (12.8 GB written + re-read per assembly
// at config 2) never exists; rows of it are handed from the warps that contract direction 1 to the warps that contract direction 2
// through shared memory.  Included by kernels.cuh (inside namespace gsb).
//
// Replaces k_sweepw<.., T3*S2, ..> + k_sweepw<.., TLast, .., FINAL> (a10/a11 contraction of directions 1 and 2, a14/a15/a16 scatter:
// gsVisitorPoisson.h:89-118, gsSparseSystem.h:972-1010, gsExprAssembler.h:553-630).
//
// CTA = one (i0, delta0) pair of direction 0 x one tile of direction 2 (NE <= S23_NEMAX knot spans and the functions whose
// support lies inside).  Warp-specialised:
//   S2 warps: thread = one quadrature point of direction 2 of the tile; marches along direction 1 with ALL output components of
//       the (p+1)^2 pairs alive on the span in registers (slot order, static rotation: no register moves); inputs (rows of A1,
//       contiguous along direction 2) through a thread-private cp.async ring.  When a function of direction 1 leaves the window,
//       its 2p+1 completed pairs (i1, delta1) x 4 components go to a shared-memory row set.
//   S3 warps: per row set, (1) thread = (pair, span of direction 2): the span's (p+1)^2 local pair integrals over its p+1 points
//       (same z-trick as the sweeps); (2) thread = (pair, owned function i2): sums the spans shared with each partner and writes
//       the 2p+1 entries K[(i0,i1,i2),(d0,d1,d2)] straight into their CSC slots (closed form / canonical / generic, as the final
//       sweep did).
// Row sets go round a two-deep ring with full/empty mbarriers; the S3 warps use a named barrier between their two phases.
// Every matrix entry is still produced exactly once, by one thread: no atomics for free rows, deterministic.
#define S23_NS2T 256          // S2 threads (one per direction-2 quadrature point of the tile): two warpgroups
#define S23_NS3T 256          // S3 threads: two warpgroups (their work is as large as the S2 side's and has less instruction-level parallelism)
#define S23_NEMAX 52          // spans of direction 2 per tile = row length of the shared-memory arrays: 4 (mod 16) keeps both access patterns conflict-free
#define S23_NSTG 3            // cp.async ring: points of direction 1 in flight
#ifndef S23_S2REG
#define S23_S2REG 184         // registers per S2 / S3 thread after the split (setmaxnreg): 256 * 184 + 256 * 72 = 64 K
#define S23_S3REG 72
#endif

struct S23Args {
    const int *first1, *nexit1; const double2 *tab1;       // direction 1: first function / exits per span, table [e][t][slot]
    const int *seg1;                                        // [gridDim.z][4] e_begin, e_end, x_min, x_max of direction 1
    const int *first2; const double2 *tabl2;               // direction 2: first function per span, table [e][t][a] (local order)
    const int *ffirst2, *flast2;                            // direction 2: first / last span of a function's support
    const int *tiles;                                       // [gridDim.y][4] e_begin, e_end, x_min, x_max of direction 2
    int e_in0;                                              // first span of direction 2 present in A1 (chunk)
    const double *a1; i64 a1_cs, a1_row, a1_q1;             // A1[o][(i0,d0)][q1][q2]: strides of component, (i0,d0) row block, q1
    int n0, W0;                                             // blockIdx.x = i0 * W0 + (d0 + p0)
    FinalArgs fin;
};

// dynamic shared memory (doubles) for a tile of ne spans
template <int P1, class T2> GSB_CX int s23_smem_doubles()
{
    const int nep = S23_NEMAX, NP = 2 * P1 - 1;
    return S23_NSTG * T2::NIN * S23_NS2T                  // A1 ring
         + 2 * NP * T2::NOUT * P1 * nep                   // row sets
         + NP * P1 * P1 * nep                             // local pair integrals
         + P1 * P1 * 2 * nep                              // direction-2 table of the tile
         + (S23_NS2T / 32) * 2 * P1 * P1 * 2              // per S2 warp: direction-1 table of the span, two buffers
         + 8                                              // barriers
         + 3 * (S23_NEMAX + 2 * GSB_MAXP + 2);            // index tables of direction 2 (ints, two per double)
}

#ifndef GSB200_EMULATE
GSB_DEVICE void s23_bar_s3() { asm volatile("bar.sync 1, %0;" ::"n"(S23_NS3T) : "memory"); }
#else
static inline void s23_bar_s3() {}
#endif

template <int P1, class T2>
struct S23Thread {
    double acc[P1][P1][T2::NOUT];     // S2: slot order
};

template <int P1, class T2>
GSB_DEVICE void s23_body(const S23Args &A, double *smem)
{
    constexpr int NIN = T2::NIN, NOUT = T2::NOUT, NT = T2::NT, NP = 2 * P1 - 1, NS2T = S23_NS2T, NS3T = S23_NS3T, NTHR = NS2T + NS3T;
    static_assert(NOUT == 4, "second-sweep tables produce the four flag combinations of the last direction");
    const FinalArgs &F = A.fin;
    // ---- what this CTA owns
    const int i0 = (int)(blockIdx.x / A.W0), d0 = (int)(blockIdx.x % A.W0) - F.p[0];
    { const int j0 = i0 + d0; if (j0 < F.plo[0][i0] || j0 > F.phi[0][i0]) return; }         // the pair never co-occurs (uniform exit)
    const int tl = blockIdx.y, sg = blockIdx.z;
    const int e2b = A.tiles[4 * tl + 0], e2e = A.tiles[4 * tl + 1], x2min = A.tiles[4 * tl + 2], x2max = A.tiles[4 * tl + 3];
    const int e1b = A.seg1[4 * sg + 0], e1e = A.seg1[4 * sg + 1], x1min = A.seg1[4 * sg + 2], x1max = A.seg1[4 * sg + 3];
    constexpr int NEP = S23_NEMAX;
    const int NE = e2e - e2b, NOWN = x2max - x2min;
    // ---- shared memory carve-up
    double *ring = smem;                                         // [stage][component][S2 thread]
    double *rows = ring + S23_NSTG * NIN * NS2T;                 // [buffer][pair][component][t2][span]
    double *local = rows + 2 * NP * NOUT * P1 * NEP;             // [pair][a * P1 + b][span]
    double *tab2s = local + NP * P1 * P1 * NEP;                  // [t2][a][value | derivative][span]
    double *tab1w = tab2s + P1 * P1 * 2 * NEP;                   // [S2 warp][buffer][t1][slot][value | derivative]
    unsigned long long *bars = (unsigned long long *)(tab1w + (NS2T / 32) * 2 * P1 * P1 * 2);     // full[2], empty[2]
    int *first2s = (int *)(bars + 8);                            // first function of every span of the tile
    int *ff2s = first2s + S23_NEMAX + 2;                         // first / last span of the functions [x2min - p, x2max + p) ...
    int *fl2s = ff2s + S23_NEMAX + 2 * GSB_MAXP + 2;
    int *dlo2s = fl2s + S23_NEMAX + 2 * GSB_MAXP + 2;            // ... and, per owned function: partner range (dlo + 8) | (dhi + 8) << 4, first span in the tile << 8,
                                                                 // number of spans << 16, local index of the function in span m at bits 20 + 3 m
    constexpr int ROWSET = NP * NOUT * P1 * NEP;
#ifdef GSB200_EMULATE
    static thread_local S23Thread<P1, T2> *states = 0; static thread_local int nstates = 0;
    if (nstates < NS2T) { delete[] states; states = new S23Thread<P1, T2>[NS2T]; nstates = NS2T; }
#define S23_TH states[tid]
#else
    S23Thread<P1, T2> state_;
#define S23_TH state_
    if (threadIdx.x == 0) { fbar_init(&bars[0], NS2T); fbar_init(&bars[1], NS2T); fbar_init(&bars[2], NS3T); fbar_init(&bars[3], NS3T); }
#endif
    // direction-2 table of the tile, transposed so that lanes = consecutive spans read consecutive words
    GSB_THREADS(tid) {
        for (int it = tid; it < NE * P1 * P1; it += NTHR) {
            const int el = it / (P1 * P1), r = it - el * (P1 * P1);      // r = t2 * P1 + a
            const double2 v = ld_keep2(A.tabl2 + (i64)(e2b + el) * P1 * P1 + r);
            tab2s[(r * 2 + 0) * NEP + el] = v.x; tab2s[(r * 2 + 1) * NEP + el] = v.y;
        }
        for (int it = tid; it < NE; it += NTHR) first2s[it] = A.first2[e2b + it];
        for (int it = tid; it < NOWN + 2 * (P1 - 1); it += NTHR) {
            const int j2 = x2min - (P1 - 1) + it;
            const bool ok = j2 >= 0 && j2 < F.n[2];
            ff2s[it] = ok ? A.ffirst2[j2] : 0; fl2s[it] = ok ? A.flast2[j2] : -1;
        }
        for (int it = tid; it < NOWN; it += NTHR) {
            const int i2 = x2min + it, ea = A.ffirst2[i2], eb = A.flast2[i2];
            unsigned w = (unsigned)(F.plo[2][i2] - i2 + 8) | ((unsigned)(F.phi[2][i2] - i2 + 8) << 4) | ((unsigned)(ea - e2b) << 8) | ((unsigned)(eb - ea + 1) << 16);
            for (int e = ea; e <= eb; ++e) w |= (unsigned)(i2 - A.first2[e]) << (20 + 3 * (e - ea));
            dlo2s[it] = (int)w;
        }
        if (tid < NS2T) {
            S23Thread<P1, T2> &th = S23_TH;
#pragma unroll
            for (int a = 0; a < P1; ++a)
#pragma unroll
                for (int b = 0; b < P1; ++b)
#pragma unroll
                    for (int g = 0; g < NOUT; ++g) th.acc[a][b][g] = 0.0;
        }
    }
    GSB_SYNCTHREADS();

    // ================================================================= S2 role
    // the thread's column of A1: point tid of the tile (threads beyond the tile work on a clamped column and are never read)
    const i64 a1_base = (i64)blockIdx.x * A.a1_row + (i64)(e2b - A.e_in0) * P1;
    auto s2_issue = [&](int tid, int pt_global) {        // pt_global = index of the direction-1 point; stage = pt mod NSTG
        const int colq = tid < NE * P1 ? tid : NE * P1 - 1;
        const double *src = A.a1 + a1_base + colq + (i64)pt_global * A.a1_q1;
        double *dst = ring + (pt_global % S23_NSTG) * NIN * NS2T + tid;
#pragma unroll
        for (int c = 0; c < NIN; ++c) cp_async8(dst + c * NS2T, src + c * A.a1_cs);
    };
    auto s2_issue_tab = [&](int tid, int e1) {           // the span's direction-1 table, one copy per warp (lanes < P1 * P1)
        const int lane = tid & 31, w = tid >> 5;
        if (lane < P1 * P1) {
            double *dst = tab1w + ((w * 2 + (e1 & 1)) * P1 * P1 + lane) * 2;
#ifndef GSB200_EMULATE
            cp_async16(reinterpret_cast<double2 *>(dst), A.tab1 + (i64)e1 * P1 * P1 + lane);
#else
            const double2 v = A.tab1[(i64)e1 * P1 * P1 + lane]; dst[0] = v.x; dst[1] = v.y;
#endif
        }
    };
    // one direction-1 point: acc[sa][sb][o] += B^(.)_sa B^(.)_sb in_c over the term table (z-trick)
    auto s2_point = [&](int tid, int e1, int t1, int pt_global) {
        S23Thread<P1, T2> &th = S23_TH;
        const double *rg = ring + (pt_global % S23_NSTG) * NIN * NS2T + tid;
        const double *tb = tab1w + (((tid >> 5) * 2 + (e1 & 1)) * P1 * P1 + t1 * P1) * 2;
        double v[NIN];
#pragma unroll
        for (int c = 0; c < NIN; ++c) v[c] = rg[c * NS2T];
        double bx[P1], by[P1];
#pragma unroll
        for (int k = 0; k < P1; ++k) { bx[k] = tb[2 * k]; by[k] = tb[2 * k + 1]; }
#pragma unroll
        for (int a = 0; a < P1; ++a) {
            double z[NOUT][2];
            static_for<0, NT>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const double bo = T2::a(k) ? by[a] : bx[a];
                if constexpr (T2::first(k)) z[T2::o(k)][T2::b(k)] = bo * v[T2::c(k)];
                else z[T2::o(k)][T2::b(k)] = fma(bo, v[T2::c(k)], z[T2::o(k)][T2::b(k)]);
            });
#pragma unroll
            for (int b = 0; b < P1; ++b)
                static_for<0, NOUT>([&](auto oc_) {
                    constexpr int o = decltype(oc_)::value;
                    if constexpr (T2::has(o, 0)) th.acc[a][b][o] = fma(bx[b], z[o][0], th.acc[a][b][o]);
                    if constexpr (T2::has(o, 1)) th.acc[a][b][o] = fma(by[b], z[o][1], th.acc[a][b][o]);
                });
        }
    };
    // exit of function x of direction 1 (slot S): its 2p+1 completed pairs go to row set rb
    //   pair k < P1: (owner x, delta +k);  pair P1-1+a: (owner x+a, delta -a)
    auto s2_exit = [&](auto sc, int tid, int rb) {
        constexpr int S = decltype(sc)::value;
        S23Thread<P1, T2> &th = S23_TH;
        if (tid < NE * P1) {
            const int el = tid / P1, t2 = tid - el * P1;
            double *rw = rows + rb * ROWSET + t2 * NEP + el;
#pragma unroll
            for (int b = 0; b < P1; ++b)
#pragma unroll
                for (int g = 0; g < NOUT; ++g) rw[(b * NOUT + g) * P1 * NEP] = th.acc[S][(S + b) % P1][g];
#pragma unroll
            for (int a = 1; a < P1; ++a)
#pragma unroll
                for (int g = 0; g < NOUT; ++g) rw[((P1 - 1 + a) * NOUT + g) * P1 * NEP] = th.acc[(S + a) % P1][S][g];
        }
#pragma unroll
        for (int k = 0; k < P1; ++k)
#pragma unroll
            for (int g = 0; g < NOUT; ++g) { th.acc[S][k][g] = 0.0; th.acc[k][S][g] = 0.0; }
    };

    // ================================================================= S3 role
    // (1) local pair integrals of one span for one pair: local[k][a * P1 + b][el] = sum_t2 sum_g B^(ag)_a B^(bg)_b row_g.
    // thread = (span el, quarter kq of the pairs): pairs kq and kq + 4
    auto s3_local = [&](int s3tid, int rb) {
        const int el = s3tid % NEP, kq = s3tid / NEP;
        if (el >= NE || kq >= 4) return;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const int k = kq + 4 * kk;
            if (k >= NP) break;
            double acc[P1][P1];
#pragma unroll
            for (int a = 0; a < P1; ++a)
#pragma unroll
                for (int b = 0; b < P1; ++b) acc[a][b] = 0.0;
            const double *rw = rows + rb * ROWSET + (k * NOUT) * P1 * NEP + el;
#pragma unroll
            for (int t2 = 0; t2 < P1; ++t2) {
                double r[NOUT], bx[P1], by[P1];
#pragma unroll
                for (int g = 0; g < NOUT; ++g) r[g] = rw[(g * P1 + t2) * NEP];
#pragma unroll
                for (int a = 0; a < P1; ++a) { bx[a] = tab2s[((t2 * P1 + a) * 2 + 0) * NEP + el]; by[a] = tab2s[((t2 * P1 + a) * 2 + 1) * NEP + el]; }
#pragma unroll
                for (int a = 0; a < P1; ++a) {          // g = 2 * (owner flag) + (partner flag), TLast
                    const double z0 = fma(by[a], r[2], bx[a] * r[0]), z1 = fma(by[a], r[3], bx[a] * r[1]);
#pragma unroll
                    for (int b = 0; b < P1; ++b) acc[a][b] = fma(by[b], z1, fma(bx[b], z0, acc[a][b]));
                }
            }
            double *lw = local + (k * P1 * P1) * NEP + el;
#pragma unroll
            for (int a = 0; a < P1; ++a)
#pragma unroll
                for (int b = 0; b < P1; ++b) lw[(a * P1 + b) * NEP] = acc[a][b];
        }
    };
    // (2) one owned function of direction 2 for one pair: sum the spans shared with each partner, scatter into the CSC arrays.
    // Index data of direction 2 comes from shared memory; the owner records of a thread's items are fetched together up front.
    auto s3_scatter = [&](int s3tid, int x) {
        const int W0 = 2 * F.p[0] + 1, W1 = 2 * F.p[1] + 1, p2 = F.p[2];
        const i64 nlow = (i64)F.n[0] * F.n[1], full = (i64)W0 * W1 * (2 * p2 + 1);
        unsigned kmask = 0;                   // pairs of this exit that exist and belong to this direction-1 segment
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const int a1o = k < P1 ? 0 : k - (P1 - 1), i1 = x + a1o, d1 = k < P1 ? k : -a1o;
            if (i1 >= x1min && i1 < x1max) { const int j1 = i1 + d1; if (j1 >= F.plo[1][i1] && j1 <= F.phi[1][i1]) kmask |= 1u << k; }
        }
        constexpr int MAXR = (NP * S23_NEMAX + NS3T - 1) / NS3T;
        i64 rec[MAXR];
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            const int item = s3tid + r * NS3T;
            rec[r] = 0;
            if (item < NP * NOWN) {
                const int k = item / NOWN, i2 = x2min + (item - k * NOWN);
                const int i1 = x + (k < P1 ? 0 : k - (P1 - 1));
                if ((kmask >> k) & 1u) rec[r] = F.ownrec[F.bcol * F.nb + (i64)i2 * nlow + (i64)i1 * F.n[0] + i0];
            }
        }
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            if (!rec[r]) continue;
            const int item = s3tid + r * NS3T;
            const int k = item / NOWN, i2l = item - k * NOWN, i2 = x2min + i2l;
            const int a1o = k < P1 ? 0 : k - (P1 - 1);
            const int i1 = x + a1o, d1 = k < P1 ? k : -a1o;
            FinalCtx fc;
            fc.li_low = (i64)i1 * F.n[0] + i0; fc.dj_low = (i64)d1 * F.n[0] + d0; fc.nlow = nlow;
            fc.r_low = d1 + F.p[1]; fc.bit0 = d0 + F.p[0];
            const int flag = (int)(rec[r] & 3);
            const unsigned dw = (unsigned)dlo2s[i2l];
            const int dlo = (int)(dw & 15u) - 8, dhi = (int)((dw >> 4) & 15u) - 8, el0 = (int)((dw >> 8) & 255u), ns = (int)((dw >> 16) & 15u);
            const double *lk = local + (k * P1 * P1) * NEP + el0;
            double *vbase = F.values + (rec[r] >> 2) + F.brow * full + (i64)fc.r_low * W0 + fc.bit0;
            // all (partner, span) combinations at once: independent shared-memory reads, no dependent index look-ups
            double val[2 * P1 - 1];
#pragma unroll
            for (int q = 0; q < 2 * P1 - 1; ++q) val[q] = 0.0;
#pragma unroll
            for (int m = 0; m < P1; ++m) {
                if (m < ns) {
                    const int la = (int)((dw >> (20 + 3 * m)) & 7u);
#pragma unroll
                    for (int q = 0; q < 2 * P1 - 1; ++q) {
                        const int lb = la + q - (P1 - 1);
                        if (lb >= 0 && lb < P1) val[q] += lk[(la * P1 + lb) * NEP + m];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2 * P1 - 1; ++q) {
                const int dd = q - (P1 - 1);
                if (dd < dlo || dd > dhi) continue;
                if (flag == 3) st_stream(vbase + (i64)(dd + p2) * W1 * W0, val[q]);
                else {
                    const int run = (dd + p2) * W1 + fc.r_low;
                    if (flag == 1) final_canonical(F, fc, i2, rec[r], dd, run, val[q]);
                    else final_slow(F, fc, i2, rec[r], dd, val[q]);
                }
            }
        }
    };

    // ================================================================= the two role loops
#ifndef GSB200_EMULATE
    if (threadIdx.x < NS2T) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(S23_S2REG));
        const int tid = threadIdx.x;
        const int npts = (e1e - e1b) * P1, pt0 = e1b * P1;
        s2_issue_tab(tid, e1b);
#pragma unroll
        for (int k = 0; k < S23_NSTG - 1; ++k) { if (k < npts) s2_issue(tid, pt0 + k); cp_async_commit(); }
        int f0 = A.first1[e1b], rb = 0; unsigned par = 0;
        auto span = [&](int e, auto exits) {
            if (e + 1 < e1e) s2_issue_tab(tid, e + 1);
#pragma unroll
            for (int t1 = 0; t1 < P1; ++t1) {
                const int pt = e * P1 + t1;
                if (pt + S23_NSTG - 1 < pt0 + npts) s2_issue(tid, pt + S23_NSTG - 1);
                cp_async_commit();
                cp_async_wait<S23_NSTG - 1>();
                if (t1 == 0) warp_sync();         // the table copied by the other lanes is in place
                s2_point(tid, e, t1, pt);
            }
            warp_sync();                          // every lane is done with this span's table before it is refilled two spans on
            exits();
        };
        int e = e1b;
        while (e < e1e) {
            bool fast = f0 % P1 == 0 && e + P1 <= e1e;
            if (fast) {
#pragma unroll
                for (int i = 0; i < P1; ++i) fast = fast && A.nexit1[e + i] == 1;
            }
            if (fast) {
                static_for<0, P1>([&](auto sc) {
                    span(e + decltype(sc)::value, [&] {
                        fbar_wait(&bars[2 + rb], par ^ 1u);
                        s2_exit(sc, tid, rb);
                        fbar_arrive(&bars[rb]);
                        if (++rb == 2) { rb = 0; par ^= 1u; }
                    });
                });
                e += P1; f0 += P1;
            } else {
                const int nx = A.nexit1[e];
                span(e, [&] {
                    for (int x = 0; x < nx; ++x) {
                        const int sx = (f0 + x) % P1;
                        fbar_wait(&bars[2 + rb], par ^ 1u);
                        static_for<0, P1>([&](auto sc) { if (sx == decltype(sc)::value) s2_exit(sc, tid, rb); });
                        fbar_arrive(&bars[rb]);
                        if (++rb == 2) { rb = 0; par ^= 1u; }
                    }
                });
                f0 += nx; ++e;
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(S23_S3REG));
        const int s3tid = threadIdx.x - NS2T;
        int f0 = A.first1[e1b], rb = 0; unsigned par = 0;
        for (int e = e1b; e < e1e; ++e) {
            const int nx = A.nexit1[e];
            for (int x = 0; x < nx; ++x) {
                fbar_wait(&bars[rb], par);
                s3_local(s3tid, rb);
                fbar_arrive(&bars[2 + rb]);       // the row set may be refilled
                s23_bar_s3();
                s3_scatter(s3tid, f0 + x);
                s23_bar_s3();                     // `local` may be overwritten
                if (++rb == 2) { rb = 0; par ^= 1u; }
            }
            f0 += nx;
        }
    }
#else
    {
        int f0 = A.first1[e1b];
        for (int e = e1b; e < e1e; ++e) {
            const int nx = A.nexit1[e];
            GSB_THREADS(tid) if (tid < NS2T) s2_issue_tab(tid, e);
            GSB_THREADS(tid) if (tid < NS2T) {
                for (int t1 = 0; t1 < P1; ++t1) { s2_issue(tid, e * P1 + t1); s2_point(tid, e, t1, e * P1 + t1); }
            }
            for (int x = 0; x < nx; ++x) {
                const int sx = (f0 + x) % P1;
                GSB_THREADS(tid) if (tid < NS2T) static_for<0, P1>([&](auto sc) { if (sx == decltype(sc)::value) s2_exit(sc, tid, 0); });
                GSB_THREADS(tid) if (tid >= NS2T) s3_local(tid - NS2T, 0);
                GSB_THREADS(tid) if (tid >= NS2T) s3_scatter(tid - NS2T, f0 + x);
            }
            f0 += nx;
        }
    }
#endif
#undef S23_TH
}

#ifndef GSB200_EMULATE
template <int P1, class T2>
GSB_GLOBAL void __launch_bounds__(S23_NS2T + S23_NS3T, 1) k_s23(const GSB_GRID_CONSTANT S23Args A)
{
    extern __shared__ __align__(16) double s23_smem[];
    s23_body<P1, T2>(A, s23_smem);
}
#else
template <int P1, class T2>
GSB_GLOBAL void k_s23(const S23Args A)
{
    static thread_local double *buf = 0;
    const size_t n = (size_t)s23_smem_doubles<P1, T2>();
    if (!buf) buf = new double[n];
    for (size_t i = 0; i < n; ++i) buf[i] = 0.0;
    s23_body<P1, T2>(A, buf);
}
#endif
