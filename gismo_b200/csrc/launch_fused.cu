// launch_fused.cu — instantiations and launch dispatch of the fused geometry + first-sweep kernel (fused.cuh).
#include "launch.h"

namespace gsb {

// K0 fused into the first sweep (fused.cuh).  One CTA = GSB_FUSE_TC columns x all output groups.
template <int DIM, int P1, class T, bool HOTOK, bool ROWS>
static int launch_fused_r(const FusedCtx &a, const FusedArgs &FA, int nseg, stream_t s, bool hot, bool rat, int pgl, const char *tname, i64 *fpp)
{
    constexpr int NG = window_ng(P1, T::NOUT), NTHR = fused_threads<P1, T, NG>();
    *fpp = (i64)P1 * (2 * T::NT - n_first<T>() + 2 * P1 * n_has<T>());
    const int tiles = (FA.ncolL + GSB_FUSE_TC - 1) / GSB_FUSE_TC;
    const dim3 grid((unsigned)((i64)tiles * FA.nrows), 1, nseg);
#ifndef GSB200_EMULATE
#define GSB_FUSED(PG_, R_, F_) { cudaKernel_t jk = (FA.nf && !dry_run()) ? jit_fused_kernel(*a.progs, a.device, DIM, P1, tname, NG, NTHR, PG_, R_, F_, ROWS) : 0; \
        if (jk) { void *kargs[] = {(void *)&FA}; GSB_TRY(dev_check(cudaLaunchKernel((const void *)jk, grid, dim3(NTHR), kargs, 0, s), "launch of the compiled fused kernel")); note_launch(); ++*a.jit_launches; } \
        else { auto kfn = k_geo_sweep<DIM, P1, T, NG, PG_, R_, F_, ROWS>; GSB_LAUNCH_CTA(kfn, grid, dim3(NTHR), s, FA); } }
#else
    (void)a; (void)tname;
#define GSB_FUSED(PG_, R_, F_) { auto kfn = k_geo_sweep<DIM, P1, T, NG, PG_, R_, F_, ROWS>; GSB_LAUNCH_CTA(kfn, grid, dim3(NTHR), s, FA); }
#endif
    if constexpr (HOTOK) {
        if (hot && !rat && pgl == 2) { GSB_FUSED(2, false, 1) return 0; }
        if (hot && !rat && pgl == 3) { GSB_FUSED(3, false, 1) return 0; }
    }
    if (rat) GSB_FUSED(0, true, 0) else GSB_FUSED(0, false, 0)
#undef GSB_FUSED
    return 0;
}
template <int DIM, int P1, class T, bool HOTOK>
static int launch_fused_t(const FusedCtx &a, const FusedArgs &FA, int nseg, stream_t s, bool hot, bool rat, int pgl, bool rows, const char *tname, i64 *fpp)
{
    // the rows layout of A1 feeds the fused second + last sweep: 3-D gradient forms up to degree 3 only
    if constexpr (DIM == 3 && P1 <= 4 && T::NOUT > 1) { if (rows) return launch_fused_r<DIM, P1, T, HOTOK, true>(a, FA, nseg, s, hot, rat, pgl, tname, fpp); }
    if (rows) { set_error("rows layout requested for a configuration without it"); return GSB200_EINVAL; }
    return launch_fused_r<DIM, P1, T, HOTOK, false>(a, FA, nseg, s, hot, rat, pgl, tname, fpp);
}
template <int DIM, class T, bool HOTOK>
static int launch_fused_p(const FusedCtx &a, int P1, const FusedArgs &FA, int nseg, stream_t s, bool hot, bool rat, int pgl, bool rows, const char *tname, i64 *fpp)
{
    switch (P1) {
    case 2: return launch_fused_t<DIM, 2, T, HOTOK>(a, FA, nseg, s, hot, rat, pgl, rows, tname, fpp);
    case 3: return launch_fused_t<DIM, 3, T, HOTOK>(a, FA, nseg, s, hot, rat, pgl, rows, tname, fpp);
    case 4: return launch_fused_t<DIM, 4, T, HOTOK>(a, FA, nseg, s, hot, rat, pgl, rows, tname, fpp);
    case 5: return launch_fused_t<DIM, 5, T, HOTOK>(a, FA, nseg, s, hot, rat, pgl, rows, tname, fpp);
    default: set_error("degree %d not supported by the sweep kernels (1..4)", P1 - 1); return GSB200_EUNSUPPORTED;
    }
}
int launch_fused(const FusedCtx &a, int kind, int dim, int P1, const FusedArgs &FA, int nseg, stream_t s, bool hot, bool rat, int pgl, bool rows, i64 *fpp)
{
    if (kind == KIND_MASS) return dim == 3 ? launch_fused_p<3, TMass, false>(a, P1, FA, nseg, s, false, rat, pgl, rows, "TMass", fpp)
                                           : launch_fused_p<2, TMass, false>(a, P1, FA, nseg, s, false, rat, pgl, rows, "TMass", fpp);
    if (dim == 3) return kind == KIND_SYM ? launch_fused_p<3, T3SymS1, true>(a, P1, FA, nseg, s, hot, rat, pgl, rows, "T3SymS1", fpp)
                                          : launch_fused_p<3, T3GenS1, false>(a, P1, FA, nseg, s, false, rat, pgl, rows, "T3GenS1", fpp);
    return kind == KIND_SYM ? launch_fused_p<2, T2SymS1, true>(a, P1, FA, nseg, s, hot, rat, pgl, rows, "T2SymS1", fpp)
                            : launch_fused_p<2, T2GenS1, false>(a, P1, FA, nseg, s, false, rat, pgl, rows, "T2GenS1", fpp);
}

// ------------------------------------------------------------------ fused second + last sweep (fused23.cuh)
bool s23_available(int kind, int P1)
{
    // opt-in (GSB200_S23=1): measured on B200 the fused kernel removes 25 GB of HBM traffic per assembly at config 2 but is bound by
    // instruction issue / shared-memory latency of the direction-2 warps (23.8 ms against 8.8 ms for the two separate sweeps):
    // profiles/r02_s23_experiment.txt
    const char *e = getenv("GSB200_S23");      // read at every assembly: the tests switch it
    const bool env = e && atoi(e) > 0;
    return env && kind != KIND_MASS && P1 >= 2 && P1 <= 4;
}
template <int P1, class T2>
static int launch_s23_t(const S23Args &A, dim3 grid, int ne_max, stream_t s, i64 *fpp2, i64 *fpp3)
{
    *fpp2 = (i64)P1 * (2 * T2::NT - n_first<T2>() + 2 * P1 * n_has<T2>());
    *fpp3 = (i64)P1 * (2 * TLast::NT - n_first<TLast>() + 2 * P1 * n_has<TLast>());
    auto kfn = k_s23<P1, T2>;
#ifndef GSB200_EMULATE
    (void)ne_max;
    const size_t smem = (size_t)s23_smem_doubles<P1, T2>() * sizeof(double);
    GSB_TRY(grant_dynamic_smem((const void *)kfn, smem));
    if (!dry_run()) { kfn<<<grid, dim3(S23_NS2T + S23_NS3T), smem, s>>>(A); note_launch(); }
#else
    (void)ne_max;
    GSB_LAUNCH_CTA(kfn, grid, dim3(S23_NS2T + S23_NS3T), s, A);
#endif
    return 0;
}
int launch_s23(int kind, int P1, const S23Args &A, dim3 grid, int ne_max, stream_t s, i64 *fpp2, i64 *fpp3)
{
    switch (P1) {
    case 2: return kind == KIND_SYM ? launch_s23_t<2, T3SymS2>(A, grid, ne_max, s, fpp2, fpp3) : launch_s23_t<2, T3GenS2>(A, grid, ne_max, s, fpp2, fpp3);
    case 3: return kind == KIND_SYM ? launch_s23_t<3, T3SymS2>(A, grid, ne_max, s, fpp2, fpp3) : launch_s23_t<3, T3GenS2>(A, grid, ne_max, s, fpp2, fpp3);
    case 4: return kind == KIND_SYM ? launch_s23_t<4, T3SymS2>(A, grid, ne_max, s, fpp2, fpp3) : launch_s23_t<4, T3GenS2>(A, grid, ne_max, s, fpp2, fpp3);
    default: set_error("fused second sweep: degree %d not available", P1 - 1); return GSB200_EUNSUPPORTED;
    }
}

} // namespace gsb
