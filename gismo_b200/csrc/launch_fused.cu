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

} // namespace gsb
