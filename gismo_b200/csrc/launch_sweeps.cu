// launch_sweeps.cu — instantiations and launch dispatch of the sum-factorisation sweep kernels (kernels.cuh):
// k_sweepw (window kernel, p+1-point rules) and k_sweep (generic quadrature sizes).
#include "launch.h"

namespace gsb {

// Generic variant (any quadrature size): one thread per column and owner-slot group, inputs straight from global memory.
template <int P1, class T, bool FINAL, int IS>
static int launch_sweep_i(const SweepArgs &A, int nseg, stream_t s, i64 *flops_per_point)
{
    *flops_per_point = (i64)P1 * (2 * T::NT - n_first<T>() + 2 * P1 * n_has<T>());
    dim3 grid((unsigned)((A.ncol + 127) / 128), P1 / IS, nseg);
    auto kfn = k_sweep<P1, T, IS, FINAL>;
    GSB_LAUNCH(kfn, grid, dim3(128), s, A);
    return 0;
}
#ifndef GSB200_EMULATE
template <class K> static int window_launch(K kfn, dim3 grid, size_t smem, stream_t s, const SweepArgs &A)
{
    GSB_TRY(grant_dynamic_smem((const void *)kfn, smem));
    if (!dry_run()) { kfn<<<grid, dim3(128), smem, s>>>(A); note_launch(); }
    return 0;
}
#endif
template <int P1, class T, bool FINAL, int NG, int GI>
static int launch_window_groups(const SweepArgs &A, int nseg, stream_t s)
{
    if constexpr (GI * NG < T::NOUT) {
        constexpr unsigned OMASK = group_mask<T, NG>(GI);
        const dim3 grid((unsigned)((A.ncol + 127) / 128), 1, nseg);
#ifndef GSB200_EMULATE
        // cp.async ring of 2 stages (one span ahead): measured best against 0 (register double buffer), 3 and 4 (profiles/r01b_layout_experiments.txt)
        GSB_TRY(window_launch(k_sweepw<P1, T, OMASK, FINAL, 2>, grid, window_smem<P1, T, OMASK, 2>(), s, A));
#else
        { auto kfn = k_sweepw<P1, T, OMASK, FINAL, 2>; GSB_LAUNCH(kfn, grid, dim3(128), s, A); }
#endif
        return launch_window_groups<P1, T, FINAL, NG, GI + 1>(A, nseg, s);
    }
    return 0;
}
static bool use_window(const SweepArgs &A, int P1) { return A.q == P1; }

template <int P1, class T, bool FINAL>
static int launch_sweep_t(const SweepArgs &A, int nseg, stream_t s, i64 *fpp)
{
    constexpr int IS = pick_is(P1, T::NOUT);     // owner slots per thread of the generic kernel: the largest that keeps the accumulators in registers
    if (use_window(A, P1)) {
        *fpp = (i64)P1 * (2 * T::NT - n_first<T>() + 2 * P1 * n_has<T>());
        return launch_window_groups<P1, T, FINAL, window_ng(P1, T::NOUT), 0>(A, nseg, s);
    }
    return launch_sweep_i<P1, T, FINAL, IS>(A, nseg, s, fpp);
}
template <class T, bool FINAL>
static int launch_sweep(int P1, const SweepArgs &A, int nseg, stream_t s, i64 *fpp)
{
    switch (P1) {
    case 2: return launch_sweep_t<2, T, FINAL>(A, nseg, s, fpp);
    case 3: return launch_sweep_t<3, T, FINAL>(A, nseg, s, fpp);
    case 4: return launch_sweep_t<4, T, FINAL>(A, nseg, s, fpp);
    case 5: return launch_sweep_t<5, T, FINAL>(A, nseg, s, fpp);
    default: set_error("degree %d not supported by the sweep kernels (1..4)", P1 - 1); return GSB200_EUNSUPPORTED;
    }
}

int dispatch_sweep(int kind, int stage, int P1, const SweepArgs &A, int nseg, stream_t s, i64 *fpp)
{
    if (kind == KIND_MASS) return stage == 2 ? launch_sweep<TMass, true>(P1, A, nseg, s, fpp) : launch_sweep<TMass, false>(P1, A, nseg, s, fpp);
    if (stage == 4) {      // middle sweep of a symmetric 3-D form on the half-stored first-sweep output (degree 3, window kernel only)
        if (kind != KIND_SYM || P1 != 4 || A.q != P1) { set_error("half-stored first-sweep output: degree 3 with the 4-point rule only"); return GSB200_EUNSUPPORTED; }
        return launch_sweep_t<4, T3SymS2U, false>(A, nseg, s, fpp);
    }
    if (stage == 2) return launch_sweep<TLast, true>(P1, A, nseg, s, fpp);
    if (stage == 0) return kind == KIND_SYM ? launch_sweep<T3SymS1, false>(P1, A, nseg, s, fpp) : launch_sweep<T3GenS1, false>(P1, A, nseg, s, fpp);
    if (stage == 1) return kind == KIND_SYM ? launch_sweep<T3SymS2, false>(P1, A, nseg, s, fpp) : launch_sweep<T3GenS2, false>(P1, A, nseg, s, fpp);
    return kind == KIND_SYM ? launch_sweep<T2SymS1, false>(P1, A, nseg, s, fpp) : launch_sweep<T2GenS1, false>(P1, A, nseg, s, fpp);
}

} // namespace gsb
