// launch_s23.cu — instantiations and launch dispatch of the fused second + last sweep kernel (fused23.cuh).
#include "launch.h"

namespace gsb {

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
