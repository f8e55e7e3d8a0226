// launch.h — declarations shared by the translation units of libgsb200.so (gsb200.cu: orchestration + C ABI;
// launch_sweeps.cu: sweep kernel instantiations; launch_fused.cu: fused geometry + first-sweep instantiations).
// The split only exists to compile the template instantiations in parallel.
#pragma once
#include "kernels.cuh"
#include <algorithm>
#include <vector>

namespace gsb {

#define GSB_TRY(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)

enum { KIND_SYM = 0, KIND_GEN = 1, KIND_MASS = 2 };
struct HostProgram { std::vector<int> ops; std::vector<double> consts; };

constexpr int pick_is(int P1, int NOUT)
{
    int best = 1;
    for (int d = 1; d <= P1; ++d) if (P1 % d == 0 && d * P1 * NOUT <= 40) best = d;
    return best;
}
template <class T> constexpr int n_has() { int n = 0; for (int o = 0; o < T::NOUT; ++o) for (int b = 0; b < 2; ++b) if (T::has(o, b)) ++n; return n; }
template <class T> constexpr int n_first() { int n = 0; for (int k = 0; k < T::NT; ++k) if (T::first(k)) ++n; return n; }
// Window kernels: one launch (fused kernel: one warp) per group of output components; a group is as many outputs as keep the
// (p+1)^2 * NG accumulators in registers.
constexpr int window_ng(int P1, int NOUT)
{
    const int per_out = P1 * P1 + (GSB_WINDOW_HOLD(P1) ? P1 * (P1 - 1) / 2 : 0);     // accumulators (+ held pairs) per output component
    int ng = 36 / per_out; if (ng < 1) ng = 1; if (ng > NOUT) ng = NOUT;
    if (NOUT % ng != 0 && ng > 1 && NOUT % (ng - 1) == 0) --ng;                       // balanced groups
    return ng;
}

// stage: 0 = first of 3-D, 1 = middle of 3-D, 2 = last; 3 = first of 2-D; 4 = middle of 3-D on the half-stored first-sweep output
int dispatch_sweep(int kind, int stage, int P1, const SweepArgs &A, int nseg, stream_t s, i64 *fpp);

// K0 fused into the first sweep (fused.cuh)
struct FusedCtx { const std::vector<HostProgram> *progs; int device; int *jit_launches; };
int launch_fused(const FusedCtx &ctx, int kind, int dim, int P1, const FusedArgs &FA, int nseg, stream_t s, bool hot, bool rat, int pgl, bool rows, i64 *fpp);

// second + last sweep of a 3-D form in one kernel (fused23.cuh); ne_max = longest tile (spans of the last direction)
int launch_s23(int kind, int P1, const S23Args &A, dim3 grid, int ne_max, stream_t s, i64 *fpp2, i64 *fpp3);
bool s23_available(int kind, int P1);

#ifndef GSB200_EMULATE
// more than 48 KB of dynamic shared memory has to be granted per kernel AND per device
int grant_dynamic_smem(const void *kfn, size_t smem);
cudaKernel_t jit_geometry_kernel(const std::vector<HostProgram> &progs, int device, int dim, int pgl, bool rational, int fspec);
cudaKernel_t jit_fused_kernel(const std::vector<HostProgram> &progs, int device, int dim, int p1, const char *table, int ng, int nthr, int pgl, bool rational, int fspec, bool rows);
#endif

} // namespace gsb
