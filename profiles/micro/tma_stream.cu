// Microbenchmark (profiling aid): how fast can ONE CTA per SM stream contiguous HBM data into a shared-memory
// ring with (a) cp.async.bulk + mbarrier, (b) cp.async 16-byte (LDGSTS), (c) plain 16-byte loads to registers?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(u64 *b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(u64 *b, unsigned n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(u64 *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(u64 *b, unsigned par) { unsigned ok; do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory"); } while (!ok); }
__device__ __forceinline__ void bulk(void *d, const void *s, unsigned n, u64 *b) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory"); }

// mode 0: bulk copies of `chunk` bytes, `nst` stages, `split` copies per stage
__global__ void k_bulk(const double *in, double *out, size_t per_cta, int chunk, int nst, int split)
{
    extern __shared__ __align__(128) unsigned char sm[];
    u64 *full = (u64 *)(sm + (size_t)nst * chunk), *empty = full + nst;
    const int nw = blockDim.x / 32, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int s = 0; s < nst; ++s) { mb_init(full + s, 1); mb_init(empty + s, nw); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const char *src = (const char *)in + (size_t)blockIdx.x * per_cta;
    const int n = (int)(per_cta / chunk);
    auto issue = [&](int i, int s) { if (lane == 0) mb_expect(full + s, chunk); __syncwarp(); const int piece = chunk / split; for (int r = lane; r < split; r += 32) bulk(sm + (size_t)s * chunk + (size_t)r * piece, src + (size_t)i * chunk + (size_t)r * piece, piece, full + s); };
    if (warp == 0) for (int i = 0; i < nst && i < n; ++i) issue(i, i);
    double acc = 0; int s = 0; unsigned par = 0;
    for (int i = 0; i < n; ++i) {
        mb_wait(full + s, par);
        const double *d = (const double *)(sm + (size_t)s * chunk);
        for (int k = threadIdx.x; k < chunk / 8; k += blockDim.x) acc += d[k];
        __syncwarp(); if (lane == 0) mb_arrive(empty + s);
        if (warp == 0 && i + nst < n) { mb_wait(empty + s, par); issue(i + nst, s); }
        if (++s == nst) { s = 0; par ^= 1; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// mode 1: plain vector loads, unrolled
__global__ void k_ldg(const double2 *in, double *out, size_t per_cta)
{
    const double2 *src = in + (size_t)blockIdx.x * (per_cta / 16);
    const size_t n = per_cta / 16;
    double acc = 0;
    for (size_t k = threadIdx.x; k + 7 * blockDim.x < n; k += 8 * blockDim.x) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(src + k + (size_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    const size_t total = (size_t)8 << 30;
    double *in, *out; cudaMalloc(&in, total); cudaMalloc(&out, 1 << 24); cudaMemset(in, 0, total);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int ctas_list[] = {148, 296, 592};
    for (int ci = 0; ci < 3; ++ci) {
        const int ctas = ctas_list[ci];
        const size_t per = (total / ctas) / (256 * 1024) * (256 * 1024);
        const int cfg[][3] = {{32768, 3, 1}, {32768, 6, 1}, {32768, 6, 32}, {16384, 6, 1}, {16384, 12, 1}, {8192, 6, 1}, {8192, 24, 1}, {4096, 12, 1}, {65536, 3, 1}};
        for (auto &c : cfg) {
            const int chunk = c[0], nst = c[1], split = c[2];
            const size_t smem = (size_t)nst * chunk + 2 * nst * 8;
            if (smem * (ctas / 148) > 220 * 1024) continue;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); k_bulk<<<ctas, 256, smem>>>(in, out, per, chunk, nst, split); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
            printf("bulk ctas %4d chunk %6d stages %2d split %2d : %7.1f GB/s  (%s)\n", ctas, chunk, nst, split, per * ctas / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); k_ldg<<<ctas, 512>>>((const double2 *)in, out, per); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        printf("ldg  ctas %4d threads 512 unroll 8                : %7.1f GB/s\n", ctas, per * ctas / best / 1e6);
    }
    return 0;
}
