// platform.cuh — thin layer between the kernels/orchestration and the CUDA runtime.
//
// Product build (nvcc, sm_100a): everything maps 1:1 onto the CUDA runtime.
// With -DGSB200_EMULATE (used ONLY by tests/emul/, never shipped, never loaded by the
// package) the same kernel bodies are interpreted thread by thread on the host so that the
// index logic can be validated in the GPU-less build container.  It is a test harness for
// the kernels' control flow, not a fallback: the product library has no such path and
// fails with GSB200_ENODEVICE when no B200 is present.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>

namespace gsb {
typedef long long i64;
void set_error(const char *fmt, ...);
}

#ifndef GSB200_EMULATE
// ------------------------------------------------------------------ real CUDA
#include <cuda_runtime.h>
#define GSB_GLOBAL static __global__     // internal linkage: kernels.cuh is included by several translation units
#define GSB_DEVICE __device__ __forceinline__
#define GSB_HD __host__ __device__ __forceinline__
#define GSB_MEMBER __device__ __forceinline__
#define GSB_LAUNCH(kernel, grid, block, stream, ...) \
    do { if (!gsb::dry_run()) { kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__); gsb::note_launch(); } } while (0)
// kernels whose CTA cooperates through shared memory (the interpreter build runs them block by block)
#define GSB_LAUNCH_CTA(kernel, grid, block, stream, ...) GSB_LAUNCH(kernel, grid, block, stream, __VA_ARGS__)
namespace gsb {
void note_launch();
bool dry_run();   // planning pass of assemble(): walk the launch sequence without launching
typedef cudaStream_t stream_t;
typedef cudaEvent_t event_t;
inline int dev_check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (e == cudaErrorMemoryAllocation) ? -5 : -4;
}
// Device memory comes from a PRIVATE stream-ordered pool per device (cudaMemPoolCreate; the process's default pool and whoever else
// uses it - e.g. torch's allocator - are left alone) with the release threshold raised, so that the multi-GB buffers of one assembly
// (pattern, values, workspace) are recycled by the next one instead of going back to the driver (a cudaMalloc/cudaFree pair of that
// size costs milliseconds).  Callers synchronise their stream before freeing (dev_free is ordered on the pool's own stream only).
// dev_pool_idle() = bytes the pool holds but nobody uses: the workspace budget counts them as available, and they are given back
// when an allocation fails (dev_malloc) or on request (dev_trim = gsb200_trim).
// GSB200_NO_POOL=1 selects plain cudaMalloc/cudaFree.
struct DevPoolState { cudaStream_t stream; cudaMemPool_t pool; bool init, usable; int users; };
inline DevPoolState &dev_pool()
{
    static DevPoolState st[64];
    int d = 0; cudaGetDevice(&d); if (d < 0 || d >= 64) d = 0;
    DevPoolState &P = st[d];
    if (!P.init) {
        P.init = true; P.usable = false; P.users = 0;
        int sup = 0;
        if (!getenv("GSB200_NO_POOL") && cudaDeviceGetAttribute(&sup, cudaDevAttrMemoryPoolsSupported, d) == cudaSuccess && sup) {
            cudaMemPoolProps props; memset(&props, 0, sizeof props);
            props.allocType = cudaMemAllocationTypePinned; props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice; props.location.id = d;
            if (cudaMemPoolCreate(&P.pool, &props) == cudaSuccess && cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking) == cudaSuccess) {
                unsigned long long thr = ~0ull;
                P.usable = cudaMemPoolSetAttribute(P.pool, cudaMemPoolAttrReleaseThreshold, &thr) == cudaSuccess;
            }
        }
        cudaGetLastError();
    }
    return P;
}
inline int dev_malloc(void **p, size_t n)
{
    DevPoolState &P = dev_pool();
    if (!P.usable) return dev_check(cudaMalloc(p, n ? n : 1), "cudaMalloc");
    cudaError_t e = cudaMallocFromPoolAsync(p, n ? n : 1, P.pool, P.stream);
    if (e != cudaSuccess) { cudaGetLastError(); cudaStreamSynchronize(P.stream); cudaMemPoolTrimTo(P.pool, 0); e = cudaMallocFromPoolAsync(p, n ? n : 1, P.pool, P.stream); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(P.stream);
    return dev_check(e, "cudaMallocFromPoolAsync");
}
inline void dev_free(void *p) { if (!p) return; DevPoolState &P = dev_pool(); if (P.usable) cudaFreeAsync(p, P.stream); else cudaFree(p); }
inline void dev_trim() { DevPoolState &P = dev_pool(); if (P.usable) { cudaStreamSynchronize(P.stream); cudaMemPoolTrimTo(P.pool, 0); } }
inline size_t dev_pool_idle()
{
    DevPoolState &P = dev_pool();
    if (!P.usable) return 0;
    cudaStreamSynchronize(P.stream);
    unsigned long long res = 0, used = 0;
    if (cudaMemPoolGetAttribute(P.pool, cudaMemPoolAttrReservedMemCurrent, &res) != cudaSuccess || cudaMemPoolGetAttribute(P.pool, cudaMemPoolAttrUsedMemCurrent, &used) != cudaSuccess) { cudaGetLastError(); return 0; }
    return res > used ? (size_t)(res - used) : 0;
}
inline int dev_h2d(void *d, const void *h, size_t n, stream_t s) { return dev_check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "H2D copy"); }
inline int dev_d2h(void *h, const void *d, size_t n, stream_t s) {
    int r = dev_check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "D2H copy");
    return r ? r : dev_check(cudaStreamSynchronize(s), "stream sync");
}
inline int dev_d2d(void *d, const void *s_, size_t n, stream_t s) { return dev_check(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s), "D2D copy"); }
inline int dev_memset(void *d, int v, size_t n, stream_t s) { return dev_check(cudaMemsetAsync(d, v, n, s), "memset"); }
inline int dev_sync(stream_t s) { return dev_check(cudaStreamSynchronize(s), "stream sync"); }
inline int dev_last_error(const char *what) { return dev_check(cudaGetLastError(), what); }
GSB_DEVICE double ld_stream(const double *p) { return __ldcs(p); }     // read-once data: evict first
GSB_DEVICE double ld_keep(const double *p) { return __ldg(p); }
GSB_DEVICE double2 ld_keep2(const double2 *p) { return __ldg(p); }
GSB_DEVICE void st_stream(double *p, double v) { __stcs(p, v); }
GSB_DEVICE void atomic_add(double *p, double v) { atomicAdd(p, v); }
GSB_DEVICE int atomic_add(int *p, int v) { return atomicAdd(p, v); }
GSB_DEVICE int popc(unsigned v) { return __popc(v); }
}
#else
// ------------------------------------------------------------------ host interpreter (tests only)
#define GSB_GLOBAL static
#define GSB_DEVICE static inline
#define GSB_HD static inline
#define GSB_MEMBER inline
struct gsb_dim3 { unsigned x, y, z; gsb_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef gsb_dim3 dim3;
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static gsb_dim3 blockIdx, threadIdx, blockDim, gridDim;
#define GSB_LAUNCH(kernel, grid, block, stream, ...)                                         \
    do { if (gsb::dry_run()) break; gridDim = gsb_dim3(grid); blockDim = gsb_dim3(block);     \
         for (unsigned bz_ = 0; bz_ < gridDim.z; ++bz_) for (unsigned by_ = 0; by_ < gridDim.y; ++by_) \
         for (unsigned bx_ = 0; bx_ < gridDim.x; ++bx_) for (unsigned tz_ = 0; tz_ < blockDim.z; ++tz_) \
         for (unsigned ty_ = 0; ty_ < blockDim.y; ++ty_) for (unsigned tx_ = 0; tx_ < blockDim.x; ++tx_) { \
             blockIdx = gsb_dim3(bx_, by_, bz_); threadIdx = gsb_dim3(tx_, ty_, tz_); kernel(__VA_ARGS__); } \
         gsb::note_launch(); } while (0)
// CTA-cooperative kernels: ONE call per block, the kernel body loops over its threads phase by phase (GSB_THREADS in fused.cuh)
#define GSB_LAUNCH_CTA(kernel, grid, block, stream, ...)                                     \
    do { if (gsb::dry_run()) break; gridDim = gsb_dim3(grid); blockDim = gsb_dim3(block);     \
         for (unsigned bz_ = 0; bz_ < gridDim.z; ++bz_) for (unsigned by_ = 0; by_ < gridDim.y; ++by_) \
         for (unsigned bx_ = 0; bx_ < gridDim.x; ++bx_) { blockIdx = gsb_dim3(bx_, by_, bz_); threadIdx = gsb_dim3(0, 0, 0); kernel(__VA_ARGS__); } \
         gsb::note_launch(); } while (0)
namespace gsb {
void note_launch();
bool dry_run();
typedef int stream_t;
typedef int event_t;
inline int dev_malloc(void **p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? 0 : -5; }
inline void dev_free(void *p) { std::free(p); }
inline void dev_trim() {}
inline size_t dev_pool_idle() { return 0; }
inline int dev_h2d(void *d, const void *h, size_t n, stream_t) { std::memcpy(d, h, n); return 0; }
inline int dev_d2h(void *h, const void *d, size_t n, stream_t) { std::memcpy(h, d, n); return 0; }
inline int dev_d2d(void *d, const void *s_, size_t n, stream_t) { std::memcpy(d, s_, n); return 0; }
inline int dev_memset(void *d, int v, size_t n, stream_t) { std::memset(d, v, n); return 0; }
inline int dev_sync(stream_t) { return 0; }
inline int dev_last_error(const char *) { return 0; }
static inline double ld_stream(const double *p) { return *p; }
static inline double ld_keep(const double *p) { return *p; }
static inline double2 ld_keep2(const double2 *p) { return *p; }
static inline void st_stream(double *p, double v) { *p = v; }
static inline void atomic_add(double *p, double v) { *p += v; }
static inline int atomic_add(int *p, int v) { int o = *p; *p += v; return o; }
static inline int popc(unsigned v) { return __builtin_popcount(v); }
}
#endif
