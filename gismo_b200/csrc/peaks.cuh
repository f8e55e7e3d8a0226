// peaks.cuh — roofline denominators MEASURED_PEAKS.json does not carry:
// FP64 FMA throughput (vector DFMA and DMMA m8n8k4) and a plain HBM copy.
#pragma once
#include "platform.cuh"

namespace gsb {

__global__ void k_peak_dfma(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_peak_dmma(double *out, int iters, double a, double b)
{
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    const double fa = a + threadIdx.x * 1e-9, fb = b;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(fa), "d"(fb));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

__global__ void k_peak_copy(const double2 *__restrict__ in, double2 *__restrict__ out, i64 n)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) out[i] = in[i];
}

} // namespace gsb

extern "C" int gsb200_measure_peaks(int device, double *fp64_tflops, double *dmma_tflops, double *hbm_gbs)
{
    using namespace gsb;
    GSB_TRY(select_device(device));
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device);
    const int nsm = prop.multiProcessorCount, blocks = nsm * 8, threads = 256, iters = 1 << 14;
    double *buf = 0;
    GSB_TRY(dev_malloc((void **)&buf, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, ms = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0); k_peak_dfma<<<blocks, threads>>>(buf, iters, 0.999999, 1e-7); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    if (fp64_tflops) *fp64_tflops = 2.0 * 8 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0); k_peak_dmma<<<blocks, threads>>>(buf, iters, 0.5, 0.25); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    // one m8n8k4 = 256 FMA per warp
    if (dmma_tflops) *dmma_tflops = 2.0 * 256 * 4 * iters * (double)blocks * (threads / 32) / (best * 1e-3) / 1e12;
    dev_free(buf);
    const i64 n = (i64)1 << 27;   // 2 GiB in + 2 GiB out
    double2 *a = 0, *b = 0;
    GSB_TRY(dev_malloc((void **)&a, sizeof(double2) * (size_t)n)); GSB_TRY(dev_malloc((void **)&b, sizeof(double2) * (size_t)n));
    cudaMemset(a, 0, sizeof(double2) * (size_t)n);
    best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0); k_peak_copy<<<nsm * 16, 512>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    if (hbm_gbs) *hbm_gbs = 2.0 * sizeof(double2) * (double)n / (best * 1e-3) / 1e9;
    dev_free(a); dev_free(b);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return dev_last_error("peak kernels");
}
