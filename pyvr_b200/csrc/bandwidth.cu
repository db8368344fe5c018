// bandwidth.cu -- measured denominators for the march kernel's roofline (no reference counterpart).
//
// SURVEY.md section 8(d) defines the gather roofline as algorithmic bytes / "peak L2->SM read bandwidth", and the
// march is bound by the L1 data path (each lane must receive 8 texels = 128 B per sample through the SM's
// 128 B/clk load-return path).  MEASURED_PEAKS.json only carries an HBM copy rate, so the two cache levels are
// measured live, on the GPU the bench runs on, by the simplest kernels that can saturate them:
//   level 1  every CTA re-reads its own 16 KiB window with coalesced LDG.128 (L1 hits after the first pass):
//            the SM's load-return bandwidth, the unit that bounds a gather served from L1;
//   level 2  every warp streams coalesced LDG.128 with .cg (L1 bypass) over a 64 MiB buffer that fits the
//            126 MB L2 several times over: the L2 -> SM fabric bandwidth.
// Both report bytes delivered to registers / CUDA-event time, best of `reps` launches.
#include "common.cuh"

namespace pyvr {
namespace {

constexpr int BW_THREADS = 256;

__global__ void __launch_bounds__(BW_THREADS)
l1_read_kernel(const float4 *__restrict__ buf, int window, int iters, float *sink) {
    const float4 *w = buf + (size_t)blockIdx.x * window;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int at = threadIdx.x;
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float4 v;
            asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(w + at));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            at = (at + BW_THREADS) & (window - 1);
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;   // keep the loads alive
}

__global__ void __launch_bounds__(BW_THREADS)
l2_read_kernel(const float4 *__restrict__ buf, size_t n, int iters, float *sink) {
    const size_t stride = (size_t)gridDim.x * BW_THREADS;
    size_t at = (size_t)blockIdx.x * BW_THREADS + threadIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + at));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            at += stride;
            if (at >= n) at -= n;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;
}

// level 3: the DRAM rate of K2's own traffic mix -- 4 bytes read and 12 bytes written per element, both streaming,
// no arithmetic: one float4 in, three float4 out per thread (the same 1536-byte store runs per warp as normals.cu).
// A stream that is 3/4 writes does not reach the 1:1 copy rate MEASURED_PEAKS.json quotes; this is the ceiling the
// normals kernel can be held against.
__global__ void __launch_bounds__(BW_THREADS)
mix13_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * BW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * BW_THREADS) {
        const float4 v = __ldcs(in + i);
        __stcs(out + 3 * i + 0, make_float4(v.x, v.y, v.z, v.w));
        __stcs(out + 3 * i + 1, make_float4(v.y, v.z, v.w, v.x));
        __stcs(out + 3 * i + 2, make_float4(v.z, v.w, v.x, v.y));
    }
}

}  // namespace

// level: 1 = L1 load-return, 2 = L2 -> SM, 3 = DRAM with K2's 1:3 read:write mix.  Returns GB/s (1e9 bytes) in *gbs.
cudaError_t measure_cache_bandwidth(int level, double *gbs) {
    int sms = 0, dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (level == 3) {
        const size_t n = ((size_t)512 << 20) / sizeof(float4);      // 512 MiB in, 1.5 GiB out: far beyond L2
        float4 *in = nullptr, *out = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        e = cudaMalloc(&in, n * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&out, 3 * n * sizeof(float4));
        if (e == cudaSuccess) e = cudaMemset(in, 0, n * sizeof(float4));
        if (e == cudaSuccess) e = cudaEventCreate(&e0);
        if (e == cudaSuccess) e = cudaEventCreate(&e1);
        double best = 0.0;
        for (int rep = 0; rep < 9 && e == cudaSuccess; ++rep) {      // rep 0 warms up; three grid sizes, best of all
            const int ctas_per_sm = rep % 3 == 0 ? 8 : rep % 3 == 1 ? 16 : 32;
            e = cudaEventRecord(e0, 0);
            mix13_kernel<<<sms * ctas_per_sm, BW_THREADS>>>(in, out, n);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaEventRecord(e1, 0);
            if (e == cudaSuccess) e = cudaEventSynchronize(e1);
            float ms = 0.0f;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
            if (e == cudaSuccess && rep > 0 && ms > 0.0f) {
                const double rate = (double)n * 64.0 / (ms * 1e-3) / 1e9;
                if (rate > best) best = rate;
            }
        }
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        cudaFree(in);
        cudaFree(out);
        if (e == cudaSuccess && gbs) *gbs = best;
        return e;
    }
    const int grid = sms * 8;                       // 8 CTAs x 8 warps per SM
    const int window = 1024;                        // float4 per CTA window = 16 KiB (8 windows = 128 KiB per SM)
    const size_t n = level == 1 ? (size_t)grid * window : ((size_t)64 << 20) / sizeof(float4);
    const int iters = level == 1 ? 2048 : 256;
    float4 *buf = nullptr;
    float *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    e = cudaMalloc(&buf, n * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&sink, sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(buf, 0, n * sizeof(float4));
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6 && e == cudaSuccess; ++rep) {   // rep 0 warms the cache
        e = cudaEventRecord(e0, 0);
        if (level == 1) l1_read_kernel<<<grid, BW_THREADS>>>(buf, window, iters, sink);
        else l2_read_kernel<<<grid, BW_THREADS>>>(buf, n, iters, sink);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(e1, 0);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0.0f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)grid * BW_THREADS * (double)iters * 8.0 * 16.0;
        if (e == cudaSuccess && rep > 0 && ms > 0.0f) {
            const double rate = bytes / (ms * 1e-3) / 1e9;
            if (rate > best) best = rate;
        }
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    if (e == cudaSuccess && gbs) *gbs = best;
    return e;
}

}  // namespace pyvr
