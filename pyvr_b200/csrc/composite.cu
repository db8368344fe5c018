// composite.cu -- sort-last compositing of partial images (sm_100a).
//
// No reference counterpart: the reference renders one volume on one device.  When the volume is split
// into bricks across GPUs (SURVEY.md section 8 e, config C5), every GPU marches its brick on the GLOBAL
// sample lattice and produces the pre-blend fragment colour of its samples alone -- premultiplied rgb and
// alpha, exactly what volume.frag.glsl:112-115 accumulates.  Front-to-back accumulation is associative:
//     acc(front ++ back) = acc(front) + (1 - acc_a(front)) * acc(back),
// so partial images are merged with `over` in visibility order (binary swap across ranks,
// pyvr_b200/multi_gpu.py), and only the final image goes through the blend + RGBA8 quantisation of
// manager.py:217-220,29 (fragment_to_rgba8, shared with the march epilogue).
//
// The shader stops a ray once accumulated alpha reaches 0.99 (volume.frag.glsl:87).  Each brick applies
// that rule to its own segment (it cannot see the alpha accumulated in front of it); `over` applies it
// again at brick granularity: a front image at or above the threshold hides the back image, and when the
// merge would cross the threshold -- the reference stops somewhere INSIDE the back brick -- only the
// fraction s of the back image that brings alpha to the threshold is added (first-order model: colour and
// alpha of the back segment grow in proportion).  The reference ends such a ray within one sample's
// contribution above 0.99, so saturating pixels agree to about one RGBA8 level (measured by the tests;
// without the clip it is up to 5).  The relay mode (pyvr_b200/multi_gpu.py) is exact when that matters.
//
// HBM/NVLink-bound streaming kernels: 32 algorithmic bytes read + 16 written per pixel for `over`
// (`back` may be a peer-mapped pointer: the load then crosses NVLink and the merge overlaps the transfer),
// 16 read + 4 written for finalize.  One float4 per thread, fully coalesced.
#include "common.cuh"

namespace pyvr {
namespace {

// `over` of one pixel (shared by the plain and the fused kernels)
__device__ __forceinline__ float4 over_pixel(float4 f, const float4 *back, size_t p, float term_alpha) {
    if (f.w < term_alpha) {
        const float4 b = __ldcs(back + p);
        float t = 1.0f - f.w;
        if (fmaf(t, b.w, f.w) > term_alpha) t = (term_alpha - f.w) / b.w;
        f.x = fmaf(t, b.x, f.x);
        f.y = fmaf(t, b.y, f.y);
        f.z = fmaf(t, b.z, f.z);
        f.w = fmaf(t, b.w, f.w);
    }
    return f;
}

__global__ void __launch_bounds__(256)
composite_over_kernel(const float4 *__restrict__ front, const float4 *__restrict__ back,
                      float4 *__restrict__ out, size_t n, float term_alpha) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const float4 f = over_pixel(front[p], back, p, term_alpha);   // streaming load of `back`: a peer buffer is read once
        out[p] = f;
    }
}

// Last round of the binary swap fused with the blend + RGBA8 quantisation: the merged floats never travel
// through memory again (out8 may be a peer-mapped frame: 4 bytes per pixel cross NVLink instead of a later gather).
__global__ void __launch_bounds__(256)
composite_finalize_kernel(const float4 *__restrict__ front, const float4 *__restrict__ back, float4 *accum_out,
                          uchar4 *__restrict__ out8, size_t n, float term_alpha, unsigned flags, int front_is_streamed) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const float4 f = over_pixel(front_is_streamed ? __ldcs(front + p) : front[p], back, p, term_alpha);
        if (accum_out) accum_out[p] = f;
        out8[p] = fragment_to_rgba8(f.x, f.y, f.z, f.w, flags);
    }
}

// Stream-ordered flags between GPUs.  signal: everything this stream did before (kernels that wrote peer memory)
// is made visible system-wide, then the flag is stored with release semantics at system scope.  wait: one thread
// per flag spins with an acquire load until it reaches `value` (monotonic counters: frame number * stages + stage).
__global__ void flag_signal_kernel(unsigned *flag, unsigned value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

struct FlagList {
    unsigned *flag[16];
    int n;
};
__global__ void flag_signal_many_kernel(FlagList list, unsigned value) {
    __threadfence_system();
    if ((int)threadIdx.x < list.n)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(list.flag[threadIdx.x]), "r"(value) : "memory");
}

__global__ void flag_wait_kernel(const unsigned *flags, int n_flags, unsigned value) {
    if ((int)threadIdx.x < n_flags) {
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
            if ((int)(v - value) < 0) __nanosleep(200);
        } while ((int)(v - value) < 0);
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(256)
finalize_rgba8_kernel(const float4 *__restrict__ accum, uchar4 *__restrict__ out, size_t n, unsigned flags) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const float4 c = accum[p];
        out[p] = fragment_to_rgba8(c.x, c.y, c.z, c.w, flags);
    }
}

inline int stream_grid(size_t n) {
    size_t g = (n + 255) / 256;
    const size_t cap = 148 * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

cudaError_t launch_composite_over(const float4 *front, const float4 *back, float4 *out, size_t n_pixels,
                                  float term_alpha, cudaStream_t stream) {
    composite_over_kernel<<<stream_grid(n_pixels), 256, 0, stream>>>(front, back, out, n_pixels, term_alpha);
    return cudaGetLastError();
}

cudaError_t launch_composite_finalize(const float4 *front, const float4 *back, float4 *accum_out, uchar4 *out8,
                                      size_t n_pixels, float term_alpha, unsigned flags, cudaStream_t stream) {
    composite_finalize_kernel<<<stream_grid(n_pixels), 256, 0, stream>>>(front, back, accum_out, out8, n_pixels, term_alpha,
                                                                         flags, 0);
    return cudaGetLastError();
}

cudaError_t launch_flag_signal(unsigned *flag, unsigned value, cudaStream_t stream) {
    flag_signal_kernel<<<1, 1, 0, stream>>>(flag, value);
    return cudaGetLastError();
}

cudaError_t launch_flag_signal_many(unsigned *const *flags, int n, unsigned value, cudaStream_t stream) {
    if (n < 0) return cudaErrorInvalidValue;
    for (int first = 0; first < n; first += 16) {      // one launch per 16 peers
        FlagList list{};
        list.n = n - first < 16 ? n - first : 16;
        for (int i = 0; i < list.n; ++i) list.flag[i] = flags[first + i];
        flag_signal_many_kernel<<<1, 32, 0, stream>>>(list, value);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_flag_wait(const unsigned *flags, int n_flags, unsigned value, cudaStream_t stream) {
    if (n_flags < 1 || n_flags > 1024) return cudaErrorInvalidValue;
    flag_wait_kernel<<<1, ((n_flags + 31) / 32) * 32, 0, stream>>>(flags, n_flags, value);
    return cudaGetLastError();
}

cudaError_t launch_finalize_rgba8(const float4 *accum, uchar4 *out, size_t n_pixels, unsigned flags,
                                  cudaStream_t stream) {
    finalize_rgba8_kernel<<<stream_grid(n_pixels), 256, 0, stream>>>(accum, out, n_pixels, flags);
    return cudaGetLastError();
}

}  // namespace pyvr
