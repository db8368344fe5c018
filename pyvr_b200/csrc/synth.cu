// synth.cu -- device-side synthetic volumes + fused scalar -> {s, nx, ny, nz} texel builder (sm_100a).
//
// SURVEY.md section 8 f-3.  The 2048^3 / 4096^3 configs cannot be built on the host and pushed through
// PCIe (68.7 GB / 550 GB packed), so the analytic shapes of the reference's input generator
// (create_sample_volume, pyvr/datasets/synthetic.py:10-106) are evaluated where they are needed and
// fused with compute_normal_volume (synthetic.py:109-122) and the texel pack:
//
//   synth_scalar_kernel   binary64 evaluation of the shape on linspace(-1, 1, size), rounded to binary32
//                         exactly where numpy's .astype(float32) rounds, for one slab of the block plus
//                         a 1-voxel halo (indices clamped to the volume, which is all the one-sided
//                         differences at the outer faces need);
//   normals_pack_kernel   np.gradient + /(norm + 1e-8) in unfused binary32 (the arithmetic of normals.cu)
//                         from the slab, written straight into the packed line/slot texel array.
//
// The block is processed in x-slabs so the binary32 scratch stays at ~1 GB whatever the volume size.
// Same meshgrid quirk as the reference: texel (ix, iy, iz) = data[ix, iy, iz] and the analytic "x" varies
// along numpy axis 1, i.e. shape(X = axis[iy], Y = axis[ix], Z = axis[iz]) (SURVEY.md trap 8).
// Both kernels are streaming: 7 scalar reads (L1/L2-served stencil) + one 8/16-byte texel write per voxel;
// the binary64 exp is the cost of the first kernel (B200 keeps full-rate FP64, unlike sm_103).
#include "common.cuh"

namespace pyvr {
namespace {

struct SynthGeom {
    int size;            // the volume is size^3
    double step;         // linspace step 2/(size-1)
    int shape;           // PYVR_SHAPE_*
};

__device__ __forceinline__ double axis_value(const SynthGeom &g, int i) {
    // np.linspace(-1, 1, size): arange * step + start, last element forced to stop
    return i == g.size - 1 ? 1.0 : (double)i * g.step + -1.0;
}

__device__ __forceinline__ double gauss(double d) { return exp(-(d * d)); }

__device__ __forceinline__ float shape_value(const SynthGeom &g, int ix, int iy, int iz) {
    const double x = axis_value(g, iy), y = axis_value(g, ix), z = axis_value(g, iz);   // 'xy' meshgrid
    double v;
    if (g.shape == PYVR_SHAPE_SPHERE) {
        v = gauss(sqrt(x * x + y * y + z * z) * 3);
    } else if (g.shape == PYVR_SHAPE_TORUS) {
        const double ring = sqrt(x * x + y * y) - 0.6;
        v = gauss(sqrt(ring * ring + z * z) / 0.3 * 4);
    } else {   // PYVR_SHAPE_DOUBLE_SPHERE
        const double xm = x - 0.3, xp = x + 0.3;
        const double a = gauss(sqrt(xm * xm + y * y + z * z) * 4);
        const double b = gauss(sqrt(xp * xp + y * y + z * z) * 4);
        v = fmax(a, b);
    }
    return (float)v;
}

// Slab scratch: (sx, sy, sz) floats covering global indices [x0-1, x0-1+sx) x [y0-1, ...) x [z0-1, ...),
// evaluated at indices clamped into the volume.
__global__ void __launch_bounds__(256)
synth_scalar_kernel(float *__restrict__ slab, SynthGeom g, int x0, int y0, int z0, int sx, int sy, int sz) {
    const long long total = (long long)sx * sy * sz;
    for (long long at = (long long)blockIdx.x * blockDim.x + threadIdx.x; at < total;
         at += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(at % sz);
        const long long r = at / sz;
        const int j = (int)(r % sy), i = (int)(r / sy);
        const int gx = min(max(x0 - 1 + i, 0), g.size - 1), gy = min(max(y0 - 1 + j, 0), g.size - 1),
                  gz = min(max(z0 - 1 + k, 0), g.size - 1);
        slab[at] = shape_value(g, gx, gy, gz);
    }
}

__device__ __forceinline__ float diff1(float lo, float mid, float hi, int idx, int n) {
    if (n < 2) return 0.0f;
    if (idx == 0) return hi - mid;
    if (idx == n - 1) return mid - lo;
    return (hi - lo) / 2.0f;
}

// One thread per stored texel of the slab's interior (nx_slab x ny x nz local texels starting at local
// x index lx0); the slab holds their values and halo.
template <bool HALF>
__global__ void __launch_bounds__(256)
normals_pack_kernel(const float *__restrict__ slab, VolumeDesc v, int gsize, int lx0, int nx_slab, int sy, int sz) {
    const int ny = v.n[1], nz = v.n[2];
    const long long total = (long long)nx_slab * ny * nz;
    const long long s0 = (long long)sy * sz, s1 = sz;
    for (long long flat = (long long)blockIdx.x * blockDim.x + threadIdx.x; flat < total;
         flat += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(flat % nz);
        const long long r = flat / nz;
        const int iy = (int)(r % ny), ixs = (int)(r / ny);
        const long long at = (long long)(ixs + 1) * s0 + (long long)(iy + 1) * s1 + (iz + 1);
        const float c = slab[at];
        const int gx = v.org[0] + lx0 + ixs, gy = v.org[1] + iy, gz = v.org[2] + iz;
        const float g0 = diff1(slab[at - s0], c, slab[at + s0], gx, gsize);
        const float g1 = diff1(slab[at - s1], c, slab[at + s1], gy, gsize);
        const float g2 = diff1(slab[at - 1], c, slab[at + 1], gz, gsize);
        const float norm = sqrtf(g0 * g0 + g1 * g1 + g2 * g2) + 1e-8f;
        const float nx = g0 / norm, nyv = g1 / norm, nzv = g2 / norm;
        // plain layout: one slot.  z-pair layout: this texel is the first half of entry(iz), the second half
        // of entry(iz - 1) and, at the top of the block, also the second half of its own entry.
        long long dsts[3];
        int n_dst = 0;
        if (!v.pair) {
            dsts[n_dst++] = texel_index(v, lx0 + ixs, iy, iz);
        } else {
            dsts[n_dst++] = texel_index(v, lx0 + ixs, iy, iz) * 2;
            if (iz > 0) dsts[n_dst++] = texel_index(v, lx0 + ixs, iy, iz - 1) * 2 + 1;
            if (iz == nz - 1) dsts[n_dst++] = texel_index(v, lx0 + ixs, iy, iz) * 2 + 1;
        }
        for (int d = 0; d < n_dst; ++d) {
            if constexpr (HALF) {
                __half2 lo = __floats2half2_rn(c, nx), hi = __floats2half2_rn(nyv, nzv);
                uint2 raw;
                raw.x = *reinterpret_cast<unsigned *>(&lo);
                raw.y = *reinterpret_cast<unsigned *>(&hi);
                reinterpret_cast<uint2 *>(const_cast<void *>(v.texels))[dsts[d]] = raw;
            } else {
                reinterpret_cast<float4 *>(const_cast<void *>(v.texels))[dsts[d]] = make_float4(c, nx, nyv, nzv);
            }
        }
    }
}

// Test/debug aid: packed texels back to the reference's two arrays (numpy C order of the stored block).
template <bool HALF>
__global__ void __launch_bounds__(256)
unpack_texels_kernel(VolumeDesc v, float *__restrict__ scalar, float *__restrict__ normals) {
    const long long total = (long long)v.n[0] * v.n[1] * v.n[2];
    for (long long flat = (long long)blockIdx.x * blockDim.x + threadIdx.x; flat < total;
         flat += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(flat % v.n[2]);
        const long long r = flat / v.n[2];
        const int iy = (int)(r % v.n[1]), ix = (int)(r / v.n[1]);
        const long long at = texel_index(v, ix, iy, iz) << v.pair;
        float4 t;
        if constexpr (HALF) {
            const uint2 raw = reinterpret_cast<const uint2 *>(v.texels)[at];
            const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
            t = make_float4(a.x, a.y, b.x, b.y);
        } else {
            t = reinterpret_cast<const float4 *>(v.texels)[at];
        }
        if (scalar) scalar[flat] = t.x;
        if (normals) { normals[3 * flat] = t.y; normals[3 * flat + 1] = t.z; normals[3 * flat + 2] = t.w; }
    }
}

// Raw copy of the stored texels of x-planes [x0, x0 + nx) into plain [x][y][z] order.
template <typename T>
__global__ void __launch_bounds__(256)
linearize_texels_kernel(VolumeDesc v, int x0, int nx, T *__restrict__ dst) {
    const long long total = (long long)nx * v.n[1] * v.n[2];
    for (long long flat = (long long)blockIdx.x * blockDim.x + threadIdx.x; flat < total;
         flat += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(flat % v.n[2]);
        const long long r = flat / v.n[2];
        const int iy = (int)(r % v.n[1]), ix = (int)(r / v.n[1]);
        dst[flat] = reinterpret_cast<const T *>(v.texels)[texel_index(v, x0 + ix, iy, iz) << v.pair];
    }
}

inline int grid_for(long long work) {
    long long g = (work + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

cudaError_t launch_synth_volume(const VolumeDesc &vol, bool half_texels, int shape, int size, float *scratch,
                                size_t scratch_floats, cudaStream_t stream) {
    SynthGeom g{size, 2.0 / (double)(size - 1), shape};
    const int sy = vol.n[1] + 2, sz = vol.n[2] + 2;
    const long long plane = (long long)sy * sz;
    int planes = (int)(scratch_floats / (size_t)plane) - 2;   // interior x-planes per slab
    if (planes < 1) return cudaErrorInvalidValue;
    if (planes > vol.n[0]) planes = vol.n[0];
    for (int lx0 = 0; lx0 < vol.n[0]; lx0 += planes) {
        const int nxs = vol.n[0] - lx0 < planes ? vol.n[0] - lx0 : planes;
        synth_scalar_kernel<<<grid_for((long long)(nxs + 2) * plane), 256, 0, stream>>>(
            scratch, g, vol.org[0] + lx0, vol.org[1], vol.org[2], nxs + 2, sy, sz);
        const long long work = (long long)nxs * vol.n[1] * vol.n[2];
        if (half_texels) normals_pack_kernel<true><<<grid_for(work), 256, 0, stream>>>(scratch, vol, size, lx0, nxs, sy, sz);
        else normals_pack_kernel<false><<<grid_for(work), 256, 0, stream>>>(scratch, vol, size, lx0, nxs, sy, sz);
    }
    return cudaGetLastError();
}

cudaError_t launch_linearize_texels(const VolumeDesc &vol, bool half_texels, int x0, int nx, void *dst, cudaStream_t stream) {
    const long long total = (long long)nx * vol.n[1] * vol.n[2];
    if (half_texels) linearize_texels_kernel<uint2><<<grid_for(total), 256, 0, stream>>>(vol, x0, nx, reinterpret_cast<uint2 *>(dst));
    else linearize_texels_kernel<float4><<<grid_for(total), 256, 0, stream>>>(vol, x0, nx, reinterpret_cast<float4 *>(dst));
    return cudaGetLastError();
}

cudaError_t launch_unpack_texels(const VolumeDesc &vol, bool half_texels, float *scalar, float *normals,
                                 cudaStream_t stream) {
    const long long total = (long long)vol.n[0] * vol.n[1] * vol.n[2];
    if (half_texels) unpack_texels_kernel<true><<<grid_for(total), 256, 0, stream>>>(vol, scalar, normals);
    else unpack_texels_kernel<false><<<grid_for(total), 256, 0, stream>>>(vol, scalar, normals);
    return cudaGetLastError();
}

}  // namespace pyvr
