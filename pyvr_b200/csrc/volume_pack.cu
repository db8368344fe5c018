// volume_pack.cu -- device-side preparation of the volume the march kernel samples.
//
//  * pack_texels: fuses the reference's two textures (R32F scalar, RGB32F normal;
//    pyvr/moderngl_renderer/manager.py:95-101,123-129) into one interleaved texel {s,nx,ny,nz}
//    (binary32 or binary16) stored in the line/slot layout of VolumeDesc (common.cuh).
//    A Volume without normals is packed with normal = (s, 0, 0): that is what the shader samples when
//    `normal_volume` is left on texture unit 0 (renderer.py:143-146).
//  * cell_minmax / cell_classify / cell_dist_relax: the macrocell grid for exact empty-space skipping.
//    A cell is inactive only if every sample whose lower taps fall in it provably has alpha_tf == 0.
//    Every cell then gets the chessboard radius of the largest cube of like cells around it, so the march
//    crosses empty space AND solid interiors with one map lookup per cube instead of one per cell.
#include "common.cuh"

namespace pyvr {
namespace {

// One thread per source voxel, z (memory-fastest in the source) across threadIdx.x: reads are
// coalesced; writes fill whole 128-byte lines (a z-run of SLOTS texels, permuted by the slot swizzle).
template <bool HALF>
__global__ void __launch_bounds__(256)
pack_texels_kernel(const float *__restrict__ scalar, const float *__restrict__ normals, VolumeDesc v) {
    const int nz = v.n[2], ny = v.n[1], nx = v.n[0];
    const long long total = (long long)nx * ny * nz;
    for (long long flat = (long long)blockIdx.x * blockDim.x + threadIdx.x; flat < total;
         flat += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(flat % nz);
        const long long rest = flat / nz;
        const int iy = (int)(rest % ny), ix = (int)(rest / ny);
        const long long at = texel_index(v, ix, iy, iz);
        // z-pair entries hold this texel and its upper z neighbour (clamped at the top of the block)
        for (int e = 0; e <= v.pair; ++e) {
            const long long src = e && iz + 1 < nz ? flat + 1 : flat;
            const float s = scalar[src];
            float a, b, c;
            if (normals) {
                a = normals[3 * src + 0]; b = normals[3 * src + 1]; c = normals[3 * src + 2];
            } else {
                a = s; b = 0.0f; c = 0.0f;
            }
            const long long dst = (at << v.pair) + e;
            if constexpr (HALF) {
                __half2 lo = __floats2half2_rn(s, a), hi = __floats2half2_rn(b, c);
                uint2 raw;
                raw.x = *reinterpret_cast<unsigned *>(&lo);
                raw.y = *reinterpret_cast<unsigned *>(&hi);
                reinterpret_cast<uint2 *>(const_cast<void *>(v.texels))[dst] = raw;
            } else {
                reinterpret_cast<float4 *>(const_cast<void *>(v.texels))[dst] = make_float4(s, a, b, c);
            }
        }
    }
}

// Apron fill, one axis at a time (z, then y, then x), so that edges and corners end up replicated too.
// AXIS 2: entries (ix, iy, -1) and (ix, iy, n2) for the interior (ix, iy); AXIS 1: (ix, -1 | n1, iz) for
// interior ix and every iz in [-1, n2]; AXIS 0: (-1 | n0, iy, iz) for every iy, iz including their aprons.
// A z-pair entry (ix, iy, iz) holds {T(cz(iz)), T(cz(iz + 1))} with cz = clamp to [0, n2 - 1]: copying the
// entry of the clamped (ix, iy) at the SAME iz is right for axes 0 and 1; on axis 2 entry(-1) = {T(0), T(0)}
// takes the first half of entry(0) twice and entry(n2) = entry(n2 - 1) = {T(n2-1), T(n2-1)}.
template <typename ENTRY, int AXIS>
__global__ void __launch_bounds__(256)
fill_apron_kernel(VolumeDesc v) {
    ENTRY *e = reinterpret_cast<ENTRY *>(const_cast<void *>(v.texels));
    const int n0 = v.n[0], n1 = v.n[1], n2 = v.n[2];
    const long long span_a = AXIS == 2 ? n0 : AXIS == 1 ? n0 : n1 + 2;
    const long long span_b = AXIS == 2 ? n1 : n2 + 2;
    const long long total = 2 * span_a * span_b;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int side = (int)(t & 1);
        const long long r = t >> 1;
        const int b = (int)(r % span_b), a = (int)(r / span_b);
        int ix, iy, iz, sx, sy, sz;
        if (AXIS == 2) { ix = sx = a; iy = sy = b; iz = side ? n2 : -1; sz = side ? n2 - 1 : 0; }
        else if (AXIS == 1) { ix = sx = a; iz = sz = b - 1; iy = side ? n1 : -1; sy = side ? n1 - 1 : 0; }
        else { iy = sy = a - 1; iz = sz = b - 1; ix = side ? n0 : -1; sx = side ? n0 - 1 : 0; }
        ENTRY val = e[texel_index(v, sx, sy, sz)];
        if (AXIS == 2 && !side && v.pair) {   // {T(0), T(0)}
            if constexpr (sizeof(ENTRY) == 32) { val.hi = val.lo; }
            else if constexpr (sizeof(ENTRY) == 16) { val.z = val.x; val.w = val.y; }
        }
        e[texel_index(v, ix, iy, iz)] = val;
    }
}

struct Entry32 { float4 lo, hi; };   // f32x4 z-pair

// scalar of entry idx (the first texel of a z-pair entry)
template <bool HALF>
__device__ __forceinline__ float load_scalar(const void *base, long long idx, int pair) {
    if constexpr (HALF) return __half2float(reinterpret_cast<const __half *>(base)[(4 * idx) << pair]);
    else return reinterpret_cast<const float *>(base)[(4 * idx) << pair];
}

// One warp per macrocell: min/max of the scalar over texels [8c, min(8c+8, n-1)]^3 -- the cell's own
// 8^3 voxels plus the +1 apron that the upper trilinear taps of its samples can reach.
template <bool HALF>
__global__ void __launch_bounds__(256)
cell_minmax_kernel(VolumeDesc v, float2 *__restrict__ out) {
    const long long n_cells = (long long)v.ncell[0] * v.ncell[1] * v.ncell[2];
    const int lane = threadIdx.x & 31;
    for (long long cell = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; cell < n_cells;
         cell += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int cz = (int)(cell % v.ncell[2]);
        const long long r = cell / v.ncell[2];
        const int cy = (int)(r % v.ncell[1]), cx = (int)(r / v.ncell[1]);
        const int x0 = kCell * cx, y0 = kCell * cy, z0 = kCell * cz;
        const int wx = min(x0 + kCell, v.n[0] - 1) - x0 + 1, wy = min(y0 + kCell, v.n[1] - 1) - y0 + 1,
                  wz = min(z0 + kCell, v.n[2] - 1) - z0 + 1;
        float lo = INFINITY, hi = -INFINITY;
        for (int t = lane; t < wx * wy * wz; t += 32) {
            const int dz = t % wz, dy = (t / wz) % wy, dx = t / (wz * wy);
            const long long at = texel_index(v, x0 + dx, y0 + dy, z0 + dz);
            const float s = load_scalar<HALF>(v.texels, at, v.pair);
            // NaN voxels: keep the cell active by poisoning the range
            if (s != s) { lo = -INFINITY; hi = INFINITY; }
            lo = fminf(lo, s); hi = fmaxf(hi, s);
        }
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) out[cell] = make_float2(lo, hi);
    }
}

// LUT taps a density can select (march.cu axis_taps): x = d*size - 0.5, taps clamp(floor(x)) and
// clamp(floor(x)+1).  Every sample of the cell has lo <= d <= hi up to a few ulps of lerp round-off,
// so [lo, hi] is widened by 1e-6 relative (~8 ulp) and x is evaluated with the very expression the
// march uses: binary32 multiply/subtract are monotonic, hence floor(x(d)) lies between the two ends.
// The cell is inactive iff every LUT alpha in that tap range is exactly zero.
__global__ void __launch_bounds__(256)
cell_classify_kernel(const float2 *__restrict__ mm, VolumeDesc v, const float4 *__restrict__ lut,
                     int lut_size, uint8_t *__restrict__ active, int *__restrict__ active_box) {
    extern __shared__ int s_nonzero_before[];  // [j] = number of nonzero alphas in lut[0..j)
    __shared__ int s_box[6];
    if (threadIdx.x == 0) {
        int run = 0;
        for (int j = 0; j < lut_size; ++j) {
            s_nonzero_before[j] = run;
            run += (lut[j].w != 0.0f) ? 1 : 0;  // NaN != 0 is true: stays active
        }
        s_nonzero_before[lut_size] = run;
        s_box[0] = s_box[1] = s_box[2] = 0x7fffffff;
        s_box[3] = s_box[4] = s_box[5] = -1;
    }
    __syncthreads();
    const size_t n_cells = (size_t)v.ncell[0] * v.ncell[1] * v.ncell[2];
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells;
         c += (size_t)gridDim.x * blockDim.x) {
        const float2 r = mm[c];
        const float slack = fmaxf(fabsf(r.x), fabsf(r.y)) * 1e-6f + 1e-30f;
        const float lo = r.x - slack, hi = r.y + slack;
        const float size = (float)lut_size;
        const float xl = floorf(lo * size - 0.5f), xh = floorf(hi * size - 0.5f) + 1.0f;
        const int jl = (int)fminf(fmaxf(xl, 0.0f), size - 1.0f);   // NaN/-inf -> 0, +inf -> size-1
        const int jh = (int)fminf(fmaxf(xh, 0.0f), size - 1.0f);
        const bool on = (s_nonzero_before[jh + 1] - s_nonzero_before[jl]) != 0;
        // seeds of the two distance maps, one per nibble: low = distance to the nearest active cell,
        // high = distance to the nearest inactive cell (0 = "is one", kCellDistCap = "far")
        active[c] = on ? (uint8_t)(kCellDistCap << 4) : (uint8_t)kCellDistCap;
        if (on) {   // bounding box of the active cells, block-local first
            const int cz = (int)(c % v.ncell[2]);
            const size_t rest = c / v.ncell[2];
            const int cy = (int)(rest % v.ncell[1]), cx = (int)(rest / v.ncell[1]);
            atomicMin(&s_box[0], cx); atomicMin(&s_box[1], cy); atomicMin(&s_box[2], cz);
            atomicMax(&s_box[3], cx); atomicMax(&s_box[4], cy); atomicMax(&s_box[5], cz);
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(active_box + threadIdx.x, s_box[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(active_box + threadIdx.x, s_box[threadIdx.x]);
}

// One relaxation sweep of both chessboard (Chebyshev) distance maps, in cells: each nibble becomes
// min(itself, 1 + min over the 26 neighbours).  Seeds are 0, everything else starts at kCellDistCap; after
// k <= kCellDistCap - 1 sweeps every value <= k is exact and the rest still hold kCellDistCap, which is then
// a valid LOWER bound of their true distance.  Cells outside the grid are neither active nor inactive.
__global__ void __launch_bounds__(256)
cell_dist_relax_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int n0, int n1, int n2) {
    const long long total = (long long)n0 * n1 * n2;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total;
         c += (long long)gridDim.x * blockDim.x) {
        const int cz = (int)(c % n2);
        const long long rest = c / n2;
        const int cy = (int)(rest % n1), cx = (int)(rest / n1);
        const int self = in[c];
        int lo = kCellDistCap, hi = kCellDistCap;
        for (int dx = -1; dx <= 1; ++dx) {
            const int x = cx + dx;
            if (x < 0 || x >= n0) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= n1) continue;
                for (int dz = -1; dz <= 1; ++dz) {
                    const int z = cz + dz;
                    if (z < 0 || z >= n2) continue;
                    const int nb = in[((long long)x * n1 + y) * n2 + z];
                    lo = min(lo, nb & 15);
                    hi = min(hi, nb >> 4);
                }
            }
        }
        lo = min(self & 15, min(lo + 1, kCellDistCap));
        hi = min(self >> 4, min(hi + 1, kCellDistCap));
        out[c] = (uint8_t)((hi << 4) | lo);
    }
}

// Nibble pair -> the byte the march reads (march.cu, phase 1):
//   inactive cell:  d      (1 .. cap)      every cell within chessboard radius d-1 is inactive
//   active cell:    128+w  (w = 0 .. cap-1) every cell within chessboard radius w is active
__global__ void __launch_bounds__(256)
cell_dist_finish_kernel(uint8_t *__restrict__ cells, long long total) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total;
         c += (long long)gridDim.x * blockDim.x) {
        const int v = cells[c], lo = v & 15, hi = v >> 4;
        cells[c] = lo == 0 ? (uint8_t)(128 + hi - 1) : (uint8_t)lo;
    }
}

__global__ void reset_active_box_kernel(int *active_box) {
    if (threadIdx.x < 3) active_box[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) active_box[threadIdx.x] = -1;
}

inline int grid_for(long long work, int block, int max_blocks = 148 * 16) {
    long long g = (work + block - 1) / block;
    return (int)(g < 1 ? 1 : (g > max_blocks ? max_blocks : g));
}

}  // namespace

cudaError_t launch_pack_texels(const float *scalar, const float *normals, const VolumeDesc &vol,
                               bool half_texels, cudaStream_t stream) {
    const long long total = (long long)vol.n[0] * vol.n[1] * vol.n[2];
    const int grid = grid_for(total, 256);
    if (half_texels) pack_texels_kernel<true><<<grid, 256, 0, stream>>>(scalar, normals, vol);
    else pack_texels_kernel<false><<<grid, 256, 0, stream>>>(scalar, normals, vol);
    return cudaGetLastError();
}

template <typename ENTRY>
static cudaError_t fill_apron_all(const VolumeDesc &vol, cudaStream_t stream) {
    const long long wz = 2LL * vol.n[0] * vol.n[1], wy = 2LL * vol.n[0] * (vol.n[2] + 2),
                    wx = 2LL * (vol.n[1] + 2) * (vol.n[2] + 2);
    fill_apron_kernel<ENTRY, 2><<<grid_for(wz, 256), 256, 0, stream>>>(vol);
    fill_apron_kernel<ENTRY, 1><<<grid_for(wy, 256), 256, 0, stream>>>(vol);
    fill_apron_kernel<ENTRY, 0><<<grid_for(wx, 256), 256, 0, stream>>>(vol);
    return cudaGetLastError();
}

cudaError_t launch_fill_apron(const VolumeDesc &vol, bool half_texels, cudaStream_t stream) {
    // entry sizes: f16x4 8 B (uint2), f16x4 pair / f32x4 16 B (uint4), f32x4 pair 32 B
    if (half_texels && !vol.pair) {
        // 8-byte entries: reuse the 16-byte kernel's logic through a dedicated instantiation
        return fill_apron_all<uint2>(vol, stream);
    }
    if (half_texels || !vol.pair) return fill_apron_all<uint4>(vol, stream);
    return fill_apron_all<Entry32>(vol, stream);
}

cudaError_t launch_cell_minmax(const VolumeDesc &vol, bool half_texels, float2 *cell_minmax,
                               cudaStream_t stream) {
    const long long n_cells = (long long)vol.ncell[0] * vol.ncell[1] * vol.ncell[2];
    const int grid = grid_for(n_cells * 32, 256);
    if (half_texels) cell_minmax_kernel<true><<<grid, 256, 0, stream>>>(vol, cell_minmax);
    else cell_minmax_kernel<false><<<grid, 256, 0, stream>>>(vol, cell_minmax);
    return cudaGetLastError();
}

cudaError_t launch_cell_classify(const float2 *cell_minmax, const VolumeDesc &vol, const float4 *lut,
                                 int lut_size, uint8_t *cell_dist, uint8_t *cell_scratch, int *active_box,
                                 cudaStream_t stream) {
    const long long n_cells = (long long)vol.ncell[0] * vol.ncell[1] * vol.ncell[2];
    const int grid = grid_for(n_cells, 256, 148 * 4);
    const size_t smem = (size_t)(lut_size + 1) * sizeof(int);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(cell_classify_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    reset_active_box_kernel<<<1, 32, 0, stream>>>(active_box);
    cell_classify_kernel<<<grid, 256, smem, stream>>>(cell_minmax, vol, lut, lut_size, cell_dist, active_box);
    // distance maps: kCellDistSweeps sweeps (even, so the result lands back in cell_dist), then the final byte
    const int grid2 = grid_for(n_cells, 256);
    static_assert(kCellDistSweeps % 2 == 0 && kCellDistSweeps < kCellDistCap && kCellDistCap <= 15, "nibble-sized distances");
    for (int it = 0; it < kCellDistSweeps; it += 2) {
        cell_dist_relax_kernel<<<grid2, 256, 0, stream>>>(cell_dist, cell_scratch, vol.ncell[0], vol.ncell[1], vol.ncell[2]);
        cell_dist_relax_kernel<<<grid2, 256, 0, stream>>>(cell_scratch, cell_dist, vol.ncell[0], vol.ncell[1], vol.ncell[2]);
    }
    cell_dist_finish_kernel<<<grid2, 256, 0, stream>>>(cell_dist, n_cells);
    return cudaGetLastError();
}

}  // namespace pyvr
