// common.cuh -- shared declarations of the sm_100a backend (device-side structs, launch prototypes).
//
// Built with -fmad=false: the compiler never contracts a*b+c on its own, so the only fused
// operations are the explicit fmaf() calls in the kernels.  That makes the arithmetic of the STRICT
// path a statement-by-statement twin of oracle/pyvr_oracle.c ("Arithmetic contract").
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pyvr_cuda.h"

namespace pyvr {

// Packed texel array.  Texels {s,nx,ny,nz} live in 128-byte lines of SLOTS = 128/texel_bytes
// consecutive-z texels (8 for f32x4, 16 for f16x4); lines are ordered x-major, then y, then z:
//   line(ix,iy,iz) = (ix*n[1] + iy) * row_lines + (iz >> slot_shift)
//   slot(ix,iy,iz) = (swz_z*iz + swz_x*ix + swz_y*iy) & (SLOTS-1)
//   texel index    = line * SLOTS + slot
// The slot rotation ("swizzle") permutes texels inside their line.  One warp-wide corner load touches
// a small planar patch of texels, typically many (x,y) rows at the same z; without the rotation they
// all sit in the same slot of different lines, i.e. on the same L1 data banks, and the load
// serialises into one data-stage wavefront per texel (measured: 18.6 wavefronts per LDG.128,
// l1tex data pipe at 99 %).  With it they spread over the banks ("L1 bank swizzle").
struct VolumeDesc {
    const void *texels;   // float4 {s,nx,ny,nz} or 4 x half
    int n[3];             // texel counts of the STORED array along world x, y, z.  NB world z is the
                          // memory-fastest axis of the array the reference uploads: (nz, ny, nx) = numpy
                          // shape (0, 1, 2).
    // Sort-last bricks (pyvr_cuda_upload_brick): the stored array is the sub-block [org, org + n) of a
    // volume of gn texels; this brick owns the samples whose voxel coordinate x satisfies
    // own_lo <= x < own_hi on every axis (own_lo = -inf / own_hi = +inf on the volume's outer faces).
    // For a whole volume gn = n, org = 0, bricked = 0.
    int gn[3], org[3];
    float own_lo[3], own_hi[3];
    int bricked;
    int slot_shift;       // log2(entries per 128-byte line): 3 for f32x4, 4 for f16x4, one less with pairs
    int pair;             // 1: every entry holds the z-pair {texel(iz), texel(min(iz+1, n-1))} (2x memory): one
                          // 256-bit (f32x4) / 128-bit (f16x4) load fetches both z-taps of a trilinear corner row
    int row_lines;        // lines per z-row = ceil(n[2] / SLOTS)
    int swz_x, swz_y, swz_z;   // slot = (swz_z*iz + swz_x*ix + swz_y*iy) mod SLOTS; swz_z odd; (0, 0, 1) = no swizzle
    float bmin[3], bmax[3];
    // fast path: voxel coordinate = world * vscale + voff  (= tc * n - 0.5)
    float vscale[3], voff[3];
    // empty-space skipping: one byte per 8^3 macrocell (a cell is "active" if some sample in it may have
    // alpha != 0): b < 128 = inactive and every cell within chessboard radius b-1 is inactive; b >= 128 =
    // active and every cell within radius b-128 is active.  Plus the bounding box of the active cells
    // {lo_x, lo_y, lo_z, hi_x, hi_y, hi_z} in cell units (hi inclusive; lo > hi when no cell is active).
    // Both are written by launch_cell_classify.
    const uint8_t *cell_dist;
    const int *active_box;
    int ncell[3];
};

__host__ __device__ __forceinline__ long long texel_index(const VolumeDesc &v, int ix, int iy, int iz) {
    const long long line = ((long long)ix * v.n[1] + iy) * v.row_lines + (iz >> v.slot_shift);
    const int slot = (v.swz_z * iz + v.swz_x * ix + v.swz_y * iy) & ((1 << v.slot_shift) - 1);
    return (line << v.slot_shift) + slot;
}

struct MarchArgs {
    VolumeDesc vol;
    const pyvr_view *views;        // device array, indexed by blockIdx.z
    cudaTextureObject_t tex;       // PYVR_FLAG_HWTEX: the same texels as a 3-D array (width = z), linear filter, clamp
    const float4 *lut;             // device, lut_size entries
    int lut_size;
    int width, height;
    float step, ref_step;
    float exp2_scale;              // -(step/ref) * log2(e), fast path
    int max_steps;
    float ambient, diffuse;
    float ldir[3];                 // normalize(light_target - light_position), binary32
    float term_alpha;
    unsigned flags;
    uchar4 *out8;                  // n_views * height * width, row 0 = bottom; may be null
    float4 *out_acc;               // same shape, pre-blend fragment colour; may be null
    const float4 *in_acc;          // relay (sort-last, exact): fragment colour accumulated by the bricks in
                                   // front of this one; the march continues from it.  May be null / == out_acc.
    unsigned long long *counters;  // [samples, fetched, rays_hit, rays_terminated]
    // Image-space sharding (multi-GPU tiles): 64x64-pixel tile groups are dealt round-robin; this launch
    // marches the groups with (group index) % shard_count == shard_rank and writes zeros elsewhere, so the
    // per-rank frames add up to the full frame.  shard_count <= 1: everything.
    int shard_rank, shard_count;
};

// Fragment colour -> what fbo.read returns: clamp to [0,1], blend SRC_ALPHA / ONE_MINUS_SRC_ALPHA onto the
// (0,0,0,0) clear (manager.py:217-220; skipped with PYVR_FLAG_NO_BLEND), clamp, round to nearest RGBA8.
__device__ __forceinline__ uchar4 fragment_to_rgba8(float cr, float cg, float cb, float ca, unsigned flags) {
    auto clamp01 = [](float v) { return fminf(fmaxf(v, 0.0f), 1.0f); };
    const float al = clamp01(ca);
    float r = clamp01(cr), g = clamp01(cg), b = clamp01(cb), o = al;
    if (!(flags & PYVR_FLAG_NO_BLEND)) { r *= al; g *= al; b *= al; o = al * al; }
    uchar4 q;
    q.x = (unsigned char)__float2uint_rn(clamp01(r) * 255.0f);
    q.y = (unsigned char)__float2uint_rn(clamp01(g) * 255.0f);
    q.z = (unsigned char)__float2uint_rn(clamp01(b) * 255.0f);
    q.w = (unsigned char)__float2uint_rn(clamp01(o) * 255.0f);
    return q;
}

enum { CNT_SAMPLES = 0, CNT_FETCHED = 1, CNT_HIT = 2, CNT_TERM = 3, CNT_N = 4 };

// Launchers (defined next to their kernels).
cudaError_t launch_march(const MarchArgs &args, int n_views, bool half_texels, bool wide_index,
                         cudaStream_t stream);
// bytes of one entry of the packed array
__host__ __device__ __forceinline__ int entry_bytes(bool half_texels, int pair) { return (half_texels ? 8 : 16) << pair; }
// composite.cu: sort-last compositing of pre-blend fragment colours (premultiplied rgb, alpha)
cudaError_t launch_composite_over(const float4 *front, const float4 *back, float4 *out, size_t n_pixels,
                                  float term_alpha, cudaStream_t stream);
cudaError_t launch_finalize_rgba8(const float4 *accum, uchar4 *out, size_t n_pixels, unsigned flags,
                                  cudaStream_t stream);
cudaError_t launch_pack_texels(const float *scalar, const float *normals, const VolumeDesc &vol,
                               bool half_texels, cudaStream_t stream);
cudaError_t launch_cell_minmax(const VolumeDesc &vol, bool half_texels, float2 *cell_minmax,
                               cudaStream_t stream);
constexpr int kCellDistCap = 15;      // distances saturate here (a nibble); a saturated value is a lower bound
constexpr int kCellDistSweeps = 14;   // relaxation sweeps: every distance <= 14 is exact
cudaError_t launch_cell_classify(const float2 *cell_minmax, const VolumeDesc &vol, const float4 *lut,
                                 int lut_size, uint8_t *cell_dist, uint8_t *cell_scratch, int *active_box,
                                 cudaStream_t stream);
// synth.cu: analytic volume -> packed texels (scalar + normals), slab by slab through `scratch`
cudaError_t launch_synth_volume(const VolumeDesc &vol, bool half_texels, int shape, int size, float *scratch,
                                size_t scratch_floats, cudaStream_t stream);
// raw texels of x-planes [x0, x0 + nx) in plain [x][y][z] order (staging for the 3-D CUDA array)
cudaError_t launch_linearize_texels(const VolumeDesc &vol, bool half_texels, int x0, int nx, void *dst, cudaStream_t stream);
cudaError_t launch_unpack_texels(const VolumeDesc &vol, bool half_texels, float *scalar, float *normals,
                                 cudaStream_t stream);
cudaError_t launch_normals(const float *in, float *out, int n0, int n1, int n2, cudaStream_t stream);

}  // namespace pyvr
