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

// Packed texel array.  Texels {s,nx,ny,nz} (float4, or 4 x binary16) are stored in plain [x][y][z] order with
//   * a one-texel APRON on every side that replicates the edge texel (CLAMP_TO_EDGE made explicit): a sample's
//     lower tap floor(x) ranges over [-1, n-1] and the upper tap is always lower + 1, so the fast march needs no
//     index clamps and the four (x, y) corner rows of a sample sit at fixed distances from the first one;
//   * PADDED PITCHES: one entry = one texel (or one z-pair, below); a row of entries along z is pitch_y entries
//     long, an x-plane pitch_x entries, with pitch_y = 2 and pitch_x = 3 modulo S = 128 / entry_bytes.  One
//     warp-wide corner load touches a small planar patch of texels, typically many (x, y) rows at the same z;
//     with aligned rows they would all sit at the same offset of different 128-byte lines, i.e. on the same L1
//     data banks, and the load would serialise (measured in round 1: 18.6 data-stage wavefronts per LDG.128, L1
//     data pipe 99 % busy; 7.9 with the rotation).  The padding rotates the bank of texel (ix, iy, iz) by
//     (3*ix + 2*iy) entries -- what round 1 did with a slot swizzle -- without a single instruction in the march;
//   * optional Z-PAIR entries (pair = 1): entry(ix, iy, iz) = {texel(iz), texel(iz + 1)} (2x memory): one 256-bit
//     (f32x4) / 128-bit (f16x4) load fetches both z-taps of a corner row.
//   entry index (ix, iy, iz) = (ix + 1) * pitch_x + (iy + 1) * pitch_y + (iz + 1),   ix in [-1, n[0]] etc.
//   * BRICK8 (f16x4 volumes whose rays will be sparse -- C4 / C5 -- or option "brick8"; never with z-pairs): the
//     array, apron included, is cut into 2x2x2-texel bricks, one brick = 8 consecutive texels = 64 bytes (f16x4: one
//     DRAM access) or 128 bytes (f32x4), so that a plane of samples cuts the fewest DRAM accesses whatever the view
//     direction: 2.4x less DRAM traffic than z-paired rows on C4, which the march turns into time by keeping several
//     samples in flight per ray (abi.cu, choose_layout; march.cu, LAYOUT 4).  With X = ix + 1 etc.:
//   entry index = (((X >> 1) * pitch_x + (Y >> 1) * pitch_y + (Z >> 1)) << 3) | (X & 1) << 2 | (Y & 1) << 1 | (Z & 1)
//     with pitch_y = bricks per z-row, pitch_x = bricks per x-plane.
struct VolumeDesc {
    const void *texels;   // allocation start = entry (-1, -1, -1)
    int n[3];             // texel counts of the STORED block along world x, y, z.  NB world z is the
                          // memory-fastest axis of the array the reference uploads: (nz, ny, nx) = numpy
                          // shape (0, 1, 2).
    // Sort-last bricks (pyvr_cuda_upload_brick): the stored array is the sub-block [org, org + n) of a
    // volume of gn texels; this brick owns the samples whose voxel coordinate x satisfies
    // own_lo <= x < own_hi on every axis (own_lo = -inf / own_hi = +inf on the volume's outer faces).
    // For a whole volume gn = n, org = 0, bricked = 0.
    int gn[3], org[3];
    float own_lo[3], own_hi[3];
    int bricked;
    int pair;             // 1: z-pair entries
    int brick8;           // 1: 2x2x2-texel bricks (pair = 0); pitch_y / pitch_x then count bricks
    int pitch_y;          // entries from (ix, iy, iz) to (ix, iy + 1, iz)
    long long pitch_x;    // entries from (ix, iy, iz) to (ix + 1, iy, iz)
    float bmin[3], bmax[3];
    // fast path: voxel coordinate = world * vscale + voff  (= tc * n - 0.5)
    float vscale[3], voff[3];
    // empty-space skipping: one byte per macrocell (kCell^3 voxels, 8^3 by default) (a cell is "active" if some sample in it may have
    // alpha != 0): b < 128 = inactive and every cell within chessboard radius b-1 is inactive; b >= 128 =
    // active and every cell within radius b-128 is active.  Plus the bounding box of the active cells
    // {lo_x, lo_y, lo_z, hi_x, hi_y, hi_z} in cell units (hi inclusive; lo > hi when no cell is active).
    // Both are written by launch_cell_classify.
    const uint8_t *cell_dist;
    const int *active_box;
    int ncell[3];
};

// Entry index of texel (ix, iy, iz) of the stored block; -1 and n address the apron.
__host__ __device__ __forceinline__ long long texel_index(const VolumeDesc &v, int ix, int iy, int iz) {
    if (v.brick8) {
        const int X = ix + 1, Y = iy + 1, Z = iz + 1;
        return ((((long long)(X >> 1) * v.pitch_x + (long long)(Y >> 1) * v.pitch_y + (Z >> 1)) << 3) |
                (long long)(((X & 1) << 2) | ((Y & 1) << 1) | (Z & 1)));
    }
    return (long long)(ix + 1) * v.pitch_x + (long long)(iy + 1) * v.pitch_y + (iz + 1);
}
// entries of the whole allocation
__host__ __device__ __forceinline__ long long entry_count(const VolumeDesc &v) {
    if (v.brick8) return (long long)((v.n[0] + 3) >> 1) * v.pitch_x * 8;
    return (long long)(v.n[0] + 2) * v.pitch_x;
}

struct MarchArgs {
    VolumeDesc vol;
    // fast path addressing: a sample with lower taps (ix, iy, iz) -- indices of the stored block, -1 = apron -- reads
    // its four (x, y) corner rows at tap_base + entry_bytes * (ix*pitch_x + iy*pitch_y + iz) + {0, stride_y,
    // stride_x, stride_x + stride_y}
    const char *tap_base;          // address of entry (0, 0, 0); brick8: the allocation start
    long long stride_y, stride_x;  // pitch_y, pitch_x in bytes; brick8: bytes per z-row / x-plane of bricks
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
    // Image-space sharding (multi-GPU tiles): groups of 2^shard_shift x 2^shard_shift CTA tiles (16x8 pixels each)
    // are dealt over the ranks, owner(gx, gy) = (gx + gy) mod shard_count; this launch marches the groups of
    // shard_rank only and leaves every other pixel untouched (the C ABI clears the frame first unless the caller
    // asked for in-place sharding into a frame that all ranks write).  shard_count <= 1: everything.
    int shard_rank, shard_count, shard_shift;
    int n_views;                   // views of this launch (filled by the launcher)
    int two_samples;               // f16x4 march: several samples of a run in flight per lane (sparse rays; LAYOUT 3 / 4)
    int first_row;                 // tile row dispatched first (the rows follow from it outwards); -1 = natural order
    unsigned *tile_counter;        // PYVR_PERSISTENT builds: ticket counter of the tile queue (zeroed per launch)
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
// last binary-swap round fused with finalize: out8 = rgba8(front over back); accum_out (may be null) keeps the floats
cudaError_t launch_composite_finalize(const float4 *front, const float4 *back, float4 *accum_out, uchar4 *out8,
                                      size_t n_pixels, float term_alpha, unsigned flags, cudaStream_t stream);
// stream-ordered flags in (peer) device memory: signal = system-scope release store, wait = spin until *flag >= value
cudaError_t launch_flag_signal(unsigned *flag, unsigned value, cudaStream_t stream);
cudaError_t launch_flag_wait(const unsigned *flags, int n_flags, unsigned value, cudaStream_t stream);
cudaError_t launch_flag_signal_many(unsigned *const *flags, int n, unsigned value, cudaStream_t stream);   // one launch per 16 flags
cudaError_t launch_pack_texels(const float *scalar, const float *normals, const VolumeDesc &vol,
                               bool half_texels, cudaStream_t stream);
// replicate the edge texels of the stored block into the one-texel apron (after every pack / generate)
cudaError_t launch_fill_apron(const VolumeDesc &vol, bool half_texels, cudaStream_t stream);
cudaError_t launch_cell_minmax(const VolumeDesc &vol, bool half_texels, float2 *cell_minmax,
                               cudaStream_t stream);
// Macrocell edge of the empty-space map: 2^PYVR_CELL_SHIFT voxels.  Smaller cells skip ~9 % more samples of C3 but
// the map walk visits more of them: 4^3 against 8^3 was a tie on views 0..15 and is +2 % over the whole turntable
// (profiles/r02_turntable_ab.txt: 481 vs 473 with z-pairs, 495 vs 486 without).
#ifndef PYVR_CELL_SHIFT
#define PYVR_CELL_SHIFT 2
#endif
constexpr int kCellShift = PYVR_CELL_SHIFT, kCell = 1 << kCellShift;
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
// relaxed: quotients as g * (1/norm) (within 2 ulp; TMA kernel only) instead of correctly rounded divisions
cudaError_t launch_normals(const float *in, float *out, int n0, int n1, int n2, bool relaxed, cudaStream_t stream);
// bandwidth.cu: measured cache bandwidths (roofline denominators); level 1 = L1 load-return, 2 = L2 -> SM
cudaError_t measure_cache_bandwidth(int level, double *gbs);

}  // namespace pyvr
