// common.cuh -- shared declarations of the sm_100a backend (device-side structs, error plumbing).
//
// Built with -fmad=false: the compiler never contracts a*b+c on its own, so the only fused
// operations are the explicit fmaf()/__fmaf_rn() calls in the kernels.  That makes the arithmetic
// of the STRICT path a statement-by-statement twin of oracle/pyvr_oracle.c (DESIGN.md,
// "Arithmetic contract").
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pyvr_cuda.h"

namespace pyvr {

// Address map of the packed texel array.  Texels live in 128-byte lines of SLOTS = 128/texel_bytes
// consecutive-z texels (8 for f32x4, 16 for f16x4):
//   line(ix,iy,iz)  = off_x(ix) + off_y(iy) + off_z(iz),   off_a(i) = (i >> shift)*outer + (i & mask)*inner
//   slot(ix,iy,iz)  = (iz + swz[0]*ix + swz[1]*iy) & (SLOTS-1)
//   texel index     = line * SLOTS + slot
// The separable line map covers plain rows (x,y linear) and 8x8 bricks of lines.  The slot rotation
// ("swizzle") permutes texels inside their line so that the texels one warp-wide load touches -- a
// small planar patch, typically many (x,y) rows at the same z -- spread over the L1 data banks
// instead of all landing on the banks of one slot ("L1 bank swizzle").
struct AxisMap {
    int shift;
    int mask;
    long long outer;
    long long inner;
};

struct VolumeDesc {
    const void *texels;   // float4 {s,nx,ny,nz} or 4 x half
    int n[3];             // texel counts along world x, y, z.  NB world z is the memory-fastest axis
                          // of the array the reference uploads: (nz, ny, nx) = numpy shape (0, 1, 2).
    AxisMap map[3];       // line offsets (see above)
    int slot_shift;       // log2(SLOTS): 3 for f32x4, 4 for f16x4
    int swz[2];           // slot rotation multipliers for ix, iy (0, 0 = no swizzle)
    float bmin[3], bmax[3];
    // fast path: voxel coordinate = world * vscale + voff  (= tc * n - 0.5)
    float vscale[3], voff[3];
    // empty-space skipping: one byte per 8^3 macrocell, 1 = some sample in it may have alpha != 0
    const uint8_t *cell_active;
    int ncell[3];
};

__host__ __device__ __forceinline__ long long axis_offset(const AxisMap &m, int i) {
    return (long long)(i >> m.shift) * m.outer + (long long)(i & m.mask) * m.inner;
}

__host__ __device__ __forceinline__ long long texel_index(const VolumeDesc &v, int ix, int iy, int iz) {
    const long long line = axis_offset(v.map[0], ix) + axis_offset(v.map[1], iy) + axis_offset(v.map[2], iz);
    const int slot = (iz + v.swz[0] * ix + v.swz[1] * iy) & ((1 << v.slot_shift) - 1);
    return (line << v.slot_shift) + slot;
}

struct MarchArgs {
    VolumeDesc vol;
    const pyvr_view *views;        // device array, indexed by blockIdx.z
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
    unsigned long long *counters;  // [samples, fetched, rays_hit, rays_terminated]
};

enum { CNT_SAMPLES = 0, CNT_FETCHED = 1, CNT_HIT = 2, CNT_TERM = 3, CNT_N = 4 };

// Launchers (defined next to their kernels).
cudaError_t launch_march(const MarchArgs &args, int n_views, bool half_texels, bool wide_index,
                         cudaStream_t stream);
cudaError_t launch_pack_texels(const float *scalar, const float *normals, const VolumeDesc &vol,
                               bool half_texels, cudaStream_t stream);
cudaError_t launch_cell_minmax(const VolumeDesc &vol, bool half_texels, float2 *cell_minmax,
                               cudaStream_t stream);
cudaError_t launch_cell_classify(const float2 *cell_minmax, size_t n_cells, const float4 *lut,
                                 int lut_size, uint8_t *cell_active, cudaStream_t stream);
cudaError_t launch_normals(const float *in, float *out, int n0, int n1, int n2, cudaStream_t stream);

}  // namespace pyvr
