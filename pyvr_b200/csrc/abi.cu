// abi.cu -- the C ABI of include/pyvr_cuda.h: context, device resources, host-side parameter
// derivation and launch sequencing.  Stands in for ModernGLManager
// (pyvr/moderngl_renderer/manager.py) of the reference; the kernels live in march.cu,
// volume_pack.cu and normals.cu.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace pyvr;

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? PYVR_ERR_NOMEM : PYVR_ERR_CUDA,        \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr double kPairBudget = 0.8;   // auto z-pair layout while the doubled array fits in this share of free HBM
constexpr int kRing = 3;   // device slots used to overlap march and device->host copies
constexpr int kSlotViews = 4;   // views marched per launch on the host-output path (one slot = kSlotViews frames)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct pyvr_ctx {
    int device = 0;
    int width = 0, height = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;

    // volume
    void *texels = nullptr;
    size_t texel_bytes = 0;
    bool half_texels = false;
    bool have_volume = false;
    VolumeDesc vol{};
    float2 *cell_minmax = nullptr;
    uint8_t *cell_dist = nullptr, *cell_scratch = nullptr;   // distance map + ping-pong buffer
    int *active_box = nullptr;     // 6 ints, see VolumeDesc
    size_t n_cells = 0;
    bool swizzle = true;  // L1 bank swizzle of the texel layout (common.cuh); off only for A/B profiling
    int shard_rank = 0, shard_count = 1;   // image-space tile sharding (pyvr_cuda_set_pixel_shard)
    int shard_shift = 1;                   // tile groups of 2^shift x 2^shift CTA tiles (32x16 pixels by default)
    bool shard_in_place = false;           // true: foreign pixels are left untouched (all ranks write one shared frame)
    int pair_option = -1;   // z-pair entries: -1 auto (when the doubled array stays under kPairBudget), 0 off, 1 on
    bool use_pair = false;  // decided per upload
    bool async_device = false;      // option "async_device_output": device-output renders return without a host sync
    bool stats_pending = false;     // counters of the last render are on their way to h_counters (stats_ready)
    size_t pending_pairs = 0;
    int pending_views = 0;
    cudaEvent_t stats_ready = nullptr;
    const pyvr_view *launch_view = nullptr;   // host copy of the first view of the launch being prepared (row order)
    int two_option = -1;    // f16x4 z-pair march two samples at a time: -1 auto (sparse rays expected), 0 off, 1 on
    int brick8_option = -1; // 2x2x2-texel bricks (common.cuh): -1 auto (f16x4 and sparse rays expected), 0 off, 1 on
    bool use_brick8 = false;
    cudaArray_t tex_array = nullptr;        // PYVR_FLAG_HWTEX: built on first use from the packed texels
    cudaTextureObject_t tex_obj = 0;

    // transfer function
    float4 *lut = nullptr;
    int lut_size = 0;

    // camera / params
    pyvr_view view{};
    pyvr_params params{};
    bool have_view = false;

    // frame resources
    pyvr_view *d_views = nullptr;
    int views_cap = 0;
    uchar4 *frames = nullptr;      // kRing slots of kSlotViews frames
    float4 *accum = nullptr;       // one frame, allocated on demand
    unsigned long long *d_counters = nullptr;
    unsigned long long *h_counters = nullptr;  // pinned
    std::vector<cudaEvent_t> ev;   // pairs (start, stop) around march launches
    cudaEvent_t slot_rendered[kRing] = {}, slot_copied[kRing] = {};

    pyvr_stats stats{};
};

namespace {

size_t frame_pixels(const pyvr_ctx *c) { return (size_t)c->width * (size_t)c->height; }

// 4x4 inverse in binary64 (Gauss-Jordan, partial pivoting).  m, out: row-major [r][c].
bool invert4(const double m[4][4], double out[4][4]) {
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            a[r][c] = m[r][c];
            a[r][4 + c] = r == c ? 1.0 : 0.0;
        }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        if (a[piv][col] == 0.0) return false;
        if (piv != col)
            for (int c = 0; c < 8; ++c) { double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
        const double inv = 1.0 / a[col][col];
        for (int c = 0; c < 8; ++c) a[col][c] *= inv;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = a[r][col];
            if (f != 0.0)
                for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) out[r][c] = a[r][4 + c];
    return true;
}

// GLSL inverse(mat4) evaluated in binary32 by cofactor expansion (column-major m[c*4+r]); the STRICT
// march applies the result per pixel.  Same expression order as the oracle's restatement, so both
// sides round identically ("Arithmetic contract").
void inverse4_f32(const float *m, float *o) {
    const float a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3], a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    const float a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11], a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    const float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10;
    const float b03 = a01 * a12 - a02 * a11, b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
    const float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30, b08 = a20 * a33 - a23 * a30;
    const float b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    const float det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
    const float id = 1.0f / det;
    o[0] = (a11 * b11 - a12 * b10 + a13 * b09) * id;
    o[1] = (a02 * b10 - a01 * b11 - a03 * b09) * id;
    o[2] = (a31 * b05 - a32 * b04 + a33 * b03) * id;
    o[3] = (a22 * b04 - a21 * b05 - a23 * b03) * id;
    o[4] = (a12 * b08 - a10 * b11 - a13 * b07) * id;
    o[5] = (a00 * b11 - a02 * b08 + a03 * b07) * id;
    o[6] = (a32 * b02 - a30 * b05 - a33 * b01) * id;
    o[7] = (a20 * b05 - a22 * b02 + a23 * b01) * id;
    o[8] = (a10 * b10 - a11 * b08 + a13 * b06) * id;
    o[9] = (a01 * b08 - a00 * b10 - a03 * b06) * id;
    o[10] = (a30 * b04 - a31 * b02 + a33 * b00) * id;
    o[11] = (a21 * b02 - a20 * b04 - a23 * b00) * id;
    o[12] = (a11 * b07 - a10 * b09 - a12 * b06) * id;
    o[13] = (a00 * b09 - a01 * b07 + a02 * b06) * id;
    o[14] = (a31 * b01 - a30 * b03 - a32 * b00) * id;
    o[15] = (a20 * b03 - a21 * b01 + a22 * b00) * id;
}

// Layout (common.cuh) for this upload.  Dense rays (C3: 512^3 at 1080p, 0.6 voxels between neighbouring rays, L1-bound):
// rows, z-paired if memory allows.  Sparse rays (C4 / C5: 2048^3 .. 4096^3 f16x4 at 2160p, 1.2-1.6 voxels between
// rays, every load a DRAM round trip): 2x2x2-texel bricks, marched with several samples in flight per ray.  Sparse =
// more than about 0.85 voxels between neighbouring rays of a fitted view, i.e. (45 degree field of view, camera three
// half-extents away) a volume edge above 0.7 x the frame height.  Evidence, C4 on one B200
// (profiles/r02_brick8_ab.txt, r02_multi_sample_ab.txt, r02_c4_ncu_summary.txt):
//   rows + z-pairs, one sample in flight     8.0-8.6 ms   45.3 GB of DRAM reads per frame, DRAM 65 % busy
//   2x2x2 bricks,   one sample in flight     9.3-10.1 ms  18.9 GB, DRAM 25 % busy: latency-bound, prefetch no help
//   rows + z-pairs, 2-4 samples in flight    6.1-6.8 ms   DRAM-bound again (45 GB at ~7 TB/s)
//   2x2x2 bricks,   4 samples in flight      3.7 ms       <- the default for this regime; half the memory of z-pairs
// The brick layout alone was a negative result; what it needed was enough loads in flight per ray to turn its 2.4x
// lower traffic into time.  f32x4 texels keep the row layouts (no multi-sample path: 32 registers per sample).
bool sparse_rays_expected(const pyvr_ctx *c, const int g[3]) {
    const int edge = g[0] > g[1] ? (g[0] > g[2] ? g[0] : g[2]) : (g[1] > g[2] ? g[1] : g[2]);
    return (double)edge > 0.7 * (double)c->height;
}

void choose_layout(pyvr_ctx *c, const int local[3], const int global[3]) {
    // sort-last bricks (global != NULL) are marched one sample at a time (make_args), and then a sample is cheaper as
    // four 16-byte requests than as eight 8-byte ones: C5 on 8 GPUs 0.96 ms per frame with z-paired rows, 1.03 with bricks
    c->use_brick8 = c->brick8_option < 0 ? (c->half_texels && !global && sparse_rays_expected(c, local))
                                         : c->brick8_option > 0;
    const size_t doubled = (size_t)local[0] * local[1] * local[2] * (c->half_texels ? 8 : 16) * 2;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;   // called after the old volume was freed
    // z-pairs by default for f16x4 only.  For f32x4 they won every A/B on views 0..15 (profiles/r02_layout_ab.txt: 428
    // vs 396 Gsamples/s with 4x1-pixel passes) but over the whole turntable, with the shipped 2x2-pixel passes, eight
    // LDG.128 over eight 16-byte slots beat four LDG.256 over four 32-byte slots (profiles/r02_turntable_ab.txt: 486 vs
    // 473) -- and need half the memory.
    c->use_pair = !c->use_brick8 &&
                  (c->pair_option < 0 ? (c->half_texels && (double)doubled <= kPairBudget * (double)free_b) : c->pair_option != 0);
}

// local[3] = stored texel counts along world x, y, z; global/org/own_* = NULL for a whole volume.
void fill_volume_desc(pyvr_ctx *c, const int local[3], const int global[3], const int org[3],
                      const int own_lo[3], const int own_hi[3], const float bmin[3], const float bmax[3]) {
    VolumeDesc &v = c->vol;
    choose_layout(c, local, global);
    v.bricked = global != nullptr;
    for (int a = 0; a < 3; ++a) {
        v.n[a] = local[a];
        v.gn[a] = global ? global[a] : local[a];
        v.org[a] = org ? org[a] : 0;
        // ownership in voxel coordinates; the volume's outer faces are open-ended
        v.own_lo[a] = (own_lo && own_lo[a] > 0) ? (float)own_lo[a] : -3.0e38f;
        v.own_hi[a] = (own_hi && own_hi[a] < v.gn[a]) ? (float)own_hi[a] : 3.0e38f;
        v.bmin[a] = bmin[a]; v.bmax[a] = bmax[a];
        const double ext = (double)bmax[a] - (double)bmin[a];
        v.vscale[a] = (float)((double)v.gn[a] / ext);
        v.voff[a] = (float)(-(double)bmin[a] * (double)v.gn[a] / ext - 0.5);
        v.ncell[a] = (v.n[a] + kCell - 1) / kCell;
    }
    // padded pitches (common.cuh): rows of n[2] + 2 entries, planes of n[1] + 2 rows, rounded up to the residue
    // that rotates the L1 slot of texel (ix, iy, iz) by rx*ix + ry*iy entries.  Measured on C3 with z-pair entries (4
    // per line) over the whole turntable (profiles/r02_turntable_ab.txt): no rotation 356 Gsamples/s; with 4x1-pixel
    // passes every odd/odd pair 459; with the shipped 2x2-pixel passes (march.cu, PYVR_LANE_ARR) (3,2) 473, (2,1) 470,
    // (3,1) 455.  What matters is that the texels the four lanes of a pass touch land in different slots.
    v.pair = c->use_pair ? 1 : 0;
    v.brick8 = c->use_brick8 ? 1 : 0;
    if (v.brick8) {   // pitches count bricks; the apron is part of the bricked array
        v.pitch_y = (v.n[2] + 3) >> 1;
        v.pitch_x = (long long)((v.n[1] + 3) >> 1) * v.pitch_y;
        return;
    }
    const int slots = 128 / entry_bytes(c->half_texels, v.pair);
    int rx = 3, ry = 2;
    if (!c->swizzle) rx = ry = 0;
    else {
        const char *env = getenv("PYVR_CUDA_SWZ");   // "x,y" override for experiments
        int ex, ey;
        if (env && sscanf(env, "%d,%d", &ex, &ey) == 2) { rx = ex; ry = ey; }
    }
    auto pad_to = [slots](long long len, int residue) {
        const int r = ((residue % slots) + slots) % slots;
        return len + (((r - len) % slots) + slots) % slots;
    };
    v.pitch_y = (int)pad_to(v.n[2] + 2, ry);
    v.pitch_x = pad_to((long long)(v.n[1] + 2) * v.pitch_y, rx);
}

size_t texel_count(const pyvr_ctx *c) { return (size_t)entry_count(c->vol); }

int classify_cells(pyvr_ctx *c) {
    if (!c->have_volume || c->lut_size <= 0) return PYVR_OK;
    CU(launch_cell_classify(c->cell_minmax, c->vol, c->lut, c->lut_size, c->cell_dist, c->cell_scratch, c->active_box, c->stream));
    return PYVR_OK;
}

void free_texture(pyvr_ctx *c) {
    if (c->tex_obj) cudaDestroyTextureObject(c->tex_obj);
    if (c->tex_array) cudaFreeArray(c->tex_array);
    c->tex_obj = 0;
    c->tex_array = nullptr;
}

// 3-D CUDA array + texture object over the stored block: width = z (memory-fastest), height = y, depth = x;
// LINEAR filter, CLAMP addressing, unnormalised coordinates -- the sampler state of manager.py:98-101.
// Filled slab by slab through a <= 256 MiB staging buffer.
int ensure_texture(pyvr_ctx *c) {
    if (c->tex_obj || !c->have_volume) return PYVR_OK;
    const VolumeDesc &v = c->vol;
    const size_t tb = c->half_texels ? 8 : 16;
    cudaChannelFormatDesc desc = c->half_texels ? cudaCreateChannelDescHalf4() : cudaCreateChannelDesc<float4>();
    CU(cudaMalloc3DArray(&c->tex_array, &desc, make_cudaExtent(v.n[2], v.n[1], v.n[0])));
    const size_t plane = (size_t)v.n[1] * v.n[2] * tb;
    size_t planes = ((size_t)256 << 20) / plane;
    if (planes < 1) planes = 1;
    if (planes > (size_t)v.n[0]) planes = v.n[0];
    void *stage = nullptr;
    cudaError_t e = cudaMalloc(&stage, planes * plane);
    for (int x0 = 0; e == cudaSuccess && x0 < v.n[0]; x0 += (int)planes) {
        const int nx = v.n[0] - x0 < (int)planes ? v.n[0] - x0 : (int)planes;
        e = launch_linearize_texels(v, c->half_texels, x0, nx, stage, c->stream);
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(stage, (size_t)v.n[2] * tb, v.n[2], v.n[1]);
        p.dstArray = c->tex_array;
        p.dstPos = make_cudaPos(0, 0, x0);
        p.extent = make_cudaExtent(v.n[2], v.n[1], nx);
        p.kind = cudaMemcpyDeviceToDevice;
        if (e == cudaSuccess) e = cudaMemcpy3DAsync(&p, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(stage);
    if (e == cudaSuccess) {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = c->tex_array;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        e = cudaCreateTextureObject(&c->tex_obj, &rd, &td, nullptr);
    }
    if (e != cudaSuccess) { free_texture(c); CU(e); }
    return PYVR_OK;
}

void free_volume(pyvr_ctx *c) {
    free_texture(c);
    cudaFree(c->texels); c->texels = nullptr;
    cudaFree(c->cell_minmax); c->cell_minmax = nullptr;
    cudaFree(c->cell_dist); c->cell_dist = nullptr;
    cudaFree(c->cell_scratch); c->cell_scratch = nullptr;
    cudaFree(c->active_box); c->active_box = nullptr;
    c->have_volume = false;
    c->vol.texels = nullptr;
    c->vol.cell_dist = nullptr;
    c->vol.active_box = nullptr;
}

int ensure_views(pyvr_ctx *c, int n) {
    if (n <= c->views_cap) return PYVR_OK;
    cudaFree(c->d_views);
    c->d_views = nullptr;
    c->views_cap = 0;
    CU(cudaMalloc(&c->d_views, sizeof(pyvr_view) * (size_t)n));
    c->views_cap = n;
    return PYVR_OK;
}

int ensure_events(pyvr_ctx *c, size_t pairs) {
    while (c->ev.size() < 2 * pairs) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        c->ev.push_back(e);
    }
    return PYVR_OK;
}

MarchArgs make_args(const pyvr_ctx *c) {
    MarchArgs a{};
    a.vol = c->vol;
    const long long eb = entry_bytes(c->half_texels, c->vol.pair);
    if (c->vol.brick8) {   // bytes per z-row / x-plane of 8-texel bricks, from the allocation start
        a.tap_base = static_cast<const char *>(c->vol.texels);
        a.stride_y = (long long)c->vol.pitch_y * eb * 8;
        a.stride_x = c->vol.pitch_x * eb * 8;
    } else {
        a.tap_base = static_cast<const char *>(c->vol.texels) + texel_index(c->vol, 0, 0, 0) * eb;
        a.stride_y = (long long)c->vol.pitch_y * eb;
        a.stride_x = c->vol.pitch_x * eb;
    }
    a.lut = c->lut;
    a.lut_size = c->lut_size;
    a.width = c->width;
    a.height = c->height;
    const pyvr_params &p = c->params;
    a.step = p.step_size;
    a.ref_step = p.reference_step_size;
    a.exp2_scale = (float)(-((double)p.step_size / (double)p.reference_step_size) * 1.4426950408889634);
    a.max_steps = p.max_steps;
    a.ambient = p.ambient;
    a.diffuse = p.diffuse;
    // normalize(light_target - light_position) in binary32, as the shader does (volume.frag.glsl:107)
    const float lx = p.light_target[0] - p.light_position[0], ly = p.light_target[1] - p.light_position[1],
                lz = p.light_target[2] - p.light_position[2];
    const float inv = 1.0f / sqrtf(lx * lx + ly * ly + lz * lz);
    a.ldir[0] = lx * inv; a.ldir[1] = ly * inv; a.ldir[2] = lz * inv;
    a.term_alpha = p.termination_alpha;
    a.flags = p.flags;
    a.counters = c->d_counters;
    a.shard_rank = c->shard_rank;
    a.shard_count = c->shard_count;
    a.shard_shift = c->shard_shift;
    a.tile_counter = reinterpret_cast<unsigned *>(c->d_counters + CNT_N);
    // several samples in flight per lane pay off when the rays are sparse (choose_layout) -- on C3 f16x4 they cost 2x --
    // and when most rays of the launch have work: a sort-last brick is crossed by a fraction of the rays, the rest only
    // pay for set-up, and 3 CTAs per SM instead of 9 then cost more than the deeper rays gain (C5 on 8 GPUs, march of
    // the slowest rank: 0.44 -> 0.54 ms; profiles/r02_c5_layout_ab.txt)
    a.two_samples = c->two_option < 0 ? (sparse_rays_expected(c, c->vol.gn) && !c->vol.bricked) : c->two_option;
    return a;
}

// Tile row in which the volume's centre appears (MarchArgs.first_row): CTAs are dispatched in block-index order, and
// the march maps consecutive blockIdx.y to rows from that one outwards, so the tiles with the longest rays start
// first and the empty rows at the top and bottom of the image finish the launch.  -1 = keep the natural order.
int centre_row(const pyvr_ctx *c, const pyvr_view *vw) {
    if (!vw) return -1;
    double cdir[3], vv = 0.0, ww = 0.0, cv = 0.0, cw = 0.0;
    for (int a = 0; a < 3; ++a) {
        cdir[a] = 0.5 * ((double)c->vol.bmin[a] + (double)c->vol.bmax[a]) - (double)vw->origin[a];
        vv += (double)vw->v[a] * vw->v[a]; ww += (double)vw->w[a] * vw->w[a];
        cv += cdir[a] * vw->v[a]; cw += cdir[a] * vw->w[a];
    }
    if (!(vv > 0.0) || !(ww > 0.0) || !(cw > 0.0)) return -1;      // no closed-form basis, or the centre is behind the camera
    const double ndy = (cv / vv) / (cw / ww);                       // dir = w + ndx*u + ndy*v, u, v, w orthogonal
    const double py = (ndy * 0.5 + 0.5) * (double)c->height;
    const int rows = (c->height + 7) / 8;                           // TILE_H = 8 (march.cu)
    int row = (int)floor(py / 8.0);
    if (row < 0) row = 0;
    if (row > rows - 1) row = rows - 1;
    return row;
}

// Launch the march for `n` views already resident in c->d_views[first..], timed with events.
int march(pyvr_ctx *c, int first, int n, uchar4 *out8, float4 *out_acc, size_t ev_pair,
          const float4 *in_acc = nullptr) {
    if ((c->params.flags & PYVR_FLAG_HWTEX) && !(c->params.flags & PYVR_FLAG_STRICT)) {
        int rc = ensure_texture(c);
        if (rc != PYVR_OK) return rc;
    }
    MarchArgs a = make_args(c);
    a.tex = c->tex_obj;
    a.views = c->d_views + first;
    a.out8 = out8;
    a.out_acc = out_acc;
    a.in_acc = in_acc;
    // measured on C4 (profiles/r02_c4_shard_probe.txt): whole frame 8.16 -> 7.83 ms, but one rank of an 8-way deal
    // 1.43 -> 1.58 ms (its few heavy tiles then all start at once), so sharded launches keep the natural order
    a.first_row = c->shard_count > 1 ? -1 : centre_row(c, c->launch_view);
    c->launch_view = nullptr;
    if (c->shard_count > 1) {
        if (in_acc) return fail(PYVR_ERR_STATE, "relay rendering cannot be combined with image-space sharding");
        if (!c->shard_in_place) {   // the launch covers this rank's tiles only: everything else reads as cleared
            if (out8) CU(cudaMemsetAsync(out8, 0, frame_pixels(c) * sizeof(uchar4) * (size_t)n, c->stream));
            if (out_acc) CU(cudaMemsetAsync(out_acc, 0, frame_pixels(c) * sizeof(float4) * (size_t)n, c->stream));
        }
    }
    CU(cudaEventRecord(c->ev[2 * ev_pair], c->stream));
    CU(launch_march(a, n, c->half_texels, c->texel_bytes / entry_bytes(c->half_texels, c->vol.pair) >= ((size_t)1 << 31), c->stream));
    CU(cudaEventRecord(c->ev[2 * ev_pair + 1], c->stream));
    return PYVR_OK;
}

int resolve_stats(pyvr_ctx *c) {
    if (!c->stats_pending) return PYVR_OK;
    c->stats_pending = false;
    CU(cudaEventSynchronize(c->stats_ready));
    float total = 0.0f;
    for (size_t i = 0; i < c->pending_pairs; ++i) {
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, c->ev[2 * i], c->ev[2 * i + 1]));
        total += ms;
    }
    c->stats.samples = c->h_counters[CNT_SAMPLES];
    c->stats.samples_fetched = c->h_counters[CNT_FETCHED];
    c->stats.rays_hit = c->h_counters[CNT_HIT];
    c->stats.rays_terminated = c->h_counters[CNT_TERM];
    c->stats.kernel_ms = total;
    c->stats.kernel_launches = (uint32_t)c->pending_pairs;
    c->stats.views = (uint32_t)c->pending_views;
    return PYVR_OK;
}

// Counters -> pyvr_stats.  defer (device output with option "async_device_output"): the copy of the counters is
// enqueued and the call returns; pyvr_cuda_get_stats waits for it.  Otherwise the stream is synchronised here.
int finish_stats(pyvr_ctx *c, size_t pairs, int views, bool defer = false) {
    CU(cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(unsigned long long) * CNT_N,
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->stats_ready, c->stream));
    c->stats_pending = true;
    c->pending_pairs = pairs;
    c->pending_views = views;
    return defer ? PYVR_OK : resolve_stats(c);
}

bool renderable(const pyvr_ctx *c) { return c->have_volume && c->lut_size > 0; }


// ---- host-array path of compute_normal_volume: slabs of axis 0 through pinned staging, pipelined -----------------
// The kernel needs 0.5 ms for 512^3; the host call is all copying (0.5 GiB in, 1.5 GiB out).  cudaMemcpy from / to
// pageable numpy memory moved that at ~6 GB/s (0.38 s).  Here the volume travels in slabs of kSlabPlanes planes (+ one
// halo plane on every cut side: interior planes then see their true neighbours, the kernel's one-sided differences only
// ever apply to the volume's own faces, and the halo planes' results are dropped -- the same split as
// multi_gpu.compute_normal_volume_sharded, bit-identical to the whole-volume call): host threads copy slab c+1 into
// pinned memory and slab c-1 out of it while slab c is on the device (H2D + kernel on one stream, D2H on another).
// The staging buffers are kept for the life of the process (pinned allocations cost ~0.3 ms per MiB).
struct NormalsStaging {
    int device = -1;
    size_t in_bytes = 0, out_bytes = 0;
    float *pin_in[2] = {nullptr, nullptr}, *pin_out[2] = {nullptr, nullptr}, *d_in[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr};
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, kernel_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
};
NormalsStaging g_normals_staging;
std::mutex g_normals_staging_mutex;      // one host-array call at a time uses the staging buffers

void parallel_copy(void *dst, const void *src, size_t bytes) {
    constexpr size_t kMinPerThread = (size_t)4 << 20;
    unsigned hw = std::thread::hardware_concurrency();
    size_t n = bytes / kMinPerThread;
    if (n > 8) n = 8;
    if (hw && n > hw) n = hw;
    if (n < 2) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> pool;
    const size_t part = ((bytes / n) + 63) & ~(size_t)63;
    for (size_t i = 1; i < n; ++i) {
        const size_t off = i * part, len = off >= bytes ? 0 : (i == n - 1 ? bytes - off : part);
        if (len) pool.emplace_back([=] { memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, len); });
    }
    memcpy(dst, src, part < bytes ? part : bytes);
    for (auto &t : pool) t.join();
}

cudaError_t ensure_normals_staging(NormalsStaging &st, int device, size_t in_bytes, size_t out_bytes) {
    cudaError_t e = cudaSuccess;
    if (st.device != device || st.in_bytes < in_bytes || st.out_bytes < out_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (st.pin_in[i]) cudaFreeHost(st.pin_in[i]);
            if (st.pin_out[i]) cudaFreeHost(st.pin_out[i]);
            if (st.d_in[i]) cudaFree(st.d_in[i]);
            if (st.d_out[i]) cudaFree(st.d_out[i]);
            st.pin_in[i] = st.pin_out[i] = st.d_in[i] = st.d_out[i] = nullptr;
        }
        st.in_bytes = st.out_bytes = 0;
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaMallocHost(&st.pin_in[i], in_bytes);
            if (e == cudaSuccess) e = cudaMallocHost(&st.pin_out[i], out_bytes);
            if (e == cudaSuccess) e = cudaMalloc(&st.d_in[i], in_bytes);
            if (e == cudaSuccess) e = cudaMalloc(&st.d_out[i], out_bytes);
        }
        if (e != cudaSuccess) return e;
        st.in_bytes = in_bytes; st.out_bytes = out_bytes; st.device = device;
    }
    if (!st.s_in) {
        e = cudaStreamCreateWithFlags(&st.s_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st.s_out, cudaStreamNonBlocking);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaEventCreateWithFlags(&st.h2d_done[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreate(&st.kernel_done[i]);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&st.d2h_done[i], cudaEventDisableTiming);
        }
    }
    return e;
}

constexpr int kSlabPlanes = 16;

// in / out: host arrays.  kernel_ms (may be NULL): sum of the slab kernels' device times.
cudaError_t normals_host_pipeline(int device, const float *in, float *out, int n0, int n1, int n2, bool relaxed, float *kernel_ms) {
    std::lock_guard<std::mutex> lock(g_normals_staging_mutex);
    NormalsStaging &st = g_normals_staging;
    const size_t plane = (size_t)n1 * n2;
    const int slab = n0 < kSlabPlanes ? n0 : kSlabPlanes;
    cudaError_t e = ensure_normals_staging(st, device, (size_t)(slab + 2) * plane * sizeof(float), (size_t)(slab + 2) * plane * 3 * sizeof(float));
    if (e != cudaSuccess) return e;
    const int n_slabs = (n0 + slab - 1) / slab;
    std::vector<cudaEvent_t> k0(n_slabs, nullptr);
    float total_ms = 0.0f;
    struct Pending { int p0, planes, slot; bool valid; } prev{0, 0, 0, false};
    auto drain = [&](const Pending &q) -> cudaError_t {      // slab q: device -> pinned is done, pinned -> caller's array
        cudaError_t d = cudaEventSynchronize(st.d2h_done[q.slot]);
        if (d != cudaSuccess) return d;
        parallel_copy(out + (size_t)q.p0 * plane * 3, st.pin_out[q.slot], (size_t)q.planes * plane * 3 * sizeof(float));
        float ms = 0.0f;
        d = cudaEventElapsedTime(&ms, k0[q.p0 / slab], st.kernel_done[q.slot]);
        total_ms += ms;
        return d;
    };
    for (int c = 0; c < n_slabs && e == cudaSuccess; ++c) {
        const int slot = c & 1, p0 = c * slab, planes = n0 - p0 < slab ? n0 - p0 : slab;
        const int lo = p0 > 0 ? 1 : 0, hi = p0 + planes < n0 ? 1 : 0, planes_in = planes + lo + hi;
        // pinned_in[slot] / d_in[slot] are free once the H2D copy / the kernel of slab c-2 are done; both are ordered on
        // s_in, and the host only has to wait for the copy before it overwrites the pinned buffer
        if (c >= 2) e = cudaEventSynchronize(st.h2d_done[slot]);
        if (e != cudaSuccess) break;
        parallel_copy(st.pin_in[slot], in + (size_t)(p0 - lo) * plane, (size_t)planes_in * plane * sizeof(float));
        e = cudaMemcpyAsync(st.d_in[slot], st.pin_in[slot], (size_t)planes_in * plane * sizeof(float), cudaMemcpyHostToDevice, st.s_in);
        if (e == cudaSuccess) e = cudaEventRecord(st.h2d_done[slot], st.s_in);
        // d_out[slot] is free once the D2H copy of slab c-2 has read it
        if (e == cudaSuccess && c >= 2) e = cudaStreamWaitEvent(st.s_in, st.d2h_done[slot], 0);
        if (e == cudaSuccess) e = cudaEventCreate(&k0[c]);
        if (e == cudaSuccess) e = cudaEventRecord(k0[c], st.s_in);
        if (e == cudaSuccess) e = launch_normals(st.d_in[slot], st.d_out[slot], planes_in, n1, n2, relaxed, st.s_in);
        if (e == cudaSuccess) e = cudaEventRecord(st.kernel_done[slot], st.s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st.s_out, st.kernel_done[slot], 0);
        // pinned_out[slot] was drained (slab c-2) in the previous iteration
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(st.pin_out[slot], st.d_out[slot] + (size_t)lo * plane * 3, (size_t)planes * plane * 3 * sizeof(float),
                                cudaMemcpyDeviceToHost, st.s_out);
        if (e == cudaSuccess) e = cudaEventRecord(st.d2h_done[slot], st.s_out);
        if (e == cudaSuccess && prev.valid) e = drain(prev);
        prev = Pending{p0, planes, slot, true};
    }
    if (e == cudaSuccess && prev.valid) e = drain(prev);
    cudaStreamSynchronize(st.s_in);
    cudaStreamSynchronize(st.s_out);
    for (cudaEvent_t ev : k0) if (ev) cudaEventDestroy(ev);
    if (e == cudaSuccess && kernel_ms) *kernel_ms = total_ms;
    return e;
}

}  // namespace

extern "C" {

int pyvr_cuda_abi_version(void) { return PYVR_CUDA_ABI_VERSION; }

const char *pyvr_cuda_last_error(void) { return g_error; }

int pyvr_cuda_device_count(int *count) {
    if (!count) return fail(PYVR_ERR_INVALID, "count is NULL");
    CU(cudaGetDeviceCount(count));
    return PYVR_OK;
}

int pyvr_cuda_create(int device, int width, int height, pyvr_ctx **out_ctx) {
    if (!out_ctx) return fail(PYVR_ERR_INVALID, "out_ctx is NULL");
    *out_ctx = nullptr;
    if (width <= 0 || height <= 0) return fail(PYVR_ERR_INVALID, "viewport %dx%d is not positive", width, height);
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(PYVR_ERR_INVALID, "device %d out of range (%d visible)", device, n_dev);
    DeviceGuard guard(device);
    pyvr_ctx *c = new (std::nothrow) pyvr_ctx();
    if (!c) return fail(PYVR_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->width = width;
    c->height = height;
    const char *layout = getenv("PYVR_CUDA_LAYOUT");
    if (layout) c->swizzle = strcmp(layout, "linear") != 0;   // "linear" = no swizzle, anything else = default
    const char *pair = getenv("PYVR_CUDA_PAIR");
    if (pair) c->pair_option = atoi(pair);
    const char *two = getenv("PYVR_CUDA_TWO_SAMPLES");
    if (two) c->two_option = atoi(two);
    const char *b8 = getenv("PYVR_CUDA_BRICK8");
    if (b8) c->brick8_option = atoi(b8);
    // defaults of the reference renderer: balanced preset, Light.default(), bounds +-0.5
    c->params.step_size = 0.01f; c->params.max_steps = 500; c->params.reference_step_size = 0.01f;
    c->params.ambient = 0.2f; c->params.diffuse = 0.8f;
    c->params.light_position[0] = c->params.light_position[1] = c->params.light_position[2] = 1.0f;
    c->params.termination_alpha = 0.99f;
    c->params.flags = PYVR_FLAG_ESS;
    cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&c->frames, frame_pixels(c) * sizeof(uchar4) * kRing * kSlotViews);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_counters, sizeof(unsigned long long) * (CNT_N + 1));   // + tile-queue ticket
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_counters, sizeof(unsigned long long) * CNT_N);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->stats_ready, cudaEventDisableTiming);
    for (int i = 0; i < kRing && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->slot_rendered[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->slot_copied[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        pyvr_cuda_destroy(c);
        return fail(e == cudaErrorMemoryAllocation ? PYVR_ERR_NOMEM : PYVR_ERR_CUDA,
                    "context creation failed: %s", cudaGetErrorString(e));
    }
    c->stream = c->own_stream;
    *out_ctx = c;
    return PYVR_OK;
}

int pyvr_cuda_destroy(pyvr_ctx *c) {
    if (!c) return PYVR_OK;
    DeviceGuard guard(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    free_volume(c);
    cudaFree(c->lut);
    cudaFree(c->d_views);
    cudaFree(c->frames);
    cudaFree(c->accum);
    cudaFree(c->d_counters);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    if (c->stats_ready) cudaEventDestroy(c->stats_ready);
    for (int i = 0; i < kRing; ++i) {
        if (c->slot_rendered[i]) cudaEventDestroy(c->slot_rendered[i]);
        if (c->slot_copied[i]) cudaEventDestroy(c->slot_copied[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
    return PYVR_OK;
}

int pyvr_cuda_set_stream(pyvr_ctx *c, void *cuda_stream) {
    if (!c) return fail(PYVR_ERR_INVALID, "ctx is NULL");
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    if (next != c->stream) {
        // work still pending on the old stream (uploads, cell classification, a march) stays ordered before
        // whatever is enqueued on the new one
        DeviceGuard guard(c->device);
        cudaEvent_t ev = nullptr;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, c->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(next, ev, 0);
        cudaEventDestroy(ev);
        CU(e);
    }
    c->stream = next;
    return PYVR_OK;
}

int pyvr_cuda_get_option(pyvr_ctx *c, const char *key, int *value) {
    if (!c || !key || !value) return fail(PYVR_ERR_INVALID, "NULL argument");
    // what is in effect for the volume that is loaded now (the automatic choices resolved)
    if (strcmp(key, "pair") == 0) *value = c->have_volume ? c->vol.pair : 0;
    else if (strcmp(key, "brick8") == 0) *value = c->have_volume ? c->vol.brick8 : 0;
    else if (strcmp(key, "two_samples") == 0)
        *value = c->have_volume && c->half_texels && (c->vol.pair || c->vol.brick8) &&
                 (c->two_option < 0 ? (sparse_rays_expected(c, c->vol.gn) && !c->vol.bricked) : c->two_option != 0);
    else if (strcmp(key, "async_device_output") == 0) *value = c->async_device;
    else return fail(PYVR_ERR_INVALID, "unknown option '%s'", key);
    return PYVR_OK;
}

int pyvr_cuda_set_option(pyvr_ctx *c, const char *key, int value) {
    if (!c || !key) return fail(PYVR_ERR_INVALID, "ctx or key is NULL");
    if (strcmp(key, "pair") == 0) {   // takes effect at the next upload
        c->pair_option = value < 0 ? -1 : (value != 0);
        return PYVR_OK;
    }
    if (strcmp(key, "async_device_output") == 0) {   // 1: renders into DEVICE buffers return without a host sync
        c->async_device = value != 0;
        return PYVR_OK;
    }
    if (strcmp(key, "two_samples") == 0) {   // f16x4 z-pair march, two samples per iteration (-1 auto, 0 off, 1 on)
        c->two_option = value < 0 ? -1 : (value != 0);
        return PYVR_OK;
    }
    if (strcmp(key, "brick8") == 0) {   // 2x2x2-texel bricks (-1 auto, 0 off, 1 on); takes effect at the next upload
        c->brick8_option = value < 0 ? -1 : (value != 0);
        return PYVR_OK;
    }
    if (strcmp(key, "shard_shift") == 0) {      // tile-group edge of the image-space sharding, in CTA tiles (log2)
        if (value < 0 || value > 6) return fail(PYVR_ERR_INVALID, "shard_shift must be in [0, 6]");
        c->shard_shift = value;
        return PYVR_OK;
    }
    if (strcmp(key, "shard_in_place") == 0) {   // 1: a sharded render leaves the other ranks' pixels untouched
        c->shard_in_place = value != 0;
        return PYVR_OK;
    }
    if (strcmp(key, "swizzle") == 0) {
        if (c->have_volume && (value != 0) != c->swizzle)
            return fail(PYVR_ERR_STATE, "swizzle must be chosen before the volume is uploaded");
        c->swizzle = value != 0;
        return PYVR_OK;
    }
    return fail(PYVR_ERR_INVALID, "unknown option '%s'", key);
}

}  // extern "C" (reopened below)

namespace {

// Device buffers of a volume whose c->vol / c->texel_bytes / c->n_cells are set.  On failure nothing stays allocated.
int alloc_volume_buffers(pyvr_ctx *c) {
    cudaError_t e = cudaMalloc(&c->texels, c->texel_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->cell_minmax, c->n_cells * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&c->cell_dist, c->n_cells);
    if (e == cudaSuccess) e = cudaMalloc(&c->cell_scratch, c->n_cells);
    if (e == cudaSuccess) e = cudaMalloc(&c->active_box, 6 * sizeof(int));
    // row / plane padding is never read; keep it defined
    if (e == cudaSuccess) e = cudaMemsetAsync(c->texels, 0, c->texel_bytes, c->stream);
    if (e != cudaSuccess) {
        free_volume(c);
        CU(e);
    }
    c->vol.texels = c->texels;
    c->vol.cell_dist = c->cell_dist;
    c->vol.active_box = c->active_box;
    return PYVR_OK;
}

// Shared tail of upload_volume / upload_brick: allocate, stage, pack, build the macrocell grid.
// c->vol (dims, ownership, bounds) and c->half_texels are already set.
int upload_packed(pyvr_ctx *c, const float *scalar, const float *normals, int src_is_device) {
    const VolumeDesc &v = c->vol;
    const size_t voxels = (size_t)v.n[0] * v.n[1] * v.n[2];
    const size_t n_tex = texel_count(c);
    c->texel_bytes = n_tex * entry_bytes(c->half_texels, c->vol.pair);
    c->n_cells = (size_t)v.ncell[0] * v.ncell[1] * v.ncell[2];
    int rc = alloc_volume_buffers(c);
    if (rc != PYVR_OK) return rc;

    // host sources go through device staging buffers; one exit frees them whatever fails (an out-of-memory here
    // must not leak multi-GB buffers for the life of the process), and a failed upload leaves no volume behind
    const float *d_scalar = scalar, *d_normals = normals;
    float *stage_s = nullptr, *stage_n = nullptr;
    cudaError_t e = cudaSuccess;
    if (!src_is_device) {
        e = cudaMalloc(&stage_s, voxels * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpyAsync(stage_s, scalar, voxels * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        d_scalar = stage_s;
        if (e == cudaSuccess && normals) {
            e = cudaMalloc(&stage_n, voxels * 3 * sizeof(float));
            if (e == cudaSuccess) e = cudaMemcpyAsync(stage_n, normals, voxels * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
            d_normals = stage_n;
        }
    }
    if (e == cudaSuccess) e = launch_pack_texels(d_scalar, d_normals, c->vol, c->half_texels, c->stream);
    if (e == cudaSuccess) e = launch_fill_apron(c->vol, c->half_texels, c->stream);
    if (e == cudaSuccess) e = launch_cell_minmax(c->vol, c->half_texels, c->cell_minmax, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    else cudaStreamSynchronize(c->stream);
    cudaFree(stage_s);
    cudaFree(stage_n);
    if (e != cudaSuccess) free_volume(c);
    CU(e);
    c->have_volume = true;
    return classify_cells(c);
}

}  // namespace

extern "C" {

int pyvr_cuda_upload_volume(pyvr_ctx *c, const float *scalar, const float *normals, int shape0, int shape1,
                            int shape2, const float bmin[3], const float bmax[3], int texel_format,
                            int src_is_device) {
    if (!c || !scalar || !bmin || !bmax) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (shape0 <= 0 || shape1 <= 0 || shape2 <= 0)
        return fail(PYVR_ERR_INVALID, "Volume data must be 3D with positive extents, got (%d, %d, %d)", shape0, shape1, shape2);
    if (texel_format != PYVR_TEXEL_F32X4 && texel_format != PYVR_TEXEL_F16X4)
        return fail(PYVR_ERR_INVALID, "unknown texel format %d", texel_format);
    for (int a = 0; a < 3; ++a)
        if (!(bmax[a] > bmin[a])) return fail(PYVR_ERR_INVALID, "max_bounds must be greater than min_bounds");
    DeviceGuard guard(c->device);
    CU(cudaStreamSynchronize(c->stream));
    free_volume(c);  // the reference leaks the previous textures (manager.py:232-236); not replicated

    c->half_texels = texel_format == PYVR_TEXEL_F16X4;
    // GL (width, height, depth) = (shape0, shape1, shape2); width is the memory-fastest axis and is
    // addressed by tex_coord.x = world z after the shader's swizzle (volume.frag.glsl:90).
    const int local[3] = {shape2, shape1, shape0};
    fill_volume_desc(c, local, nullptr, nullptr, nullptr, nullptr, bmin, bmax);
    return upload_packed(c, scalar, normals, src_is_device);
}

int pyvr_cuda_upload_brick(pyvr_ctx *c, const float *scalar, const float *normals, const int local_dims[3],
                           const int global_dims[3], const int origin[3], const int own_lo[3],
                           const int own_hi[3], const float bmin[3], const float bmax[3], int texel_format,
                           int src_is_device) {
    if (!c || !scalar || !local_dims || !global_dims || !origin || !own_lo || !own_hi || !bmin || !bmax)
        return fail(PYVR_ERR_INVALID, "NULL argument");
    if (texel_format != PYVR_TEXEL_F32X4 && texel_format != PYVR_TEXEL_F16X4)
        return fail(PYVR_ERR_INVALID, "unknown texel format %d", texel_format);
    for (int a = 0; a < 3; ++a) {
        if (!(bmax[a] > bmin[a])) return fail(PYVR_ERR_INVALID, "max_bounds must be greater than min_bounds");
        if (local_dims[a] <= 0 || global_dims[a] <= 0 || origin[a] < 0 || origin[a] + local_dims[a] > global_dims[a])
            return fail(PYVR_ERR_INVALID, "brick [%d, %d) does not fit axis %d of extent %d", origin[a],
                        origin[a] + local_dims[a], a, global_dims[a]);
        if (own_lo[a] < origin[a] || own_hi[a] <= own_lo[a] || own_hi[a] > global_dims[a])
            return fail(PYVR_ERR_INVALID, "ownership [%d, %d) on axis %d is outside the brick", own_lo[a], own_hi[a], a);
        // samples with x in [own_hi - 1, own_hi) read texel own_hi: it must be stored unless it is clamped away
        const int need_hi = own_hi[a] < global_dims[a] ? own_hi[a] : global_dims[a] - 1;
        if (need_hi > origin[a] + local_dims[a] - 1)
            return fail(PYVR_ERR_INVALID, "brick lacks the +1 ghost layer on axis %d", a);
    }
    DeviceGuard guard(c->device);
    CU(cudaStreamSynchronize(c->stream));
    free_volume(c);
    c->half_texels = texel_format == PYVR_TEXEL_F16X4;
    fill_volume_desc(c, local_dims, global_dims, origin, own_lo, own_hi, bmin, bmax);
    return upload_packed(c, scalar, normals, src_is_device);
}

int pyvr_cuda_generate_volume(pyvr_ctx *c, int shape, int size, const int local_dims[3], const int origin[3],
                              const int own_lo[3], const int own_hi[3], const float bmin[3], const float bmax[3],
                              int texel_format, float *kernel_ms) {
    if (!c || !bmin || !bmax) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (shape < PYVR_SHAPE_SPHERE || shape > PYVR_SHAPE_DOUBLE_SPHERE) return fail(PYVR_ERR_INVALID, "unknown shape %d", shape);
    if (size < 2) return fail(PYVR_ERR_INVALID, "size must be at least 2");
    if (texel_format != PYVR_TEXEL_F32X4 && texel_format != PYVR_TEXEL_F16X4)
        return fail(PYVR_ERR_INVALID, "unknown texel format %d", texel_format);
    const bool brick = local_dims || origin || own_lo || own_hi;
    if (brick && !(local_dims && origin && own_lo && own_hi))
        return fail(PYVR_ERR_INVALID, "a brick needs local_dims, origin, own_lo and own_hi");
    const int global[3] = {size, size, size};
    for (int a = 0; a < 3; ++a) {
        if (!(bmax[a] > bmin[a])) return fail(PYVR_ERR_INVALID, "max_bounds must be greater than min_bounds");
        if (!brick) continue;
        if (local_dims[a] <= 0 || origin[a] < 0 || origin[a] + local_dims[a] > size)
            return fail(PYVR_ERR_INVALID, "brick [%d, %d) does not fit axis %d of extent %d", origin[a],
                        origin[a] + local_dims[a], a, size);
        const int need_hi = own_hi[a] < size ? own_hi[a] : size - 1;
        if (own_lo[a] < origin[a] || own_hi[a] <= own_lo[a] || own_hi[a] > size || need_hi > origin[a] + local_dims[a] - 1)
            return fail(PYVR_ERR_INVALID, "ownership [%d, %d) on axis %d does not fit the brick (+1 ghost layer)", own_lo[a], own_hi[a], a);
    }
    DeviceGuard guard(c->device);
    CU(cudaStreamSynchronize(c->stream));
    free_volume(c);
    c->half_texels = texel_format == PYVR_TEXEL_F16X4;
    if (brick) fill_volume_desc(c, local_dims, global, origin, own_lo, own_hi, bmin, bmax);
    else fill_volume_desc(c, global, nullptr, nullptr, nullptr, nullptr, bmin, bmax);

    const VolumeDesc &v = c->vol;
    const size_t n_tex = texel_count(c);
    c->texel_bytes = n_tex * entry_bytes(c->half_texels, c->vol.pair);
    c->n_cells = (size_t)v.ncell[0] * v.ncell[1] * v.ncell[2];
    int rc = alloc_volume_buffers(c);
    if (rc != PYVR_OK) return rc;

    // slab scratch: about 1 GiB, at least 3 planes
    const size_t plane = (size_t)(v.n[1] + 2) * (size_t)(v.n[2] + 2);
    size_t planes = ((size_t)1 << 28) / plane;
    if (planes < 3) planes = 3;
    if (planes > (size_t)v.n[0] + 2) planes = (size_t)v.n[0] + 2;
    float *scratch = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaMalloc(&scratch, planes * plane * sizeof(float));
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaEventRecord(e0, c->stream);
    if (e == cudaSuccess) e = launch_synth_volume(c->vol, c->half_texels, shape, size, scratch, planes * plane, c->stream);
    if (e == cudaSuccess) e = launch_fill_apron(c->vol, c->half_texels, c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(e1, c->stream);
    if (e == cudaSuccess) e = launch_cell_minmax(c->vol, c->half_texels, c->cell_minmax, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess && kernel_ms) e = cudaEventElapsedTime(kernel_ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(scratch);
    if (e != cudaSuccess) free_volume(c);
    CU(e);
    c->have_volume = true;
    return classify_cells(c);
}

int pyvr_cuda_read_texels(pyvr_ctx *c, float *scalar, float *normals) {
    if (!c) return fail(PYVR_ERR_INVALID, "ctx is NULL");
    if (!c->have_volume) return fail(PYVR_ERR_STATE, "no volume loaded");
    DeviceGuard guard(c->device);
    const size_t voxels = (size_t)c->vol.n[0] * c->vol.n[1] * c->vol.n[2];
    float *d_s = nullptr, *d_n = nullptr;
    cudaError_t e = cudaSuccess;
    if (scalar) e = cudaMalloc(&d_s, voxels * sizeof(float));
    if (e == cudaSuccess && normals) e = cudaMalloc(&d_n, voxels * 3 * sizeof(float));
    if (e == cudaSuccess) e = launch_unpack_texels(c->vol, c->half_texels, d_s, d_n, c->stream);
    if (e == cudaSuccess && scalar) e = cudaMemcpyAsync(scalar, d_s, voxels * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && normals) e = cudaMemcpyAsync(normals, d_n, voxels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_s);
    cudaFree(d_n);
    CU(e);
    return PYVR_OK;
}

int pyvr_cuda_set_pixel_shard(pyvr_ctx *c, int rank, int count) {
    if (!c) return fail(PYVR_ERR_INVALID, "ctx is NULL");
    if (count < 1 || rank < 0 || rank >= count) return fail(PYVR_ERR_INVALID, "shard %d of %d", rank, count);
    c->shard_rank = rank;
    c->shard_count = count;
    return PYVR_OK;
}

int pyvr_cuda_set_lut(pyvr_ctx *c, const float *rgba, int size) {
    if (!c || !rgba) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (size < 1) return fail(PYVR_ERR_INVALID, "LUT size must be at least 1, got %d", size);
    if (((size_t)size + 1) * 2 * sizeof(float4) > 200 * 1024)   // pair-packed in shared memory (march.cu, stage_lut)
        return fail(PYVR_ERR_INVALID, "LUT of %d entries does not fit in shared memory (max 6399)", size);
    DeviceGuard guard(c->device);
    CU(cudaStreamSynchronize(c->stream));
    if (size != c->lut_size) {
        cudaFree(c->lut);
        c->lut = nullptr;
        c->lut_size = 0;
        CU(cudaMalloc(&c->lut, (size_t)size * sizeof(float4)));
    }
    CU(cudaMemcpyAsync(c->lut, rgba, (size_t)size * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    c->lut_size = size;
    CU(cudaStreamSynchronize(c->stream));
    return classify_cells(c);
}

int pyvr_cuda_view_from_matrices(const float view[16], const float proj[16], const float cam_pos[3],
                                 pyvr_view *out) {
    if (!view || !proj || !cam_pos || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    // GL reads the 16 floats column-major: M(r, c) = m[c*4 + r].
    double V[4][4], P[4][4], iV[4][4], iP[4][4];
    for (int r = 0; r < 4; ++r)
        for (int col = 0; col < 4; ++col) {
            V[r][col] = view[col * 4 + r];
            P[r][col] = proj[col * 4 + r];
        }
    if (!invert4(V, iV)) return fail(PYVR_ERR_INVALID, "view matrix is singular");
    if (!invert4(P, iP)) return fail(PYVR_ERR_INVALID, "projection matrix is singular");
    // eye.xy = (inverse(P) * (ndc.x, ndc.y, -1, 1)).xy ; world = inverse(V) * (eye.x, eye.y, -1, 0)
    const double ex_c = iP[0][3] - iP[0][2], ey_c = iP[1][3] - iP[1][2];
    for (int r = 0; r < 3; ++r) {
        out->u[r] = (float)(iV[r][0] * iP[0][0] + iV[r][1] * iP[1][0]);
        out->v[r] = (float)(iV[r][0] * iP[0][1] + iV[r][1] * iP[1][1]);
        out->w[r] = (float)(iV[r][0] * ex_c + iV[r][1] * ey_c - iV[r][2]);
        out->origin[r] = cam_pos[r];
    }
    inverse4_f32(proj, out->inv_proj);
    inverse4_f32(view, out->inv_view);
    out->has_matrices = 1;
    return PYVR_OK;
}

int pyvr_cuda_set_camera(pyvr_ctx *c, const float view[16], const float proj[16], const float cam_pos[3]) {
    if (!c) return fail(PYVR_ERR_INVALID, "ctx is NULL");
    pyvr_view v;
    int rc = pyvr_cuda_view_from_matrices(view, proj, cam_pos, &v);
    if (rc != PYVR_OK) return rc;
    c->view = v;
    c->have_view = true;
    return PYVR_OK;
}

int pyvr_cuda_set_view(pyvr_ctx *c, const pyvr_view *view) {
    if (!c || !view) return fail(PYVR_ERR_INVALID, "NULL argument");
    c->view = *view;
    c->have_view = true;
    return PYVR_OK;
}

int pyvr_cuda_set_params(pyvr_ctx *c, const pyvr_params *p) {
    if (!c || !p) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (!(p->step_size > 0.0f)) return fail(PYVR_ERR_INVALID, "step_size must be positive");
    if (p->max_steps < 1) return fail(PYVR_ERR_INVALID, "max_steps must be at least 1");
    if (!(p->reference_step_size > 0.0f)) return fail(PYVR_ERR_INVALID, "reference_step_size must be positive");
    c->params = *p;
    return PYVR_OK;
}

int pyvr_cuda_render_batch(pyvr_ctx *c, const pyvr_view *views, int n, uint8_t *out, int out_is_device) {
    if (!c || !views || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (n < 1) return fail(PYVR_ERR_INVALID, "batch of %d views", n);
    DeviceGuard guard(c->device);
    const size_t frame_bytes = frame_pixels(c) * sizeof(uchar4);
    memset(&c->stats, 0, sizeof c->stats);
    c->stats_pending = false;      // stats of an earlier asynchronous render that nobody asked for
    if (!renderable(c)) {  // cleared framebuffer
        if (out_is_device) {
            CU(cudaMemsetAsync(out, 0, frame_bytes * n, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        } else {
            memset(out, 0, frame_bytes * n);
        }
        c->stats.views = (uint32_t)n;
        return PYVR_OK;
    }
    int rc = ensure_views(c, n);
    if (rc != PYVR_OK) return rc;
    CU(cudaMemcpyAsync(c->d_views, views, sizeof(pyvr_view) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned long long) * CNT_N, c->stream));

    size_t pairs = 0;
    if (out_is_device) {
        // frames land directly in the caller's device buffer, up to 16 views per launch
        const int chunk = 16;
        rc = ensure_events(c, (size_t)(n + chunk - 1) / chunk);
        if (rc != PYVR_OK) return rc;
        for (int first = 0; first < n; first += chunk) {
            const int m = n - first < chunk ? n - first : chunk;
            c->launch_view = views + first;
            rc = march(c, first, m, reinterpret_cast<uchar4 *>(out) + (size_t)first * frame_pixels(c), nullptr, pairs++);
            if (rc != PYVR_OK) return rc;
        }
    } else {
        // ring of device slots: the march of group g overlaps the device->host copy of group g-1.  A group is
        // up to kSlotViews views in one launch (fewer launch tails than one launch per view); the groups taper
        // towards the end (.., 4, 2, 1, 1) so that the copy left exposed after the last march is one frame.
        rc = ensure_events(c, (size_t)n);
        if (rc != PYVR_OK) return rc;
        int first = 0;
        for (int g = 0; first < n; ++g) {
            const int remaining = n - first;
            int m = remaining / 2 < 1 ? 1 : remaining / 2;
            if (m > kSlotViews) m = kSlotViews;
            const int slot = g % kRing;
            if (g >= kRing) CU(cudaStreamWaitEvent(c->stream, c->slot_copied[slot], 0));
            uchar4 *frames = c->frames + (size_t)slot * kSlotViews * frame_pixels(c);
            c->launch_view = views + first;
            rc = march(c, first, m, frames, nullptr, pairs++);
            if (rc != PYVR_OK) return rc;
            CU(cudaEventRecord(c->slot_rendered[slot], c->stream));
            CU(cudaStreamWaitEvent(c->copy_stream, c->slot_rendered[slot], 0));
            CU(cudaMemcpyAsync(out + (size_t)first * frame_bytes, frames, frame_bytes * m, cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaEventRecord(c->slot_copied[slot], c->copy_stream));
            first += m;
        }
        CU(cudaStreamSynchronize(c->copy_stream));
    }
    return finish_stats(c, pairs, n, c->async_device && out_is_device);
}

int pyvr_cuda_render(pyvr_ctx *c, uint8_t *out, int out_is_device) {
    if (!c || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (!c->have_view) {  // the reference renders garbage/black without a camera; return the cleared frame
        DeviceGuard guard(c->device);
        const size_t frame_bytes = frame_pixels(c) * sizeof(uchar4);
        memset(&c->stats, 0, sizeof c->stats);
        if (out_is_device) {
            CU(cudaMemsetAsync(out, 0, frame_bytes, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        } else {
            memset(out, 0, frame_bytes);
        }
        return PYVR_OK;
    }
    return pyvr_cuda_render_batch(c, &c->view, 1, out, out_is_device);
}

int pyvr_cuda_render_accum(pyvr_ctx *c, float *out, int out_is_device) {
    if (!c || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    DeviceGuard guard(c->device);
    const size_t bytes = frame_pixels(c) * sizeof(float4);
    memset(&c->stats, 0, sizeof c->stats);
    c->stats_pending = false;
    if (!renderable(c) || !c->have_view) {
        if (out_is_device) {
            CU(cudaMemsetAsync(out, 0, bytes, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        } else {
            memset(out, 0, bytes);
        }
        return PYVR_OK;
    }
    float4 *target = reinterpret_cast<float4 *>(out);
    if (!out_is_device) {
        if (!c->accum) CU(cudaMalloc(&c->accum, bytes));
        target = c->accum;
    }
    int rc = ensure_views(c, 1);
    if (rc == PYVR_OK) rc = ensure_events(c, 1);
    if (rc != PYVR_OK) return rc;
    CU(cudaMemcpyAsync(c->d_views, &c->view, sizeof(pyvr_view), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned long long) * CNT_N, c->stream));
    c->launch_view = &c->view;
    rc = march(c, 0, 1, nullptr, target, 0);
    if (rc != PYVR_OK) return rc;
    if (!out_is_device) CU(cudaMemcpyAsync(out, c->accum, bytes, cudaMemcpyDeviceToHost, c->stream));
    return finish_stats(c, 1, 1, c->async_device && out_is_device);
}

int pyvr_cuda_render_accum_relay(pyvr_ctx *c, const float *in_accum, float *out_accum) {
    if (!c || !out_accum) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (!renderable(c) || !c->have_view) return fail(PYVR_ERR_STATE, "relay rendering needs a volume, a LUT and a camera");
    if (c->params.flags & PYVR_FLAG_STRICT) return fail(PYVR_ERR_STATE, "relay rendering is a fast-path feature");
    DeviceGuard guard(c->device);
    memset(&c->stats, 0, sizeof c->stats);
    c->stats_pending = false;
    int rc = ensure_views(c, 1);
    if (rc == PYVR_OK) rc = ensure_events(c, 1);
    if (rc != PYVR_OK) return rc;
    CU(cudaMemcpyAsync(c->d_views, &c->view, sizeof(pyvr_view), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned long long) * CNT_N, c->stream));
    c->launch_view = &c->view;
    rc = march(c, 0, 1, nullptr, reinterpret_cast<float4 *>(out_accum), 0, reinterpret_cast<const float4 *>(in_accum));
    if (rc != PYVR_OK) return rc;
    return finish_stats(c, 1, 1, c->async_device);   // relay buffers are device buffers
}

int pyvr_cuda_get_stats(pyvr_ctx *c, pyvr_stats *out) {
    if (!c || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    DeviceGuard guard(c->device);
    int rc = resolve_stats(c);
    if (rc != PYVR_OK) return rc;
    *out = c->stats;
    return PYVR_OK;
}

int pyvr_cuda_compute_normals(int device, const float *in, float *out, int n0, int n1, int n2,
                              int flags, float *kernel_ms) {
    const bool buffers_are_device = (flags & PYVR_NORMALS_DEVICE_BUFFERS) != 0, relaxed = (flags & PYVR_NORMALS_RELAXED) != 0;
    if (!in || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (n0 <= 0 || n1 <= 0 || n2 <= 0) return fail(PYVR_ERR_INVALID, "Volume data must be 3D with positive extents");
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(PYVR_ERR_INVALID, "device %d out of range (%d visible)", device, n_dev);
    DeviceGuard guard(device);
    const size_t voxels = (size_t)n0 * n1 * n2;
    // host arrays of some size: slabs through pinned staging (PYVR_NORMALS_HOST_PIPELINE=0: the plain staged copy below)
    static const bool pipeline = !(getenv("PYVR_NORMALS_HOST_PIPELINE") && atoi(getenv("PYVR_NORMALS_HOST_PIPELINE")) == 0);
    if (!buffers_are_device && pipeline && voxels >= ((size_t)1 << 22) && (size_t)n1 * n2 * (kSlabPlanes + 2) * 16 <= ((size_t)1 << 31)) {
        CU(normals_host_pipeline(device, in, out, n0, n1, n2, relaxed, kernel_ms));
        return PYVR_OK;
    }
    const float *d_in = in;
    float *d_out = out, *stage_in = nullptr, *stage_out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaSuccess;
    if (!buffers_are_device) {
        e = cudaMalloc(&stage_in, voxels * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&stage_out, voxels * 3 * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(stage_in, in, voxels * sizeof(float), cudaMemcpyHostToDevice);
        d_in = stage_in;
        d_out = stage_out;
    }
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaEventRecord(e0, 0);
    if (e == cudaSuccess) e = launch_normals(d_in, d_out, n0, n1, n2, relaxed, 0);
    if (e == cudaSuccess) e = cudaEventRecord(e1, 0);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess && kernel_ms) e = cudaEventElapsedTime(kernel_ms, e0, e1);
    if (e == cudaSuccess && !buffers_are_device)
        e = cudaMemcpy(out, stage_out, voxels * 3 * sizeof(float), cudaMemcpyDeviceToHost);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(stage_in);
    cudaFree(stage_out);
    CU(e);
    return PYVR_OK;
}

int pyvr_cuda_measure_cache_bandwidth(int device, int level, double *gbs) {
    if (!gbs) return fail(PYVR_ERR_INVALID, "gbs is NULL");
    if (level < 1 || level > 3) return fail(PYVR_ERR_INVALID, "level must be 1 (L1), 2 (L2) or 3 (DRAM, 1:3 read:write mix)");
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(PYVR_ERR_INVALID, "device %d out of range (%d visible)", device, n_dev);
    DeviceGuard guard(device);
    CU(measure_cache_bandwidth(level, gbs));
    return PYVR_OK;
}

int pyvr_cuda_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(PYVR_ERR_INVALID, "out is NULL");
    CU(cudaMallocHost(out, bytes));
    return PYVR_OK;
}

int pyvr_cuda_host_free(void *p) {
    if (p) CU(cudaFreeHost(p));
    return PYVR_OK;
}

int pyvr_cuda_composite_over(int device, const float *front, const float *back, float *out, size_t n_pixels,
                             float termination_alpha, void *cuda_stream) {
    if (!front || !back || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    DeviceGuard guard(device);
    CU(launch_composite_over(reinterpret_cast<const float4 *>(front), reinterpret_cast<const float4 *>(back),
                             reinterpret_cast<float4 *>(out), n_pixels, termination_alpha, (cudaStream_t)cuda_stream));
    return PYVR_OK;
}

int pyvr_cuda_finalize_rgba8(int device, const float *accum, uint8_t *out, size_t n_pixels, uint32_t flags,
                             void *cuda_stream) {
    if (!accum || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    DeviceGuard guard(device);
    CU(launch_finalize_rgba8(reinterpret_cast<const float4 *>(accum), reinterpret_cast<uchar4 *>(out), n_pixels,
                             flags, (cudaStream_t)cuda_stream));
    return PYVR_OK;
}

int pyvr_cuda_composite_finalize(int device, const float *front, const float *back, float *accum_out, uint8_t *out,
                                 size_t n_pixels, float termination_alpha, uint32_t flags, void *cuda_stream) {
    if (!front || !back || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    DeviceGuard guard(device);
    CU(launch_composite_finalize(reinterpret_cast<const float4 *>(front), reinterpret_cast<const float4 *>(back),
                                 reinterpret_cast<float4 *>(accum_out), reinterpret_cast<uchar4 *>(out), n_pixels,
                                 termination_alpha, flags, (cudaStream_t)cuda_stream));
    return PYVR_OK;
}

int pyvr_cuda_flag_signal(int device, uint32_t *flag, uint32_t value, void *cuda_stream) {
    if (!flag) return fail(PYVR_ERR_INVALID, "flag is NULL");
    DeviceGuard guard(device);
    CU(launch_flag_signal(flag, value, (cudaStream_t)cuda_stream));
    return PYVR_OK;
}

int pyvr_cuda_flag_signal_many(int device, uint32_t *const *flags, int n_flags, uint32_t value, void *cuda_stream) {
    if (!flags || n_flags < 1) return fail(PYVR_ERR_INVALID, "flags is NULL or empty");
    for (int i = 0; i < n_flags; ++i)
        if (!flags[i]) return fail(PYVR_ERR_INVALID, "flags[%d] is NULL", i);
    DeviceGuard guard(device);
    CU(launch_flag_signal_many(flags, n_flags, value, (cudaStream_t)cuda_stream));
    return PYVR_OK;
}

int pyvr_cuda_binary_swap(int device, const pyvr_swap_round *rounds, int n_rounds, uint32_t value,
                          float termination_alpha, uint32_t flags, void *cuda_stream) {
    if (!rounds || n_rounds < 0) return fail(PYVR_ERR_INVALID, "rounds is NULL");
    for (int r = 0; r < n_rounds; ++r) {
        const pyvr_swap_round &q = rounds[r];
        if (!q.front || !q.back || (!q.out && !q.out8)) return fail(PYVR_ERR_INVALID, "round %d: NULL image pointer", r);
    }
    DeviceGuard guard(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    for (int r = 0; r < n_rounds; ++r) {
        const pyvr_swap_round &q = rounds[r];
        if (q.signal_before) CU(launch_flag_signal(q.signal_before, value, st));
        if (q.wait_flag) CU(launch_flag_wait(q.wait_flag, 1, value, st));
        if (q.out8)
            CU(launch_composite_finalize(reinterpret_cast<const float4 *>(q.front), reinterpret_cast<const float4 *>(q.back),
                                         reinterpret_cast<float4 *>(q.out), reinterpret_cast<uchar4 *>(q.out8), (size_t)q.n_pixels,
                                         termination_alpha, flags, st));
        else
            CU(launch_composite_over(reinterpret_cast<const float4 *>(q.front), reinterpret_cast<const float4 *>(q.back),
                                     reinterpret_cast<float4 *>(q.out), (size_t)q.n_pixels, termination_alpha, st));
        unsigned *after[2];
        int n_after = 0;
        if (q.signal_done) after[n_after++] = q.signal_done;
        if (q.signal_next) after[n_after++] = q.signal_next;
        if (n_after) CU(launch_flag_signal_many(after, n_after, value, st));
    }
    return PYVR_OK;
}

int pyvr_cuda_flag_wait(int device, const uint32_t *flags, int n_flags, uint32_t value, void *cuda_stream) {
    if (!flags) return fail(PYVR_ERR_INVALID, "flags is NULL");
    if (n_flags < 1 || n_flags > 1024) return fail(PYVR_ERR_INVALID, "n_flags must be in [1, 1024]");
    DeviceGuard guard(device);
    CU(launch_flag_wait(flags, n_flags, value, (cudaStream_t)cuda_stream));
    return PYVR_OK;
}

// The driver packs cudaMalloc requests below 2 MiB into shared 2 MiB blocks, and a CUDA IPC handle names the BLOCK:
// the process that opens it gets the block's base address, not the buffer's (round 2: a 1.9 MB partial image read by
// its binary-swap partner at the wrong offset).  Every buffer handed out here is therefore a whole number of 2 MiB,
// i.e. an allocation of its own, and pyvr_cuda_ipc_export refuses pointers that are not the start of one.
constexpr size_t kIpcGranule = (size_t)2 << 20;

int pyvr_cuda_device_alloc(int device, size_t bytes, void **out) {
    if (!out) return fail(PYVR_ERR_INVALID, "out is NULL");
    DeviceGuard guard(device);
    const size_t rounded = ((bytes > 0 ? bytes : 1) + kIpcGranule - 1) / kIpcGranule * kIpcGranule;
    CU(cudaMalloc(out, rounded));
    return PYVR_OK;
}

int pyvr_cuda_device_free(int device, void *ptr) {
    DeviceGuard guard(device);
    if (ptr) CU(cudaFree(ptr));
    return PYVR_OK;
}

int pyvr_cuda_ipc_export(int device, void *ptr, uint8_t handle[PYVR_IPC_HANDLE_BYTES]) {
    if (!ptr || !handle) return fail(PYVR_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == PYVR_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(device);
    {   // the handle would map the enclosing allocation: insist that `ptr` starts one (cuMemGetAddressRange)
        typedef int (*RangeFn)(unsigned long long *, size_t *, unsigned long long);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess && fn) {
            unsigned long long base = 0;
            size_t size = 0;
            if (reinterpret_cast<RangeFn>(fn)(&base, &size, (unsigned long long)(uintptr_t)ptr) == 0 &&
                base != (unsigned long long)(uintptr_t)ptr)
                return fail(PYVR_ERR_INVALID, "IPC export needs the start of an allocation (use pyvr_cuda_device_alloc)");
        }
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, sizeof h);
    return PYVR_OK;
}

int pyvr_cuda_ipc_open(int device, const uint8_t handle[PYVR_IPC_HANDLE_BYTES], void **out) {
    if (!handle || !out) return fail(PYVR_ERR_INVALID, "NULL argument");
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CU(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return PYVR_OK;
}

int pyvr_cuda_ipc_close(int device, void *ptr) {
    DeviceGuard guard(device);
    if (ptr) CU(cudaIpcCloseMemHandle(ptr));
    return PYVR_OK;
}

int pyvr_cuda_memcpy(int device, void *dst, const void *src, size_t bytes, int kind, void *cuda_stream) {
    if (!dst || !src) return fail(PYVR_ERR_INVALID, "NULL argument");
    if (kind < 1 || kind > 3) return fail(PYVR_ERR_INVALID, "kind must be 1 (H2D), 2 (D2H) or 3 (D2D)");
    DeviceGuard guard(device);
    const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CU(cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)cuda_stream));
    CU(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    return PYVR_OK;
}

int pyvr_cuda_stream_synchronize(int device, void *cuda_stream) {
    DeviceGuard guard(device);
    CU(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    return PYVR_OK;
}

}  // extern "C"
