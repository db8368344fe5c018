// march.cu -- K1: the volume ray-march kernel (sm_100a).
//
// Replaces the fragment shader pyvr/shaders/volume.frag.glsl:70-122 of the reference together with the
// fixed-function stages after it: blend SRC_ALPHA/ONE_MINUS_SRC_ALPHA onto a cleared target
// (pyvr/moderngl_renderer/manager.py:217-220), RGBA8 quantisation (manager.py:29) and the bottom-up
// row order of fbo.read (manager.py:228-230).
//
// Mapping: one thread per ray; a warp owns an 8x4 pixel tile, a CTA (4 warps) a 16x8 tile, so the 8
// texel gathers of neighbouring rays land in the same 128-byte lines of the packed {s,nx,ny,nz}
// texel array (L1 does the gather amplification, L2/HBM stream each touched line once).  The RGBA
// transfer-function LUT is staged in shared memory once per CTA.
//
// Ray set-up (direction, slab test, entry point, validity of the entry sample) uses the oracle's
// arithmetic in every mode.  After that (MarchArgs.flags, template parameters):
//   STRICT  statement-by-statement twin of the oracle: incremental position p += dir*step, IEEE
//           divide / sqrt / expf, no skipping.  Used to pin the kernel against oracle/pyvr_oracle.c.
//   fast    (default) the same sample lattice evaluated directly in voxel space,
//           x(i) = X0 + i*DX (one FMA per axis): the in-volume index range [i_lo, i_hi] is found once
//           per ray and clipped to the bounding box of the active macrocells; the ray then walks the
//           macrocell map once and records its active index intervals (exact empty-space skipping), and
//           a warp-synchronous loop samples them; binary32 software trilinear, ex2.approx / rsqrt.approx.
//           Differences to STRICT are ~1e-6 relative.
//   TEX     the fast march with the gather + filter replaced by one tex3D fetch (PYVR_FLAG_HWTEX).
//   BRICK   the fast march restricted to the samples a sort-last brick owns (pyvr_cuda_upload_brick);
//   PAIR    z-pair entries: one LDG.256 / LDG.128 per corner row;  HALF  binary16 texels;
//   in_acc  continue an incoming accumulation (relay);  shard_*  image-space tile sharding.
#include "common.cuh"

namespace pyvr {
namespace {

// Warp tile WARP_W x WARP_H pixels (one ray per lane), CTA = CTA_WX x CTA_WY warps.  Measured on C3: 8x4 warp
// tiles in 2x2-warp CTAs (16x8 pixels, 128 threads) beat 1x2, 2x4, 4x2 and 4x4 warps per CTA by 2-29 %.
#ifndef PYVR_WARP_W
#define PYVR_WARP_W 8
#endif
#ifndef PYVR_CTA_WARPS_X
#define PYVR_CTA_WARPS_X 2
#define PYVR_CTA_WARPS_Y 2
#endif
constexpr int WARP_W = PYVR_WARP_W, WARP_H = 32 / WARP_W;
constexpr int CTA_WX = PYVR_CTA_WARPS_X, CTA_WY = PYVR_CTA_WARPS_Y;
constexpr int TILE_W = CTA_WX * WARP_W, TILE_H = CTA_WY * WARP_H, CTA_THREADS = TILE_W * TILE_H;

struct Taps {
    int i0, i1;
    float f;
};

// One axis of a LINEAR / CLAMP_TO_EDGE fetch: texel centres at integer + 0.5 (GL 3.3 spec 3.8.8).
__device__ __forceinline__ Taps axis_taps(float u, int n) {
    float x = u * (float)n - 0.5f;
    float fl = floorf(x);
    Taps t;
    t.f = x - fl;
    int i = (int)fl;
    t.i0 = min(max(i, 0), n - 1);
    t.i1 = min(max(i + 1, 0), n - 1);
    return t;
}

__device__ __forceinline__ float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }

// ---- packed binary32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 work on 64-bit register pairs, one issue slot for two
// IEEE operations).  A texel {s, nx, ny, nz} is two such pairs, so the trilinear filter of all four channels is
// 14 packed lerps = 28 instructions instead of 56; every lane of a pair rounds exactly like the scalar fmaf().
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 lerp2(f32x2 a, f32x2 b, f32x2 t) {   // fma(t, b - a, a) per lane
    f32x2 d, r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(b), "l"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(t), "l"(d), "l"(a));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

struct Texel2 {
    f32x2 sn, yz;   // {s, nx}, {ny, nz}
};

__device__ __forceinline__ f32x2 half2_to_f32x2(unsigned raw) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&raw));
    return pack2(f.x, f.y);
}

// Both z-taps of one (x, y) corner row: entry(iz) and entry(iz + 1), or the two halves of one z-pair entry --
// a single LDG.E.256 (f32x4 pairs, sm_100+) or LDG.E.128 (f16x4 pairs).
template <bool HALF, bool PAIR>
__device__ __forceinline__ void load_row(const char *p, Texel2 &lo, Texel2 &hi) {
    if constexpr (HALF) {
        if constexpr (PAIR) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p));
            lo.sn = half2_to_f32x2(raw.x); lo.yz = half2_to_f32x2(raw.y);
            hi.sn = half2_to_f32x2(raw.z); hi.yz = half2_to_f32x2(raw.w);
        } else {
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(p)), b = __ldg(reinterpret_cast<const uint2 *>(p) + 1);
            lo.sn = half2_to_f32x2(a.x); lo.yz = half2_to_f32x2(a.y);
            hi.sn = half2_to_f32x2(b.x); hi.yz = half2_to_f32x2(b.y);
        }
    } else {
        if constexpr (PAIR) {
            asm("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];"
                : "=l"(lo.sn), "=l"(lo.yz), "=l"(hi.sn), "=l"(hi.yz) : "l"(p));
        } else {
            asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(lo.sn), "=l"(lo.yz) : "l"(p));
            asm("ld.global.nc.v2.b64 {%0, %1}, [%2+16];" : "=l"(hi.sn), "=l"(hi.yz) : "l"(p));
        }
    }
}

// f16x4 z-pair entry, raw: the 16 bytes of one corner row, converted when the sample is filtered (PAIRED samples below
// keep a second sample's four rows in flight in this form: 16 registers instead of 32).
__device__ __forceinline__ uint4 load_row_raw(const char *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void unpack_row(const uint4 &raw, Texel2 &lo, Texel2 &hi) {
    lo.sn = half2_to_f32x2(raw.x); lo.yz = half2_to_f32x2(raw.y);
    hi.sn = half2_to_f32x2(raw.z); hi.yz = half2_to_f32x2(raw.w);
}

// One texel (brick8 layout): LDG.64 (f16x4) or LDG.128 (f32x4).
template <bool HALF>
__device__ __forceinline__ void load_one(const char *p, Texel2 &t) {
    if constexpr (HALF) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(p));
        t.sn = half2_to_f32x2(raw.x); t.yz = half2_to_f32x2(raw.y);
    } else {
        asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(t.sn), "=l"(t.yz) : "l"(p));
    }
}

// STRICT path: one texel by index (clamped taps, never the apron), unpacked.
template <bool HALF>
__device__ __forceinline__ float4 load_texel(const VolumeDesc &v, int ix, int iy, int iz) {
    const long long e = texel_index(v, ix, iy, iz) << v.pair;    // first half of a z-pair entry = the texel itself
    if constexpr (HALF) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(v.texels) + e);
        const float2 fa = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
        const float2 fb = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    } else {
        return __ldg(reinterpret_cast<const float4 *>(v.texels) + e);
    }
}

struct Corner8 {
    float4 c[8];  // index = (x_tap << 2) | (y_tap << 1) | z_tap
};

template <bool HALF>
__device__ __forceinline__ Corner8 gather_strict(const VolumeDesc &v, const Taps &tx, const Taps &ty, const Taps &tz) {
    Corner8 r;
    r.c[0] = load_texel<HALF>(v, tx.i0, ty.i0, tz.i0); r.c[1] = load_texel<HALF>(v, tx.i0, ty.i0, tz.i1);
    r.c[2] = load_texel<HALF>(v, tx.i0, ty.i1, tz.i0); r.c[3] = load_texel<HALF>(v, tx.i0, ty.i1, tz.i1);
    r.c[4] = load_texel<HALF>(v, tx.i1, ty.i0, tz.i0); r.c[5] = load_texel<HALF>(v, tx.i1, ty.i0, tz.i1);
    r.c[6] = load_texel<HALF>(v, tx.i1, ty.i1, tz.i0); r.c[7] = load_texel<HALF>(v, tx.i1, ty.i1, tz.i1);
    return r;
}

// Filter order of the oracle: GL x (= world z, the fastest axis) first, then GL y (world y), then GL z (world x).
#define PYVR_TRILERP(field)                                                                       \
    lerpf(lerpf(lerpf(k.c[0].field, k.c[1].field, wz), lerpf(k.c[2].field, k.c[3].field, wz), wy), \
          lerpf(lerpf(k.c[4].field, k.c[5].field, wz), lerpf(k.c[6].field, k.c[7].field, wz), wy), wx)

__device__ __forceinline__ float rsqrt_approx(float x) {   // one MUFU.RSQ
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Accum {
    float r, g, b, a;
};

// The transfer-function LUT in shared memory with one apron entry on each side: entry j = lut[clamp(j - 1)], so a
// fetch is one clamp of floor(x) to [-1, size-1] and two LDS.128 sixteen bytes apart, and out-of-range densities
// land on two equal taps (CLAMP_TO_EDGE).  PYVR_LUT_PAIRED=1 (A/B only) pair-packs it like the z-pair texels: entry
// j holds {lut[max(j-1, 0)], lut[min(j, size-1)]}.
#ifndef PYVR_LUT_PAIRED
#define PYVR_LUT_PAIRED 0   // unpaired: half the shared memory and fewer LDS bank conflicts (+1 % on C3, profiles/r02_lane_ab.txt)
#endif
__host__ __device__ constexpr size_t lut_smem_bytes(int size) {
    return PYVR_LUT_PAIRED ? ((size_t)size + 1) * 2 * sizeof(float4) : ((size_t)size + 2) * sizeof(float4);
}
__device__ __forceinline__ void stage_lut(float4 *s_lut, const float4 *lut, int size, int n_threads) {
#if PYVR_LUT_PAIRED
    for (int j = threadIdx.x; j <= size; j += n_threads) {
        s_lut[2 * j] = lut[max(j - 1, 0)];
        s_lut[2 * j + 1] = lut[min(j, size - 1)];
    }
#else
    for (int j = threadIdx.x; j < size + 2; j += n_threads) s_lut[j] = lut[min(max(j - 1, 0), size - 1)];
#endif
}
// entry of lower tap i in [-1, size-1]: its two taps are at [0] and [1]
__device__ __forceinline__ const float4 *lut_entry(const float4 *s_lut, int i) {
    return PYVR_LUT_PAIRED ? s_lut + 2 * (i + 1) : s_lut + (i + 1);
}
__device__ __forceinline__ float4 lds128(const float4 *p) {   // one LDS.128 (keeps the compiler from splitting the fetch)
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

// Shade + composite one sample, STRICT arithmetic (volume.frag.glsl:96-115).  `k` holds the 8 corner texels.
__device__ __forceinline__ void shade_strict(const MarchArgs &a, const float4 *s_lut, const Corner8 &k,
                                             float wx, float wy, float wz, Accum &acc) {
    const float density = PYVR_TRILERP(x);
    // texture(transfer_function_lut, vec2(density, 0.5)): linear, clamp-to-edge, row axis degenerate
    const Taps tl = axis_taps(density, a.lut_size);
    const float4 l0 = lut_entry(s_lut, tl.i0)[0], l1 = lut_entry(s_lut, tl.i1)[0];
    const float alpha_tf = lerpf(l0.w, l1.w, tl.f);
    const float alpha = 1.0f - expf(-alpha_tf * a.step / a.ref_step);
    const float cr = lerpf(l0.x, l1.x, tl.f), cg = lerpf(l0.y, l1.y, tl.f), cb = lerpf(l0.z, l1.z, tl.f);
    float nx = PYVR_TRILERP(y), ny = PYVR_TRILERP(z), nz = PYVR_TRILERP(w);
    const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
    nx *= inv; ny *= inv; nz *= inv;
    const float ndotl = nx * a.ldir[0] + ny * a.ldir[1] + nz * a.ldir[2];
    // fmaxf(NaN, 0) = 0: a zero-length normal gives the ambient term only (oracle header).
    const float diff = fmaxf(ndotl, 0.0f);
    const float light = a.ambient + a.diffuse * diff;
    const float t = 1.0f - acc.a;
    acc.r = fmaf(t, cr * light * alpha, acc.r);
    acc.g = fmaf(t, cg * light * alpha, acc.g);
    acc.b = fmaf(t, cb * light * alpha, acc.b);
    acc.a = fmaf(t, alpha, acc.a);
}

// LUT fetch + shading + front-to-back compositing of the fast paths.  density and the (not yet normalised)
// normal are already filtered.  Returns without touching acc when alpha_tf == 0: such a sample contributes
// exactly +0 to every accumulator.
__device__ __forceinline__ bool shade_fast(const MarchArgs &a, const float4 *s_lut, float density, float nx, float ny,
                                           float nz, Accum &acc) {
    const float x = density * (float)a.lut_size - 0.5f;     // the oracle's expression (cell_classify uses the same)
    const int i = __float2int_rd(x);
    const float f = x - (float)i;
    const float4 *e = lut_entry(s_lut, min(max(i, -1), a.lut_size - 1));
    const float4 l0 = lds128(e), l1 = lds128(e + 1);
    const float alpha_tf = lerpf(l0.w, l1.w, f);
    if (alpha_tf == 0.0f) return false;
    const float alpha = 1.0f - ex2_approx(alpha_tf * a.exp2_scale);
    const f32x2 f2 = pack2(f, f);
    float cr, cg;
    unpack2(lerp2(pack2(l0.x, l0.y), pack2(l1.x, l1.y), f2), cr, cg);
    const float cb = lerpf(l0.z, l1.z, f);
    const float inv = rsqrt_approx(fmaf(nz, nz, fmaf(ny, ny, nx * nx)));
    const float ndotl = fmaf(nz, a.ldir[2], fmaf(ny, a.ldir[1], nx * a.ldir[0])) * inv;
    // fmaxf(NaN, 0) = 0: a zero-length normal (0 * inf) gives the ambient term only
    const float light = fmaf(a.diffuse, fmaxf(ndotl, 0.0f), a.ambient);
    const float t = 1.0f - acc.a;
    acc.r = fmaf(t, cr * light * alpha, acc.r);
    acc.g = fmaf(t, cg * light * alpha, acc.g);
    acc.b = fmaf(t, cb * light * alpha, acc.b);
    acc.a = fmaf(t, alpha, acc.a);
    return true;
}

// Opacity of a sample from its density alone (the same expressions as shade_fast): two 4-byte LUT reads.
__device__ __forceinline__ float lut_alpha(const MarchArgs &a, const float4 *s_lut, float density) {
    const float x = density * (float)a.lut_size - 0.5f;
    const int i = __float2int_rd(x);
    const float f = x - (float)i;
    const float4 *e = lut_entry(s_lut, min(max(i, -1), a.lut_size - 1));
    return lerpf(e[0].w, e[1].w, f);
}

// The scalar of the texel at p (binary32, or binary16 widened): 4 (2) bytes instead of the 16 (8) of the texel.
template <bool HALF>
__device__ __forceinline__ float load_scalar_at(const char *p) {
    if constexpr (HALF) return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short *>(p))));
    else return __ldg(reinterpret_cast<const float *>(p));
}


// Index interval on which lo <= X0 + i*D <= hi, intersected into [enter, exit].
__device__ __forceinline__ void index_slab(float X0, float D, float lo, float hi, float &enter, float &exit) {
    if (D != 0.0f) {
        const float r = 1.0f / D, a = (lo - X0) * r, b = (hi - X0) * r;
        enter = fmaxf(enter, fminf(a, b));
        exit = fminf(exit, fmaxf(a, b));
    } else if (X0 < lo || X0 > hi) {
        enter = 3.0e38f;
        exit = -3.0e38f;
    }
}

#ifndef PYVR_PERSISTENT
#define PYVR_PERSISTENT 0   // 1: one wave of CTAs, warps pull tiles off an atomic queue (not with image-space sharding)
#endif
#ifndef PYVR_LANE_ARR
#define PYVR_LANE_ARR 4     // measured on C3 over the WHOLE turntable (profiles/r02_turntable_ab.txt), each with its best pitch
                            // residues: rows (0) 447, 4x2 quarter blocks (1) 459, 2x2 pass blocks (3 / 4) 471 / 473 Gsamples/s
#endif
#ifndef PYVR_DENSITY_FIRST
#define PYVR_DENSITY_FIRST 0   // measured (profiles/r02_density_first_ab.txt): 452 vs 482 Gsamples/s with ESS, 122 vs 130 dense
#endif
#ifndef PYVR_PF_DIST
#define PYVR_PF_DIST 0      // samples ahead to prefetch (0 = off); PYVR_PF_LEVEL 1 = L1, 2 = L2
#endif
#ifndef PYVR_PF_LEVEL
#define PYVR_PF_LEVEL 2
#endif
// LAYOUT 3 / 4 (f16x4, sparse rays): samples of a run in flight per lane, and the occupancy that leaves room for their
// raw texels (16 registers per sample).  Measured on C4 with 2x2x2 bricks (profiles/r02_multi_sample_ab.txt): one
// sample 9.4 ms per frame; K = 2 at 5 CTAs/SM 5.35; K = 3 at 4 CTAs/SM 4.38; K = 4 at 3 CTAs/SM 3.69; K = 5, 6, 8 at
// 3 or 2 CTAs/SM 3.69-3.72 (DRAM: 18.9 GB per frame).  Spilling variants lose everything (K = 2 at 7-9 CTAs/SM: 9.5-12.7).
#ifndef PYVR_TWO_K
#define PYVR_TWO_K 4
#endif
#ifndef PYVR_MARCH_MIN_BLOCKS_TWO
#define PYVR_MARCH_MIN_BLOCKS_TWO 3   // up to 168 registers; K = 4 uses 148, no spills
#endif
#ifndef PYVR_B8_PF_DIST
#define PYVR_B8_PF_DIST 0   // brick8 layout: prefetch the bricks of the sample this many steps ahead (0 = off)
#endif
__device__ __forceinline__ void prefetch_line(const char *p) {
#if PYVR_PF_LEVEL == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

#ifndef PYVR_MAX_IV
#define PYVR_MAX_IV 6
#endif
constexpr int MAX_IV = PYVR_MAX_IV;   // active-interval table entries per ray (shared memory); refilled when exhausted

#ifndef PYVR_MARCH_MIN_BLOCKS
#define PYVR_MARCH_MIN_BLOCKS 7   // 7 CTAs x 4 warps per SM <=> at most 72 registers per thread
#endif

#ifndef PYVR_MARCH_MIN_BLOCKS_F16
#define PYVR_MARCH_MIN_BLOCKS_F16 9   // f16x4 corners need half the registers: 56 per thread, 36 warps per SM (+5 %)
#endif
#ifndef PYVR_MARCH_MIN_BLOCKS_TEX
#define PYVR_MARCH_MIN_BLOCKS_TEX 12   // the texture-unit variant holds no corner texels and is latency-bound: 40 registers,
                                       // 48 warps per SM (measured: 7 -> 699, 10 -> 890, 12 -> 939, 14 -> 583 Gsamples/s on C3)
#endif

// LAYOUT: 0 = rows, 1 = rows of z-pair entries, 2 = 2x2x2-texel bricks (common.cuh), 3 = z-pair entries marched two
// samples at a time (f16x4 only; MarchArgs.two_samples)
template <bool STRICT, bool HALF, typename IDX, bool BRICK, int LAYOUT, bool TEX>
__global__ void __launch_bounds__(CTA_THREADS, TEX ? PYVR_MARCH_MIN_BLOCKS_TEX : LAYOUT >= 3 ? PYVR_MARCH_MIN_BLOCKS_TWO
                                                   : (HALF && !STRICT) ? PYVR_MARCH_MIN_BLOCKS_F16 : PYVR_MARCH_MIN_BLOCKS)
march_kernel(const __grid_constant__ MarchArgs a) {
    constexpr bool PAIR = LAYOUT == 1 || LAYOUT == 3;
    // Several samples per iteration (f16x4, the format of the 2048^3 / 4096^3 configs): with sparse rays every warp
    // sits on its loads for a DRAM round trip (ncu on C4: 52 warp-cycles of long-scoreboard stall per issued
    // instruction, issue slots 14 % busy), and the longest ray alone bounds a sharded launch.  Requesting the texels of
    // samples i .. i+K-1 of a run together multiplies the loads in flight per lane and divides the round trips of a
    // ray; the samples are still composited in order, and those after the one that saturates the ray are dropped, so
    // the arithmetic -- and every pixel -- is unchanged (test_two_samples_per_iteration_is_bit_identical).
    // Chosen per launch (launch_march): on C3, where the L1 data stage binds, the extra registers cost more than the
    // round trips save.
    constexpr bool TWO = LAYOUT == 3;
    constexpr bool TWO_B8 = LAYOUT == 4;       // the same for the 2x2x2-brick layout
    static_assert(!(TWO || TWO_B8) || (HALF && !STRICT && !TEX), "multi-sample march: f16x4 fast path only");
    constexpr int TEXEL_BYTES = HALF ? 8 : 16, ENTRY_BYTES = TEXEL_BYTES << (PAIR ? 1 : 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // lane -> pixel inside the warp's 8x4 tile.  The L1 data stage serves a warp-wide LDG.256 in passes of four
    // consecutive lanes (tools/l1_gather_probe.cu); how many distinct entries those lanes touch, and in which slots of
    // their lines, decides how many cycles a pass takes (tools/bank_sim.py): a 2x2 pixel block per pass conflicts
    // least.  The first A/Bs of round 2 marched views 0..15 only and preferred 4x1 passes (profiles/r02_lane_ab.txt);
    // over all 360 views the 2x2 blocks win by 3 %.
#if PYVR_LANE_ARR == 1      // quarter-warp = 4x2 pixel block
    const int lane_x = (lane & 3) + 4 * ((lane >> 3) & 1), lane_y = ((lane >> 2) & 1) + 2 * (lane >> 4);
#elif PYVR_LANE_ARR == 2    // quarter-warp = 2x4 pixel block
    const int lane_x = (lane & 1) + 2 * (lane >> 3), lane_y = (lane >> 1) & 3;
#elif PYVR_LANE_ARR == 3    // 4 consecutive lanes (one LDG.256 data-stage pass) = 2x2 pixels, quarter-warp = 4x2
    const int lane_x = (lane & 1) + 2 * ((lane >> 2) & 1) + 4 * ((lane >> 3) & 1), lane_y = ((lane >> 1) & 1) + 2 * (lane >> 4);
#elif PYVR_LANE_ARR == 4    // 4 consecutive lanes = 2x2 pixels, quarter-warp = 2x4
    const int lane_x = (lane & 1) + 2 * (lane >> 3), lane_y = ((lane >> 1) & 1) + 2 * ((lane >> 2) & 1);
#else                       // quarter-warp = one 8-pixel row
    const int lane_x = lane % WARP_W, lane_y = lane / WARP_W;
#endif
    extern __shared__ float4 s_lut[];
    stage_lut(s_lut, a.lut, a.lut_size, CTA_THREADS);
    __syncthreads();
    const VolumeDesc &vol = a.vol;
    __shared__ int s_iv[2 * MAX_IV][CTA_THREADS];   // per-ray table of active index intervals, see the pre-walk below

#if PYVR_PERSISTENT
    // Persistent warps with an atomic tile queue: the grid is one wave of CTAs, every
    // WARP pulls 8x4-pixel tiles off a global counter until none is left.  Consecutive tickets are the four warp
    // tiles of one 16x8 CTA tile, then the next CTA tile along the image row, so warps that run together still
    // share lines.  No warp waits for the slowest warp of its CTA and the LUT is staged once per SM slot.
    const unsigned tiles_cx = (unsigned)((a.width + TILE_W - 1) / TILE_W), tiles_cy = (unsigned)((a.height + TILE_H - 1) / TILE_H);
    const unsigned total_tickets = tiles_cx * tiles_cy * (unsigned)(CTA_WX * CTA_WY) * (unsigned)a.n_views;
    for (;;) {
    unsigned ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.tile_counter, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= total_tickets) break;
    const int wsub = (int)(ticket % (unsigned)(CTA_WX * CTA_WY));
    const unsigned cta_tile = ticket / (unsigned)(CTA_WX * CTA_WY);
    const int tile_x = (int)(cta_tile % tiles_cx), tile_y = (int)((cta_tile / tiles_cx) % tiles_cy);
    const int view_index = (int)(cta_tile / (tiles_cx * tiles_cy));
#else
    // image-space sharding (multi-GPU tiles): groups of 2^shift x 2^shift CTA tiles are dealt over the ranks along
    // image rows, every row of groups shifted by one rank against the one below: owner(gx, gy) = (gx + gy) mod P.
    // Only the owned tiles are launched: blockIdx.x enumerates this rank's groups of the row.
    // Row order: CTAs start in block-index order.  Consecutive blockIdx.y are mapped to tile rows from first_row (where
    // the volume's centre projects, abi.cu) outwards, alternating below / above, so the rows with the longest rays
    // start first and the launch ends on rows that miss the volume (a single 16x8 tile through the middle of C4 runs
    // for about a millisecond): 8.16 -> 7.83 ms per C4 frame on one GPU, nothing on C3 (profiles/r02_c4_shard_probe.txt).
    int block_y = blockIdx.y;
    if (a.first_row >= 0) {
        const int rows = gridDim.y, cy = a.first_row, m = min(cy, rows - 1 - cy), k = blockIdx.y;
        block_y = k <= 2 * m ? cy + ((k & 1) ? (k + 1) / 2 : -(k / 2)) : (cy <= rows - 1 - cy ? k : rows - 1 - k);
    }
    int tile_x = blockIdx.x;
    if (a.shard_count > 1) {
        const int sh = a.shard_shift, gy = block_y >> sh, k = blockIdx.x >> sh;
        const int first = (a.shard_rank - gy % a.shard_count + a.shard_count) % a.shard_count;
        tile_x = ((first + k * a.shard_count) << sh) + (blockIdx.x & ((1 << sh) - 1));
        if (tile_x * TILE_W >= a.width) return;
    }
    const int tile_y = block_y, view_index = blockIdx.z, wsub = warp;
    {
#endif
    const int px = tile_x * TILE_W + (wsub % CTA_WX) * WARP_W + lane_x;
    const int py = tile_y * TILE_H + (wsub / CTA_WX) * WARP_H + lane_y;
    const bool in_image = px < a.width && py < a.height;
    const pyvr_view &vw = a.views[view_index];

    Accum acc = {0.0f, 0.0f, 0.0f, 0.0f};
    unsigned n_samples = 0, n_fetched = 0;
    bool hit = false, terminated = false;

    // State of the fast march.  It lives outside the set-up conditionals because the fast loop is
    // warp-synchronous: all 32 lanes stay in it (idle lanes predicated off) until the whole warp is done.
    float X0 = 0.0f, Y0 = 0.0f, Z0 = 0.0f, DX = 0.0f, DY = 0.0f, DZ = 0.0f;
    int i = 0, i_lo = 0, j_hi = -1, last = -1;
    bool alive = false;

    if (in_image && vol.texels != nullptr && a.lut_size > 0) {
        // ---- ray set-up: the oracle's arithmetic in BOTH modes.  The first sample sits exactly on the
        // box surface, so whether it passes the validity test is decided by the last bit of these
        // operations; any shortcut here would flip that coin for volumes that are opaque at the faces.
        // volume.vert.glsl:8 at the pixel centre, volume.frag.glsl:33-35,47-54
        const float ndx = ((float)px + 0.5f) / (float)a.width * 2.0f - 1.0f;
        const float ndy = ((float)py + 0.5f) / (float)a.height * 2.0f - 1.0f;
        float dx, dy, dz;
        if (vw.has_matrices) {
            // eye = inverse(P) * (ndc, -1, 1); eye.zw = (-1, 0); world = inverse(V) * eye
            const float *ip = vw.inv_proj, *iv = vw.inv_view;
            const float ex = ip[0] * ndx + ip[4] * ndy + ip[8] * -1.0f + ip[12] * 1.0f;
            const float ey = ip[1] * ndx + ip[5] * ndy + ip[9] * -1.0f + ip[13] * 1.0f;
            dx = iv[0] * ex + iv[4] * ey + iv[8] * -1.0f + iv[12] * 0.0f;
            dy = iv[1] * ex + iv[5] * ey + iv[9] * -1.0f + iv[13] * 0.0f;
            dz = iv[2] * ex + iv[6] * ey + iv[10] * -1.0f + iv[14] * 0.0f;
        } else {
            dx = fmaf(vw.u[0], ndx, fmaf(vw.v[0], ndy, vw.w[0]));
            dy = fmaf(vw.u[1], ndx, fmaf(vw.v[1], ndy, vw.w[1]));
            dz = fmaf(vw.u[2], ndx, fmaf(vw.v[2], ndy, vw.w[2]));
        }
        const float dinv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        dx *= dinv; dy *= dinv; dz *= dinv;
        const float ox = vw.origin[0], oy = vw.origin[1], oz = vw.origin[2];

        // intersect_box, volume.frag.glsl:56-68 (IEEE divide: 1/0 = inf is relied on)
        const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
        const float ax = (vol.bmin[0] - ox) * ix, bx = (vol.bmax[0] - ox) * ix;
        const float ay = (vol.bmin[1] - oy) * iy, by = (vol.bmax[1] - oy) * iy;
        const float az = (vol.bmin[2] - oz) * iz, bz = (vol.bmax[2] - oz) * iz;
        float t_near = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        const float t_far = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        hit = t_near <= t_far && t_far > 0.0f;

        if (hit) {
            t_near = fmaxf(t_near, 0.0f);
            // The shader runs to max_steps; everything past t_far + 2 steps is outside the box
            // (the validity test below still decides every sample), so stop there.
            const float span = (t_far - t_near) / a.step;
            const int n_steps = (int)fminf((float)a.max_steps, ceilf(span) + 2.0f);
            const float p0x = ox + dx * t_near, p0y = oy + dy * t_near, p0z = oz + dz * t_near;
            const float sx = dx * a.step, sy = dy * a.step, sz = dz * a.step;

            if (STRICT) {
                float pxw = p0x, pyw = p0y, pzw = p0z;
                int k = 0;
                for (; k < n_steps && acc.a < a.term_alpha; ++k) {
                    const float tcx = (pxw - vol.bmin[0]) / (vol.bmax[0] - vol.bmin[0]);
                    const float tcy = (pyw - vol.bmin[1]) / (vol.bmax[1] - vol.bmin[1]);
                    const float tcz = (pzw - vol.bmin[2]) / (vol.bmax[2] - vol.bmin[2]);
                    if (tcx >= 0.0f && tcx <= 1.0f && tcy >= 0.0f && tcy <= 1.0f && tcz >= 0.0f && tcz <= 1.0f) {
                        ++n_samples; ++n_fetched;
                        const Taps tx = axis_taps(tcx, vol.gn[0]), ty = axis_taps(tcy, vol.gn[1]),
                                   tz = axis_taps(tcz, vol.gn[2]);
                        const Corner8 c8 = gather_strict<HALF>(vol, tx, ty, tz);
                        shade_strict(a, s_lut, c8, tx.f, ty.f, tz.f, acc);
                    }
                    pxw += sx; pyw += sy; pzw += sz;
                }
                terminated = (k < n_steps) || (n_steps < a.max_steps && acc.a >= a.term_alpha);
            } else {
                // ---- voxel-space lattice x(i) = X0 + i*DX; sample i is valid iff -0.5 <= x <= n-0.5 on
                // every axis, except sample 0 (on the box surface), which takes the reference's own test.
                const float t0x = (p0x - vol.bmin[0]) / (vol.bmax[0] - vol.bmin[0]);
                const float t0y = (p0y - vol.bmin[1]) / (vol.bmax[1] - vol.bmin[1]);
                const float t0z = (p0z - vol.bmin[2]) / (vol.bmax[2] - vol.bmin[2]);
                const bool valid0 = t0x >= 0.0f && t0x <= 1.0f && t0y >= 0.0f && t0y <= 1.0f && t0z >= 0.0f && t0z <= 1.0f;
                const float hx = (float)vol.gn[0] - 0.5f, hy = (float)vol.gn[1] - 0.5f, hz = (float)vol.gn[2] - 0.5f;
                // The entry point lies on the box surface up to round-off (~1e-5 voxel): pull it inside so
                // that the taps of sample 0 stay in the array.
                X0 = fminf(fmaxf(fmaf(p0x, vol.vscale[0], vol.voff[0]), -0.5f), hx); DX = sx * vol.vscale[0];
                Y0 = fminf(fmaxf(fmaf(p0y, vol.vscale[1], vol.voff[1]), -0.5f), hy); DY = sy * vol.vscale[1];
                Z0 = fminf(fmaxf(fmaf(p0z, vol.vscale[2], vol.voff[2]), -0.5f), hz); DZ = sz * vol.vscale[2];
                auto valid = [&](int k) -> bool {
                    if (k <= 0) return k == 0 && valid0;
                    if (k >= n_steps) return false;
                    const float fk = (float)k;
                    const float x = fmaf(fk, DX, X0), y = fmaf(fk, DY, Y0), z = fmaf(fk, DZ, Z0);
                    return x >= -0.5f && x <= hx && y >= -0.5f && y <= hy && z >= -0.5f && z <= hz;
                };

                // in-volume index range [i_lo, i_hi]: slab estimate, then the predicate is the arbiter
                float enter = -3.0e38f, exit = 3.0e38f;
                index_slab(X0, DX, -0.5f, hx, enter, exit);
                index_slab(Y0, DY, -0.5f, hy, enter, exit);
                index_slab(Z0, DZ, -0.5f, hz, enter, exit);
                i_lo = (int)fminf(fmaxf(ceilf(enter), 0.0f), (float)n_steps);
                int i_hi = (int)fminf(fmaxf(floorf(exit), -1.0f), (float)(n_steps - 1));
                if (valid(i_lo - 1)) --i_lo; else if (!valid(i_lo)) ++i_lo;
                if (valid(i_hi + 1)) ++i_hi; else if (!valid(i_hi)) --i_hi;
                if (valid0) i_lo = 0;
                else if (i_lo == 0) i_lo = 1;

                if (BRICK) {
                    // Sort-last brick: keep the samples this brick owns.  Ownership is a convex box in voxel
                    // space, so the owned indices form one interval; the slab estimate is corrected with the
                    // ownership predicate itself, evaluated on the very x(k) every brick computes, so each
                    // sample of the ray ends up in exactly one brick.
                    auto owned = [&](int k) -> bool {
                        const float fk = (float)k;
                        const float x = fmaf(fk, DX, X0), y = fmaf(fk, DY, Y0), z = fmaf(fk, DZ, Z0);
                        return x >= vol.own_lo[0] && x < vol.own_hi[0] && y >= vol.own_lo[1] && y < vol.own_hi[1] &&
                               z >= vol.own_lo[2] && z < vol.own_hi[2];
                    };
                    float en = -3.0e38f, exi = 3.0e38f;
                    index_slab(X0, DX, vol.own_lo[0], vol.own_hi[0], en, exi);
                    index_slab(Y0, DY, vol.own_lo[1], vol.own_hi[1], en, exi);
                    index_slab(Z0, DZ, vol.own_lo[2], vol.own_hi[2], en, exi);
                    int o_lo = max(i_lo, (int)fminf(fmaxf(ceilf(en) - 2.0f, 0.0f), 2.0e9f));
                    int o_hi = min(i_hi, (int)fminf(fmaxf(floorf(exi) + 2.0f, -1.0f), 2.0e9f));
                    for (int g = 0; g < 6 && o_lo <= o_hi && !owned(o_lo); ++g) ++o_lo;
                    for (int g = 0; g < 6 && o_lo <= o_hi && !owned(o_hi); ++g) --o_hi;
                    if (o_lo <= o_hi && !(owned(o_lo) && owned(o_hi))) o_hi = o_lo - 1;   // grazing ray
                    i_lo = o_lo;
                    i_hi = o_hi;
                }

                // clip to the bounding box of the active macrocells (samples outside add exactly zero)
                int j_lo = i_lo;
                j_hi = i_hi;
                if ((a.flags & PYVR_FLAG_ESS) && vol.cell_dist != nullptr) {
                    const int *ab = vol.active_box;
                    const int cx0 = __ldg(ab + 0), cy0 = __ldg(ab + 1), cz0 = __ldg(ab + 2);
                    const int cx1 = __ldg(ab + 3), cy1 = __ldg(ab + 4), cz1 = __ldg(ab + 5);
                    if (cx0 > cx1) {
                        j_hi = j_lo - 1;
                    } else {
                        float en = -3.0e38f, exi = 3.0e38f;
                        const int fx0 = vol.org[0] + kCell * cx0, fy0 = vol.org[1] + kCell * cy0, fz0 = vol.org[2] + kCell * cz0;
                        index_slab(X0, DX, fx0 == 0 ? -0.5f : (float)fx0, fminf((float)(vol.org[0] + kCell * cx1 + kCell), hx), en, exi);
                        index_slab(Y0, DY, fy0 == 0 ? -0.5f : (float)fy0, fminf((float)(vol.org[1] + kCell * cy1 + kCell), hy), en, exi);
                        index_slab(Z0, DZ, fz0 == 0 ? -0.5f : (float)fz0, fminf((float)(vol.org[2] + kCell * cz1 + kCell), hz), en, exi);
                        j_lo = max(j_lo, (int)fminf(fmaxf(floorf(en) - 1.0f, 0.0f), (float)n_steps));
                        j_hi = min(j_hi, (int)fminf(fmaxf(ceilf(exi) + 1.0f, -1.0f), (float)n_steps));
                    }
                }
                i = j_lo;
                last = i_hi;        // index of the last sample the reference executes
                alive = j_lo <= j_hi;
            }
        }
    }

    if constexpr (!STRICT) {
        // relay: continue the accumulation of the bricks in front.  A ray that arrives saturated executes no
        // sample here, exactly as the single pass would not (volume.frag.glsl:87 tests before each sample).
        if (a.in_acc != nullptr && in_image) {
            const float4 in = a.in_acc[((size_t)view_index * a.height + py) * a.width + px];
            acc.r = in.x; acc.g = in.y; acc.b = in.z; acc.a = in.w;
            if (acc.a >= a.term_alpha) { alive = false; last = i_lo - 1; }
        }
        // ---- fast march.  The ray first WALKS the cell map from j_lo to j_hi (cubes of like cells, see
        // volume_pack.cu) and records the index intervals that lie in active cells -- adjacent active cubes
        // merge, so a compact object is one interval -- in a small per-thread table in shared memory.  All 32
        // lanes walk at once, each its own ray; doing the lookups inside the sampling loop instead made every
        // round wait for the few lanes (8 of 32 on C3) that had just run off their cube.  The sampling loop is
        // then warp-synchronous: a vote separates "advance to the next interval" from "gather + shade" and
        // reconverges the warp (without it the compiler ran the gather as two half-empty groups).
        const bool ess = (a.flags & PYVR_FLAG_ESS) && vol.cell_dist != nullptr;
        const float hx = (float)vol.gn[0] - 0.5f, hy = (float)vol.gn[1] - 0.5f, hz = (float)vol.gn[2] - 0.5f;
        const int ogx = BRICK ? vol.org[0] : 0, ogy = BRICK ? vol.org[1] : 0, ogz = BRICK ? vol.org[2] : 0;
        const int max_last = a.max_steps - 1;
        const IDX PX8 = (IDX)(vol.pitch_x * 8), PY8 = (IDX)vol.pitch_y * 8;   // brick8: texels per x-plane / z-row of bricks
        (void)PX8; (void)PY8;
        // s_iv[2k] = first index, s_iv[2k+1] = last index of interval k
        int n_iv = 0, iv_next = 0;
        int walk_i = i;            // where the walk resumes when the table has been consumed
        bool walked = true;        // the walk has reached j_hi

        // Walk from walk_i and (re)fill the table.
        auto fill_table = [&]() {
            // Per-ray constants of the walk.  Along an axis the ray leaves a cube of cells [c - r, c + r]
            // through the face at voxel 8*(c + r + 1) (moving up) or 8*(c - r) (moving down):
            //   face = 8*c + r*fstep + fbase,  fstep = +-8,  fbase = 8 or 0 (plus the brick origin);
            // a zero direction component counts as "up" with an infinite reciprocal, i.e. it never exits.
            const float rDX = DX != 0.0f ? 1.0f / DX : 3.0e38f, rDY = DY != 0.0f ? 1.0f / DY : 3.0e38f,
                        rDZ = DZ != 0.0f ? 1.0f / DZ : 3.0e38f;
            const int fstep_x = DX < 0.0f ? -kCell : kCell, fstep_y = DY < 0.0f ? -kCell : kCell, fstep_z = DZ < 0.0f ? -kCell : kCell;
            const int fbase_x = (DX < 0.0f ? 0 : kCell) + ogx, fbase_y = (DY < 0.0f ? 0 : kCell) + ogy, fbase_z = (DZ < 0.0f ? 0 : kCell) + ogz;
            n_iv = 0; iv_next = 0; walked = true;
            int cs = -1, ce = -1;          // the interval being built
            int k = walk_i;
            while (k <= j_hi) {
                // Cell byte: b < 128: inactive, every cell within chessboard radius b-1 is inactive too;
                // b >= 128: active, every cell within radius b-128 is active.  Either way the ray may run to
                // the faces of that cube of cells: whole steps that stay inside it (and inside the volume) on
                // every axis, conservative by 0.01 step.
                const float fk = (float)k;
                const float x = fmaf(fk, DX, X0), y = fmaf(fk, DY, Y0), z = fmaf(fk, DZ, Z0);
                const int lx = max(__float2int_rd(x), 0) - ogx, ly = max(__float2int_rd(y), 0) - ogy,
                          lz = max(__float2int_rd(z), 0) - ogz;     // lower tap, local to the stored block
                const int b = __ldg(vol.cell_dist + ((lx >> kCellShift) * vol.ncell[1] + (ly >> kCellShift)) * vol.ncell[2] + (lz >> kCellShift));
                const bool active = b >= 128;
                const int r = (b & 127) - (active ? 0 : 1);
                const float fx = fminf(fmaxf((float)((lx & ~(kCell - 1)) + r * fstep_x + fbase_x), -0.5f), hx);
                const float fy = fminf(fmaxf((float)((ly & ~(kCell - 1)) + r * fstep_y + fbase_y), -0.5f), hy);
                const float fz = fminf(fmaxf((float)((lz & ~(kCell - 1)) + r * fstep_z + fbase_z), -0.5f), hz);
                const float tmin = fminf(fminf((fx - x) * rDX, (fy - y) * rDY), (fz - z) * rDZ);
                const int stay = min(max(__float2int_rd(tmin - 0.01f), 1), 1 << 24);
                if (active) {
                    if (cs >= 0 && k > ce + 1) {         // a gap: close the interval being built
                        if (n_iv == MAX_IV) { walked = false; break; }   // table full: resume here later
                        s_iv[2 * n_iv][threadIdx.x] = cs; s_iv[2 * n_iv + 1][threadIdx.x] = ce; ++n_iv;
                        cs = -1;
                    }
                    if (cs < 0) cs = k;
                    ce = min(k + stay - 1, j_hi);
                }
                k += stay;
            }
            if (cs >= 0) {
                if (walked && n_iv < MAX_IV) {
                    s_iv[2 * n_iv][threadIdx.x] = cs; s_iv[2 * n_iv + 1][threadIdx.x] = ce; ++n_iv;
                    k = j_hi + 1;
                } else if (walked) {      // finished the walk with one interval too many
                    walked = false; k = cs;
                } else {
                    k = cs;               // stopped at a gap: the open interval is re-walked on the refill
                }
            }
            walk_i = k;
        };

        if (alive) {
            if (ess) {
                fill_table();
            } else {                      // no skipping: one interval, the whole in-volume range
                s_iv[0][threadIdx.x] = i; s_iv[1][threadIdx.x] = j_hi; n_iv = 1;
            }
        }
        int run_end = i;                  // first index past the current interval
        bool warp_transparent = !ess;     // no lane of the warp saw a visible sample in the previous round
        while (true) {
            bool have = false;
            if (alive) {
                if (i < run_end) {
                    have = true;
                } else {
                    if (iv_next == n_iv && !walked) fill_table();      // rare: more than MAX_IV intervals
                    if (iv_next < n_iv) {
                        i = s_iv[2 * iv_next][threadIdx.x];
                        run_end = s_iv[2 * iv_next + 1][threadIdx.x] + 1;
                        ++iv_next;
                        have = true;
                    } else {
                        alive = false;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, have)) break;
            bool visible = false;
            if (have) {
                ++n_fetched;
                const float fi = (float)i;
                const float x = fmaf(fi, DX, X0), y = fmaf(fi, DY, Y0), z = fmaf(fi, DZ, Z0);
                if constexpr (TEX) {
                    // texel centres sit at integer + 0.5 in unnormalised texture space; width = z
                    const float4 t = tex3D<float4>(a.tex, z + 0.5f - (float)ogz, y + 0.5f - (float)ogy, x + 0.5f - (float)ogx);
                    visible = shade_fast(a, s_lut, t.x, t.y, t.z, t.w, acc);
                } else {
                    // lower taps floor(x) in [-1, n-1] (the apron holds the clamped texels), upper taps = lower + 1
                    const int ix = __float2int_rd(x), iy = __float2int_rd(y), iz = __float2int_rd(z);
                    const float wx = x - (float)ix, wy = y - (float)iy, wz = z - (float)iz;
                    Texel2 c000, c001, c010, c011, c100, c101, c110, c111;
                    bool fetch = true;
                    if constexpr (TWO_B8) {
                        // 2x2x2-texel bricks, K samples of the run in flight: eight single-texel loads per sample, raw
                        constexpr int K = PYVR_TWO_K;
                        const int cnt = min(K, run_end - i);
                        uint2 raw[K][8];
                        float wgt[K][3];
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            if (k < cnt) {
                                const float fk = fi + (float)k;
                                const float xk = fmaf(fk, DX, X0), yk = fmaf(fk, DY, Y0), zk = fmaf(fk, DZ, Z0);
                                const int jx = __float2int_rd(xk), jy = __float2int_rd(yk), jz = __float2int_rd(zk);
                                wgt[k][0] = xk - (float)jx; wgt[k][1] = yk - (float)jy; wgt[k][2] = zk - (float)jz;
                                const int X = jx - ogx + 1, Y = jy - ogy + 1, Z = jz - ogz + 1;
                                const IDX ex0 = (IDX)(X >> 1) * PX8 + (IDX)((X & 1) << 2), ex1 = (IDX)((X + 1) >> 1) * PX8 + (IDX)(((X + 1) & 1) << 2);
                                const IDX ey0 = (IDX)(Y >> 1) * PY8 + (IDX)((Y & 1) << 1), ey1 = (IDX)((Y + 1) >> 1) * PY8 + (IDX)(((Y + 1) & 1) << 1);
                                const IDX ez0 = (IDX)((Z >> 1) << 3) + (IDX)(Z & 1), ez1 = (IDX)(((Z + 1) >> 1) << 3) + (IDX)((Z + 1) & 1);
                                const IDX e4[4] = {ex0 + ey0, ex0 + ey1, ex1 + ey0, ex1 + ey1};
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    raw[k][2 * c] = __ldg(reinterpret_cast<const uint2 *>(a.tap_base + (long long)(e4[c] + ez0) * TEXEL_BYTES));
                                    raw[k][2 * c + 1] = __ldg(reinterpret_cast<const uint2 *>(a.tap_base + (long long)(e4[c] + ez1) * TEXEL_BYTES));
                                }
                            }
                        }
                        auto filter_shade = [&](const uint2 (&q)[8], const float (&f)[3]) -> bool {
                            const f32x2 tz2 = pack2(f[2], f[2]), ty2 = pack2(f[1], f[1]), tx2 = pack2(f[0], f[0]);
                            f32x2 sn[4], yz[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {     // z-lerp of corner row c = (x tap, y tap)
                                sn[c] = lerp2(half2_to_f32x2(q[2 * c].x), half2_to_f32x2(q[2 * c + 1].x), tz2);
                                yz[c] = lerp2(half2_to_f32x2(q[2 * c].y), half2_to_f32x2(q[2 * c + 1].y), tz2);
                            }
                            float density, nx, ny, nz;
                            unpack2(lerp2(lerp2(sn[0], sn[1], ty2), lerp2(sn[2], sn[3], ty2), tx2), density, nx);
                            unpack2(lerp2(lerp2(yz[0], yz[1], ty2), lerp2(yz[2], yz[3], ty2), tx2), ny, nz);
                            return shade_fast(a, s_lut, density, nx, ny, nz, acc);
                        };
                        visible = filter_shade(raw[0], wgt[0]);
                        fetch = false;
#pragma unroll
                        for (int k = 1; k < K; ++k) {
                            if (k < cnt && acc.a < a.term_alpha) {      // volume.frag.glsl:87: alpha is tested before every sample
                                ++i;
                                ++n_fetched;
                                visible |= filter_shade(raw[k], wgt[k]);
                            }
                        }
                    } else if constexpr (LAYOUT == 2) {
                        // 2x2x2-texel bricks: per axis the brick term and the in-brick bit of the lower and the upper tap
                        // (coordinates shifted by the apron), then eight single-texel loads.
                        const int X = ix - ogx + 1, Y = iy - ogy + 1, Z = iz - ogz + 1;
                        const IDX ex0 = (IDX)(X >> 1) * PX8 + (IDX)((X & 1) << 2), ex1 = (IDX)((X + 1) >> 1) * PX8 + (IDX)(((X + 1) & 1) << 2);
                        const IDX ey0 = (IDX)(Y >> 1) * PY8 + (IDX)((Y & 1) << 1), ey1 = (IDX)((Y + 1) >> 1) * PY8 + (IDX)(((Y + 1) & 1) << 1);
                        const IDX ez0 = (IDX)((Z >> 1) << 3) + (IDX)(Z & 1), ez1 = (IDX)(((Z + 1) >> 1) << 3) + (IDX)((Z + 1) & 1);
                        const IDX e00 = ex0 + ey0, e01 = ex0 + ey1, e10 = ex1 + ey0, e11 = ex1 + ey1;
                        load_one<HALF>(a.tap_base + (long long)(e00 + ez0) * TEXEL_BYTES, c000);
                        load_one<HALF>(a.tap_base + (long long)(e00 + ez1) * TEXEL_BYTES, c001);
                        load_one<HALF>(a.tap_base + (long long)(e01 + ez0) * TEXEL_BYTES, c010);
                        load_one<HALF>(a.tap_base + (long long)(e01 + ez1) * TEXEL_BYTES, c011);
                        load_one<HALF>(a.tap_base + (long long)(e10 + ez0) * TEXEL_BYTES, c100);
                        load_one<HALF>(a.tap_base + (long long)(e10 + ez1) * TEXEL_BYTES, c101);
                        load_one<HALF>(a.tap_base + (long long)(e11 + ez0) * TEXEL_BYTES, c110);
                        load_one<HALF>(a.tap_base + (long long)(e11 + ez1) * TEXEL_BYTES, c111);
#if PYVR_B8_PF_DIST > 0
                        if (i + PYVR_B8_PF_DIST < run_end) {
                            // A/B builds only (measured: no help, abi.cu choose_layout): ask L2 now for the bricks of a later
                            // sample of this run.  Four taps of alternating parity reach every brick the sample touches
                            // unless it straddles brick faces on all three axes (1 in 8).
                            const float fp = fi + (float)PYVR_B8_PF_DIST;
                            const int QX = __float2int_rd(fmaf(fp, DX, X0)) - ogx + 1, QY = __float2int_rd(fmaf(fp, DY, Y0)) - ogy + 1,
                                      QZ = __float2int_rd(fmaf(fp, DZ, Z0)) - ogz + 1;
                            const IDX qx0 = (IDX)(QX >> 1) * PX8, qx1 = (IDX)((QX + 1) >> 1) * PX8;
                            const IDX qy0 = (IDX)(QY >> 1) * PY8, qy1 = (IDX)((QY + 1) >> 1) * PY8;
                            const IDX qz0 = (IDX)((QZ >> 1) << 3), qz1 = (IDX)(((QZ + 1) >> 1) << 3);
                            prefetch_line(a.tap_base + (long long)(qx0 + qy0 + qz0) * TEXEL_BYTES);
                            prefetch_line(a.tap_base + (long long)(qx1 + qy1 + qz0) * TEXEL_BYTES);
                            prefetch_line(a.tap_base + (long long)(qx1 + qy0 + qz1) * TEXEL_BYTES);
                            prefetch_line(a.tap_base + (long long)(qx0 + qy1 + qz1) * TEXEL_BYTES);
                        }
#endif
                    } else {
                    if constexpr (TWO) {
                        // the rows of samples i .. i + K - 1 of the run are requested now (raw: 16 registers per sample)
                        // and filtered one after the other
                        constexpr int K = PYVR_TWO_K;
                        const int cnt = min(K, run_end - i);
                        uint4 raw[K][4];
                        float wgt[K][3];
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            if (k < cnt) {
                                const float fk = fi + (float)k;
                                const float xk = fmaf(fk, DX, X0), yk = fmaf(fk, DY, Y0), zk = fmaf(fk, DZ, Z0);
                                const int jx = __float2int_rd(xk), jy = __float2int_rd(yk), jz = __float2int_rd(zk);
                                wgt[k][0] = xk - (float)jx; wgt[k][1] = yk - (float)jy; wgt[k][2] = zk - (float)jz;
                                const IDX ek = (IDX)(jx - ogx) * (IDX)vol.pitch_x + (IDX)(jy - ogy) * (IDX)vol.pitch_y + (IDX)(jz - ogz);
                                const char *q00 = a.tap_base + (long long)ek * ENTRY_BYTES;
                                raw[k][0] = load_row_raw(q00); raw[k][1] = load_row_raw(q00 + a.stride_y);
                                raw[k][2] = load_row_raw(q00 + a.stride_x); raw[k][3] = load_row_raw(q00 + a.stride_x + a.stride_y);
                            }
                        }
                        // z-lerp each corner row as it is unpacked (8 live values per row, not 32 per sample), then y, then x:
                        // the same operations in the same order as the one-sample path
                        auto filter_shade = [&](const uint4 (&q)[4], const float (&f)[3]) -> bool {
                            const f32x2 tz2 = pack2(f[2], f[2]), ty2 = pack2(f[1], f[1]), tx2 = pack2(f[0], f[0]);
                            Texel2 lo, hi;
                            unpack_row(q[0], lo, hi);
                            const f32x2 a00 = lerp2(lo.sn, hi.sn, tz2), b00 = lerp2(lo.yz, hi.yz, tz2);
                            unpack_row(q[1], lo, hi);
                            const f32x2 a01 = lerp2(lo.sn, hi.sn, tz2), b01 = lerp2(lo.yz, hi.yz, tz2);
                            unpack_row(q[2], lo, hi);
                            const f32x2 a10 = lerp2(lo.sn, hi.sn, tz2), b10 = lerp2(lo.yz, hi.yz, tz2);
                            unpack_row(q[3], lo, hi);
                            const f32x2 a11 = lerp2(lo.sn, hi.sn, tz2), b11 = lerp2(lo.yz, hi.yz, tz2);
                            float density, nx, ny, nz;
                            unpack2(lerp2(lerp2(a00, a01, ty2), lerp2(a10, a11, ty2), tx2), density, nx);
                            unpack2(lerp2(lerp2(b00, b01, ty2), lerp2(b10, b11, ty2), tx2), ny, nz);
                            return shade_fast(a, s_lut, density, nx, ny, nz, acc);
                        };
                        visible = filter_shade(raw[0], wgt[0]);
                        fetch = false;                         // all K samples are handled here
#pragma unroll
                        for (int k = 1; k < K; ++k) {
                            // the shader tests alpha before every sample (volume.frag.glsl:87): once the ray is saturated the
                            // samples already requested are dropped
                            if (k < cnt && acc.a < a.term_alpha) {
                                ++i;
                                ++n_fetched;
                                visible |= filter_shade(raw[k], wgt[k]);
                            }
                        }
                    } else {
                    const IDX e = (IDX)(ix - ogx) * (IDX)vol.pitch_x + (IDX)(iy - ogy) * (IDX)vol.pitch_y + (IDX)(iz - ogz);
                    const char *p00 = a.tap_base + (long long)e * ENTRY_BYTES;
                    const char *p01 = p00 + a.stride_y, *p10 = p00 + a.stride_x, *p11 = p10 + a.stride_y;
                    // Density first while the warp travels through transparent space (PYVR_DENSITY_FIRST, off): the 8
                    // scalars alone (32 of the 128 gather bytes) decide whether the sample is visible; only visible
                    // samples fetch their texels.  Same arithmetic on the same values as the full path, so the
                    // skipped samples are exactly the ones that would have added +0.  A NEGATIVE result on the B200:
                    // eight 4-byte gathers cost the L1 data stage about as much as the four 32-byte ones they
                    // replace (the stage is paid per quarter-warp pass, not per byte), so even the dense march,
                    // where 79 % of the samples are transparent, got 6 % slower.
#if PYVR_DENSITY_FIRST
                    if (warp_transparent) {
                        constexpr int ZOFF = HALF ? 8 : 16;    // texel(iz + 1): the next entry, or the second half of a z-pair
                        const float d = lerpf(lerpf(lerpf(load_scalar_at<HALF>(p00), load_scalar_at<HALF>(p00 + ZOFF), wz),
                                                    lerpf(load_scalar_at<HALF>(p01), load_scalar_at<HALF>(p01 + ZOFF), wz), wy),
                                              lerpf(lerpf(load_scalar_at<HALF>(p10), load_scalar_at<HALF>(p10 + ZOFF), wz),
                                                    lerpf(load_scalar_at<HALF>(p11), load_scalar_at<HALF>(p11 + ZOFF), wz), wy), wx);
                        fetch = lut_alpha(a, s_lut, d) != 0.0f;
                    }
#endif
                    if (fetch) {
                    load_row<HALF, PAIR>(p00, c000, c001);
                    load_row<HALF, PAIR>(p01, c010, c011);
                    load_row<HALF, PAIR>(p10, c100, c101);
                    load_row<HALF, PAIR>(p11, c110, c111);
#if PYVR_PF_DIST > 0
                    if (i + PYVR_PF_DIST < run_end) {
                        // A/B builds only: ask L2 for the four corner rows of the sample PYVR_PF_DIST steps ahead (the
                        // lattice makes its address known now).  A NEGATIVE result on both regimes
                        // (profiles/r02_c4_prefetch_ab.txt): C3 477 -> 398 Gsamples/s (prefetches take data-stage slots
                        // too), C4 8.7 -> 16-19 ms per frame (sparse rays: DRAM is already the bound and the extra
                        // sectors are evicted before use).  As a run-time option the branch also cost the f16 kernel
                        // 80 bytes of spills inside the loop, so it is compile-time only.
                        const float fp = fi + (float)PYVR_PF_DIST;
                        const int jx = __float2int_rd(fmaf(fp, DX, X0)), jy = __float2int_rd(fmaf(fp, DY, Y0)),
                                  jz = __float2int_rd(fmaf(fp, DZ, Z0));
                        const IDX ep = (IDX)(jx - ogx) * (IDX)vol.pitch_x + (IDX)(jy - ogy) * (IDX)vol.pitch_y + (IDX)(jz - ogz);
                        const char *q = a.tap_base + (long long)ep * ENTRY_BYTES;
                        prefetch_line(q);
                        prefetch_line(q + a.stride_y);
                        prefetch_line(q + a.stride_x);
                        prefetch_line(q + a.stride_x + a.stride_y);
                    }
#endif
                    }   // fetch
                    }   // one sample per iteration
                    }   // row layouts
                    if (fetch) {
                    // filter order of the oracle: z (memory-fastest) first, then y, then x; {s, nx} and {ny, nz} packed
                    const f32x2 tz2 = pack2(wz, wz), ty2 = pack2(wy, wy), tx2 = pack2(wx, wx);
                    float density, nx, ny, nz;
                    unpack2(lerp2(lerp2(lerp2(c000.sn, c001.sn, tz2), lerp2(c010.sn, c011.sn, tz2), ty2),
                                  lerp2(lerp2(c100.sn, c101.sn, tz2), lerp2(c110.sn, c111.sn, tz2), ty2), tx2), density, nx);
                    unpack2(lerp2(lerp2(lerp2(c000.yz, c001.yz, tz2), lerp2(c010.yz, c011.yz, tz2), ty2),
                                  lerp2(lerp2(c100.yz, c101.yz, tz2), lerp2(c110.yz, c111.yz, tz2), ty2), tx2), ny, nz);
                    visible = shade_fast(a, s_lut, density, nx, ny, nz, acc);
                    }
                }
                if (acc.a >= a.term_alpha) {
                    terminated = i < max_last;
                    last = i;
                    alive = false;
                }
                ++i;
            }
#if PYVR_DENSITY_FIRST
            warp_transparent = !__any_sync(0xffffffffu, visible);
#endif
        }
        if (hit) n_samples = (unsigned)max(last - i_lo + 1, 0);
    }

    // fragment colour -> [0,1] clamp -> blend onto the (0,0,0,0) clear -> clamp -> round to RGBA8
    if (in_image) {
        const size_t pix = ((size_t)view_index * a.height + py) * a.width + px;
        if (a.out_acc) a.out_acc[pix] = make_float4(acc.r, acc.g, acc.b, acc.a);
        if (a.out8) a.out8[pix] = fragment_to_rgba8(acc.r, acc.g, acc.b, acc.a, a.flags);
    }

    // per-warp reduction of the work counters, one atomic per counter per warp
    if (a.counters) {
        unsigned s = n_samples, f = n_fetched;
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            f += __shfl_xor_sync(0xffffffffu, f, o);
        }
        const unsigned h = __popc(__ballot_sync(0xffffffffu, hit));
        const unsigned t = __popc(__ballot_sync(0xffffffffu, terminated));
        if (lane == 0) {
            if (s) atomicAdd(a.counters + CNT_SAMPLES, (unsigned long long)s);
            if (f) atomicAdd(a.counters + CNT_FETCHED, (unsigned long long)f);
            if (h) atomicAdd(a.counters + CNT_HIT, (unsigned long long)h);
            if (t) atomicAdd(a.counters + CNT_TERM, (unsigned long long)t);
        }
    }
    }   // one tile (PYVR_PERSISTENT: next ticket)
}

template <bool STRICT, bool HALF, typename IDX, bool BRICK, int LAYOUT, bool TEX = false>
cudaError_t launch_one(const MarchArgs &a, int n_views, cudaStream_t stream) {
    const size_t smem = lut_smem_bytes(a.lut_size);
    auto kern = march_kernel<STRICT, HALF, IDX, BRICK, LAYOUT, TEX>;
    // dynamic + static shared memory above the 48 KiB default needs the opt-in (static = the interval table)
    static size_t static_smem = ~(size_t)0;
    if (static_smem == ~(size_t)0) {
        cudaFuncAttributes attr;
        cudaError_t e = cudaFuncGetAttributes(&attr, kern);
        if (e != cudaSuccess) return e;
        static_smem = attr.sharedSizeBytes;
    }
    if (smem + static_smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int tiles_x = (a.width + TILE_W - 1) / TILE_W;
    if (a.shard_count > 1) {   // this rank's share of every row of tile groups (march_kernel, "image-space sharding")
        const int group = 1 << a.shard_shift, groups_x = (tiles_x + group - 1) / group;
        tiles_x = ((groups_x + a.shard_count - 1) / a.shard_count) * group;
    }
    dim3 grid(tiles_x, (a.height + TILE_H - 1) / TILE_H, n_views);
#if PYVR_PERSISTENT
    if (a.shard_count <= 1 && a.tile_counter != nullptr) {
        static int slots = 0;       // resident CTAs of this instantiation on the whole device
        if (slots == 0) {
            int per_sm = 0, sms = 0, dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CTA_THREADS, smem);
            slots = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
        }
        cudaError_t e = cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned), stream);
        if (e != cudaSuccess) return e;
        grid = dim3(slots, 1, 1);
    }
#endif
    MarchArgs b = a;
    b.n_views = n_views;
    kern<<<grid, CTA_THREADS, smem, stream>>>(b);
    return cudaGetLastError();
}

template <bool HALF, int LAYOUT>
cudaError_t launch_fast(const MarchArgs &a, int n_views, bool wide, cudaStream_t stream) {
    if (a.vol.bricked)
        return wide ? launch_one<false, HALF, long long, true, LAYOUT>(a, n_views, stream)
                    : launch_one<false, HALF, int, true, LAYOUT>(a, n_views, stream);
    return wide ? launch_one<false, HALF, long long, false, LAYOUT>(a, n_views, stream)
                : launch_one<false, HALF, int, false, LAYOUT>(a, n_views, stream);
}

}  // namespace

cudaError_t launch_march(const MarchArgs &a, int n_views, bool half_texels, bool wide_index, cudaStream_t stream) {
    const int layout = a.vol.brick8 ? 2 : a.vol.pair ? 1 : 0;
    // bricks are marched by the fast path only (the STRICT twin of the oracle has no notion of ownership);
    // STRICT is a test mode, always uses 64-bit indices and finds its texels through texel_index() whatever the layout
    if ((a.flags & PYVR_FLAG_STRICT) != 0 && !a.vol.bricked)
        return half_texels ? launch_one<true, true, long long, false, 0>(a, n_views, stream)
                           : launch_one<true, false, long long, false, 0>(a, n_views, stream);
    if ((a.flags & PYVR_FLAG_HWTEX) != 0 && a.tex != 0)   // the texture object hides texel format and layout
        return a.vol.bricked ? launch_one<false, false, int, true, 0, true>(a, n_views, stream)
                             : launch_one<false, false, int, false, 0, true>(a, n_views, stream);
    // 2x2x2 bricks multiply pitches by 8 before the index type is chosen: keep 32-bit indices to a quarter of the range
    if (layout == 2) {
        const bool wide = wide_index || (long long)((a.vol.n[0] + 3) >> 1) * a.vol.pitch_x * 8 >= (1LL << 30);
        if (half_texels && a.two_samples) return launch_fast<true, 4>(a, n_views, wide, stream);
        return half_texels ? launch_fast<true, 2>(a, n_views, wide, stream) : launch_fast<false, 2>(a, n_views, wide, stream);
    }
    if (half_texels && layout == 1 && a.two_samples) return launch_fast<true, 3>(a, n_views, wide_index, stream);
    if (half_texels) return layout == 1 ? launch_fast<true, 1>(a, n_views, wide_index, stream)
                                        : launch_fast<true, 0>(a, n_views, wide_index, stream);
    return layout == 1 ? launch_fast<false, 1>(a, n_views, wide_index, stream)
                       : launch_fast<false, 0>(a, n_views, wide_index, stream);
}

}  // namespace pyvr
