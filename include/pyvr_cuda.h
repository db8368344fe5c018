/*
 * pyvr_cuda.h -- C ABI of the B200-native (sm_100a) backend for PyVR's volume ray-march path.
 *
 * The reference (JixianLi/pyvr v0.4.1) has no FFI: its renderer is a Python class that drives an
 * OpenGL context through moderngl.  The boundary this library replaces is therefore the
 * ModernGLManager resource layer (pyvr/moderngl_renderer/manager.py) plus the fragment shader
 * (pyvr/shaders/volume.frag.glsl); each entry point below names the manager method / uniform /
 * shader stage it stands in for.  The Python side (pyvr_b200/cuda_renderer) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success and a negative pyvr_status on failure;
 * pyvr_cuda_last_error() returns a thread-local description of the last failure.  All pointers
 * are plain host pointers unless a *_is_device argument says otherwise.  A context is bound to
 * one CUDA device and must not be used from two threads at once (the reference's GL context
 * is thread-affine in the same way).
 */
#ifndef PYVR_CUDA_H
#define PYVR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYVR_CUDA_ABI_VERSION 2
#define PYVR_IPC_HANDLE_BYTES 64   /* sizeof(cudaIpcMemHandle_t) */

typedef enum {
    PYVR_OK = 0,
    PYVR_ERR_INVALID = -1,   /* bad argument */
    PYVR_ERR_CUDA = -2,      /* a CUDA runtime call failed */
    PYVR_ERR_STATE = -3,     /* e.g. render before a volume / LUT was uploaded is fine, but
                                a batch render without a volume is not */
    PYVR_ERR_NOMEM = -4
} pyvr_status;

/* Texel storage of the packed scalar+normal volume. */
typedef enum {
    PYVR_TEXEL_F32X4 = 0,    /* {scalar, nx, ny, nz} binary32, 16 B / voxel (parity configs) */
    PYVR_TEXEL_F16X4 = 1     /* same in binary16, 8 B / voxel (2048^3 / 4096^3 configs) */
} pyvr_texel_format;

/* pyvr_params.flags */
#define PYVR_FLAG_STRICT   0x1u  /* reference-faithful arithmetic: incremental position, IEEE div/sqrt/exp,
                                    no empty-space skipping (used to compare against the oracle bit for bit) */
#define PYVR_FLAG_ESS      0x2u  /* macrocell empty-space skipping (exact: skips only zero contributions) */
#define PYVR_FLAG_NO_BLEND 0x4u  /* RGBA8 output holds (C, A) instead of the reference's blended (C*A, A*A) */
#define PYVR_FLAG_HWTEX    0x8u  /* sample through the texture unit (3-D CUDA array, hardware trilinear filter with
                                    8-bit fixed-point weights) instead of binary32 software trilinear.  What the
                                    reference's own sampler3D does on a real GPU; NOT bit-comparable with the oracle,
                                    offered only because it passes the stated tolerance (tests/test_hwtex_gpu.py).
                                    Ignored with PYVR_FLAG_STRICT. */

typedef struct pyvr_ctx pyvr_ctx;

/*
 * One view.  The march derives the ray of pixel centre (px+0.5, py+0.5), ndc = 2*(p/size) - 1, exactly
 * as ray_direction() does (volume.frag.glsl:47-54) when has_matrices != 0:
 *   eye = inv_proj * (ndc, -1, 1); eye.zw = (-1, 0); dir = normalize((inv_view * eye).xyz)
 * with inv_proj / inv_view = inverse(projection_matrix) / inverse(view_matrix) evaluated once on the
 * host in binary32 (the shader re-evaluates them per fragment).  With has_matrices == 0 the closed form
 *   dir = normalize(u * ndc.x + v * ndc.y + w)
 * is used instead (u, v, w derived in binary64; differs from the matrix path by ~1e-7 relative).
 * origin = camera_pos.  py = 0 is the BOTTOM row, as in the GL framebuffer the reference reads back.
 */
typedef struct {
    float origin[3];
    float u[3];
    float v[3];
    float w[3];
    float inv_proj[16];      /* column-major, as GL holds the uniform */
    float inv_view[16];
    int32_t has_matrices;
} pyvr_view;

/* Uniforms of the shader other than camera and bounds (volume.frag.glsl:10-19). */
typedef struct {
    float step_size;            /* uniform step_size            (renderer.py:305) */
    int32_t max_steps;          /* uniform max_steps            (renderer.py:306) */
    float reference_step_size;  /* uniform reference_step_size  (renderer.py:307-309) */
    float ambient;              /* uniform ambient_light        (renderer.py:313) */
    float diffuse;              /* uniform diffuse_light        (renderer.py:314) */
    float light_position[3];    /* uniform light_position       (renderer.py:315) */
    float light_target[3];      /* uniform light_target         (renderer.py:316) */
    float termination_alpha;    /* the shader hard-codes 0.99 (volume.frag.glsl:87); pass 0.99 for parity */
    uint32_t flags;             /* PYVR_FLAG_* */
} pyvr_params;

typedef struct {
    uint64_t samples;           /* in-box loop bodies of volume.frag.glsl:92-116 the reference would execute */
    uint64_t samples_fetched;   /* of those, how many actually fetched texels (== samples without ESS) */
    uint64_t rays_hit;          /* rays that pass intersect_box */
    uint64_t rays_terminated;   /* rays stopped by the accumulated-alpha test */
    float kernel_ms;            /* device time of the march kernel(s) of the last render call (CUDA events) */
    uint32_t kernel_launches;   /* kernels launched by the last render call */
    uint32_t views;             /* views rendered by the last render call */
} pyvr_stats;

/* --- life cycle: ModernGLManager.__init__ / cleanup (manager.py:14-41, 238-256) ------------------ */
int pyvr_cuda_create(int device, int width, int height, pyvr_ctx **out_ctx);
int pyvr_cuda_destroy(pyvr_ctx *ctx);
/* Run all work of this context on an existing cudaStream_t.  0 = the context's own non-blocking stream; the legacy
 * default stream has to be named by its explicit handle cudaStreamLegacy ((cudaStream_t)0x1). */
int pyvr_cuda_set_stream(pyvr_ctx *ctx, void *cuda_stream);

/* --- create_volume_texture + create_normal_texture + bounds uniforms
 *     (manager.py:77-135, renderer.py:131-146) ----------------------------------------------------
 * scalar: shape0*shape1*shape2 floats in numpy C order.  normals: the same voxels x 3 floats, or
 * NULL (the shader then sees normal = (density, 0, 0)).
 * The reference hands `shape` to moderngl as (width, height, depth) over the C-order bytes; the
 * library reproduces that addressing exactly (identical to data[ix,iy,iz] for cubic volumes).
 * src_is_device != 0: both pointers are device pointers on this context's device. */
int pyvr_cuda_upload_volume(pyvr_ctx *ctx, const float *scalar, const float *normals,
                            int shape0, int shape1, int shape2,
                            const float bmin[3], const float bmax[3],
                            int texel_format, int src_is_device);

/* Sort-last brick (no reference counterpart; SURVEY.md section 8 e, config C5): this context holds the
 * sub-block [origin, origin + local_dims) of a volume of global_dims texels and renders only the samples
 * whose voxel coordinate lies in [own_lo, own_hi) on every axis (the volume's outer faces are open-ended),
 * on the sample lattice of the WHOLE volume, so the partial images of all bricks composite to the
 * single-GPU result.  All index triples are in world order (x, y, z) = numpy axes (0, 1, 2) of
 * data[ix, iy, iz] (z memory-fastest); bmin/bmax are the bounds of the whole volume.  The sub-block must
 * include texel own_hi on every axis where own_hi < global_dims (the +1 ghost layer the upper trilinear
 * taps reach).  scalar/normals: local_dims[0]*local_dims[1]*local_dims[2] (x3) floats. */
int pyvr_cuda_upload_brick(pyvr_ctx *ctx, const float *scalar, const float *normals,
                           const int local_dims[3], const int global_dims[3], const int origin[3],
                           const int own_lo[3], const int own_hi[3],
                           const float bmin[3], const float bmax[3], int texel_format, int src_is_device);

/* Device-side synthetic volume: what
 *   v = create_sample_volume(size, shape); Volume(v, compute_normal_volume(v), bmin, bmax) -> load_volume
 * would upload (pyvr/datasets/synthetic.py:10-122), generated and packed on the device without touching
 * the host.  With origin/local_dims/own_lo/own_hi non-NULL only that sub-block is generated, as a
 * sort-last brick (arguments as pyvr_cuda_upload_brick); all NULL = the whole volume.  kernel_ms (may be
 * NULL) receives the device time of generation + pack. */
#define PYVR_SHAPE_SPHERE 0
#define PYVR_SHAPE_TORUS 1
#define PYVR_SHAPE_DOUBLE_SPHERE 2
int pyvr_cuda_generate_volume(pyvr_ctx *ctx, int shape, int size, const int local_dims[3], const int origin[3],
                              const int own_lo[3], const int own_hi[3], const float bmin[3], const float bmax[3],
                              int texel_format, float *kernel_ms);
/* Test aid: unpack the stored texels into the reference's two arrays (host pointers, numpy C order of the
 * stored block; either may be NULL). */
int pyvr_cuda_read_texels(pyvr_ctx *ctx, float *scalar, float *normals);

/* Image-space sharding (SURVEY.md section 8 e, config C4): this context marches only its own tile groups --
 * groups of 2^s x 2^s CTA tiles of 16x8 pixels (option "shard_shift", default s = 1), owner(gx, gy) =
 * (gx + gy) mod count -- and only those CTAs are launched.  By default the rest of the frame is cleared first, so
 * the frames of all ranks add up (e.g. ncclReduce SUM over uint8) to the full frame, bit for bit; with option
 * "shard_in_place" = 1 the other pixels are left untouched, so that all ranks can write one shared (peer-mapped)
 * frame directly.  count = 1 restores normal rendering. */
int pyvr_cuda_set_pixel_shard(pyvr_ctx *ctx, int rank, int count);

/* --- create_rgba_transfer_function_texture (manager.py:137-186): (size,4) RGBA binary32 ---------- */
int pyvr_cuda_set_lut(pyvr_ctx *ctx, const float *rgba, int size);

/* --- set_camera (renderer.py:148-172): the three uniforms as the reference writes them:
 *     16 floats of view_matrix.tobytes(), 16 of projection_matrix.tobytes() (GL reads both
 *     column-major) and camera_pos.  Derives the pyvr_view on the host in binary64. ---------------- */
int pyvr_cuda_set_camera(pyvr_ctx *ctx, const float view[16], const float proj[16], const float cam_pos[3]);
/* Same derivation without a context (used to build batches). */
int pyvr_cuda_view_from_matrices(const float view[16], const float proj[16], const float cam_pos[3],
                                 pyvr_view *out);
int pyvr_cuda_set_view(pyvr_ctx *ctx, const pyvr_view *view);

/* --- _update_render_config + _update_light (renderer.py:303-316) --------------------------------- */
int pyvr_cuda_set_params(pyvr_ctx *ctx, const pyvr_params *params);

/* --- render(): clear + blend + draw + fbo.read (renderer.py:209-219, manager.py:212-230) ---------
 * out: width*height*4 bytes RGBA8, row 0 = bottom.  out_is_device != 0: device pointer.
 * Rendering with no volume or no LUT loaded yields the cleared frame (all zero), as the reference
 * tolerates (tests/test_moderngl_renderer/test_volume_renderer.py:306-327). */
int pyvr_cuda_render(pyvr_ctx *ctx, uint8_t *out, int out_is_device);
/* n views in one call; out holds n consecutive frames. */
int pyvr_cuda_render_batch(pyvr_ctx *ctx, const pyvr_view *views, int n, uint8_t *out, int out_is_device);
/* The fragment colour BEFORE blending/quantisation, width*height*4 floats (acc_rgb, acc_a). */
int pyvr_cuda_render_accum(pyvr_ctx *ctx, float *out, int out_is_device);

/* Sort-last relay (exact): continue the accumulation `in_accum` (device, width*height*4 floats, the
 * fragment colours of the bricks IN FRONT of this context's brick; NULL = nothing in front) through this
 * brick and write the result to `out_accum` (device; may alias in_accum).  Passing the image through the
 * bricks in visibility order reproduces the single-GPU march operation for operation, stop rule included. */
int pyvr_cuda_render_accum_relay(pyvr_ctx *ctx, const float *in_accum, float *out_accum);

int pyvr_cuda_get_stats(pyvr_ctx *ctx, pyvr_stats *out);

/* --- compute_normal_volume (pyvr/datasets/synthetic.py:109-122) ----------------------------------
 * in: n0*n1*n2 floats (C order); out: n0*n1*n2*3 floats.  flags: PYVR_NORMALS_DEVICE_BUFFERS = both are device
 * pointers on `device`; PYVR_NORMALS_RELAXED = quotients as g * (1/norm), within 2 ulp of the reference's divisions
 * (the path's tolerance is 1e-5 relative) instead of bit-identical to them.  kernel_ms (may be NULL) receives the
 * stencil kernel's device time. */
#define PYVR_NORMALS_DEVICE_BUFFERS 1
#define PYVR_NORMALS_RELAXED 2
int pyvr_cuda_compute_normals(int device, const float *in, float *out, int n0, int n1, int n2,
                              int flags, float *kernel_ms);

/* --- sort-last compositing (no reference counterpart) ----------------------------------------------
 * Partial images are the pre-blend fragment colours pyvr_cuda_render_accum produces: n_pixels * 4 floats
 * (premultiplied rgb, alpha), all DEVICE pointers here.  `back` may be a peer-mapped pointer
 * (pyvr_cuda_ipc_open): the merge then reads it across NVLink.  cuda_stream: a cudaStream_t or NULL. */
/* out = front over back; a front pixel at or above termination_alpha hides the back pixel (the shader's
 * stop rule, volume.frag.glsl:87, at brick granularity).  out may alias front. */
int pyvr_cuda_composite_over(int device, const float *front, const float *back, float *out, size_t n_pixels,
                             float termination_alpha, void *cuda_stream);
/* Blend onto the cleared target + RGBA8 quantisation of a composited image (manager.py:217-220, 29). */
int pyvr_cuda_finalize_rgba8(int device, const float *accum, uint8_t *out, size_t n_pixels, uint32_t flags,
                             void *cuda_stream);
/* Last round of a binary swap fused with finalize: out = rgba8(front over back) for n_pixels; accum_out (may be
 * NULL, may alias front) also receives the merged floats.  out may be a peer-mapped frame. */
int pyvr_cuda_composite_finalize(int device, const float *front, const float *back, float *accum_out, uint8_t *out,
                                 size_t n_pixels, float termination_alpha, uint32_t flags, void *cuda_stream);
/* Stream-ordered flags between the GPUs of a node (uint32 counters in pyvr_cuda_device_alloc / peer-mapped memory).
 * signal: everything enqueued on the stream before is visible system-wide, then *flag = value (release, system
 * scope).  wait: the stream stalls until every one of flags[0..n_flags) has reached `value` (counters only grow;
 * comparison is wrap-safe).  They replace host barriers / NCCL fences around peer reads and writes. */
int pyvr_cuda_flag_signal(int device, uint32_t *flag, uint32_t value, void *cuda_stream);
int pyvr_cuda_flag_wait(int device, const uint32_t *flags, int n_flags, uint32_t value, void *cuda_stream);
/* The same release store to n flags (one per peer) in one launch. */
int pyvr_cuda_flag_signal_many(int device, uint32_t *const *flags, int n_flags, uint32_t value, void *cuda_stream);
/* A whole binary swap in one call: per round [signal_before] -> wait(wait_flag >= value) -> merge -> signal(signal_done,
 * signal_next), enqueued back to back on the stream (a Python loop over the single calls leaves the GPU idle between
 * them: 8 GPUs, 3 rounds, ~0.3 ms per frame).  A round with out8 != NULL is the fused last round
 * (pyvr_cuda_composite_finalize); otherwise pyvr_cuda_composite_over into `out`.  NULL flag pointers are skipped. */
typedef struct {
    const float *front, *back;   /* n_pixels float4 each; either may be peer-mapped */
    float *out;                  /* merged floats (may alias front or back); may be NULL when out8 is set */
    uint8_t *out8;               /* NULL, or the RGBA8 destination of the fused last round (may be peer-mapped) */
    uint64_t n_pixels;
    uint32_t *signal_before;     /* partner's counter: "my image of this round is complete" (round 0) */
    const uint32_t *wait_flag;   /* own counter: the partner's image of this round is complete */
    uint32_t *signal_done;       /* partner's counter: "I have finished reading your image" */
    uint32_t *signal_next;       /* next partner's counter: "my image of the next round is complete" */
} pyvr_swap_round;
int pyvr_cuda_binary_swap(int device, const pyvr_swap_round *rounds, int n_rounds, uint32_t value,
                          float termination_alpha, uint32_t flags, void *cuda_stream);
/* Plain cudaMalloc memory, rounded up to whole 2 MiB blocks so that the buffer is an allocation of its own: a CUDA
 * IPC handle names the enclosing allocation, and the driver packs smaller requests into shared blocks. */
int pyvr_cuda_device_alloc(int device, size_t bytes, void **out);
int pyvr_cuda_device_free(int device, void *ptr);
/* CUDA IPC: share a pyvr_cuda_device_alloc buffer with the other single-GPU processes of the node.  export fails
 * with PYVR_ERR_INVALID if ptr is not the start of an allocation. */
int pyvr_cuda_ipc_export(int device, void *ptr, uint8_t handle[PYVR_IPC_HANDLE_BYTES]);
int pyvr_cuda_ipc_open(int device, const uint8_t handle[PYVR_IPC_HANDLE_BYTES], void **out);
int pyvr_cuda_ipc_close(int device, void *ptr);
/* Synchronous copy; kind: 1 = host->device, 2 = device->host, 3 = device->device (peer pointers allowed). */
int pyvr_cuda_memcpy(int device, void *dst, const void *src, size_t bytes, int kind, void *cuda_stream);
int pyvr_cuda_stream_synchronize(int device, void *cuda_stream);

/* --- misc ------------------------------------------------------------------------------------------ */
/* Tuning knobs (no reference counterpart).  "swizzle": 1 (default) = bank-rotating row/plane padding of the packed
 * texel layout, 0 = aligned rows (for A/B profiling); must be set before pyvr_cuda_upload_volume.  "pair": z-pair
 * entries (-1 auto, 0 off, 1 on; next upload).  "shard_shift", "shard_in_place": see pyvr_cuda_set_pixel_shard.
 * "brick8": 2x2x2-texel brick layout (-1 auto = f16x4 texels and a volume edge above 0.7 x the frame height, i.e.
 * sparse rays; 0 off; 1 on; next upload).  "two_samples": f16x4 march with several samples of a run in flight per ray
 * (-1 auto = the same sparse-ray rule, 0 off, 1 on).
 * "async_device_output": 1 = renders into device buffers return without a host synchronisation (the pixels are valid
 * in stream order; pyvr_cuda_get_stats waits for the counters). */
int pyvr_cuda_set_option(pyvr_ctx *ctx, const char *key, int value);
/* What is in effect for the loaded volume, automatic choices resolved: "pair", "brick8", "two_samples",
 * "async_device_output". */
int pyvr_cuda_get_option(pyvr_ctx *ctx, const char *key, int *value);
/* Roofline denominators measured on the spot (no reference counterpart; csrc/bandwidth.cu): bytes per second
 * delivered to registers by coalesced 128-bit loads that hit L1 (level 1: the SM load-return path that bounds
 * the march's texel gather) or stream from L2 with L1 bypassed (level 2); level 3: the DRAM rate of a streaming
 * kernel with the normals stencil's traffic mix (4 bytes read, 12 written per element, no arithmetic), which a
 * 1:1 copy rate overstates.  *gbs in 1e9 bytes/s. */
int pyvr_cuda_measure_cache_bandwidth(int device, int level, double *gbs);
/* Page-locked host memory for frame read-back at full PCIe rate (cudaMallocHost / cudaFreeHost). */
int pyvr_cuda_host_alloc(size_t bytes, void **out);
int pyvr_cuda_host_free(void *ptr);
int pyvr_cuda_device_count(int *count);
int pyvr_cuda_abi_version(void);
const char *pyvr_cuda_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PYVR_CUDA_H */
