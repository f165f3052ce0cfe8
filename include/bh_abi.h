/* bh_abi.h — C ABI of libbhray.so: bhusie's per-pixel geodesic ray pass on B200 (sm_100a).
 *
 * The reference has no FFI surface; its boundary for this path is the Rust pass-object API in
 * src/renderer/pipelines/ (paths below are relative to the cleggacus/bhusie tree):
 *
 *   RayPipelineDescriptor{device,queue,resolution,camera_buffer,black_hole_buffer,material_buffer,
 *                         model_buffer,ray_details_buffer,prev_texture_view}   ray_pipeline.rs:16-26
 *   RayPipeline::new(descriptor) -> Self                                       ray_pipeline.rs:36-295
 *   RayPipeline::output_view(&self) -> &TextureView                            ray_pipeline.rs:297-299
 *   RayPipeline::pass(&mut self, &mut ComputePass)                             ray_pipeline.rs:301-309
 *   SkyPipeline::{new, output_view, pass}                                      sky_pipeline.rs:18,136,140
 *
 * Each entry point below names the reference item it replaces.  Conventions: every function
 * returns 0 on success or a negative errno-style code (never throws or aborts across the
 * boundary); bh_last_error() gives the text for the calling thread.  The caller owns all host
 * memory; the library owns all device memory unless bh_*_bind_output is used.  One context per
 * CUDA device; calls on one context are not re-entrant (the reference drives its passes from one
 * thread, src/app.rs:108-114); different contexts may be driven from different threads.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with BH_ERR_NODEV.
 *
 * The three small uniforms cross the ABI as the verbatim bytes of the reference's #[repr(C)]
 * structs (what it passes to queue.write_buffer at src/renderer/mod.rs:386-388).
 */
#ifndef BH_ABI_H
#define BH_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BH_ABI_VERSION 2

/* error codes (negative errno values) */
#define BH_OK            0
#define BH_ERR_INVALID  (-22)   /* EINVAL: bad argument */
#define BH_ERR_NOMEM    (-12)   /* ENOMEM: host or device allocation failed */
#define BH_ERR_NODEV    (-19)   /* ENODEV: no usable CUDA device */
#define BH_ERR_CUDA     (-5)    /* EIO:    a CUDA call failed (text in bh_last_error) */
#define BH_ERR_STATE    (-1)    /* EPERM:  call sequence error (e.g. pass before textures are set) */
#define BH_ERR_NOENT    (-2)    /* ENOENT: file not found */
#define BH_ERR_TOOBIG   (-7)    /* E2BIG:  mesh exceeds MAX_MODEL_VERTICES */
#define BH_ERR_NUMERIC  (-34)   /* ERANGE: the RK accept loop of ray.wgsl:425-451 would not terminate (e_max > 1) */
#define BH_ERR_TIMEOUT  (-110)  /* ETIMEDOUT: a cross-GPU / cross-process wait gave up (a peer never signalled) */

/* ---- uniform byte layouts (verbatim) ------------------------------------ */

typedef struct bh_camera_uniform {          /* CameraUniform, src/scene/camera.rs:66-73 */
    float position[3]; uint32_t _padding;
    float forward[3];  float fov;
} bh_camera_uniform;                        /* 32 B */

typedef struct bh_black_hole_uniform {      /* BlackHoleUniform, src/scene/blackhole.rs:37-51 */
    float accretion_disk_inner, accretion_disk_outer, rotation_speed, relativity_sphere_radius;
    float position[3]; int32_t show_disk_texture;
    float normal[3];   int32_t show_red_shift;
    float rotation_matrix[12];              /* three columns (right, up, forward), each vec3 + pad */
    float feather_amount;
    int32_t pad[8];
} bh_black_hole_uniform;                    /* 132 B */

typedef struct bh_ray_details {             /* RayDetails, src/renderer/pipelines/ray_pipeline.rs:3-14 */
    int32_t material_count, model_count;
    float   time;
    int32_t integration_method;             /* 0 Euler, 1 Cash–Karp RK (ray.wgsl:29) */
    float   step_size;
    int32_t max_iterations;
    float   angle_division_threshold;
    int32_t highlight_interpolation;
} bh_ray_details;                           /* 32 B */

#define BH_MAX_MODEL_VERTICES 524288        /* src/renderer/triangle.rs:7 */
#define BH_MAX_MODELS         1             /* src/renderer/triangle.rs:6 */
#define BH_MODEL_UNIFORM_SIZE 48234572      /* sizeof(ModelUniform), src/renderer/triangle.rs:268-285 */

typedef enum bh_texture_slot {              /* the three include_bytes! textures, ray_pipeline.rs:63-70 */
    BH_TEX_COLOR = 0,                       /* color.png  256x256   (t_temp, ray.wgsl:13)  */
    BH_TEX_DISK  = 1,                       /* disk.png   1000x1000 (t_disk, ray.wgsl:15)  */
    BH_TEX_SKY   = 2                        /* sky.png    6000x3000 (t_sky,  ray.wgsl:17)  */
} bh_texture_slot;

typedef struct bh_ctx          bh_ctx;
typedef struct bh_ray_pipeline bh_ray_pipeline;
typedef struct bh_sky_pipeline bh_sky_pipeline;

/* ---- library ------------------------------------------------------------------------------- */
int         bh_abi_version(void);
const char *bh_last_error(void);            /* thread-local text of the last failure ("" if none) */

/* ---- context: replaces the wgpu Device/Queue pair + Renderer-owned scene buffers
 *      (src/renderer/mod.rs:63-89,113-114) ---------------------------------------------------- */
int  bh_ctx_create(int cuda_device, bh_ctx **out);
void bh_ctx_destroy(bh_ctx *ctx);            /* destroy every pipeline / post pass of the context first */

/* Numeric mode of every kernel launched through this context.  WGSL leaves contraction and the
 * rounding of `/` (2.5 ulp) to the shader compiler; the two modes are the two ends of that latitude:
 *   LITERAL  one IEEE-754 binary32 operation per WGSL expression node, in source order; true division.
 *            Bit-comparable with the oracle's "contract" flavour.
 *   FUSED    (default) x*y+z contracted to one fma where the expression tree has that shape, vec3/f32 computed as
 *            vec3 * (1/f32).  ~1.5x fewer instructions.  Bit-comparable with the oracle's "fused" flavour. */
typedef enum bh_numeric_mode { BH_NUMERIC_LITERAL = 0, BH_NUMERIC_FUSED = 1 } bh_numeric_mode;
int  bh_ctx_set_numeric_mode(bh_ctx *ctx, bh_numeric_mode mode);
int  bh_ctx_get_numeric_mode(const bh_ctx *ctx);

/* texture::Texture::from_bytes (src/renderer/texture.rs:10-80) minus the PNG decode: RGBA8 unorm
 * (no sRGB), bilinear, clamp-to-edge, one mip.  Copies rgba8 (w*h*4 bytes) to the device. */
int  bh_ctx_set_texture(bh_ctx *ctx, bh_texture_slot slot, const uint8_t *rgba8, uint32_t w, uint32_t h);

/* SURVEY §8 f4 — the reference's offline generator of disk.png (perlin/src/main.rs:6-148: 4 octaves of hash-gradient
 * Perlin noise, spiral-warped, merged) run on the GPU: w x h RGBA8 with r=g=b=a (the tool uses 1000 x 1000).
 * host_rgba8 (nullable) receives the texels; install != 0 also makes it the context's BH_TEX_DISK, so a scene needs
 * no binary disk asset. */
int  bh_ctx_generate_disk_texture(bh_ctx *ctx, uint32_t w, uint32_t h, uint8_t *host_rgba8, int install);

/* ModelArrayBuffer::update_buffer (src/renderer/array_buffer.rs:71-79): `bytes` is
 * ModelUniform[count] verbatim, nbytes == count*BH_MODEL_UNIFORM_SIZE, count <= BH_MAX_MODELS.
 * Synchronous (returns after the copy).  bh_ctx_upload_models_async enqueues the same copy on
 * `cuda_stream` from caller-pinned memory (the reference re-sends the blob every frame). */
int  bh_ctx_upload_models(bh_ctx *ctx, const void *bytes, size_t nbytes);
int  bh_ctx_upload_models_async(bh_ctx *ctx, const void *pinned_bytes, size_t nbytes, void *cuda_stream);
/* cheap per-frame update of ModelUniform.position/.visible (the fields the UI edits,
 * src/ui/model_settings.rs:39-48) without re-sending 48 MB */
int  bh_ctx_set_model_header(bh_ctx *ctx, uint32_t index, const float position[3], int32_t visible);

/* ---- ray pass: replaces RayPipeline -------------------------------------------------------- */
/* RayPipeline::new.  prev == NULL is the 1x1 "base" texture of mod.rs:151-168 (every pixel traced);
 * otherwise prev's output is this level's t_prev (ray.wgsl:19) and must outlive this object. */
int  bh_ray_pipeline_create(bh_ctx *ctx, uint32_t width, uint32_t height,
                            const bh_ray_pipeline *prev, bh_ray_pipeline **out);
void bh_ray_pipeline_destroy(bh_ray_pipeline *p);

/* Image-space sharding for multi-GPU (no reference counterpart): this pipeline renders
 * only the cyclic row bands  { y : (y / band_rows) % n_ranks == rank }  of the width x height frame
 * into a compact band-major buffer of bh_ray_pipeline_local_rows() rows.  Default (1 rank): whole frame. */
int      bh_ray_pipeline_set_tiling(bh_ray_pipeline *p, uint32_t band_rows, uint32_t rank, uint32_t n_ranks);
uint32_t bh_ray_pipeline_local_rows(const bh_ray_pipeline *p);

/* Optional per-pixel aux buffers for validation: bit 0 = BVH hit triangle index (int32, value of
 * `index` at ray.wgsl:333 for the composited triangle, -1 if none), bit 1 = integrator step count
 * (uint32), bit 2 = pixel class (uint8: 0 traced/base 1 copied 2 interpolated 3 traced/fine). */
#define BH_AUX_HIT   1u
#define BH_AUX_STEPS 2u
#define BH_AUX_CLASS 4u
int  bh_ray_pipeline_enable_aux(bh_ray_pipeline *p, uint32_t aux_mask);

/* Render into caller-owned device memory (local_rows*width*16 B, 16-byte aligned) instead of the
 * pipeline's own buffer — lets the host hand in a torch tensor / NCCL send buffer.  NULL restores. */
int  bh_ray_pipeline_bind_output(bh_ray_pipeline *p, void *device_rgba32f);

/* Multi-GPU without a gather step: render this pipeline's rows directly into a FULL width x height RGBA32F frame at
 * their global row index.  The frame may be local, or another GPU's memory mapped through bh_shared_frame_open — the
 * kernel's 16-byte stores then travel over NVLink while it is still tracing, so the "collective" is fused into the
 * pass.  NULL restores the compact local buffer.  (With a frame bound, bh_ray_pipeline_read cannot return RGBA.) */
int  bh_ray_pipeline_bind_frame(bh_ray_pipeline *p, void *device_frame_rgba32f);
/* CUDA-IPC plumbing for that frame: the owner allocates and exports a 64-byte handle (sent to the other processes of
 * the node by any means, e.g. a torch.distributed broadcast); the others map it; both release it when done. */
int  bh_shared_frame_create(bh_ctx *ctx, size_t nbytes, void **device_ptr, uint8_t handle_out[64]);
int  bh_shared_frame_open(bh_ctx *ctx, const uint8_t handle[64], void **device_ptr);
int  bh_shared_frame_release(bh_ctx *ctx, void *device_ptr, int owner);

/* RayPipeline::pass: enqueue the pass on `cuda_stream` (a cudaStream_t; NULL = default stream) and
 * return — same contract as recording into a ComputePass.  The three pointers are HOST bytes, read
 * before return. */
int  bh_ray_pipeline_pass(bh_ray_pipeline *p, const bh_camera_uniform *camera,
                          const bh_black_hole_uniform *black_hole, const bh_ray_details *details,
                          void *cuda_stream);

/* pass + read-back in one call for callers whose consumer lives on the host side of the bus (the reference's wgpu
 * post chain): same as bh_ray_pipeline_pass, then the RGBA32F output is copied to `pinned_host_rgba32f`
 * (local_rows*width*16 B of page-locked memory).  On the base level the frame is traced as n_chunks (1..16) row bands and
 * each band's D2H copy overlaps the tracing of the next.  n_chunks == 0 selects ZERO-COPY: the kernel stores finished pixels
 * directly into the (mapped) page-locked buffer over PCIe while it traces, so no copy is enqueued at all (the pipeline's
 * own device buffer is not written by that pass).  Asynchronous; bh_ray_pipeline_sync() waits for kernels and copies. */
int  bh_ray_pipeline_pass_to_host(bh_ray_pipeline *p, const bh_camera_uniform *camera,
                                  const bh_black_hole_uniform *black_hole, const bh_ray_details *details,
                                  float *pinned_host_rgba32f, uint32_t n_chunks, void *cuda_stream);
int  bh_ray_pipeline_sync(bh_ray_pipeline *p);

/* RayPipeline::output_view: device pointer, row-major RGBA32F, pitch = width*16 B, local_rows rows.
 * alpha==1: finished colour; alpha==0: rgb is the escaped-ray direction (ray.wgsl:592-595). */
const float *bh_ray_pipeline_output(const bh_ray_pipeline *p);
uint32_t     bh_ray_pipeline_width(const bh_ray_pipeline *p);
uint32_t     bh_ray_pipeline_height(const bh_ray_pipeline *p);

/* Synchronous read-back (waits for the last pass).  Every pointer is nullable. */
int  bh_ray_pipeline_read(bh_ray_pipeline *p, float *host_rgba32f, int32_t *host_hit,
                          uint32_t *host_steps, uint8_t *host_class);

typedef struct bh_pass_stats {              /* totals of the last pass (waits for it) */
    uint64_t ray_steps;                     /* integrator calls: next_ray_rk / next_ray_euler (ray.wgsl:525-531) */
    uint64_t px_traced, px_copied, px_interp;
    uint64_t node_visits, tri_tests;        /* BVH inner-node visits / hit_triangle calls actually made: the walk is bounded by the
                                             * relativity-sphere hit when there is one (a triangle behind it is discarded anyway,
                                             * ray.wgsl:562), so these can be below the literal traversal's counts */
    uint64_t tex_samples;                   /* bilinear samples taken (disk, LUT, sky) */
    uint64_t rk_reject;                     /* RK steps whose error norm exceeded 1 (reference would spin; Q5) */
    uint64_t stack_overflow;                /* BVH pushes beyond the reference's 19-entry stack (Q16) */
} bh_pass_stats;
/* Fills *out; returns BH_ERR_NUMERIC (with *out still valid) when rk_reject != 0. */
int  bh_ray_pipeline_stats(bh_ray_pipeline *p, bh_pass_stats *out);

/* ---- sky resolve: replaces SkyPipeline ----------------------------------------------------- */
typedef enum bh_sky_format {
    BH_SKY_RGBA16F = 0,                     /* the reference's Rgba16Float (sky_pipeline.rs:34) */
    BH_SKY_RGBA32F = 1                      /* the value before the f16 store (parity) */
} bh_sky_format;
int  bh_sky_pipeline_create(bh_ctx *ctx, const bh_ray_pipeline *prev, bh_sky_format format, bh_sky_pipeline **out);
/* Same pass over a raw width x height RGBA32F frame in device memory instead of a pipeline's output — the frame a
 * multi-GPU render assembles on one device (bh_ray_pipeline_bind_frame). */
int  bh_sky_pipeline_create_for_frame(bh_ctx *ctx, const void *device_frame_rgba32f, uint32_t width, uint32_t height,
                                      bh_sky_format format, bh_sky_pipeline **out);
void bh_sky_pipeline_destroy(bh_sky_pipeline *p);
int  bh_sky_pipeline_bind_output(bh_sky_pipeline *p, void *device_rgba);
int  bh_sky_pipeline_pass(bh_sky_pipeline *p, void *cuda_stream);
const void *bh_sky_pipeline_output(const bh_sky_pipeline *p);
int  bh_sky_pipeline_read(bh_sky_pipeline *p, void *host_rgba);   /* local_rows*width*(8|16) B */

/* ---- multi-GPU (no reference counterpart: bhusie is single-GPU) ----------------------------------------
 * Pixels of a level are independent (ray.wgsl:167-243), so the frame is cut into cyclic row bands, band b -> device
 * b mod N; the scene is replicated; the coarse levels of the adaptive grid (12.5 % of the pixels) are replicated too, the
 * last level is tiled, and every device's ray kernel stores its finished pixels straight into ONE frame on device 0 over
 * NVLink (peer stores, fused into the pass).  The sky resolve then runs on device 0 over the assembled frame.
 *
 * (1) bh_frame_multi: one process, ONE host thread, N devices — what a Rust host like bhusie's (one thread,
 *     src/app.rs:108-114) can drive.  No torch, no NCCL: peer access + CUDA events.  `ctxs` are N contexts on N distinct
 *     devices, each with the same textures / models uploaded by the caller; ctxs[0] owns the frame. */
typedef struct bh_frame_multi bh_frame_multi;
typedef struct bh_frame_multi_desc {
    uint32_t base_width, base_height;       /* level 0 (levels == 1: the frame itself) */
    uint32_t levels;                        /* 1 = single level, every pixel traced; n = the reference's adaptive grid, */
    uint32_t multiplier;                    /*     S_k = multiplier*S_(k-1) - (multiplier-1)  (mod.rs:177-206: base 72x41, x3, 4 levels) */
    uint32_t band_rows;                     /* rows per cyclic band of the last level (8 is a good default) */
    int32_t  sky_format;                    /* -1: no sky resolve; else bh_sky_format: resolve the assembled frame on ctxs[0] */
} bh_frame_multi_desc;
int  bh_frame_multi_create(bh_ctx *const *ctxs, uint32_t n_devices, const bh_frame_multi_desc *desc, bh_frame_multi **out);
void bh_frame_multi_destroy(bh_frame_multi *fm);
uint32_t bh_frame_multi_width(const bh_frame_multi *fm);       /* final level */
uint32_t bh_frame_multi_height(const bh_frame_multi *fm);
/* Renderer::render's compute pass (mod.rs:406-421) on N devices: enqueue every level on every device (the object's own
 * streams), join on device 0, resolve the sky there.  Asynchronous; a later pass waits for the previous frame's consumers. */
int  bh_frame_multi_pass(bh_frame_multi *fm, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                         const bh_ray_details *details);
/* Same, but the LAST level's pixels are stored by each device straight into `pinned_host_rgba32f` (width*height*16 B of
 * page-locked, mapped, portable memory) over that device's own PCIe link; no device frame, no sky resolve. */
int  bh_frame_multi_pass_to_host(bh_frame_multi *fm, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                                 const bh_ray_details *details, float *pinned_host_rgba32f);
int  bh_frame_multi_sync(bh_frame_multi *fm);
const float *bh_frame_multi_output(const bh_frame_multi *fm);      /* device 0: assembled RGBA32F frame of the last level */
const void  *bh_frame_multi_sky_output(const bh_frame_multi *fm);  /* device 0: resolved frame (NULL when sky_format < 0) */
int  bh_frame_multi_read(bh_frame_multi *fm, float *host_rgba32f, void *host_sky);     /* sync; both nullable */
/* totals of the last frame: coarse levels counted once (device 0's copy), last level summed over the devices;
 * *elapsed_ms (nullable) = device time of the frame from its first launch on device 0 to the end of the join / sky */
int  bh_frame_multi_stats(bh_frame_multi *fm, bh_pass_stats *out, float *elapsed_ms);

/* (2) one process per GPU (torch.distributed / MPI launchers): rank 0 exports its frame with bh_shared_frame_create, the
 *     others map it (bh_shared_frame_open) and bind it (bh_ray_pipeline_bind_frame).  Ordering needs no collective: the
 *     shared allocation carries BH_SHARED_FLAGS 32-bit flags after the frame; a rank signals "my rows of frame k are
 *     stored" by a stream-ordered store of k into its flag (after its ray kernel, so the pixels are visible first), and
 *     rank 0 waits — stream-ordered, on the device — until the flags of all ranks reach k.  The same pair orders
 *     "rank 0 has consumed frame k" back to the writers.  Waits give up after timeout_ms (the error is reported by
 *     bh_ctx_check_async) instead of hanging the GPU. */
#define BH_SHARED_FLAGS 64
uint32_t *bh_shared_frame_flags(void *device_ptr, size_t nbytes);   /* device pointer of flag 0 (nbytes as given to create) */
int  bh_stream_signal(bh_ctx *ctx, uint32_t *device_flag, uint32_t value, void *cuda_stream);
int  bh_stream_wait(bh_ctx *ctx, const uint32_t *device_flags, uint32_t n_flags, uint32_t value, uint32_t timeout_ms,
                    void *cuda_stream);
int  bh_ctx_check_async(bh_ctx *ctx);        /* BH_ERR_TIMEOUT if a bh_stream_wait on this context has given up since the last call */

/* (3) end to end on N GPUs: a page-locked host frame in POSIX shared memory that every rank maps and registers with CUDA,
 *     so each rank's kernel stores its own bands into the caller's frame over its own PCIe link (no NVLink hop, no D2H
 *     copy).  The segment carries BH_SHARED_FLAGS host-side flags behind the frame for the processes to order themselves. */
typedef struct bh_host_frame bh_host_frame;
int   bh_host_frame_create(bh_ctx *ctx, const char *shm_name, size_t nbytes, int create, bh_host_frame **out);
void  bh_host_frame_destroy(bh_host_frame *hf, int unlink_name);
float *bh_host_frame_ptr(const bh_host_frame *hf);
int   bh_host_frame_signal(bh_host_frame *hf, uint32_t slot, uint32_t value);
int   bh_host_frame_wait(bh_host_frame *hf, uint32_t first_slot, uint32_t n_slots, uint32_t value, uint32_t timeout_ms);
/* A (tiled) pipeline's pass with its rows stored at their GLOBAL row index into a full width x height page-locked host frame
 * (bh_host_frame_ptr, or any cudaHostRegister'ed / cudaHostAlloc'ed mapped memory).  Asynchronous on `cuda_stream`. */
int   bh_ray_pipeline_pass_to_host_frame(bh_ray_pipeline *p, const bh_camera_uniform *camera,
                                         const bh_black_hole_uniform *black_hole, const bh_ray_details *details,
                                         float *mapped_host_frame_rgba32f, void *cuda_stream);

/* ---- post chain: the passes that consume the sky pass's output.  One object per pass, the same
 *      new / pass / output_view triple as the reference's BloomPipeline (bloom_pipline.rs:20,156,160; shaders
 *      bloom_down.wgsl / bloom_up.wgsl), MixPipeline (mix_pipeline.rs:24; mix.wgsl), HDRPipeline (hdr_pipeline.rs;
 *      hdr.wgsl ACES) and FXAAPipeline (fxaa_pipline.rs; fxaa.wgsl).  Inputs and outputs are device images:
 *      RGBA16F (8 B/pixel) everywhere, RGBA8 with sRGB-encoded rgb (4 B/pixel) out of FXAA (fxaa_pipline.rs:120). ---- */
typedef enum bh_post_kind {
    BH_POST_BLOOM_DOWN = 0, BH_POST_BLOOM_UP = 1, BH_POST_MIX = 2, BH_POST_HDR = 3, BH_POST_FXAA = 4
} bh_post_kind;
typedef struct bh_mix_details  { float mix_ratio; } bh_mix_details;                       /* mix_pipeline.rs:5-7  */
typedef struct bh_fxaa_details { float edge_threshold_min, edge_threshold_max; int32_t iterations; float subpixel_quality; } bh_fxaa_details; /* fxaa_pipline.rs:74-80 */
typedef struct bh_post_pass bh_post_pass;
/* in1: the pass's input image (MIX: texture_view_1, the sky output); in2: MIX only (texture_view_2, the bloom output),
 * same size as in1 and as the output.  The pass allocates its own out_w x out_h output. */
int  bh_post_pass_create(bh_ctx *ctx, bh_post_kind kind, uint32_t out_w, uint32_t out_h,
                         const void *in1_device, uint32_t in1_w, uint32_t in1_h, const void *in2_device, bh_post_pass **out);
void bh_post_pass_destroy(bh_post_pass *p);
/* details: bh_mix_details for MIX, bh_fxaa_details for FXAA (host bytes, read before return), NULL otherwise */
int  bh_post_pass_run(bh_post_pass *p, const void *details, void *cuda_stream);
const void *bh_post_pass_output(const bh_post_pass *p);
int  bh_post_pass_read(bh_post_pass *p, void *host);

/* ---- host-side scene preparation: replaces load_model + Model::build_bvh ---------------------
 * (src/renderer/model.rs:7-87, src/renderer/triangle.rs:143-259).  Pure host code; `model_uniform`
 * is a caller-owned BH_MODEL_UNIFORM_SIZE-byte buffer that receives the verbatim ModelUniform. */
typedef struct bh_model_info {
    int32_t point_count, normal_count, triangle_count;
    int32_t nodes_used, max_depth, leaf_count, max_leaf_size;
} bh_model_info;
int  bh_model_load_obj(const char *path, void *model_uniform, bh_model_info *info);
/* Checks every index trace_ray_model (ray.wgsl:287-363) would follow from the root of a ModelUniform blob: children inside
 * the node array and numbered after their parent, leaf ranges, lookup entries, point / normal indices.  The reference's
 * WGSL clamps out-of-range indices (naga's Restrict policy); this library refuses such a blob: bh_ctx_upload_models calls
 * this, bh_ctx_upload_models_async (the per-frame re-send) trusts its caller. */
int  bh_model_validate(const void *model_uniform);
/* points/normals: n*3 floats already in model space; tris: m*6 int32 (p1,p2,p3,n1,n2,n3) */
int  bh_model_from_arrays(const float *points, int32_t n_points, const float *normals, int32_t n_normals,
                          const int32_t *tris, int32_t n_tris, const float position[3], int32_t visible,
                          void *model_uniform, bh_model_info *info);
int  bh_model_build_bvh(void *model_uniform, int32_t triangle_count, bh_model_info *info);

/* ---- frame dump: the reference's "Save Image" (src/renderer/mod.rs:460-486): RGBA8 -> PNG, alpha
 *      forced to 255 when force_opaque != 0 (mod.rs:479 writes 255).  Pure host code. ------------------------------ */
int  bh_save_png(const char *path, const uint8_t *rgba8, uint32_t w, uint32_t h, int force_opaque);

/* ---- device math probe (tests only): evaluates the kernel's det-math on the device so that
 *      tests can compare it bit-for-bit with the oracle's contract flavour.
 *      fn: 0 pow(a,b) 1 pow5(a) 2 pow4(a) 3 sin 4 cos 5 tan 6 atan2(a,b) 7 acos ------------------ */
int  bh_ctx_math_probe(bh_ctx *ctx, int fn, const float *host_a, const float *host_b, float *host_out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* BH_ABI_H */
