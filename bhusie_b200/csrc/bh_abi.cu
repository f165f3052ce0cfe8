// bh_abi.cu — the C ABI of libbhray.so (include/bh_abi.h): context, ray pass and sky pass objects.
//
// Mirrors the reference's pass-object triple new / pass / output_view
// (src/renderer/pipelines/ray_pipeline.rs:36,297,301; sky_pipeline.rs:18,136,140) with plain
// pointers and sizes.  pass() only ENQUEUES work on the caller's CUDA stream, the same contract as
// recording into a wgpu ComputePass (execution starts at queue.submit, src/renderer/mod.rs:453).
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <set>

#include "bh_objects.h"

namespace bh {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (e == cudaErrorMemoryAllocation) ? BH_ERR_NOMEM
         : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) ? BH_ERR_NODEV
         : BH_ERR_CUDA;
}

static std::mutex g_ctx_mutex;
static std::set<const bh_ctx *> g_ctx_live;
bool ctx_alive(const bh_ctx *ctx)
{
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    return g_ctx_live.count(ctx) != 0;
}

}  // namespace bh

using namespace bh;

static_assert(sizeof(bh_camera_uniform) == 32, "CameraUniform is 32 bytes (camera.rs:66-73)");
static_assert(sizeof(bh_black_hole_uniform) == 132, "BlackHoleUniform is 132 bytes (blackhole.rs:37-51)");
static_assert(sizeof(bh_ray_details) == 32, "RayDetails is 32 bytes (ray_pipeline.rs:3-14)");

static uint32_t count_local_rows(uint32_t h, uint32_t band_rows, uint32_t rank, uint32_t n_ranks)
{
    if (n_ranks <= 1) return h;
    uint32_t rows = 0;
    const uint32_t bands = (h + band_rows - 1) / band_rows;
    for (uint32_t b = rank; b < bands; b += n_ranks) {
        const uint32_t begin = b * band_rows;
        const uint32_t end = begin + band_rows < h ? begin + band_rows : h;
        rows += end - begin;
    }
    return rows;
}

static int realloc_pipeline_buffers(bh_ray_pipeline *p)
{
    const size_t px = (size_t)p->local_rows * p->w;
    if (p->own_out_rows != p->local_rows) {
        if (p->own_out) cudaFree(p->own_out);
        if (p->queue) cudaFree(p->queue);
        p->own_out = nullptr; p->queue = nullptr;
        BH_CUDA(cudaMalloc(&p->own_out, (px ? px : 1) * sizeof(float4)));
        BH_CUDA(cudaMalloc(&p->queue, (px ? px : 1) * sizeof(unsigned)));
        p->own_out_rows = p->local_rows;
        if (p->aux_hit) { cudaFree(p->aux_hit); p->aux_hit = nullptr; }
        if (p->aux_steps) { cudaFree(p->aux_steps); p->aux_steps = nullptr; }
        if (p->aux_class) { cudaFree(p->aux_class); p->aux_class = nullptr; }
    }
    if ((p->aux_mask & BH_AUX_HIT) && !p->aux_hit) BH_CUDA(cudaMalloc(&p->aux_hit, (px ? px : 1) * sizeof(int32_t)));
    if ((p->aux_mask & BH_AUX_STEPS) && !p->aux_steps) BH_CUDA(cudaMalloc(&p->aux_steps, (px ? px : 1) * sizeof(uint32_t)));
    if ((p->aux_mask & BH_AUX_CLASS) && !p->aux_class) BH_CUDA(cudaMalloc(&p->aux_class, (px ? px : 1) * sizeof(uint8_t)));
    return BH_OK;
}

extern "C" {

int bh_abi_version(void) { return BH_ABI_VERSION; }
const char *bh_last_error(void) { return g_error; }

// ------------------------------------------------------------------------------------------ context
int bh_ctx_create(int cuda_device, bh_ctx **out)
{
    if (!out) { set_error("bh_ctx_create: out is NULL"); return BH_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("bh_ctx_create: no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        cudaGetLastError();
        return BH_ERR_NODEV;
    }
    if (cuda_device < 0 || cuda_device >= n) { set_error("bh_ctx_create: device %d out of range [0,%d)", cuda_device, n); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(cuda_device));
    cudaDeviceProp prop;
    BH_CUDA(cudaGetDeviceProperties(&prop, cuda_device));
    if (prop.major != 10) {
        set_error("bh_ctx_create: device %d is sm_%d%d; libbhray is built for sm_100a (B200) only", cuda_device, prop.major, prop.minor);
        return BH_ERR_NODEV;
    }
    bh_ctx *c = new (std::nothrow) bh_ctx();
    if (!c) { set_error("bh_ctx_create: out of host memory"); return BH_ERR_NOMEM; }
    c->device = cuda_device;
    c->sm_count = prop.multiProcessorCount;
    { std::lock_guard<std::mutex> lock(g_ctx_mutex); g_ctx_live.insert(c); }
    *out = c;
    return BH_OK;
}

void bh_ctx_destroy(bh_ctx *ctx)
{
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        if (!g_ctx_live.erase(ctx)) return;              // not (or no longer) one of ours
    }
    cudaSetDevice(ctx->device);
    for (auto &t : ctx->tex) if (t) cudaFree(t);
    if (ctx->models) cudaFree(ctx->models);
    if (ctx->async_err) cudaFreeHost(ctx->async_err);
    delete ctx;
}

int bh_ctx_set_numeric_mode(bh_ctx *ctx, bh_numeric_mode mode)
{
    if (!ctx || ((int)mode != BH_NUMERIC_LITERAL && (int)mode != BH_NUMERIC_FUSED)) { set_error("bh_ctx_set_numeric_mode: bad argument"); return BH_ERR_INVALID; }
    ctx->numeric_mode = (int)mode;
    return BH_OK;
}

int bh_ctx_get_numeric_mode(const bh_ctx *ctx) { return ctx ? ctx->numeric_mode : BH_ERR_INVALID; }

int bh_ctx_set_texture(bh_ctx *ctx, bh_texture_slot slot, const uint8_t *rgba8, uint32_t w, uint32_t h)
{
    if (!ctx || !rgba8 || (int)slot < 0 || (int)slot > 2 || w == 0 || h == 0 || w > 32768 || h > 32768) {
        set_error("bh_ctx_set_texture: bad argument (slot=%d, %ux%u)", (int)slot, w, h);
        return BH_ERR_INVALID;
    }
    BH_CUDA(cudaSetDevice(ctx->device));
    if (ctx->tex[slot]) { cudaFree(ctx->tex[slot]); ctx->tex[slot] = nullptr; }
    BH_CUDA(cudaMalloc(&ctx->tex[slot], (size_t)w * h * 4));
    BH_CUDA(cudaMemcpy(ctx->tex[slot], rgba8, (size_t)w * h * 4, cudaMemcpyHostToDevice));
    ctx->tex_w[slot] = (int)w; ctx->tex_h[slot] = (int)h;
    return BH_OK;
}

int bh_ctx_generate_disk_texture(bh_ctx *ctx, uint32_t w, uint32_t h, uint8_t *host_rgba8, int install)
{
    if (!ctx || w == 0 || h == 0 || w > 32768 || h > 32768) { set_error("bh_ctx_generate_disk_texture: bad argument"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(ctx->device));
    uchar4 *dev = nullptr;
    BH_CUDA(cudaMalloc(&dev, (size_t)w * h * 4));
    cudaError_t e = launch_disk_texture((int)w, (int)h, dev, nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    if (e == cudaSuccess && host_rgba8) e = cudaMemcpy(host_rgba8, dev, (size_t)w * h * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { cudaFree(dev); return cuda_fail(e, "bh_ctx_generate_disk_texture"); }
    if (install) {
        if (ctx->tex[BH_TEX_DISK]) cudaFree(ctx->tex[BH_TEX_DISK]);
        ctx->tex[BH_TEX_DISK] = dev; ctx->tex_w[BH_TEX_DISK] = (int)w; ctx->tex_h[BH_TEX_DISK] = (int)h;
    } else {
        cudaFree(dev);
    }
    return BH_OK;
}

static int upload_models(bh_ctx *ctx, const void *bytes, size_t nbytes, bool async, cudaStream_t stream)
{
    if (!ctx || (!bytes && nbytes) || nbytes % BH_MODEL_UNIFORM_SIZE != 0 || nbytes / BH_MODEL_UNIFORM_SIZE > BH_MAX_MODELS) {
        set_error("bh_ctx_upload_models: nbytes=%zu is not count*%d with count <= %d", nbytes, BH_MODEL_UNIFORM_SIZE, BH_MAX_MODELS);
        return BH_ERR_INVALID;
    }
    BH_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->models) BH_CUDA(cudaMalloc(&ctx->models, (size_t)BH_MAX_MODELS * kModelStride));
    const int count = (int)(nbytes / BH_MODEL_UNIFORM_SIZE);
    for (int i = 0; i < count; ++i) {
        const unsigned char *src = static_cast<const unsigned char *>(bytes) + (size_t)i * BH_MODEL_UNIFORM_SIZE;
        // the kernel follows every index of the blob unchecked (the reference's WGSL clamps them, naga Restrict): refuse a
        // malformed blob here.  The per-frame async re-send trusts its caller (bh_model_validate is exported for it).
        if (!async) { const int vrc = bh_model_validate(src); if (vrc != BH_OK) return vrc; }
        if (async) BH_CUDA(cudaMemcpyAsync(ctx->models + (size_t)i * kModelStride, src, BH_MODEL_UNIFORM_SIZE, cudaMemcpyHostToDevice, stream));
        else BH_CUDA(cudaMemcpy(ctx->models + (size_t)i * kModelStride, src, BH_MODEL_UNIFORM_SIZE, cudaMemcpyHostToDevice));
    }
    ctx->models_uploaded = count;
    return BH_OK;
}

int bh_ctx_upload_models(bh_ctx *ctx, const void *bytes, size_t nbytes) { return upload_models(ctx, bytes, nbytes, false, nullptr); }
int bh_ctx_upload_models_async(bh_ctx *ctx, const void *pinned_bytes, size_t nbytes, void *cuda_stream)
{
    return upload_models(ctx, pinned_bytes, nbytes, true, static_cast<cudaStream_t>(cuda_stream));
}

int bh_ctx_set_model_header(bh_ctx *ctx, uint32_t index, const float position[3], int32_t visible)
{
    if (!ctx || !position || (int)index >= ctx->models_uploaded) { set_error("bh_ctx_set_model_header: bad argument"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(ctx->device));
    unsigned char hdr[16];
    memcpy(hdr, position, 12);
    memcpy(hdr + 12, &visible, 4);
    BH_CUDA(cudaMemcpy(ctx->models + (size_t)index * kModelStride, hdr, 16, cudaMemcpyHostToDevice));
    return BH_OK;
}

// ------------------------------------------------------------------------------------------ ray pass
int bh_ray_pipeline_create(bh_ctx *ctx, uint32_t width, uint32_t height, const bh_ray_pipeline *prev, bh_ray_pipeline **out)
{
    if (!out) { set_error("bh_ray_pipeline_create: out is NULL"); return BH_ERR_INVALID; }
    *out = nullptr;
    if (!ctx || width < 2 || height < 2 || width > 65535 || height > 65535) {
        set_error("bh_ray_pipeline_create: resolution %ux%u out of range [2,65535]", width, height);
        return BH_ERR_INVALID;
    }
    if (prev && (prev->ctx != ctx || prev->w < 2 || prev->h < 2)) { set_error("bh_ray_pipeline_create: prev belongs to another context"); return BH_ERR_INVALID; }
    if (prev && prev->n_ranks != 1) { set_error("bh_ray_pipeline_create: prev level must be untiled (coarse levels are replicated)"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(ctx->device));
    bh_ray_pipeline *p = new (std::nothrow) bh_ray_pipeline();
    if (!p) { set_error("bh_ray_pipeline_create: out of host memory"); return BH_ERR_NOMEM; }
    p->ctx = ctx; p->w = width; p->h = height; p->prev = prev;
    p->band_rows = height; p->rank = 0; p->n_ranks = 1; p->local_rows = height;
    int rc = realloc_pipeline_buffers(p);
    if (rc == BH_OK) {
        cudaError_t e = cudaMalloc(&p->stats, sizeof(unsigned long long) * kStatCount);
        if (e == cudaSuccess) e = cudaMalloc(&p->work, sizeof(unsigned) * kWorkCount);
        if (e == cudaSuccess) e = cudaMemset(p->stats, 0, sizeof(unsigned long long) * kStatCount);
        if (e != cudaSuccess) rc = cuda_fail(e, "bh_ray_pipeline_create: cudaMalloc");
    }
    if (rc != BH_OK) { bh_ray_pipeline_destroy(p); return rc; }
    *out = p;
    return BH_OK;
}

void bh_ray_pipeline_destroy(bh_ray_pipeline *p)
{
    if (!p) return;
    if (!ctx_alive(p->ctx)) { delete p; return; }        // context destroyed first: nothing of it may be touched
    cudaSetDevice(p->ctx->device);
    if (p->ran) cudaStreamSynchronize(p->last_stream);
    if (p->copy_stream) {
        cudaStreamSynchronize(p->copy_stream);
        for (auto &e : p->chunk_done) if (e) cudaEventDestroy(e);
        if (p->copy_done) cudaEventDestroy(p->copy_done);
        cudaStreamDestroy(p->copy_stream);
    }
    for (void *q : { (void *)p->own_out, (void *)p->aux_hit, (void *)p->aux_steps, (void *)p->aux_class, (void *)p->stats, (void *)p->work, (void *)p->queue })
        if (q) cudaFree(q);
    delete p;
}

int bh_ray_pipeline_set_tiling(bh_ray_pipeline *p, uint32_t band_rows, uint32_t rank, uint32_t n_ranks)
{
    if (!p || band_rows == 0 || n_ranks == 0 || rank >= n_ranks) { set_error("bh_ray_pipeline_set_tiling: bad argument"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(p->ctx->device));
    if (p->ran) BH_CUDA(cudaStreamSynchronize(p->last_stream));
    p->band_rows = band_rows; p->rank = rank; p->n_ranks = n_ranks;
    p->local_rows = count_local_rows(p->h, band_rows, rank, n_ranks);
    return realloc_pipeline_buffers(p);
}

uint32_t bh_ray_pipeline_local_rows(const bh_ray_pipeline *p) { return p ? p->local_rows : 0; }
uint32_t bh_ray_pipeline_width(const bh_ray_pipeline *p) { return p ? p->w : 0; }
uint32_t bh_ray_pipeline_height(const bh_ray_pipeline *p) { return p ? p->h : 0; }

int bh_ray_pipeline_enable_aux(bh_ray_pipeline *p, uint32_t aux_mask)
{
    if (!p || (aux_mask & ~7u)) { set_error("bh_ray_pipeline_enable_aux: bad argument"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(p->ctx->device));
    if (p->ran) BH_CUDA(cudaStreamSynchronize(p->last_stream));
    p->aux_mask = aux_mask;
    if (!(aux_mask & BH_AUX_HIT) && p->aux_hit) { cudaFree(p->aux_hit); p->aux_hit = nullptr; }
    if (!(aux_mask & BH_AUX_STEPS) && p->aux_steps) { cudaFree(p->aux_steps); p->aux_steps = nullptr; }
    if (!(aux_mask & BH_AUX_CLASS) && p->aux_class) { cudaFree(p->aux_class); p->aux_class = nullptr; }
    return realloc_pipeline_buffers(p);
}

int bh_ray_pipeline_bind_output(bh_ray_pipeline *p, void *device_rgba32f)
{
    if (!p || ((uintptr_t)device_rgba32f & 15u)) { set_error("bh_ray_pipeline_bind_output: pointer must be 16-byte aligned"); return BH_ERR_INVALID; }
    p->bound_out = static_cast<float4 *>(device_rgba32f);
    return BH_OK;
}

}  // extern "C"

int bh::build_pass_params(bh_ray_pipeline *p, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                             const bh_ray_details *details, const char *who, PassParams &P)
{
    if (!p || !camera || !black_hole || !details) { set_error("%s: NULL argument", who); return BH_ERR_INVALID; }
    bh_ctx *c = p->ctx;
    if (!c->tex[0] || !c->tex[1] || !c->tex[2]) { set_error("%s: textures not set (bh_ctx_set_texture for COLOR, DISK and SKY)", who); return BH_ERR_STATE; }
    if (details->model_count < 0 || details->model_count > c->models_uploaded) {
        set_error("%s: model_count=%d but %d model(s) uploaded", who, details->model_count, c->models_uploaded);
        return BH_ERR_INVALID;
    }
    if (p->prev && p->prev->host_only) {
        set_error("%s: the previous level's last pass wrote its output to host memory only (bh_ray_pipeline_pass_to_host, n_chunks=0)", who);
        return BH_ERR_STATE;
    }
    memset(&P, 0, sizeof P);
    P.cam = *camera; P.hole = *black_hole; P.det = *details;
    if (P.det.integration_method != 0) P.det.integration_method = 1;        // ray.wgsl:525: anything non-zero takes the RK branch
    P.color = DevTexture{ c->tex[0], c->tex_w[0], c->tex_h[0] };
    P.disk = DevTexture{ c->tex[1], c->tex_w[1], c->tex_h[1] };
    P.sky = DevTexture{ c->tex[2], c->tex_w[2], c->tex_h[2] };
    P.models = c->models;
    P.out = p->out();
    P.out_global_rows = p->bound_frame ? 1 : 0;
    P.prev = p->prev ? p->prev->out() : nullptr;
    P.w = (int)p->w; P.h = (int)p->h;
    P.pw = p->prev ? (int)p->prev->w : 1; P.ph = p->prev ? (int)p->prev->h : 1;
    P.band_rows = (int)p->band_rows; P.rank = (int)p->rank; P.n_ranks = (int)p->n_ranks; P.local_rows = (int)p->local_rows;
    P.aux_hit = p->aux_hit; P.aux_steps = p->aux_steps; P.aux_class = p->aux_class;
    P.stats = p->stats; P.work = p->work; P.queue = p->queue;
    P.tiles_x = (int)((p->w + 7) / 8);
    P.tile_rows = 4;
    P.item_begin = 0;
    P.n_items = (unsigned)P.tiles_x * (unsigned)((p->local_rows + 3) / 4);
    derive_pass_constants(P);
    return BH_OK;
}

extern "C" {

int bh_ray_pipeline_bind_frame(bh_ray_pipeline *p, void *device_frame_rgba32f)
{
    if (!p || ((uintptr_t)device_frame_rgba32f & 15u)) { set_error("bh_ray_pipeline_bind_frame: pointer must be 16-byte aligned"); return BH_ERR_INVALID; }
    p->bound_frame = static_cast<float4 *>(device_frame_rgba32f);
    return BH_OK;
}

// ---- frames shared between the processes of one node (CUDA IPC): rank 0 owns, the others map it over NVLink
int bh_shared_frame_create(bh_ctx *ctx, size_t nbytes, void **device_ptr, uint8_t handle_out[64])
{
    if (!ctx || !device_ptr || !handle_out || nbytes == 0) { set_error("bh_shared_frame_create: bad argument"); return BH_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    BH_CUDA(cudaSetDevice(ctx->device));
    void *ptr = nullptr;
    // the frame, then BH_SHARED_FLAGS 32-bit flags at the next 256-byte boundary (bh_shared_frame_flags), zeroed
    const size_t flag_off = (nbytes + 255) / 256 * 256;
    BH_CUDA(cudaMalloc(&ptr, flag_off + BH_SHARED_FLAGS * sizeof(uint32_t)));
    cudaError_t e = cudaMemset(static_cast<unsigned char *>(ptr) + flag_off, 0, BH_SHARED_FLAGS * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(ptr); return cuda_fail(e, "bh_shared_frame_create: cudaMemset"); }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { cudaFree(ptr); return cuda_fail(e, "cudaIpcGetMemHandle"); }
    memcpy(handle_out, &h, 64);
    *device_ptr = ptr;
    return BH_OK;
}

int bh_shared_frame_open(bh_ctx *ctx, const uint8_t handle[64], void **device_ptr)
{
    if (!ctx || !handle || !device_ptr) { set_error("bh_shared_frame_open: bad argument"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    BH_CUDA(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BH_OK;
}

int bh_shared_frame_release(bh_ctx *ctx, void *device_ptr, int owner)
{
    if (!ctx || !device_ptr) { set_error("bh_shared_frame_release: bad argument"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(ctx->device));
    if (owner) BH_CUDA(cudaFree(device_ptr));
    else BH_CUDA(cudaIpcCloseMemHandle(device_ptr));
    return BH_OK;
}

int bh_ray_pipeline_pass(bh_ray_pipeline *p, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                         const bh_ray_details *details, void *cuda_stream)
{
    PassParams P;
    const int rc = build_pass_params(p, camera, black_hole, details, "bh_ray_pipeline_pass", P);
    if (rc != BH_OK) return rc;
    bh_ctx *c = p->ctx;
    BH_CUDA(cudaSetDevice(c->device));
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    LaunchConfig cfg{ c->sm_count, c->numeric_mode };
    p->last_stream = stream; p->ran = true; p->host_only = false;
    if (p->local_rows == 0) return BH_OK;
    BH_CUDA(launch_ray_pass(P, cfg, stream));
    return BH_OK;
}

int bh_ray_pipeline_pass_to_host(bh_ray_pipeline *p, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                                 const bh_ray_details *details, float *pinned_host_rgba32f, uint32_t n_chunks, void *cuda_stream)
{
    PassParams P;
    const int rc = build_pass_params(p, camera, black_hole, details, "bh_ray_pipeline_pass_to_host", P);
    if (rc != BH_OK) return rc;
    if (!pinned_host_rgba32f) { set_error("bh_ray_pipeline_pass_to_host: host buffer is NULL"); return BH_ERR_INVALID; }
    if (p->bound_frame) { set_error("bh_ray_pipeline_pass_to_host: output is bound to an external frame (bh_ray_pipeline_bind_frame)"); return BH_ERR_STATE; }
    if (n_chunks > 16) n_chunks = 16;
    bh_ctx *c = p->ctx;
    BH_CUDA(cudaSetDevice(c->device));
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    if (n_chunks == 0) {
        // zero-copy: the kernel's 16-byte pixel stores go straight to the caller's page-locked buffer over PCIe while it
        // is still tracing (a 4K frame is 133 MB per ~20 ms = 6.5 GB/s, a fraction of the link), so there is no D2H step
        void *mapped = nullptr;
        cudaError_t e = cudaHostGetDevicePointer(&mapped, pinned_host_rgba32f, 0);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("bh_ray_pipeline_pass_to_host: n_chunks=0 needs page-locked, mapped host memory (%s)", cudaGetErrorString(e)); return BH_ERR_INVALID; }
        P.out = static_cast<float4 *>(mapped);
        LaunchConfig cfg0{ c->sm_count, c->numeric_mode };
        p->last_stream = stream; p->ran = true; p->host_only = true;
        if (p->local_rows == 0) return BH_OK;
        BH_CUDA(launch_ray_pass(P, cfg0, stream));
        return BH_OK;
    }
    if (!p->copy_stream) {
        BH_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
        for (auto &e : p->chunk_done) BH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        BH_CUDA(cudaEventCreateWithFlags(&p->copy_done, cudaEventDisableTiming));
    }
    LaunchConfig cfg{ c->sm_count, c->numeric_mode };
    p->last_stream = stream; p->ran = true; p->host_only = false;
    if (p->local_rows == 0) return BH_OK;
    const size_t row_bytes = (size_t)p->w * sizeof(float4);
    if (P.prev != nullptr) n_chunks = 1;                 // fine levels: classification + queue cover the whole level at once
    if (n_chunks == 1) {
        BH_CUDA(launch_ray_pass(P, cfg, stream));
        BH_CUDA(cudaEventRecord(p->chunk_done[0], stream));
        BH_CUDA(cudaStreamWaitEvent(p->copy_stream, p->chunk_done[0], 0));
        BH_CUDA(cudaMemcpyAsync(pinned_host_rgba32f, p->out(), row_bytes * p->local_rows, cudaMemcpyDeviceToHost, p->copy_stream));
    } else {
        // base level: tiles are ordered row-major, so a tile-row range is a contiguous band of output rows.  Each band is a
        // separate launch; its D2H copy runs on the copy stream while the next band is traced.
        const unsigned tile_rows = (p->local_rows + 3) / 4;
        bool first = true;
        for (uint32_t k = 0; k < n_chunks; ++k) {
            const unsigned tr0 = (unsigned)((uint64_t)tile_rows * k / n_chunks), tr1 = (unsigned)((uint64_t)tile_rows * (k + 1) / n_chunks);
            if (tr1 == tr0) continue;
            PassParams Q = P;
            Q.item_begin = tr0 * (unsigned)P.tiles_x;
            Q.n_items = tr1 * (unsigned)P.tiles_x;
            BH_CUDA(launch_trace_range(Q, cfg, first, stream));
            first = false;
            BH_CUDA(cudaEventRecord(p->chunk_done[k], stream));
            BH_CUDA(cudaStreamWaitEvent(p->copy_stream, p->chunk_done[k], 0));
            const size_t r0 = (size_t)tr0 * 4, r1 = (size_t)tr1 * 4 < p->local_rows ? (size_t)tr1 * 4 : p->local_rows;
            BH_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(pinned_host_rgba32f) + r0 * row_bytes,
                                    reinterpret_cast<const char *>(p->out()) + r0 * row_bytes, (r1 - r0) * row_bytes,
                                    cudaMemcpyDeviceToHost, p->copy_stream));
        }
    }
    BH_CUDA(cudaEventRecord(p->copy_done, p->copy_stream));
    BH_CUDA(cudaStreamWaitEvent(stream, p->copy_done, 0));          // work enqueued later on `stream` sees the copy finished
    p->copy_pending = true;
    return BH_OK;
}

int bh_ray_pipeline_sync(bh_ray_pipeline *p)
{
    if (!p) { set_error("bh_ray_pipeline_sync: NULL pipeline"); return BH_ERR_INVALID; }
    if (!p->ran) return BH_OK;
    BH_CUDA(cudaSetDevice(p->ctx->device));
    BH_CUDA(cudaStreamSynchronize(p->last_stream));
    if (p->copy_pending) { BH_CUDA(cudaStreamSynchronize(p->copy_stream)); p->copy_pending = false; }
    return BH_OK;
}

const float *bh_ray_pipeline_output(const bh_ray_pipeline *p) { return p ? reinterpret_cast<const float *>(p->out()) : nullptr; }

int bh_ray_pipeline_read(bh_ray_pipeline *p, float *host_rgba32f, int32_t *host_hit, uint32_t *host_steps, uint8_t *host_class)
{
    if (!p) { set_error("bh_ray_pipeline_read: NULL pipeline"); return BH_ERR_INVALID; }
    if (!p->ran) { set_error("bh_ray_pipeline_read: no pass has been enqueued"); return BH_ERR_STATE; }
    if ((host_hit && !p->aux_hit) || (host_steps && !p->aux_steps) || (host_class && !p->aux_class)) {
        set_error("bh_ray_pipeline_read: aux buffer requested but not enabled (bh_ray_pipeline_enable_aux)");
        return BH_ERR_STATE;
    }
    if (host_rgba32f && p->host_only) { set_error("bh_ray_pipeline_read: the last pass wrote its RGBA to host memory only (bh_ray_pipeline_pass_to_host, n_chunks=0)"); return BH_ERR_STATE; }
    if (host_rgba32f && p->bound_frame) { set_error("bh_ray_pipeline_read: output is bound to an external frame (bh_ray_pipeline_bind_frame); read the frame instead"); return BH_ERR_STATE; }
    BH_CUDA(cudaSetDevice(p->ctx->device));
    BH_CUDA(cudaStreamSynchronize(p->last_stream));
    const size_t px = (size_t)p->local_rows * p->w;
    if (host_rgba32f) BH_CUDA(cudaMemcpy(host_rgba32f, p->out(), px * sizeof(float4), cudaMemcpyDeviceToHost));
    if (host_hit) BH_CUDA(cudaMemcpy(host_hit, p->aux_hit, px * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (host_steps) BH_CUDA(cudaMemcpy(host_steps, p->aux_steps, px * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (host_class) BH_CUDA(cudaMemcpy(host_class, p->aux_class, px * sizeof(uint8_t), cudaMemcpyDeviceToHost));
    return BH_OK;
}

int bh_ray_pipeline_stats(bh_ray_pipeline *p, bh_pass_stats *out)
{
    if (!p || !out) { set_error("bh_ray_pipeline_stats: NULL argument"); return BH_ERR_INVALID; }
    if (!p->ran) { set_error("bh_ray_pipeline_stats: no pass has been enqueued"); return BH_ERR_STATE; }
    BH_CUDA(cudaSetDevice(p->ctx->device));
    BH_CUDA(cudaStreamSynchronize(p->last_stream));
    unsigned long long s[kStatCount];
    BH_CUDA(cudaMemcpy(s, p->stats, sizeof s, cudaMemcpyDeviceToHost));
    out->ray_steps = s[kStatSteps]; out->px_traced = s[kStatTraced]; out->px_copied = s[kStatCopied]; out->px_interp = s[kStatInterp];
    out->node_visits = s[kStatNodeVisits]; out->tri_tests = s[kStatTriTests]; out->tex_samples = s[kStatTexSamples];
    out->rk_reject = s[kStatRkReject]; out->stack_overflow = s[kStatStackOverflow];
    if (out->rk_reject) {
        set_error("bh_ray_pipeline_stats: %llu RK step(s) had an error norm > 1; the reference's accept loop (ray.wgsl:425-451) would not terminate there",
                  (unsigned long long)out->rk_reject);
        return BH_ERR_NUMERIC;
    }
    return BH_OK;
}

// ------------------------------------------------------------------------------------------ sky pass
int bh_sky_pipeline_create(bh_ctx *ctx, const bh_ray_pipeline *prev, bh_sky_format format, bh_sky_pipeline **out)
{
    if (!out) { set_error("bh_sky_pipeline_create: out is NULL"); return BH_ERR_INVALID; }
    *out = nullptr;
    if (!ctx || !prev || prev->ctx != ctx || ((int)format != 0 && (int)format != 1)) { set_error("bh_sky_pipeline_create: bad argument"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(ctx->device));
    bh_sky_pipeline *s = new (std::nothrow) bh_sky_pipeline();
    if (!s) { set_error("bh_sky_pipeline_create: out of host memory"); return BH_ERR_NOMEM; }
    s->ctx = ctx; s->prev = prev; s->format = format;
    // sized for the whole frame so that a later set_tiling on prev cannot outgrow it
    cudaError_t e = cudaMalloc(&s->own_out, (size_t)prev->w * prev->h * s->texel_bytes());
    if (e == cudaSuccess) e = cudaMalloc(&s->stats, sizeof(unsigned long long) * kStatCount);
    if (e == cudaSuccess) e = cudaMemset(s->stats, 0, sizeof(unsigned long long) * kStatCount);
    if (e != cudaSuccess) { const int rc = cuda_fail(e, "bh_sky_pipeline_create: cudaMalloc"); bh_sky_pipeline_destroy(s); return rc; }
    *out = s;
    return BH_OK;
}

int bh_sky_pipeline_create_for_frame(bh_ctx *ctx, const void *device_frame_rgba32f, uint32_t width, uint32_t height,
                                     bh_sky_format format, bh_sky_pipeline **out)
{
    if (!out) { set_error("bh_sky_pipeline_create_for_frame: out is NULL"); return BH_ERR_INVALID; }
    *out = nullptr;
    if (!ctx || !device_frame_rgba32f || ((uintptr_t)device_frame_rgba32f & 15u) || width == 0 || height == 0 || width > 65535 ||
        height > 65535 || ((int)format != 0 && (int)format != 1)) {
        set_error("bh_sky_pipeline_create_for_frame: bad argument");
        return BH_ERR_INVALID;
    }
    BH_CUDA(cudaSetDevice(ctx->device));
    bh_sky_pipeline *s = new (std::nothrow) bh_sky_pipeline();
    if (!s) { set_error("bh_sky_pipeline_create_for_frame: out of host memory"); return BH_ERR_NOMEM; }
    s->ctx = ctx; s->format = format;
    s->raw_prev = static_cast<const float4 *>(device_frame_rgba32f); s->raw_w = width; s->raw_h = height;
    cudaError_t e = cudaMalloc(&s->own_out, (size_t)width * height * s->texel_bytes());
    if (e == cudaSuccess) e = cudaMalloc(&s->stats, sizeof(unsigned long long) * kStatCount);
    if (e == cudaSuccess) e = cudaMemset(s->stats, 0, sizeof(unsigned long long) * kStatCount);
    if (e != cudaSuccess) { const int rc = cuda_fail(e, "bh_sky_pipeline_create_for_frame: cudaMalloc"); bh_sky_pipeline_destroy(s); return rc; }
    *out = s;
    return BH_OK;
}

void bh_sky_pipeline_destroy(bh_sky_pipeline *s)
{
    if (!s) return;
    if (!ctx_alive(s->ctx)) { delete s; return; }
    cudaSetDevice(s->ctx->device);
    if (s->ran) cudaStreamSynchronize(s->last_stream);
    if (s->own_out) cudaFree(s->own_out);
    if (s->stats) cudaFree(s->stats);
    delete s;
}

int bh_sky_pipeline_bind_output(bh_sky_pipeline *s, void *device_rgba)
{
    if (!s || ((uintptr_t)device_rgba & 15u)) { set_error("bh_sky_pipeline_bind_output: pointer must be 16-byte aligned"); return BH_ERR_INVALID; }
    s->bound_out = device_rgba;
    return BH_OK;
}

int bh_sky_pipeline_pass(bh_sky_pipeline *s, void *cuda_stream)
{
    if (!s) { set_error("bh_sky_pipeline_pass: NULL pipeline"); return BH_ERR_INVALID; }
    bh_ctx *c = s->ctx;
    if (!c->tex[2]) { set_error("bh_sky_pipeline_pass: sky texture not set"); return BH_ERR_STATE; }
    if (s->prev && s->prev->host_only) { set_error("bh_sky_pipeline_pass: the ray level's last pass wrote its output to host memory only (pass_to_host, n_chunks=0)"); return BH_ERR_STATE; }
    if (s->prev && s->prev->bound_frame && s->prev->n_ranks > 1) {
        // prev->out() is then a whole (possibly remote) frame addressed by global row, not this rank's bands
        set_error("bh_sky_pipeline_pass: the ray level renders tiled into a bound frame; resolve the assembled frame instead (bh_frame_multi does)");
        return BH_ERR_STATE;
    }
    BH_CUDA(cudaSetDevice(c->device));
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    SkyParams S;
    memset(&S, 0, sizeof S);
    S.sky = DevTexture{ c->tex[2], c->tex_w[2], c->tex_h[2] };
    S.prev = s->prev ? s->prev->out() : s->raw_prev;
    S.out = s->out();
    S.n_pixels = (int)s->pixels();
    S.format = (int)s->format;
    S.stats = s->stats;
    s->last_stream = stream; s->ran = true;
    if (S.n_pixels == 0) return BH_OK;
    BH_CUDA(cudaMemsetAsync(s->stats, 0, sizeof(unsigned long long) * kStatCount, stream));
    LaunchConfig cfg{ c->sm_count, c->numeric_mode };
    BH_CUDA(launch_sky_pass(S, cfg, stream));
    return BH_OK;
}

const void *bh_sky_pipeline_output(const bh_sky_pipeline *s) { return s ? s->out() : nullptr; }

int bh_sky_pipeline_read(bh_sky_pipeline *s, void *host_rgba)
{
    if (!s || !host_rgba) { set_error("bh_sky_pipeline_read: NULL argument"); return BH_ERR_INVALID; }
    if (!s->ran) { set_error("bh_sky_pipeline_read: no pass has been enqueued"); return BH_ERR_STATE; }
    if (s->prev && s->prev->bound_frame && s->prev->n_ranks > 1) { set_error("bh_sky_pipeline_read: the ray level renders tiled into a bound frame"); return BH_ERR_STATE; }
    BH_CUDA(cudaSetDevice(s->ctx->device));
    BH_CUDA(cudaStreamSynchronize(s->last_stream));
    BH_CUDA(cudaMemcpy(host_rgba, s->out(), s->pixels() * s->texel_bytes(), cudaMemcpyDeviceToHost));
    return BH_OK;
}

// ------------------------------------------------------------------------------------------ post chain
struct bh_post_pass {
    bh_ctx *ctx = nullptr;
    int kind = 0;
    uint32_t w = 0, h = 0, in_w = 0, in_h = 0;
    const void *in1 = nullptr, *in2 = nullptr;
    void *out = nullptr;
    cudaStream_t last_stream = nullptr;
    bool ran = false;
    size_t out_bytes() const { return (size_t)w * h * (kind == BH_POST_FXAA ? 4 : 8); }
};

int bh_post_pass_create(bh_ctx *ctx, bh_post_kind kind, uint32_t out_w, uint32_t out_h, const void *in1_device, uint32_t in1_w,
                        uint32_t in1_h, const void *in2_device, bh_post_pass **out)
{
    if (!out) { set_error("bh_post_pass_create: out is NULL"); return BH_ERR_INVALID; }
    *out = nullptr;
    if (!ctx || (int)kind < 0 || (int)kind > BH_POST_FXAA || !in1_device || out_w == 0 || out_h == 0 || in1_w == 0 || in1_h == 0 ||
        out_w > 65535 || out_h > 65535 || ((uintptr_t)in1_device & 7u) || ((uintptr_t)in2_device & 7u)) {
        set_error("bh_post_pass_create: bad argument");
        return BH_ERR_INVALID;
    }
    if (kind == BH_POST_MIX && !in2_device) { set_error("bh_post_pass_create: MIX needs two inputs"); return BH_ERR_INVALID; }
    if (kind >= BH_POST_MIX && (in1_w != out_w || in1_h != out_h)) { set_error("bh_post_pass_create: mix/hdr/fxaa run at their input's resolution"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(ctx->device));
    bh_post_pass *p = new (std::nothrow) bh_post_pass();
    if (!p) { set_error("bh_post_pass_create: out of host memory"); return BH_ERR_NOMEM; }
    p->ctx = ctx; p->kind = (int)kind; p->w = out_w; p->h = out_h; p->in_w = in1_w; p->in_h = in1_h; p->in1 = in1_device; p->in2 = in2_device;
    cudaError_t e = cudaMalloc(&p->out, p->out_bytes());
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "bh_post_pass_create: cudaMalloc"); }
    *out = p;
    return BH_OK;
}

void bh_post_pass_destroy(bh_post_pass *p)
{
    if (!p) return;
    if (!ctx_alive(p->ctx)) { delete p; return; }
    cudaSetDevice(p->ctx->device);
    if (p->ran) cudaStreamSynchronize(p->last_stream);
    if (p->out) cudaFree(p->out);
    delete p;
}

int bh_post_pass_run(bh_post_pass *p, const void *details, void *cuda_stream)
{
    if (!p) { set_error("bh_post_pass_run: NULL pass"); return BH_ERR_INVALID; }
    if ((p->kind == BH_POST_MIX || p->kind == BH_POST_FXAA) && !details) { set_error("bh_post_pass_run: this pass needs its details uniform"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(p->ctx->device));
    PostParams P;
    memset(&P, 0, sizeof P);
    P.in1 = HalfImage{ static_cast<const uint2 *>(p->in1), (int)p->in_w, (int)p->in_h };
    P.in2 = HalfImage{ static_cast<const uint2 *>(p->in2), (int)p->in_w, (int)p->in_h };
    P.out = p->out; P.w = (int)p->w; P.h = (int)p->h;
    if (p->kind == BH_POST_MIX) P.mix_ratio = static_cast<const bh_mix_details *>(details)->mix_ratio;
    if (p->kind == BH_POST_FXAA) {
        const bh_fxaa_details *d = static_cast<const bh_fxaa_details *>(details);
        P.edge_min = d->edge_threshold_min; P.edge_max = d->edge_threshold_max; P.iterations = d->iterations; P.subpix = d->subpixel_quality;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    LaunchConfig cfg{ p->ctx->sm_count, p->ctx->numeric_mode };
    p->last_stream = stream; p->ran = true;
    BH_CUDA(launch_post_pass(p->kind, P, cfg, stream));
    return BH_OK;
}

const void *bh_post_pass_output(const bh_post_pass *p) { return p ? p->out : nullptr; }

int bh_post_pass_read(bh_post_pass *p, void *host)
{
    if (!p || !host) { set_error("bh_post_pass_read: NULL argument"); return BH_ERR_INVALID; }
    if (!p->ran) { set_error("bh_post_pass_read: the pass has not run"); return BH_ERR_STATE; }
    BH_CUDA(cudaSetDevice(p->ctx->device));
    BH_CUDA(cudaStreamSynchronize(p->last_stream));
    BH_CUDA(cudaMemcpy(host, p->out, p->out_bytes(), cudaMemcpyDeviceToHost));
    return BH_OK;
}

// ------------------------------------------------------------------------------------------ math probe
int bh_ctx_math_probe(bh_ctx *ctx, int fn, const float *host_a, const float *host_b, float *host_out, size_t n)
{
    if (!ctx || !host_a || !host_out || fn < 0 || fn > 7) { set_error("bh_ctx_math_probe: bad argument"); return BH_ERR_INVALID; }
    if (n == 0) return BH_OK;
    BH_CUDA(cudaSetDevice(ctx->device));
    float *a = nullptr, *b = nullptr, *o = nullptr;
    cudaError_t e = cudaMalloc(&a, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&b, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&o, n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(a, host_a, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = host_b ? cudaMemcpy(b, host_b, n * 4, cudaMemcpyHostToDevice) : cudaMemset(b, 0, n * 4);
    if (e == cudaSuccess) e = launch_math_probe(fn, a, b, o, n, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(host_out, o, n * 4, cudaMemcpyDeviceToHost);
    if (a) cudaFree(a);
    if (b) cudaFree(b);
    if (o) cudaFree(o);
    if (e != cudaSuccess) return cuda_fail(e, "bh_ctx_math_probe");
    return BH_OK;
}

}  // extern "C"
