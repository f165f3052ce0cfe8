// bh_multi.cu — the multi-GPU part of the C ABI (include/bh_abi.h, "multi-GPU").
//
// The reference is single-GPU (one wgpu queue, src/renderer/mod.rs:63-89,453).  Pixels of a level are independent
// (ray.wgsl:167-243), so a frame shards by cyclic row bands with no data-path collective: every device's ray kernel stores
// its finished pixels straight into one frame — device 0's memory over NVLink, or the caller's page-locked host frame over
// the device's own PCIe link — and what is left to do is ORDERING, which is done with CUDA events inside one process
// (bh_frame_multi) and with stream-ordered flag stores / waits in peer memory between processes (bh_stream_signal / _wait).
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <new>
#include <vector>

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "bh_objects.h"

using namespace bh;

// ------------------------------------------------------------------------------------------------ stream-ordered flags
namespace {

__global__ void signal_kernel(uint32_t *flag, uint32_t value)
{
    // everything this stream wrote before (the ray kernel's pixel stores, possibly to a peer GPU) is ordered before the flag
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void wait_kernel(const uint32_t *flags, uint32_t n, uint32_t value, unsigned long long timeout_ns, unsigned *gave_up)
{
    const unsigned i = threadIdx.x;
    if (i >= n) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        if ((int32_t)(v - value) >= 0) break;                       // wrap-safe "v has reached value"
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) { atomicExch(gave_up, 1u); break; }   // never hang the GPU on a peer that died
        __nanosleep(256);
    }
}

int ensure_async_word(bh_ctx *ctx)
{
    if (ctx->async_err) return BH_OK;
    BH_CUDA(cudaSetDevice(ctx->device));
    BH_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->async_err), sizeof(unsigned), cudaHostAllocMapped));
    *ctx->async_err = 0u;
    return BH_OK;
}

size_t align_up(size_t n, size_t a) { return (n + a - 1) / a * a; }

}  // namespace

extern "C" {

uint32_t *bh_shared_frame_flags(void *device_ptr, size_t nbytes)
{
    if (!device_ptr) return nullptr;
    return reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(device_ptr) + align_up(nbytes, 256));
}

int bh_stream_signal(bh_ctx *ctx, uint32_t *device_flag, uint32_t value, void *cuda_stream)
{
    if (!ctx || !device_flag) { set_error("bh_stream_signal: NULL argument"); return BH_ERR_INVALID; }
    BH_CUDA(cudaSetDevice(ctx->device));
    signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(cuda_stream)>>>(device_flag, value);
    BH_CUDA(cudaGetLastError());
    return BH_OK;
}

int bh_stream_wait(bh_ctx *ctx, const uint32_t *device_flags, uint32_t n_flags, uint32_t value, uint32_t timeout_ms, void *cuda_stream)
{
    if (!ctx || !device_flags || n_flags == 0 || n_flags > BH_SHARED_FLAGS) { set_error("bh_stream_wait: bad argument"); return BH_ERR_INVALID; }
    const int rc = ensure_async_word(ctx);
    if (rc != BH_OK) return rc;
    BH_CUDA(cudaSetDevice(ctx->device));
    unsigned *dev_word = nullptr;
    BH_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&dev_word), ctx->async_err, 0));
    wait_kernel<<<1, BH_SHARED_FLAGS, 0, static_cast<cudaStream_t>(cuda_stream)>>>(device_flags, n_flags, value,
                                                                                   (unsigned long long)timeout_ms * 1000000ull, dev_word);
    BH_CUDA(cudaGetLastError());
    return BH_OK;
}

int bh_ctx_check_async(bh_ctx *ctx)
{
    if (!ctx) { set_error("bh_ctx_check_async: NULL context"); return BH_ERR_INVALID; }
    if (!ctx->async_err) return BH_OK;
    if (__atomic_exchange_n(ctx->async_err, 0u, __ATOMIC_ACQ_REL) != 0u) {
        set_error("bh_ctx_check_async: a cross-GPU wait on device %d gave up (a peer never signalled)", ctx->device);
        return BH_ERR_TIMEOUT;
    }
    return BH_OK;
}

// ------------------------------------------------------------------------------------------------ tiled pass into a host frame
int bh_ray_pipeline_pass_to_host_frame(bh_ray_pipeline *p, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                                       const bh_ray_details *details, float *mapped_host_frame_rgba32f, void *cuda_stream)
{
    PassParams P;
    const int rc = build_pass_params(p, camera, black_hole, details, "bh_ray_pipeline_pass_to_host_frame", P);
    if (rc != BH_OK) return rc;
    if (!mapped_host_frame_rgba32f) { set_error("bh_ray_pipeline_pass_to_host_frame: host frame is NULL"); return BH_ERR_INVALID; }
    bh_ctx *c = p->ctx;
    BH_CUDA(cudaSetDevice(c->device));
    void *mapped = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&mapped, mapped_host_frame_rgba32f, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("bh_ray_pipeline_pass_to_host_frame: needs page-locked, mapped host memory (%s)", cudaGetErrorString(e));
        return BH_ERR_INVALID;
    }
    P.out = static_cast<float4 *>(mapped);
    P.out_global_rows = 1;                    // rows land at their global index of the full frame
    LaunchConfig cfg{ c->sm_count, c->numeric_mode };
    p->last_stream = static_cast<cudaStream_t>(cuda_stream); p->ran = true; p->host_only = true;
    if (p->local_rows == 0) return BH_OK;
    BH_CUDA(launch_ray_pass(P, cfg, p->last_stream));
    return BH_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ shared host frame
struct bh_host_frame {
    bh_ctx *ctx = nullptr;
    void *base = nullptr;
    size_t frame_bytes = 0, map_bytes = 0;
    bool registered = false;
    char name[256] = "";
    bool is_path = false;
    uint32_t *flags() const { return reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(base) + align_up(frame_bytes, 4096)); }
};

extern "C" {

int bh_host_frame_create(bh_ctx *ctx, const char *shm_name, size_t nbytes, int create, bh_host_frame **out)
{
    if (!out) { set_error("bh_host_frame_create: out is NULL"); return BH_ERR_INVALID; }
    *out = nullptr;
    if (!ctx || !shm_name || !shm_name[0] || strlen(shm_name) >= sizeof(bh_host_frame::name) || nbytes == 0) {
        set_error("bh_host_frame_create: bad argument");
        return BH_ERR_INVALID;
    }
    // "/name" is a POSIX shared-memory object; anything with a second '/' is a file path (e.g. on a tmpfs of the caller's choice)
    const bool is_path = strchr(shm_name + 1, '/') != nullptr;
    const size_t map_bytes = align_up(nbytes, 4096) + 4096;
    const int oflag = create ? (O_CREAT | O_RDWR) : O_RDWR;
    const int fd = is_path ? open(shm_name, oflag, 0600) : shm_open(shm_name, oflag, 0600);
    if (fd < 0) { set_error("bh_host_frame_create: cannot open %s: %s", shm_name, strerror(errno)); return BH_ERR_NOENT; }
    if (create) {
        // posix_fallocate reserves the pages: a tmpfs that is too small answers ENOSPC here instead of SIGBUS at the first store
        const int frc = posix_fallocate(fd, 0, (off_t)map_bytes);
        if (frc != 0) {
            close(fd);
            if (is_path) unlink(shm_name); else shm_unlink(shm_name);
            set_error("bh_host_frame_create: cannot reserve %zu bytes for %s: %s", map_bytes, shm_name, strerror(frc));
            return BH_ERR_NOMEM;
        }
    } else {
        struct stat st;
        if (fstat(fd, &st) != 0 || (size_t)st.st_size < map_bytes) {
            close(fd);
            set_error("bh_host_frame_create: %s is smaller than %zu bytes (created by another rank with a different size?)", shm_name, map_bytes);
            return BH_ERR_INVALID;
        }
    }
    void *base = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (base == MAP_FAILED) { set_error("bh_host_frame_create: mmap failed: %s", strerror(errno)); return BH_ERR_NOMEM; }
    bh_host_frame *hf = new (std::nothrow) bh_host_frame();
    if (!hf) { munmap(base, map_bytes); set_error("bh_host_frame_create: out of host memory"); return BH_ERR_NOMEM; }
    hf->ctx = ctx; hf->base = base; hf->frame_bytes = nbytes; hf->map_bytes = map_bytes; hf->is_path = is_path;
    snprintf(hf->name, sizeof hf->name, "%s", shm_name);
    if (create) memset(hf->flags(), 0, 4096);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaHostRegister(base, map_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        const int rc = cuda_fail(e, "bh_host_frame_create: cudaHostRegister");
        bh_host_frame_destroy(hf, create);
        return rc;
    }
    hf->registered = true;
    *out = hf;
    return BH_OK;
}

void bh_host_frame_destroy(bh_host_frame *hf, int unlink_name)
{
    if (!hf) return;
    if (hf->registered) { if (ctx_alive(hf->ctx)) cudaSetDevice(hf->ctx->device); cudaHostUnregister(hf->base); }
    if (hf->base) munmap(hf->base, hf->map_bytes);
    if (unlink_name) { if (hf->is_path) unlink(hf->name); else shm_unlink(hf->name); }
    delete hf;
}

float *bh_host_frame_ptr(const bh_host_frame *hf) { return hf ? static_cast<float *>(hf->base) : nullptr; }

int bh_host_frame_signal(bh_host_frame *hf, uint32_t slot, uint32_t value)
{
    if (!hf || slot >= BH_SHARED_FLAGS) { set_error("bh_host_frame_signal: bad argument"); return BH_ERR_INVALID; }
    __atomic_store_n(hf->flags() + slot, value, __ATOMIC_RELEASE);
    return BH_OK;
}

int bh_host_frame_wait(bh_host_frame *hf, uint32_t first_slot, uint32_t n_slots, uint32_t value, uint32_t timeout_ms)
{
    if (!hf || n_slots == 0 || first_slot + n_slots > BH_SHARED_FLAGS) { set_error("bh_host_frame_wait: bad argument"); return BH_ERR_INVALID; }
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t i = 0; i < n_slots; ++i) {
        unsigned spins = 0;
        while ((int32_t)(__atomic_load_n(hf->flags() + first_slot + i, __ATOMIC_ACQUIRE) - value) < 0) {
            if ((++spins & 63u) == 0u) {
                timespec t;
                clock_gettime(CLOCK_MONOTONIC, &t);
                const double ms = (t.tv_sec - t0.tv_sec) * 1e3 + (t.tv_nsec - t0.tv_nsec) * 1e-6;
                if (ms > (double)timeout_ms) { set_error("bh_host_frame_wait: flag %u never reached %u", first_slot + i, value); return BH_ERR_TIMEOUT; }
                sched_yield();
            }
        }
    }
    return BH_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ bh_frame_multi
struct bh_frame_multi {
    uint32_t n = 0, levels = 0, w = 0, h = 0, band_rows = 0;
    int sky_format = -1;
    std::vector<bh_ctx *> ctx;
    std::vector<std::vector<bh_ray_pipeline *>> lv;      // [device][level]
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> done;                       // per device: its rows of the frame are stored
    cudaEvent_t consumed = nullptr;                      // device 0: the object's own consumers of the frame have finished
    cudaEvent_t t0 = nullptr, t1 = nullptr;              // device 0: frame timing
    float4 *frame = nullptr;                             // device 0: the assembled last level
    bh_sky_pipeline *sky = nullptr;
    bool ran = false;
};

static int frame_multi_enqueue(bh_frame_multi *fm, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                               const bh_ray_details *details, float *host_frame)
{
    if (!fm || !camera || !black_hole || !details) { set_error("bh_frame_multi_pass: NULL argument"); return BH_ERR_INVALID; }
    for (uint32_t d = 0; d < fm->n; ++d) {
        BH_CUDA(cudaSetDevice(fm->ctx[d]->device));
        // the previous frame is still being read by the sky resolve / a pending read on device 0: writers wait for it
        if (fm->ran) BH_CUDA(cudaStreamWaitEvent(fm->stream[d], fm->consumed, 0));
        if (d == 0) BH_CUDA(cudaEventRecord(fm->t0, fm->stream[0]));
        for (uint32_t k = 0; k < fm->levels; ++k) {
            int rc;
            if (host_frame && k + 1 == fm->levels)
                rc = bh_ray_pipeline_pass_to_host_frame(fm->lv[d][k], camera, black_hole, details, host_frame, fm->stream[d]);
            else
                rc = bh_ray_pipeline_pass(fm->lv[d][k], camera, black_hole, details, fm->stream[d]);
            if (rc != BH_OK) return rc;
        }
        BH_CUDA(cudaEventRecord(fm->done[d], fm->stream[d]));
    }
    BH_CUDA(cudaSetDevice(fm->ctx[0]->device));
    for (uint32_t d = 1; d < fm->n; ++d) BH_CUDA(cudaStreamWaitEvent(fm->stream[0], fm->done[d], 0));
    if (fm->sky && !host_frame) {
        const int rc = bh_sky_pipeline_pass(fm->sky, fm->stream[0]);
        if (rc != BH_OK) return rc;
    }
    BH_CUDA(cudaEventRecord(fm->t1, fm->stream[0]));
    BH_CUDA(cudaEventRecord(fm->consumed, fm->stream[0]));
    fm->ran = true;
    return BH_OK;
}

extern "C" {

int bh_frame_multi_create(bh_ctx *const *ctxs, uint32_t n_devices, const bh_frame_multi_desc *desc, bh_frame_multi **out)
{
    if (!out) { set_error("bh_frame_multi_create: out is NULL"); return BH_ERR_INVALID; }
    *out = nullptr;
    if (!ctxs || !desc || n_devices == 0 || n_devices > 64 || desc->levels == 0 || desc->levels > 8 || desc->band_rows == 0 ||
        desc->base_width < 2 || desc->base_height < 2 || (desc->levels > 1 && desc->multiplier < 2) || desc->sky_format > 1) {
        set_error("bh_frame_multi_create: bad argument");
        return BH_ERR_INVALID;
    }
    for (uint32_t d = 0; d < n_devices; ++d) {
        if (!ctxs[d]) { set_error("bh_frame_multi_create: context %u is NULL", d); return BH_ERR_INVALID; }
        for (uint32_t e = 0; e < d; ++e)
            if (ctxs[e]->device == ctxs[d]->device) { set_error("bh_frame_multi_create: contexts %u and %u share device %d", e, d, ctxs[d]->device); return BH_ERR_INVALID; }
    }
    // level sizes: mod.rs:177-206 (f32 arithmetic, cast to u32 at use)
    std::vector<uint32_t> lw, lh;
    {
        float cw = (float)desc->base_width, ch = (float)desc->base_height;
        const float m = (float)desc->multiplier;
        for (uint32_t k = 0; k < desc->levels; ++k) {
            lw.push_back((uint32_t)cw); lh.push_back((uint32_t)ch);
            cw = cw * m - (m - 1.0f); ch = ch * m - (m - 1.0f);
        }
        if (lw.back() > 65535u || lh.back() > 65535u) { set_error("bh_frame_multi_create: final level %ux%u out of range", lw.back(), lh.back()); return BH_ERR_INVALID; }
    }
    bh_frame_multi *fm = new (std::nothrow) bh_frame_multi();
    if (!fm) { set_error("bh_frame_multi_create: out of host memory"); return BH_ERR_NOMEM; }
    fm->n = n_devices; fm->levels = desc->levels; fm->w = lw.back(); fm->h = lh.back(); fm->band_rows = desc->band_rows;
    fm->sky_format = desc->sky_format;
    fm->ctx.assign(ctxs, ctxs + n_devices);
    fm->lv.resize(n_devices); fm->stream.assign(n_devices, nullptr); fm->done.assign(n_devices, nullptr);
    int rc = BH_OK;
    auto fail = [&](int code) { bh_frame_multi_destroy(fm); return code; };
    cudaError_t e = cudaSetDevice(ctxs[0]->device);
    if (e == cudaSuccess) e = cudaMalloc(&fm->frame, (size_t)fm->w * fm->h * sizeof(float4));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fm->consumed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&fm->t0);
    if (e == cudaSuccess) e = cudaEventCreate(&fm->t1);
    if (e != cudaSuccess) return fail(cuda_fail(e, "bh_frame_multi_create: device 0 resources"));
    for (uint32_t d = 0; d < n_devices; ++d) {
        e = cudaSetDevice(ctxs[d]->device);
        if (e == cudaSuccess && d > 0) {
            int can = 0;
            e = cudaDeviceCanAccessPeer(&can, ctxs[d]->device, ctxs[0]->device);
            if (e == cudaSuccess && !can) {
                set_error("bh_frame_multi_create: device %d cannot access device %d's memory (no NVLink/PCIe peer path)", ctxs[d]->device, ctxs[0]->device);
                return fail(BH_ERR_NODEV);
            }
            if (e == cudaSuccess) {
                e = cudaDeviceEnablePeerAccess(ctxs[0]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            }
        }
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&fm->stream[d], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fm->done[d], cudaEventDisableTiming);
        if (e != cudaSuccess) return fail(cuda_fail(e, "bh_frame_multi_create: per-device resources"));
        for (uint32_t k = 0; k < desc->levels; ++k) {
            bh_ray_pipeline *p = nullptr;
            rc = bh_ray_pipeline_create(ctxs[d], lw[k], lh[k], k ? fm->lv[d][k - 1] : nullptr, &p);
            if (rc != BH_OK) return fail(rc);
            fm->lv[d].push_back(p);
        }
        bh_ray_pipeline *last = fm->lv[d].back();
        if (n_devices > 1) { rc = bh_ray_pipeline_set_tiling(last, desc->band_rows, d, n_devices); if (rc != BH_OK) return fail(rc); }
        rc = bh_ray_pipeline_bind_frame(last, fm->frame);
        if (rc != BH_OK) return fail(rc);
    }
    if (desc->sky_format >= 0) {
        rc = bh_sky_pipeline_create_for_frame(ctxs[0], fm->frame, fm->w, fm->h, (bh_sky_format)desc->sky_format, &fm->sky);
        if (rc != BH_OK) return fail(rc);
    }
    *out = fm;
    return BH_OK;
}

void bh_frame_multi_destroy(bh_frame_multi *fm)
{
    if (!fm) return;
    for (uint32_t d = 0; d < fm->n; ++d)
        if (!ctx_alive(fm->ctx[d])) { delete fm; return; }       // a context went first: leak rather than touch it
    for (uint32_t d = 0; d < fm->n; ++d) {
        cudaSetDevice(fm->ctx[d]->device);
        if (fm->stream[d]) cudaStreamSynchronize(fm->stream[d]);
    }
    if (fm->sky) bh_sky_pipeline_destroy(fm->sky);
    for (uint32_t d = 0; d < fm->n; ++d) {
        for (size_t k = fm->lv[d].size(); k-- > 0;) bh_ray_pipeline_destroy(fm->lv[d][k]);
        cudaSetDevice(fm->ctx[d]->device);
        if (fm->done[d]) cudaEventDestroy(fm->done[d]);
        if (fm->stream[d]) cudaStreamDestroy(fm->stream[d]);
    }
    if (fm->n) cudaSetDevice(fm->ctx[0]->device);
    if (fm->consumed) cudaEventDestroy(fm->consumed);
    if (fm->t0) cudaEventDestroy(fm->t0);
    if (fm->t1) cudaEventDestroy(fm->t1);
    if (fm->frame) cudaFree(fm->frame);
    delete fm;
}

uint32_t bh_frame_multi_width(const bh_frame_multi *fm) { return fm ? fm->w : 0; }
uint32_t bh_frame_multi_height(const bh_frame_multi *fm) { return fm ? fm->h : 0; }

int bh_frame_multi_pass(bh_frame_multi *fm, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                        const bh_ray_details *details)
{
    return frame_multi_enqueue(fm, camera, black_hole, details, nullptr);
}

int bh_frame_multi_pass_to_host(bh_frame_multi *fm, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                                const bh_ray_details *details, float *pinned_host_rgba32f)
{
    if (!pinned_host_rgba32f) { set_error("bh_frame_multi_pass_to_host: host frame is NULL"); return BH_ERR_INVALID; }
    return frame_multi_enqueue(fm, camera, black_hole, details, pinned_host_rgba32f);
}

int bh_frame_multi_sync(bh_frame_multi *fm)
{
    if (!fm) { set_error("bh_frame_multi_sync: NULL object"); return BH_ERR_INVALID; }
    for (uint32_t d = 0; d < fm->n; ++d) {
        BH_CUDA(cudaSetDevice(fm->ctx[d]->device));
        BH_CUDA(cudaStreamSynchronize(fm->stream[d]));
    }
    return BH_OK;
}

const float *bh_frame_multi_output(const bh_frame_multi *fm) { return fm ? reinterpret_cast<const float *>(fm->frame) : nullptr; }
const void *bh_frame_multi_sky_output(const bh_frame_multi *fm) { return (fm && fm->sky) ? bh_sky_pipeline_output(fm->sky) : nullptr; }

int bh_frame_multi_read(bh_frame_multi *fm, float *host_rgba32f, void *host_sky)
{
    if (!fm) { set_error("bh_frame_multi_read: NULL object"); return BH_ERR_INVALID; }
    if (!fm->ran) { set_error("bh_frame_multi_read: no pass has been enqueued"); return BH_ERR_STATE; }
    if (host_sky && !fm->sky) { set_error("bh_frame_multi_read: created without a sky resolve"); return BH_ERR_STATE; }
    if (host_rgba32f && fm->lv[0].back()->host_only) { set_error("bh_frame_multi_read: the last pass wrote to host memory (bh_frame_multi_pass_to_host)"); return BH_ERR_STATE; }
    const int rc = bh_frame_multi_sync(fm);
    if (rc != BH_OK) return rc;
    BH_CUDA(cudaSetDevice(fm->ctx[0]->device));
    if (host_rgba32f) BH_CUDA(cudaMemcpy(host_rgba32f, fm->frame, (size_t)fm->w * fm->h * sizeof(float4), cudaMemcpyDeviceToHost));
    if (host_sky) return bh_sky_pipeline_read(fm->sky, host_sky);
    return BH_OK;
}

int bh_frame_multi_stats(bh_frame_multi *fm, bh_pass_stats *out, float *elapsed_ms)
{
    if (!fm || !out) { set_error("bh_frame_multi_stats: NULL argument"); return BH_ERR_INVALID; }
    if (!fm->ran) { set_error("bh_frame_multi_stats: no pass has been enqueued"); return BH_ERR_STATE; }
    int rc = bh_frame_multi_sync(fm);
    if (rc != BH_OK) return rc;
    memset(out, 0, sizeof *out);
    int worst = BH_OK;
    auto add = [&](bh_ray_pipeline *p) {
        bh_pass_stats s;
        const int r = bh_ray_pipeline_stats(p, &s);
        if (r != BH_OK && r != BH_ERR_NUMERIC) return r;
        if (r == BH_ERR_NUMERIC) worst = r;
        out->ray_steps += s.ray_steps; out->px_traced += s.px_traced; out->px_copied += s.px_copied; out->px_interp += s.px_interp;
        out->node_visits += s.node_visits; out->tri_tests += s.tri_tests; out->tex_samples += s.tex_samples;
        out->rk_reject += s.rk_reject; out->stack_overflow += s.stack_overflow;
        return BH_OK;
    };
    for (uint32_t k = 0; k + 1 < fm->levels; ++k) { rc = add(fm->lv[0][k]); if (rc != BH_OK) return rc; }
    for (uint32_t d = 0; d < fm->n; ++d) { rc = add(fm->lv[d].back()); if (rc != BH_OK) return rc; }
    if (elapsed_ms) {
        BH_CUDA(cudaSetDevice(fm->ctx[0]->device));
        BH_CUDA(cudaEventElapsedTime(elapsed_ms, fm->t0, fm->t1));
    }
    return worst;
}

}  // extern "C"
