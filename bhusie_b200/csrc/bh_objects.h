// bh_objects.h — the objects behind the opaque handles of include/bh_abi.h, shared by bh_abi.cu and bh_multi.cu.
#pragma once
#include "bh_device.h"

namespace bh {
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
// true while `ctx` is a context created by bh_ctx_create and not yet destroyed.  The destroy functions of pipelines / passes
// ask before they touch their context: a host that drops objects in the wrong order (bh_ctx_destroy first) then leaks the
// object's device memory instead of dereferencing a freed context.
bool ctx_alive(const struct ::bh_ctx *ctx);
}  // namespace bh

#define BH_CUDA(call)                                             \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return bh::cuda_fail(e_, #call);   \
    } while (0)

struct bh_ctx {
    int device = 0;
    int sm_count = 0;
    uchar4 *tex[3] = { nullptr, nullptr, nullptr };
    int tex_w[3] = { 0, 0, 0 }, tex_h[3] = { 0, 0, 0 };
    unsigned char *models = nullptr;     // BH_MAX_MODELS * kModelStride
    int models_uploaded = 0;
    int numeric_mode = BH_NUMERIC_FUSED;
    unsigned *async_err = nullptr;       // page-locked word set by a bh_stream_wait that gave up (bh_multi.cu)
};

struct bh_ray_pipeline {
    bh_ctx *ctx = nullptr;
    uint32_t w = 0, h = 0;
    const bh_ray_pipeline *prev = nullptr;
    uint32_t band_rows = 0, rank = 0, n_ranks = 1, local_rows = 0;
    float4 *own_out = nullptr;
    size_t own_out_rows = 0;
    float4 *bound_out = nullptr;
    float4 *bound_frame = nullptr;                   // full-frame target (global row addressing), local or peer memory
    int32_t *aux_hit = nullptr;
    uint32_t *aux_steps = nullptr;
    uint8_t *aux_class = nullptr;
    uint32_t aux_mask = 0;
    unsigned long long *stats = nullptr;
    unsigned int *work = nullptr;
    unsigned int *queue = nullptr;
    cudaStream_t last_stream = nullptr;
    bool ran = false;
    bool host_only = false;                          // the last pass stored its RGBA straight into caller host memory (pass_to_host,
                                                     // n_chunks == 0): out() holds nothing of it, so dependants must not read it
    cudaStream_t copy_stream = nullptr;              // chunked read-back (bh_ray_pipeline_pass_to_host)
    cudaEvent_t chunk_done[16] = {};
    cudaEvent_t copy_done = nullptr;
    bool copy_pending = false;
    float4 *out() const { return bound_frame ? bound_frame : (bound_out ? bound_out : own_out); }
};

struct bh_sky_pipeline {
    bh_ctx *ctx = nullptr;
    const bh_ray_pipeline *prev = nullptr;           // nullptr: resolves a raw frame (bh_sky_pipeline_create_for_frame)
    const float4 *raw_prev = nullptr;
    uint32_t raw_w = 0, raw_h = 0;
    bh_sky_format format = BH_SKY_RGBA16F;
    void *own_out = nullptr;
    void *bound_out = nullptr;
    unsigned long long *stats = nullptr;
    cudaStream_t last_stream = nullptr;
    bool ran = false;
    void *out() const { return bound_out ? bound_out : own_out; }
    size_t texel_bytes() const { return format == BH_SKY_RGBA32F ? 16 : 8; }
    size_t pixels() const { return prev ? (size_t)prev->local_rows * prev->w : (size_t)raw_w * raw_h; }
};


namespace bh {
// fills the launch parameters of one ray pass from the pipeline's state and the three uniforms (bh_abi.cu)
int build_pass_params(bh_ray_pipeline *p, const bh_camera_uniform *camera, const bh_black_hole_uniform *black_hole,
                      const bh_ray_details *details, const char *who, PassParams &P);
}  // namespace bh
