// bh_device.h — structures shared by the C-ABI host code (bh_abi.cu) and the kernels (ray_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bh_abi.h"

namespace bh {

// RGBA8 unorm texture in linear device memory (texture.rs:32: Rgba8Unorm; filtering is done in
// fp32 in the kernel — CUDA's hardware filter has 9-bit weights).
struct DevTexture {
    const uchar4 *texels;
    int w, h;
};

// ModelUniform byte offsets (src/renderer/triangle.rs:268-285)
constexpr size_t kMuPosition  = 0;
constexpr size_t kMuVisible   = 12;
constexpr size_t kMuPoints    = 48;
constexpr size_t kMuNormals   = 8388656;
constexpr size_t kMuTriangles = 16777264;
constexpr size_t kMuNodes     = 29360176;
constexpr size_t kMuLookup    = 46137392;
// device copies of consecutive models are placed at a 256-byte-aligned stride so that the
// 16-byte vector loads of nodes/points stay aligned for every model index
constexpr size_t kModelStride = (BH_MODEL_UNIFORM_SIZE + 255) / 256 * 256;

enum StatIndex : int {
    kStatSteps = 0, kStatTraced, kStatCopied, kStatInterp, kStatNodeVisits, kStatTriTests,
    kStatTexSamples, kStatRkReject, kStatStackOverflow, kStatCount
};

enum WorkIndex : int { kWorkNext = 0, kWorkQueueLen = 1, kWorkCount = 4 };

struct PassParams {
    bh_camera_uniform cam;
    bh_black_hole_uniform hole;
    bh_ray_details det;
    DevTexture color, disk, sky;
    const unsigned char *models;     // kModelStride per model
    float4 *out;                     // local_rows x w (band-major), or the whole h x w frame when out_global_rows != 0
    int out_global_rows;             // write row gy of a full frame (possibly a peer GPU's, over NVLink) instead of local row ly
    const float4 *prev;              // ph x pw (nullptr: base level)
    int w, h, pw, ph;
    int band_rows, rank, n_ranks, local_rows;
    int32_t *aux_hit;                // nullable
    uint32_t *aux_steps;             // nullable
    uint8_t *aux_class;              // nullable
    unsigned long long *stats;       // kStatCount
    unsigned int *work;              // kWorkCount
    unsigned int *queue;             // local pixel indices awaiting a trace (fine levels)
    unsigned int n_items;            // tile mode: items [item_begin, n_items) are 8x4 warp tiles of this launch
    unsigned int item_begin;
    int tiles_x;
    unsigned int tile_rows;          // rows of a tile-mode work item: 4 (8x4 pixels, 32 rays per warp), or 2 / 1 on launches that do not fill the GPU
    float disk_k;                    // 1.0021 * |hole.normal| (+inf when degenerate): fast disk-plane rejection, ray_impl.cuh hot_iteration
};

struct SkyParams {
    DevTexture sky;
    const float4 *prev;
    void *out;
    int n_pixels;
    int format;                      // bh_sky_format
    unsigned long long *stats;
};

// ---- post chain
struct HalfImage { const uint2 *px; int w, h; };     // RGBA16F, 8 B per texel

struct PostParams {
    HalfImage in1, in2;
    void *out;                       // RGBA16F (uint2 per pixel); FXAA: RGBA8 sRGB (uint per pixel)
    int w, h;
    float mix_ratio;                 // MixDetails (mix_pipeline.rs:5-7)
    float edge_min, edge_max, subpix;    // FXAADetailsUniform (fxaa_pipline.rs:74-80)
    int iterations;
};

struct LaunchConfig {
    int sm_count;
    int numeric_mode;                // bh_numeric_mode
};

// launchers (ray_kernels.cu)
cudaError_t launch_ray_pass(const PassParams &p, const LaunchConfig &cfg, cudaStream_t stream);
// base level only: trace the tile range [p.item_begin, p.n_items) without resetting the pass statistics
cudaError_t launch_trace_range(const PassParams &p, const LaunchConfig &cfg, bool reset_stats, cudaStream_t stream);
cudaError_t launch_sky_pass(const SkyParams &p, const LaunchConfig &cfg, cudaStream_t stream);
cudaError_t launch_post_pass(int kind, const PostParams &p, const LaunchConfig &cfg, cudaStream_t stream);
cudaError_t launch_disk_texture(int w, int h, uchar4 *out, cudaStream_t stream);
cudaError_t launch_math_probe(int fn, const float *a, const float *b, float *out, size_t n, cudaStream_t stream);

}  // namespace bh
