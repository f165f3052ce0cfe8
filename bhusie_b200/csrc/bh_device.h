// bh_device.h — structures shared by the C-ABI host code (bh_abi.cu) and the kernels (ray_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
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
    float disk_far;                  // 1.001 * accretion_disk_outer (+inf when unusable): a segment that starts farther than
                                     // disk_far + 1.01 h from the hole cannot reach the annulus, ray_impl.cuh hot_iteration
    int angle_fast;                  // classification (ray.wgsl:221-226): 1 when cos_hi / cos_lo bracket cos(angle_division_threshold)
    float cos_hi, cos_lo;            // cosine above cos_hi: the angle is provably below the threshold; below cos_lo: provably not
                                     // (classify_kernel evaluates the literal acos only in between), derive_pass_constants
};

// The two constants above, derived once per pass on the host.  They only ever SKIP a test whose outcome is then provably
// "miss", so any value at least as large as stated is valid; degenerate inputs disable the shortcut (+inf).
inline void derive_pass_constants(PassParams &P)
{
    const float *n = P.hole.normal;
    const float nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    P.disk_k = (nn > 1e-30f && nn < 1e30f) ? 1.0021f * sqrtf(nn) : INFINITY;
    // |ip - bh| >= |p - bh| - t |d| for every point of the segment; the 0.1 % margin covers the rounding of the computed
    // distances as long as the hole is not placed absurdly far from the origin (absolute rounding error ~ 1e-7 |bh|)
    const float outer = P.hole.accretion_disk_outer;
    const float *b = P.hole.position;
    const float bmax = fmaxf(fmaxf(fabsf(b[0]), fabsf(b[1])), fabsf(b[2]));
    P.disk_far = (outer > 1e-3f && outer < 1e30f && bmax <= 64.0f * outer) ? 1.001f * outer : INFINITY;
    // `acos(c) < thr` (angle_between, ray.wgsl:221-226) without the acos for all but the few c next to cos(thr).  The kernel's
    // acos is atan2 in float64 rounded to f32: relative error < 7e-8.  So acos(c) < thr (1 - 1e-6) makes the rounded value < thr
    // and acos(c) > thr (1 + 1e-6) makes it > thr (rounding is monotonic, and the margins are 8 f32 ulps wide); in terms of c,
    // with cos decreasing on [0, pi]: c > cos(thr (1 - 1e-6)) and c < cos(thr (1 + 1e-6)).  One more f32 ulp either way covers
    // the host's own cos().  For thr = 0.02 the two bounds are 3 f32 values apart.
    const float thr = P.det.angle_division_threshold;
    P.angle_fast = 0; P.cos_hi = INFINITY; P.cos_lo = -INFINITY;
    if (thr >= 1e-3f && thr <= 3.0f) {
        const double t = (double)thr;
        const float hi = (float)cos(t * (1.0 - 1e-6)), lo = (float)cos(t * (1.0 + 1e-6));
        P.cos_hi = nextafterf(nextafterf(hi, 2.0f), 2.0f);
        P.cos_lo = nextafterf(nextafterf(lo, -2.0f), -2.0f);
        P.angle_fast = 1;
    }
}

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
