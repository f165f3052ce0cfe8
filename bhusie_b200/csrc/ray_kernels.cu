// ray_kernels.cu — bhusie's per-pixel geodesic ray pass for B200 (sm_100a).
//
// What it computes is ray.wgsl::main (src/renderer/shaders/ray.wgsl:167-243) and sky.wgsl::main
// (src/renderer/shaders/sky.wgsl:8-30); how it runs is not the WGSL's one-invocation-per-pixel:
//
//   * classify_kernel (fine pyramid levels only) decides copy / interpolate / trace per pixel
//     (ray.wgsl:185-241), writes the first two directly and appends the third to a device queue
//     with one warp-ballot-compacted atomic per warp.
//   * trace_kernel is a persistent grid (4 CTAs per SM); each WARP pulls work items — an 8x4
//     pixel tile on the base level, 32 queue entries on fine levels (fewer rays per warp on launches
//     that do not fill the GPU) — from a global counter.  A ray's integrator state (9 scalars) lives in
//     registers; the other variables of trace_ray (ray.wgsl:482-596) live in a per-thread row of
//     shared memory and are only written when something happens to the ray.  The per-ray state
//     machine is run phase-sorted inside the warp: all lanes that want an integration step run the
//     hot loop together (one basic block per step, hot_iteration in ray_impl.cuh); lanes that
//     crossed the disk are shaded together; lanes that left the relativity sphere wait and are
//     then served together by the flat-space branch (BVH + sphere re-entry).  Every ray still
//     executes exactly its own sequence of operations; only the interleaving across lanes differs.
//
// Numerics: compiled with --fmad=false, IEEE div/sqrt; one float op per WGSL expression node,
// left to right; transcendentals from detmath.cuh.  See DESIGN.md §4 ("numeric contract").
#include "bh_device.h"
#include "detmath.cuh"
#include <cuda_fp16.h>
#include <string.h>

namespace bh {

// ------------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers: global -> shared staging of the BVH top
// ------------------------------------------------------------------------------------------------
namespace tma {
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
}  // namespace tma

// BVH top staged per CTA: ModelUniform header (position, visible, ...: 48 B) + the first kTopNodes nodes in the
// reference's pre-order numbering (root, its two children, the left spine ...).  A ray that misses the mesh — the
// common case — touches only nodes 0,1,2, so its whole flat-space test runs out of shared memory.
constexpr int kTopNodes = 256;
constexpr int kTopHeaderBytes = 64;                       // 48 used, padded to keep the nodes 16-byte aligned
constexpr int kTopBytes = kTopHeaderBytes + kTopNodes * 32;

// Resident CTAs per SM (tuning: -DBH_OCC_RK=n -DBH_OCC_EULER=n).  Measured at 4K (tools/gpu_c3_probe.py): the Cash-Karp
// kernel is fastest spill-free at 4 CTAs/SM (14.7 ms; 5: 15.1, 6: 15.3): its hot loop runs at ~70 % of three coincident
// limits (issue slots, FMA pipe, register operand bandwidth), so extra warps buy nothing once spills appear.  Euler — a short
// dependent chain per step, latency- rather than pipe-bound — gains from a fifth CTA since the kernel shrank in round 2:
// 7.72 -> 7.47 ms (254 -> 263 G ray-steps/s) at 5, 7.70 at 6 (round 1: flat from 4 to 6, 9.36 ms at 8 with spills).  A value
// other than 4 for Euler adds a second build that launch_trace_mode uses on tile-mode frames that saturate the GPU.
#ifndef BH_OCC_RK
#define BH_OCC_RK 4
#endif
#ifndef BH_OCC_EULER
#define BH_OCC_EULER 5
#endif
// lanes of a 32-ray warp that must wait for disk shading before the hot phase is interrupted for them (1 = serve every crossing at
// once).  Reference frame, RK: 4 -> 2.72 ms, 8 -> 2.59, 16 -> 2.53, 24 -> 2.46, 32 -> 2.43 (profiles/r2_10_*, r2_11_*_quick.json): in
// effect a crossing is shaded when nobody steps any more or when it has waited BH_SHADE_PATIENCE votes (8 -> 2.43 ms, 12 -> 2.45 ms,
// Euler 2.00 -> 1.98 ms, 4K grid 6.50 -> 6.47 ms); narrower warps scale the batch down.
#ifndef BH_SHADE_BATCH
#define BH_SHADE_BATCH 32
#endif
// ... or one crossing has waited this many warp votes (two steps each)
#ifndef BH_SHADE_PATIENCE
#define BH_SHADE_PATIENCE 12
#endif

// one-copy hot loop in queue-mode kernels (see trace_warp in ray_impl.cuh; 0 = the two-copy loop everywhere)
#ifndef BH_QUEUE_COMPACT
#define BH_QUEUE_COMPACT 1
#endif

#define BH_NUM_NS lit
#define BH_FUSED 0
#include "ray_impl.cuh"
#include "post_impl.cuh"
#undef BH_NUM_NS
#undef BH_FUSED
#define BH_NUM_NS fus
#define BH_FUSED 1
#include "ray_impl.cuh"
#include "post_impl.cuh"
#undef BH_NUM_NS
#undef BH_FUSED

namespace {
// ------------------------------------------------------------------------------------------------
// disk texture generator: perlin/src/main.rs:6-148, the offline tool behind disk.png.
// Rust never contracts f32 arithmetic, so this kernel is numeric-mode independent: plain IEEE ops (the TU is
// compiled with --fmad=false) and det-math transcendentals.  One thread per texel: the spiral map is the same for
// all four octaves, so the multi-pass generate -> spiral -> merge of the tool collapses into one gather-free pass.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned rotl32(unsigned v, unsigned s) { return (v << s) | (v >> (32u - s)); }
__device__ __forceinline__ unsigned sat_u8(float v) { return !(v == v) ? 0u : (v <= 0.0f ? 0u : (v >= 255.0f ? 255u : (unsigned)v)); }
__device__ __forceinline__ unsigned sat_u32(float v) { return !(v == v) ? 0u : (v <= 0.0f ? 0u : (v >= 4294967296.0f ? 4294967295u : (unsigned)v)); }

__device__ __forceinline__ float dot_grid_gradient(unsigned ix, unsigned iy, float x, float y)      // main.rs:6-32
{
    unsigned a = ix, b = iy;
    a *= 3284157443u;
    b ^= rotl32(a, 16);
    b *= 1911520717u;
    a ^= rotl32(b, 16);
    a *= 2048419325u;
    const float random = (float)a * (3.14159265358979323846f / (float)0xffffffffu);
    float gx, gy;
    detmath::sincos_f(random, gy, gx);
    const float dx = x - (float)ix, dy = y - (float)iy;
    return dx * gx + dy * gy;
}
__device__ __forceinline__ float perlin_interp(float a0, float a1, float w)                          // main.rs:34-37
{
    return (a1 - a0) * ((w * (w * 6.0f - 15.0f) + 10.0f) * w * w * w) + a0;
}
__device__ float perlin2(float x, float y)                                                            // main.rs:39-57
{
    const unsigned x0 = sat_u32(floorf(x)), x1 = x0 + 1u, y0 = sat_u32(floorf(y)), y1 = y0 + 1u;
    const float sx = x - (float)x0, sy = y - (float)y0;
    float n0 = dot_grid_gradient(x0, y0, x, y), n1 = dot_grid_gradient(x1, y0, x, y);
    const float ix0 = perlin_interp(n0, n1, sx);
    n0 = dot_grid_gradient(x0, y1, x, y); n1 = dot_grid_gradient(x1, y1, x, y);
    const float ix1 = perlin_interp(n0, n1, sx);
    return perlin_interp(ix0, ix1, sy) * 0.5f + 0.5f;
}

__global__ void __launch_bounds__(256) disk_texture_kernel(int w, int h, uchar4 *out)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const float PI32 = 3.14159265358979323846f;
    // spiral(amount 2, power 0.5), main.rs:79-110
    const float rx = ((float)x / (float)w) * 2.0f - 1.0f, ry = ((float)y / (float)h) * 2.0f - 1.0f;
    const float r = sqrtf(rx * rx + ry * ry);
    float theta = detmath::atan2_f(ry, rx);
    theta = fmodf(theta + PI32 + detmath::pow_f(r, 0.5f) * PI32 * 2.0f, 2.0f * PI32) - PI32;
    float sn, cs;
    detmath::sincos_f(theta, sn, cs);
    const unsigned nx = sat_u32((r * cs * 0.5f + 0.5f) * (float)w) % (unsigned)w;
    const unsigned ny = sat_u32((r * sn * 0.5f + 0.5f) * (float)h) % (unsigned)h;
    // generate() at the warped texel for the four densities (main.rs:60-77,134-141), merged finest-first (143-145)
    const float dens[4] = { 4.0f / (float)w, 20.0f / (float)w, 50.0f / (float)w, 100.0f / (float)w };
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = sat_u8(perlin2((float)nx * dens[k], (float)ny * dens[k]) * 256.0f);
    unsigned m = sat_u8((float)v[3] * 0.5f + (float)v[2] * (1.0f - 0.5f));
    m = sat_u8((float)m * 0.5f + (float)v[1] * (1.0f - 0.5f));
    m = sat_u8((float)m * 0.5f + (float)v[0] * (1.0f - 0.5f));
    out[(size_t)y * (size_t)w + (size_t)x] = make_uchar4((unsigned char)m, (unsigned char)m, (unsigned char)m, (unsigned char)m);
}

__global__ void math_probe_kernel(int fn, const float *a, const float *b, float *out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r;
    switch (fn) {
    case 0: r = detmath::pow_f(a[i], b[i]); break;
    case 1: r = detmath::pow5_f(a[i]); break;
    case 2: r = detmath::pow4_f(a[i]); break;
    case 3: r = detmath::sin_f(a[i]); break;
    case 4: r = detmath::cos_f(a[i]); break;
    case 5: r = detmath::tan_f(a[i]); break;
    case 6: r = detmath::atan2_f(a[i], b[i]); break;
    case 7: r = detmath::acos_f(a[i]); break;
    default: r = CUDART_NAN_F;
    }
    out[i] = r;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
template <typename K>
static cudaError_t launch_trace(K kernel, unsigned items, const PassParams &p, const LaunchConfig &cfg, cudaStream_t stream)
{
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned grid = (unsigned)(cfg.sm_count * per_sm);      // persistent: one wave of resident CTAs
    // never more CTAs than there can be work (`items` = warp work items of this launch, an upper bound in queue mode)
    const unsigned need = (items + 3u) / 4u;
    if (need < grid) grid = need ? need : 1u;
    kernel<<<grid, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

template <bool QUEUE>
static cudaError_t launch_trace_mode(const PassParams &p, const LaunchConfig &cfg, cudaStream_t stream)
{
    const bool euler = p.det.integration_method == 0;
    // warp work items: 8x4 tiles (tile mode) / 32 queue entries
    // (queue mode: an upper bound, in the narrowest items trace_kernel may choose, so that the grid is not what limits it)
    const unsigned items = QUEUE ? (unsigned)(((size_t)p.local_rows * (size_t)p.w + 7) / 8) : p.n_items - p.item_begin;
    // Euler only: the higher-occupancy build when every warp of it would still get several items (tile mode knows the count)
    const bool hi = !QUEUE && euler && BH_OCC_EULER != 4 && items >= 6u * (unsigned)cfg.sm_count * 4u * (unsigned)BH_OCC_EULER;
    if (!QUEUE && p.tile_rows == 4 && p.item_begin == 0 && p.n_items == (unsigned)p.tiles_x * (unsigned)((p.local_rows + 3) / 4)) {
        // whole-frame tile launch with fewer 8x4 tiles than warp slots: 8x2 or 8x1 tiles (see trace_kernel)
        const unsigned slots = (unsigned)cfg.sm_count * 4u * 4u;
        for (unsigned r = 1; r <= 2; r *= 2) {
            const unsigned n = (unsigned)p.tiles_x * (unsigned)((p.local_rows + (int)r - 1) / (int)r);
            if (n <= slots) {
                PassParams q = p;
                q.tile_rows = r;
                q.n_items = n;
                return launch_trace_mode<QUEUE>(q, cfg, stream);
            }
        }
    }
    if (cfg.numeric_mode == BH_NUMERIC_LITERAL) {
        if (!euler) return launch_trace(lit::trace_kernel<1, QUEUE, 4, false>, items, p, cfg, stream);
        if (!QUEUE && hi) return launch_trace(lit::trace_kernel<0, false, BH_OCC_EULER, false>, items, p, cfg, stream);
        return launch_trace(lit::trace_kernel<0, QUEUE, 4, false>, items, p, cfg, stream);
    }
    // hole at exactly (+0,+0,+0) (bit pattern 0 in all three words; the reference's default): the build without the
    // `- bh.position` subtractions, which are bit-for-bit no-ops there
    unsigned pos_bits[3];
    memcpy(pos_bits, p.hole.position, sizeof pos_bits);
    const bool origin = (pos_bits[0] | pos_bits[1] | pos_bits[2]) == 0u;
    if (origin) {
        if (!euler) return launch_trace(fus::trace_kernel<1, QUEUE, BH_OCC_RK, true>, items, p, cfg, stream);
        if (!QUEUE && hi) return launch_trace(fus::trace_kernel<0, false, BH_OCC_EULER, true>, items, p, cfg, stream);
        return launch_trace(fus::trace_kernel<0, QUEUE, 4, true>, items, p, cfg, stream);
    }
    if (!euler) return launch_trace(fus::trace_kernel<1, QUEUE, BH_OCC_RK, false>, items, p, cfg, stream);
    if (!QUEUE && hi) return launch_trace(fus::trace_kernel<0, false, BH_OCC_EULER, false>, items, p, cfg, stream);
    return launch_trace(fus::trace_kernel<0, QUEUE, 4, false>, items, p, cfg, stream);
}

cudaError_t launch_ray_pass(const PassParams &p, const LaunchConfig &cfg, cudaStream_t stream)
{
    (void)cudaGetLastError();          // a stale error of an unrelated earlier call is not this launch's

    cudaError_t e = cudaMemsetAsync(p.work, 0, sizeof(unsigned) * kWorkCount, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(p.stats, 0, sizeof(unsigned long long) * kStatCount, stream);
    if (e != cudaSuccess) return e;
    if (p.prev == nullptr) return launch_trace_mode<false>(p, cfg, stream);
    const size_t total = (size_t)p.local_rows * (size_t)p.w;
    unsigned grid = (unsigned)((total + 255) / 256);
    const unsigned cap = (unsigned)cfg.sm_count * 8u;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    if (cfg.numeric_mode == BH_NUMERIC_LITERAL) lit::classify_kernel<<<grid, 256, 0, stream>>>(p);
    else fus::classify_kernel<<<grid, 256, 0, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_trace_mode<true>(p, cfg, stream);
}

cudaError_t launch_trace_range(const PassParams &p, const LaunchConfig &cfg, bool reset_stats, cudaStream_t stream)
{
    (void)cudaGetLastError();          // a stale error of an unrelated earlier call is not this launch's

    cudaError_t e = cudaMemsetAsync(p.work, 0, sizeof(unsigned) * kWorkCount, stream);
    if (e != cudaSuccess) return e;
    if (reset_stats) {
        e = cudaMemsetAsync(p.stats, 0, sizeof(unsigned long long) * kStatCount, stream);
        if (e != cudaSuccess) return e;
    }
    if (p.item_begin >= p.n_items) return cudaSuccess;
    return launch_trace_mode<false>(p, cfg, stream);
}

cudaError_t launch_sky_pass(const SkyParams &s, const LaunchConfig &cfg, cudaStream_t stream)
{
    (void)cudaGetLastError();          // a stale error of an unrelated earlier call is not this launch's

    unsigned grid = (unsigned)((s.n_pixels + 255) / 256);
    const unsigned cap = (unsigned)cfg.sm_count * 8u;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    if (cfg.numeric_mode == BH_NUMERIC_LITERAL) lit::sky_kernel<<<grid, 256, 0, stream>>>(s);
    else fus::sky_kernel<<<grid, 256, 0, stream>>>(s);
    return cudaGetLastError();
}

cudaError_t launch_post_pass(int kind, const PostParams &p, const LaunchConfig &cfg, cudaStream_t stream)
{
    (void)cudaGetLastError();          // a stale error of an unrelated earlier call is not this launch's

    const dim3 grid((unsigned)((p.w + 31) / 32), (unsigned)((p.h + 7) / 8));
    const bool lit_mode = cfg.numeric_mode == BH_NUMERIC_LITERAL;
    switch (kind) {
    case BH_POST_BLOOM_DOWN: if (lit_mode) lit::bloom_down_kernel<<<grid, 256, 0, stream>>>(p); else fus::bloom_down_kernel<<<grid, 256, 0, stream>>>(p); break;
    case BH_POST_BLOOM_UP:   if (lit_mode) lit::bloom_up_kernel<<<grid, 256, 0, stream>>>(p);   else fus::bloom_up_kernel<<<grid, 256, 0, stream>>>(p); break;
    case BH_POST_MIX:        if (lit_mode) lit::mix_kernel<<<grid, 256, 0, stream>>>(p);        else fus::mix_kernel<<<grid, 256, 0, stream>>>(p); break;
    case BH_POST_HDR:        if (lit_mode) lit::hdr_kernel<<<grid, 256, 0, stream>>>(p);        else fus::hdr_kernel<<<grid, 256, 0, stream>>>(p); break;
    case BH_POST_FXAA:       if (lit_mode) lit::fxaa_kernel<<<grid, 256, 0, stream>>>(p);       else fus::fxaa_kernel<<<grid, 256, 0, stream>>>(p); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_disk_texture(int w, int h, uchar4 *out, cudaStream_t stream)
{
    const dim3 grid((unsigned)((w + 31) / 32), (unsigned)((h + 7) / 8));
    disk_texture_kernel<<<grid, 256, 0, stream>>>(w, h, out);
    return cudaGetLastError();
}

cudaError_t launch_math_probe(int fn, const float *a, const float *b, float *out, size_t n, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(fn, a, b, out, n);
    return cudaGetLastError();
}

}  // namespace bh
