// host_warp.cpp — the PRODUCT's kernels (trace_kernel, classify_kernel, sky_kernel of bhusie_b200/csrc/ray_impl.cuh) run on
// the CPU as ONE WARP OF 32 HOST THREADS in lock-step: every full-mask warp collective (__any_sync, __ballot_sync,
// __shfl_sync, __reduce_add_sync, __syncwarp, __syncthreads) is a barrier + exchange among the 32 threads, shared memory
// is ordinary globals shared by them, atomics are host atomics.  This executes the kernels' real control flow — the
// persistent work loop, phase sorting, batched disk shading, narrow work items, ballot compaction of the trace queue —
// so a whole ray level (and a pyramid, level by level) can be checked against the oracle without a GPU
// (tests/test_host_kernel.py).  TEST INFRASTRUCTURE ONLY.  See host_kernel.cpp for what the arithmetic shims are.
#define BH_HOST_EMULATION 1
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math_constants.h>
using std::max;
using std::min;
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif

// ---- one CTA of one warp
struct Idx { unsigned x, y, z; };
static thread_local Idx threadIdx = { 0, 0, 0 };
static const Idx blockIdx = { 0, 0, 0 }, blockDim = { 32, 1, 1 }, gridDim = { 1, 1, 1 };
static pthread_barrier_t g_bar;
static unsigned g_slot[32];
static inline void warp_barrier() { pthread_barrier_wait(&g_bar); }
static inline unsigned exchange_or(unsigned bit_value)           // ballot: every lane contributes one bit
{
    g_slot[threadIdx.x & 31u] = bit_value;
    warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (g_slot[i] & 1u) << i;
    warp_barrier();
    return m;
}
static inline unsigned __ballot_sync(unsigned mask, int p) { (void)mask; return exchange_or(p ? 1u : 0u); }
static inline bool __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0u; }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src)
{
    g_slot[threadIdx.x & 31u] = v;
    warp_barrier();
    const unsigned r = g_slot[src & 31];
    warp_barrier();
    return r;
}
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31u); }    // divergent code: a lane only knows itself
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v)
{
    if (mask != 0xffffffffu) return v;             // stat_add under __activemask(): lane-local sums add up to the same totals
    g_slot[threadIdx.x & 31u] = v;
    warp_barrier();
    unsigned s = 0;
    for (int i = 0; i < 32; ++i) s += g_slot[i];
    warp_barrier();
    return s;
}
static inline void __syncwarp() { warp_barrier(); }
static inline void __syncthreads() { warp_barrier(); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
// ---- arithmetic intrinsics: one IEEE operation each
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __double2hiint(double d) { long long u; memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline int __double2loint(double d) { long long u; memcpy(&u, &d, 8); return (int)(u & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) { const unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double d; memcpy(&d, &u, 8); return d; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }

#include "../../bhusie_b200/csrc/bh_device.h"
#include "../../bhusie_b200/csrc/detmath.cuh"

namespace bh {
namespace tma {            // cp.async.bulk + mbarrier: the elected thread copies, the __syncthreads() that follows publishes
static inline void mbar_init(unsigned long long *, unsigned) {}
static inline void mbar_expect_tx(unsigned long long *, unsigned) {}
static inline void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *) { memcpy(dst, src, bytes); }
static inline void mbar_wait(unsigned long long *, unsigned) {}
}
constexpr int kTopNodes = 256;
constexpr int kTopHeaderBytes = 64;
constexpr int kTopBytes = kTopHeaderBytes + kTopNodes * 32;
#define BH_SHADE_BATCH 32
#define BH_SHADE_PATIENCE 12
#define BH_QUEUE_COMPACT 1
#define BH_NUM_NS lit
#define BH_FUSED 0
#include "../../bhusie_b200/csrc/ray_impl.cuh"
#undef BH_NUM_NS
#undef BH_FUSED
#define BH_NUM_NS fus
#define BH_FUSED 1
#include "../../bhusie_b200/csrc/ray_impl.cuh"
#undef BH_NUM_NS
#undef BH_FUSED
}  // namespace bh

template <typename F>
static void run_warp(F body)
{
    pthread_barrier_init(&g_bar, nullptr, 32);
    std::vector<std::thread> lanes;
    for (unsigned l = 0; l < 32; ++l) lanes.emplace_back([l, &body] { threadIdx.x = l; body(); });
    for (auto &t : lanes) t.join();
    pthread_barrier_destroy(&g_bar);
}

// One RayPipeline::pass (launch_ray_pass of ray_kernels.cu): work/stats reset, classify_kernel on fine levels, trace_kernel in
// tile or queue mode — with the kernels the launcher would pick (numeric mode, integrator, hole-at-origin build) — then,
// optionally, sky_kernel (RGBA32F) on the result.  tile_rows: 4, or 2 / 1 for the narrow work items.  Returns 0.
extern "C" int bh_host_warp_level(int mode, const void *camera, const void *hole, const void *details,
                                  const uint8_t *color, int cw, int ch, const uint8_t *disk, int dw, int dh,
                                  const uint8_t *sky, int sw, int sh, const unsigned char *models,
                                  int w, int h, const float *prev, int pw, int ph, int tile_rows,
                                  float *out_rgba, int32_t *out_hit, uint32_t *out_steps, uint8_t *out_class,
                                  unsigned long long *stats9, float *sky_rgba32f,
                                  int band_rows, int rank, int n_ranks, int local_rows, int out_global_rows)
{
    using namespace bh;
    PassParams P;
    memset(&P, 0, sizeof P);
    memcpy(&P.cam, camera, sizeof P.cam);
    memcpy(&P.hole, hole, sizeof P.hole);
    memcpy(&P.det, details, sizeof P.det);
    if (P.det.integration_method != 0) P.det.integration_method = 1;
    P.color = DevTexture{ reinterpret_cast<const uchar4 *>(color), cw, ch };
    P.disk = DevTexture{ reinterpret_cast<const uchar4 *>(disk), dw, dh };
    P.sky = DevTexture{ reinterpret_cast<const uchar4 *>(sky), sw, sh };
    P.models = models;
    P.out = reinterpret_cast<float4 *>(out_rgba);
    P.prev = reinterpret_cast<const float4 *>(prev);
    P.w = w; P.h = h; P.pw = prev ? pw : 1; P.ph = prev ? ph : 1;
    // image-space sharding (bh_ray_pipeline_set_tiling / bind_frame): n_ranks <= 1 is the whole frame.  With out_global_rows the
    // pixels go to their global row of a full w x h frame (the peer-store exchange); the aux planes stay band-major.
    if (n_ranks <= 1) { band_rows = h; rank = 0; n_ranks = 1; local_rows = h; out_global_rows = 0; }
    P.band_rows = band_rows; P.rank = rank; P.n_ranks = n_ranks; P.local_rows = local_rows;
    P.out_global_rows = out_global_rows;
    P.aux_hit = out_hit; P.aux_steps = out_steps; P.aux_class = out_class;
    P.tiles_x = (w + 7) / 8;
    P.tile_rows = (unsigned)tile_rows;
    P.item_begin = 0;
    P.n_items = (unsigned)P.tiles_x * (unsigned)((local_rows + tile_rows - 1) / tile_rows);
    unsigned long long stats[kStatCount] = { 0 };
    unsigned work[kWorkCount] = { 0 };
    std::vector<unsigned> queue((size_t)w * (size_t)local_rows + 1);
    P.stats = stats; P.work = work; P.queue = queue.data();
    derive_pass_constants(P);
    unsigned pos_bits[3];
    memcpy(pos_bits, P.hole.position, sizeof pos_bits);
    const bool origin = mode == 1 && (pos_bits[0] | pos_bits[1] | pos_bits[2]) == 0u;
    const bool rk = P.det.integration_method != 0;
    const bool fine = prev != nullptr;
    run_warp([&] {
        if (fine) {
            if (mode == 0) lit::classify_kernel(P); else fus::classify_kernel(P);
            __syncthreads();                                    // kernel boundary
        }
        if (mode == 0) {
            if (fine) { if (rk) lit::trace_kernel<1, true, 4, false>(P); else lit::trace_kernel<0, true, 4, false>(P); }
            else      { if (rk) lit::trace_kernel<1, false, 4, false>(P); else lit::trace_kernel<0, false, 4, false>(P); }
        } else if (origin) {
            if (fine) { if (rk) fus::trace_kernel<1, true, 4, true>(P); else fus::trace_kernel<0, true, 4, true>(P); }
            else      { if (rk) fus::trace_kernel<1, false, 4, true>(P); else fus::trace_kernel<0, false, 4, true>(P); }
        } else {
            if (fine) { if (rk) fus::trace_kernel<1, true, 4, false>(P); else fus::trace_kernel<0, true, 4, false>(P); }
            else      { if (rk) fus::trace_kernel<1, false, 4, false>(P); else fus::trace_kernel<0, false, 4, false>(P); }
        }
    });
    if (sky_rgba32f) {
        SkyParams S;
        memset(&S, 0, sizeof S);
        S.sky = P.sky; S.prev = P.out; S.out = sky_rgba32f; S.n_pixels = w * (out_global_rows ? h : local_rows); S.format = BH_SKY_RGBA32F; S.stats = stats;
        run_warp([&] { if (mode == 0) lit::sky_kernel(S); else fus::sky_kernel(S); });
    }
    if (stats9) memcpy(stats9, stats, sizeof stats);
    return 0;
}

// classify_kernel's angle test (cos_below_threshold) against its literal form acos(c) < thr: every f32 within `span` ulps of
// cos(thr), plus a sweep of [-1, 1] and the special values.  Returns the number of disagreements (0 expected);
// *n_literal = how many of the probed cosines <= 1 still needed the literal acos.
extern "C" long bh_host_angle_shortcut_mismatches(float thr, int span, long *n_literal)
{
    using namespace bh;
    PassParams P;
    memset(&P, 0, sizeof P);
    P.det.angle_division_threshold = thr;
    derive_pass_constants(P);
    long bad = 0, lit_n = 0;
    auto probe = [&](float c) {
        const bool fast = fus::cos_below_threshold(P, c);
        const bool ref = detmath::acos_f(c) < thr;
        if (fast != ref) ++bad;
        if (c <= 1.0f && !(P.angle_fast && (c > P.cos_hi || c < P.cos_lo))) ++lit_n;     // (c > 1 and NaN are literal by design)
    };
    const float centre = (float)cos((double)thr);
    float up = centre, dn = centre;
    probe(centre);
    for (int i = 0; i < span; ++i) { up = nextafterf(up, 4.0f); dn = nextafterf(dn, -4.0f); probe(up); probe(dn); }
    for (int i = -100000; i <= 100000; ++i) probe((float)i * 1e-5f);
    const float specials[] = { 1.0f, nextafterf(1.0f, 2.0f), nextafterf(1.0f, 0.0f), -1.0f, nextafterf(-1.0f, -2.0f), 0.0f, -0.0f,
                               INFINITY, -INFINITY, NAN, 1e-30f, -1e-30f, 2.0f, -2.0f };
    for (float c : specials) probe(c);
    if (n_literal) *n_literal = lit_n;
    return bad;
}
