// host_kernel.cpp — the PRODUCT's device source (bhusie_b200/csrc/ray_impl.cuh, detmath.cuh) compiled for the CPU, one
// lane per "warp", so that the state machine of trace_warp / hot_iteration / hot_tail can be checked against the oracle
// bit for bit without a GPU (tests/test_host_kernel.py).  TEST INFRASTRUCTURE ONLY: nothing in the package loads this.
//
// What is emulated: warp votes and shuffles of a one-lane warp, atomics, __ldg, the packed-FP32 intrinsics (each lane is
// an IEEE fmaf / mul / add, which is what FFMA2 / FMUL2 / FADD2 do), shared memory as plain globals, and the MUFU-seeded
// sqrt_spec / rcp_fast as sqrtf / 1.0f/x (their in-range results are the correctly rounded ones; that equivalence is what
// the GPU suite checks).  What is not: TMA (the BVH top is memcpy'd), scheduling, anything about performance.
#define BH_HOST_EMULATION 1
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <algorithm>
using std::min;
using std::max;
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif

// ---- execution-model shims (one thread: lane 0 of warp 0 of CTA 0)
static const struct { unsigned x, y, z; } threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 }, blockDim = { 128, 1, 1 }, gridDim = { 1, 1, 1 };
static inline bool __any_sync(unsigned, int p) { return p != 0; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
static inline unsigned __activemask() { return 1u; }
static inline void __syncwarp() {}
static inline void __syncthreads() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
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
namespace tma {            // the kernel bodies reference these; the host driver below never calls the kernels
static inline void mbar_init(unsigned long long *, unsigned) {}
static inline void mbar_expect_tx(unsigned long long *, unsigned) {}
static inline void bulk_g2s(void *, const void *, unsigned, unsigned long long *) {}
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

// One single-level ray pass (prev == NULL) over a w x h frame through the product's trace_warp, pixel by pixel.
// mode: 0 LITERAL, 1 FUSED.  models: kModelStride-strided ModelUniform blobs (model_count of them).  Returns 0.
extern "C" int bh_host_kernel_pass(int mode, const void *camera, const void *hole, const void *details,
                                   const uint8_t *color, int cw, int ch, const uint8_t *disk, int dw, int dh,
                                   const uint8_t *sky, int sw, int sh, const unsigned char *models,
                                   int w, int h, float *out_rgba, int32_t *out_hit, uint32_t *out_steps, unsigned long long *stats9)
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
    P.w = w; P.h = h; P.pw = 1; P.ph = 1;
    P.band_rows = h; P.rank = 0; P.n_ranks = 1; P.local_rows = h;
    P.tiles_x = (w + 7) / 8; P.tile_rows = 4;
    unsigned long long stats[kStatCount] = { 0 };
    P.stats = stats;
    derive_pass_constants(P);
    unsigned pos_bits[3];
    memcpy(pos_bits, P.hole.position, sizeof pos_bits);
    const bool origin = mode == 1 && (pos_bits[0] | pos_bits[1] | pos_bits[2]) == 0u;      // launch_trace_mode's choice
    const bool rk = P.det.integration_method != 0;
    const bool origin2 = (pos_bits[0] | pos_bits[1] | pos_bits[2]) == 0u;
    (void)origin2;
    unsigned long long steps_total = 0, traced = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            // what trace_kernel does per CTA / per item around trace_warp
            if (mode == 0) lit::fill_unorm_table(0, 1); else fus::fill_unorm_table(0, 1);
            if (mode == 0) {
                memset(lit::s_warp_stats, 0, sizeof lit::s_warp_stats);
                if (P.det.model_count > 0) { memcpy(lit::s_model_top, models, 48); memcpy(lit::s_model_top + kTopHeaderBytes, models + kMuNodes, kTopNodes * 32); }
            } else {
                memset(fus::s_warp_stats, 0, sizeof fus::s_warp_stats);
                if (P.det.model_count > 0) { memcpy(fus::s_model_top, models, 48); memcpy(fus::s_model_top + kTopHeaderBytes, models + kMuNodes, kTopNodes * 32); }
            }
            float4 rgba; int tri; unsigned steps;
            if (mode == 0) {
                const lit::LaneOut o = rk ? lit::trace_warp<1, false>(P, true, x, y) : lit::trace_warp<0, false>(P, true, x, y);
                rgba = o.rgba; tri = o.tri; steps = o.steps;
                for (int k = 0; k < kStatCount; ++k) stats[k] += lit::s_warp_stats[0][k];
            } else {
                const fus::LaneOut o = origin ? (rk ? fus::trace_warp<1, true>(P, true, x, y) : fus::trace_warp<0, true>(P, true, x, y))
                                              : (rk ? fus::trace_warp<1, false>(P, true, x, y) : fus::trace_warp<0, false>(P, true, x, y));
                rgba = o.rgba; tri = o.tri; steps = o.steps;
                for (int k = 0; k < kStatCount; ++k) stats[k] += fus::s_warp_stats[0][k];
            }
            const size_t idx = (size_t)y * (size_t)w + (size_t)x;
            memcpy(out_rgba + 4 * idx, &rgba, 16);
            if (out_hit) out_hit[idx] = tri;
            if (out_steps) out_steps[idx] = steps;
            steps_total += steps; ++traced;
        }
    stats[kStatSteps] += steps_total;
    stats[kStatTraced] += traced;
    if (stats9) memcpy(stats9, stats, sizeof stats);
    return 0;
}
