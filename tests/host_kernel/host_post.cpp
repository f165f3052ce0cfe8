// host_post.cpp — the PRODUCT's post-chain kernels (bhusie_b200/csrc/post_impl.cuh: bloom down / up, mix, ACES, FXAA) run on
// the CPU, thread by thread over the launch grid, so that every stage can be checked against the oracle without a GPU
// (tests/test_host_kernel.py).  These kernels have no warp collectives; the shim only supplies threadIdx / blockIdx and the
// one-IEEE-operation intrinsics.  TEST INFRASTRUCTURE ONLY.
#define BH_HOST_EMULATION 1
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
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

struct Idx { unsigned x, y, z; };
static Idx threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 };
static const Idx blockDim = { 256, 1, 1 }, gridDim = { 1, 1, 1 };
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
namespace tma {
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
#include "../../bhusie_b200/csrc/post_impl.cuh"
#undef BH_NUM_NS
#undef BH_FUSED
#define BH_NUM_NS fus
#define BH_FUSED 1
#include "../../bhusie_b200/csrc/ray_impl.cuh"
#include "../../bhusie_b200/csrc/post_impl.cuh"
#undef BH_NUM_NS
#undef BH_FUSED
}  // namespace bh

// One post pass (launch_post_pass of ray_kernels.cu: grid of 32x8-pixel blocks, 256 threads).  kind: bh_post_kind.
// in1 / in2: RGBA16F images (uint16 bits); out: RGBA16F, or RGBA8 for FXAA.  mode: 0 LITERAL, 1 FUSED.
extern "C" int bh_host_post_pass(int mode, int kind, const uint16_t *in1, int in1_w, int in1_h, const uint16_t *in2,
                                 void *out, int out_w, int out_h, float mix_ratio, const void *fxaa_details16)
{
    using namespace bh;
    PostParams P;
    memset(&P, 0, sizeof P);
    P.in1 = HalfImage{ reinterpret_cast<const uint2 *>(in1), in1_w, in1_h };
    P.in2 = HalfImage{ reinterpret_cast<const uint2 *>(in2), in1_w, in1_h };
    P.out = out; P.w = out_w; P.h = out_h;
    P.mix_ratio = mix_ratio;
    if (fxaa_details16) {
        bh_fxaa_details d;
        memcpy(&d, fxaa_details16, sizeof d);
        P.edge_min = d.edge_threshold_min; P.edge_max = d.edge_threshold_max; P.subpix = d.subpixel_quality; P.iterations = d.iterations;
    }
    const unsigned gx = (unsigned)((out_w + 31) / 32), gy = (unsigned)((out_h + 7) / 8);
    for (blockIdx.y = 0; blockIdx.y < gy; ++blockIdx.y)
        for (blockIdx.x = 0; blockIdx.x < gx; ++blockIdx.x)
            for (threadIdx.x = 0; threadIdx.x < 256; ++threadIdx.x) {
                switch (kind) {
                case BH_POST_BLOOM_DOWN: if (mode == 0) lit::bloom_down_kernel(P); else fus::bloom_down_kernel(P); break;
                case BH_POST_BLOOM_UP:   if (mode == 0) lit::bloom_up_kernel(P);   else fus::bloom_up_kernel(P); break;
                case BH_POST_MIX:        if (mode == 0) lit::mix_kernel(P);        else fus::mix_kernel(P); break;
                case BH_POST_HDR:        if (mode == 0) lit::hdr_kernel(P);        else fus::hdr_kernel(P); break;
                case BH_POST_FXAA:       if (mode == 0) lit::fxaa_kernel(P);       else fus::fxaa_kernel(P); break;
                default: return -22;
                }
            }
    return 0;
}
