// post_impl.cuh — the post chain that follows the ray path: bloom down/up, mix, ACES, FXAA.
// Included by ray_kernels.cu inside namespace lit / fus right after ray_impl.cuh (shares its madd / V3 layer).
// What each kernel computes is the fragment shader named on it (paths relative to the reference tree); the
// rasteriser is made explicit: one thread per output pixel, coordinate ((px+0.5)/W, (py+0.5)/H).  Implementation-
// defined choices (implicit LOD -> filter, Nearest/Linear rules, f16 / sRGB stores) are listed in DESIGN.md §8 and
// restated independently in oracle/bh_oracle_post.inc.
namespace BH_NUM_NS {

__device__ __forceinline__ float4 himg_texel(const HalfImage &t, int x, int y)
{
    const uint2 v = __ldg(t.px + (size_t)y * (size_t)t.w + (size_t)x);
    const __half2 lo = *reinterpret_cast<const __half2 *>(&v.x), hi = *reinterpret_cast<const __half2 *>(&v.y);
    const float2 a = __half22float2(lo), b = __half22float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 himg_nearest(const HalfImage &t, float u, float v)
{
    return himg_texel(t, tex_index(floorf(u * (float)t.w), t.w), tex_index(floorf(v * (float)t.h), t.h));
}
__device__ __forceinline__ float4 himg_linear_off(const HalfImage &t, float u, float v, int ox, int oy)
{
    const float x = msub(u, (float)t.w, 0.5f) + (float)ox;
    const float y = msub(v, (float)t.h, 0.5f) + (float)oy;
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int ix0 = tex_index(x0, t.w), ix1 = tex_index(x0 + 1.0f, t.w);
    const int iy0 = tex_index(y0, t.h), iy1 = tex_index(y0 + 1.0f, t.h);
    const float4 t00 = himg_texel(t, ix0, iy0), t10 = himg_texel(t, ix1, iy0), t01 = himg_texel(t, ix0, iy1), t11 = himg_texel(t, ix1, iy1);
    const float ux = 1.0f - fx, uy = 1.0f - fy;
    const float4 top = make_float4(lerp2(t00.x, t10.x, ux, fx), lerp2(t00.y, t10.y, ux, fx), lerp2(t00.z, t10.z, ux, fx), lerp2(t00.w, t10.w, ux, fx));
    const float4 bot = make_float4(lerp2(t01.x, t11.x, ux, fx), lerp2(t01.y, t11.y, ux, fx), lerp2(t01.z, t11.z, ux, fx), lerp2(t01.w, t11.w, ux, fx));
    return make_float4(lerp2(top.x, bot.x, uy, fy), lerp2(top.y, bot.y, uy, fy), lerp2(top.z, bot.z, uy, fy), lerp2(top.w, bot.w, uy, fy));
}
__device__ __forceinline__ float4 himg_linear(const HalfImage &t, float u, float v) { return himg_linear_off(t, u, v, 0, 0); }
__device__ __forceinline__ V3 rgb(float4 c) { return mk(c.x, c.y, c.z); }
__device__ __forceinline__ void store_h4(uint2 *dst, size_t o, float r, float g, float b, float a)
{
    const unsigned short hr = __half_as_ushort(__float2half_rn(r)), hg = __half_as_ushort(__float2half_rn(g));
    const unsigned short hb = __half_as_ushort(__float2half_rn(b)), ha = __half_as_ushort(__float2half_rn(a));
    dst[o] = make_uint2((unsigned)hr | ((unsigned)hg << 16), (unsigned)hb | ((unsigned)ha << 16));
}

// bloom_down.wgsl:32-60 — 13 Nearest taps (the pass renders at half the source resolution: minification)
__global__ void __launch_bounds__(256) bloom_down_kernel(const __grid_constant__ PostParams P)
{
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= P.w || py >= P.h) return;
    const HalfImage &t = P.in1;
    const float x = 1.0f / (float)t.w, y = 1.0f / (float)t.h;
    const float u = ((float)px + 0.5f) / (float)P.w, v = ((float)py + 0.5f) / (float)P.h;
    const float x2 = 2.0f * x, y2 = 2.0f * y;
    const V3 a = rgb(himg_nearest(t, u - x2, v + y2)), b = rgb(himg_nearest(t, u, v + y2)), c = rgb(himg_nearest(t, u + x2, v + y2));
    const V3 d = rgb(himg_nearest(t, u - x2, v)), e = rgb(himg_nearest(t, u, v)), f = rgb(himg_nearest(t, u + x2, v));
    const V3 g = rgb(himg_nearest(t, u - x2, v - y2)), h = rgb(himg_nearest(t, u, v - y2)), i = rgb(himg_nearest(t, u + x2, v - y2));
    const V3 j = rgb(himg_nearest(t, u - x, v + y)), k = rgb(himg_nearest(t, u + x, v + y));
    const V3 l = rgb(himg_nearest(t, u - x, v - y)), m = rgb(himg_nearest(t, u + x, v - y));
    V3 ds = e * 0.125f;
    ds = vmadd(((a + c) + g) + i, 0.03125f, ds);
    ds = vmadd(((b + d) + f) + h, 0.0625f, ds);
    ds = vmadd(((j + k) + l) + m, 0.125f, ds);
    store_h4(static_cast<uint2 *>(P.out), (size_t)py * (size_t)P.w + (size_t)px, ds.x, ds.y, ds.z, 1.0f);
}

// bloom_up.wgsl:31-54 — 9 Linear taps at fixed 0.005 offsets (magnification), tent weights
__global__ void __launch_bounds__(256) bloom_up_kernel(const __grid_constant__ PostParams P)
{
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= P.w || py >= P.h) return;
    const HalfImage &t = P.in1;
    const float x = 0.005f, y = 0.005f;
    const float u = ((float)px + 0.5f) / (float)P.w, v = ((float)py + 0.5f) / (float)P.h;
    const V3 a = rgb(himg_linear(t, u - x, v + y)), b = rgb(himg_linear(t, u, v + y)), c = rgb(himg_linear(t, u + x, v + y));
    const V3 d = rgb(himg_linear(t, u - x, v)), e = rgb(himg_linear(t, u, v)), f = rgb(himg_linear(t, u + x, v));
    const V3 g = rgb(himg_linear(t, u - x, v - y)), h = rgb(himg_linear(t, u, v - y)), i = rgb(himg_linear(t, u + x, v - y));
    V3 us = e * 4.0f;
    us = vmadd(((b + d) + f) + h, 2.0f, us);
    us = us + (((a + c) + g) + i);
    us = us * (1.0f / 16.0f);
    store_h4(static_cast<uint2 *>(P.out), (size_t)py * (size_t)P.w + (size_t)px, us.x, us.y, us.z, 1.0f);
}

// mix.wgsl:31-35
__global__ void __launch_bounds__(256) mix_kernel(const __grid_constant__ PostParams P)
{
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= P.w || py >= P.h) return;
    const float u = ((float)px + 0.5f) / (float)P.w, v = ((float)py + 0.5f) / (float)P.h;
    const float4 a = himg_linear(P.in1, u, v), b = himg_linear(P.in2, u, v);
    const float r = P.mix_ratio, k = 1.0f - r;
    store_h4(static_cast<uint2 *>(P.out), (size_t)py * (size_t)P.w + (size_t)px, madd(k, b.x, r * a.x), madd(k, b.y, r * a.y),
             madd(k, b.z, r * a.z), madd(k, b.w, r * a.w));
}

// hdr.wgsl:1-16 — ACES fit; mat3x3(...) lists columns
__device__ __forceinline__ V3 aces_tone_map(V3 hdr)
{
    const V3 m1c0 = mk(0.59719f, 0.07600f, 0.02840f), m1c1 = mk(0.35458f, 0.90834f, 0.13383f), m1c2 = mk(0.04823f, 0.01566f, 0.83777f);
    const V3 m2c0 = mk(1.60475f, -0.10208f, -0.00327f), m2c1 = mk(-0.53108f, 1.10813f, -0.07276f), m2c2 = mk(-0.07367f, -0.00605f, 1.07602f);
    const V3 v = vmadd(m1c2, hdr.z, vmadd(m1c1, hdr.y, m1c0 * hdr.x));
    const V3 a = mk(msub(v.x, v.x + 0.0245786f, 0.000090537f), msub(v.y, v.y + 0.0245786f, 0.000090537f), msub(v.z, v.z + 0.0245786f, 0.000090537f));
    const V3 b = mk(madd(v.x, madd(0.983729f, v.x, 0.4329510f), 0.238081f), madd(v.y, madd(0.983729f, v.y, 0.4329510f), 0.238081f),
                    madd(v.z, madd(0.983729f, v.z, 0.4329510f), 0.238081f));
    const V3 q = mk(a.x / b.x, a.y / b.y, a.z / b.z);
    const V3 o = vmadd(m2c2, q.z, vmadd(m2c1, q.y, m2c0 * q.x));
    return mk(clampf(o.x, 0.0f, 1.0f), clampf(o.y, 0.0f, 1.0f), clampf(o.z, 0.0f, 1.0f));
}

__global__ void __launch_bounds__(256) hdr_kernel(const __grid_constant__ PostParams P)
{
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= P.w || py >= P.h) return;
    const float u = ((float)px + 0.5f) / (float)P.w, v = ((float)py + 0.5f) / (float)P.h;
    const float4 c = himg_linear(P.in1, u, v);
    const V3 s = aces_tone_map(rgb(c));
    store_h4(static_cast<uint2 *>(P.out), (size_t)py * (size_t)P.w + (size_t)px, s.x, s.y, s.z, c.w);
}

// fxaa.wgsl:30-40,42-200
__device__ __forceinline__ float fxaa_quality(int q)
{
    switch (q) { case 5: return 1.5f; case 6: case 7: case 8: case 9: return 2.0f; case 10: return 4.0f; case 11: return 8.0f; default: return 1.0f; }
}
__device__ __forceinline__ float rgb2luma(V3 c) { return sqrtf(dot(c, mk(0.299f, 0.587f, 0.114f))); }
__device__ __forceinline__ unsigned srgb8(float lin)
{
    float c = clampf(lin, 0.0f, 1.0f);
    if (!(c == c)) c = 0.0f;
    const float s = c <= 0.0031308f ? 12.92f * c : msub(1.055f, detmath::pow_f(c, (float)(1.0 / 2.4)), 0.055f);
    return (unsigned)(int)floorf(madd(clampf(s, 0.0f, 1.0f), 255.0f, 0.5f));
}
__device__ __forceinline__ unsigned unorm8(float x)
{
    float c = clampf(x, 0.0f, 1.0f);
    if (!(c == c)) c = 0.0f;
    return (unsigned)(int)floorf(madd(c, 255.0f, 0.5f));
}

__global__ void __launch_bounds__(256) fxaa_kernel(const __grid_constant__ PostParams P)
{
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= P.w || py >= P.h) return;
    const HalfImage &t = P.in1;
    const float isx = 1.0f / (float)P.w, isy = 1.0f / (float)P.h;
    const float tu = ((float)px + 0.5f) * isx, tv = ((float)py + 0.5f) * isy;
    const float4 center = himg_linear(t, tu, tv);
    const float lumaCenter = rgb2luma(rgb(center));
    const float lumaDown = rgb2luma(rgb(himg_linear_off(t, tu, tv, 0, -1))), lumaUp = rgb2luma(rgb(himg_linear_off(t, tu, tv, 0, 1)));
    const float lumaLeft = rgb2luma(rgb(himg_linear_off(t, tu, tv, -1, 0))), lumaRight = rgb2luma(rgb(himg_linear_off(t, tu, tv, 1, 0)));
    const float lumaMin = fminf(lumaCenter, fminf(fminf(lumaDown, lumaUp), fminf(lumaLeft, lumaRight)));
    const float lumaMax = fmaxf(lumaCenter, fmaxf(fmaxf(lumaDown, lumaUp), fmaxf(lumaLeft, lumaRight)));
    const float lumaRange = lumaMax - lumaMin;
    V3 out = rgb(center);
    if (!(lumaRange < fmaxf(P.edge_min, lumaMax * P.edge_max))) {
        const float lumaDownLeft = rgb2luma(rgb(himg_linear_off(t, tu, tv, -1, -1))), lumaUpRight = rgb2luma(rgb(himg_linear_off(t, tu, tv, 1, 1)));
        const float lumaUpLeft = rgb2luma(rgb(himg_linear_off(t, tu, tv, -1, 1))), lumaDownRight = rgb2luma(rgb(himg_linear_off(t, tu, tv, 1, -1)));
        const float lumaDownUp = lumaDown + lumaUp, lumaLeftRight = lumaLeft + lumaRight;
        const float lumaLeftCorners = lumaDownLeft + lumaUpLeft, lumaDownCorners = lumaDownLeft + lumaDownRight;
        const float lumaRightCorners = lumaDownRight + lumaUpRight, lumaUpCorners = lumaUpRight + lumaUpLeft;
        const float edgeHorizontal = madd(fabsf(madd(-2.0f, lumaCenter, lumaDownUp)), 2.0f, fabsf(madd(-2.0f, lumaLeft, lumaLeftCorners))) +
                                     fabsf(madd(-2.0f, lumaRight, lumaRightCorners));
        const float edgeVertical = madd(fabsf(madd(-2.0f, lumaCenter, lumaLeftRight)), 2.0f, fabsf(madd(-2.0f, lumaUp, lumaUpCorners))) +
                                   fabsf(madd(-2.0f, lumaDown, lumaDownCorners));
        const bool isHorizontal = edgeHorizontal >= edgeVertical;
        float stepLength = isHorizontal ? isy : isx;
        const float luma1 = isHorizontal ? lumaDown : lumaLeft, luma2 = isHorizontal ? lumaUp : lumaRight;
        const float gradient1 = luma1 - lumaCenter, gradient2 = luma2 - lumaCenter;
        const bool is1Steepest = fabsf(gradient1) >= fabsf(gradient2);
        const float gradientScaled = 0.25f * fmaxf(fabsf(gradient1), fabsf(gradient2));
        float lumaLocalAverage;
        if (is1Steepest) { stepLength = -stepLength; lumaLocalAverage = 0.5f * (luma1 + lumaCenter); }
        else lumaLocalAverage = 0.5f * (luma2 + lumaCenter);
        float cu = tu, cv = tv, offx = 0.0f, offy = 0.0f;
        if (isHorizontal) { cv = madd(stepLength, 0.5f, cv); offx = isx; }
        else { cu = madd(stepLength, 0.5f, cu); offy = isy; }
        float u1 = cu - offx, v1 = cv - offy, u2 = cu + offx, v2 = cv + offy;
        float lumaEnd1 = rgb2luma(rgb(himg_linear(t, u1, v1))) - lumaLocalAverage;
        float lumaEnd2 = rgb2luma(rgb(himg_linear(t, u2, v2))) - lumaLocalAverage;
        bool reached1 = fabsf(lumaEnd1) >= gradientScaled, reached2 = fabsf(lumaEnd2) >= gradientScaled;
        bool reachedBoth = reached1 && reached2;
        if (!reached1) { u1 -= offx; v1 -= offy; }
        if (!reached2) { u2 += offx; v2 += offy; }
        if (!reachedBoth) {
            for (int i = 2; i < P.iterations; ++i) {
                if (!reached1) lumaEnd1 = rgb2luma(rgb(himg_linear(t, u1, v1))) - lumaLocalAverage;
                if (!reached2) lumaEnd2 = rgb2luma(rgb(himg_linear(t, u2, v2))) - lumaLocalAverage;
                reached1 = fabsf(lumaEnd1) >= gradientScaled; reached2 = fabsf(lumaEnd2) >= gradientScaled;
                reachedBoth = reached1 && reached2;
                const float q = fxaa_quality(i);
                if (!reached1) { u1 = nmadd(offx, q, u1); v1 = nmadd(offy, q, v1); }
                if (!reached2) { u2 = madd(offx, q, u2); v2 = madd(offy, q, v2); }
                if (reachedBoth) break;
            }
        }
        const float distance1 = isHorizontal ? tu - u1 : tv - v1, distance2 = isHorizontal ? u2 - tu : v2 - tv;
        const bool isDirection1 = distance1 < distance2;
        const float distanceFinal = fminf(distance1, distance2), edgeThickness = distance1 + distance2;
        const bool isLumaCenterSmaller = lumaCenter < lumaLocalAverage;
        const bool correctVariation1 = (lumaEnd1 < 0.0f) != isLumaCenterSmaller, correctVariation2 = (lumaEnd2 < 0.0f) != isLumaCenterSmaller;
        const bool correctVariation = isDirection1 ? correctVariation1 : correctVariation2;
        const float pixelOffset = -distanceFinal / edgeThickness + 0.5f;
        float finalOffset = correctVariation ? pixelOffset : 0.0f;
        const float lumaAverage = (1.0f / 12.0f) * (madd(2.0f, lumaDownUp + lumaLeftRight, lumaLeftCorners) + lumaRightCorners);
        const float sp1 = clampf(fabsf(lumaAverage - lumaCenter) / lumaRange, 0.0f, 1.0f);
        const float sp2 = madd(-2.0f, sp1, 3.0f) * sp1 * sp1;
        const float spFinal = sp2 * sp2 * P.subpix;
        finalOffset = fmaxf(finalOffset, spFinal);
        float fu = tu, fv = tv;
        if (isHorizontal) fv = madd(finalOffset, stepLength, fv); else fu = madd(finalOffset, stepLength, fu);
        out = rgb(himg_linear(t, fu, fv));
    }
    const unsigned r = srgb8(out.x), g = srgb8(out.y), b = srgb8(out.z), a = unorm8(center.w);
    static_cast<unsigned *>(P.out)[(size_t)py * (size_t)P.w + (size_t)px] = r | (g << 8) | (b << 16) | (a << 24);
}

}  // namespace BH_NUM_NS
