// detmath.cuh — the kernel's deterministic transcendentals ("det-math").
//
// WGSL leaves pow/sin/cos/tan/atan2/acos (ray.wgsl:402,419,444,459,550,586,623,632-634,657;
// sky.wgsl:20-24) to the shader compiler and driver.  This build pins them: every function is
// evaluated in binary64 with explicit fma() and fixed polynomials, then rounded once to binary32.
// Only correctly-rounded IEEE operations are used (+ - * / sqrt fma rint, conversions), so the
// result is a pure function of the input bits on any conforming machine — which is what lets
// tests compare the kernel bit-for-bit with a CPU evaluation of the same contract.
// Must be compiled with --fmad=false (no implicit contraction).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace detmath {

constexpr double kPi      = 3.141592653589793;
constexpr double kPio2Hi  = 1.5707963267948966;
constexpr double kPio2Lo  = 6.123233995736766e-17;
constexpr double kTwoOPi  = 0.6366197723675814;
constexpr double kLn2Hi   = 0.6931471805599453;
constexpr double kLn2Lo   = 2.3190468138462996e-17;
constexpr double kInvLn2  = 1.4426950408889634;
constexpr double kSqrt2   = 1.4142135623730951;

// Horner evaluation, highest coefficient first: ((c0*x + c1)*x + c2)...
template <int N>
__device__ __forceinline__ double horner(const double (&c)[N], double x)
{
    double p = c[0];
#pragma unroll
    for (int i = 1; i < N; ++i) p = fma(p, x, c[i]);
    return p;
}

// sin/cos for |x| < 1e9 (NaN beyond): reduce by pi/2 in two fma steps, Taylor to r^15 / r^16.
__device__ __forceinline__ void sincos_d(double x, double &s, double &c)
{
    if (!(fabs(x) < 1.0e9)) { s = CUDART_NAN; c = CUDART_NAN; return; }
    const double kd = rint(x * kTwoOPi);
    double r = fma(-kd, kPio2Hi, x);
    r = fma(-kd, kPio2Lo, r);
    const long long q = (long long)kd;
    const double r2 = r * r;
    const double S[7] = { -1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0,
                          -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0 };
    const double Cc[8] = { 1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0,
                           1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5 };
    const double sr = fma(r * r2, horner(S, r2), r);
    const double cr = fma(horner(Cc, r2), r2, 1.0);
    switch ((int)(q & 3)) {
    case 0:  s = sr;  c = cr;  break;
    case 1:  s = cr;  c = -sr; break;
    case 2:  s = -sr; c = -cr; break;
    default: s = -cr; c = sr;  break;
    }
}

__device__ __forceinline__ double atan_tab(int i)
{
    // atan(i/8), i = 0..8
    switch (i) {
    case 0: return 0.0;
    case 1: return 0.12435499454676144;
    case 2: return 0.24497866312686414;
    case 3: return 0.35877067027057225;
    case 4: return 0.4636476090008061;
    case 5: return 0.5585993153435624;
    case 6: return 0.6435011087932844;
    case 7: return 0.7188299996216245;
    default: return 0.7853981633974483;
    }
}

__device__ __forceinline__ bool sign_bit(double v) { return __double2hiint(v) < 0; }

// atan2: a = min/max in [0,1]; split at multiples of 1/8; odd series through t^13; octant fix-up.
__device__ __forceinline__ double atan2_d(double y, double x)
{
    if (x != x || y != y) return CUDART_NAN;
    const double ax = fabs(x), ay = fabs(y);
    const double mx = ax > ay ? ax : ay;
    const double mn = ax > ay ? ay : ax;
    double a;
    if (mx == 0.0) a = 0.0;
    else if (mx == CUDART_INF) a = (mn == CUDART_INF) ? 1.0 : 0.0;
    else a = mn / mx;
    const int idx = (int)(a * 8.0 + 0.5);
    const double cc = (double)idx * 0.125;
    const double t = (a - cc) / fma(a, cc, 1.0);
    const double t2 = t * t;
    const double A[6] = { 1.0 / 13.0, -1.0 / 11.0, 1.0 / 9.0, -1.0 / 7.0, 1.0 / 5.0, -1.0 / 3.0 };
    double r = atan_tab(idx) + fma(t * t2, horner(A, t2), t);
    if (ay > ax) r = kPio2Hi - r;
    if (sign_bit(x)) r = kPi - r;
    return sign_bit(y) ? -r : r;
}

// pow(x,y) = exp(y ln x): ln via 2 atanh((m-1)/(m+1)) through s^23, exp via Taylor through r^13.
__device__ __forceinline__ double pow_d(double x, double y)
{
    if (x != x || y != y) return CUDART_NAN;
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (x < 0.0) return CUDART_NAN;
    if (x == 0.0) return y > 0.0 ? 0.0 : CUDART_INF;
    if (x == CUDART_INF) return y > 0.0 ? CUDART_INF : 0.0;
    if (y == CUDART_INF) return x > 1.0 ? CUDART_INF : 0.0;
    if (y == -CUDART_INF) return x > 1.0 ? 0.0 : CUDART_INF;
    const unsigned long long ub = (unsigned long long)__double_as_longlong(x);
    int e = (int)((ub >> 52) & 0x7ffULL) - 1023;
    double m = __longlong_as_double((long long)((ub & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL));
    if (m > kSqrt2) { m *= 0.5; e += 1; }
    const double s = (m - 1.0) / (m + 1.0);
    const double s2 = s * s;
    const double L[11] = { 1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0,
                           1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0 };
    const double lnm = 2.0 * fma(s * s2, horner(L, s2), s);
    const double ed = (double)e;
    const double lnx = fma(ed, kLn2Hi, fma(ed, kLn2Lo, lnm));
    const double t = y * lnx;
    if (t > 90.0) return CUDART_INF;
    if (t < -105.0) return 0.0;
    const double kd = rint(t * kInvLn2);
    double r = fma(-kd, kLn2Hi, t);
    r = fma(-kd, kLn2Lo, r);
    const double E[14] = { 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                           1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0,
                           1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0 };
    const double q = horner(E, r);
    const int k = (int)kd;
    const double scale = __longlong_as_double((long long)(k + 1023) << 52);
    return q * scale;
}

// ---- binary32 API used by the kernels -------------------------------------------------------
__device__ __forceinline__ float sin_f(float x)  { double s, c; sincos_d((double)x, s, c); return (float)s; }
__device__ __forceinline__ float cos_f(float x)  { double s, c; sincos_d((double)x, s, c); return (float)c; }
__device__ __forceinline__ void sincos_f(float x, float &s, float &c)
{
    double sd, cd; sincos_d((double)x, sd, cd); s = (float)sd; c = (float)cd;
}
__device__ __forceinline__ float tan_f(float x)  { double s, c; sincos_d((double)x, s, c); return (float)(s / c); }
__device__ __forceinline__ float atan2_f(float y, float x) { return (float)atan2_d((double)y, (double)x); }
__device__ __forceinline__ float acos_f(float x)
{
    const double d = (double)x;
    if (!(fabs(d) <= 1.0)) return CUDART_NAN_F;
    return (float)atan2_d(sqrt((1.0 - d) * (1.0 + d)), d);
}
__device__ __forceinline__ float pow_f(float x, float y) { return (float)pow_d((double)x, (double)y); }
// constant exponents on the path: exact-product forms
__device__ __forceinline__ float pow2_f(float x) { return x * x; }
__device__ __forceinline__ float pow4_f(float x) { const double d = (double)x * (double)x; return (float)(d * d); }
__device__ __forceinline__ float pow5_f(float x) { const double d = (double)x * (double)x; return (float)((d * d) * (double)x); }

}  // namespace detmath
