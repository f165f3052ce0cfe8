/* bho_math.h — scalar math layer of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference (cleggacus/bhusie, src/renderer/shaders/ray.wgsl) leaves the rounding of
 * its WGSL built-ins (pow, sin, cos, tan, atan2, acos) to naga 0.19.2 + the Vulkan driver
 *  — there is nothing in /root/reference that pins them.  The oracle therefore
 * exists in two numeric flavours, selected at compile time:
 *
 *   BHO_FLAVOUR_STRICT   (default)  glibc libm float functions.  Neutral, independent of the
 *                                   product.  Also the CPU baseline that bench.py times.
 *   BHO_SHADOW                      float64 SHADOW: binary64 glibc functions; bh_oracle.c widens every arithmetic
 *                                   `float` to double with it (inputs, constants and outputs stay binary32).  Marks
 *                                   the pixels whose value is decided by f32 rounding (ill-conditioned).
 *   BHO_FLAVOUR_CONTRACT            "det-math": every transcendental is evaluated in IEEE
 *                                   binary64 with explicit fma() and a fixed polynomial, then
 *                                   rounded once to binary32.  Uses only operations that are
 *                                   correctly rounded on both x86-64 and sm_100a, so the CUDA
 *                                   kernel (which carries its own, separately written copy of
 *                                   the same numeric contract, bhusie_b200/csrc/detmath.cuh)
 *                                   can be compared BIT-EXACTLY against this flavour.
 *
 * The numeric contract (polynomial degrees, constants, reduction) is specified in DESIGN.md §4;
 * both files implement that spec.  tests/test_detmath_gpu.py compares them bit-for-bit.
 */
#ifndef BHO_MATH_H
#define BHO_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if !defined(BHO_FLAVOUR_CONTRACT) && !defined(BHO_FLAVOUR_STRICT)
#define BHO_FLAVOUR_STRICT 1
#endif

/* ---------------------------------------------------------------- det-math (binary64 core) */

static inline double bho_dm_bits2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t bho_dm_d2bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

#define BHO_DM_PI      3.141592653589793
#define BHO_DM_PIO2    1.5707963267948966
#define BHO_DM_PIO2_LO 6.123233995736766e-17
#define BHO_DM_2OPI    0.6366197723675814
#define BHO_DM_LN2     0.6931471805599453
#define BHO_DM_LN2_LO  2.3190468138462996e-17
#define BHO_DM_INVLN2  1.4426950408889634

/* sin and cos of a binary64 argument, |x| < 1e9 (else NaN).  Cody–Waite reduction by pi/2
 * with two fma steps, Taylor polynomials through r^15 (sin) and r^16 (cos) in Horner form. */
static inline void bho_dm_sincos(double x, double *s_out, double *c_out)
{
    if (!(fabs(x) < 1.0e9)) { *s_out = NAN; *c_out = NAN; return; }
    double kd = rint(x * BHO_DM_2OPI);
    double r = fma(-kd, BHO_DM_PIO2, x);
    r = fma(-kd, BHO_DM_PIO2_LO, r);
    long long q = (long long)kd;
    double r2 = r * r;
    /* sin(r) = r + r*r2*(S1 + r2*(S2 + ...)), S_i = (-1)^i / (2i+1)! */
    double ps = -1.0 / 1307674368000.0;              /* 1/15! */
    ps = fma(ps, r2, 1.0 / 6227020800.0);            /* 1/13! */
    ps = fma(ps, r2, -1.0 / 39916800.0);             /* 1/11! */
    ps = fma(ps, r2, 1.0 / 362880.0);                /* 1/9!  */
    ps = fma(ps, r2, -1.0 / 5040.0);                 /* 1/7!  */
    ps = fma(ps, r2, 1.0 / 120.0);                   /* 1/5!  */
    ps = fma(ps, r2, -1.0 / 6.0);                    /* 1/3!  */
    double sr = fma(r * r2, ps, r);
    /* cos(r) = 1 + r2*(C1 + r2*(C2 + ...)), C_i = (-1)^i / (2i)! */
    double pc = 1.0 / 20922789888000.0;              /* 1/16! */
    pc = fma(pc, r2, -1.0 / 87178291200.0);          /* 1/14! */
    pc = fma(pc, r2, 1.0 / 479001600.0);             /* 1/12! */
    pc = fma(pc, r2, -1.0 / 3628800.0);              /* 1/10! */
    pc = fma(pc, r2, 1.0 / 40320.0);                 /* 1/8!  */
    pc = fma(pc, r2, -1.0 / 720.0);                  /* 1/6!  */
    pc = fma(pc, r2, 1.0 / 24.0);                    /* 1/4!  */
    pc = fma(pc, r2, -0.5);                          /* 1/2!  */
    double cr = fma(pc, r2, 1.0);
    switch ((int)(q & 3)) {
    case 0: *s_out = sr;  *c_out = cr;  break;
    case 1: *s_out = cr;  *c_out = -sr; break;
    case 2: *s_out = -sr; *c_out = -cr; break;
    default: *s_out = -cr; *c_out = sr; break;
    }
}

static const double BHO_DM_ATAN_TAB[9] = {
    0.0, 0.12435499454676144, 0.24497866312686414, 0.35877067027057225, 0.4636476090008061,
    0.5585993153435624, 0.6435011087932844, 0.7188299996216245, 0.7853981633974483
};

/* atan2 in binary64: a = min/max in [0,1]; c = round(8a)/8; atan(a) = atan(c) + atan(t),
 * t = (a-c)/(1+a*c), |t| <= 1/16; odd Taylor series through t^13; octant fix-up. */
static inline double bho_dm_atan2(double y, double x)
{
    if (x != x || y != y) return NAN;
    double ax = fabs(x), ay = fabs(y);
    double mx = ax > ay ? ax : ay;
    double mn = ax > ay ? ay : ax;
    double a;
    if (mx == 0.0) a = 0.0;
    else if (mx == INFINITY) a = (mn == INFINITY) ? 1.0 : 0.0;
    else a = mn / mx;
    int idx = (int)(a * 8.0 + 0.5);
    double c = (double)idx * 0.125;
    double t = (a - c) / fma(a, c, 1.0);
    double t2 = t * t;
    double p = 1.0 / 13.0;
    p = fma(p, t2, -1.0 / 11.0);
    p = fma(p, t2, 1.0 / 9.0);
    p = fma(p, t2, -1.0 / 7.0);
    p = fma(p, t2, 1.0 / 5.0);
    p = fma(p, t2, -1.0 / 3.0);
    double r = BHO_DM_ATAN_TAB[idx] + fma(t * t2, p, t);
    if (ay > ax) r = BHO_DM_PIO2 - r;
    if (signbit(x)) r = BHO_DM_PI - r;
    return signbit(y) ? -r : r;
}

/* pow(x, y) for binary32 x, y evaluated in binary64: exp(y * ln x).
 * ln: x = 2^e * m, m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1); ln m = 2 s (1 + s^2/3 + ... + s^22/23).
 * exp: k = rint(t/ln2); r = t - k ln2 (two fma); Taylor through r^13; scale by 2^k. */
static inline double bho_dm_pow(double x, double y)
{
    if (x != x || y != y) return NAN;
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (x < 0.0) return NAN;
    if (x == 0.0) return y > 0.0 ? 0.0 : INFINITY;
    if (x == INFINITY) return y > 0.0 ? INFINITY : 0.0;
    if (y == INFINITY) return x > 1.0 ? INFINITY : 0.0;
    if (y == -INFINITY) return x > 1.0 ? 0.0 : INFINITY;
    uint64_t ub = bho_dm_d2bits(x);
    int e = (int)((ub >> 52) & 0x7ff) - 1023;
    double m = bho_dm_bits2d((ub & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, s2, 1.0 / 21.0);
    p = fma(p, s2, 1.0 / 19.0);
    p = fma(p, s2, 1.0 / 17.0);
    p = fma(p, s2, 1.0 / 15.0);
    p = fma(p, s2, 1.0 / 13.0);
    p = fma(p, s2, 1.0 / 11.0);
    p = fma(p, s2, 1.0 / 9.0);
    p = fma(p, s2, 1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0);
    p = fma(p, s2, 1.0 / 3.0);
    double lnm = 2.0 * fma(s * s2, p, s);
    double ed = (double)e;
    double lnx = fma(ed, BHO_DM_LN2, fma(ed, BHO_DM_LN2_LO, lnm));
    double t = y * lnx;
    if (t > 90.0) return INFINITY;
    if (t < -105.0) return 0.0;
    double kd = rint(t * BHO_DM_INVLN2);
    double r = fma(-kd, BHO_DM_LN2, t);
    r = fma(-kd, BHO_DM_LN2_LO, r);
    double q = 1.0 / 6227020800.0;                   /* 1/13! */
    q = fma(q, r, 1.0 / 479001600.0);
    q = fma(q, r, 1.0 / 39916800.0);
    q = fma(q, r, 1.0 / 3628800.0);
    q = fma(q, r, 1.0 / 362880.0);
    q = fma(q, r, 1.0 / 40320.0);
    q = fma(q, r, 1.0 / 5040.0);
    q = fma(q, r, 1.0 / 720.0);
    q = fma(q, r, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    int k = (int)kd;
    double scale = bho_dm_bits2d((uint64_t)(k + 1023) << 52);
    return q * scale;
}

/* ---------------------------------------------------------------- flavour dispatch (binary32 API) */

#if defined(BHO_SHADOW)

/* float64 shadow flavour (bh_oracle.c): glibc's binary64 functions on binary64 operands */
static inline double bho_sin(double x)  { return sin(x); }
static inline double bho_cos(double x)  { return cos(x); }
static inline double bho_tan(double x)  { return tan(x); }
static inline double bho_atan2(double y, double x) { return atan2(y, x); }
static inline double bho_acos(double x) { return acos(x); }
static inline double bho_pow(double x, double y) { return pow(x, y); }
static inline double bho_pow2(double x) { return pow(x, 2.0); }
static inline double bho_pow4(double x) { return pow(x, 4.0); }
static inline double bho_pow5(double x) { return pow(x, 5.0); }
static inline double bho_min(double a, double b) { return fmin(a, b); }
static inline double bho_max(double a, double b) { return fmax(a, b); }
static inline double bho_clamp(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

#elif defined(BHO_FLAVOUR_CONTRACT)

static inline float bho_sin(float x)  { double s, c; bho_dm_sincos((double)x, &s, &c); return (float)s; }
static inline float bho_cos(float x)  { double s, c; bho_dm_sincos((double)x, &s, &c); return (float)c; }
static inline float bho_tan(float x)  { double s, c; bho_dm_sincos((double)x, &s, &c); return (float)(s / c); }
static inline float bho_atan2(float y, float x) { return (float)bho_dm_atan2((double)y, (double)x); }
static inline float bho_acos(float x)
{
    double d = (double)x;
    if (!(fabs(d) <= 1.0)) return NAN;
    return (float)bho_dm_atan2(sqrt((1.0 - d) * (1.0 + d)), d);
}
static inline float bho_pow(float x, float y) { return (float)bho_dm_pow((double)x, (double)y); }
/* the three constant exponents on the path get exact-product forms (correctly rounded or
 * one binary64 rounding away from it), see DESIGN.md §4 */
static inline float bho_pow2(float x) { return x * x; }
static inline float bho_pow4(float x) { double d = (double)x * (double)x; return (float)(d * d); }
static inline float bho_pow5(float x) { double d = (double)x * (double)x; return (float)((d * d) * (double)x); }

#else /* BHO_FLAVOUR_STRICT: glibc */

static inline float bho_sin(float x)  { return sinf(x); }
static inline float bho_cos(float x)  { return cosf(x); }
static inline float bho_tan(float x)  { return tanf(x); }
static inline float bho_atan2(float y, float x) { return atan2f(y, x); }
static inline float bho_acos(float x) { return acosf(x); }
static inline float bho_pow(float x, float y) { return powf(x, y); }
static inline float bho_pow2(float x) { return powf(x, 2.0f); }
static inline float bho_pow4(float x) { return powf(x, 4.0f); }
static inline float bho_pow5(float x) { return powf(x, 5.0f); }

#endif

/* min/max/clamp: NaN-ignoring (C fminf/fmaxf == PTX min.f32/max.f32), DESIGN.md §4 */
#if !defined(BHO_SHADOW)
static inline float bho_min(float a, float b) { return fminf(a, b); }
static inline float bho_max(float a, float b) { return fmaxf(a, b); }
static inline float bho_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
#endif

#endif /* BHO_MATH_H */
