/* bh_oracle.c — CPU restatement of bhusie's per-pixel geodesic ray pass.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bhusie_b200/ may include, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the reference (cleggacus/bhusie) ships no tests, golden images or known-answer
 * vectors for this path, and it cannot be executed in this environment (no Rust toolchain, no
 * Vulkan ICD).  This file is a literal float32 transliteration of the WGSL
 * and Rust sources named below; it is the only executable statement of the reference semantics
 * available.  It is pinned by self-derived known-answer tests (tests/test_oracle_kat.py), by
 * brute-force cross-checks, and by golden frame digests committed under tests/golden/.
 *
 * Sources followed (paths relative to /root/reference/):
 *   src/renderer/shaders/ray.wgsl          whole file   -> bho_ray_pass and helpers
 *   src/renderer/shaders/sky.wgsl          :8-38        -> bho_sky_pass
 *   src/renderer/triangle.rs               :143-259     -> bho_build_bvh (BVH builder)
 *   src/renderer/triangle.rs               :268-285     -> ModelUniform byte layout (MU_* offsets)
 *   src/renderer/model.rs                  :7-87        -> bho_load_obj
 *   src/renderer/texture.rs                :16,32,61-69 -> RGBA8 unorm, bilinear, clamp-to-edge
 *   src/scene/camera.rs                    :66-73       -> CameraUniform bytes
 *   src/scene/blackhole.rs                 :37-51       -> BlackHoleUniform bytes
 *   src/renderer/pipelines/ray_pipeline.rs :3-14        -> RayDetails bytes
 *
 * Compile with -ffp-contract=off (no implicit FMA) so every +,-,*,/ and sqrtf is a single IEEE
 * binary32 operation in the order the WGSL expression tree gives.  Numeric flavour of the
 * transcendentals: see bho_math.h.  Build choices for implementation-defined WGSL corners are
 * marked "CHOICE" and listed in DESIGN.md §4.
 */
#include "bho_math.h"

#include <stdio.h>
#include <stdlib.h>
#include <errno.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Storage type of everything that crosses the oracle's API or lives in a reference byte layout: always binary32. */
typedef float f32;

#ifdef BHO_SHADOW
/* float64 SHADOW flavour (SURVEY.md §8c parity protocol): the SAME source with every arithmetic `float` widened to
 * binary64 and libm's double functions — inputs, constants and stored outputs stay binary32 (f32), so the only thing
 * that changes is the rounding of the intermediate operations.  A pixel whose shadow result differs from the strict
 * flavour's by more than the parity tolerance is ILL-CONDITIONED: its value is decided by f32 rounding (a ray grazing the
 * photon sphere, a disk / mesh / horizon edge, a star edge in the sky map), and no two conforming WGSL implementations
 * need agree on it.  tests/ and bench.py use it to classify the outliers of the device-vs-strict comparison.
 * The BVH builder and OBJ loader below are written on f32 and are unaffected. */
#define float double
#define sqrtf sqrt
#define floorf floor
#define fabsf fabs
#define truncf trunc
#define fminf fmin
#define fmaxf fmax
#define fmaf fma
#endif

/* ------------------------------------------------------------------ byte layouts */
#define MAX_MODEL_VERTICES 524288          /* triangle.rs:7, ray.wgsl:1 */
#define MU_SIZE        48234572ULL         /* sizeof(ModelUniform), triangle.rs:268-285 */
#define MU_POSITION    0
#define MU_VISIBLE     12
#define MU_POINTS      48ULL
#define MU_NORMALS     8388656ULL
#define MU_TRIANGLES   16777264ULL
#define MU_NODES       29360176ULL
#define MU_LOOKUP      46137392ULL

typedef struct { float x, y, z; } v3;

typedef struct {                            /* ray.wgsl:85-90 / triangle.rs:45-52 */
    f32 min_corner[3]; int32_t left_child;
    f32 max_corner[3]; int32_t obj_count;
} bho_node;

typedef struct { int32_t p1, p2, p3, n1, n2, n3; } bho_tri;   /* triangle.rs:54-63 */

typedef struct {                            /* ray.wgsl:25-34 */
    int32_t material_count, model_count; float time; int32_t integration_method;
    float step_size; int32_t max_iterations; float angle_division_threshold; int32_t highlight_interpolation;
} bho_details;

typedef struct { v3 position; v3 forward; float fov; } bho_camera;     /* ray.wgsl:41-45 */

typedef struct {                            /* ray.wgsl:112-123 */
    float inner_radius, outer_radius, rotation_speed, relativity_radius;
    v3 position; int32_t show_disk_texture;
    v3 normal; int32_t show_red_shift;
    v3 m0, m1, m2;                          /* rotation_matrix columns */
    float feather_amount;
} bho_black_hole;

typedef struct { const uint8_t *rgba; int32_t w, h; } bho_texture;

typedef struct {
    bho_texture color, disk, sky;           /* t_temp, t_disk, t_sky (ray.wgsl:13-17) */
    const uint8_t *models;                  /* ModelUniform[model_capacity], verbatim bytes */
    int32_t model_capacity;
} bho_scene;

/* counters (summed over the pixels of one pass) */
typedef struct {
    uint64_t steps;            /* integrator calls (ray.wgsl:525-531) */
    uint64_t loop_iters;       /* trace_ray loop iterations */
    uint64_t node_visits;      /* inner-node visits in trace_ray_model (two children each) */
    uint64_t tri_tests;        /* hit_triangle calls */
    uint64_t tex_samples;      /* bilinear samples (disk, LUT, sky) */
    uint64_t stack_overflow;   /* pushes at stack_location >= 19 (Q16) */
    uint64_t rk_reject;        /* e_max > 1 or NaN in next_ray_rk (Q5; reference would spin) */
    uint64_t px_traced, px_copied, px_interp;
    uint64_t bvh_calls;        /* trace_ray_model invocations */
} bho_counters;

typedef struct { bho_counters c; } tls_t;

/* ------------------------------------------------------------------ vec helpers (CHOICE: WGSL built-ins
 * expanded left to right, one IEEE op per node) */
static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }

/* madd(a,b,c) = a*b + c.  LITERAL flavours: two IEEE operations (what a non-contracting WGSL compiler
 * emits).  FUSED flavour: one fmaf — WGSL permits contracting x*y+z into fma (naga emits no
 * NoContraction decoration), and real drivers do.  Every helper below is written through madd with
 * the association order of the WGSL expression, so the literal flavours are unchanged by it:
 * a*b + (-c) == a*b - c and (-a)*b + c == c - a*b exactly in IEEE arithmetic. */
#if defined(BHO_FUSED)
static inline float madd(float a, float b, float c) { return fmaf(a, b, c); }
#else
static inline float madd(float a, float b, float c) { return a * b + c; }
#endif
static inline float msub(float a, float b, float c) { return madd(a, b, -c); }     /* a*b - c */
static inline float nmadd(float a, float b, float c) { return madd(-a, b, c); }    /* c - a*b */

static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 sscale(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
/* vec3 / f32.  LITERAL: three IEEE divisions.  FUSED: multiply by the correctly rounded reciprocal
 * (<= 1.5 ulp, inside WGSL's 2.5 ulp bound for `/`; what GPU drivers emit for vector/scalar). */
#if defined(BHO_FUSED)
static inline v3 vdivs(v3 a, float s) { const float r = 1.0f / s; return V(a.x * r, a.y * r, a.z * r); }
#else
static inline v3 vdivs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
#endif
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return madd(a.z, b.z, madd(a.y, b.y, a.x * b.x)); }
static inline v3 vcross(v3 a, v3 b)
{
    return V(msub(a.y, b.z, a.z * b.y), msub(a.z, b.x, a.x * b.z), msub(a.x, b.y, a.y * b.x));
}
/* p + v*s and acc + s*k (the same thing; two spellings keep call sites readable) */
static inline v3 vmadd(v3 v, float s, v3 p) { return V(madd(v.x, s, p.x), madd(v.y, s, p.y), madd(v.z, s, p.z)); }
static inline float vlength(v3 a) { return sqrtf(vdot(a, a)); }
static inline float vdistance(v3 a, v3 b) { return vlength(vsub(a, b)); }
static inline v3 vnormalize(v3 a) { return vdivs(a, vlength(a)); }
static inline v3 vmix(v3 a, v3 b, float t)   /* WGSL mix: e1*(1-e3) + e2*e3 */
{
    float u = 1.0f - t;
    return V(madd(b.x, t, a.x * u), madd(b.y, t, a.y * u), madd(b.z, t, a.z * u));
}
static inline v3 vmin(v3 a, v3 b) { return V(bho_min(a.x, b.x), bho_min(a.y, b.y), bho_min(a.z, b.z)); }
static inline v3 vmax(v3 a, v3 b) { return V(bho_max(a.x, b.x), bho_max(a.y, b.y), bho_max(a.z, b.z)); }
static inline float smoothstep_f(float lo, float hi, float x)
{
    float t = bho_clamp((x - lo) / (hi - lo), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
/* determinant(mat3x3(c0,c1,c2)) — CHOICE: cofactor expansion in GLM order */
static inline float det3(v3 c0, v3 c1, v3 c2)
{
    const float m0 = msub(c1.y, c2.z, c2.y * c1.z);
    const float m1 = msub(c0.y, c2.z, c2.y * c0.z);
    const float m2 = msub(c0.y, c1.z, c1.y * c0.z);
    return madd(c2.x, m2, nmadd(c1.x, m1, c0.x * m0));
}
static inline v3 mat3_mul(v3 c0, v3 c1, v3 c2, v3 v)   /* M*v = c0*v.x + c1*v.y + c2*v.z */
{
    return vmadd(c2, v.z, vmadd(c1, v.y, vscale(c0, v.x)));
}

/* ------------------------------------------------------------------ uniforms from raw bytes */
static float rd_f32(const uint8_t *p) { f32 f; memcpy(&f, p, 4); return f; }
static int32_t rd_i32(const uint8_t *p) { int32_t i; memcpy(&i, p, 4); return i; }
static v3 rd_v3(const uint8_t *p) { return V(rd_f32(p), rd_f32(p + 4), rd_f32(p + 8)); }

static bho_camera parse_camera(const uint8_t *b)      /* camera.rs:66-73 */
{
    bho_camera c; c.position = rd_v3(b); c.forward = rd_v3(b + 16); c.fov = rd_f32(b + 28); return c;
}
static bho_details parse_details(const uint8_t *b)    /* ray_pipeline.rs:5-14 */
{
    bho_details d;
    d.material_count = rd_i32(b); d.model_count = rd_i32(b + 4); d.time = rd_f32(b + 8);
    d.integration_method = rd_i32(b + 12); d.step_size = rd_f32(b + 16); d.max_iterations = rd_i32(b + 20);
    d.angle_division_threshold = rd_f32(b + 24); d.highlight_interpolation = rd_i32(b + 28);
    return d;
}
static bho_black_hole parse_black_hole(const uint8_t *b)  /* blackhole.rs:37-51 */
{
    bho_black_hole h;
    h.inner_radius = rd_f32(b); h.outer_radius = rd_f32(b + 4); h.rotation_speed = rd_f32(b + 8);
    h.relativity_radius = rd_f32(b + 12); h.position = rd_v3(b + 16); h.show_disk_texture = rd_i32(b + 28);
    h.normal = rd_v3(b + 32); h.show_red_shift = rd_i32(b + 44);
    h.m0 = rd_v3(b + 48); h.m1 = rd_v3(b + 64); h.m2 = rd_v3(b + 80); h.feather_amount = rd_f32(b + 96);
    return h;
}

/* ------------------------------------------------------------------ texture sampling (Q20)
 * RGBA8 unorm, bilinear, clamp-to-edge, level 0 (texture.rs:32,61-69).
 * CHOICE: Vulkan unnormalised-coordinate rule x = u*W - 0.5, i0 = floor(x), frac = x - i0,
 * indices clamped, two-stage lerp (top/bottom rows, then vertical) with weights (1-f), f. */
typedef struct { float r, g, b, a; } v4;

static inline int tex_index(float f, int n)
{
    /* float->int after clamping in float so NaN/inf cannot overflow the conversion */
    float c = bho_min(bho_max(f, -1.0f), (float)n);
    int i = (int)c;
    return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}
static inline v4 texel(const bho_texture *t, int x, int y)
{
    const uint8_t *p = t->rgba + 4 * ((size_t)y * (size_t)t->w + (size_t)x);
    v4 r = { (float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f };
    return r;
}
static v4 sample_bilinear(const bho_texture *t, float u, float v, tls_t *tls)
{
    tls->c.tex_samples++;
    float x = msub(u, (float)t->w, 0.5f);
    float y = msub(v, (float)t->h, 0.5f);
    float x0 = floorf(x), y0 = floorf(y);
    float fx = x - x0, fy = y - y0;
    int ix0 = tex_index(x0, t->w), ix1 = tex_index(x0 + 1.0f, t->w);
    int iy0 = tex_index(y0, t->h), iy1 = tex_index(y0 + 1.0f, t->h);
    v4 t00 = texel(t, ix0, iy0), t10 = texel(t, ix1, iy0);
    v4 t01 = texel(t, ix0, iy1), t11 = texel(t, ix1, iy1);
    float ux = 1.0f - fx, uy = 1.0f - fy;
    v4 top = { madd(t10.r, fx, t00.r * ux), madd(t10.g, fx, t00.g * ux), madd(t10.b, fx, t00.b * ux), madd(t10.a, fx, t00.a * ux) };
    v4 bot = { madd(t11.r, fx, t01.r * ux), madd(t11.g, fx, t01.g * ux), madd(t11.b, fx, t01.b * ux), madd(t11.a, fx, t01.a * ux) };
    v4 o = { madd(bot.r, fy, top.r * uy), madd(bot.g, fy, top.g * uy), madd(bot.b, fy, top.b * uy), madd(bot.a, fy, top.a * uy) };
    return o;
}

/* ------------------------------------------------------------------ shader structs */
typedef struct { v3 position, direction; } ray_t;

typedef struct {          /* ray.wgsl:92-98, plus the triangle index for the aux hit buffer */
    v3 color; float opacity; float t; v3 normal; int hit; int32_t tri;
} render_state;

static inline render_state rs_zero(float t)   /* WGSL zero-init, then .t = t_max */
{
    render_state r; memset(&r, 0, sizeof r); r.t = t; r.tri = -1; return r;
}

typedef struct {
    const bho_scene *scene;
    bho_camera camera;
    bho_details details;
    bho_black_hole bh;
} ctx_t;

static const float PI_F = 3.1415926f;   /* ray.wgsl:131, sky.wgsl:6 (Q19) */

/* Cash–Karp tableau, ray.wgsl:133-165.  The WGSL `const a_21 = 1.0/5.0;` declarations are
 * AbstractFloat const-expressions (binary64) converted to f32 where they meet an f32 operand;
 * CHOICE: follow that rule, including the (b_i - b_a_i) differences at ray.wgsl:435. */
#define A21 ((f32)(1.0 / 5.0))
#define A31 ((f32)(3.0 / 40.0))
#define A32 ((f32)(9.0 / 40.0))
#define A41 ((f32)(3.0 / 10.0))
#define A42 ((f32)(-9.0 / 10.0))
#define A43 ((f32)(6.0 / 5.0))
#define A51 ((f32)(-11.0 / 54.0))
#define A52 ((f32)(5.0 / 2.0))
#define A53 ((f32)(-70.0 / 27.0))
#define A54 ((f32)(35.0 / 27.0))
#define A61 ((f32)(1631.0 / 55296.0))
#define A62 ((f32)(175.0 / 512.0))
#define A63 ((f32)(575.0 / 13824.0))
#define A64 ((f32)(44275.0 / 110592.0))
#define A65 ((f32)(253.0 / 4096.0))
#define B1  (37.0 / 378.0)
#define B2  (0.0)
#define B3  (250.0 / 621.0)
#define B4  (125.0 / 594.0)
#define B5  (0.0)
#define B6  (512.0 / 1771.0)
#define BA1 (2825.0 / 27648.0)
#define BA2 (0.0)
#define BA3 (18575.0 / 48384.0)
#define BA4 (13525.0 / 55296.0)
#define BA5 (277.0 / 14336.0)
#define BA6 (1.0 / 4.0)

/* ------------------------------------------------------------------ ray.wgsl:725-766 hit_sphere */
static render_state hit_sphere(ray_t ray, float radius, v3 center, v3 color, float t_min, float t_max)
{
    render_state rs = rs_zero(t_max);
    v3 oc = vsub(ray.position, center);
    float a = vdot(ray.direction, ray.direction);
    float b = 2.0f * vdot(oc, ray.direction);
    float c = nmadd(radius, radius, vdot(oc, oc));
    float discriminant = msub(b, b, 4.0f * a * c);
    if (discriminant > 0.0f) {
        float t1 = (-b - sqrtf(discriminant)) / (2.0f * a);
        float t2 = (-b + sqrtf(discriminant)) / (2.0f * a);
        float t_closest = t_max;
        if (t1 > t_min && t1 < t_max) t_closest = t1;
        if (t2 > t_min && t2 < t_max && t2 < t_closest) t_closest = t2;
        if (t_closest < t_max && t_closest > t_min) {
            v3 ip = vmadd(ray.direction, t_closest, ray.position);
            rs.color = color;
            rs.opacity = 1.0f;
            rs.t = t_closest;
            rs.normal = vnormalize(vsub(ip, center));
            rs.hit = 1;
            return rs;
        }
    }
    return rs;
}

/* ------------------------------------------------------------------ ray.wgsl:668-701 hit_torus2d */
static render_state hit_torus2d(ray_t ray, float inner, float outer, v3 pos, v3 normal, float t_min, float t_max)
{
    float denom = vdot(normal, ray.direction);
    render_state rs = rs_zero(t_max);
    v3 dist = vsub(pos, ray.position);
    float t = vdot(dist, normal) / denom;
    if (t < t_max && t > t_min) {
        rs.normal = denom < 0.0f ? vneg(normal) : normal;
        v3 ip = vmadd(ray.direction, t, ray.position);
        float dc = vdistance(pos, ip);
        if (dc >= inner && dc <= outer) {
            rs.color = V(1.0f, 1.0f, 1.0f);
            rs.opacity = 1.0f;
            rs.t = t;
            rs.hit = 1;
            return rs;
        }
    }
    return rs;
}

/* ------------------------------------------------------------------ ray.wgsl:598-666 hit_black_hole */
static render_state hit_black_hole(const ctx_t *cx, ray_t ray, float t_min, float t_max, float total_distance, tls_t *tls)
{
    const bho_black_hole *bh = &cx->bh;
    render_state rs = hit_sphere(ray, 1.0f, bh->position, V(0, 0, 0), t_min, t_max);
    render_state disk_hit = hit_torus2d(ray, bh->inner_radius, bh->outer_radius, bh->position, bh->normal, t_min, t_max);

    if (disk_hit.hit && disk_hit.t < rs.t) {
        rs = disk_hit;
        v3 intersection = vmadd(ray.direction, rs.t, ray.position);
        float dist = vdistance(bh->position, intersection);
        /* disk_displacement (ray.wgsl:618) is dead */
        float disk_density = 1.0f - vlength(vdivs(intersection, bh->outer_radius));
        disk_density *= smoothstep_f(bh->inner_radius, bh->inner_radius + 1.0f, dist);
        disk_density *= 1.0f / sqrtf(dist);                       /* inverseSqrt — CHOICE: 1/sqrt */
        float optical_depth = bho_pow(30.0f * disk_density, 1.3f);

        rs.opacity = bho_clamp(optical_depth * 0.2f, 0.0f, 1.0f);
        rs.color = V(optical_depth, optical_depth, optical_depth);

        if (bh->show_disk_texture != 0) {
            float r = (dist - bh->inner_radius) / (bh->outer_radius - bh->inner_radius);
            v3 relative_pos = vdivs(vsub(intersection, bh->position), bh->outer_radius);
            v3 rotated_pos = mat3_mul(bh->m0, bh->m1, bh->m2, relative_pos);
            float angle = -bho_atan2(rotated_pos.z, rotated_pos.x);
            float arg = madd(cx->details.time, bh->rotation_speed, angle);
            float u = madd(bho_sin(arg), r, 1.0f) / 2.0f;
            float v = madd(bho_cos(arg), r, 1.0f) / 2.0f;
            v4 dc = sample_bilinear(&cx->scene->disk, u, v, tls);
            rs.opacity *= bho_clamp(madd(dc.a, 0.5f, 0.7f), 0.0f, 1.0f);
            rs.color = vmul(rs.color, V(dc.r * dc.a, dc.g * dc.a, dc.b * dc.a));
        }

        if (bh->show_red_shift != 0) {
            const float temp_max = 100000.0f, temp_min = 10000.0f, temp = 15000.0f;
            float y = 1.0f - (temp - temp_min) / (temp_max - temp_min);
            /* normalize(vec3(0,-1,0)) == (0,-1,0) exactly */
            v3 shift_vector = sscale(0.6f, vcross(vnormalize(intersection), V(0.0f, -1.0f, 0.0f)));
            float velocity = vdot(ray.direction, shift_vector);
            float doppler_shift = sqrtf((1.0f - velocity) / (1.0f + velocity));
            float gravitational_shift = sqrtf((1.0f - 2.0f / dist) / (1.0f - 2.0f / total_distance));
            float shift = bho_pow2(bho_clamp(gravitational_shift * doppler_shift, 0.0f, 1.0f));
            v4 sc = sample_bilinear(&cx->scene->color, shift, y, tls);
            rs.color = vmul(rs.color, V(sc.r, sc.g, sc.b));
        }
    }
    return rs;
}

/* ------------------------------------------------------------------ ray.wgsl:703-723 hit_aabb */
static float hit_aabb(ray_t ray, const bho_node *node, v3 offset)
{
    v3 inv = V(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z);
    v3 mn = vadd(V(node->min_corner[0], node->min_corner[1], node->min_corner[2]), offset);
    v3 mx = vadd(V(node->max_corner[0], node->max_corner[1], node->max_corner[2]), offset);
    v3 t1 = vmul(vsub(mn, ray.position), inv);
    v3 t2 = vmul(vsub(mx, ray.position), inv);
    v3 tmin = vmin(t1, t2), tmax = vmax(t1, t2);
    float t_min_axis = bho_max(bho_max(tmin.x, tmin.y), tmin.z);
    float t_max_axis = bho_min(bho_min(tmax.x, tmax.y), tmax.z);
    if (t_min_axis > t_max_axis || t_max_axis < 0.0f) return 1e8f;
    return t_min_axis;
}

/* ------------------------------------------------------------------ ray.wgsl:768-847 hit_triangle */
static render_state hit_triangle(ray_t ray, float t_min, float t_max, v3 pa, v3 pb, v3 pc, v3 n1, v3 n2, v3 n3)
{
    render_state rs = rs_zero(t_max);
    v3 edge_ab = vsub(pb, pa), edge_ac = vsub(pc, pa);
    v3 n = vnormalize(vcross(edge_ab, edge_ac));
    float ray_dot_tri = vdot(ray.direction, n);
    if (ray_dot_tri > 0.0f) {
        ray_dot_tri = ray_dot_tri * -1.0f;
        n = vscale(n, -1.0f);
    }
    if (fabsf(ray_dot_tri) < 0.00001f) return rs;

    float denominator = det3(ray.direction, vsub(pa, pb), vsub(pa, pc));
    if (fabsf(denominator) < 0.00001f) return rs;

    float u = det3(ray.direction, vsub(pa, ray.position), vsub(pa, pc)) / denominator;
    if (u < 0.0f || u > 1.0f) return rs;

    float v = det3(ray.direction, vsub(pa, pb), vsub(pa, ray.position)) / denominator;
    if (v < 0.0f || u + v > 1.0f) return rs;

    float t = det3(vsub(pa, ray.position), vsub(pa, pb), vsub(pa, pc)) / denominator;
    if (t > t_min && t < t_max) {
        v3 normal = vmadd(n3, v, vmadd(n2, u, sscale(1.0f - u - v, n1)));
        rs.normal = n;
        rs.color = V(madd(-normal.x, 0.5f, 0.5f), madd(-normal.y, 0.5f, 0.5f), madd(-normal.z, 0.5f, 0.5f));
        rs.opacity = 1.0f;
        rs.t = t;
        rs.hit = 1;
        return rs;
    }
    return rs;
}

/* ------------------------------------------------------------------ ray.wgsl:287-363 trace_ray_model
 * The WGSL keeps a stack of 19 full Node values.  Nodes are immutable, so a stack of node
 * indices is observationally identical.  Q16: naga's Restrict bounds policy clamps an
 * out-of-range dynamic index into a fixed-size array to the last element; CHOICE: reproduce
 * that (index clamped to 18 on both push and pop) and count every such push. */
#define BVH_STACK 19
static render_state trace_ray_model(const ctx_t *cx, ray_t ray, int model_index, float t_min, float t_max, tls_t *tls)
{
    const uint8_t *mu = cx->scene->models + (size_t)model_index * MU_SIZE;
    const bho_node *nodes = (const bho_node *)(mu + MU_NODES);
    const bho_tri *tris = (const bho_tri *)(mu + MU_TRIANGLES);
    const f32 *points = (const f32 *)(mu + MU_POINTS);
    const f32 *normals = (const f32 *)(mu + MU_NORMALS);
    const int32_t *lookup = (const int32_t *)(mu + MU_LOOKUP);
    v3 mpos = rd_v3(mu + MU_POSITION);

    tls->c.bvh_calls++;
    render_state closest = rs_zero(t_max);
    int32_t node = 0;
    int32_t stack[BVH_STACK];
    uint32_t stack_location = 0;

    for (;;) {
        int32_t obj_count = nodes[node].obj_count;
        int32_t contents = nodes[node].left_child;

        if (obj_count == 0) {
            tls->c.node_visits++;
            int32_t child_1 = contents, child_2 = contents + 1;
            float distance_1 = hit_aabb(ray, &nodes[child_1], mpos);
            float distance_2 = hit_aabb(ray, &nodes[child_2], mpos);
            if (distance_1 > distance_2) {
                float td = distance_1; distance_1 = distance_2; distance_2 = td;
                int32_t tc = child_1; child_1 = child_2; child_2 = tc;
            }
            if (distance_1 > closest.t) {
                if (stack_location == 0) break;
                stack_location -= 1;
                node = stack[stack_location > BVH_STACK - 1 ? BVH_STACK - 1 : stack_location];
            } else {
                node = child_1;
                if (distance_2 < closest.t) {
                    if (stack_location > BVH_STACK - 1) tls->c.stack_overflow++;
                    stack[stack_location > BVH_STACK - 1 ? BVH_STACK - 1 : stack_location] = child_2;
                    stack_location += 1;
                }
            }
        } else {
            for (int32_t i = 0; i < obj_count; i++) {
                int32_t index = lookup[contents + i];
                bho_tri ti = tris[index];
                v3 p1 = vadd(V(points[4 * ti.p1], points[4 * ti.p1 + 1], points[4 * ti.p1 + 2]), mpos);
                v3 p2 = vadd(V(points[4 * ti.p2], points[4 * ti.p2 + 1], points[4 * ti.p2 + 2]), mpos);
                v3 p3 = vadd(V(points[4 * ti.p3], points[4 * ti.p3 + 1], points[4 * ti.p3 + 2]), mpos);
                v3 n1 = V(normals[4 * ti.n1], normals[4 * ti.n1 + 1], normals[4 * ti.n1 + 2]);
                v3 n2 = V(normals[4 * ti.n2], normals[4 * ti.n2 + 1], normals[4 * ti.n2 + 2]);
                v3 n3 = V(normals[4 * ti.n3], normals[4 * ti.n3 + 1], normals[4 * ti.n3 + 2]);
                tls->c.tri_tests++;
                render_state rs = hit_triangle(ray, t_min, t_max, p1, p2, p3, n1, n2, n3);
                if (rs.hit && rs.t < closest.t) {
                    closest = rs;
                    closest.tri = index;
                }
            }
            if (stack_location == 0) break;
            stack_location -= 1;
            node = stack[stack_location > BVH_STACK - 1 ? BVH_STACK - 1 : stack_location];
        }
    }
    return closest;
}

/* ------------------------------------------------------------------ ray.wgsl:365-393 hit_ray */
static render_state hit_ray(const ctx_t *cx, ray_t ray, float t_min, float t_max, float ray_distance,
                            int render_triangles, int render_black_hole, tls_t *tls)
{
    render_state closest = rs_zero(t_max);
    /* Q13: hit_black_hole is evaluated unconditionally in the WGSL; it is pure, so when the
     * result is discarded we skip it and (to keep tex_samples a count of *used* samples) */
    if (render_black_hole) {
        render_state rs = hit_black_hole(cx, ray, t_min, t_max, ray_distance, tls);
        if (rs.hit && rs.t < closest.t) closest = rs;
    }
    if (render_triangles) {
        for (int i = 0; i < cx->details.model_count; i++) {
            const uint8_t *mu = cx->scene->models + (size_t)i * MU_SIZE;
            if (rd_i32(mu + MU_VISIBLE) != 0) {
                render_state rs = trace_ray_model(cx, ray, i, t_min, t_max, tls);
                if (rs.hit && rs.t < closest.t) {
                    closest = rs;
                    v3 light = vnormalize(V(0.2f, 0.2f, -1.0f));
                    float diffuse = vdot(closest.normal, light);
                    closest.color = vscale(closest.color, diffuse);
                }
            }
        }
    }
    return closest;
}

/* ------------------------------------------------------------------ ray.wgsl:401-403 f */
static inline v3 accel(const ctx_t *cx, v3 ray_pos, float h2, float r5)
{
    /* -1.5 * h2 * (rayPos - bh) / pow(dist, 5.0); pow(dist,5) is loop-invariant (pure), hoisted */
    return vdivs(sscale(-1.5f * h2, vsub(ray_pos, cx->bh.position)), r5);
}

typedef struct { float h, e_max; ray_t ray; } rk_state_t;

/* ------------------------------------------------------------------ ray.wgsl:405-465 next_ray_rk */
static rk_state_t next_ray_rk(const ctx_t *cx, rk_state_t st, tls_t *tls)
{
    ray_t ray = st.ray;
    float dist = vlength(vsub(ray.position, cx->bh.position));
    float h2 = bho_pow2(vlength(vcross(ray.position, ray.direction)));     /* Q1: p, not p-bh */
    float r5 = bho_pow5(dist);
    v3 dydx = accel(cx, ray.position, h2, r5);

    /* Q5: the accept loop body runs once; e_max > 1 would spin forever in the reference */
    float h = st.h;
    v3 k_1 = dydx;
    v3 k_2 = accel(cx, vmadd(sscale(A21, k_1), h, ray.position), h2, r5);
    v3 k_3 = accel(cx, vmadd(vmadd(k_2, A32, sscale(A31, k_1)), h, ray.position), h2, r5);
    /* Q4: a_43 multiplies k_2 */
    v3 k_4 = accel(cx, vmadd(vmadd(k_2, A43, vmadd(k_2, A42, sscale(A41, k_1))), h, ray.position), h2, r5);
    v3 k_5 = accel(cx, vmadd(vmadd(k_4, A54, vmadd(k_3, A53, vmadd(k_2, A52, sscale(A51, k_1)))), h, ray.position), h2, r5);
    v3 k_6 = accel(cx, vmadd(vmadd(k_5, A65, vmadd(k_4, A64, vmadd(k_3, A63, vmadd(k_2, A62, sscale(A61, k_1))))), h, ray.position), h2, r5);

    v3 esum = sscale((f32)(B1 - BA1), k_1);
    esum = vmadd(k_2, (f32)(B2 - BA2), esum);
    esum = vmadd(k_3, (f32)(B3 - BA3), esum);
    esum = vmadd(k_4, (f32)(B4 - BA4), esum);
    esum = vmadd(k_5, (f32)(B5 - BA5), esum);
    esum = vmadd(k_6, (f32)(B6 - BA6), esum);
    v3 e = sscale(h, esum);
    /* yscal = 1, eps = 1: x/1 == x */
    st.e_max = bho_max(bho_max(fabsf(e.x), fabsf(e.y)), fabsf(e.z));
    if (!(st.e_max <= 1.0f)) tls->c.rk_reject++;

    v3 dsum = sscale((f32)BA1, k_1);
    dsum = vmadd(k_2, (f32)BA2, dsum);
    dsum = vmadd(k_3, (f32)BA3, dsum);
    dsum = vmadd(k_4, (f32)BA4, dsum);
    dsum = vmadd(k_5, (f32)BA5, dsum);
    dsum = vmadd(k_6, (f32)BA6, dsum);
    st.ray.direction = vnormalize(vmadd(dsum, st.h, st.ray.direction));
    st.ray.position = vmadd(ray.direction, st.h, st.ray.position);                 /* Q6: OLD direction */

    if (st.e_max > 0.00002f) st.h *= 0.9f * bho_pow(st.e_max, -0.001f);
    else st.h *= 1.0001f;
    return st;
}

/* ------------------------------------------------------------------ ray.wgsl:467-480 next_ray_euler */
static ray_t next_ray_euler(const ctx_t *cx, ray_t ray, float step_size)
{
    float h2 = bho_pow2(vlength(vcross(ray.position, ray.direction)));
    float dist = vlength(vsub(ray.position, cx->bh.position));
    float r5 = bho_pow5(dist);
    ray.direction = vnormalize(vmadd(accel(cx, ray.position, h2, r5), step_size, ray.direction));
    ray.position = vmadd(ray.direction, step_size, ray.position);          /* Q8: NEW direction */
    return ray;
}

/* ------------------------------------------------------------------ ray.wgsl:255-261 / sky.wgsl:32-38 */
static inline void sky_uv(v3 dir, float *u, float *v)
{
    /* cartesian_to_spherical(dir.xzy) then the uv of ray.wgsl:586 / sky.wgsl:21 (Q19) */
    v3 c = V(dir.x, dir.z, dir.y);
    float theta = bho_atan2(sqrtf(madd(c.y, c.y, c.x * c.x)), c.z);
    float phi = bho_atan2(c.y, c.x);
    float uu = (phi + 2.6f * PI_F) / (2.0f * PI_F);
    float vv = (PI_F - theta) / PI_F;
    /* WGSL f32 %: e1 - e2*trunc(e1/e2) with e2 = 1 */
    *u = uu - 1.0f * truncf(uu / 1.0f);
    *v = vv - 1.0f * truncf(vv / 1.0f);
}

typedef struct { float r, g, b, a; int32_t hit_tri; uint32_t steps; } trace_out;

/* ------------------------------------------------------------------ ray.wgsl:482-596 trace_ray */
static trace_out trace_ray(const ctx_t *cx, ray_t ray, tls_t *tls)
{
    const v3 bh_position = cx->bh.position;
    const float bh_radius = cx->bh.relativity_radius;
    int relativity = 0;
    if (vdistance(ray.position, bh_position) < bh_radius) relativity = 1;

    const float t_max = 1e5f, t_min = 1e-8f;
    ray_t curr_ray = ray, prev_ray = ray;
    float color_amount = 1.0f;
    v3 color = V(0, 0, 0);
    float step_size = cx->details.step_size;
    rk_state_t rk_state = { step_size, 0.0f, curr_ray };      /* Q3: separate copy */
    const float ray_distance = vdistance(ray.position, bh_position);
    int hit = 0;
    int i = 0;
    float closest_to_bh = vdistance(curr_ray.position, bh_position);
    trace_out out; out.hit_tri = -1; out.steps = 0;

    for (; i < cx->details.max_iterations; i++) {
        render_state closest = rs_zero(t_max);
        tls->c.loop_iters++;

        if (relativity) {
            prev_ray = curr_ray;
            if (cx->details.integration_method == 0) {
                curr_ray = next_ray_euler(cx, curr_ray, step_size);
            } else {
                rk_state = next_ray_rk(cx, rk_state, tls);
                curr_ray = rk_state.ray;
                step_size = rk_state.h;
            }
            out.steps++; tls->c.steps++;
            float curr_distance_to_bh = vdistance(curr_ray.position, bh_position);
            if (curr_distance_to_bh < closest_to_bh) closest_to_bh = curr_distance_to_bh;
            prev_ray.direction = curr_ray.direction;                     /* Q7 */
            closest = hit_ray(cx, prev_ray, t_min, step_size, ray_distance, 0, 1, tls);
            if (curr_distance_to_bh > bh_radius) {
                relativity = 0;
                float feather_width = bh_radius * cx->bh.feather_amount;
                float feather_start = bh_radius - feather_width;
                float linear_mix_amount = bho_clamp((closest_to_bh - feather_start) / feather_width, 0.0f, 1.0f);
                float mix_amount = bho_pow2(linear_mix_amount);
                curr_ray.direction = vmix(curr_ray.direction, ray.direction, mix_amount);   /* Q9 */
            }
        } else {
            render_state rs = hit_ray(cx, curr_ray, t_min, t_max, ray_distance, 1, 0, tls);
            render_state hs = hit_sphere(prev_ray, bh_radius, bh_position, V(0, 0, 0), t_min, t_max);   /* Q10 */
            if (!hs.hit && !rs.hit) break;
            if (hs.hit && hs.t < rs.t) {
                curr_ray.position = vmadd(curr_ray.direction, hs.t, curr_ray.position);
                relativity = 1;
            } else {
                closest = rs;
            }
        }

        if (closest.hit) {
            curr_ray.position = vmadd(prev_ray.direction, closest.t, curr_ray.position);          /* Q11 */
            v3 cc = V(bho_clamp(closest.color.x, 0.0f, 1.0f), bho_clamp(closest.color.y, 0.0f, 1.0f),
                      bho_clamp(closest.color.z, 0.0f, 1.0f));
            float w = color_amount * closest.opacity;
            color = vmadd(cc, w, color);
            color_amount *= 1.0f - closest.opacity;
            hit = 1;
            if (closest.tri >= 0) out.hit_tri = closest.tri;
        }
        if (color_amount < 0.005f) break;
    }

    if (hit || i <= 5) {                                             /* Q12 */
        if (color_amount > 0.001f) {
            float u, v;
            sky_uv(curr_ray.direction, &u, &v);
            v4 s = sample_bilinear(&cx->scene->sky, u, v, tls);
            v3 miss = V(bho_pow4(s.r), bho_pow4(s.g), bho_pow4(s.b));
            color = vmadd(miss, color_amount, color);
        }
        out.r = color.x; out.g = color.y; out.b = color.z; out.a = 1.0f;
        return out;
    }
    out.r = curr_ray.direction.x; out.g = curr_ray.direction.y; out.b = curr_ray.direction.z; out.a = 0.0f;
    return out;
}

/* ------------------------------------------------------------------ ray.wgsl:269-285 create_ray */
static ray_t create_ray(const ctx_t *cx, int px, int py, int sw, int sh)
{
    int sm = (sw - 1) < (sh - 1) ? (sw - 1) : (sh - 1);
    float increment = 1.0f / (float)sm;
    float posx = 2.0f * ((float)px - (float)(sw - 1) / 2.0f) * increment;
    float posy = 2.0f * ((float)py - (float)(sh - 1) / 2.0f) * increment;
    v3 plane_up = V(0.0f, -1.0f, 0.0f);
    v3 right = vnormalize(vcross(cx->camera.forward, plane_up));
    v3 up = vnormalize(vcross(cx->camera.forward, right));
    float fov_factor = 1.0f / bho_tan(cx->camera.fov / 2.0f);
    v3 d = vmadd(cx->camera.forward, fov_factor, vmadd(up, posy, sscale(posx, right)));               /* Q21 */
    ray_t r; r.position = cx->camera.position; r.direction = vnormalize(d);
    return r;
}

#ifdef BHO_SHADOW
/* Conditioning probe of the shadow flavour: rotate every camera ray by `g_shadow_perturb` radians (about y, and 0.618x
 * that about x) before it is traced.  |shadow(delta) - shadow(0)| > tolerance for a delta of the size of the rounding
 * error a binary32 integration accumulates (measured: median 2e-7, 99th percentile 8e-7 on the exit direction) marks a
 * pixel whose value no binary32 implementation can be expected to reproduce to that tolerance. */
static double g_shadow_perturb = 0.0;
void bho_shadow_set_perturbation(double radians) { g_shadow_perturb = radians; }
static ray_t perturb_ray(ray_t r)
{
    if (g_shadow_perturb != 0.0) {
        const double a = g_shadow_perturb, b = 0.618 * g_shadow_perturb;
        v3 d = r.direction;
        d = V(d.x + a * d.z, d.y, d.z - a * d.x);
        d = V(d.x, d.y + b * d.z, d.z - b * d.y);
        r.direction = vnormalize(d);
    }
    return r;
}
#else
#define perturb_ray(r) (r)
#endif

/* ------------------------------------------------------------------ ray.wgsl:263-267 angle_between */
static inline float angle_between(v3 a, v3 b)
{
    float d = vdot(a, b);
    float c = d / (vlength(a) * vlength(b));
    return bho_acos(c);
}

/* textureLoad on the previous level; CHOICE (Q18): out-of-range coordinates return a zero texel */
static inline v4 prev_load(const f32 *prev, int pw, int ph, int x, int y)
{
    v4 z = { 0, 0, 0, 0 };
    if (x < 0 || y < 0 || x >= pw || y >= ph) return z;
    const f32 *p = prev + 4 * ((size_t)y * (size_t)pw + (size_t)x);
    v4 r = { p[0], p[1], p[2], p[3] };
    return r;
}

static void add_counters(bho_counters *a, const bho_counters *b)
{
    a->steps += b->steps; a->loop_iters += b->loop_iters; a->node_visits += b->node_visits;
    a->tri_tests += b->tri_tests; a->tex_samples += b->tex_samples; a->stack_overflow += b->stack_overflow;
    a->rk_reject += b->rk_reject; a->px_traced += b->px_traced; a->px_copied += b->px_copied;
    a->px_interp += b->px_interp; a->bvh_calls += b->bvh_calls;
}

/* ------------------------------------------------------------------ ray.wgsl:167-243 main
 * One call = one RayPipeline::pass (ray_pipeline.rs:301-309) over rows [row_begin,row_end).
 * prev == NULL (or 1x1) is the base case.  Aux outputs are nullable.
 * out_class: 0 traced(base) 1 copied 2 interpolated 3 traced(fine level). */
int bho_ray_pass(const bho_scene *scene, int32_t w, int32_t h,
                 const f32 *prev, int32_t pw, int32_t ph,
                 const uint8_t *camera32, const uint8_t *black_hole132, const uint8_t *details32,
                 int32_t row_begin, int32_t row_end,
                 f32 *out_rgba, int32_t *out_hit, uint32_t *out_steps, uint8_t *out_class,
                 bho_counters *counters, int32_t nthreads)
{
    if (!scene || !camera32 || !black_hole132 || !details32 || !out_rgba) return -EINVAL;
    if (w < 2 || h < 2 || row_begin < 0 || row_end > h || row_begin > row_end) return -EINVAL;
    ctx_t cx;
    cx.scene = scene;
    cx.camera = parse_camera(camera32);
    cx.bh = parse_black_hole(black_hole132);
    cx.details = parse_details(details32);
    if (cx.details.model_count < 0 || cx.details.model_count > scene->model_capacity) return -EINVAL;
    int base = (prev == NULL) || (pw == 1 && ph == 1);
    if (!base && (pw < 2 || ph < 2)) return -EINVAL;
    bho_counters total; memset(&total, 0, sizeof total);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif

#pragma omp parallel num_threads(nthreads)
    {
        tls_t tls; memset(&tls, 0, sizeof tls);
#pragma omp for schedule(dynamic, 1)
        for (int32_t y = row_begin; y < row_end; y++) {
            for (int32_t x = 0; x < w; x++) {
                size_t o = (size_t)y * (size_t)w + (size_t)x;
                trace_out to; to.hit_tri = -1; to.steps = 0;
                uint8_t cls;
                if (base) {
                    to = trace_ray(&cx, perturb_ray(create_ray(&cx, x, y, w, h)), &tls);
                    cls = 0; tls.c.px_traced++;
                } else {
                    int sfx = (w - 1) / (pw - 1), sfy = (h - 1) / (ph - 1);
                    float rx = (float)pw / (float)(w + (sfx - 1));
                    float ry = (float)ph / (float)(h + (sfy - 1));
                    float ppx = (float)x * rx, ppy = (float)y * ry;
                    float tlx = floorf(ppx), tly = floorf(ppy);
                    v4 c_tl = prev_load(prev, pw, ph, (int)tlx, (int)tly);
                    if (fabsf(tlx - ppx) < 0.001f && fabsf(tly - ppy) < 0.001f) {
                        to.r = c_tl.r; to.g = c_tl.g; to.b = c_tl.b; to.a = c_tl.a;
                        cls = 1; tls.c.px_copied++;
                    } else {
                        v4 c_bl = prev_load(prev, pw, ph, (int)(tlx + 0.0f), (int)(tly + 1.0f));
                        v4 c_tr = prev_load(prev, pw, ph, (int)(tlx + 1.0f), (int)(tly + 0.0f));
                        v4 c_br = prev_load(prev, pw, ph, (int)(tlx + 1.0f), (int)(tly + 1.0f));
                        v3 tl = V(c_tl.r, c_tl.g, c_tl.b), tr = V(c_tr.r, c_tr.g, c_tr.b);
                        v3 bl = V(c_bl.r, c_bl.g, c_bl.b), br = V(c_br.r, c_br.g, c_br.b);
                        /* spherical_to_cartesian results (ray.wgsl:203-206) are dead */
                        float a0 = angle_between(bl, tl), a1 = angle_between(br, tr);
                        float a2 = angle_between(tl, tr), a3 = angle_between(bl, br);
                        float thr = cx.details.angle_division_threshold;
                        if (c_tl.a == 0.0f && c_tr.a == 0.0f && c_bl.a == 0.0f && c_br.a == 0.0f &&
                            a0 < thr && a1 < thr && a2 < thr && a3 < thr) {
                            float tx = ppx - tlx, ty = ppy - tly;
                            v3 uv_t = vmix(tl, tr, tx), uv_b = vmix(bl, br, tx);
                            v3 p = vmix(uv_t, uv_b, ty);
                            to.r = p.x; to.g = p.y; to.b = p.z; to.a = 0.0f;
                            cls = 2; tls.c.px_interp++;
                        } else {
                            to = trace_ray(&cx, perturb_ray(create_ray(&cx, x, y, w, h)), &tls);
                            cls = 3; tls.c.px_traced++;
                        }
                    }
                }
                out_rgba[4 * o + 0] = to.r; out_rgba[4 * o + 1] = to.g;
                out_rgba[4 * o + 2] = to.b; out_rgba[4 * o + 3] = to.a;
                if (out_hit) out_hit[o] = to.hit_tri;
                if (out_steps) out_steps[o] = to.steps;
                if (out_class) out_class[o] = cls;
            }
        }
#pragma omp critical
        add_counters(&total, &tls.c);
    }
    if (counters) *counters = total;
    return 0;
}

/* ------------------------------------------------------------------ float -> binary16, round to nearest even */
static uint16_t f32_to_f16(f32 f)
{
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (ax > 0x7f800000u ? 0x0200u | ((ax >> 13) & 0x3ffu) : 0));
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                /* rounds to inf (>= 65520) */
    if (ax < 0x33000001u) return (uint16_t)sign;                               /* <= 2^-25 rounds to zero */
    int32_t e = (int32_t)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    int shift = (e < -14) ? (13 + (-14 - e)) : 13;
    uint32_t half_m = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_m & 1u))) half_m++;
    uint32_t out;
    if (e < -14) out = half_m;                                                 /* subnormal (may carry into normal) */
    else out = ((uint32_t)(e + 15) << 10) + (half_m - 0x400u);                /* carry propagates into exponent */
    return (uint16_t)(sign | out);
}

/* ------------------------------------------------------------------ sky.wgsl:8-30 main
 * One call = one SkyPipeline::pass (sky_pipeline.rs:140-148).  out_f32 (RGBA32F, before the
 * f16 store) and out_f16 (the reference's Rgba16Float, sky_pipeline.rs:34) are each nullable. */
int bho_sky_pass(const bho_scene *scene, int32_t w, int32_t h, const f32 *prev,
                 int32_t row_begin, int32_t row_end,
                 f32 *out_f32, uint16_t *out_f16, bho_counters *counters, int32_t nthreads)
{
    if (!scene || !prev || w < 1 || h < 1 || row_begin < 0 || row_end > h || row_begin > row_end) return -EINVAL;
    bho_counters total; memset(&total, 0, sizeof total);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        tls_t tls; memset(&tls, 0, sizeof tls);
#pragma omp for schedule(dynamic, 4)
        for (int32_t y = row_begin; y < row_end; y++) {
            for (int32_t x = 0; x < w; x++) {
                size_t o = (size_t)y * (size_t)w + (size_t)x;
                const f32 *p = prev + 4 * o;
                float r, g, b, a;
                if (p[3] == 0.0f) {
                    float u, v;
                    sky_uv(V(p[0], p[1], p[2]), &u, &v);
                    v4 s = sample_bilinear(&scene->sky, u, v, &tls);
                    r = bho_pow4(s.r); g = bho_pow4(s.g); b = bho_pow4(s.b); a = 1.0f;
                } else {
                    r = p[0]; g = p[1]; b = p[2]; a = p[3];
                }
                if (out_f32) { out_f32[4 * o] = r; out_f32[4 * o + 1] = g; out_f32[4 * o + 2] = b; out_f32[4 * o + 3] = a; }
                if (out_f16) {
                    out_f16[4 * o] = f32_to_f16((f32)r); out_f16[4 * o + 1] = f32_to_f16((f32)g);
                    out_f16[4 * o + 2] = f32_to_f16((f32)b); out_f16[4 * o + 3] = f32_to_f16((f32)a);
                }
            }
        }
#pragma omp critical
        add_counters(&total, &tls.c);
    }
    if (counters) *counters = total;
    return 0;
}

/* ------------------------------------------------------------------ triangle.rs:143-259 BVH builder
 * Operates in place on a verbatim ModelUniform blob whose points/triangles are already filled.
 * Literal recursion: children allocated as a consecutive pair, left subtree fully built first. */
typedef struct {
    f32 *points; bho_tri *tris; bho_node *nodes; int32_t *lookup; size_t nodes_used; int max_depth;
} bvh_build_t;

static void bvh_update_bounds(bvh_build_t *b, size_t ni)       /* triangle.rs:159-194 */
{
    bho_node *node = &b->nodes[ni];
    const f32 flt_max = 3.40282347e+38f;                         /* f32::MAX / f32::MIN */
    for (int a = 0; a < 3; a++) { node->min_corner[a] = flt_max; node->max_corner[a] = -flt_max; }
    for (int32_t i = 0; i < node->obj_count; i++) {
        const bho_tri *t = &b->tris[b->lookup[node->left_child + i]];
        const int32_t idx[3] = { t->p1, t->p2, t->p3 };
        for (int k = 0; k < 3; k++) {
            const f32 *p = b->points + 4 * (size_t)idx[k];
            for (int a = 0; a < 3; a++) {
                /* Rust f32::min/max: NaN-ignoring, same as fminf/fmaxf */
                node->min_corner[a] = fminf(node->min_corner[a], p[a]);
                node->max_corner[a] = fmaxf(node->max_corner[a], p[a]);
            }
        }
    }
}

static void bvh_subdivide(bvh_build_t *b, size_t ni, int depth)   /* triangle.rs:196-259 */
{
    if (depth > b->max_depth) b->max_depth = depth;
    if (b->nodes[ni].obj_count <= 2) return;
    f32 extent[3];
    for (int a = 0; a < 3; a++) extent[a] = b->nodes[ni].max_corner[a] - b->nodes[ni].min_corner[a];
    int axis = 0;
    if (extent[1] > extent[axis]) axis = 1;
    if (extent[2] > extent[axis]) axis = 2;
    f32 split_position = b->nodes[ni].min_corner[axis] + extent[axis] / 2.0f;

    int32_t i = b->nodes[ni].left_child;
    int32_t j = i + b->nodes[ni].obj_count - 1;
    while (i <= j) {
        const bho_tri *t = &b->tris[b->lookup[i]];
        f32 c = (b->points[4 * (size_t)t->p1 + axis] + b->points[4 * (size_t)t->p2 + axis] + b->points[4 * (size_t)t->p3 + axis]) / 3.0f;
        if (c < split_position) {
            i += 1;
        } else {
            int32_t tmp = b->lookup[i]; b->lookup[i] = b->lookup[j]; b->lookup[j] = tmp;
            j -= 1;
        }
    }
    int32_t left_count = i - b->nodes[ni].left_child;
    if (left_count == 0 || left_count == b->nodes[ni].obj_count) return;

    size_t left = b->nodes_used++;
    size_t right = b->nodes_used++;
    b->nodes[left].left_child = b->nodes[ni].left_child;
    b->nodes[left].obj_count = left_count;
    b->nodes[right].left_child = i;
    b->nodes[right].obj_count = b->nodes[ni].obj_count - left_count;
    b->nodes[ni].left_child = (int32_t)left;
    b->nodes[ni].obj_count = 0;
    bvh_update_bounds(b, left);
    bvh_update_bounds(b, right);
    bvh_subdivide(b, left, depth + 1);
    bvh_subdivide(b, right, depth + 1);
}

/* returns nodes_used (>0) or a negative errno; max_depth_out nullable */
int64_t bho_build_bvh(uint8_t *model_uniform, int32_t triangle_count, int32_t *max_depth_out)
{
    if (!model_uniform || triangle_count < 0 || triangle_count > MAX_MODEL_VERTICES) return -EINVAL;
    bvh_build_t b;
    b.points = (f32 *)(model_uniform + MU_POINTS);
    b.tris = (bho_tri *)(model_uniform + MU_TRIANGLES);
    b.nodes = (bho_node *)(model_uniform + MU_NODES);
    b.lookup = (int32_t *)(model_uniform + MU_LOOKUP);
    b.nodes_used = 0; b.max_depth = 0;
    for (int32_t i = 0; i < triangle_count; i++) b.lookup[i] = i;     /* triangle.rs:146-148 */
    b.nodes[0].left_child = 0;
    b.nodes[0].obj_count = triangle_count;
    b.nodes_used = 1;
    bvh_update_bounds(&b, 0);
    bvh_subdivide(&b, 0, 0);
    if (max_depth_out) *max_depth_out = b.max_depth;
    return (int64_t)b.nodes_used;
}

/* ------------------------------------------------------------------ model.rs:7-87 load_model
 * tobj 4.0.2 (Cargo.lock:2951, un-vendored) with LoadOptions::default(): separate position and
 * normal index streams, positions/normals re-indexed in first-use order per object, f32 parse
 * (correctly rounded — strtof is too).  Restated for the OBJ subset bhusie's asset uses
 * (`v`, `vn`, `f a//b` / `f a` / `f a/t/n`, triangles only).  Fills a zeroed ModelUniform blob:
 * position (-10,0,30) and visible=1 (triangle.rs:100,108), points scaled (0.5,-0.5,0.5), w=0,
 * per-face normals when the file has none, then the BVH.  Returns triangle count or -errno. */
typedef struct { int32_t *keys; int32_t *vals; size_t cap; } imap_t;
static int imap_init(imap_t *m, size_t cap) {
    m->cap = 1; while (m->cap < cap * 2) m->cap <<= 1;
    m->keys = (int32_t *)malloc(m->cap * sizeof(int32_t)); m->vals = (int32_t *)malloc(m->cap * sizeof(int32_t));
    if (!m->keys || !m->vals) return -1;
    for (size_t i = 0; i < m->cap; i++) m->keys[i] = -1;
    return 0;
}
static void imap_free(imap_t *m) { free(m->keys); free(m->vals); }
static int32_t imap_get_or_add(imap_t *m, int32_t key, int32_t next, int *added) {
    size_t h = ((uint32_t)key * 2654435761u) & (m->cap - 1);
    while (m->keys[h] != -1) { if (m->keys[h] == key) { *added = 0; return m->vals[h]; } h = (h + 1) & (m->cap - 1); }
    m->keys[h] = key; m->vals[h] = next; *added = 1; return next;
}

int64_t bho_load_obj(const char *path, uint8_t *model_uniform, int32_t *point_count_out, int64_t *nodes_used_out, int32_t *max_depth_out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -ENOENT;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    char *text = (char *)malloc((size_t)sz + 1);
    if (!text) { fclose(f); return -ENOMEM; }
    if (fread(text, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(text); return -EIO; }
    fclose(f); text[sz] = 0;

    /* pass 1: raw v / vn pools (file order) */
    size_t nv = 0, nvn = 0, nf = 0;
    for (char *p = text; *p; ) {
        if (p[0] == 'v' && p[1] == ' ') nv++;
        else if (p[0] == 'v' && p[1] == 'n' && p[2] == ' ') nvn++;
        else if (p[0] == 'f' && p[1] == ' ') nf++;
        while (*p && *p != '\n') p++;
        if (*p) p++;
    }
    f32 *pos = (f32 *)malloc((nv + 1) * 3 * sizeof(f32));
    f32 *nrm = (f32 *)malloc((nvn + 1) * 3 * sizeof(f32));
    imap_t vmap, nmap;
    if (!pos || !nrm || imap_init(&vmap, nv + 1) || imap_init(&nmap, nvn + 1)) { free(text); return -ENOMEM; }

    memset(model_uniform, 0, MU_SIZE);
    f32 *points = (f32 *)(model_uniform + MU_POINTS);
    f32 *normals = (f32 *)(model_uniform + MU_NORMALS);
    bho_tri *tris = (bho_tri *)(model_uniform + MU_TRIANGLES);
    int32_t point_count = 0, normal_count = 0, triangle_count = 0;
    int64_t rc = 0;
    size_t iv = 0, ivn = 0;
    /* per-object (tobj "model") state: first-use re-indexing restarts, offsets per model.rs:22-23 */
    int32_t mesh_offset = 0, normal_offset = 0;
    int32_t obj_points = 0, obj_normals = 0;
    int obj_has_faces = 0;

    for (char *p = text; *p; ) {
        char *line = p;
        while (*p && *p != '\n') p++;
        char *next = *p ? p + 1 : p;
        if (line[0] == 'v' && line[1] == ' ') {
            char *q = line + 2;
            pos[3 * iv] = strtof(q, &q); pos[3 * iv + 1] = strtof(q, &q); pos[3 * iv + 2] = strtof(q, &q); iv++;
        } else if (line[0] == 'v' && line[1] == 'n' && line[2] == ' ') {
            char *q = line + 3;
            nrm[3 * ivn] = strtof(q, &q); nrm[3 * ivn + 1] = strtof(q, &q); nrm[3 * ivn + 2] = strtof(q, &q); ivn++;
        } else if ((line[0] == 'o' || line[0] == 'g') && (line[1] == ' ' || line[1] == '\n' || line[1] == '\r' || line[1] == 0)) {
            if (obj_has_faces) {
                /* tobj closes the current model; bhusie (model.rs:22-23, Q17) offsets the next one by
                 * triangle_count / normal_count */
                mesh_offset = triangle_count; normal_offset = normal_count;
                obj_points = 0; obj_normals = 0; obj_has_faces = 0;
                for (size_t i = 0; i < vmap.cap; i++) vmap.keys[i] = -1;
                for (size_t i = 0; i < nmap.cap; i++) nmap.keys[i] = -1;
            }
        } else if (line[0] == 'f' && line[1] == ' ') {
            char *q = line + 2;
            int32_t vi[3], ni[3]; int has_n = 0, cnt = 0;
            while (cnt < 3) {
                while (*q == ' ') q++;
                if (*q == '\n' || *q == '\r' || *q == 0) break;
                long a = strtol(q, &q, 10), c = 0; int hn = 0;
                if (*q == '/') { q++; if (*q != '/') strtol(q, &q, 10); if (*q == '/') { q++; c = strtol(q, &q, 10); hn = 1; } }
                vi[cnt] = (int32_t)(a < 0 ? (long)iv + a : a - 1);
                ni[cnt] = hn ? (int32_t)(c < 0 ? (long)ivn + c : c - 1) : -1;
                has_n |= hn; cnt++;
            }
            if (cnt != 3) { rc = -EINVAL; break; }
            if (triangle_count >= MAX_MODEL_VERTICES) { rc = -E2BIG; break; }
            obj_has_faces = 1;
            int32_t pi[3], nn[3];
            for (int k = 0; k < 3; k++) {
                int added;
                if (vi[k] < 0 || (size_t)vi[k] >= iv) { rc = -EINVAL; break; }
                int32_t li = imap_get_or_add(&vmap, vi[k], obj_points, &added);
                if (added) {
                    if (point_count >= MAX_MODEL_VERTICES) { rc = -E2BIG; break; }
                    /* model.rs:36-42: scale (0.5,-0.5,0.5), w = 0 */
                    points[4 * (size_t)point_count + 0] = pos[3 * (size_t)vi[k] + 0] * 0.5f;
                    points[4 * (size_t)point_count + 1] = pos[3 * (size_t)vi[k] + 1] * -0.5f;
                    points[4 * (size_t)point_count + 2] = pos[3 * (size_t)vi[k] + 2] * 0.5f;
                    point_count++; obj_points++;
                }
                pi[k] = li;
                if (has_n) {
                    if (ni[k] < 0 || (size_t)ni[k] >= ivn) { rc = -EINVAL; break; }
                    int32_t ln = imap_get_or_add(&nmap, ni[k], obj_normals, &added);
                    if (added) {
                        if (normal_count >= MAX_MODEL_VERTICES) { rc = -E2BIG; break; }
                        normals[4 * (size_t)normal_count + 0] = nrm[3 * (size_t)ni[k] + 0];
                        normals[4 * (size_t)normal_count + 1] = nrm[3 * (size_t)ni[k] + 1];
                        normals[4 * (size_t)normal_count + 2] = nrm[3 * (size_t)ni[k] + 2];
                        normal_count++; obj_normals++;
                    }
                    nn[k] = ln;
                }
            }
            if (rc) break;
            if (!has_n) {
                /* model.rs:56-68: face normal from the already-scaled points; index = running normal_count.
                 * NOTE: the lookup uses the un-offset index p (model.points[p1]) exactly like the reference. */
                const f32 *a = points + 4 * (size_t)pi[0], *b2 = points + 4 * (size_t)pi[1], *c2 = points + 4 * (size_t)pi[2];
                const f32 e1[3] = { b2[0] - a[0], b2[1] - a[1], b2[2] - a[2] }, e2[3] = { c2[0] - a[0], c2[1] - a[1], c2[2] - a[2] };
                /* cgmath 0.18 Vector3::cross: (y*oz - z*oy, z*ox - x*oz, x*oy - y*ox), plain f32 ops (Rust never contracts) */
                const f32 cr[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
                /* cgmath 0.18 InnerSpace::normalize == self * (1 / magnitude) (recalled from the crate source,
                 * which is not vendored; only reached for OBJ files without `vn`, which lucy.obj is not) */
                const f32 inv = 1.0f / (f32)sqrt((double)(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]));
                const struct { f32 x, y, z; } dir = { cr[0] * inv, cr[1] * inv, cr[2] * inv };
                if (normal_count >= MAX_MODEL_VERTICES) { rc = -E2BIG; break; }
                int32_t index = normal_count;
                normals[4 * (size_t)normal_count + 0] = dir.x; normals[4 * (size_t)normal_count + 1] = dir.y;
                normals[4 * (size_t)normal_count + 2] = dir.z; normal_count++;
                nn[0] = nn[1] = nn[2] = index;                   /* model.rs:67; normal_offset is added below (Q17) */
            }
            bho_tri t = { pi[0] + mesh_offset, pi[1] + mesh_offset, pi[2] + mesh_offset,
                          nn[0] + normal_offset, nn[1] + normal_offset, nn[2] + normal_offset };
            tris[triangle_count++] = t;
        }
        p = next;
    }
    free(text); free(pos); free(nrm); imap_free(&vmap); imap_free(&nmap);
    if (rc) return rc;

    /* header: Model::new defaults (triangle.rs:100,108) via ModelUniform::update (triangle.rs:309-324) */
    f32 hdr_pos[3] = { -10.0f, 0.0f, 30.0f };
    memcpy(model_uniform + MU_POSITION, hdr_pos, 12);
    int32_t one = 1; memcpy(model_uniform + MU_VISIBLE, &one, 4);
    memcpy(model_uniform + 32, &point_count, 4);            /* point_count */
    /* normal_count @36 is never copied by the reference (Q17) -> stays 0 */
    memcpy(model_uniform + 40, &triangle_count, 4);
    int32_t depth = 0;
    int64_t used = bho_build_bvh(model_uniform, triangle_count, &depth);
    if (used < 0) return used;
    if (point_count_out) *point_count_out = point_count;
    if (nodes_used_out) *nodes_used_out = used;
    if (max_depth_out) *max_depth_out = depth;
    return triangle_count;
}

#ifndef BHO_SHADOW                /* the post chain, the disk generator and the leaf-function KAT entry points exist in the f32 flavours only */
#include "bh_oracle_post.inc"   /* post chain: bloom, mix, ACES, FXAA */

/* ------------------------------------------------------------------ leaf-function entry points for KATs */
void bho_kat_euler(const uint8_t *black_hole132, const float *pos_dir6, float step, float *out6)
{
    ctx_t cx; memset(&cx, 0, sizeof cx); cx.bh = parse_black_hole(black_hole132);
    ray_t r = { V(pos_dir6[0], pos_dir6[1], pos_dir6[2]), V(pos_dir6[3], pos_dir6[4], pos_dir6[5]) };
    r = next_ray_euler(&cx, r, step);
    out6[0] = r.position.x; out6[1] = r.position.y; out6[2] = r.position.z;
    out6[3] = r.direction.x; out6[4] = r.direction.y; out6[5] = r.direction.z;
}
void bho_kat_rk(const uint8_t *black_hole132, const float *pos_dir6, float h, float *out8)
{
    ctx_t cx; memset(&cx, 0, sizeof cx); cx.bh = parse_black_hole(black_hole132);
    tls_t tls; memset(&tls, 0, sizeof tls);
    rk_state_t st = { h, 0.0f, { V(pos_dir6[0], pos_dir6[1], pos_dir6[2]), V(pos_dir6[3], pos_dir6[4], pos_dir6[5]) } };
    st = next_ray_rk(&cx, st, &tls);
    out8[0] = st.ray.position.x; out8[1] = st.ray.position.y; out8[2] = st.ray.position.z;
    out8[3] = st.ray.direction.x; out8[4] = st.ray.direction.y; out8[5] = st.ray.direction.z;
    out8[6] = st.h; out8[7] = st.e_max;
}
/* out: hit, t */
void bho_kat_hit_sphere(const float *pos_dir6, float radius, const float *center3, float t_min, float t_max, float *out2)
{
    ray_t r = { V(pos_dir6[0], pos_dir6[1], pos_dir6[2]), V(pos_dir6[3], pos_dir6[4], pos_dir6[5]) };
    render_state rs = hit_sphere(r, radius, V(center3[0], center3[1], center3[2]), V(0, 0, 0), t_min, t_max);
    out2[0] = (float)rs.hit; out2[1] = rs.t;
}
void bho_kat_hit_torus2d(const float *pos_dir6, float inner, float outer, const float *center3, const float *normal3,
                         float t_min, float t_max, float *out2)
{
    ray_t r = { V(pos_dir6[0], pos_dir6[1], pos_dir6[2]), V(pos_dir6[3], pos_dir6[4], pos_dir6[5]) };
    render_state rs = hit_torus2d(r, inner, outer, V(center3[0], center3[1], center3[2]), V(normal3[0], normal3[1], normal3[2]), t_min, t_max);
    out2[0] = (float)rs.hit; out2[1] = rs.t;
}
/* out: hit, t, color rgb */
void bho_kat_hit_triangle(const float *pos_dir6, const float *tri18, float t_min, float t_max, float *out5)
{
    ray_t r = { V(pos_dir6[0], pos_dir6[1], pos_dir6[2]), V(pos_dir6[3], pos_dir6[4], pos_dir6[5]) };
    const float *t = tri18;
    render_state rs = hit_triangle(r, t_min, t_max, V(t[0], t[1], t[2]), V(t[3], t[4], t[5]), V(t[6], t[7], t[8]),
                                   V(t[9], t[10], t[11]), V(t[12], t[13], t[14]), V(t[15], t[16], t[17]));
    out5[0] = (float)rs.hit; out5[1] = rs.t; out5[2] = rs.color.x; out5[3] = rs.color.y; out5[4] = rs.color.z;
}
void bho_kat_sample(const uint8_t *rgba, int32_t w, int32_t h, float u, float v, float *out4)
{
    bho_texture t = { rgba, w, h }; tls_t tls; memset(&tls, 0, sizeof tls);
    v4 s = sample_bilinear(&t, u, v, &tls);
    out4[0] = s.r; out4[1] = s.g; out4[2] = s.b; out4[3] = s.a;
}
void bho_kat_sky_uv(const float *dir3, float *uv2) { sky_uv(V(dir3[0], dir3[1], dir3[2]), &uv2[0], &uv2[1]); }
void bho_kat_create_ray(const uint8_t *camera32, int32_t px, int32_t py, int32_t w, int32_t h, float *out6)
{
    ctx_t cx; memset(&cx, 0, sizeof cx); cx.camera = parse_camera(camera32);
    ray_t r = create_ray(&cx, px, py, w, h);
    out6[0] = r.position.x; out6[1] = r.position.y; out6[2] = r.position.z;
    out6[3] = r.direction.x; out6[4] = r.direction.y; out6[5] = r.direction.z;
}
/* brute-force closest triangle (no BVH), same acceptance rule as trace_ray_model's leaf loop in
 * bvh_lookup order — the cross-check of SURVEY §8c */
void bho_kat_closest_triangle(const uint8_t *model_uniform, int32_t triangle_count, const float *pos_dir6,
                              int use_bvh, int32_t *tri_out, float *t_out)
{
    bho_scene sc; memset(&sc, 0, sizeof sc); sc.models = model_uniform; sc.model_capacity = 1;
    ctx_t cx; memset(&cx, 0, sizeof cx); cx.scene = &sc;
    tls_t tls; memset(&tls, 0, sizeof tls);
    ray_t r = { V(pos_dir6[0], pos_dir6[1], pos_dir6[2]), V(pos_dir6[3], pos_dir6[4], pos_dir6[5]) };
    const float t_min = 1e-8f, t_max = 1e5f;
    if (use_bvh) {
        render_state rs = trace_ray_model(&cx, r, 0, t_min, t_max, &tls);
        *tri_out = rs.hit ? rs.tri : -1; *t_out = rs.t;
        return;
    }
    const bho_tri *tris = (const bho_tri *)(model_uniform + MU_TRIANGLES);
    const float *points = (const float *)(model_uniform + MU_POINTS);
    const float *normals = (const float *)(model_uniform + MU_NORMALS);
    const int32_t *lookup = (const int32_t *)(model_uniform + MU_LOOKUP);
    v3 mpos = rd_v3(model_uniform + MU_POSITION);
    render_state best = rs_zero(t_max);
    /* lookup-position of the current best, to resolve exact-t ties the way a traversal cannot:
     * ties are reported by the caller comparing t only */
    for (int32_t k = 0; k < triangle_count; k++) {
        int32_t index = lookup[k];
        bho_tri ti = tris[index];
        v3 p1 = vadd(V(points[4 * ti.p1], points[4 * ti.p1 + 1], points[4 * ti.p1 + 2]), mpos);
        v3 p2 = vadd(V(points[4 * ti.p2], points[4 * ti.p2 + 1], points[4 * ti.p2 + 2]), mpos);
        v3 p3 = vadd(V(points[4 * ti.p3], points[4 * ti.p3 + 1], points[4 * ti.p3 + 2]), mpos);
        v3 n1 = V(normals[4 * ti.n1], normals[4 * ti.n1 + 1], normals[4 * ti.n1 + 2]);
        v3 n2 = V(normals[4 * ti.n2], normals[4 * ti.n2 + 1], normals[4 * ti.n2 + 2]);
        v3 n3 = V(normals[4 * ti.n3], normals[4 * ti.n3 + 1], normals[4 * ti.n3 + 2]);
        render_state rs = hit_triangle(r, t_min, t_max, p1, p2, p3, n1, n2, n3);
        if (rs.hit && rs.t < best.t) { best = rs; best.tri = index; }
    }
    *tri_out = best.hit ? best.tri : -1; *t_out = best.t;
}
/* scalar math entry points (flavour under test) */
float bho_kat_pow(float x, float y) { return bho_pow(x, y); }
float bho_kat_pow5(float x) { return bho_pow5(x); }
float bho_kat_pow4(float x) { return bho_pow4(x); }
float bho_kat_sin(float x) { return bho_sin(x); }
float bho_kat_cos(float x) { return bho_cos(x); }
float bho_kat_tan(float x) { return bho_tan(x); }
float bho_kat_atan2(float y, float x) { return bho_atan2(y, x); }
float bho_kat_acos(float x) { return bho_acos(x); }
uint16_t bho_kat_f16(float x) { return f32_to_f16(x); }
void bho_kat_math_array(int fn, const float *a, const float *b, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        switch (fn) {
        case 0: out[i] = bho_pow(a[i], b[i]); break;
        case 1: out[i] = bho_pow5(a[i]); break;
        case 2: out[i] = bho_pow4(a[i]); break;
        case 3: out[i] = bho_sin(a[i]); break;
        case 4: out[i] = bho_cos(a[i]); break;
        case 5: out[i] = bho_tan(a[i]); break;
        case 6: out[i] = bho_atan2(a[i], b[i]); break;
        case 7: out[i] = bho_acos(a[i]); break;
        default: out[i] = NAN;
        }
    }
}
#endif /* !BHO_SHADOW */
int bho_flavour(void)
{
#if defined(BHO_SHADOW)
    return 3;
#endif
#if defined(BHO_FLAVOUR_CONTRACT) && defined(BHO_FUSED)
    return 2;
#elif defined(BHO_FLAVOUR_CONTRACT)
    return 1;
#else
    return 0;
#endif
}
int bho_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
