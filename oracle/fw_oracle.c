/*
 * fw_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See fw_oracle.h.
 *
 * Restates, in scalar f32 C (build with -O2 -ffp-contract=off, no fast-math), the
 * reference's AoS, sequential, order-preserving algorithm. Every function cites the
 * reference file:line it follows ("ref" = /root/reference/).
 *
 * The update loop is kept in its "reference-faithful" form on purpose, because this file is
 * also the CPU baseline that bench.py times: AoS records, one heap-allocated
 * last_emitted_age vector per particle that is cloned per particle per frame
 * (ref src/core.rs:592,320), a fresh output vector per stream per frame (:589-659), one task
 * per spawner on a thread pool (:583-585), sequential spawn (:377).
 */
#define _GNU_SOURCE
#include "fw_oracle.h"
/* sine / cosine: the library's own IEEE-only definition, compiled into both sides (see the header) */
#include "../include/fw_sincos.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small vector math
 * glam 0.32.1 scalar-math formulas (crate not on disk; restated, PARITY UNPINNED). */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } q4;

static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_div(v3 a, float s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline float v3_length(v3 a) { return sqrtf(v3_dot(a, a)); }
static inline int v3_is_zero(v3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
/* glam Vec3::normalize: self * self.length_recip() */
static inline v3 v3_normalize(v3 a) { return v3_mul(a, 1.0f / v3_length(a)); }
/* glam Vec3::normalize_or_zero: rcp finite and > 0, else ZERO */
static inline v3 v3_normalize_or_zero(v3 a) {
    float rcp = 1.0f / v3_length(a);
    if (isfinite(rcp) && rcp > 0.0f) return v3_mul(a, rcp);
    return v3_make(0.0f, 0.0f, 0.0f);
}
/* glam Vec3::project_onto: rhs * self.dot(rhs) * rhs.dot(rhs).recip() */
static inline v3 v3_project_onto(v3 a, v3 rhs) {
    float other_len_sq_rcp = 1.0f / v3_dot(rhs, rhs);
    return v3_mul(v3_mul(rhs, v3_dot(a, rhs)), other_len_sq_rcp);
}
static inline v3 v3_reject_from(v3 a, v3 rhs) { return v3_sub(a, v3_project_onto(a, rhs)); }

/* glam Quat::mul_quat, scalar form */
static inline q4 q_mul(q4 a, q4 b) {
    q4 r;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return r;
}
/* glam Quat::mul_vec3, scalar form */
static inline v3 q_mul_v3(q4 q, v3 v) {
    float w = q.w;
    v3 b = v3_make(q.x, q.y, q.z);
    float b2 = v3_dot(b, b);
    v3 r = v3_mul(v, w * w - b2);
    r = v3_add(r, v3_mul(b, v3_dot(v, b) * 2.0f));
    r = v3_add(r, v3_mul(v3_cross(b, v), w * 2.0f));
    return r;
}
static inline q4 q_conj(q4 q) { q4 r = {-q.x, -q.y, -q.z, q.w}; return r; }
/* glam Quat::from_axis_angle */
static inline q4 q_from_axis_angle(v3 axis, float angle) {
    float s, c;
    fw_sincosf(angle * 0.5f, &s, &c);
    q4 r = {axis.x * s, axis.y * s, axis.z * s, c};
    return r;
}
/* glam Quat::from_scaled_axis: zero vector -> identity (used at ref src/core.rs:646) */
static inline q4 q_from_scaled_axis(v3 v) {
    float length = v3_length(v);
    if (length == 0.0f) { q4 id = {0.0f, 0.0f, 0.0f, 1.0f}; return id; }
    return q_from_axis_angle(v3_div(v, length), length);
}
static inline q4 q_from_rotation_y(float angle) {
    float s, c;
    fw_sincosf(angle * 0.5f, &s, &c);
    q4 r = {0.0f, s, 0.0f, c};
    return r;
}
/* glam Vec3::any_orthonormal_vector (for the 180-degree branch of from_rotation_arc) */
static inline v3 v3_any_orthonormal(v3 n) {
    float sign = copysignf(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    return v3_make(b, sign + n.y * n.y * a, -n.y);
}
/* glam Quat::from_rotation_arc */
static inline q4 q_from_rotation_arc(v3 from, v3 to) {
    const float ONE_MINUS_EPS = 1.0f - 2.0f * FLT_EPSILON;
    float dot = v3_dot(from, to);
    if (dot > ONE_MINUS_EPS) { q4 id = {0.0f, 0.0f, 0.0f, 1.0f}; return id; }
    if (dot < -ONE_MINUS_EPS) return q_from_axis_angle(v3_any_orthonormal(from), 3.14159265358979323846f);
    v3 c = v3_cross(from, to);
    q4 q = {c.x, c.y, c.z, 1.0f + dot};
    float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    float rcp = 1.0f / len;
    q.x *= rcp; q.y *= rcp; q.z *= rcp; q.w *= rcp;
    return q;
}

void fwo_sincosf(float x, float *s, float *c) { fw_sincosf(x, s, c); }
void fwo_sincosf_array(const float *x, uint64_t n, float *s, float *c) {
    for (uint64_t i = 0; i < n; i++) fw_sincosf(x[i], &s[i], &c[i]);
}
void fwo_quat_from_scaled_axis(const float v[3], float out[4]) {
    q4 q = q_from_scaled_axis(v3_make(v[0], v[1], v[2]));
    out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
}
void fwo_quat_mul(const float a[4], const float b[4], float out[4]) {
    q4 qa = {a[0], a[1], a[2], a[3]}, qb = {b[0], b[1], b[2], b[3]};
    q4 q = q_mul(qa, qb);
    out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
}

/* the glam / bevy_utilitarian restatements one by one, for tests/test_reference_vectors.py (known answers of the
 * crates themselves, rust/gen_golden.rs) and the host-compiled kernel harness (scripts/probes/host_math.cu) */
void fwo_quat_from_rotation_arc(const float from[3], const float to[3], float out[4]) {
    q4 q = q_from_rotation_arc(v3_make(from[0], from[1], from[2]), v3_make(to[0], to[1], to[2]));
    out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
}
void fwo_quat_mul_vec3(const float q[4], const float v[3], float out[3]) {
    q4 qq = {q[0], q[1], q[2], q[3]};
    v3 r = q_mul_v3(qq, v3_make(v[0], v[1], v[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void fwo_vec3_normalize_or_zero(const float v[3], float out[3]) {
    v3 r = v3_normalize_or_zero(v3_make(v[0], v[1], v[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void fwo_vec3_project_onto(const float a[3], const float b[3], float out[3]) {
    v3 r = v3_project_onto(v3_make(a[0], a[1], a[2]), v3_make(b[0], b[1], b[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void fwo_vec3_reject_from(const float a[3], const float b[3], float out[3]) {
    v3 r = v3_reject_from(v3_make(a[0], a[1], a[2]), v3_make(b[0], b[1], b[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
/* PitchYaw(u, v).to_unit_vec() as fwo_generate_point uses it: (sin v cos u, cos v, sin v sin u) */
void fwo_pitch_yaw_to_unit_vec(float u, float v, float out[3]) {
    float su, cu, sv, cv;
    fw_sincosf(u, &su, &cu);
    fw_sincosf(v, &sv, &cv);
    out[0] = sv * cu; out[1] = cv; out[2] = sv * su;
}

/* ------------------------------------------------------------------ Rust f32 helpers */
/* core::f32::rem_euclid */
float fwo_rem_euclid(float a, float b) {
    float r = fmodf(a, b);
    return (r < 0.0f) ? r + fabsf(b) : r;
}
/* core::f32::div_euclid */
float fwo_div_euclid(float a, float b) {
    float q = truncf(a / b);
    if (fmodf(a, b) < 0.0f) return (b > 0.0f) ? q - 1.0f : q + 1.0f;
    return q;
}
/* `f32 as usize`: saturating, NaN -> 0 */
static inline uint64_t f32_as_usize(float f) {
    if (!(f > 0.0f)) return 0; /* NaN, negatives, zero */
    if (f >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)f;
}

/* ------------------------------------------------------------------ emission pacing
 * ref src/core.rs:553-575 compute_emission_count, line by line. */
void fwo_compute_emission_count(float time_passed_in_cycle, float last_emission,
                                float cycle_duration, float emission_offset_start,
                                float emission_offset_end, float particles_per_cycle,
                                uint64_t *n_emit, float *next_last) {
    float percent_passed = time_passed_in_cycle / cycle_duration;
    float last_emission_percent = last_emission / cycle_duration;
    float percent_passed_since_emission = fminf(percent_passed, emission_offset_end) -
                                          fmaxf(last_emission_percent, emission_offset_start);
    float percent_between_emissions =
        (emission_offset_end - emission_offset_start) / particles_per_cycle;
    float times_needed_to_emit =
        fwo_div_euclid(percent_passed_since_emission, percent_between_emissions);
    uint64_t times_needed_to_emit_usize = f32_as_usize(times_needed_to_emit);
    float next_last_emission_percent = fmaxf(last_emission_percent, emission_offset_start) +
                                       times_needed_to_emit * percent_between_emissions;
    float next_last_emission = next_last_emission_percent * cycle_duration;
    *n_emit = times_needed_to_emit_usize;
    *next_last = next_last_emission;
}

/* ------------------------------------------------------------------ curves
 * ref src/curve.rs:8-75 (FireworkCurve<f32>) over bevy_math 0.19.0 EvenCore / UnevenCore
 * (crate not on disk; even_interp / uneven_interp restated, PARITY UNPINNED beyond the three
 * knots of ref src/curve.rs:245-258). Returns 0 = exact/tail at *lo, 1 = between lo,hi with s. */
static int even_interp(uint32_t n, float t, uint32_t *lo, uint32_t *hi, float *s) {
    uint32_t subdivs = n - 1u;
    float step = (1.0f - 0.0f) / (float)subdivs; /* domain.length() / subdivs */
    float t_shifted = t - 0.0f;
    float steps_taken = t_shifted / step;
    if (!(steps_taken > 0.0f)) { *lo = 0; return 0; }               /* LeftTail  (NaN too) */
    if (steps_taken >= (float)subdivs) { *lo = n - 1u; return 0; }  /* RightTail */
    float fl = floorf(steps_taken);
    *lo = (uint32_t)fl;
    *hi = *lo + 1u;
    *s = steps_taken - fl; /* f32::fract for positive values */
    if (*s == 0.0f) return 0; /* Exact(lower) */
    return 1;
}
static int uneven_interp(const float *times, uint32_t n, float t, uint32_t *lo, uint32_t *hi,
                         float *s) {
    /* binary_search_by(partial_cmp): first index whose time is >= t */
    uint32_t idx = 0;
    while (idx < n && times[idx] < t) idx++;
    if (idx < n && times[idx] == t) { *lo = idx; return 0; } /* Exact */
    if (idx == 0) { *lo = 0; return 0; }                      /* LeftTail  */
    if (idx >= n) { *lo = n - 1u; return 0; }                 /* RightTail */
    float t_lower = times[idx - 1u], t_upper = times[idx];
    *lo = idx - 1u;
    *hi = idx;
    *s = (t - t_lower) / (t_upper - t_lower);
    return 1;
}
static inline float clamp01(float t) { /* Interval::clamp = f32::clamp(0,1); NaN -> 0 here */
    if (!(t > 0.0f)) return 0.0f;
    if (t > 1.0f) return 1.0f;
    return t;
}
/* Curve::sample_clamped of FireworkCurve<f32> (called at ref src/core.rs:603).
 * f32 interpolation: a + (b - a) * s. */
float fwo_sample_curve(const fw_curve_f32 *c, float t) {
    t = clamp01(t);
    uint32_t lo = 0, hi = 0;
    float s = 0.0f;
    int between;
    switch (c->kind) {
    case FW_CURVE_EVEN: between = even_interp(c->n, t, &lo, &hi, &s); break;
    case FW_CURVE_UNEVEN: between = uneven_interp(c->times, c->n, t, &lo, &hi, &s); break;
    default: return c->values[0];
    }
    if (!between) return c->values[lo];
    float a = c->values[lo], b = c->values[hi];
    return a + (b - a) * s;
}
/* Curve::sample_clamped of FireworkGradient<LinearRgba> (ref src/curve.rs:171-239, called at
 * src/core.rs:460-461,653,655); interpolation = bevy_color Mix: a*(1-s) + b*s per channel. */
void fwo_sample_gradient(const fw_gradient *g, float t, float out[4]) {
    t = clamp01(t);
    uint32_t lo = 0, hi = 0;
    float s = 0.0f;
    int between = 0;
    switch (g->kind) {
    case FW_CURVE_EVEN: between = even_interp(g->n, t, &lo, &hi, &s); break;
    case FW_CURVE_UNEVEN: between = uneven_interp(g->times, g->n, t, &lo, &hi, &s); break;
    default: lo = 0; break;
    }
    if (!between) { memcpy(out, g->colors[lo], 4 * sizeof(float)); return; }
    float n_factor = 1.0f - s;
    for (int k = 0; k < 4; k++) out[k] = g->colors[lo][k] * n_factor + g->colors[hi][k] * s;
}

/* ------------------------------------------------------------------ RNG protocol
 * The reference draws from rand 0.9.4's unseedable thread-local generator
 * (ref src/emission_shape.rs:23-25,33), so replay is DEFINED by this build:
 * Philox4x32-10 (Salmon et al., SC'11), key = seed, counter =
 * (serial_lo, serial_hi, spawner_key, emitter<<8 | block); draw d of a particle is lane d&3
 * of block d>>2, converted like rand's StandardUniform f32: (x >> 8) * 2^-24 in [0,1). */
void fwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
float fwo_uniform(uint64_t seed, uint32_t spawner_key, uint32_t emitter, uint64_t serial,
                  uint32_t draw) {
    uint32_t ctr[4] = {(uint32_t)serial, (uint32_t)(serial >> 32), spawner_key,
                       (emitter << 8) | (draw >> 2)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    fwo_philox4x32_10(ctr, key, out);
    return (float)(out[draw & 3u] >> 8) * (1.0f / 16777216.0f);
}
/* draw indices, in the reference's draw order (ref src/core.rs:438-466) */
enum {
    DRAW_SHAPE0 = 0, DRAW_SHAPE1 = 1, DRAW_SHAPE2 = 2,
    DRAW_VEL_ANGLE = 3, DRAW_VEL_RADIUS = 4, DRAW_VEL_MAG = 5,
    DRAW_RADIAL = 6, DRAW_SCALE = 7, DRAW_LIFETIME = 8,
    DRAW_ANG_ANGLE = 9, DRAW_ANG_RADIUS = 10, DRAW_ANG_MAG = 11
};

/* ------------------------------------------------------------------ samplers
 * bevy_utilitarian 0.10.0 is not on disk: these formulas are THE BUILD'S DEFINITION
 * (PARITY UNPINNED). */
static inline float rand_f32(const fw_rand_f32 *r, float u) { return u * (r->max - r->min) + r->min; }
#define FW_PI 3.14159265358979323846f
/* RandVec3::generate: direction tilted by a polar angle u_radius*spread at azimuth
 * u_angle*2pi (cone around `direction`), times a uniform magnitude. spread <= 0: direction
 * is used as given. */
void fwo_rand_vec3(const fw_rand_vec3 *r, float u_angle, float u_radius, float u_mag, float out[3]) {
    v3 dir = v3_make(r->direction[0], r->direction[1], r->direction[2]);
    if (r->spread > 0.0f) {
        float a = u_angle * 2.0f * FW_PI;
        float p = u_radius * r->spread;
        float sp, cp, sa, ca;
        fw_sincosf(p, &sp, &cp);
        fw_sincosf(a, &sa, &ca);
        v3 local = v3_make(sp * ca, cp, sp * sa);
        q4 arc = q_from_rotation_arc(v3_make(0.0f, 1.0f, 0.0f), v3_normalize_or_zero(dir));
        dir = q_mul_v3(arc, local);
    }
    float m = rand_f32(&r->magnitude, u_mag);
    v3 o = v3_mul(dir, m);
    out[0] = o.x; out[1] = o.y; out[2] = o.z;
}
/* ref src/emission_shape.rs:18-39 EmissionShape::generate_point. PitchYaw::to_unit_vec
 * (bevy_utilitarian) is DEFINED here as (sin v cos u, cos v, sin v sin u). */
void fwo_generate_point(const fw_emission_settings *e, float u0, float u1, float u2, float out[3]) {
    v3 p = v3_make(0.0f, 0.0f, 0.0f);
    if (e->shape_kind == FW_SHAPE_SPHERE) {
        float u = u0 * 2.0f * FW_PI, v = u1 * FW_PI, r = u2; /* :23-25 */
        float su, cu, sv, cv;
        fw_sincosf(u, &su, &cu);
        fw_sincosf(v, &sv, &cv);
        v3 unit = v3_make(sv * cu, cv, sv * su);
        p = v3_mul(v3_mul(unit, r), e->shape_radius); /* :30 */
    } else if (e->shape_kind == FW_SHAPE_CIRCLE) {
        float u = u0 * 2.0f * FW_PI, r = u1; /* :33 */
        q4 arc = q_from_rotation_arc(v3_make(0.0f, 1.0f, 0.0f),
                                     v3_make(e->shape_normal[0], e->shape_normal[1], e->shape_normal[2]));
        q4 q = q_mul(arc, q_from_rotation_y(u)); /* :34-35, left-to-right */
        p = q_mul_v3(q, v3_make(r * e->shape_radius, 0.0f, 0.0f)); /* :36 */
    }
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}

/* ------------------------------------------------------------------ ray casting
 * avian3d 0.7.0 SpatialQuery::cast_ray over parry3d 0.27.0 shapes (crates not on disk;
 * closest hit, solid = true; restated from parry's clip_aabb_line / ray_toi_with_ball,
 * PARITY UNPINNED). Ties on distance keep the lowest collider index. */
static int ray_cuboid_local(v3 he, v3 o, v3 d, float max_toi, float *toi, v3 *normal) {
    float tmax = FLT_MAX, tmin = -FLT_MAX;
    int near_side = 0, near_diag = 0;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, hh[3] = {he.x, he.y, he.z};
    for (int i = 0; i < 3; i++) {
        float mn = -hh[i], mx = hh[i];
        if (dd[i] == 0.0f) {
            if (oo[i] < mn || oo[i] > mx) return 0;
        } else {
            float denom = 1.0f / dd[i];
            int flip;
            float t_near = (mn - oo[i]) * denom;
            float t_far = (mx - oo[i]) * denom;
            if (t_near > t_far) { flip = 1; float tmp = t_near; t_near = t_far; t_far = tmp; }
            else flip = 0;
            if (t_near > tmin) { tmin = t_near; near_side = flip ? -(i + 1) : (i + 1); near_diag = 0; }
            else if (t_near == tmin) near_diag = 1;
            if (t_far < tmax) tmax = t_far;
            if (tmax < 0.0f || tmin > tmax) return 0;
        }
    }
    if (tmin < 0.0f) { /* origin inside, solid: toi 0, zero normal */
        *toi = 0.0f;
        *normal = v3_make(0.0f, 0.0f, 0.0f);
        return 1;
    }
    if (tmin <= max_toi) {
        float nn[3] = {0.0f, 0.0f, 0.0f};
        if (near_diag) { /* edge / corner hit: parry's clip_aabb_line returns -dir.normalize() */
            v3 nd = v3_normalize(d);
            nn[0] = -nd.x; nn[1] = -nd.y; nn[2] = -nd.z;
        } else if (near_side != 0) {
            if (near_side < 0) nn[-near_side - 1] = 1.0f; else nn[near_side - 1] = -1.0f;
        }
        *toi = tmin;
        *normal = v3_make(nn[0], nn[1], nn[2]);
        return 1;
    }
    return 0;
}
static int ray_ball_local(float radius, v3 o, v3 d, float max_toi, float *toi, v3 *normal) {
    float a = v3_dot(d, d);
    float b = v3_dot(o, d);
    float c = v3_dot(o, o) - radius * radius;
    float t;
    int inside = 0;
    if (a == 0.0f) {
        if (c > 0.0f) return 0;
        t = 0.0f; inside = 1;
    } else if (c > 0.0f && b > 0.0f) {
        return 0;
    } else {
        float delta = b * b - a * c;
        if (delta < 0.0f) return 0;
        t = (-b - sqrtf(delta)) / a;
        if (t <= 0.0f) { t = 0.0f; inside = 1; }
    }
    if (!(t <= max_toi)) return 0;
    v3 pos = v3_add(o, v3_mul(d, t));
    v3 n = v3_normalize(pos);
    if (inside) n = v3_make(-n.x, -n.y, -n.z);
    *toi = t;
    *normal = n;
    return 1;
}
/* Cylinder and cone (parry: support-map shapes cast with GJK -- not restated; DEFINED BY THIS
 * BUILD as the analytic solid of revolution about +Y whose radius goes linearly from r0 at
 * y = -h to r1 at y = +h: cylinder r0 = r1, cone r1 = 0). Solid: an origin inside gives toi 0
 * and a zero normal, like the cuboid. Candidates in the order bottom cap, top cap, the two side
 * roots; the first smallest one wins. */
static int ray_frustum_local(float r0, float r1, float h, v3 o, v3 d, float max_toi, float *toi, v3 *normal) {
    const float s = (r1 - r0) / (2.0f * h), c0 = (r0 + r1) * 0.5f;
    const float ro = c0 + s * o.y; /* radius of the solid at the origin's height */
    if (o.y >= -h && o.y <= h && ro >= 0.0f && o.x * o.x + o.z * o.z <= ro * ro) {
        *toi = 0.0f;
        *normal = v3_make(0.0f, 0.0f, 0.0f);
        return 1;
    }
    int found = 0;
    float best = 0.0f;
    v3 bn = v3_make(0.0f, 0.0f, 0.0f);
    if (d.y != 0.0f) {
        const float inv = 1.0f / d.y;
        for (int cap = 0; cap < 2; cap++) {
            const float yc = cap ? h : -h, rc = cap ? r1 : r0;
            const float t = (yc - o.y) * inv;
            const float x = o.x + d.x * t, z = o.z + d.z * t;
            if (rc > 0.0f && t >= 0.0f && x * x + z * z <= rc * rc && (!found || t < best)) {
                found = 1;
                best = t;
                bn = v3_make(0.0f, cap ? 1.0f : -1.0f, 0.0f);
            }
        }
    }
    const float A = d.x * d.x + d.z * d.z - (s * s) * (d.y * d.y);
    const float B = o.x * d.x + o.z * d.z - (s * ro) * d.y;
    const float C = o.x * o.x + o.z * o.z - ro * ro;
    float roots[2];
    int n_roots = 0;
    if (A != 0.0f) {
        const float disc = B * B - A * C;
        if (disc >= 0.0f) {
            /* q = -(B + sign(B) sqrt(disc)); roots q / A and C / q. The textbook (-B -+ sqrt) / A cancels
             * to 0 / A when A C is tiny against B B -- a ray almost along a cone's slant -- and reported
             * a hit at distance (-)0 for a cone the ray never comes near. Same order: [0] = the "-" root. */
            const float sq = sqrtf(disc);
            const float q = B < 0.0f ? sq - B : -(B + sq);
            if (q != 0.0f) {
                roots[0] = B < 0.0f ? C / q : q / A;
                roots[1] = B < 0.0f ? q / A : C / q;
            } else { /* B = 0 and disc = 0: the double root 0 */
                roots[0] = 0.0f;
                roots[1] = 0.0f;
            }
            n_roots = 2;
        }
    } else if (B != 0.0f) {
        roots[0] = -C / (2.0f * B);
        n_roots = 1;
    }
    for (int k = 0; k < n_roots; k++) {
        const float t = roots[k];
        const float y = o.y + d.y * t, rr = c0 + s * y;
        if (t >= 0.0f && y >= -h && y <= h && rr >= 0.0f && (!found || t < best)) {
            const float x = o.x + d.x * t, z = o.z + d.z * t;
            v3 g = v3_make(x, -(s * rr), z);
            const float len = v3_length(g);
            found = 1;
            best = t;
            bn = len > 0.0f ? v3_mul(g, 1.0f / len) : v3_make(0.0f, s < 0.0f ? 1.0f : -1.0f, 0.0f);
        }
    }
    if (!found || !(best <= max_toi)) return 0;
    *toi = best;
    *normal = bn;
    return 1;
}
/* Capsule: the segment (0,-h,0)..(0,h,0) swept by a ball of radius r (parry's Capsule; parry casts
 * rays against it through its support map with GJK -- crate not on disk, PARITY UNPINNED -- this is
 * the analytic solid). Solid: an origin inside gives toi 0 and a zero normal. Candidates in the
 * order side, bottom cap, top cap; only the entering roots (the origin is outside); the first
 * smallest one wins. */
static int ray_capsule_local(float r, float h, v3 o, v3 d, float max_toi, float *toi, v3 *normal) {
    const float yc = fminf(fmaxf(o.y, -h), h); /* nearest point of the segment */
    const float dy0 = o.y - yc;
    if (o.x * o.x + dy0 * dy0 + o.z * o.z <= r * r) {
        *toi = 0.0f;
        *normal = v3_make(0.0f, 0.0f, 0.0f);
        return 1;
    }
    int found = 0;
    float best = 0.0f;
    v3 bn = v3_make(0.0f, 0.0f, 0.0f);
    const float inv_r = 1.0f / r;
    {
        const float A = d.x * d.x + d.z * d.z;
        const float B = o.x * d.x + o.z * d.z;
        const float C = o.x * o.x + o.z * o.z - r * r;
        if (A != 0.0f) {
            const float disc = B * B - A * C;
            if (disc >= 0.0f) {
                const float t = (-B - sqrtf(disc)) / A;
                const float y = o.y + d.y * t;
                if (t >= 0.0f && y >= -h && y <= h) {
                    found = 1;
                    best = t;
                    bn = v3_make((o.x + d.x * t) * inv_r, 0.0f, (o.z + d.z * t) * inv_r);
                }
            }
        }
    }
    for (int cap = 0; cap < 2; cap++) {
        const float cy = cap ? h : -h;
        const v3 oc = v3_make(o.x, o.y - cy, o.z);
        const float a = v3_dot(d, d), b = v3_dot(oc, d), c = v3_dot(oc, oc) - r * r;
        if (a != 0.0f) {
            const float disc = b * b - a * c;
            if (disc >= 0.0f) {
                const float t = (-b - sqrtf(disc)) / a;
                const v3 p = v3_add(oc, v3_mul(d, t)); /* relative to the cap's centre */
                const int outer = cap ? p.y >= 0.0f : p.y <= 0.0f;
                if (t >= 0.0f && outer && (!found || t < best)) {
                    found = 1;
                    best = t;
                    bn = v3_mul(p, inv_r);
                }
            }
        }
    }
    if (!found || !(best <= max_toi)) return 0;
    *toi = best;
    *normal = bn;
    return 1;
}
/* TEST HELPER (not part of the restatement): conservative boxes that let the full-size collision
 * parity tests finish in seconds. boxes[6i..6i+5] = min.xyz, max.xyz of collider i's bounding
 * sphere, inflated far beyond any fp32 rounding of the exact tests below. A collider is skipped
 * only when the box of the ray segment [o, o + d*max_distance] misses its box -- the exact test
 * could not have reported a hit within max_distance -- so the result is the brute-force result
 * (tests/test_oracle_golden.py compares the two on random scenes). Non-finite input never culls. */
static void collider_cull_box(const fw_collider *c, float box[6]) {
    float r;
    if (c->kind == FW_COLLIDER_SPHERE) r = fabsf(c->half_extents[0]);
    else if (c->kind == FW_COLLIDER_CUBOID)
        r = sqrtf(c->half_extents[0] * c->half_extents[0] + c->half_extents[1] * c->half_extents[1] + c->half_extents[2] * c->half_extents[2]);
    else if (c->kind == FW_COLLIDER_CAPSULE) r = fabsf(c->half_extents[0]) + fabsf(c->half_extents[1]);
    else r = sqrtf(c->half_extents[0] * c->half_extents[0] + c->half_extents[1] * c->half_extents[1]);
    for (int a = 0; a < 3; a++) {
        const float m = 1e-3f + 1e-3f * (fabsf(c->translation[a]) + r);
        box[a] = c->translation[a] - r - m;
        box[3 + a] = c->translation[a] + r + m;
        if (!isfinite(box[a]) || !isfinite(box[3 + a])) { box[a] = -FLT_MAX; box[3 + a] = FLT_MAX; }
    }
}
static int cast_ray_impl(const fw_collider *c, uint32_t n, const float *boxes, const fw_collision_settings *filter, const float origin[3],
                         const float dir[3], float max_distance, float *distance, float normal[3],
                         uint32_t *index) {
    const uint32_t filter_mask = filter->filter_mask; /* SpatialQueryFilter::mask */
    int found = 0;
    float best = 0.0f;
    v3 best_n = v3_make(0.0f, 0.0f, 0.0f);
    uint32_t best_i = 0;
    v3 o = v3_make(origin[0], origin[1], origin[2]), d = v3_make(dir[0], dir[1], dir[2]);
    float slo[3], shi[3];
    int cull = boxes != NULL;
    for (int a = 0; a < 3 && cull; a++) {
        const float e = origin[a] + dir[a] * max_distance;
        if (!isfinite(e) || !isfinite(origin[a])) cull = 0;
        slo[a] = fminf(origin[a], e);
        shi[a] = fmaxf(origin[a], e);
    }
    for (uint32_t i = 0; i < n; i++) {
        if ((c[i].layers & filter_mask) == 0u) continue;
        { /* SpatialQueryFilter::excluded_entities (ref src/core.rs:247,764) */
            int excluded = 0;
            for (uint32_t x = 0; x < filter->n_excluded && x < FW_MAX_EXCLUDED; x++) excluded |= filter->excluded_keys[x] == c[i].key;
            if (excluded) continue;
        }
        if (cull) {
            const float *b = boxes + 6 * (size_t)i;
            if (shi[0] < b[0] || slo[0] > b[3] || shi[1] < b[1] || slo[1] > b[4] || shi[2] < b[2] || slo[2] > b[5]) continue;
        }
        q4 rot = {c[i].rotation[0], c[i].rotation[1], c[i].rotation[2], c[i].rotation[3]};
        q4 inv = q_conj(rot);
        v3 tr = v3_make(c[i].translation[0], c[i].translation[1], c[i].translation[2]);
        v3 ol = q_mul_v3(inv, v3_sub(o, tr));
        v3 dl = q_mul_v3(inv, d);
        float toi;
        v3 nl;
        int hit;
        if (c[i].kind == FW_COLLIDER_SPHERE)
            hit = ray_ball_local(c[i].half_extents[0], ol, dl, max_distance, &toi, &nl);
        else if (c[i].kind == FW_COLLIDER_CYLINDER)
            hit = ray_frustum_local(c[i].half_extents[0], c[i].half_extents[0], c[i].half_extents[1], ol, dl, max_distance, &toi, &nl);
        else if (c[i].kind == FW_COLLIDER_CONE)
            hit = ray_frustum_local(c[i].half_extents[0], 0.0f, c[i].half_extents[1], ol, dl, max_distance, &toi, &nl);
        else if (c[i].kind == FW_COLLIDER_CAPSULE)
            hit = ray_capsule_local(c[i].half_extents[0], c[i].half_extents[1], ol, dl, max_distance, &toi, &nl);
        else
            hit = ray_cuboid_local(v3_make(c[i].half_extents[0], c[i].half_extents[1], c[i].half_extents[2]),
                                   ol, dl, max_distance, &toi, &nl);
        if (hit && (!found || toi < best)) {
            found = 1;
            best = toi;
            best_n = q_mul_v3(rot, nl);
            best_i = i;
        }
    }
    if (!found) return 0;
    *distance = best;
    normal[0] = best_n.x; normal[1] = best_n.y; normal[2] = best_n.z;
    if (index) *index = best_i;
    return 1;
}
int fwo_cast_ray(const fw_collider *c, uint32_t n, uint32_t filter_mask, const float origin[3],
                 const float dir[3], float max_distance, float *distance, float normal[3],
                 uint32_t *index) {
    fw_collision_settings f;
    memset(&f, 0, sizeof(f));
    f.filter_mask = filter_mask;
    return cast_ray_impl(c, n, NULL, &f, origin, dir, max_distance, distance, normal, index);
}
int fwo_cast_ray_culled(const fw_collider *c, uint32_t n, uint32_t filter_mask, const float origin[3],
                        const float dir[3], float max_distance, float *distance, float normal[3],
                        uint32_t *index) {
    float *boxes = (float *)malloc(sizeof(float) * 6 * (n ? n : 1));
    for (uint32_t i = 0; i < n; i++) collider_cull_box(&c[i], boxes + 6 * (size_t)i);
    fw_collision_settings f;
    memset(&f, 0, sizeof(f));
    f.filter_mask = filter_mask;
    const int r = cast_ray_impl(c, n, boxes, &f, origin, dir, max_distance, distance, normal, index);
    free(boxes);
    return r;
}

/* ref src/core.rs:744-800 particle_collision, line by line (including the
 * time-minus-distance subtraction at :786 and the undiminished delta at :766-775). */
static void particle_collision_impl(const fw_collider *colliders, uint32_t n_colliders, const float *boxes,
                                    const fw_collision_settings *cs, float pos_io[3], float vel_io[3],
                                    float delta, uint32_t *should_destroy_out) {
    v3 pos = v3_make(pos_io[0], pos_io[1], pos_io[2]);
    v3 vel = v3_make(vel_io[0], vel_io[1], vel_io[2]);
    float orig_delta = delta;
    int n_steps = 0;
    uint32_t should_destroy = 0;
    while (delta > 0.0f && n_steps < 4) {
        /* Dir3::try_from(vel), fallback Dir3::Y (:758-761) */
        float len = v3_length(vel);
        v3 dir = (isfinite(len) && len > 0.0f) ? v3_div(vel, len) : v3_make(0.0f, 1.0f, 0.0f);
        float o[3] = {pos.x, pos.y, pos.z}, d[3] = {dir.x, dir.y, dir.z}, nrm[3], distance;
        if (cast_ray_impl(colliders, n_colliders, boxes, cs, o, d, v3_length(vel) * delta,
                          &distance, nrm, NULL)) {
            v3 hit_normal = v3_make(nrm[0], nrm[1], nrm[2]);
            if (distance == 0.0f) {
                v3 normal = hit_normal;
                if (v3_is_zero(normal)) {
                    if (!v3_is_zero(vel)) normal = v3_normalize(vel);
                    else normal = v3_make(0.0f, 1.0f, 0.0f);
                }
                /* pos += vel.length().max(1.) * normal * delta  (:775) */
                pos = v3_add(pos, v3_mul(v3_mul(normal, fmaxf(v3_length(vel), 1.0f)), delta));
            } else {
                pos = v3_add(pos, v3_mul(v3_normalize_or_zero(vel), distance)); /* :777 */
                v3 vel_reject = v3_reject_from(vel, hit_normal);                /* :778 */
                v3 vel_project = v3_project_onto(vel, hit_normal);              /* :779 */
                float friction_dv =
                    fminf(v3_length(vel_project), v3_length(vel_reject)) * cs->friction; /* :780-781 */
                vel = v3_sub(v3_sub(vel_reject, v3_mul(v3_normalize_or_zero(vel_reject), friction_dv)),
                             v3_mul(vel_project, cs->restitution));             /* :782-784 */
                pos = v3_add(pos, v3_mul(hit_normal, 0.0001f));                 /* :785 */
                delta = delta - distance;                                       /* :786 */
                if (delta < 0.0f) delta = 0.0f;
                if (delta > orig_delta) delta = orig_delta;
            }
            should_destroy = cs->destroy_on_collision ? 1u : 0u; /* :788 */
            if (should_destroy) break;                           /* :789-791 */
        } else {
            pos = v3_add(pos, v3_mul(vel, delta)); /* :793 */
            delta = 0.0f;
        }
        n_steps += 1;
    }
    pos_io[0] = pos.x; pos_io[1] = pos.y; pos_io[2] = pos.z;
    vel_io[0] = vel.x; vel_io[1] = vel.y; vel_io[2] = vel.z;
    *should_destroy_out = should_destroy;
}
void fwo_particle_collision(const fw_collider *colliders, uint32_t n_colliders,
                            const fw_collision_settings *cs, float pos_io[3], float vel_io[3],
                            float delta, uint32_t *should_destroy_out) {
    particle_collision_impl(colliders, n_colliders, NULL, cs, pos_io, vel_io, delta, should_destroy_out);
}

/* ------------------------------------------------------------------ world model
 * ref src/core.rs:261-321 EmissionData / ParticleSpawnerData / ParticleData */
typedef struct {
    fw_particle_data d;
    float *last_emitted_age; /* Vec<f32>, one per emitter (:320) */
} particle;

typedef struct {
    particle *p;
    size_t len, cap;
} pvec;

typedef struct {
    float last_emission;
    float time_passed_in_cycle;
    int enabled;
    int emits_on_other_particles;
    uint64_t serial; /* particles this emitter has spawned since reset (RNG protocol) */
} emission_data;

typedef struct {
    uint32_t key;
    uint32_t n_types, n_emitters;
    fw_particle_settings *ps;
    fw_emission_settings *es;
    emission_data *emission;
    pvec *particles; /* Vec<Vec<ParticleData>> */
    pvec *destroyed; /* what the particles_destroyed handler of the last frame received */
    v3 parent_velocity;
    v3 origin_translation;
    q4 origin_rotation;
    float modifier_scale, modifier_speed;
    uint64_t manual_queued_count;
    int initialized, finished_notified;
} spawner;

struct fwo_world {
    uint64_t seed;
    spawner **sp;
    size_t n_sp, cap_sp;
    fw_collider *colliders;
    uint32_t n_colliders;
    float *cull_boxes; /* test helper, see collider_cull_box; NULL = brute force */
    int cull;
};

static void pvec_push(pvec *v, particle p) {
    if (v->len == v->cap) { /* RawVec growth: max(4, 2*cap) */
        size_t nc = v->cap ? v->cap * 2 : 4;
        v->p = (particle *)realloc(v->p, nc * sizeof(particle));
        v->cap = nc;
    }
    v->p[v->len++] = p;
}
static void pvec_drop(pvec *v) {
    for (size_t i = 0; i < v->len; i++) free(v->p[i].last_emitted_age);
    free(v->p);
    v->p = NULL;
    v->len = v->cap = 0;
}

fwo_world *fwo_create(uint64_t seed) {
    fwo_world *w = (fwo_world *)calloc(1, sizeof(*w));
    w->seed = seed;
    return w;
}
static void spawner_free(spawner *s) {
    for (uint32_t i = 0; i < s->n_types; i++) { pvec_drop(&s->particles[i]); pvec_drop(&s->destroyed[i]); }
    free(s->particles); free(s->destroyed); free(s->ps); free(s->es); free(s->emission); free(s);
}
void fwo_destroy(fwo_world *w) {
    if (!w) return;
    for (size_t i = 0; i < w->n_sp; i++) spawner_free(w->sp[i]);
    free(w->sp); free(w->colliders); free(w->cull_boxes); free(w);
}
static spawner *find_spawner(const fwo_world *w, uint32_t key) {
    for (size_t i = 0; i < w->n_sp; i++) if (w->sp[i]->key == key) return w->sp[i];
    return NULL;
}

/* ref src/core.rs:343-365 sync_spawner_data for one changed spawner */
int fwo_spawner_reset(fwo_world *w, uint32_t key, const fw_particle_settings *ps, uint32_t n_types,
                      const fw_emission_settings *es, uint32_t n_emitters, uint32_t starts_enabled) {
    spawner *s = find_spawner(w, key);
    if (!s) {
        s = (spawner *)calloc(1, sizeof(*s));
        s->key = key;
        s->origin_rotation.w = 1.0f;
        s->modifier_scale = 1.0f;
        s->modifier_speed = 1.0f;
        if (w->n_sp == w->cap_sp) {
            w->cap_sp = w->cap_sp ? w->cap_sp * 2 : 16;
            w->sp = (spawner **)realloc(w->sp, w->cap_sp * sizeof(spawner *));
        }
        w->sp[w->n_sp++] = s;
    } else {
        for (uint32_t i = 0; i < s->n_types; i++) { pvec_drop(&s->particles[i]); pvec_drop(&s->destroyed[i]); }
        free(s->particles); free(s->destroyed); free(s->ps); free(s->es); free(s->emission);
    }
    s->n_types = n_types;
    s->n_emitters = n_emitters;
    s->ps = (fw_particle_settings *)malloc(sizeof(*ps) * (n_types ? n_types : 1));
    memcpy(s->ps, ps, sizeof(*ps) * n_types);
    s->es = (fw_emission_settings *)malloc(sizeof(*es) * (n_emitters ? n_emitters : 1));
    memcpy(s->es, es, sizeof(*es) * n_emitters);
    s->emission = (emission_data *)calloc(n_emitters ? n_emitters : 1, sizeof(emission_data));
    for (uint32_t i = 0; i < n_emitters; i++) { /* :350-358 */
        s->emission[i].last_emission = 0.0f;
        s->emission[i].time_passed_in_cycle = 0.0f;
        s->emission[i].enabled = starts_enabled ? 1 : 0;
        s->emission[i].emits_on_other_particles = (es[i].mode == FW_MODE_NESTED);
        s->emission[i].serial = 0;
    }
    s->particles = (pvec *)calloc(n_types ? n_types : 1, sizeof(pvec)); /* :360 */
    s->destroyed = (pvec *)calloc(n_types ? n_types : 1, sizeof(pvec));
    s->initialized = 1; /* :361-363 */
    return 0;
}
int fwo_spawner_remove(fwo_world *w, uint32_t key) {
    for (size_t i = 0; i < w->n_sp; i++) {
        if (w->sp[i]->key == key) {
            spawner_free(w->sp[i]);
            memmove(&w->sp[i], &w->sp[i + 1], (w->n_sp - i - 1) * sizeof(spawner *));
            w->n_sp--;
            return 0;
        }
    }
    return 1;
}
void fwo_set_colliders(fwo_world *w, const fw_collider *c, uint32_t n) {
    free(w->colliders);
    w->colliders = (fw_collider *)malloc(sizeof(*c) * (n ? n : 1));
    memcpy(w->colliders, c, sizeof(*c) * n);
    w->n_colliders = n;
    free(w->cull_boxes);
    w->cull_boxes = NULL;
    if (w->cull) {
        w->cull_boxes = (float *)malloc(sizeof(float) * 6 * (n ? n : 1));
        for (uint32_t i = 0; i < n; i++) collider_cull_box(&c[i], w->cull_boxes + 6 * (size_t)i);
    }
}
/* test helper: from the next fwo_set_colliders on, skip colliders whose box the ray segment misses */
void fwo_set_cull(fwo_world *w, int on) { w->cull = on; }

/* ref src/core.rs:288-302 ParticleSpawnerData::active */
static int spawner_active(const spawner *s) {
    int enabled = 0;
    for (uint32_t i = 0; i < s->n_emitters; i++) {
        const emission_data *e = &s->emission[i];
        if (e->emits_on_other_particles) {
            int any = 0;
            for (uint32_t t = 0; t < s->n_types; t++) if (s->particles[t].len != 0) any = 1;
            enabled |= (e->enabled && any);
        } else {
            enabled |= e->enabled;
        }
    }
    return enabled;
}

/* one new particle: ref src/core.rs:437-469 (Global) and :506-544 (Nested) share this body;
 * origin_* are the spawner transform (Global) or the parent particle (Nested). */
static void spawn_one(fwo_world *w, spawner *s, uint32_t emitter, v3 origin_translation,
                      q4 origin_rotation, v3 inherited_velocity) {
    const fw_emission_settings *es = &s->es[emitter];
    const fw_particle_settings *ps = &s->ps[es->particle_index];
    uint64_t serial = s->emission[emitter].serial++;
#define U(draw) fwo_uniform(w->seed, s->key, emitter, serial, (draw))
    float so[3];
    fwo_generate_point(es, U(DRAW_SHAPE0), U(DRAW_SHAPE1), U(DRAW_SHAPE2), so); /* :438 */
    v3 spawn_offset = v3_make(so[0], so[1], so[2]);
    float iv[3];
    fwo_rand_vec3(&es->initial_velocity, U(DRAW_VEL_ANGLE), U(DRAW_VEL_RADIUS), U(DRAW_VEL_MAG), iv);
    float radial = rand_f32(&es->initial_velocity_radial, U(DRAW_RADIAL));
    /* :440-448 */
    v3 velocity = v3_mul(v3_add(q_mul_v3(origin_rotation, v3_make(iv[0], iv[1], iv[2])),
                                v3_mul(v3_normalize_or_zero(spawn_offset), radial)),
                         s->modifier_speed);
    velocity = v3_add(velocity, es->inherit_parent_velocity ? inherited_velocity : v3_make(0.0f, 0.0f, 0.0f));
    float initial_scale = rand_f32(&ps->initial_scale, U(DRAW_SCALE)) * s->modifier_scale; /* :450-451 */
    particle p;
    memset(&p, 0, sizeof(p));
    v3 position = v3_add(origin_translation, spawn_offset); /* :454 */
    p.d.position[0] = position.x; p.d.position[1] = position.y; p.d.position[2] = position.z;
    p.d.lifetime = rand_f32(&ps->lifetime, U(DRAW_LIFETIME)); /* :455 */
    p.d.initial_scale = initial_scale;
    p.d.scale = initial_scale;
    p.d.velocity[0] = velocity.x; p.d.velocity[1] = velocity.y; p.d.velocity[2] = velocity.z;
    p.d.age = 0.0f;
    fwo_sample_gradient(&ps->base_color, 0.0f, p.d.base_color);         /* :460 */
    fwo_sample_gradient(&ps->emissive_color, 0.0f, p.d.emissive_color); /* :461 */
    p.d.pbr = ps->pbr;
    memcpy(p.d.rotation, es->initial_rotation, sizeof(float) * 4); /* :463 */
    float av[3];
    fwo_rand_vec3(&es->initial_angular_velocity, U(DRAW_ANG_ANGLE), U(DRAW_ANG_RADIUS), U(DRAW_ANG_MAG), av);
    memcpy(p.d.angular_velocity, av, sizeof(av)); /* :464-466 */
#undef U
    p.last_emitted_age = (float *)malloc(sizeof(float) * (s->n_emitters ? s->n_emitters : 1)); /* :467 */
    for (uint32_t k = 0; k < s->n_emitters; k++) p.last_emitted_age[k] = -FLT_MAX; /* f32::MIN */
    pvec_push(&s->particles[es->particle_index], p); /* :453 */
}

/* ref src/core.rs:367-551 spawn_particles */
void fwo_spawn_only(fwo_world *w, float dt, const fw_spawner_frame_input *in, uint32_t n_in) {
    for (uint32_t k = 0; k < n_in; k++) {
        spawner *s = find_spawner(w, in[k].spawner_key);
        if (!s) continue;
        s->origin_translation = v3_make(in[k].origin_translation[0], in[k].origin_translation[1], in[k].origin_translation[2]);
        s->origin_rotation.x = in[k].origin_rotation[0]; s->origin_rotation.y = in[k].origin_rotation[1];
        s->origin_rotation.z = in[k].origin_rotation[2]; s->origin_rotation.w = in[k].origin_rotation[3];
        s->parent_velocity = v3_make(in[k].parent_velocity[0], in[k].parent_velocity[1], in[k].parent_velocity[2]);
        s->modifier_scale = in[k].modifier_scale;
        s->modifier_speed = in[k].modifier_speed;
        s->manual_queued_count += in[k].queue_particles; /* :284-286 */
    }
    for (size_t si = 0; si < w->n_sp; si++) { /* :377 */
        spawner *s = w->sp[si];
        if (!spawner_active(s)) continue; /* :378 */
        for (uint32_t i = 0; i < s->n_emitters; i++) { /* :386 */
            const fw_emission_settings *es = &s->es[i];
            emission_data *ed = &s->emission[i];
            if (!ed->enabled) continue; /* :388-390 */
            if (es->mode == FW_MODE_GLOBAL) {
                uint64_t particles_to_spawn = 0;
                if (es->pacing_kind == FW_PACING_ONE_SHOT) { /* :397-400 */
                    ed->enabled = 0;
                    particles_to_spawn = es->one_shot_count;
                } else if (es->pacing_kind == FW_PACING_ON_DEMAND) { /* :401-405 */
                    particles_to_spawn = s->manual_queued_count;
                    s->manual_queued_count = 0;
                } else { /* :406-427 */
                    ed->time_passed_in_cycle = fwo_rem_euclid(ed->time_passed_in_cycle + dt, es->duration);
                    float next_last;
                    fwo_compute_emission_count(ed->time_passed_in_cycle, ed->last_emission, es->duration,
                                               es->offset_start, es->offset_end, es->count,
                                               &particles_to_spawn, &next_last);
                    ed->last_emission = next_last;
                }
                for (uint64_t n = 0; n < particles_to_spawn; n++) /* :437 */
                    spawn_one(w, s, i, s->origin_translation, s->origin_rotation, s->parent_velocity);
            } else { /* Nested :471-546 */
                if (es->pacing_kind != FW_PACING_COUNT_OVER_DURATION) continue; /* :474-485 */
                uint32_t target = es->target_particle_type;
                if (target >= s->n_types) continue;
                size_t n_parents = s->particles[target].len; /* range evaluated once, :488 */
                for (size_t p_i = 0; p_i < n_parents; p_i++) {
                    particle *other = &s->particles[target].p[p_i];
                    uint64_t times;
                    float next_last;
                    fwo_compute_emission_count(other->d.age, other->last_emitted_age[i], other->d.lifetime,
                                               es->offset_start, es->offset_end, es->count, &times, &next_last);
                    other->last_emitted_age[i] = next_last; /* :500 */
                    v3 op = v3_make(other->d.position[0], other->d.position[1], other->d.position[2]);
                    q4 orot = {other->d.rotation[0], other->d.rotation[1], other->d.rotation[2], other->d.rotation[3]};
                    v3 ov = v3_make(other->d.velocity[0], other->d.velocity[1], other->d.velocity[2]);
                    for (uint64_t n = 0; n < times; n++) spawn_one(w, s, i, op, orot, ov); /* :506-544 */
                    /* pvec_push may have moved the target vector if particle_index == target */
                }
            }
        }
    }
}

/* ref src/core.rs:586-668: the body of the par_iter_mut closure for one spawner */
static void update_spawner(const fwo_world *w, spawner *s, float dt) {
    for (uint32_t i = 0; i < s->n_types; i++) { /* :586 */
        const fw_particle_settings *ps = &s->ps[i];
        pvec destroyed = {0, 0, 0}; /* :588 */
        pvec out = {0, 0, 0};       /* collect() into a fresh Vec, :589-659 */
        pvec *src = &s->particles[i];
        for (size_t k = 0; k < src->len; k++) {
            /* let mut particle = particle.clone();  (:592) -- including the heap Vec<f32> */
            particle p = src->p[k];
            size_t lea_bytes = sizeof(float) * (s->n_emitters ? s->n_emitters : 1);
            p.last_emitted_age = (float *)malloc(lea_bytes);
            memcpy(p.last_emitted_age, src->p[k].last_emitted_age, lea_bytes);

            p.d.age += dt; /* :594 */
            if (p.d.age >= p.d.lifetime) { /* :596-599 */
                pvec_push(&destroyed, p);
                continue;
            }
            float age_percent = p.d.age / p.d.lifetime;                       /* :601 */
            float scale_factor = fwo_sample_curve(&ps->scale_curve, age_percent); /* :602-603 */
            p.d.scale = p.d.initial_scale * scale_factor;                      /* :605 */

            v3 pos = v3_make(p.d.position[0], p.d.position[1], p.d.position[2]);
            v3 vel = v3_make(p.d.velocity[0], p.d.velocity[1], p.d.velocity[2]);
            uint32_t should_destroy = 0;
            if (ps->collision.enabled) { /* :608-617 */
                float pp[3] = {pos.x, pos.y, pos.z}, vv[3] = {vel.x, vel.y, vel.z};
                particle_collision_impl(w->colliders, w->n_colliders, w->cull_boxes, &ps->collision, pp, vv, dt, &should_destroy);
                pos = v3_make(pp[0], pp[1], pp[2]);
                vel = v3_make(vv[0], vv[1], vv[2]);
            } else { /* :619-623 */
                pos = v3_add(pos, v3_mul(vel, dt));
            }
            p.d.position[0] = pos.x; p.d.position[1] = pos.y; p.d.position[2] = pos.z; /* :633 */
            p.d.velocity[0] = vel.x; p.d.velocity[1] = vel.y; p.d.velocity[2] = vel.z; /* :634 */
            if (should_destroy) { /* :636-639 */
                pvec_push(&destroyed, p);
                continue;
            }
            /* :641-643 velocity += (acceleration - velocity * linear_drag) * dt */
            v3 acc = v3_make(ps->acceleration[0], ps->acceleration[1], ps->acceleration[2]);
            vel = v3_add(vel, v3_mul(v3_sub(acc, v3_mul(vel, ps->linear_drag)), dt));
            p.d.velocity[0] = vel.x; p.d.velocity[1] = vel.y; p.d.velocity[2] = vel.z;
            /* :645-647 rotation = from_scaled_axis(angular_velocity * dt) * rotation */
            v3 av = v3_make(p.d.angular_velocity[0], p.d.angular_velocity[1], p.d.angular_velocity[2]);
            q4 rot = {p.d.rotation[0], p.d.rotation[1], p.d.rotation[2], p.d.rotation[3]};
            rot = q_mul(q_from_scaled_axis(v3_mul(av, dt)), rot);
            p.d.rotation[0] = rot.x; p.d.rotation[1] = rot.y; p.d.rotation[2] = rot.z; p.d.rotation[3] = rot.w;
            /* :648-650 angular_velocity += (angular_acceleration - angular_drag * angular_velocity) * dt */
            v3 aacc = v3_make(ps->angular_acceleration[0], ps->angular_acceleration[1], ps->angular_acceleration[2]);
            av = v3_add(av, v3_mul(v3_sub(aacc, v3_mul(av, ps->angular_drag)), dt));
            p.d.angular_velocity[0] = av.x; p.d.angular_velocity[1] = av.y; p.d.angular_velocity[2] = av.z;
            /* :652-655 */
            fwo_sample_gradient(&ps->base_color, age_percent, p.d.base_color);
            fwo_sample_gradient(&ps->emissive_color, age_percent, p.d.emissive_color);
            pvec_push(&out, p); /* Some(particle) */
        }
        pvec_drop(src); /* the old Vec is dropped when data.particles[i] is assigned */
        *src = out;
        /* :660-667: the handler would receive `destroyed`; keep it readable for one frame when
         * a handler is registered, drop it otherwise */
        pvec_drop(&s->destroyed[i]);
        if (ps->capture_destroyed) s->destroyed[i] = destroyed;
        else pvec_drop(&destroyed);
    }
}

typedef struct {
    fwo_world *w;
    float dt;
    atomic_size_t next;
} update_job;

static void *update_worker(void *arg) {
    update_job *j = (update_job *)arg;
    for (;;) {
        size_t i = atomic_fetch_add(&j->next, 1);
        if (i >= j->w->n_sp) break;
        update_spawner(j->w, j->w->sp[i], j->dt);
    }
    return NULL;
}

/* ref src/core.rs:577-670 update_particles */
void fwo_update_only(fwo_world *w, float dt, uint32_t n_threads) {
    update_job job;
    job.w = w;
    job.dt = dt;
    atomic_init(&job.next, 0);
    if (n_threads <= 1 || w->n_sp <= 1) {
        update_worker(&job);
        return;
    }
    if (n_threads > 1024) n_threads = 1024;
    if (n_threads > w->n_sp) n_threads = (uint32_t)w->n_sp;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    uint32_t started = 0;
    for (uint32_t t = 0; t + 1 < n_threads; t++)
        if (pthread_create(&th[started], NULL, update_worker, &job) == 0) started++;
    update_worker(&job);
    for (uint32_t t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
}

/* ref src/plugin.rs:55-56: spawn_particles, then update_particles */
void fwo_frame(fwo_world *w, float dt, const fw_spawner_frame_input *in, uint32_t n_in,
               uint32_t n_threads) {
    fwo_spawn_only(w, dt, in, n_in);
    fwo_update_only(w, dt, n_threads);
}

uint64_t fwo_count(const fwo_world *w, uint32_t key, uint32_t type) {
    const spawner *s = find_spawner(w, key);
    if (!s || type >= s->n_types) return 0;
    return s->particles[type].len;
}
uint64_t fwo_total_live(const fwo_world *w) {
    uint64_t n = 0;
    for (size_t i = 0; i < w->n_sp; i++)
        for (uint32_t t = 0; t < w->sp[i]->n_types; t++) n += w->sp[i]->particles[t].len;
    return n;
}
static int read_vec(const pvec *v, fw_particle_data *out, uint64_t cap, uint64_t *n) {
    *n = v->len;
    if (v->len > cap) return 1;
    for (size_t i = 0; i < v->len; i++) out[i] = v->p[i].d;
    return 0;
}
int fwo_read_particles(const fwo_world *w, uint32_t key, uint32_t type, fw_particle_data *out,
                       uint64_t cap, uint64_t *n) {
    const spawner *s = find_spawner(w, key);
    if (!s || type >= s->n_types) return 2;
    return read_vec(&s->particles[type], out, cap, n);
}
int fwo_read_destroyed(const fwo_world *w, uint32_t key, uint32_t type, fw_particle_data *out,
                       uint64_t cap, uint64_t *n) {
    const spawner *s = find_spawner(w, key);
    if (!s || type >= s->n_types) return 2;
    return read_vec(&s->destroyed[type], out, cap, n);
}
int fwo_write_particles(fwo_world *w, uint32_t key, uint32_t type, const fw_particle_data *in, uint64_t n) {
    spawner *s = find_spawner(w, key);
    if (!s || type >= s->n_types) return 2;
    pvec_drop(&s->particles[type]);
    for (uint64_t i = 0; i < n; i++) {
        particle p;
        p.d = in[i];
        p.last_emitted_age = (float *)malloc(sizeof(float) * (s->n_emitters ? s->n_emitters : 1));
        for (uint32_t k = 0; k < s->n_emitters; k++) p.last_emitted_age[k] = -FLT_MAX;
        pvec_push(&s->particles[type], p);
    }
    return 0;
}
int fwo_mark_finished_notified(fwo_world *w, uint32_t key) {
    spawner *s = find_spawner(w, key);
    if (!s) return 2;
    s->finished_notified = 1; /* ref src/core.rs:685 */
    return 0;
}
/* ref src/core.rs:674-688 notify_finished_particle_spawners: the condition of :679-682 */
int fwo_status(fwo_world *w, uint32_t key, fw_spawner_status *out) {
    spawner *s = find_spawner(w, key);
    if (!s) return 2;
    memset(out, 0, sizeof(*out));
    int all_empty = 1;
    uint64_t live = 0;
    for (uint32_t t = 0; t < s->n_types; t++) { if (s->particles[t].len) all_empty = 0; live += s->particles[t].len; }
    out->active = (uint32_t)spawner_active(s);
    out->all_empty = (uint32_t)all_empty;
    out->live_particles = live;
    out->finished = (all_empty && !out->active && s->initialized && !s->finished_notified) ? 1u : 0u;
    out->finished_notified = (uint32_t)s->finished_notified;
    return 0;
}
/* ref src/render.rs:677-692 update_aabbs (world-space min/max of position -/+ scale) */
int fwo_read_aabb(const fwo_world *w, uint32_t key, float mn[3], float mx[3], uint32_t *empty) {
    const spawner *s = find_spawner(w, key);
    if (!s) return 2;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint64_t n = 0;
    for (uint32_t t = 0; t < s->n_types; t++)
        for (size_t i = 0; i < s->particles[t].len; i++) {
            const fw_particle_data *d = &s->particles[t].p[i].d;
            for (int k = 0; k < 3; k++) {
                lo[k] = fminf(lo[k], d->position[k] - d->scale);
                hi[k] = fmaxf(hi[k], d->position[k] + d->scale);
            }
            n++;
        }
    memcpy(mn, lo, sizeof(lo)); memcpy(mx, hi, sizeof(hi));
    *empty = (n == 0);
    return 0;
}
