// fw_math.cuh -- device-side fp32 helpers of the particle kernels.
//
// This translation unit is compiled with -fmad=false: every a*b+c below is two correctly
// rounded IEEE operations, in the operation order of the reference's Rust expressions
// (glam 0.32 scalar formulas), so position / velocity / age / scale / colours can be compared
// bit-for-bit with a CPU evaluation of the same expressions. Sine and cosine (rotation, spawn
// shapes) are the library's own IEEE-only definition, include/fw_sincos.h, not CUDA's sincosf:
// a CPU replay that compiles the same header gets the same bits.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/fw_sincos.h"
#include "fw_internal.h"

namespace fw {

struct V3 {
    float x, y, z;
};
struct Q4 {
    float x, y, z, w;
};

__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
// glam Vec3::normalize = self * length_recip()
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / length(a)); }
// glam Vec3::normalize_or_zero
__device__ __forceinline__ V3 normalize_or_zero(V3 a) {
    float rcp = 1.0f / length(a);
    if (isfinite(rcp) && rcp > 0.0f) return a * rcp;
    return v3(0.0f, 0.0f, 0.0f);
}
// glam Vec3::project_onto / reject_from
__device__ __forceinline__ V3 project_onto(V3 a, V3 rhs) {
    float other_len_sq_rcp = 1.0f / dot(rhs, rhs);
    return (rhs * dot(a, rhs)) * other_len_sq_rcp;
}
__device__ __forceinline__ V3 reject_from(V3 a, V3 rhs) { return a - project_onto(a, rhs); }

// glam Quat::mul_quat (scalar form)
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    Q4 r;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return r;
}
// glam Quat::mul_vec3 (scalar form)
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
    float w = q.w;
    V3 b = v3(q.x, q.y, q.z);
    float b2 = dot(b, b);
    V3 r = v * (w * w - b2);
    r = r + b * (dot(v, b) * 2.0f);
    r = r + cross(b, v) * (w * 2.0f);
    return r;
}
__device__ __forceinline__ Q4 qconj(Q4 q) { return Q4{-q.x, -q.y, -q.z, q.w}; }
__device__ __forceinline__ Q4 q_from_axis_angle(V3 axis, float angle) {
    float s, c;
    fw_sincosf(angle * 0.5f, &s, &c);
    return Q4{axis.x * s, axis.y * s, axis.z * s, c};
}
// glam Quat::from_scaled_axis (reference src/core.rs:646)
__device__ __forceinline__ Q4 q_from_scaled_axis(V3 v) {
    float len = length(v);
    if (len == 0.0f) return Q4{0.0f, 0.0f, 0.0f, 1.0f};
    return q_from_axis_angle(v / len, len);
}
__device__ __forceinline__ Q4 q_from_rotation_y(float angle) {
    float s, c;
    fw_sincosf(angle * 0.5f, &s, &c);
    return Q4{0.0f, s, 0.0f, c};
}
__device__ __forceinline__ V3 any_orthonormal(V3 n) {
    float sign = copysignf(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    return v3(b, sign + n.y * n.y * a, -n.y);
}
// glam Quat::from_rotation_arc
__device__ __forceinline__ Q4 q_from_rotation_arc(V3 from, V3 to) {
    const float kOneMinusEps = 1.0f - 2.0f * FLT_EPSILON;
    float d = dot(from, to);
    if (d > kOneMinusEps) return Q4{0.0f, 0.0f, 0.0f, 1.0f};
    if (d < -kOneMinusEps) return q_from_axis_angle(any_orthonormal(from), 3.14159265358979323846f);
    V3 c = cross(from, to);
    Q4 q{c.x, c.y, c.z, 1.0f + d};
    float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    float rcp = 1.0f / len;
    q.x *= rcp;
    q.y *= rcp;
    q.z *= rcp;
    q.w *= rcp;
    return q;
}

// ---- curve cores (bevy_math EvenCore / UnevenCore semantics; reference src/curve.rs)
struct Interp {
    uint32_t lo, hi;
    float s;
    bool between;
};
__device__ __forceinline__ float clamp01(float t) {
    if (!(t > 0.0f)) return 0.0f;
    if (t > 1.0f) return 1.0f;
    return t;
}
__device__ __forceinline__ Interp even_interp(uint32_t n, float t) {
    Interp r{0u, 0u, 0.0f, false};
    uint32_t subdivs = n - 1u;
    float step = (1.0f - 0.0f) / (float)subdivs;
    float steps_taken = (t - 0.0f) / step;
    if (!(steps_taken > 0.0f)) return r;
    if (steps_taken >= (float)subdivs) {
        r.lo = n - 1u;
        return r;
    }
    float fl = floorf(steps_taken);
    r.lo = (uint32_t)fl;
    r.hi = r.lo + 1u;
    r.s = steps_taken - fl;
    r.between = (r.s != 0.0f);
    return r;
}
__device__ __forceinline__ Interp uneven_interp(const float *times, uint32_t n, float t) {
    Interp r{0u, 0u, 0.0f, false};
    uint32_t idx = 0;
    while (idx < n && times[idx] < t) idx++;
    if (idx < n && times[idx] == t) {
        r.lo = idx;
        return r;
    }
    if (idx == 0u) return r;
    if (idx >= n) {
        r.lo = n - 1u;
        return r;
    }
    float t_lower = times[idx - 1u], t_upper = times[idx];
    r.lo = idx - 1u;
    r.hi = idx;
    r.s = (t - t_lower) / (t_upper - t_lower);
    r.between = true;
    return r;
}
// FireworkCurve<f32>::sample_clamped (reference src/core.rs:603)
__device__ __forceinline__ float sample_curve(const DevCurve &c, float t) {
    if (c.kind == FW_CURVE_CONSTANT) return c.values[0];
    t = clamp01(t);
    Interp it = (c.kind == FW_CURVE_EVEN) ? even_interp(c.n, t) : uneven_interp(c.times, c.n, t);
    float a = c.values[it.lo];
    if (!it.between) return a;
    float b = c.values[it.hi];
    return a + (b - a) * it.s;
}
// FireworkGradient<LinearRgba>::sample_clamped (reference src/core.rs:460-461,653,655);
// bevy_color Mix: a*(1-s) + b*s per channel
__device__ __forceinline__ float4 sample_gradient(const DevGradient &g, float t) {
    const float4 *colors = g.colors;
    if (g.kind == FW_CURVE_CONSTANT) return colors[0];
    t = clamp01(t);
    Interp it = (g.kind == FW_CURVE_EVEN) ? even_interp(g.n, t) : uneven_interp(g.times, g.n, t);
    float4 a = colors[it.lo];
    if (!it.between) return a;
    float4 b = colors[it.hi];
    float nf = 1.0f - it.s;
    return make_float4(a.x * nf + b.x * it.s, a.y * nf + b.y * it.s, a.z * nf + b.z * it.s,
                       a.w * nf + b.w * it.s);
}

// ---- Philox4x32-10 (Salmon et al. SC'11) -- the spawn RNG protocol of DESIGN.md
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ---- order-preserving float <-> uint (for atomicMin/Max and REDUX on AABB bounds)
__device__ __forceinline__ uint32_t enc_f32(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- ray casting against the static colliders (avian cast_ray / parry shapes semantics,
// closest hit, solid = true). Lowest collider index wins ties.
__device__ __forceinline__ bool ray_cuboid_local(V3 he, V3 o, V3 d, float max_toi, float &toi, V3 &normal) {
    float tmax = FLT_MAX, tmin = -FLT_MAX;
    int near_side = 0;
    bool near_diag = false;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, hh[3] = {he.x, he.y, he.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float mn = -hh[i], mx = hh[i];
        if (dd[i] == 0.0f) {
            if (oo[i] < mn || oo[i] > mx) return false;
        } else {
            float denom = 1.0f / dd[i];
            float t_near = (mn - oo[i]) * denom;
            float t_far = (mx - oo[i]) * denom;
            bool flip = t_near > t_far;
            if (flip) {
                float tmp = t_near;
                t_near = t_far;
                t_far = tmp;
            }
            if (t_near > tmin) {
                tmin = t_near;
                near_side = flip ? -(i + 1) : (i + 1);
                near_diag = false;
            } else if (t_near == tmin) {
                near_diag = true;
            }
            if (t_far < tmax) tmax = t_far;
            if (tmax < 0.0f || tmin > tmax) return false;
        }
    }
    if (tmin < 0.0f) {
        toi = 0.0f;
        normal = v3(0.0f, 0.0f, 0.0f);
        return true;
    }
    if (tmin <= max_toi) {
        float nx = 0.0f, ny = 0.0f, nz = 0.0f;
        if (near_diag) {
            // the ray enters through an edge or a corner (two slabs at the same parameter): parry's
            // clip_aabb_line reports -dir.normalize() (a zero normal here would turn the bounce of
            // src/core.rs:778-784 into NaNs)
            const V3 nd = normalize(d);
            nx = -nd.x;
            ny = -nd.y;
            nz = -nd.z;
        } else if (near_side != 0) {
            float sgn = near_side < 0 ? 1.0f : -1.0f;
            int ax = (near_side < 0 ? -near_side : near_side) - 1;
            if (ax == 0) nx = sgn;
            else if (ax == 1) ny = sgn;
            else nz = sgn;
        }
        toi = tmin;
        normal = v3(nx, ny, nz);
        return true;
    }
    return false;
}
__device__ __forceinline__ bool ray_ball_local(float radius, V3 o, V3 d, float max_toi, float &toi, V3 &normal) {
    float a = dot(d, d);
    float b = dot(o, d);
    float c = dot(o, o) - radius * radius;
    float t;
    bool inside = false;
    if (a == 0.0f) {
        if (c > 0.0f) return false;
        t = 0.0f;
        inside = true;
    } else if (c > 0.0f && b > 0.0f) {
        return false;
    } else {
        float delta = b * b - a * c;
        if (delta < 0.0f) return false;
        t = (-b - sqrtf(delta)) / a;
        if (t <= 0.0f) {
            t = 0.0f;
            inside = true;
        }
    }
    if (!(t <= max_toi)) return false;
    V3 pos = o + d * t;
    V3 n = normalize(pos);
    if (inside) n = v3(-n.x, -n.y, -n.z);
    toi = t;
    normal = n;
    return true;
}
// Cylinder and cone: the analytic solid of revolution about +Y whose radius goes linearly from r0
// at y = -h to r1 at y = +h (cylinder r0 = r1, cone r1 = 0); defined by this build, see the oracle.
// (not inlined: the cuboid / sphere scenes should not pay registers for it)
__device__ __noinline__ bool ray_frustum_local(float r0, float r1, float h, V3 o, V3 d, float max_toi, float &toi, V3 &normal) {
    const float s = (r1 - r0) / (2.0f * h), c0 = (r0 + r1) * 0.5f;
    const float ro = c0 + s * o.y;
    if (o.y >= -h && o.y <= h && ro >= 0.0f && o.x * o.x + o.z * o.z <= ro * ro) {
        toi = 0.0f;
        normal = v3(0.0f, 0.0f, 0.0f);
        return true;
    }
    bool found = false;
    float best = 0.0f;
    V3 bn = v3(0.0f, 0.0f, 0.0f);
    if (d.y != 0.0f) {
        const float inv = 1.0f / d.y;
#pragma unroll
        for (int cap = 0; cap < 2; cap++) {
            const float yc = cap ? h : -h, rc = cap ? r1 : r0;
            const float t = (yc - o.y) * inv;
            const float x = o.x + d.x * t, z = o.z + d.z * t;
            if (rc > 0.0f && t >= 0.0f && x * x + z * z <= rc * rc && (!found || t < best)) {
                found = true;
                best = t;
                bn = v3(0.0f, cap ? 1.0f : -1.0f, 0.0f);
            }
        }
    }
    const float A = d.x * d.x + d.z * d.z - (s * s) * (d.y * d.y);
    const float B = o.x * d.x + o.z * d.z - (s * ro) * d.y;
    const float C = o.x * o.x + o.z * o.z - ro * ro;
    float roots[2] = {0.0f, 0.0f};
    int n_roots = 0;
    if (A != 0.0f) {
        const float disc = B * B - A * C;
        if (disc >= 0.0f) {
            // the cancellation-free form (see the oracle): q = -(B + sign(B) sqrt(disc)), roots q / A and C / q
            const float sq = sqrtf(disc);
            const float q = B < 0.0f ? sq - B : -(B + sq);
            if (q != 0.0f) {
                roots[0] = B < 0.0f ? C / q : q / A;
                roots[1] = B < 0.0f ? q / A : C / q;
            } else {
                roots[0] = 0.0f;
                roots[1] = 0.0f;
            }
            n_roots = 2;
        }
    } else if (B != 0.0f) {
        roots[0] = -C / (2.0f * B);
        n_roots = 1;
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float t = roots[k];
        const float y = o.y + d.y * t, rr = c0 + s * y;
        if (k < n_roots && t >= 0.0f && y >= -h && y <= h && rr >= 0.0f && (!found || t < best)) {
            const float x = o.x + d.x * t, z = o.z + d.z * t;
            const V3 g = v3(x, -(s * rr), z);
            const float len = length(g);
            found = true;
            best = t;
            bn = len > 0.0f ? g * (1.0f / len) : v3(0.0f, s < 0.0f ? 1.0f : -1.0f, 0.0f);
        }
    }
    if (!found || !(best <= max_toi)) return false;
    toi = best;
    normal = bn;
    return true;
}
// Capsule: the segment (0,-h,0)..(0,h,0) swept by a ball of radius r; the oracle's ray_capsule_local
// operation by operation (candidates: side, bottom cap, top cap; entering roots only).
__device__ __noinline__ bool ray_capsule_local(float r, float h, V3 o, V3 d, float max_toi, float &toi, V3 &normal) {
    const float yc = fminf(fmaxf(o.y, -h), h);
    const float dy0 = o.y - yc;
    if (o.x * o.x + dy0 * dy0 + o.z * o.z <= r * r) {
        toi = 0.0f;
        normal = v3(0.0f, 0.0f, 0.0f);
        return true;
    }
    bool found = false;
    float best = 0.0f;
    V3 bn = v3(0.0f, 0.0f, 0.0f);
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
                    found = true;
                    best = t;
                    bn = v3((o.x + d.x * t) * inv_r, 0.0f, (o.z + d.z * t) * inv_r);
                }
            }
        }
    }
#pragma unroll
    for (int cap = 0; cap < 2; cap++) {
        const float cy = cap ? h : -h;
        const V3 oc = v3(o.x, o.y - cy, o.z);
        const float a = dot(d, d), b = dot(oc, d), c = dot(oc, oc) - r * r;
        if (a != 0.0f) {
            const float disc = b * b - a * c;
            if (disc >= 0.0f) {
                const float t = (-b - sqrtf(disc)) / a;
                const V3 p = oc + d * t;
                const bool outer = cap ? p.y >= 0.0f : p.y <= 0.0f;
                if (t >= 0.0f && outer && (!found || t < best)) {
                    found = true;
                    best = t;
                    bn = p * inv_r;
                }
            }
        }
    }
    if (!found || !(best <= max_toi)) return false;
    toi = best;
    normal = bn;
    return true;
}
// Broad phase (blob layout: BroadPhaseHeader in fw_internal.h). Every candidate passes a box test
// of the ray segment's AABB against the collider's inflated world AABB; the boxes are inflated on
// the host by far more than any fp32 rounding of the exact test, so a collider is skipped only
// when the exact test below could not report a hit within max_distance: the result is identical
// to testing every collider in index order (the CPU oracle does exactly that; equal distances
// resolve to the lowest collider index). NaN never culls. Two ways to enumerate candidates:
//   * grid: a segment that spans at most 2 cells per axis (the usual case: |v| dt is a fraction of
//     a collider) reads the CSR lists of its <= 8 cells plus the short list of "big" colliders;
//     a collider listed in several of those cells is taken in the first one its box overlaps;
//   * BVH: anything else (long segments, non-finite input, grid disabled) walks the BVH in
//     depth-first order without a stack: a node whose box misses the segment's box jumps to its
//     skip link (a leaf's is k + 1).
//
// Warp-synchronous: ALL 32 lanes of the warp must call (lanes without a ray pass act = false).
// The lanes of a warp see different candidates, so enumeration only collects them into a small
// per-lane queue in shared memory (queue[j * kUpdateThreads], j < kCandQueue); the long exact
// test then runs over the queues with the warp converged: a warp pays max-over-lanes(candidates)
// exact tests instead of one per divergent loop trip (ncu on C5 before the split: 2.7 active
// lanes per instruction in the exact test, profiles/r1_tuning.md).
// REVOLVED: the collider set contains cylinders / cones (their exact test costs the cuboid /
// sphere scenes registers, so it is compiled in only when the set needs it: C5 0.123 vs 0.129 ms).
constexpr uint32_t kCandQueue = 4;
#ifdef FW_COLLIDE_STATS
__device__ unsigned long long g_collide_stats[16];
#define FW_STAT(i, v) atomicAdd(&g_collide_stats[i], (unsigned long long)(v))
#else
#define FW_STAT(i, v)
#endif
__device__ __forceinline__ float grid_coord(float x, float lo, float inv_cell) { return floorf((x - lo) * inv_cell); }
template <bool REVOLVED>
__device__ __forceinline__ bool cast_ray(const fw_collider *__restrict__ colliders, const uint8_t *__restrict__ bp,
                                         const fw_collision_settings &cs, bool act, V3 o, V3 d, float max_distance, uint32_t *queue,
                                         float &distance, V3 &normal) {
    const uint32_t filter_mask = cs.filter_mask;
    if (bp == nullptr) { // no collider set was ever uploaded
        distance = 0.0f;
        normal = v3(0.0f, 0.0f, 0.0f);
        return false;
    }
    const BroadPhaseHeader &h = *reinterpret_cast<const BroadPhaseHeader *>(bp);
    const float4 *__restrict__ bvh = reinterpret_cast<const float4 *>(bp + h.nodes_off);
    const float4 *__restrict__ leaf = reinterpret_cast<const float4 *>(bp + h.leaf_off);
    const uint32_t n_nodes = h.n_nodes;
    bool found = false;
    float best = 0.0f;
    uint32_t best_i = 0xFFFFFFFFu;
    V3 best_n = v3(0.0f, 0.0f, 0.0f);
    const V3 e = o + d * max_distance;
    const V3 slo = v3(fminf(o.x, e.x), fminf(o.y, e.y), fminf(o.z, e.z));
    const V3 shi = v3(fmaxf(o.x, e.x), fmaxf(o.y, e.y), fmaxf(o.z, e.z));
    const bool cull_ok = isfinite(e.x) && isfinite(e.y) && isfinite(e.z) && isfinite(o.x) && isfinite(o.y) && isfinite(o.z);

    // ---- which enumeration; grid state: base cell (8 bits per axis) | span bits << 24
    bool use_grid = false;
    uint32_t base = 0, sub = 8u /* next cell sub-index, 8 = no more cells */, cur_sub = 0, big_i = 0, p = 0, p_end = 0;
    if (act && cull_ok && h.use_grid != 0u) {
        const float dx = (float)h.dim[0], dy = (float)h.dim[1], dz = (float)h.dim[2];
        // cell coordinates as floats, clamped to [-1, dim] so that the int conversion is safe
        const float ax = fminf(fmaxf(grid_coord(slo.x, h.lo[0], h.inv_cell[0]), -1.0f), dx), bx = fminf(fmaxf(grid_coord(shi.x, h.lo[0], h.inv_cell[0]), -1.0f), dx);
        const float ay = fminf(fmaxf(grid_coord(slo.y, h.lo[1], h.inv_cell[1]), -1.0f), dy), by = fminf(fmaxf(grid_coord(shi.y, h.lo[1], h.inv_cell[1]), -1.0f), dy);
        const float az = fminf(fmaxf(grid_coord(slo.z, h.lo[2], h.inv_cell[2]), -1.0f), dz), bz = fminf(fmaxf(grid_coord(shi.z, h.lo[2], h.inv_cell[2]), -1.0f), dz);
        const int x0 = max((int)ax, 0), x1 = min((int)bx, (int)h.dim[0] - 1);
        const int y0 = max((int)ay, 0), y1 = min((int)by, (int)h.dim[1] - 1);
        const int z0 = max((int)az, 0), z1 = min((int)bz, (int)h.dim[2] - 1);
        if (x1 - x0 <= 1 && y1 - y0 <= 1 && z1 - z0 <= 1) {
            use_grid = true;
            if (x1 >= x0 && y1 >= y0 && z1 >= z0) { // else: the segment misses the grid, only the big list
                base = (uint32_t)x0 | ((uint32_t)y0 << 8) | ((uint32_t)z0 << 16) |
                       ((uint32_t)(x1 - x0) << 24) | ((uint32_t)(y1 - y0) << 25) | ((uint32_t)(z1 - z0) << 26);
                sub = 0u;
            }
        }
    }
    const uint32_t *__restrict__ big = reinterpret_cast<const uint32_t *>(bp + h.big_off);
    const uint32_t *__restrict__ cell_start = reinterpret_cast<const uint32_t *>(bp + h.cell_off);
    const uint32_t *__restrict__ items = reinterpret_cast<const uint32_t *>(bp + h.items_off);
    const uint32_t n_big = h.n_big, span = base >> 24;
    uint32_t k = (act && !use_grid) ? 0u : n_nodes;
    bool more;
#ifdef FW_COLLIDE_STATS
    uint32_t st_total = 0, st_nonempty = 0;
    FW_STAT(0, act ? 1 : 0);                 // rays
    FW_STAT(1, (act && use_grid) ? 1 : 0);   // rays on the grid path
    if ((threadIdx.x & 31u) == 0) FW_STAT(2, 1); // warp-level casts
    if (act && use_grid && sub < 8u) {
        for (uint32_t s2 = 0;;) {
            const uint32_t cx = (base & 255u) + (s2 & 1u), cy = ((base >> 8) & 255u) + ((s2 >> 1) & 1u), cz = ((base >> 16) & 255u) + (s2 >> 2);
            const uint32_t cell = (cz * h.dim[1] + cy) * h.dim[0] + cx;
            if (cell_start[cell + 1] > cell_start[cell]) st_nonempty = 1;
            s2 = (s2 - span) & span;
            if (s2 == 0u) break;
        }
    }
    FW_STAT(3, st_nonempty);                 // rays with a non-empty cell
#endif
    do {
        uint32_t cnt = 0;
        if (use_grid) {
            while (cnt < kCandQueue) {
                uint32_t cand;
                bool from_cell = false;
                if (big_i < n_big) {
                    cand = __ldg(big + big_i++);
                } else if (p < p_end) {
                    cand = __ldg(items + p++);
                    from_cell = true;
                } else if (sub < 8u) {
                    const uint32_t cx = (base & 255u) + (sub & 1u), cy = ((base >> 8) & 255u) + ((sub >> 1) & 1u),
                                   cz = ((base >> 16) & 255u) + (sub >> 2);
                    const uint32_t cell = (cz * h.dim[1] + cy) * h.dim[0] + cx;
                    p = __ldg(cell_start + cell);
                    p_end = __ldg(cell_start + cell + 1u);
                    cur_sub = sub;
                    sub = (sub - span) & span; // next subset of the span bits; back at 0 = all cells seen
                    if (sub == 0u) sub = 8u;
                    continue;
                } else {
                    break;
                }
                const float4 blo = __ldg(leaf + 2u * cand), bhi = __ldg(leaf + 2u * cand + 1u);
                const bool disjoint = (shi.x < blo.x) | (slo.x > bhi.x) | (shi.y < blo.y) | (slo.y > bhi.y) | (shi.z < blo.z) | (slo.z > bhi.z);
                if (disjoint | ((__float_as_uint(blo.w) & filter_mask) == 0u)) continue;
                if (from_cell && span != 0u) {
                    // listed in several of our cells: take it in the first one its box overlaps
                    const uint32_t fx = grid_coord(blo.x, h.lo[0], h.inv_cell[0]) > (float)(base & 255u) ? 1u : 0u;
                    const uint32_t fy = grid_coord(blo.y, h.lo[1], h.inv_cell[1]) > (float)((base >> 8) & 255u) ? 1u : 0u;
                    const uint32_t fz = grid_coord(blo.z, h.lo[2], h.inv_cell[2]) > (float)((base >> 16) & 255u) ? 1u : 0u;
                    if (((fx | (fy << 1) | (fz << 2)) & span) != cur_sub) continue;
                }
                queue[(cnt++) * kUpdateThreads] = cand;
            }
            more = (big_i < n_big) | (p < p_end) | (sub < 8u);
        } else {
            while (k < n_nodes && cnt < kCandQueue) {
                const float4 blo = __ldg(bvh + 2u * k), bhi = __ldg(bvh + 2u * k + 1u);
                const uint32_t leaf_i = __float_as_uint(bhi.w), link = __float_as_uint(blo.w);
                // (bitwise, not short-circuit: one predicate chain instead of six branches)
                const bool disjoint = cull_ok & ((shi.x < blo.x) | (slo.x > bhi.x) | (shi.y < blo.y) | (slo.y > bhi.y) | (shi.z < blo.z) | (slo.z > bhi.z));
                const bool inner = leaf_i == 0xFFFFFFFFu;
                k = (inner & disjoint) ? link : k + 1u;
                if (!inner & !disjoint & ((link & filter_mask) != 0u)) queue[(cnt++) * kUpdateThreads] = leaf_i;
            }
            more = k < n_nodes;
        }
        const uint32_t rounds = __reduce_max_sync(0xffffffffu, cnt);
#ifdef FW_COLLIDE_STATS
        st_total += cnt;
        FW_STAT(4, cnt);                              // exact tests
        if ((threadIdx.x & 31u) == 0) { FW_STAT(5, rounds); FW_STAT(6, 1); } // warp exact rounds, warp enumeration rounds
#endif
        for (uint32_t j = 0; j < rounds; j++) {
            if (j < cnt) {
                const uint32_t cand = queue[j * kUpdateThreads];
                const fw_collider &c = colliders[cand];
                Q4 rot{c.rotation[0], c.rotation[1], c.rotation[2], c.rotation[3]};
                Q4 inv = qconj(rot);
                V3 tr = v3(c.translation[0], c.translation[1], c.translation[2]);
                V3 ol = qrot(inv, o - tr);
                V3 dl = qrot(inv, d);
                float toi;
                V3 nl;
                bool hit;
                if (c.kind == FW_COLLIDER_SPHERE) hit = ray_ball_local(c.half_extents[0], ol, dl, max_distance, toi, nl);
                else if (REVOLVED && c.kind == FW_COLLIDER_CYLINDER) hit = ray_frustum_local(c.half_extents[0], c.half_extents[0], c.half_extents[1], ol, dl, max_distance, toi, nl);
                else if (REVOLVED && c.kind == FW_COLLIDER_CONE) hit = ray_frustum_local(c.half_extents[0], 0.0f, c.half_extents[1], ol, dl, max_distance, toi, nl);
                else if (REVOLVED && c.kind == FW_COLLIDER_CAPSULE) hit = ray_capsule_local(c.half_extents[0], c.half_extents[1], ol, dl, max_distance, toi, nl);
                else hit = ray_cuboid_local(v3(c.half_extents[0], c.half_extents[1], c.half_extents[2]), ol, dl, max_distance, toi, nl);
                // SpatialQueryFilter::excluded_entities (src/core.rs:247,764): a listed collider is not seen
                for (uint32_t x = 0; x < cs.n_excluded; x++) hit = hit && cs.excluded_keys[x] != c.key;
                if (hit && (!found || toi < best || (toi == best && cand < best_i))) {
                    found = true;
                    best = toi;
                    best_i = cand;
                    best_n = qrot(rot, nl);
                }
            }
            __syncwarp();
        }
    } while (__any_sync(0xffffffffu, more));
#ifdef FW_COLLIDE_STATS
    FW_STAT(7, st_total > 0 ? 1 : 0); // rays with at least one exact test
    FW_STAT(8, found ? 1 : 0);        // rays that hit
    FW_STAT(9, __any_sync(0xffffffffu, st_total > 0) && (threadIdx.x & 31u) == 0 ? 1 : 0); // warp casts with any exact test
#endif
    distance = best;
    normal = best_n;
    return found;
}

// reference src/core.rs:744-800 particle_collision. Warp-synchronous like cast_ray: every lane of
// the warp calls, lanes without a live particle pass active = false (their pos / vel stay untouched).
template <bool REVOLVED>
__device__ __forceinline__ void particle_collision(const fw_collider *__restrict__ colliders, const uint8_t *__restrict__ broadphase,
                                                   const fw_collision_settings &cs, bool active, V3 &pos,
                                                   V3 &vel, float delta, uint32_t *queue, bool &should_destroy) {
    const float orig_delta = delta;
    int n_steps = 0;
    should_destroy = false;
    bool go = active && delta > 0.0f; // `while delta > 0 && n_steps < 4` (:755)
    while (__any_sync(0xffffffffu, go)) {
        float len = length(vel);
        V3 dir = (isfinite(len) && len > 0.0f) ? vel / len : v3(0.0f, 1.0f, 0.0f);
        float distance;
        V3 hit_normal;
        const bool hit = cast_ray<REVOLVED>(colliders, broadphase, cs, go, pos, dir, length(vel) * delta, queue, distance, hit_normal);
        if (go) {
            if (hit) {
                if (distance == 0.0f) {
                    V3 normal = hit_normal;
                    if (is_zero(normal)) {
                        if (!is_zero(vel)) normal = normalize(vel);
                        else normal = v3(0.0f, 1.0f, 0.0f);
                    }
                    pos = pos + (normal * fmaxf(length(vel), 1.0f)) * delta;
                } else {
                    pos = pos + normalize_or_zero(vel) * distance;
                    V3 vel_reject = reject_from(vel, hit_normal);
                    V3 vel_project = project_onto(vel, hit_normal);
                    float friction_dv = fminf(length(vel_project), length(vel_reject)) * cs.friction;
                    vel = (vel_reject - normalize_or_zero(vel_reject) * friction_dv) - vel_project * cs.restitution;
                    pos = pos + hit_normal * 0.0001f;
                    delta = delta - distance;
                    if (delta < 0.0f) delta = 0.0f;
                    if (delta > orig_delta) delta = orig_delta;
                }
                should_destroy = cs.destroy_on_collision != 0u;
            } else {
                pos = pos + vel * delta;
                delta = 0.0f;
            }
            n_steps += 1;
            go = !should_destroy && delta > 0.0f && n_steps < 4; // early return at :788-791
        }
    }
}

} // namespace fw
