// fw_kernels.cu -- sm_100a kernels of the particle path.
//
//   plan_kernel     applies last frame's deaths to every ring (head/count), appends a phase's
//                   Global spawn counts, builds the per-variant tile prefix tables
//   spawn_kernel    reference src/core.rs:437-469 + src/emission_shape.rs:18-39, one thread per
//                   new particle, Philox4x32-10 counter-based draws
//   nested_*        reference src/core.rs:471-546: per-parent emission counts, scan, children
//   update_kernel   reference src/core.rs:591-658 (+ :744-800 when COLLIDE), fused with death
//                   handling (FIFO ring advance, or stable compaction out of place inside the ring
//                   with deaths precounted by count_kernel / scan_kernel), the destroyed-
//                   particle stream (:588,597,637) and the per-stream AABB (src/render.rs:677-703);
//                   tiles dealt round robin, settings staged by bulk async copies: rotating and
//                   colliding streams
//   update_static_kernel  the same for static streams without a sweep (C1..C4 of BASELINE.json):
//                   32 B in / 48 B out per particle, stream segments instead of tiles, no barrier
//   pack kernels    assemble the 64-byte ParticleInstance rows (src/render.rs:95-115) of the
//                   live particles into one contiguous buffer (render extract / all-gather)
//   gather/scatter  ParticleData rows of one stream <-> the SoA packs (host mirror)
//
// Built with -fmad=false (see fw_math.cuh).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "fw_math.cuh"

// tuning knobs (see profiles/r1_tuning.md for the measurements behind the defaults)
#ifndef FW_MINB
// minimum resident CTAs per SM asked of ptxas for the streaming update kernels: 5 x 256 threads
// = 40 warps/SM at <= 48 registers, no spills. Measured on C3 (10 M particles): 4 CTAs/SM
// 0.320 ms, 5 CTAs/SM 0.261 ms, 6 CTAs/SM (40 regs, spills) 0.275 ms.
#define FW_MINB 5
#endif
#ifndef FW_MINB_STATIC
// static FIFO streams (32 B read + 48 B written per particle): half the bytes in flight per thread,
// so more resident threads are asked for
#define FW_MINB_STATIC 5
#endif

#ifndef FW_MINB_COMPACT
#define FW_MINB_COMPACT 4 // 64 registers; C3r (precounted path): 4 CTAs/SM 0.412 ms, 5 (48 regs, spills) 0.421 ms, 6 0.468 ms
#endif
#ifndef FW_MINB_COLLIDE
#define FW_MINB_COLLIDE 3 // the collision variants are compute-bound and need ~80 registers
#endif
#ifndef FW_CS
#define FW_CS 0 // cache operators on the particle packs: 0 default (generic ld/st), 1 .cs, 2 .cg, 3 .ca/.wb
#endif

namespace fw {

template <typename T>
__device__ __forceinline__ T ld_pack(const T *p) {
#if FW_CS == 1
    return __ldcs(p);
#elif FW_CS == 2
    return __ldcg(p);
#elif FW_CS == 3
    return __ldca(p);
#else
    return *p;
#endif
}
template <typename T>
__device__ __forceinline__ void st_pack(T *p, T v) {
#if FW_CS == 1
    __stcs(p, v);
#elif FW_CS == 2
    __stcg(p, v);
#elif FW_CS == 3
    __stwb(p, v);
#else
    *p = v;
#endif
}

__device__ __forceinline__ uint32_t wrap(uint32_t x, uint32_t cap) { return x >= cap ? x - cap : x; }
__device__ __forceinline__ bool is_fifo(uint32_t variant) { return variant_is_fifo(variant); }
// A compacting ring compacts OUT OF PLACE inside its own ring: frame f reads [head, head + n) and
// writes the survivors behind it, to [head + n, ...), so it may only ever be half full. After the
// frame its live particles start at head + count (a FIFO ring's: at head + dead).
__device__ __forceinline__ uint32_t usable_capacity(const StreamDesc &d) { return is_fifo(d.variant) ? d.capacity : d.capacity / 2u; }
__device__ __forceinline__ uint32_t live_first(const StreamDesc &d, const StreamState &st) {
    return wrap(st.head + (is_fifo(d.variant) ? st.dead : st.count), d.capacity);
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
constexpr float kF32Min = -3.402823466e+38f; // f32::MIN, initial last_emitted_age (src/core.rs:467)

// ------------------------------------------------------------------------------------------
// block-wide inclusive scan of one uint per thread (1024 threads max)
__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t &total) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31u) >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += n;
    }
    if (lane == 31u) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nwarps ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += n;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    total = warp_sums[nwarps - 1u];
    uint32_t r = v + (warp ? warp_sums[warp - 1u] : 0u);
    __syncthreads();
    return r;
}

constexpr uint32_t kErrOverflow = 1u;   // a ring was full: spawns were dropped
constexpr uint32_t kErrLookback = 2u;   // look-back table too small
constexpr uint32_t kErrNestedCap = 4u;  // a parent wanted to emit more than the planned bound

__global__ void __launch_bounds__(1024) plan_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t variant_mask,
                                                    uint32_t what, uint32_t phase) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t n_slots = f.header->n_slots;
    const uint32_t stride = t.slots_cap + 1u;
    if (what & (kPlanDeaths | kPlanAppend)) {
        const uint32_t *spawn_per_slot = f.spawn_per_slot + (size_t)phase * n_slots;
        for (uint32_t s = threadIdx.x; s < n_slots; s += blockDim.x) {
            const StreamDesc d = t.descs[s];
            if (d.capacity == 0u) continue;
            StreamState st = (what & kPlanDeaths) ? t.states_prev[s] : t.states[s];
            if (what & kPlanDeaths) { // last frame's buffer -> this frame's buffer
                st.head = live_first(d, st);
                st.count -= st.dead;
                st.dead = 0u;
                st.aabb_min_inv[0] = st.aabb_min_inv[1] = st.aabb_min_inv[2] = 0u;
                st.aabb_max[0] = st.aabb_max[1] = st.aabb_max[2] = 0u;
            }
            if (what & kPlanAppend) {
                uint32_t spawn = spawn_per_slot[s];
                const uint32_t room = usable_capacity(d) - st.count;
                if (spawn > room) {
                    st.overflow += spawn - room;
                    spawn = room;
                    atomicOr(&t.plan->error_flags, kErrOverflow);
                }
                st.spawn_base = st.count;
                st.count += spawn;
            }
            t.states[s] = st;
        }
        __syncthreads();
    }
    if (!(what & kPlanTiles)) return;
    uint32_t base_total = 0, my_total = 0;
    for (uint32_t v = 0; v < kNumVariants; v++) {
        uint32_t carry = 0;
        if (variant_mask & (1u << v)) {
            uint32_t *prefix = t.tile_prefix + (size_t)v * stride;
            for (uint32_t chunk = 0; chunk < n_slots; chunk += blockDim.x) {
                const uint32_t s = chunk + threadIdx.x;
                uint32_t tiles = 0;
                if (s < n_slots) {
                    const StreamDesc d = t.descs[s];
                    if (d.capacity != 0u && d.variant == v) {
                        const uint32_t n = t.states[s].count;
                        tiles = n ? (n + (t.states[s].head & 31u) + kTile - 1u) / kTile : 0u; // slot-aligned tiles, see update_kernel
                        my_total += n;
                    }
                }
                uint32_t total;
                const uint32_t incl = block_inclusive_scan(tiles, warp_sums, total);
                if (s < n_slots) prefix[s] = carry + incl - tiles;
                carry += total;
            }
            if (threadIdx.x == 0) {
                prefix[n_slots] = carry; // sentinel: total tiles of the variant
                if (!variant_is_fifo(v) && base_total + carry > t.lookback_capacity)
                    atomicOr(&t.plan->error_flags, kErrLookback);
            }
        }
        if (threadIdx.x == 0) {
            t.plan->tile_base[v] = base_total;
            t.plan->n_tiles[v] = carry;
        }
        base_total += carry;
    }
    uint32_t total;
    block_inclusive_scan(my_total, warp_sums, total);
    if (threadIdx.x == 0) t.plan->total_update = total;
}

// ------------------------------------------------------------------------------------------
// The state of a stream for THIS frame as a pure function of last frame's buffer and this
// frame's spawn count (the fast path has no plan kernel): deaths of the last update advance the
// head of a FIFO ring (a compacting ring wrote its survivors behind the particles it read), spawns are
// appended, clamped to the ring's room (the host grows rings before that can happen).
struct Derived {
    uint32_t head, c0, n_update, dropped;
};
__device__ __forceinline__ Derived derive_state(const StreamState &old, const StreamDesc &d, uint32_t spawn) {
    Derived r;
    r.head = live_first(d, old);
    r.c0 = old.count - old.dead;
    const uint32_t room = usable_capacity(d) - r.c0;
    r.dropped = spawn > room ? spawn - room : 0u;
    r.n_update = r.c0 + (spawn - r.dropped);
    return r;
}

// ------------------------------------------------------------------------------------------
// One new particle: the shared body of reference src/core.rs:437-469 (Global: origin = the
// spawner transform, inherited velocity = parent_velocity) and :506-544 (Nested: origin = the
// parent particle). Draws 0..11 in the reference's draw order.
struct ParticleRegs { // one particle, every field of ParticleData (src/core.rs:305-321)
    V3 pos, vel, av;
    Q4 rot;
    float age, lifetime, iscale, scale;
    float4 c0, c1;
};
// ---- one particle <-> the packs its stream keeps (layout: fw_internal.h)
__device__ __forceinline__ void store_particle(const StreamDesc &d, uint32_t slot, const ParticleRegs &p, bool init_lea) {
    const StreamArrays a = stream_arrays(d.base, d.capacity);
    const bool rot = variant_rotates(d.variant);
    a.m0[slot] = make_float4(p.pos.x, p.pos.y, p.pos.z, p.age);
    a.m2[slot] = make_float4(p.vel.x, p.vel.y, p.vel.z, rot ? p.av.x : p.iscale);
    if (rot) {
        a.m1[slot] = make_float4(p.rot.x, p.rot.y, p.rot.z, p.rot.w);
        a.m3[slot] = make_float2(p.av.y, p.av.z);
        a.k[slot] = make_float2(p.lifetime, p.iscale);
    } else if (d.flags & kStoreLife) {
        a.k[slot] = make_float2(p.lifetime, p.age);
    }
    if (d.flags & kStoreBase) a.o0[slot] = p.c0;
    if (d.flags & kStoreEmi) a.o1[slot] = p.c1;
    if (d.flags & kStoreScale) a.o2[slot] = p.scale;
    if (init_lea)
        for (uint32_t j = 0; j < d.n_lea; j++) lea_array(d.base, d.capacity, j)[slot] = kF32Min; // :467
}
__device__ __forceinline__ float load_lifetime(const StreamDesc &d, const StreamArrays &a, const DevParticleSettings &ps, uint32_t slot) {
    return (variant_rotates(d.variant) || (d.flags & kStoreLife)) ? a.k[slot].x : ps.const_lifetime;
}
__device__ __forceinline__ Q4 load_rotation(const StreamDesc &d, const StreamArrays &a, const DevParticleSettings &ps, uint32_t slot) {
    if (variant_rotates(d.variant)) {
        const float4 r = a.m1[slot];
        return Q4{r.x, r.y, r.z, r.w};
    }
    return Q4{ps.const_rotation[0], ps.const_rotation[1], ps.const_rotation[2], ps.const_rotation[3]};
}
__device__ __forceinline__ ParticleRegs load_particle(const StreamDesc &d, const DevParticleSettings &ps, uint32_t slot) {
    const StreamArrays a = stream_arrays(d.base, d.capacity);
    ParticleRegs p;
    const float4 A = a.m0[slot], V = a.m2[slot];
    p.pos = v3(A.x, A.y, A.z);
    p.age = A.w;
    p.vel = v3(V.x, V.y, V.z);
    p.rot = load_rotation(d, a, ps, slot);
    if (variant_rotates(d.variant)) {
        const float2 W = a.m3[slot], K = a.k[slot];
        p.av = v3(V.w, W.x, W.y);
        p.lifetime = K.x;
        p.iscale = K.y;
    } else {
        p.av = v3(0.0f, 0.0f, 0.0f);
        p.iscale = V.w;
        p.lifetime = (d.flags & kStoreLife) ? a.k[slot].x : ps.const_lifetime;
    }
    p.c0 = (d.flags & kStoreBase) ? a.o0[slot] : ps.base_color.colors[0];
    p.c1 = (d.flags & kStoreEmi) ? a.o1[slot] : ps.emissive_color.colors[0];
    p.scale = (d.flags & kStoreScale) ? a.o2[slot] : p.iscale * ps.scale_curve.values[0]; // :602-605 with a constant curve
    return p;
}
__device__ __forceinline__ ParticleRegs make_particle(const DeviceTables &t, const fw_emission_settings &es, const DevParticleSettings &ps,
                                                      V3 origin_translation, Q4 origin_rotation,
                                                      V3 inherited_velocity, float modifier_scale, float modifier_speed,
                                                      uint32_t spawner_key, uint32_t emitter_local, uint64_t serial) {
    const uint2 key = make_uint2((uint32_t)t.seed, (uint32_t)(t.seed >> 32));
    const uint32_t c0 = (uint32_t)serial, c1 = (uint32_t)(serial >> 32), c2 = spawner_key;
    const uint4 r0 = philox4x32_10(make_uint4(c0, c1, c2, (emitter_local << 8) | 0u), key);
    const uint4 r1 = philox4x32_10(make_uint4(c0, c1, c2, (emitter_local << 8) | 1u), key);
    const uint4 r2 = philox4x32_10(make_uint4(c0, c1, c2, (emitter_local << 8) | 2u), key);
    const float u_shape0 = u01(r0.x), u_shape1 = u01(r0.y), u_shape2 = u01(r0.z);
    const float u_vel_angle = u01(r0.w), u_vel_radius = u01(r1.x), u_vel_mag = u01(r1.y);
    const float u_radial = u01(r1.z), u_scale = u01(r1.w), u_life = u01(r2.x);
    const float u_ang_angle = u01(r2.y), u_ang_radius = u01(r2.z), u_ang_mag = u01(r2.w);
    const float kPi = 3.14159265358979323846f;

    // EmissionShape::generate_point (src/emission_shape.rs:18-39)
    V3 spawn_offset = v3(0.0f, 0.0f, 0.0f);
    if (es.shape_kind == FW_SHAPE_SPHERE) {
        const float u = u_shape0 * 2.0f * kPi, v = u_shape1 * kPi, r = u_shape2;
        float su, cu, sv, cv;
        fw_sincosf(u, &su, &cu);
        fw_sincosf(v, &sv, &cv);
        const V3 unit = v3(sv * cu, cv, sv * su);
        spawn_offset = (unit * r) * es.shape_radius;
    } else if (es.shape_kind == FW_SHAPE_CIRCLE) {
        const float u = u_shape0 * 2.0f * kPi, r = u_shape1;
        const Q4 arc = q_from_rotation_arc(v3(0.0f, 1.0f, 0.0f), v3(es.shape_normal[0], es.shape_normal[1], es.shape_normal[2]));
        const Q4 q = qmul(arc, q_from_rotation_y(u));
        spawn_offset = qrot(q, v3(r * es.shape_radius, 0.0f, 0.0f));
    }
    // RandVec3::generate (bevy_utilitarian; definition in DESIGN.md section 4)
    auto rand_vec3 = [&](const fw_rand_vec3 &rv, float ua, float ur, float um) -> V3 {
        V3 dir = v3(rv.direction[0], rv.direction[1], rv.direction[2]);
        if (rv.spread > 0.0f) {
            const float a = ua * 2.0f * kPi;
            const float p = ur * rv.spread;
            float sp, cp, sa, ca;
            fw_sincosf(p, &sp, &cp);
            fw_sincosf(a, &sa, &ca);
            const V3 local = v3(sp * ca, cp, sp * sa);
            const Q4 arc = q_from_rotation_arc(v3(0.0f, 1.0f, 0.0f), normalize_or_zero(dir));
            dir = qrot(arc, local);
        }
        const float m = um * (rv.magnitude.max - rv.magnitude.min) + rv.magnitude.min;
        return dir * m;
    };
    const V3 iv = rand_vec3(es.initial_velocity, u_vel_angle, u_vel_radius, u_vel_mag);
    const float radial = u_radial * (es.initial_velocity_radial.max - es.initial_velocity_radial.min) + es.initial_velocity_radial.min;
    // src/core.rs:440-448 / :509-518
    V3 velocity = (qrot(origin_rotation, iv) + normalize_or_zero(spawn_offset) * radial) * modifier_speed;
    velocity = velocity + (es.inherit_parent_velocity ? inherited_velocity : v3(0.0f, 0.0f, 0.0f));
    const float initial_scale = (u_scale * (ps.initial_scale.max - ps.initial_scale.min) + ps.initial_scale.min) * modifier_scale;
    const V3 position = origin_translation + spawn_offset;
    const float lifetime = u_life * (ps.lifetime.max - ps.lifetime.min) + ps.lifetime.min;
    const V3 av = rand_vec3(es.initial_angular_velocity, u_ang_angle, u_ang_radius, u_ang_mag);

    ParticleRegs p;
    p.pos = position;
    p.age = 0.0f;
    p.rot = Q4{es.initial_rotation[0], es.initial_rotation[1], es.initial_rotation[2], es.initial_rotation[3]}; // :463
    p.vel = velocity;
    p.av = av;
    p.lifetime = lifetime;
    p.iscale = initial_scale;
    p.c0 = sample_gradient(ps.base_color, 0.0f);     // :460
    p.c1 = sample_gradient(ps.emissive_color, 0.0f); // :461
    p.scale = initial_scale;                         // scale = initial_scale (:457)
    return p;
}
__device__ __forceinline__ void emit_particle(const DeviceTables &t, const fw_emission_settings &es, const DevParticleSettings &ps,
                                              const StreamDesc &d, uint32_t slot, V3 origin_translation, Q4 origin_rotation,
                                              V3 inherited_velocity, float modifier_scale, float modifier_speed,
                                              uint32_t spawner_key, uint32_t emitter_local, uint64_t serial) {
    store_particle(d, slot, make_particle(t, es, ps, origin_translation, origin_rotation, inherited_velocity, modifier_scale,
                                          modifier_speed, spawner_key, emitter_local, serial), true);
}

// ------------------------------------------------------------------------------------------
// One particle, one frame: reference src/core.rs:591-658 in the reference's order. Returns
// whether the particle survives; p is updated in place (a particle that dies of age keeps its old
// state but the bumped age, :594-598; one destroyed by a collision has position / velocity / scale
// updated, :633-639).
// COLLIDE (0 none, 1 cuboid / sphere colliders, 2 + cylinders / cones): warp-synchronous (every
// lane of the warp must call; cand_queue = this thread's column of the CTA's candidate queue, see
// cast_ray). ROT = false: the stream is static (fw_internal.h) -- rotation and angular velocity
// are per-stream constants that :645-650 map to themselves, so they are not evaluated.
template <int COLLIDE, bool ROT>
__device__ __forceinline__ bool step_particle(const DeviceTables &t, const DevParticleSettings &ps, float dt, bool valid, ParticleRegs &p,
                                              bool &destroyed_by_collision, uint32_t *cand_queue = nullptr, bool sweeps = true) {
    const float age = p.age + dt;               // :594
    bool alive = valid && !(age >= p.lifetime); // :596-599
    p.age = age;
    destroyed_by_collision = false;
    V3 pos = p.pos, vel = p.vel;
    bool should_destroy = false;
    if (COLLIDE) // :608-617, hoisted out of the branch so that the warp stays converged inside
        particle_collision<(COLLIDE == 2)>(t.colliders, t.broadphase, ps.collision, alive && sweeps, pos, vel, dt, cand_queue, should_destroy);
    if (alive) {
        const float age_percent = age / p.lifetime;                     // :601
        p.scale = p.iscale * sample_curve(ps.scale_curve, age_percent); // :602-605
        if (!COLLIDE || !sweeps) pos = pos + vel * dt;                  // :619-623
        p.pos = pos;                                                    // :633
        if (should_destroy) {
            alive = false; // :636-639: position, velocity and scale are already updated
            destroyed_by_collision = true;
            p.vel = vel;
        } else {
            const V3 acc = v3(ps.acceleration[0], ps.acceleration[1], ps.acceleration[2]);
            p.vel = vel + (acc - vel * ps.linear_drag) * dt; // :641-643
            if (ROT) {
                p.rot = qmul(q_from_scaled_axis(p.av * dt), p.rot); // :645-647
                const V3 aacc = v3(ps.angular_acceleration[0], ps.angular_acceleration[1], ps.angular_acceleration[2]);
                p.av = p.av + (aacc - p.av * ps.angular_drag) * dt; // :648-650
            }
            p.c0 = sample_gradient(ps.base_color, age_percent);     // :652-653
            p.c1 = sample_gradient(ps.emissive_color, age_percent); // :654-655
        }
    }
    return alive;
}
// per-stream AABB of position -/+ scale (reference src/render.rs:681-692) for the spawn+step
// kernel, where the lanes of a warp may belong to different streams: reduce over the lanes of
// `group` (all of the same stream; every lane of the group must call)
__device__ __forceinline__ void accumulate_aabb(StreamState *stp, uint32_t group, bool alive, V3 pos, float scale) {
    uint32_t mn[3], mx[3];
    mn[0] = alive ? enc_f32(pos.x - scale) : 0xFFFFFFFFu;
    mn[1] = alive ? enc_f32(pos.y - scale) : 0xFFFFFFFFu;
    mn[2] = alive ? enc_f32(pos.z - scale) : 0xFFFFFFFFu;
    mx[0] = alive ? enc_f32(pos.x + scale) : 0u;
    mx[1] = alive ? enc_f32(pos.y + scale) : 0u;
    mx[2] = alive ? enc_f32(pos.z + scale) : 0u;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        mn[k] = __reduce_min_sync(group, mn[k]);
        mx[k] = __reduce_max_sync(group, mx[k]);
    }
    // zero = empty, so the minimum is kept as the max of the inverted encoding
    if ((group & ((1u << lane_id()) - 1u)) == 0u) { // the group's first lane
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (~mn[k] > stp->aabb_min_inv[k]) atomicMax(&stp->aabb_min_inv[k], ~mn[k]);
            if (mx[k] > stp->aabb_max[k]) atomicMax(&stp->aabb_max[k], mx[k]);
        }
    }
}

// spawn (Global emitters): one thread per new particle of the phase
__device__ __forceinline__ uint32_t find_cmd(const FrameDeviceInputs &f, uint32_t begin, uint32_t end, uint32_t g) {
    uint32_t lo = begin, hi = end; // last c with cmds[c].first <= g
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (f.cmds[mid].first <= g) lo = mid; else hi = mid;
    }
    return lo;
}
// the same by a whole warp, 32 probes per round (all 32 lanes must call; every lane gets the result)
__device__ __forceinline__ uint32_t find_cmd_warp(const FrameDeviceInputs &f, uint32_t begin, uint32_t end, uint32_t g) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t lo = begin, n = end - begin; // the answer is in [lo, lo + n), cmds[lo].first <= g
    while (n > 1u) {
        const uint32_t step = (n + 31u) >> 5;
        const uint32_t idx = lo + lane * step;
        const bool ok = lane == 0u || (idx < lo + n && f.cmds[idx].first <= g); // monotone in the lane
        const uint32_t k = 31u - (uint32_t)__clz((int)__ballot_sync(0xffffffffu, ok));
        const uint32_t hi = lo + n;
        lo += k * step;
        n = min(step, hi - lo);
    }
    return lo;
}
// grid-stride so the launch configuration is frame-independent (CUDA-graph friendly): the
// number of new particles is read from the frame header on the device
// 5 CTAs/SM: the C3 frame spawns 163 k particles = 1.08 waves at 4 CTAs/SM (ncu: 1.10 waves, the
// tail wave doubled the kernel time); at 5 CTAs/SM it is a single wave.
// STEP: the kernel also applies this frame's update (src/core.rs:591-658) to the particles it
// creates, so that it can run CONCURRENTLY with update_kernel, which then only touches older
// particles (FIFO streams without collision only; the host decides, header.step_in_spawn).
template <bool COLLIDE>
struct CandQueue { // broad-phase candidates of cast_ray, one column per thread
    uint32_t q[kCandQueue * kUpdateThreads];
};
template <>
struct CandQueue<false> {
    uint32_t q[1];
};
// COLLIDE (only with STEP): some of the streams sweep their particles against the colliders
// (src/core.rs:608-617), so the first step does too (warp-synchronous, see cast_ray).
template <bool STEP, int COLLIDE>
__global__ void __launch_bounds__(256, COLLIDE ? FW_MINB_COLLIDE : 5) spawn_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t phase) {
    static_assert(kUpdateThreads == 256, "the candidate queue is laid out for 256-thread CTAs");
    __shared__ uint32_t s_cmd;
    __shared__ uint32_t s_first[256];
    __shared__ CandQueue<(STEP && COLLIDE != 0)> cq;
    const PhaseInfo ph = f.header->phase[phase];
    const float dt = f.header->dt;
    const uint32_t n_slots = f.header->n_slots;
    for (uint32_t base = blockIdx.x * blockDim.x; base < ph.total_spawn; base += gridDim.x * blockDim.x) {
        // one binary search per CTA chunk, then the `first` of the (at most 256: every command has
        // count > 0) commands that begin inside the chunk go to shared memory, where every thread
        // finds its own. (A scene of many slow emitters has one command per particle: walking the
        // command list per thread was 90 us for 256 particles from 512 spawners.)
        if (threadIdx.x < 32u) { // warp 0: 32-ary search (2 rounds of loads for <= 1024 commands instead of 10)
            const uint32_t c0 = find_cmd_warp(f, ph.cmd_begin, ph.cmd_end, base);
            if (threadIdx.x == 0) s_cmd = c0;
        }
        __syncthreads();
        {
            const uint32_t c = s_cmd + threadIdx.x;
            s_first[threadIdx.x] = c < ph.cmd_end ? f.cmds[c].first : 0xFFFFFFFFu;
        }
        __syncthreads();
        const uint32_t g = base + threadIdx.x;
        const bool in_range = g < ph.total_spawn;
        bool have = false;
        uint32_t stream = 0xFFFFFFFFu, slot = 0;
        StreamDesc d{};
        ParticleRegs p{};
        if (in_range) {
            uint32_t lo = 0, hi = 256; // last i with s_first[i] <= g (s_first[0] <= base <= g)
            while (hi - lo > 1u) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_first[mid] <= g) lo = mid; else hi = mid;
            }
            const uint32_t c = s_cmd + lo;
            const SpawnCmd cmd = f.cmds[c];
            stream = cmd.stream;
            d = t.descs[stream];
            uint32_t head, spawn_base, count;
            if (f.header->derive) {
                const StreamState old = t.states_prev[stream];
                const Derived dv = derive_state(old, d, f.spawn_per_slot[stream]);
                head = dv.head;
                spawn_base = dv.c0;
                count = dv.n_update;
                if (STEP && cmd.dst_off == 0u && g == cmd.first) {
                    // a stream without update tiles this frame (no older particles): its first
                    // new particle publishes the stream state instead of update tile 0
                    const uint32_t *pv = f.host_tile_prefix + (size_t)d.variant * (n_slots + 1u);
                    if (pv[stream + 1u] == pv[stream]) {
                        StreamState *stp = &t.states[stream];
                        stp->head = dv.head;
                        stp->count = dv.n_update;
                        stp->spawn_base = dv.c0;
                        stp->overflow = old.overflow + dv.dropped;
                        atomicAdd(&t.plan->total_update, dv.n_update);
                        if (dv.dropped) atomicOr(&t.plan->error_flags, kErrOverflow);
                    }
                }
            } else {
                const StreamState st = t.states[stream];
                head = st.head;
                spawn_base = st.spawn_base;
                count = st.count;
            }
            const uint32_t logical = spawn_base + cmd.dst_off + (g - cmd.first);
            have = logical < count; // else dropped by the overflow clamp
            if (have) {
                slot = wrap(head + logical, d.capacity);
                const SpawnerInput in = f.inputs[cmd.input_idx];
                p = make_particle(t, t.emitters[cmd.emitter_idx], t.settings[stream],
                                  v3(in.translation[0], in.translation[1], in.translation[2]),
                                  Q4{in.rotation[0], in.rotation[1], in.rotation[2], in.rotation[3]},
                                  v3(in.parent_velocity[0], in.parent_velocity[1], in.parent_velocity[2]), in.modifier_scale,
                                  in.modifier_speed, cmd.spawner_key, cmd.emitter_local, cmd.serial_base + (g - cmd.first));
            }
        }
        if (!STEP) {
            if (have) store_particle(d, slot, p, true);
        } else {
            // first update step of the new particle, then one store of the final state. (Static
            // streams run the rotating body too: on their constants it is the identity, fw_internal.h.)
            bool by_collision;
            const DevParticleSettings &ps = t.settings[have ? stream : 0u];
            bool alive;
            if (COLLIDE) { // per lane: only the streams with collision settings sweep
                const bool sweeps = have && variant_collides(d.variant);
                alive = step_particle<COLLIDE, true>(t, ps, dt, have, p, by_collision, cq.q + threadIdx.x, sweeps);
            } else {
                alive = step_particle<0, true>(t, ps, dt, have, p, by_collision);
            }
            if (alive) store_particle(d, slot, p, true);
            // AABB and death count per stream: lanes of a warp may belong to different streams
            const uint32_t group = __match_any_sync(0xffffffffu, stream);
            if (stream != 0xFFFFFFFFu) {
                StreamState *stp = &t.states[stream];
                accumulate_aabb(stp, group, alive, p.pos, p.scale);
                const uint32_t dead_mask = __ballot_sync(group, have && !alive) & group;
                if (dead_mask && (group & ((1u << lane_id()) - 1u)) == 0u) atomicAdd(&stp->dead, __popc(dead_mask));
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Nested emission, reference src/core.rs:471-546.
__device__ __forceinline__ float div_euclid_f32(float a, float b) { // core::f32::div_euclid
    const float q = truncf(a / b);
    if (fmodf(a, b) < 0.0f) return b > 0.0f ? q - 1.0f : q + 1.0f;
    return q;
}
// step 1: compute_emission_count (:553-575) for every parent particle (:490-500)
__global__ void __launch_bounds__(256) nested_count_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t phase) {
    const PhaseInfo ph = f.header->phase[phase];
    const uint32_t ci = ph.nested_begin + blockIdx.y;
    if (ci >= ph.nested_end) return;
    const NestedCmd cmd = f.nested[ci];
    const fw_emission_settings &es = t.emitters[cmd.emitter_idx];
    const StreamDesc d = t.descs[cmd.parent_stream];
    const StreamState st = t.states[cmd.parent_stream];
    const StreamArrays a = stream_arrays(d.base, d.capacity);
    const DevParticleSettings &pps = t.settings[cmd.parent_stream];
    float *lea = lea_array(d.base, d.capacity, cmd.lea_index);
    uint32_t *counts = t.nested_scratch + cmd.scratch_off;
    const float start = es.offset_start, end = es.offset_end, per_cycle = es.count;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < st.count; j += gridDim.x * blockDim.x) {
        const uint32_t slot = wrap(st.head + j, d.capacity);
        const float age = a.m0[slot].w, lifetime = load_lifetime(d, a, pps, slot), last = lea[slot];
        const float percent_passed = age / lifetime;
        const float last_emission_percent = last / lifetime;
        const float lo = fmaxf(last_emission_percent, start);
        const float percent_passed_since_emission = fminf(percent_passed, end) - lo;
        const float percent_between_emissions = (end - start) / per_cycle;
        const float times = div_euclid_f32(percent_passed_since_emission, percent_between_emissions);
        uint32_t n = 0; // `as usize`: NaN / negatives -> 0, saturating
        if (times > 0.0f) n = times >= 4294967040.0f ? 0xFFFFFFFFu : (uint32_t)times;
        if (n > cmd.per_parent_cap) {
            n = cmd.per_parent_cap;
            atomicOr(&t.plan->error_flags, kErrNestedCap);
        }
        lea[slot] = (lo + times * percent_between_emissions) * lifetime; // :500
        counts[j] = n;
    }
}
// step 2: exclusive scan of the per-parent counts (parents emit in Vec order), then append the
// children to the child stream and reserve their serial numbers. One CTA per command.
__global__ void __launch_bounds__(1024) nested_scan_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t phase) {
    __shared__ uint32_t warp_sums[32];
    const PhaseInfo ph = f.header->phase[phase];
    const uint32_t ci = ph.nested_begin + blockIdx.x;
    if (ci >= ph.nested_end) return;
    const NestedCmd cmd = f.nested[ci];
    const uint32_t n_parents = t.states[cmd.parent_stream].count;
    uint32_t *counts = t.nested_scratch + cmd.scratch_off;
    uint32_t carry = 0;
    for (uint32_t chunk = 0; chunk < n_parents; chunk += blockDim.x) {
        const uint32_t j = chunk + threadIdx.x;
        const uint32_t v = j < n_parents ? counts[j] : 0u;
        uint32_t total;
        const uint32_t incl = block_inclusive_scan(v, warp_sums, total);
        if (j < n_parents) counts[j] = carry + incl - v;
        carry += total;
    }
    if (threadIdx.x == 0) {
        const StreamDesc cd = t.descs[cmd.child_stream];
        StreamState *cs = &t.states[cmd.child_stream];
        uint32_t total = carry;
        const uint32_t room = usable_capacity(cd) - cs->count;
        if (total > room) {
            cs->overflow += total - room;
            total = room;
            atomicOr(&t.plan->error_flags, kErrOverflow);
        }
        NestedOut o;
        o.total = total;
        o.spawn_base = cs->count;
        o.serial_base = t.nested_serial[cmd.emitter_idx];
        t.nested_out[ci] = o;
        cs->spawn_base = cs->count;
        cs->count += total;
        t.nested_serial[cmd.emitter_idx] = o.serial_base + carry;
    }
}
// step 3: the children (:506-544): origin = the parent particle's position / rotation / velocity
__global__ void __launch_bounds__(256) nested_spawn_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t phase) {
    const PhaseInfo ph = f.header->phase[phase];
    const uint32_t ci = ph.nested_begin + blockIdx.y;
    if (ci >= ph.nested_end) return;
    const NestedCmd cmd = f.nested[ci];
    const NestedOut out = t.nested_out[ci];
    if (out.total == 0u) return;
    const StreamDesc pd = t.descs[cmd.parent_stream], cd = t.descs[cmd.child_stream];
    const StreamState pst = t.states[cmd.parent_stream], cst = t.states[cmd.child_stream];
    // when parent and child are the same stream the parents are the first n_parents particles
    const uint32_t n_parents = cmd.parent_stream == cmd.child_stream ? out.spawn_base : pst.count;
    const uint32_t *offsets = t.nested_scratch + cmd.scratch_off;
    const StreamArrays pa = stream_arrays(pd.base, pd.capacity);
    const SpawnerInput in = f.inputs[cmd.input_idx];
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < out.total; g += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n_parents; // parent: last j with offsets[j] <= g
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if (offsets[mid] <= g) lo = mid; else hi = mid;
        }
        const uint32_t pslot = wrap(pst.head + lo, pd.capacity);
        const float4 p0 = pa.m0[pslot], p2 = pa.m2[pslot];
        emit_particle(t, t.emitters[cmd.emitter_idx], t.settings[cmd.child_stream], cd,
                      wrap(cst.head + out.spawn_base + g, cd.capacity), v3(p0.x, p0.y, p0.z),
                      load_rotation(pd, pa, t.settings[cmd.parent_stream], pslot),
                      v3(p2.x, p2.y, p2.z), in.modifier_scale, in.modifier_speed, cmd.spawner_key, cmd.emitter_local,
                      out.serial_base + g);
    }
}

// ------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA unit; SASS UBLKCP) for the per-stream settings block
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#ifndef FW_STATIC_UNROLL
#define FW_STATIC_UNROLL 1 // static update: tiles per loop trip (2: twice the loads in flight per thread)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// look-back status word of a tile: epoch<<34 | flag<<32 | value. The whole message is this one
// 64-bit word, so relaxed gpu-scope accesses are enough (as in CUB's single-word tile status): a
// release store would first drain the CTA's previous tile of particle stores, an acquire load
// invalidates L1 -- measured on C3r as 16 % + 38 % of the kernel (profiles/r1_tuning.md). No
// other data is ordered through it: survivors are written out of place, where no tile reads.
constexpr unsigned long long kFlagAgg = 1ull, kFlagPrefix = 2ull;
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct TileRef {
    uint32_t stream;
    uint32_t tile;     // tile index inside the stream
    uint32_t head;     // ring head for this frame
    uint32_t n_update; // particles of the stream entering this frame's update
};
// stream owning update tile `tile` of a variant: last s with prefix[s] <= tile (prefix[n] > tile)
__device__ __forceinline__ TileRef find_tile(const uint32_t *__restrict__ prefix, uint32_t n_slots, uint32_t tile) {
    uint32_t lo = 0, hi = n_slots;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (prefix[mid] <= tile) lo = mid; else hi = mid;
    }
    return TileRef{lo, tile - prefix[lo], 0u, 0u};
}
// find_tile by a whole warp: 32-ary search, one round of loads per factor 32 of the table (the
// binary search above is ~9 dependent loads for 512 streams: microseconds on the critical path of a
// CTA that starts a new stream segment). Every lane of the warp must call; all return the result.
__device__ __forceinline__ TileRef find_tile_warp(const uint32_t *__restrict__ prefix, uint32_t n_slots, uint32_t tile) {
    const uint32_t lane = lane_id();
    uint32_t lo = 0, n = n_slots; // answer in [lo, lo + n): the last s with prefix[s] <= tile
    while (n > 1u) {
        const uint32_t step = (n + 31u) / 32u;
        const uint32_t s = lo + lane * step;
        const bool le = lane * step < n && prefix[s] <= tile;
        const uint32_t m = __ballot_sync(0xffffffffu, le); // lanes 0..k (prefix is non-decreasing)
        const uint32_t k = m ? 31u - (uint32_t)__clz((int)m) : 0u;
        lo += k * step;
        n = min(step, n - k * step);
    }
    return TileRef{lo, tile - prefix[lo], 0u, 0u};
}
// the stream, ring head and particle count behind update tile `tile` (no side effects)
__device__ __forceinline__ TileRef resolve_tile(const DeviceTables &t, const FrameDeviceInputs &f, const uint32_t *prefix,
                                                uint32_t n_slots, uint32_t tile, bool derive) {
    TileRef r = find_tile(prefix, n_slots, tile);
    if (derive) {
        const Derived dv = derive_state(t.states_prev[r.stream], t.descs[r.stream], f.spawn_per_slot[r.stream]);
        r.head = dv.head;
        // concurrent spawn+step (header.step_in_spawn): this frame's new particles get their first
        // update inside the spawn kernel, the update kernel only covers the older ones
        r.n_update = f.header->step_in_spawn ? dv.c0 : dv.n_update;
    } else {
        r.head = t.states[r.stream].head;
        r.n_update = t.states[r.stream].count;
    }
    return r;
}
// thread 0 of a CTA: everything the CTA needs to know about an upcoming tile. On the derive
// path the first tile of a stream also publishes the stream's state for this frame.
__device__ __forceinline__ TileRef prepare_found_tile(const DeviceTables &t, const FrameDeviceInputs &f, TileRef r, bool derive);
__device__ __forceinline__ TileRef prepare_tile(const DeviceTables &t, const FrameDeviceInputs &f, const uint32_t *prefix,
                                                uint32_t n_slots, uint32_t tile, bool derive) {
    return prepare_found_tile(t, f, find_tile(prefix, n_slots, tile), derive);
}
__device__ __forceinline__ TileRef prepare_found_tile(const DeviceTables &t, const FrameDeviceInputs &f, TileRef r, bool derive) {
    StreamState *stp = &t.states[r.stream];
    if (derive) {
        const StreamState old = t.states_prev[r.stream];
        const Derived dv = derive_state(old, t.descs[r.stream], f.spawn_per_slot[r.stream]);
        r.head = dv.head;
        r.n_update = f.header->step_in_spawn ? dv.c0 : dv.n_update; // (see resolve_tile)
        if (r.tile == 0u) {
            stp->head = dv.head;
            stp->count = dv.n_update;
            stp->spawn_base = dv.c0;
            stp->overflow = old.overflow + dv.dropped;
            atomicAdd(&t.plan->total_update, dv.n_update);
            if (dv.dropped) atomicOr(&t.plan->error_flags, kErrOverflow);
        }
    } else {
        r.head = stp->head;
        r.n_update = stp->count;
    }
    return r;
}

// ------------------------------------------------------------------------------------------
// Compaction without collisions, pass 1 and 2. A particle dies this frame iff age + dt >=
// lifetime (src/core.rs:594-599), which only needs the `m0` and `k` packs: count_kernel writes
// every tile's death count and the running count in front of each of its 8 warps (one warp per
// tile, 24 B per particle), scan_kernel turns the tile counts of each stream into exclusive
// prefixes (one warp per stream) and publishes the stream's total.
// The update kernel then knows where its survivors go before it has loaded anything: no tile
// waits for another one (the one-pass look-back version was bound by exactly that wait: load ->
// aggregate -> poll per tile, profiles/r1_tuning.md section j).
__global__ void __launch_bounds__(256) count_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t variant) {
    const bool derive = f.header->derive != 0u;
    const uint32_t n_tiles = derive ? f.header->host_n_tiles[variant] : t.plan->n_tiles[variant];
    const uint32_t tile_base = derive ? f.header->host_tile_base[variant] : t.plan->tile_base[variant];
    const uint32_t n_slots = f.header->n_slots;
    const uint32_t *prefix = derive ? f.host_tile_prefix + (size_t)variant * (n_slots + 1u)
                                    : t.tile_prefix + (size_t)variant * (t.slots_cap + 1u);
    const float dt = f.header->dt;
    // every warp owns a contiguous share of the tiles: consecutive tiles belong to the same stream, so
    // the stream lookup (32-ary search by the whole warp) and its ring geometry are resolved once per
    // segment, and what is left per tile is eight independent loads, eight ballots and two stores
    const uint32_t lane = lane_id(), warps = (gridDim.x * blockDim.x) >> 5, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t per_warp = (n_tiles + warps - 1u) / warps;
    uint32_t tile = w * per_warp;
    const uint32_t tile_end = min(n_tiles, tile + per_warp);
    static_assert(kTile == 256, "the per-warp death counts of a tile are packed as 7 x 8 bits");
    while (tile < tile_end) {
        TileRef e = find_tile_warp(prefix, n_slots, tile);
        const StreamDesc d = t.descs[e.stream];
        if (derive) {
            const Derived dv = derive_state(t.states_prev[e.stream], d, f.spawn_per_slot[e.stream]);
            e.head = dv.head;
            e.n_update = f.header->step_in_spawn ? dv.c0 : dv.n_update; // (see resolve_tile)
        } else {
            e.head = t.states[e.stream].head;
            e.n_update = t.states[e.stream].count;
        }
        const uint32_t seg_end = min(tile_end, prefix[e.stream + 1u]);
        const StreamArrays a = stream_arrays(d.base, d.capacity);
        const uint32_t shift = e.head & 31u;
        // where (age, lifetime) live: rotating stream m0.w / k.x; static stream k = (lifetime, age)
        // -- 8 instead of 24 bytes per particle -- or, with a constant lifetime, m0.w alone
        const bool rot = variant_rotates(d.variant), klife = (d.flags & kStoreLife) != 0u;
        const float const_life = t.settings[e.stream].const_lifetime;
        for (uint32_t tis = e.tile; tile < seg_end; tile++, tis++) {
            const uint32_t tile_first = tis * kTile;
            // all eight rows of the tile are loaded before the first ballot: eight requests in flight per lane
            float age[kTile / 32], life[kTile / 32];
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)kTile / 32u; j++) {
                const uint32_t p = tile_first + j * 32u + lane, i = p - shift;
                const bool valid = p >= shift && i < e.n_update;
                age[j] = 0.0f;
                life[j] = __int_as_float(0x7f800000); // +inf: never dies
                if (valid) {
                    const uint32_t slot = wrap(e.head + i, d.capacity);
                    if (rot) {
                        age[j] = a.m0[slot].w;
                        life[j] = a.k[slot].x;
                    } else if (klife) {
                        const float2 K = a.k[slot];
                        age[j] = K.y;
                        life[j] = K.x;
                    } else {
                        age[j] = a.m0[slot].w;
                        life[j] = const_life;
                    }
                }
            }
            uint32_t dead = 0;
            unsigned long long before_warp = 0; // dead particles of the tile in front of warp j, 8 bits each (j = 1..7)
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)kTile / 32u; j++) {
                if (j) before_warp |= (unsigned long long)dead << (8u * (j - 1u));
                dead += __popc(__ballot_sync(0xffffffffu, age[j] + dt >= life[j])); // src/core.rs:594-599
            }
            if (lane == 0) {
                t.lookback[tile_base + tile] = dead;
                t.lookback[t.lookback_capacity + tile_base + tile] = before_warp;
            }
        }
    }
}
__global__ void __launch_bounds__(256) scan_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t variant) {
    const bool derive = f.header->derive != 0u;
    const uint32_t tile_base = derive ? f.header->host_tile_base[variant] : t.plan->tile_base[variant];
    const uint32_t n_slots = f.header->n_slots;
    const uint32_t *prefix = derive ? f.host_tile_prefix + (size_t)variant * (n_slots + 1u)
                                    : t.tile_prefix + (size_t)variant * (t.slots_cap + 1u);
    const uint32_t lane = lane_id(), warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slots; s += warps) {
        const uint32_t begin = prefix[s], end = prefix[s + 1u];
        if (begin == end) continue; // not a stream of this variant, or nothing to update
        uint32_t carry = 0;
        for (uint32_t k = begin; k < end; k += 32u) {
            unsigned long long *w = t.lookback + tile_base + k + lane;
            const uint32_t v = k + lane < end ? (uint32_t)*w : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += n;
            }
            if (k + lane < end) *w = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) t.states[s].dead = carry;
    }
}

struct alignas(16) UpdateSmem {
    DevParticleSettings settings[2];
    uint64_t bar[2];
    TileRef ref[2];
    uint32_t warp_alive[kUpdateThreads / 32];
    uint32_t team_cut[16];                       // first tile of every team's share (compact variants)
    uint32_t lb_sum[kUpdateThreads / 32];        // look-back partial sums, one per warp
    uint32_t lb_has_prefix[kUpdateThreads / 32]; // that warp's 32 predecessors include an inclusive prefix
};

// The fused per-frame update. One CTA processes whole tiles of 256 consecutive particles of one
// stream; persistent grid, tile = blockIdx.x + k*gridDim.x in increasing order (required by the
// look-back of the compact variants: a tile only ever waits on lower-numbered tiles, and every
// CTA of the grid is resident). Thread 0 looks up the next tile's stream and prefetches its
// settings block with a bulk async copy while the CTA works on the current tile.
// ROT: the streams of this launch keep rotation / angular velocity per particle (fw_internal.h).
template <bool COMPACT, int COLLIDE, bool ROT>
__global__ void __launch_bounds__(kUpdateThreads, COLLIDE ? FW_MINB_COLLIDE : (COMPACT ? FW_MINB_COMPACT : (ROT ? FW_MINB : FW_MINB_STATIC)))
    update_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t variant, uint32_t team_size) {
    __shared__ UpdateSmem sm;
    __shared__ CandQueue<(COLLIDE != 0)> cq;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool derive = f.header->derive != 0u;
    const uint32_t all_tiles = derive ? f.header->host_n_tiles[variant] : t.plan->n_tiles[variant];
    // Tile order. Independent tiles (FIFO; compaction with precounted deaths): CTA b takes tiles b,
    // b + grid, ... With collisions a compacting tile
    // waits for the aggregates of the preceding tiles of its stream, which couples the CTAs that
    // share a stream; with one global round robin the stream boundaries shift every round and the
    // whole grid falls into lock-step (everybody loads, then everybody computes, then everybody
    // stores: nothing overlaps, C3r 0.60 ms). So the grid is split into TEAMS of team_size CTAs
    // (one CTA slot of every SM), each team owns a contiguous share of the tiles and deals it
    // round robin among its members: teams only meet at the one stream that straddles a boundary,
    // drift apart in phase, and the CTAs resident on an SM are again in different phases.
    // Team shares start at stream boundaries (a stream that straddled two teams would make the
    // later team wait for the END of the earlier team's share); if that leaves the shares badly
    // unbalanced (few, huge streams) everybody falls back to the global round robin.
    const uint32_t n_slots = f.header->n_slots;
    const uint32_t *prefix = derive ? f.host_tile_prefix + (size_t)variant * (n_slots + 1u)
                                    : t.tile_prefix + (size_t)variant * (t.slots_cap + 1u);
    uint32_t tile_begin = blockIdx.x, tile_stride = gridDim.x, n_tiles = all_tiles;
    if (COMPACT && COLLIDE && team_size != 0u && gridDim.x >= 2u * team_size && all_tiles != 0u) {
        const uint32_t n_teams = min(gridDim.x / team_size, 15u);
        if (tid <= n_teams) {
            uint32_t cut = (uint32_t)((uint64_t)all_tiles * tid / n_teams);
            if (tid != 0u && tid != n_teams) cut = prefix[find_tile(prefix, n_slots, cut).stream]; // snap down to the stream's first tile
            sm.team_cut[tid] = cut;
        }
        __syncthreads();
        uint32_t biggest = 0;
        for (uint32_t k = 0; k < n_teams; k++) biggest = max(biggest, sm.team_cut[k + 1u] - sm.team_cut[k]);
        if (biggest <= 2u * (all_tiles / n_teams + 1u)) {
            const uint32_t team = blockIdx.x / team_size;
            if (team >= n_teams) return; // (grid not a multiple of the team size)
            n_tiles = sm.team_cut[team + 1u];
            tile_begin = sm.team_cut[team] + blockIdx.x % team_size;
            tile_stride = team_size;
        }
    }
    if (tile_begin >= n_tiles) return;
    const uint32_t tile_base = derive ? f.header->host_tile_base[variant] : t.plan->tile_base[variant];
    const float dt = f.header->dt;
    const uint32_t epoch = f.header->epoch;
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        const TileRef r = prepare_tile(t, f, prefix, n_slots, tile_begin, derive);
        sm.ref[0] = r;
        bulk_load(&sm.settings[0], &t.settings[r.stream], sizeof(DevParticleSettings), &sm.bar[0]);
    }
    __syncthreads();
    uint32_t it = 0;
#ifdef FW_DEBUG_TIMING
    long long dbg[6] = {0, 0, 0, 0, 0, 0}, dbg_t = clock64();
#define FW_DBG(k) { const long long now__ = clock64(); dbg[k] += now__ - dbg_t; dbg_t = now__; }
#else
#define FW_DBG(k)
#endif
    for (uint32_t tile = tile_begin; tile < n_tiles; tile += tile_stride, it++) {
        const uint32_t buf = it & 1u;
        const TileRef e = sm.ref[buf];
        const StreamDesc d = t.descs[e.stream];
        StreamState *stp = &t.states[e.stream];
        const uint32_t head = e.head, n_update = e.n_update;
        const StreamArrays a = stream_arrays(d.base, d.capacity);
        // FIFO rings: tiles are aligned to the ring's physical slots, not to the logical index --
        // the head moves by an arbitrary count every frame, and a warp whose 32 slots start at a
        // multiple of 32 touches 4 full 128-byte lines per float4 pack instead of straddling 5
        // (the first `shift` lanes of a stream's tile 0 idle).
        const uint32_t shift = head & 31u;
        const uint32_t tile_first = e.tile * kTile;
        const uint32_t i = tile_first + tid - shift;
        const bool valid = tile_first + tid >= shift && i < n_update;
        const uint32_t slot = wrap(head + (valid ? i : 0u), d.capacity);
        // compacting rings: logical index of the tile's first particle; the survivors go OUT OF
        // PLACE, behind the particles this frame reads (usable_capacity keeps the ring half empty)
        const uint32_t first_logical = tile_first > shift ? tile_first - shift : 0u;
        const uint32_t dst_base = wrap(head + n_update, d.capacity);
        // without collisions the death counts were taken by count_kernel / scan_kernel: dead
        // particles of the stream before this tile
        constexpr bool PRECOUNT = COMPACT && !COLLIDE;
        uint32_t pre_excl = 0;
        if (PRECOUNT) { // ... plus the dead ones of this tile in front of this warp
            const unsigned long long in_tile = t.lookback[t.lookback_capacity + tile_base + tile];
            pre_excl = (uint32_t)t.lookback[tile_base + tile] + (warp ? (uint32_t)(in_tile >> (8u * (warp - 1u))) & 255u : 0u);
        }

        // ---- loads: 64 B per particle of a rotating stream (five independent coalesced requests
        // per thread), 32 B of a static one (+ 8 B when its lifetime varies)
        float4 M0 = make_float4(0.f, 0.f, 0.f, 0.f), M1 = M0, M2 = M0;
        float2 M3 = make_float2(0.f, 0.f), K = make_float2(1.f, 0.f);
        const uint32_t flags = d.flags;
        if (valid) {
            M0 = ld_pack(a.m0 + slot);
            M2 = ld_pack(a.m2 + slot);
            if (ROT) {
                M1 = ld_pack(a.m1 + slot);
                M3 = ld_pack(a.m3 + slot);
                K = ld_pack(a.k + slot);
            } else if (flags & kStoreLife) {
                K = ld_pack(a.k + slot);
            }
        }
        // next tile: stream lookup + settings prefetch into the other buffer (every thread left
        // that buffer at the __syncthreads closing the previous iteration). ~2.5k cycles of
        // dependent loads for one thread: on thread 0, except where the tile's look-back aggregate
        // waits for warp 0 (compaction with collisions: the last warp does it, after the aggregate)
        constexpr bool LOOKBACK = COMPACT && COLLIDE;
        const uint32_t prep_tid = LOOKBACK ? kUpdateThreads - 32u : 0u;
        auto prepare_next = [&]() {
            if (tid == prep_tid && tile + tile_stride < n_tiles) {
                const TileRef r = prepare_tile(t, f, prefix, n_slots, tile + tile_stride, derive);
                sm.ref[buf ^ 1u] = r;
                bulk_load(&sm.settings[buf ^ 1u], &t.settings[r.stream], sizeof(DevParticleSettings), &sm.bar[buf ^ 1u]);
            }
        };
        if (!LOOKBACK) prepare_next();
        mbar_wait(&sm.bar[buf], (it >> 1) & 1u);
        const DevParticleSettings &ps = sm.settings[buf];

        // ---- reference src/core.rs:591-658
        ParticleRegs p;
        p.pos = v3(M0.x, M0.y, M0.z);
        p.age = M0.w;
        p.vel = v3(M2.x, M2.y, M2.z);
        if (ROT) {
            p.rot = Q4{M1.x, M1.y, M1.z, M1.w};
            p.av = v3(M2.w, M3.x, M3.y);
            p.lifetime = K.x;
            p.iscale = K.y;
        } else {
            p.rot = Q4{0.f, 0.f, 0.f, 1.f}; // (not used: per-stream constants)
            p.av = v3(0.f, 0.f, 0.f);
            p.iscale = M2.w;
            p.lifetime = (flags & kStoreLife) ? K.x : ps.const_lifetime;
        }
        p.c0 = make_float4(0.f, 0.f, 0.f, 0.f);
        p.c1 = p.c0;
        p.scale = 0.f;
        bool destroyed_by_collision;
        bool alive = step_particle<COLLIDE, ROT>(t, ps, dt, valid, p, destroyed_by_collision, COLLIDE ? cq.q + tid : nullptr);
        // destroyed-particle stream (:588,597,637): the record the handler receives keeps the old
        // colours (and the old scale unless a collision destroyed it). Capturing streams keep every
        // pack (fw_api.cu: stream_proofs).
        const bool capture = COMPACT && d.destroyed_base != nullptr && valid && !alive;
        if (capture) {
            p.c0 = a.o0[slot];
            p.c1 = a.o1[slot];
            if (!destroyed_by_collision) p.scale = a.o2[slot];
        }
        FW_DBG(0)
        const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
        const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);

        // ---- destination slot of a survivor
        uint32_t dslot = slot;
        if (COMPACT) {
            // LOOKBACK: rank of this particle among the survivors of its tile: warp ballot + popc,
            // then the warps' counts through shared memory. PRECOUNT knows the dead in front of
            // its warp already and needs neither shared memory nor a barrier.
            uint32_t before = 0, tile_alive = 0;
            if (LOOKBACK) {
                if (lane == 0) sm.warp_alive[warp] = __popc(alive_mask);
                __syncthreads();
#pragma unroll
                for (uint32_t w = 0; w < kUpdateThreads / 32; w++) {
                    const uint32_t n = sm.warp_alive[w];
                    if (w < warp) before += n;
                    tile_alive += n;
                }
            }
            FW_DBG(1)
            uint32_t excl = pre_excl; // dead particles of the stream before this tile (PRECOUNT: before this warp)
            if (LOOKBACK) {
                // destroy_on_collision: deaths are only known now, so the tiles of a stream chain
                // through a decoupled look-back. One status word per tile (epoch | flag | count).
                // (the host's tile table holds upper bounds: a tile may lie entirely past the count)
                const uint32_t tile_end = min(n_update, tile_first + (uint32_t)kTile - shift);
                const uint32_t tile_dead = (tile_end > first_logical ? tile_end - first_logical : 0u) - tile_alive;
                unsigned long long *status = t.lookback + tile_base + tile;
                const unsigned long long tag = (unsigned long long)epoch << 34;
                if (tid == 0 && e.tile != 0u) st_status(status, tag | (kFlagAgg << 32) | tile_dead);
                prepare_next();
                // 256 predecessors at a time, one THREAD per predecessor, so a stream of up to 256
                // tiles needs a single round of status loads (a serial walk by one thread cost
                // ~0.35 us per step; profiles/r1_tuning.md)
                excl = 0;
                if (e.tile != 0u) {
                    for (uint32_t nearest = e.tile - 1u;; nearest -= (uint32_t)kTile) {
                        const bool in_stream = tid <= nearest; // predecessor nearest - tid exists
                        unsigned long long w;
                        bool ready;
                        do {
                            w = in_stream ? ld_status(status - 1 - tid - (e.tile - 1u - nearest)) : (tag | (kFlagPrefix << 32));
                            ready = (w >> 34) == (unsigned long long)epoch && ((w >> 32) & 3ull) != 0ull;
                        } while (!__all_sync(0xffffffffu, ready));
                        FW_DBG(2)
                        const uint32_t prefix_lanes = __ballot_sync(0xffffffffu, ((w >> 32) & 3ull) == kFlagPrefix);
                        const uint32_t upto = prefix_lanes ? (uint32_t)__ffs((int)prefix_lanes) - 1u : 31u;
                        const uint32_t part = __reduce_add_sync(0xffffffffu, lane <= upto ? (uint32_t)w : 0u);
                        if (lane == 0) {
                            sm.lb_sum[warp] = part;
                            sm.lb_has_prefix[warp] = prefix_lanes != 0u;
                        }
                        __syncthreads();
                        bool found = false;
#pragma unroll
                        for (uint32_t w8 = 0; w8 < kUpdateThreads / 32; w8++) {
                            if (!found) {
                                excl += sm.lb_sum[w8];
                                found = sm.lb_has_prefix[w8] != 0u;
                            }
                        }
                        if (found) break;
                        __syncthreads(); // the warp slots are rewritten by the next round
                    }
                }
                if (tid == 0) {
                    st_status(status, tag | (kFlagPrefix << 32) | (excl + tile_dead));
                    if (tile_first + kTile - shift >= n_update) stp->dead = excl + tile_dead; // last tile of the stream
                }
            }
            FW_DBG(3)
            // dead particles of the stream in front of this one -> its rank among the survivors
            const uint32_t lanes_lt = (1u << lane) - 1u;
            const uint32_t dead_before = PRECOUNT ? excl + __popc(valid_mask & ~alive_mask & lanes_lt)
                                                  : excl + ((i - first_logical) - (before + __popc(alive_mask & lanes_lt)));
            dslot = wrap(dst_base + (i - dead_before), d.capacity);
            if (capture) { // destroyed particles, in Vec order, into the side block
                StreamDesc dd = d;
                dd.base = d.destroyed_base;
                dd.n_lea = 0u;
                store_particle(dd, dead_before, p, false);
            }
        } else {
            const uint32_t n_dead_w = __popc(valid_mask & ~alive_mask);
            if (lane == 0 && n_dead_w) atomicAdd(&stp->dead, n_dead_w);
        }

        // ---- stores: 92 B per survivor of a rotating stream, 28 B + the colours / scale that vary
        // of a static one (more when a compacting stream moves its constants)
        if (alive) {
            st_pack(a.m0 + dslot, make_float4(p.pos.x, p.pos.y, p.pos.z, p.age));
            st_pack(a.m2 + dslot, make_float4(p.vel.x, p.vel.y, p.vel.z, ROT ? p.av.x : p.iscale));
            if (ROT) {
                st_pack(a.m1 + dslot, make_float4(p.rot.x, p.rot.y, p.rot.z, p.rot.w));
                st_pack(a.m3 + dslot, make_float2(p.av.y, p.av.z));
                if (COMPACT) st_pack(a.k + dslot, K);
            } else if (flags & kStoreLife) {
                st_pack(a.k + dslot, make_float2(p.lifetime, p.age)); // (lifetime, copy of age): what count_kernel reads
            }
            if (COMPACT) // last_emitted_age moves with the particle (out of place: the source is still intact)
                for (uint32_t j = 0; j < d.n_lea; j++) lea_array(d.base, d.capacity, j)[dslot] = lea_array(d.base, d.capacity, j)[slot];
            if (flags & kStoreBase) st_pack(a.o0 + dslot, p.c0);
            if (flags & kStoreEmi) st_pack(a.o1 + dslot, p.c1);
            if (flags & kStoreScale) st_pack(a.o2 + dslot, p.scale);
        }

        // ---- per-stream AABB of position -/+ scale (reference src/render.rs:681-692), after the
        // stores so that nothing on the way to them waits for it. One axis per lane, bounds
        // pre-checked through L1 (stale but cheap; reading them from L2 first cost 10 %,
        // profiles/r1_tuning.md)
        {
            uint32_t mn[3], mx[3];
            mn[0] = alive ? enc_f32(p.pos.x - p.scale) : 0xFFFFFFFFu;
            mn[1] = alive ? enc_f32(p.pos.y - p.scale) : 0xFFFFFFFFu;
            mn[2] = alive ? enc_f32(p.pos.z - p.scale) : 0xFFFFFFFFu;
            mx[0] = alive ? enc_f32(p.pos.x + p.scale) : 0u;
            mx[1] = alive ? enc_f32(p.pos.y + p.scale) : 0u;
            mx[2] = alive ? enc_f32(p.pos.z + p.scale) : 0u;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
                mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
            }
            if (lane < 3u) {
                const uint32_t lo_inv = ~(lane == 0 ? mn[0] : (lane == 1 ? mn[1] : mn[2]));
                const uint32_t hi = lane == 0 ? mx[0] : (lane == 1 ? mx[1] : mx[2]);
                if (lo_inv > stp->aabb_min_inv[lane]) atomicMax(&stp->aabb_min_inv[lane], lo_inv);
                if (hi > stp->aabb_max[lane]) atomicMax(&stp->aabb_max[lane], hi);
            }
        }
        FW_DBG(4)
        __syncthreads(); // settings / tile-ref buffers are reused by the next iterations
        FW_DBG(5)
    }
#ifdef FW_DEBUG_TIMING
    if (COMPACT && (tid == 0 || tid == 200) && (blockIdx.x == 5 || blockIdx.x == 400) && f.header->epoch % 50u == 0u)
        printf("cta %u tid %u iters %u | loads+math %lld ranks %lld spin %lld lookback %lld stores+aabb %lld endbar %lld\n", blockIdx.x, tid, it,
               dbg[0] / it, dbg[1] / it, dbg[2] / it, dbg[3] / it, dbg[4] / it, dbg[5] / it);
#endif
}

// ------------------------------------------------------------------------------------------
// The fused per-frame update of STATIC streams without a collision sweep (FIFO rings and compaction
// with precounted deaths) -- C1..C4 of BASELINE.json. A static stream moves 32 B in and 48 B out per
// particle instead of 64 + 92, so what bounds the tile-scheduled update_kernel above (0.94 of the HBM
// peak on a rotating stream) is no longer memory: ncu on C3 showed 335 warp instructions per 32
// particles and the issue slots 62 % busy (profiles/r2/c_c3_static_tile_schedule_ncu.txt). Same
// per-particle arithmetic (src/core.rs:591-658 in the reference's order), another schedule:
//   * a CTA takes GROUPS of consecutive tiles, so consecutive tiles belong to the same stream: the
//     stream lookup, the ring geometry and all addressing are done once per stream segment, not
//     once per tile;
//   * there is no CTA-wide barrier and no shared memory at all: the 8 warps of a CTA free-run, each
//     looks its segment up itself and reads the per-stream settings through the read-only path;
//   * the per-stream AABB (src/render.rs:681-692) and the FIFO death count are accumulated in
//     registers over the whole segment and reduced (REDUX + atomics) once at its end;
//   * outputs the stream does not keep (constant gradients) are not evaluated; the knot interval of
//     the colour gradient is carried from tile to tile (ages along a ring change slowly) and only
//     re-searched when it no longer brackets the particle's age.
// FireworkGradient::sample_clamped as sample_gradient (fw_math.cuh), with the uneven core's knot
// search started from the interval the previous call found (same result: the first knot that is
// not < t is unique)
__device__ __forceinline__ float4 sample_gradient_hint(const DevGradient &g, float t, uint32_t &hint) {
    const float4 *colors = g.colors;
    if (g.kind == FW_CURVE_CONSTANT) return colors[0];
    t = clamp01(t);
    Interp it;
    if (g.kind == FW_CURVE_EVEN) {
        it = even_interp(g.n, t);
    } else {
        const uint32_t n = g.n;
        uint32_t idx = hint;
        const bool ok = (idx == 0u || g.times[idx - 1u] < t) && (idx >= n || !(g.times[idx] < t));
        if (!ok) {
            idx = 0;
            while (idx < n && g.times[idx] < t) idx++;
        }
        hint = idx;
        it = Interp{0u, 0u, 0.0f, false};
        if (idx < n && g.times[idx] == t) {
            it.lo = idx;
        } else if (idx == 0u) {
        } else if (idx >= n) {
            it.lo = n - 1u;
        } else {
            const float t_lower = g.times[idx - 1u], t_upper = g.times[idx];
            it.lo = idx - 1u;
            it.hi = idx;
            it.s = (t - t_lower) / (t_upper - t_lower);
            it.between = true;
        }
    }
    const float4 a = colors[it.lo];
    if (!it.between) return a;
    const float4 b = colors[it.hi];
    const float nf = 1.0f - it.s;
    return make_float4(a.x * nf + b.x * it.s, a.y * nf + b.y * it.s, a.z * nf + b.z * it.s, a.w * nf + b.w * it.s);
}
template <bool COMPACT>
__global__ void __launch_bounds__(kUpdateThreads, COMPACT ? FW_MINB_COMPACT : FW_MINB_STATIC)
    update_static_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t variant, uint32_t group_tiles) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool derive = f.header->derive != 0u;
    const uint32_t all_tiles = derive ? f.header->host_n_tiles[variant] : t.plan->n_tiles[variant];
    const uint32_t n_slots = f.header->n_slots;
    const uint32_t *prefix = derive ? f.host_tile_prefix + (size_t)variant * (n_slots + 1u)
                                    : t.tile_prefix + (size_t)variant * (t.slots_cap + 1u);
    // Tile order: groups of G consecutive tiles are dealt round robin to the CTAs, so that at any time
    // the grid sweeps one moving window of the particle arrays while a warp still stays inside one
    // stream for a whole group (never fewer groups than CTAs: a small scene spreads over the grid).
    // There is NO CTA-wide barrier and no shared memory: every warp looks its group's stream up itself
    // (32-ary search, two rounds of loads for 512 streams) and reads the per-stream settings through
    // the read-only path -- warp-uniform loads that hit in L1 after the first tile.
    const uint32_t share = max(1u, (all_tiles + gridDim.x - 1u) / gridDim.x);
    const uint32_t G = group_tiles ? min(group_tiles, share) : share;
    const uint32_t tile_base = derive ? f.header->host_tile_base[variant] : t.plan->tile_base[variant];
    const float dt = f.header->dt;
    const float kInf = __int_as_float(0x7f800000);
    for (uint64_t g0 = (uint64_t)blockIdx.x * G; g0 < all_tiles; g0 += (uint64_t)gridDim.x * G) {
        uint32_t tile = (uint32_t)g0;
        const uint32_t tile_end = (uint32_t)min((uint64_t)all_tiles, g0 + G);
        while (tile < tile_end) {
            // ---- a segment: the tiles [tile, seg_end) of one stream
            TileRef e = find_tile_warp(prefix, n_slots, tile);
            if (warp == 0u && lane == 0u) {
                e = prepare_found_tile(t, f, e, derive); // (publishes the stream's state with its tile 0)
            } else if (derive) {
                const Derived dv = derive_state(t.states_prev[e.stream], t.descs[e.stream], f.spawn_per_slot[e.stream]);
                e.head = dv.head;
                e.n_update = f.header->step_in_spawn ? dv.c0 : dv.n_update; // (see resolve_tile)
            }
            if (!derive) { // the plan kernel wrote this frame's state already
                e.head = t.states[e.stream].head;
                e.n_update = t.states[e.stream].count;
            }
            e.head = __shfl_sync(0xffffffffu, e.head, 0);
            e.n_update = __shfl_sync(0xffffffffu, e.n_update, 0);
            const uint32_t seg_end = min(tile_end, prefix[e.stream + 1u]);
            const StreamDesc d = t.descs[e.stream];
            StreamState *stp = &t.states[e.stream];
            // pack addresses from ONE base pointer: pack p of slot s sits at base + cap16 * (16-byte units of
            // the pack's offset) + s * element size (fw_internal.h) -- eight hoisted 64-bit pointers would
            // not fit the register budget of 5 CTAs per SM
            uint8_t *const base = d.base;
            const size_t cap16 = (size_t)d.capacity * 16u;
            const DevParticleSettings &ps = t.settings[e.stream];
            const uint32_t head = e.head, n_update = e.n_update, flags = d.flags, cap = d.capacity;
            // FIFO rings: tiles are aligned to the ring's physical slots, not to the logical index --
            // the head moves by an arbitrary count every frame, and a warp whose 32 slots start at a
            // multiple of 32 touches 4 full 128-byte lines per float4 pack instead of straddling 5
            // (the first `shift` lanes of a stream's tile 0 idle). q = offset from the aligned head:
            // a particle's logical index is q - shift, it exists iff shift <= q < q_end.
            const uint32_t shift = head & 31u, head_aligned = head - shift, q_end = n_update + shift;
            // compacting rings: the survivors go OUT OF PLACE, behind the particles this frame reads
            // (usable_capacity keeps the ring half empty); the death counts in front of every tile /
            // warp were taken by count_kernel / scan_kernel
            const uint32_t dst_base = wrap(head + n_update, cap);
            const bool klife = (flags & kStoreLife) != 0u;

            float mn0 = kInf, mn1 = kInf, mn2 = kInf, mx0 = -kInf, mx1 = -kInf, mx2 = -kInf; // AABB of this thread's survivors
            uint32_t n_alive = 0, n_dead = 0, hint = 0;
            uint32_t q = e.tile * (uint32_t)kTile + tid;
            struct TileIn { // one particle of one tile, as loaded
                bool valid;
                uint32_t slot, q, tile;
                float4 A, V;
                float2 K;
                unsigned long long pre_tile, pre_warps;
            };
            auto load_tile = [&](uint32_t tile_i, uint32_t qi) {
                TileIn in;
                in.q = qi;
                in.tile = tile_i;
                in.valid = qi >= shift && qi < q_end;
                in.slot = head_aligned + qi;
                if (in.slot >= cap) in.slot -= cap;
                // ---- loads: 32 B per particle (+ 8 B when the lifetime varies)
                in.A = make_float4(0.f, 0.f, 0.f, 0.f);
                in.V = in.A;
                in.K = make_float2(1.f, 0.f);
                const uint8_t *const row = base + (size_t)in.slot * 16u; // m0[slot]
                if (in.valid) {
                    in.A = ld_pack((const float4 *)row);
                    in.V = ld_pack((const float4 *)(row + 2u * cap16));
                    if (klife) in.K = ld_pack((const float2 *)(base + 3u * cap16 + cap16 / 2u) + in.slot);
                }
                in.pre_tile = 0;
                in.pre_warps = 0;
                if (COMPACT) {
                    in.pre_tile = t.lookback[tile_base + tile_i];
                    in.pre_warps = t.lookback[t.lookback_capacity + tile_base + tile_i];
                }
                return in;
            };
            auto process_tile = [&](const TileIn &in) {
                const bool valid = in.valid;
                const uint32_t slot = in.slot;
                // ---- reference src/core.rs:591-658 for a static stream (rotation and angular velocity
                // are per-stream constants that :645-650 map to themselves)
                const float lifetime = klife ? in.K.x : ps.const_lifetime, iscale = in.V.w;
                const float age = in.A.w + dt;                   // :594
                const bool alive = valid && !(age >= lifetime);  // :596-599
                uint32_t dslot = slot;
                if (COMPACT) {
                    const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
                    const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
                    const uint32_t excl = (uint32_t)in.pre_tile + (warp ? (uint32_t)(in.pre_warps >> (8u * (warp - 1u))) & 255u : 0u);
                    const uint32_t dead_before = excl + __popc(valid_mask & ~alive_mask & ((1u << lane) - 1u));
                    dslot = wrap(dst_base + ((in.q - shift) - dead_before), cap);
                } else {
                    n_dead += (valid && !alive) ? 1u : 0u;
                }
                if (alive) {
                    const float age_percent = age / lifetime;                                // :601
                    const float scale = iscale * sample_curve(ps.scale_curve, age_percent);  // :602-605
                    const V3 vel = v3(in.V.x, in.V.y, in.V.z);
                    const V3 pos = v3(in.A.x, in.A.y, in.A.z) + vel * dt;                    // :619-623
                    const V3 acc = v3(ps.acceleration[0], ps.acceleration[1], ps.acceleration[2]);
                    const V3 nvel = vel + (acc - vel * ps.linear_drag) * dt;                 // :641-643
                    // ---- stores: 28 B + the colours / scale that vary (+ 8 B when the lifetime does)
                    uint8_t *const drow = base + (size_t)dslot * 16u;
                    st_pack((float4 *)drow, make_float4(pos.x, pos.y, pos.z, age));
                    st_pack((float4 *)(drow + 2u * cap16), make_float4(nvel.x, nvel.y, nvel.z, iscale));
                    if (klife) st_pack((float2 *)(base + 3u * cap16 + cap16 / 2u) + dslot, make_float2(lifetime, age)); // (lifetime, copy of age): what count_kernel reads
                    if (COMPACT) // last_emitted_age moves with the particle (out of place: the source is still intact)
                        for (uint32_t j = 0; j < d.n_lea; j++) lea_array(d.base, cap, j)[dslot] = lea_array(d.base, cap, j)[slot];
                    if (flags & kStoreBase) st_pack((float4 *)(drow + 4u * cap16), sample_gradient_hint(ps.base_color, age_percent, hint)); // :652-653
                    if (flags & kStoreEmi) st_pack((float4 *)(drow + 5u * cap16), sample_gradient(ps.emissive_color, age_percent));         // :654-655
                    if (flags & kStoreScale) st_pack((float *)(base + 6u * cap16) + dslot, scale);
                    // AABB of position -/+ scale (reference src/render.rs:681-692; fminf / fmaxf like it)
                    mn0 = fminf(mn0, pos.x - scale);
                    mn1 = fminf(mn1, pos.y - scale);
                    mn2 = fminf(mn2, pos.z - scale);
                    mx0 = fmaxf(mx0, pos.x + scale);
                    mx1 = fmaxf(mx1, pos.y + scale);
                    mx2 = fmaxf(mx2, pos.z + scale);
                    n_alive++;
                }
            };
            while (tile < seg_end) {
                if (q - tid >= q_end) { // tiles past the stream's real count (the host's tile table holds upper bounds)
                    tile = seg_end;
                    break;
                }
#if FW_STATIC_UNROLL == 2
                // two tiles per trip: four LDG.128 in flight per thread before the first is used
                const bool two = tile + 1u < seg_end && q - tid + (uint32_t)kTile < q_end;
                const TileIn a = load_tile(tile, q);
                TileIn b = a;
                if (two) b = load_tile(tile + 1u, q + (uint32_t)kTile);
                process_tile(a);
                if (two) process_tile(b);
                tile += two ? 2u : 1u;
                q += two ? 2u * (uint32_t)kTile : (uint32_t)kTile;
#else
                process_tile(load_tile(tile, q));
                tile++;
                q += (uint32_t)kTile;
#endif
            }
            // ---- end of the segment: one reduction per warp
            const uint32_t warp_alive = __reduce_add_sync(0xffffffffu, n_alive);
            if (!COMPACT) {
                const uint32_t warp_dead = __reduce_add_sync(0xffffffffu, n_dead);
                if (lane == 0 && warp_dead) atomicAdd(&stp->dead, warp_dead);
            }
            if (warp_alive) {
                uint32_t lo[3], hi[3];
                lo[0] = __reduce_min_sync(0xffffffffu, enc_f32(mn0));
                lo[1] = __reduce_min_sync(0xffffffffu, enc_f32(mn1));
                lo[2] = __reduce_min_sync(0xffffffffu, enc_f32(mn2));
                hi[0] = __reduce_max_sync(0xffffffffu, enc_f32(mx0));
                hi[1] = __reduce_max_sync(0xffffffffu, enc_f32(mx1));
                hi[2] = __reduce_max_sync(0xffffffffu, enc_f32(mx2));
                if (lane < 3u) { // zero = empty, so the minimum is kept as the max of the inverted encoding
                    const uint32_t lo_inv = ~(lane == 0 ? lo[0] : (lane == 1 ? lo[1] : lo[2]));
                    const uint32_t h = lane == 0 ? hi[0] : (lane == 1 ? hi[1] : hi[2]);
                    if (lo_inv > stp->aabb_min_inv[lane]) atomicMax(&stp->aabb_min_inv[lane], lo_inv);
                    if (h > stp->aabb_max[lane]) atomicMax(&stp->aabb_max[lane], h);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// pack: live ParticleInstance rows of the streams [slot_begin, slot_end), creation order, Vec
// order inside a stream. out[0] = total rows, out[1 + k] = first row of stream slot_begin + k.
// slot_list != nullptr: the streams slot_list[slot_begin .. slot_end) instead of the slots themselves
// (render extract of a subset of the spawners)
__global__ void pack_prefix_kernel(DeviceTables t, const uint32_t *slot_list, uint32_t slot_begin, uint32_t slot_end, unsigned long long *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long acc = 0;
        for (uint32_t k = slot_begin; k < slot_end; k++) {
            const uint32_t s = slot_list ? slot_list[k] : k;
            out[1 + (k - slot_begin)] = acc;
            if (t.descs[s].capacity) acc += t.states[s].count - t.states[s].dead;
        }
        out[0] = acc;
    }
}
// PackDst: where the rows go -- one caller buffer, or (multi-GPU render extract) this rank's
// region of the gather buffer of EVERY rank: peer buffers are mapped over NVLink, so the all-gather
// is just these stores (one HBM read, n_dst coalesced 16-byte writes per chunk).
__global__ void __launch_bounds__(256) pack_copy_kernel(DeviceTables t, const uint32_t *slot_list, uint32_t slot_begin, uint32_t slot_end,
                                                        const unsigned long long *offsets, PackDst dst, uint64_t cap_rows) {
    if (slot_begin + blockIdx.y >= slot_end) return;
    const uint32_t s = slot_list ? slot_list[slot_begin + blockIdx.y] : slot_begin + blockIdx.y;
    const StreamDesc d = t.descs[s];
    if (d.capacity == 0u) return;
    const StreamState st = t.states[s];
    const uint32_t live = st.count - st.dead;
    const uint32_t first = live_first(d, st);
    const unsigned long long off = offsets[1 + blockIdx.y];
    const StreamArrays a = stream_arrays(d.base, d.capacity);
    const DevParticleSettings &ps = t.settings[s];
    const bool rot = variant_rotates(d.variant);
    // one 16-byte chunk of a row per thread: fully coalesced 64-byte row writes; what the stream
    // does not keep per particle comes from its constants (fw_internal.h)
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < (uint64_t)live * 4u; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(q >> 2), c = (uint32_t)(q & 3u);
        if (off + r >= cap_rows) return;
        const uint32_t slot = wrap(first + r, d.capacity);
        float4 v;
        if (c == 0u) { // position.xyz, scale
            v = a.m0[slot];
            if (d.flags & kStoreScale) v.w = a.o2[slot];
            else v.w = (rot ? a.k[slot].y : a.m2[slot].w) * ps.scale_curve.values[0];
        } else if (c == 1u) {
            v = rot ? a.m1[slot] : make_float4(ps.const_rotation[0], ps.const_rotation[1], ps.const_rotation[2], ps.const_rotation[3]);
        } else if (c == 2u) {
            v = (d.flags & kStoreBase) ? a.o0[slot] : ps.base_color.colors[0];
        } else {
            v = (d.flags & kStoreEmi) ? a.o1[slot] : ps.emissive_color.colors[0];
        }
        for (uint32_t k = 0; k < dst.n; k++) dst.rows[k][(off + r) * 4u + c] = v;
    }
}

// Device-side barrier between the ranks of a gather (one warp, lane r <-> rank r): publish
// `epoch` in slot [which][my_rank] of every rank's header (with this rank's row count when
// which == kGatherDone), then wait until every rank has published it in OUR header. System-scope
// release / acquire; the pack kernel that precedes a kGatherDone signal in stream order has
// completed, the fence makes its peer stores visible before the flag. A rank that never shows
// up is a timeout (error word set), never a hang.
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void __launch_bounds__(32) gather_signal_kernel(GatherPeers p, uint32_t which, unsigned long long epoch,
                                                           const unsigned long long *rows_src, unsigned long long timeout_ns) {
    const uint32_t r = threadIdx.x;
    if (r >= p.n_ranks) return;
    GatherHeader *peer = (GatherHeader *)p.base[r];
    GatherHeader *mine = (GatherHeader *)p.base[p.my_rank];
    if (which == kGatherDone) {
        unsigned long long n = rows_src[0];
        if (n > p.cap_rows_per_rank) n = p.cap_rows_per_rank; // the pack kernel dropped the rest
        *(volatile unsigned long long *)&peer->rows[p.my_rank] = n;
    }
    __threadfence_system();
    st_release_sys(which == kGatherDone ? &peer->done[p.my_rank] : &peer->ready[p.my_rank], epoch);
    const unsigned long long *flag = which == kGatherDone ? &mine->done[r] : &mine->ready[r];
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < epoch) {
        if (global_timer_ns() - t0 > timeout_ns) {
            *(volatile unsigned long long *)&mine->error = 1ull + r;
            break;
        }
        __nanosleep(256);
    }
}

// ParticleData rows of one block (host mirror, destroyed stream, tests)
__global__ void __launch_bounds__(256) gather_particles_kernel(StreamDesc d, const DevParticleSettings *psp, uint32_t first, uint32_t n,
                                                               uint32_t pbr, fw_particle_data *dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ParticleRegs r = load_particle(d, *psp, wrap(first + i, d.capacity));
    fw_particle_data p;
    p.position[0] = r.pos.x; p.position[1] = r.pos.y; p.position[2] = r.pos.z;
    p.velocity[0] = r.vel.x; p.velocity[1] = r.vel.y; p.velocity[2] = r.vel.z;
    p.rotation[0] = r.rot.x; p.rotation[1] = r.rot.y; p.rotation[2] = r.rot.z; p.rotation[3] = r.rot.w;
    p.angular_velocity[0] = r.av.x; p.angular_velocity[1] = r.av.y; p.angular_velocity[2] = r.av.z;
    p.initial_scale = r.iscale;
    p.scale = r.scale;
    p.age = r.age;
    p.lifetime = r.lifetime;
    p.base_color[0] = r.c0.x; p.base_color[1] = r.c0.y; p.base_color[2] = r.c0.z; p.base_color[3] = r.c0.w;
    p.emissive_color[0] = r.c1.x; p.emissive_color[1] = r.c1.y; p.emissive_color[2] = r.c1.z; p.emissive_color[3] = r.c1.w;
    p.pbr = pbr;
    dst[i] = p;
}
__global__ void __launch_bounds__(256) scatter_particles_kernel(StreamDesc d, uint32_t n, const fw_particle_data *src) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fw_particle_data s = src[i];
    ParticleRegs p;
    p.pos = v3(s.position[0], s.position[1], s.position[2]);
    p.vel = v3(s.velocity[0], s.velocity[1], s.velocity[2]);
    p.rot = Q4{s.rotation[0], s.rotation[1], s.rotation[2], s.rotation[3]};
    p.av = v3(s.angular_velocity[0], s.angular_velocity[1], s.angular_velocity[2]);
    p.age = s.age;
    p.lifetime = s.lifetime;
    p.iscale = s.initial_scale;
    p.scale = s.scale;
    p.c0 = make_float4(s.base_color[0], s.base_color[1], s.base_color[2], s.base_color[3]);
    p.c1 = make_float4(s.emissive_color[0], s.emissive_color[1], s.emissive_color[2], s.emissive_color[3]);
    store_particle(d, i, p, true); // (what the stream does not keep was checked against its constants by the host)
}
__global__ void __launch_bounds__(256) sincos_kernel(const float *x, uint64_t n, float *s, float *c) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        fw_sincosf(x[i], &s[i], &c[i]);
}
__global__ void __launch_bounds__(256) ring_copy_kernel(StreamDesc src, uint32_t first, uint32_t n, StreamDesc dst) {
    const StreamArrays a = stream_arrays(src.base, src.capacity), b = stream_arrays(dst.base, dst.capacity);
    // only the packs this stream keeps (store_particle): the others were never written
    const bool rot = variant_rotates(src.variant);
    const uint32_t flags = src.flags;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = wrap(first + i, src.capacity);
        b.m0[i] = a.m0[s];
        b.m2[i] = a.m2[s];
        if (rot) {
            b.m1[i] = a.m1[s];
            b.m3[i] = a.m3[s];
        }
        if (rot || (flags & kStoreLife)) b.k[i] = a.k[s];
        if (flags & kStoreBase) b.o0[i] = a.o0[s];
        if (flags & kStoreEmi) b.o1[i] = a.o1[s];
        if (flags & kStoreScale) b.o2[i] = a.o2[s];
        for (uint32_t j = 0; j < src.n_lea; j++) lea_array(dst.base, dst.capacity, j)[i] = lea_array(src.base, src.capacity, j)[s];
    }
}

// ------------------------------------------------------------------------------------------
cudaError_t launch_plan(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant_mask, uint32_t what, uint32_t phase, cudaStream_t s) {
    plan_kernel<<<1, 1024, 0, s>>>(t, f, variant_mask, what, phase);
    return cudaGetLastError();
}
cudaError_t launch_spawn(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t phase, uint32_t total_spawn, bool step, int collide,
                         cudaStream_t s) {
    if (total_spawn == 0) return cudaSuccess;
    const uint32_t fixed = 148u * 5u; // one resident wave; larger counts stride
    const uint32_t blocks = total_spawn == 0xFFFFFFFFu ? fixed : std::min(fixed, (total_spawn + 255u) / 256u);
    if (step && collide == 2) spawn_kernel<true, 2><<<blocks, 256, 0, s>>>(t, f, phase);
    else if (step && collide == 1) spawn_kernel<true, 1><<<blocks, 256, 0, s>>>(t, f, phase);
    else if (step) spawn_kernel<true, 0><<<blocks, 256, 0, s>>>(t, f, phase);
    else spawn_kernel<false, 0><<<blocks, 256, 0, s>>>(t, f, phase);
    return cudaGetLastError();
}
cudaError_t launch_nested(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t phase, uint32_t n_cmds, cudaStream_t s) {
    if (n_cmds == 0) return cudaSuccess;
    if (n_cmds > 65535u) return cudaErrorInvalidValue;
    const uint32_t gx = std::max(1u, std::min(148u * 4u, (148u * 8u) / n_cmds));
    nested_count_kernel<<<dim3(gx, n_cmds), 256, 0, s>>>(t, f, phase);
    nested_scan_kernel<<<n_cmds, 1024, 0, s>>>(t, f, phase);
    nested_spawn_kernel<<<dim3(gx, n_cmds), 256, 0, s>>>(t, f, phase);
    return cudaGetLastError();
}
#ifndef FW_GROUP_TILES
// 0 = one contiguous share of the tiles per CTA. Measured on C3 (10 M particles, 80 B each, update kernel
// alone): contiguous 0.145 ms; groups of 32 / 16 / 8 / 4 / 2 tiles dealt round robin 0.157 / 0.159 / 0.161 /
// 0.163 / 0.186 ms, odd group sizes (3 .. 27) the same -- the moving window buys nothing, the per-group
// lookups cost (profiles/r2/tuning.md)
#define FW_GROUP_TILES 0
#endif
// tiles per group of the static update (see update_static_kernel); FW_GROUP_TILES in the environment
// overrides the built-in value (tuning runs)
static uint32_t group_tiles_setting() {
    static const uint32_t v = [] {
        const char *e = getenv("FW_GROUP_TILES");
        return e ? (uint32_t)strtoul(e, nullptr, 10) : (uint32_t)FW_GROUP_TILES;
    }();
    return v;
}
template <bool COMPACT, bool ROT>
static void launch_update_t(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant, int grid, uint32_t ts, int collide,
                            cudaStream_t s) {
    if (collide == 2) update_kernel<COMPACT, 2, ROT><<<grid, kUpdateThreads, 0, s>>>(t, f, variant, ts);
    else if (collide == 1) update_kernel<COMPACT, 1, ROT><<<grid, kUpdateThreads, 0, s>>>(t, f, variant, ts);
    else if constexpr (ROT) update_kernel<COMPACT, 0, true><<<grid, kUpdateThreads, 0, s>>>(t, f, variant, ts);
    else update_static_kernel<COMPACT><<<grid, kUpdateThreads, 0, s>>>(t, f, variant, group_tiles_setting()); // static, no sweep: C1..C4
}
cudaError_t launch_update(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant, int grid, int team_size, bool revolved,
                          cudaStream_t s) {
    if (variant >= kNumVariants) return cudaErrorInvalidValue;
    const uint32_t ts = (uint32_t)team_size;
    const int collide = variant_collides(variant) ? (revolved ? 2 : 1) : 0;
    if (variant_rotates(variant)) {
        if (variant_is_fifo(variant)) launch_update_t<false, true>(t, f, variant, grid, ts, collide, s);
        else launch_update_t<true, true>(t, f, variant, grid, ts, collide, s);
    } else {
        if (variant_is_fifo(variant)) launch_update_t<false, false>(t, f, variant, grid, ts, collide, s);
        else launch_update_t<true, false>(t, f, variant, grid, ts, collide, s);
    }
    return cudaGetLastError();
}
cudaError_t launch_count_scan(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant, uint32_t n_slots, cudaStream_t s) {
    count_kernel<<<148 * 6, 256, 0, s>>>(t, f, variant); // one resident wave (38 registers: 6 CTAs per SM); a contiguous share of the tiles per warp
    scan_kernel<<<std::max(1u, std::min(148u * 8u, (n_slots + 7u) / 8u)), 256, 0, s>>>(t, f, variant); // one warp per stream
    return cudaGetLastError();
}
cudaError_t update_grid_size(int device, int *grids, int *team_size) {
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    int occ[kNumVariants] = {0};
    // (collision variants: the build with the cylinder / cone tests is the larger one)
#define FW_OCC(v, K) \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[v], K, kUpdateThreads, 0); \
    if (e != cudaSuccess) return e;
    FW_OCC(kFifo, (update_static_kernel<false>))
    FW_OCC(kCompact, (update_static_kernel<true>))
    FW_OCC(kFifoCollide, (update_kernel<false, 2, false>))
    FW_OCC(kCompactCollide, (update_kernel<true, 2, false>))
    FW_OCC(kFifo | kVarRot, (update_kernel<false, 0, true>))
    FW_OCC(kCompact | kVarRot, (update_kernel<true, 0, true>))
    FW_OCC(kFifoCollide | kVarRot, (update_kernel<false, 2, true>))
    FW_OCC(kCompactCollide | kVarRot, (update_kernel<true, 2, true>))
#undef FW_OCC
    // persistent grids: every CTA must be resident (the look-back of the compact variants
    // spins on lower-numbered tiles)
    for (int v = 0; v < (int)kNumVariants; v++) grids[v] = sms * (occ[v] > 0 ? occ[v] : 1);
    *team_size = sms; // one CTA slot of every SM (see update_kernel)
    return cudaSuccess;
}
static cudaError_t launch_pack_any(const DeviceTables &t, const uint32_t *slot_list, uint32_t slot_begin, uint32_t slot_end, const PackDst &dst,
                                   uint64_t cap_rows, unsigned long long *out, cudaStream_t s) {
    pack_prefix_kernel<<<1, 32, 0, s>>>(t, slot_list, slot_begin, slot_end, out);
    if (slot_end > slot_begin) {
        const uint32_t n = slot_end - slot_begin;
        for (uint32_t y0 = 0; y0 < n; y0 += 32768u) { // gridDim.y limit is 65535
            dim3 grid(n == 1 ? 1184 : 16, std::min(32768u, n - y0));
            pack_copy_kernel<<<grid, 256, 0, s>>>(t, slot_list, slot_begin + y0, slot_end, out + y0, dst, cap_rows);
        }
    }
    return cudaGetLastError();
}
cudaError_t launch_pack_instances(const DeviceTables &t, uint32_t slot_begin, uint32_t slot_end, const PackDst &dst,
                                  uint64_t cap_rows, unsigned long long *out, cudaStream_t s) {
    return launch_pack_any(t, nullptr, slot_begin, slot_end, dst, cap_rows, out, s);
}
cudaError_t launch_pack_listed(const DeviceTables &t, const uint32_t *d_slot_list, uint32_t n, float4 *dst, uint64_t cap_rows,
                               unsigned long long *out, cudaStream_t s) {
    PackDst d{};
    d.rows[0] = dst;
    d.n = 1;
    return launch_pack_any(t, d_slot_list, 0, n, d, cap_rows, out, s);
}
cudaError_t launch_pack_instances(const DeviceTables &t, uint32_t slot_begin, uint32_t slot_end, float4 *dst,
                                  uint64_t cap_rows, unsigned long long *out, cudaStream_t s) {
    PackDst d{};
    d.rows[0] = dst;
    d.n = 1;
    return launch_pack_instances(t, slot_begin, slot_end, d, cap_rows, out, s);
}
cudaError_t launch_gather_signal(const GatherPeers &p, uint32_t which, unsigned long long epoch, const unsigned long long *rows_src,
                                 unsigned long long timeout_ns, cudaStream_t s) {
    gather_signal_kernel<<<1, 32, 0, s>>>(p, which, epoch, rows_src, timeout_ns);
    return cudaGetLastError();
}
cudaError_t launch_gather_particles(const StreamDesc &d, const DevParticleSettings *ps, uint32_t first, uint32_t n, uint32_t pbr,
                                    fw_particle_data *dst, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    gather_particles_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(d, ps, first, n, pbr, dst);
    return cudaGetLastError();
}
cudaError_t launch_scatter_particles(const StreamDesc &d, uint32_t n, const fw_particle_data *src, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    scatter_particles_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(d, n, src);
    return cudaGetLastError();
}
cudaError_t launch_sincos(const float *x, uint64_t n, float *s, float *c, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sincos_kernel<<<(unsigned)std::min<uint64_t>((n + 255u) / 256u, 148u * 8u), 256, 0, st>>>(x, n, s, c);
    return cudaGetLastError();
}
cudaError_t launch_ring_copy(const StreamDesc &src, uint32_t first, uint32_t n, const StreamDesc &dst, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint32_t blocks = std::min<uint32_t>((n + 255u) / 256u, 148u * 8u);
    ring_copy_kernel<<<blocks, 256, 0, s>>>(src, first, n, dst);
    return cudaGetLastError();
}

} // namespace fw

#ifdef FW_COLLIDE_STATS
extern "C" int fw_debug_collide_stats(unsigned long long *out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, fw::g_collide_stats, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(fw::g_collide_stats, z, sizeof z); }
    return 0;
}
#endif
