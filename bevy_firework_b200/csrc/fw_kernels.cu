// fw_kernels.cu -- sm_100a kernels of the particle path.
//
//   plan_kernel    applies last frame's deaths to every ring (head/count), appends this
//                  frame's spawn counts, and builds the per-variant tile tables
//   spawn_kernel   reference src/core.rs:437-469 + src/emission_shape.rs:18-39, one thread per
//                  new particle, Philox4x32-10 counter-based draws
//   update_kernel  reference src/core.rs:591-658 (+ :744-800 when COLLIDE), fused with the
//                  ParticleInstance row conversion (src/render.rs:105-115), death handling
//                  and the per-stream AABB reduction (src/render.rs:677-703)
//   pack_kernel    gathers the live instance rows of all streams into one contiguous buffer
//
// Built with -fmad=false (see fw_math.cuh).
#include "fw_math.cuh"

namespace fw {

__device__ __forceinline__ uint32_t wrap(uint32_t x, uint32_t cap) { return x >= cap ? x - cap : x; }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

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

constexpr uint32_t kErrOverflow = 1u;
constexpr uint32_t kErrTileTable = 2u;

__global__ void __launch_bounds__(1024) plan_kernel(DeviceTables t, FrameDeviceInputs f) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t n_slots = f.header->n_slots;
    uint32_t my_total = 0;
    for (uint32_t s = threadIdx.x; s < n_slots; s += blockDim.x) {
        const StreamDesc d = t.descs[s];
        if (d.capacity == 0u) continue;
        StreamState st = t.states[s];
        const bool fifo = (d.variant == kFifo || d.variant == kFifoCollide);
        if (fifo) st.head = wrap(st.head + st.dead, d.capacity);
        st.count -= st.dead;
        st.dead = 0u;
        uint32_t spawn = f.spawn_per_slot[s];
        const uint32_t room = d.capacity - st.count;
        if (spawn > room) {
            st.overflow += spawn - room;
            spawn = room;
            atomicOr(&t.plan->error_flags, kErrOverflow);
        }
        st.spawn_base = st.count;
        st.count += spawn;
        st.aabb_min[0] = st.aabb_min[1] = st.aabb_min[2] = 0xFFFFFFFFu;
        st.aabb_max[0] = st.aabb_max[1] = st.aabb_max[2] = 0u;
        t.states[s] = st;
        my_total += st.count;
    }
    __syncthreads();
    uint32_t base_total = 0;
    for (uint32_t v = 0; v < kNumVariants; v++) {
        uint32_t carry = 0;
        for (uint32_t chunk = 0; chunk < n_slots; chunk += blockDim.x) {
            const uint32_t s = chunk + threadIdx.x;
            uint32_t tiles = 0;
            if (s < n_slots) {
                const StreamDesc d = t.descs[s];
                if (d.capacity != 0u && d.variant == v) tiles = (t.states[s].count + kTile - 1u) / kTile;
            }
            uint32_t total;
            const uint32_t incl = block_inclusive_scan(tiles, warp_sums, total);
            uint32_t at = base_total + carry + incl - tiles;
            if (at + tiles > t.tiles_capacity) {
                if (tiles) atomicOr(&t.plan->error_flags, kErrTileTable);
            } else {
                for (uint32_t j = 0; j < tiles; j++) t.tiles[at + j] = TileEntry{s, j};
            }
            carry += total;
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
// spawn: one thread per new particle
__global__ void __launch_bounds__(256) spawn_kernel(DeviceTables t, FrameDeviceInputs f) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = f.header->total_spawn;
    if (g >= total) return;
    // command of this particle: last c with cmds[c].first <= g
    uint32_t lo = 0, hi = f.header->n_cmds;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (f.cmds[mid].first <= g) lo = mid; else hi = mid;
    }
    const SpawnCmd cmd = f.cmds[lo];
    const uint32_t j = g - cmd.first;
    const StreamDesc d = t.descs[cmd.stream];
    const StreamState st = t.states[cmd.stream];
    const uint32_t logical = st.spawn_base + cmd.dst_off + j;
    if (logical >= st.count) return; // dropped by the overflow clamp of the plan kernel
    const uint32_t slot = wrap(st.head + logical, d.capacity);

    const fw_emission_settings &es = t.emitters[cmd.emitter_idx];
    const DevParticleSettings &ps = t.settings[d.settings_idx];
    const SpawnerInput in = f.inputs[cmd.input_idx];

    // draws 0..11 in the reference's draw order (src/core.rs:438-466)
    const uint64_t serial = cmd.serial_base + j;
    const uint2 key = make_uint2((uint32_t)t.seed, (uint32_t)(t.seed >> 32));
    const uint32_t c0 = (uint32_t)serial, c1 = (uint32_t)(serial >> 32), c2 = cmd.spawner_key;
    const uint4 r0 = philox4x32_10(make_uint4(c0, c1, c2, (cmd.emitter_local << 8) | 0u), key);
    const uint4 r1 = philox4x32_10(make_uint4(c0, c1, c2, (cmd.emitter_local << 8) | 1u), key);
    const uint4 r2 = philox4x32_10(make_uint4(c0, c1, c2, (cmd.emitter_local << 8) | 2u), key);
    const float u_shape0 = u01(r0.x), u_shape1 = u01(r0.y), u_shape2 = u01(r0.z);
    const float u_vel_angle = u01(r0.w), u_vel_radius = u01(r1.x), u_vel_mag = u01(r1.y);
    const float u_radial = u01(r1.z), u_scale = u01(r1.w), u_life = u01(r2.x);
    const float u_ang_angle = u01(r2.y), u_ang_radius = u01(r2.z), u_ang_mag = u01(r2.w);
    const float kPi = 3.14159265358979323846f;

    // EmissionShape::generate_point (src/emission_shape.rs:18-39)
    V3 spawn_offset = v3(0.0f, 0.0f, 0.0f);
    if (es.shape_kind == FW_SHAPE_SPHERE) {
        const float u = u_shape0 * 2.0f * kPi, v = u_shape1 * kPi, r = u_shape2;
        const float sv = sinf(v);
        const V3 unit = v3(sv * cosf(u), cosf(v), sv * sinf(u));
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
            const float sp = sinf(p), cp = cosf(p);
            const V3 local = v3(sp * cosf(a), cp, sp * sinf(a));
            const Q4 arc = q_from_rotation_arc(v3(0.0f, 1.0f, 0.0f), normalize_or_zero(dir));
            dir = qrot(arc, local);
        }
        const float m = um * (rv.magnitude.max - rv.magnitude.min) + rv.magnitude.min;
        return dir * m;
    };
    const V3 iv = rand_vec3(es.initial_velocity, u_vel_angle, u_vel_radius, u_vel_mag);
    const float radial = u_radial * (es.initial_velocity_radial.max - es.initial_velocity_radial.min) + es.initial_velocity_radial.min;
    const Q4 orot{in.rotation[0], in.rotation[1], in.rotation[2], in.rotation[3]};
    // src/core.rs:440-448
    V3 velocity = (qrot(orot, iv) + normalize_or_zero(spawn_offset) * radial) * in.modifier_speed;
    const V3 inherit = es.inherit_parent_velocity ? v3(in.parent_velocity[0], in.parent_velocity[1], in.parent_velocity[2]) : v3(0.0f, 0.0f, 0.0f);
    velocity = velocity + inherit;
    const float initial_scale = (u_scale * (ps.initial_scale.max - ps.initial_scale.min) + ps.initial_scale.min) * in.modifier_scale;
    const V3 position = v3(in.translation[0], in.translation[1], in.translation[2]) + spawn_offset;
    const float lifetime = u_life * (ps.lifetime.max - ps.lifetime.min) + ps.lifetime.min;
    const V3 av = rand_vec3(es.initial_angular_velocity, u_ang_angle, u_ang_radius, u_ang_mag);
    const float4 base = sample_gradient(ps.base_color, 0.0f);
    const float4 emissive = sample_gradient(ps.emissive_color, 0.0f);

    float4 *row = d.rows + (size_t)slot * 4u;
    row[0] = make_float4(position.x, position.y, position.z, initial_scale);
    row[1] = make_float4(es.initial_rotation[0], es.initial_rotation[1], es.initial_rotation[2], es.initial_rotation[3]);
    row[2] = base;
    row[3] = emissive;
    d.s0[slot] = make_float4(velocity.x, velocity.y, velocity.z, 0.0f);
    d.s1[slot] = make_float4(av.x, av.y, av.z, lifetime);
    d.s2[slot] = initial_scale;
}

// ------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA unit; SASS UBLKCP) for the per-stream settings block
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_settings(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
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

// look-back status word of a tile: epoch<<34 | flag<<32 | value
constexpr unsigned long long kFlagAgg = 1ull, kFlagPrefix = 2ull;
__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct alignas(16) UpdateSmem {
    DevParticleSettings settings[2];
    float4 rows[kUpdateThreads / 32][32 * 4]; // per-warp 2 KB staging for the AoS row transposes
    uint64_t bar[2];
    uint32_t warp_alive[kUpdateThreads / 32];
    uint32_t excl_dead; // dead particles of the stream before this tile (compact variants)
};

// The fused per-frame update. One CTA processes whole tiles of 256 consecutive particles of one
// stream; persistent grid, tile = blockIdx.x + k*gridDim.x in increasing order (required by the
// look-back of the compact variants: a tile only ever waits on lower-numbered tiles).
template <bool COMPACT, bool COLLIDE>
__global__ void __launch_bounds__(kUpdateThreads) update_kernel(DeviceTables t, FrameDeviceInputs f, uint32_t variant) {
    __shared__ UpdateSmem sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_tiles = t.plan->n_tiles[variant];
    const uint32_t tile_base = t.plan->tile_base[variant];
    if (blockIdx.x >= n_tiles) return;
    const float dt = f.header->dt;
    const uint32_t epoch = f.header->epoch;

    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
    }
    __syncthreads();
    // prefetch the settings of the first tile
    if (tid == 0) {
        const TileEntry e = t.tiles[tile_base + blockIdx.x];
        bulk_load_settings(&sm.settings[0], &t.settings[t.descs[e.stream].settings_idx], sizeof(DevParticleSettings), &sm.bar[0]);
    }
    float4 *wrows = sm.rows[warp];
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
        const uint32_t buf = it & 1u;
        const TileEntry e = t.tiles[tile_base + tile];
        const StreamDesc d = t.descs[e.stream];
        StreamState *stp = &t.states[e.stream];
        const uint32_t head = stp->head, n_update = stp->count;
        // prefetch the next tile's settings into the other buffer (all threads left it at the
        // __syncthreads that closed the previous iteration)
        if (tid == 0 && tile + gridDim.x < n_tiles) {
            const TileEntry en = t.tiles[tile_base + tile + gridDim.x];
            bulk_load_settings(&sm.settings[buf ^ 1u], &t.settings[t.descs[en.stream].settings_idx], sizeof(DevParticleSettings), &sm.bar[buf ^ 1u]);
        }
        const uint32_t tile_first = e.tile * kTile;
        const uint32_t warp_first = tile_first + warp * 32u;
        const uint32_t i = warp_first + lane;
        const bool valid = i < n_update;
        const uint32_t slot = wrap(head + (valid ? i : 0u), d.capacity);

        // ---- loads: 5 independent 16-byte (one 4-byte) requests per thread in flight
        float4 S0 = make_float4(0.f, 0.f, 0.f, 0.f), S1 = make_float4(0.f, 0.f, 0.f, 1.f);
        float iscale = 0.f;
        if (valid) {
            S0 = d.s0[slot];
            S1 = d.s1[slot];
            iscale = d.s2[slot];
        }
        // rows: the warp's 32 rows are 2 KB contiguous (mod ring wrap); only bytes 0..31 of each
        // row (position/scale, rotation) are state. Lane q loads 16-byte chunk q of the 64
        // input chunks, fully coalesced, and they are transposed through shared memory.
        float4 in0 = make_float4(0.f, 0.f, 0.f, 0.f), in1 = in0;
        {
            const uint32_t p0 = lane >> 1, c = lane & 1u;
            const uint32_t ia = warp_first + p0, ib = warp_first + 16u + p0;
            if (ia < n_update) in0 = d.rows[(size_t)wrap(head + ia, d.capacity) * 4u + c];
            if (ib < n_update) in1 = d.rows[(size_t)wrap(head + ib, d.capacity) * 4u + c];
            // 2-chunk layout, 16-byte unit index = 2p + (c ^ ((p>>2)&1)): conflict-free both ways
            wrows[2u * p0 + (c ^ ((p0 >> 2) & 1u))] = in0;
            const uint32_t p1 = 16u + p0;
            wrows[2u * p1 + (c ^ ((p1 >> 2) & 1u))] = in1;
        }
        __syncwarp();
        const uint32_t sw2 = (lane >> 2) & 1u;
        const float4 P0 = wrows[2u * lane + (0u ^ sw2)];
        const float4 P1 = wrows[2u * lane + (1u ^ sw2)];
        __syncwarp();

        mbar_wait(&sm.bar[buf], (it >> 1) & 1u);
        const DevParticleSettings &ps = sm.settings[buf];

        // ---- reference src/core.rs:591-658, same order
        const float lifetime = S1.w;
        const float age = S0.w + dt;                 // :594
        bool alive = valid && !(age >= lifetime);    // :596-599
        float4 o0 = P0, o1 = P1, o2, o3, oS0 = S0, oS1 = S1;
        o2 = o3 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (alive) {
            const float age_percent = age / lifetime;                        // :601
            const float scale = iscale * sample_curve(ps.scale_curve, age_percent); // :602-605
            V3 pos = v3(P0.x, P0.y, P0.z), vel = v3(S0.x, S0.y, S0.z);
            bool should_destroy = false;
            if (COLLIDE) {
                particle_collision(t.colliders, t.n_colliders, ps.collision, pos, vel, dt, should_destroy); // :608-617
            } else {
                pos = pos + vel * dt;                                        // :619-623
            }
            if (should_destroy) {
                alive = false;                                               // :636-639
            } else {
                const V3 acc = v3(ps.acceleration[0], ps.acceleration[1], ps.acceleration[2]);
                vel = vel + (acc - vel * ps.linear_drag) * dt;               // :641-643
                V3 av = v3(S1.x, S1.y, S1.z);
                const Q4 rot = qmul(q_from_scaled_axis(av * dt), Q4{P1.x, P1.y, P1.z, P1.w}); // :645-647
                const V3 aacc = v3(ps.angular_acceleration[0], ps.angular_acceleration[1], ps.angular_acceleration[2]);
                av = av + (aacc - av * ps.angular_drag) * dt;                // :648-650
                o0 = make_float4(pos.x, pos.y, pos.z, scale);
                o1 = make_float4(rot.x, rot.y, rot.z, rot.w);
                o2 = sample_gradient(ps.base_color, age_percent);            // :652-653
                o3 = sample_gradient(ps.emissive_color, age_percent);        // :654-655
                oS0 = make_float4(vel.x, vel.y, vel.z, age);
                oS1 = make_float4(av.x, av.y, av.z, lifetime);
            }
        }
        const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
        const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
        const uint32_t n_alive_w = __popc(alive_mask);

        // ---- per-stream AABB of position -/+ scale (reference src/render.rs:681-692)
        {
            uint32_t mn[3], mx[3];
            mn[0] = alive ? enc_f32(o0.x - o0.w) : 0xFFFFFFFFu;
            mn[1] = alive ? enc_f32(o0.y - o0.w) : 0xFFFFFFFFu;
            mn[2] = alive ? enc_f32(o0.z - o0.w) : 0xFFFFFFFFu;
            mx[0] = alive ? enc_f32(o0.x + o0.w) : 0u;
            mx[1] = alive ? enc_f32(o0.y + o0.w) : 0u;
            mx[2] = alive ? enc_f32(o0.z + o0.w) : 0u;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
                mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
            }
            if (lane < 3u) {
                const uint32_t a = lane == 0 ? mn[0] : (lane == 1 ? mn[1] : mn[2]);
                const uint32_t b = lane == 0 ? mx[0] : (lane == 1 ? mx[1] : mx[2]);
                if (a < stp->aabb_min[lane]) atomicMin(&stp->aabb_min[lane], a);
                if (b > stp->aabb_max[lane]) atomicMax(&stp->aabb_max[lane], b);
            }
        }

        // ---- destination of the survivors
        uint32_t dst_first; // logical index the warp's first stored row goes to
        uint32_t my_rank;   // row position of this lane inside the warp's stored block
        if (COMPACT) {
            if (lane == 0) sm.warp_alive[warp] = n_alive_w;
            __syncthreads(); // every load of this tile has been consumed by now
            uint32_t before = 0, tile_alive = 0;
#pragma unroll
            for (uint32_t w = 0; w < kUpdateThreads / 32; w++) {
                const uint32_t a = sm.warp_alive[w];
                if (w < warp) before += a;
                tile_alive += a;
            }
            if (tid == 0) {
                const uint32_t tile_valid = min(n_update - tile_first, (uint32_t)kTile);
                const uint32_t tile_dead = tile_valid - tile_alive;
                unsigned long long *status = t.lookback + tile_base + tile;
                const unsigned long long tag = (unsigned long long)epoch << 34;
                uint32_t excl = 0;
                if (e.tile != 0u) {
                    st_release(status, tag | (kFlagAgg << 32) | tile_dead);
                    // decoupled look-back over the preceding tiles of the same stream
                    const unsigned long long *p = status - 1;
                    for (uint32_t back = 0; back < e.tile;) {
                        const unsigned long long w = ld_acquire(p);
                        if ((w >> 34) != (unsigned long long)epoch || ((w >> 32) & 3ull) == 0ull) continue; // not published yet
                        excl += (uint32_t)w;
                        if (((w >> 32) & 3ull) == kFlagPrefix) break;
                        back++;
                        p--;
                    }
                }
                st_release(status, tag | (kFlagPrefix << 32) | (excl + tile_dead));
                sm.excl_dead = excl;
                if (tile_first + kTile >= n_update) stp->dead = excl + tile_dead; // last tile of the stream
            }
            __syncthreads();
            dst_first = tile_first - sm.excl_dead + before;
            my_rank = __popc(alive_mask & ((1u << lane) - 1u));
        } else {
            dst_first = warp_first;
            my_rank = lane;
            const uint32_t n_dead_w = __popc(valid_mask & ~alive_mask);
            if (lane == 0 && n_dead_w) atomicAdd(&stp->dead, n_dead_w);
        }

        // ---- stores. s0/s1 (and s2 when particles move) go straight from the owning lane;
        // rows are transposed back through shared memory so each store instruction writes
        // whole 64-byte rows.
        if (alive) {
            const uint32_t dslot = wrap(head + dst_first + my_rank, d.capacity);
            d.s0[dslot] = oS0;
            d.s1[dslot] = oS1;
            if (COMPACT) d.s2[dslot] = iscale;
            // 4-chunk layout, 16-byte unit index = 4r + (c ^ ((r>>1)&3)): conflict-free both ways
            const uint32_t sw4 = (my_rank >> 1) & 3u;
            wrows[4u * my_rank + (0u ^ sw4)] = o0;
            wrows[4u * my_rank + (1u ^ sw4)] = o1;
            wrows[4u * my_rank + (2u ^ sw4)] = o2;
            wrows[4u * my_rank + (3u ^ sw4)] = o3;
        }
        __syncwarp();
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) {
            const uint32_t q = j * 32u + lane, r = q >> 2, c = q & 3u;
            const bool on = COMPACT ? (r < n_alive_w) : ((alive_mask >> r) & 1u);
            if (on) {
                const float4 v = wrows[4u * r + (c ^ ((r >> 1) & 3u))];
                d.rows[(size_t)wrap(head + dst_first + r, d.capacity) * 4u + c] = v;
            }
        }
        __syncthreads(); // settings buffer + staging reuse
    }
}

// ------------------------------------------------------------------------------------------
// pack: live instance rows of every stream, creation order, Vec order inside a stream
__global__ void __launch_bounds__(256) pack_prefix_kernel(DeviceTables t, uint32_t n_slots, unsigned long long *offsets, unsigned long long *n_rows) {
    // single thread-block serial prefix over streams (n_slots is small)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long acc = 0;
        for (uint32_t s = 0; s < n_slots; s++) {
            offsets[s] = acc;
            if (t.descs[s].capacity) acc += t.states[s].count - t.states[s].dead;
        }
        *n_rows = acc;
    }
}
__global__ void __launch_bounds__(256) pack_copy_kernel(DeviceTables t, uint32_t n_slots, const unsigned long long *offsets, float4 *dst, uint64_t cap_rows) {
    const uint32_t s = blockIdx.y;
    if (s >= n_slots) return;
    const StreamDesc d = t.descs[s];
    if (d.capacity == 0u) return;
    const StreamState st = t.states[s];
    const bool fifo = (d.variant == kFifo || d.variant == kFifoCollide);
    const uint32_t live = st.count - st.dead;
    const uint32_t first = wrap(st.head + (fifo ? st.dead : 0u), d.capacity);
    const unsigned long long off = offsets[s];
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < (uint64_t)live * 4u; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(q >> 2), c = (uint32_t)(q & 3u);
        if (off + r >= cap_rows) return;
        dst[(off + r) * 4u + c] = d.rows[(size_t)wrap(first + r, d.capacity) * 4u + c];
    }
}

// ------------------------------------------------------------------------------------------
cudaError_t launch_plan(const DeviceTables &t, const FrameDeviceInputs &f, cudaStream_t s) {
    plan_kernel<<<1, 1024, 0, s>>>(t, f);
    return cudaGetLastError();
}
cudaError_t launch_spawn(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t total_spawn, cudaStream_t s) {
    if (total_spawn == 0) return cudaSuccess;
    spawn_kernel<<<(total_spawn + 255u) / 256u, 256, 0, s>>>(t, f);
    return cudaGetLastError();
}
cudaError_t launch_update(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant, int grid, cudaStream_t s) {
    switch (variant) {
    case kFifo: update_kernel<false, false><<<grid, kUpdateThreads, 0, s>>>(t, f, variant); break;
    case kCompact: update_kernel<true, false><<<grid, kUpdateThreads, 0, s>>>(t, f, variant); break;
    case kFifoCollide: update_kernel<false, true><<<grid, kUpdateThreads, 0, s>>>(t, f, variant); break;
    case kCompactCollide: update_kernel<true, true><<<grid, kUpdateThreads, 0, s>>>(t, f, variant); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t update_grid_size(int device, int *grids) {
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    int occ[kNumVariants] = {0, 0, 0, 0};
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[kFifo], update_kernel<false, false>, kUpdateThreads, 0);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[kCompact], update_kernel<true, false>, kUpdateThreads, 0);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[kFifoCollide], update_kernel<false, true>, kUpdateThreads, 0);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[kCompactCollide], update_kernel<true, true>, kUpdateThreads, 0);
    if (e != cudaSuccess) return e;
    for (int v = 0; v < (int)kNumVariants; v++) grids[v] = sms * (occ[v] > 0 ? occ[v] : 1);
    return cudaSuccess;
}
cudaError_t launch_pack_instances(const DeviceTables &t, uint32_t n_slots, float4 *dst, uint64_t cap_rows, unsigned long long *n_rows, cudaStream_t s) {
    // offsets scratch lives right behind n_rows (the host allocates n_slots + 1 words)
    unsigned long long *offsets = n_rows + 1;
    pack_prefix_kernel<<<1, 32, 0, s>>>(t, n_slots, offsets, n_rows);
    if (n_slots) {
        dim3 grid(64, n_slots);
        pack_copy_kernel<<<grid, 256, 0, s>>>(t, n_slots, offsets, dst, cap_rows);
    }
    return cudaGetLastError();
}

} // namespace fw
