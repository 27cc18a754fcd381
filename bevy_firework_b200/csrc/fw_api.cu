// fw_api.cu -- host side of libfirework_b200.so: the C ABI of include/firework_b200.h.
//
// What lives here is the part of the reference's two systems that is inherently sequential
// host logic -- which emitters are enabled, emission pacing (reference src/core.rs:396-428,
// 553-575), OneShot/OnDemand bookkeeping, the finished condition (:674-688) -- plus the
// management of device memory. Everything per-particle runs in fw_kernels.cu.
//
// There is no CPU fallback: every entry point needs a live CUDA context.
#include <unistd.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstddef>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "fw_internal.h"

using namespace fw;

namespace {

thread_local std::string g_global_error;

constexpr uint32_t kRing = 4; // frames in flight (pinned parameter blocks + state readbacks)
constexpr uint32_t kGraphWarmFrames = 8; // frames of unchanged topology before frames are replayed as graphs

struct Emitter {
    fw_emission_settings es;
    float last_emission = 0.f;
    float time_passed_in_cycle = 0.f;
    bool enabled = true;
    bool emits_on_other_particles = false;
    uint64_t serial = 0;   // particles spawned since reset (RNG protocol; Global emitters)
    uint32_t dev_idx = 0xFFFFFFFFu; // index into the device emitter array (and nested_serial); ~0 = none yet
    uint32_t lea_index = 0; // Nested: which last_emitted_age array of the target stream
};

struct Block { // one device allocation holding the packs of a stream
    void *base = nullptr;
    uint32_t capacity = 0;
    uint32_t n_lea = 0;
};

struct Stream {
    uint32_t slot = 0;
    uint32_t type = 0;
    Block block;
    Block destroyed; // particles destroyed by the last update (capture_destroyed types only)
    uint32_t n_lea = 0;     // nested emitters targeting this type
    bool injected = false;  // state was written by the host: nested emitters may catch up at once
    bool accounted = false; // counted in tiles_needed / variant_streams (a failed reset may stop before that)
    uint32_t variant = kFifo;
    uint32_t flags = kStoreAll; // kStore*: which derived / constant fields the stream keeps per particle
    DevParticleSettings dev;    // device form of ps incl. the constants of the packs it does not keep
    uint64_t n_hi = 0;        // host-side upper bound of the live count
    uint64_t born_frame = 0;  // readbacks of older frames do not describe this stream
    fw_particle_settings ps;
};

struct Spawner {
    uint32_t key = 0;
    std::vector<Stream> streams; // one per particle type
    std::vector<Emitter> emitters;
    SpawnerInput input;
    uint64_t manual_queued_count = 0;
    bool initialized = false;
    bool finished_notified = false;
};

struct FrameSlot {
    uint8_t *host = nullptr; // pinned
    uint8_t *dev = nullptr;
    size_t bytes = 0;
    uint8_t *readback = nullptr;        // pinned copy of the frame's state buffer [PlanOut | states]
    StreamState *states_host = nullptr; // = readback + sizeof(PlanOut)
    uint32_t states_slots = 0;
    PlanOut *plan_host = nullptr;       // = readback
    cudaEvent_t done = nullptr;       // the frame's state readback has landed (recorded on the copy stream)
    cudaEvent_t ev_h2d = nullptr;     // the frame's parameter block is on the device (copy stream)
    cudaEvent_t ev_kernels = nullptr; // the frame's kernels are done (main stream)
    bool in_flight = false;
    uint64_t frame = 0;
    std::vector<uint32_t> spawn_per_slot; // host copy (all phases summed), for the n_hi bound
    // profiling
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool profiled = false;        // events ev[0..3] were recorded around the kernels
    bool pending_account = false; // not yet added to the profile sums
    uint64_t particles_spawned = 0;
    uint32_t launches = 0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    // the frame as an instantiated CUDA graph (valid while the topology version matches)
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_version = 0;
    uint32_t graph_launches = 0;
};

} // namespace

// most update tiles a ring of `capacity` slots can have (slot-aligned FIFO tiles: one extra)
static inline uint64_t max_tiles_of(uint64_t capacity) { return (capacity + kTile - 1) / kTile + 1; }

struct fw_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    // forked branch of a frame: spawn_kernel<STEP> runs here concurrently with update_kernel
    cudaStream_t side_stream = nullptr;
    // parameter uploads and state readbacks run here, overlapping the neighbouring frames' kernels
    cudaStream_t copy_stream = nullptr; // parameter uploads (H2D) only: never waits for kernels
    cudaStream_t rb_stream = nullptr;   // per-frame state readbacks (D2H), each behind its frame's kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool concurrent_spawn = true;
    uint64_t seed = 0;
    uint32_t flags = 0;
    std::string error;

    std::vector<std::unique_ptr<Spawner>> spawners; // creation order
    std::unordered_map<uint32_t, Spawner *> by_key;

    // device tables
    uint32_t slots_cap = 0, n_slots = 0; // n_slots = high-water mark of used slots
    std::vector<uint32_t> free_slots;
    std::vector<Stream *> slot_owner;
    StreamDesc *d_descs = nullptr;
    // double-buffered [PlanOut | StreamState x slots_cap]; frame f writes buffer f&1 and reads the
    // other one, so the buffer holding the current state is always frame_no & 1
    uint8_t *d_statebuf[2] = {nullptr, nullptr};
    DevParticleSettings *d_settings = nullptr;
    std::vector<StreamDesc> h_descs;
    uint32_t emitters_cap = 0, n_emitters = 0;
    std::vector<uint32_t> free_emitters;
    fw_emission_settings *d_emitters = nullptr;
    std::vector<fw_emission_settings> h_emitters; // host copy, indexed like d_emitters
    fw_collider *d_colliders = nullptr;
    uint8_t *d_broadphase = nullptr; // BroadPhaseHeader blob, broadphase_bytes(n_colliders)
    bool colliders_revolved = false; // the set contains cylinders / cones: kernels built with their exact test
    uint32_t n_colliders = 0;
    // pinned staging of (colliders | BVH nodes) for asynchronous re-uploads, a small ring guarded by events
    uint8_t *h_col_stage[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t col_stage_bytes = 0;
    cudaEvent_t col_stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t col_stage_next = 0;
    uint32_t *d_tile_prefix = nullptr; // kNumVariants x (slots_cap + 1)
    uint8_t *d_stage = nullptr;        // staging of ParticleData rows (host mirror reads/writes)
    size_t stage_bytes = 0;
    unsigned long long *d_lookback = nullptr;
    uint32_t tiles_cap = 0;
    uint64_t tiles_needed = 0; // sum over streams of max_tiles_of(capacity)
    uint32_t device_error_flags = 0; // accumulated from the per-frame plan readbacks
    unsigned long long *d_pack = nullptr; // n_rows + per-slot offsets
    uint32_t pack_cap = 0;
    unsigned long long *h_pack = nullptr; // pinned
    float4 *d_extract = nullptr;          // staging of fw_extract_instances
    uint64_t extract_cap = 0;
    // fw_extract_begin / fw_extract_wait: two staging slots alternate; the D2H copies run on xt_stream
    struct ExtractSlot {
        float4 *d_rows = nullptr;
        uint64_t cap_rows = 0;
        unsigned long long *d_offsets = nullptr, *h_offsets = nullptr; // [0] = n rows, [1 + k] = first row of listed stream k
        uint32_t offsets_cap = 0;
        std::vector<uint32_t> slots; // stream slots packed, in order
        uint32_t *d_slots = nullptr;
        void *host_dst = nullptr;
        uint64_t host_cap_rows = 0;
        cudaEvent_t packed = nullptr, landed = nullptr;
        bool pending = false;
    } xt[2];
    cudaStream_t xt_stream = nullptr;
    uint64_t xt_issued = 0, xt_waited = 0;
    // fw_export_instances_fd: a VMM allocation exported as a POSIX file descriptor
    unsigned long long vmm_handle = 0;
    unsigned long long vmm_ptr = 0;
    size_t vmm_bytes = 0;
    // multi-GPU render extract over peer memory (fw_gather_*)
    uint8_t *d_gather = nullptr; // this rank's gather buffer: [GatherHeader | n_ranks regions]
    GatherPeers gather{};        // base[r] = rank r's buffer as mapped here (base[my_rank] = d_gather)
    bool gather_ipc[kMaxGatherRanks] = {false}; // base[r] came from cudaIpcOpenMemHandle
    bool gather_connected = false;
    uint64_t gather_epoch = 0;
    GatherHeader *h_gather = nullptr; // pinned copy of our header (fw_gather_result)
    cudaEvent_t user_events[16] = {};

    std::map<size_t, std::vector<void *>> block_cache; // bytes -> free device blocks
    FrameSlot ring[kRing];
    uint64_t frame_no = 0; // frames submitted so far; epoch of the next frame = frame_no + 1
    // topology = everything a captured frame graph bakes in (table pointers, slot / emitter /
    // spawner counts, active variants, colliders). Any change bumps the version.
    uint64_t topo_version = 1;
    uint32_t topo_stable_frames = 0;
    uint32_t live_emitters = 0;
    uint32_t live_nested = 0; // nested emitters over all spawners
    uint32_t n_phases = 1;    // 1 + max nested emitters per spawner
    bool use_graphs = true;
    // nested emission scratch
    uint32_t *d_nested_scratch = nullptr;
    uint64_t nested_scratch_cap = 0;
    unsigned long long *d_nested_serial = nullptr; // per device emitter
    NestedOut *d_nested_out = nullptr;
    uint32_t nested_out_cap = 0;
    int grids[kNumVariants] = {0};
    int team_size = 0; // CTAs per team of the compacting update kernels (= SM count)
    uint32_t variant_streams[kNumVariants] = {0};
    float prev_dt = 0.f; // dt of the previous frame (nested emission reads ages that frame advanced)
    // per-frame scratch of fw_frame, kept between frames so that a steady-state frame allocates nothing
    std::vector<SpawnCmd> fr_cmds[kMaxPhases];
    std::vector<NestedCmd> fr_nested[kMaxPhases];
    std::vector<uint32_t> fr_spawn_per_slot;
    std::vector<uint64_t> fr_add;
    std::vector<SpawnerInput> fr_inputs;
    std::vector<Spawner *> fr_targets;
    struct PacingUndo { // what fw_frame changed in an emitter / spawner before the frame was certain to be submitted
        Emitter *e;
        Spawner *sp;
        float last_emission, time_passed_in_cycle;
        bool enabled;
        uint64_t serial, manual_queued_count;
    };
    std::vector<PacingUndo> fr_undo;

    // profile accumulators
    fw_frame_profile prof_last{};
    fw_frame_profile prof_sum{};
    uint32_t prof_frames = 0;
    uint32_t prof_timed_frames = 0;
    bool profiling = false; // timed frames are launched kernel by kernel (events inside a graph do not time)
    // last exact state snapshot (valid after refresh_exact)
    std::vector<StreamState> snapshot;
    bool snapshot_valid = false; // no device work was enqueued since `snapshot` was read
    bool readback_is_current = false; // the newest frame's pinned state readback == device state
};

namespace {

int fail(fw_context *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    g_global_error = buf;
    return code;
}

inline void topo_changed(fw_context *ctx) {
    ctx->snapshot_valid = false;
    ctx->readback_is_current = false;
    ctx->topo_version++;
    ctx->topo_stable_frames = 0;
}

// wait for everything the context has enqueued (kernels on the main stream, then the readbacks
// that follow them on the readback stream)
inline cudaError_t sync_all(fw_context *ctx) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && ctx->copy_stream) e = cudaStreamSynchronize(ctx->copy_stream);
    if (e == cudaSuccess && ctx->rb_stream) e = cudaStreamSynchronize(ctx->rb_stream);
    return e;
}
inline size_t statebuf_bytes(uint32_t slots) { return sizeof(PlanOut) + sizeof(StreamState) * (size_t)slots; }
inline int cur_buf(const fw_context *ctx) { return (int)(ctx->frame_no & 1u); }
inline StreamState *states_of(fw_context *ctx, int i) { return (StreamState *)(ctx->d_statebuf[i] + sizeof(PlanOut)); }
inline PlanOut *plan_of(fw_context *ctx, int i) { return (PlanOut *)ctx->d_statebuf[i]; }
inline StreamState *cur_states(fw_context *ctx) { return states_of(ctx, cur_buf(ctx)); }

#define CU(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? FW_ERR_OUT_OF_MEMORY : FW_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

inline int enter(fw_context *ctx) {
    if (!ctx) return fail(nullptr, FW_ERR_INVALID_ARGUMENT, "null context");
    cudaError_t e = cudaSetDevice(ctx->device); // callable from any thread
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    return FW_OK;
}
#define ENTER(ctx)                      \
    do {                                \
        int rc__ = enter(ctx);          \
        if (rc__ != FW_OK) return rc__; \
    } while (0)

// ---- Rust f32 helpers used by emission pacing (core::f32::rem_euclid / div_euclid, `as usize`)
// (fmod is exact and costs ~25 ns; the common cases need none: 0 <= a < b is its own remainder, and a
// remainder is negative only when a is -- 512 emitters per frame pay for the difference)
inline float rem_euclid_f32(float a, float b) {
    if (a >= 0.0f && a < b) return a;
    const float r = std::fmod(a, b);
    return r < 0.0f ? r + std::fabs(b) : r;
}
inline float div_euclid_f32(float a, float b) {
    const float q = std::trunc(a / b);
    if (a < 0.0f && std::fmod(a, b) < 0.0f) return b > 0.0f ? q - 1.0f : q + 1.0f;
    return q;
}
inline uint64_t f32_as_usize(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)f;
}
// reference src/core.rs:553-575 compute_emission_count
inline void compute_emission_count(float time_passed_in_cycle, float last_emission, float cycle_duration,
                                   float offset_start, float offset_end, float particles_per_cycle,
                                   uint64_t &times, float &next_last_emission) {
    const float percent_passed = time_passed_in_cycle / cycle_duration;
    const float last_emission_percent = last_emission / cycle_duration;
    const float lo = std::fmax(last_emission_percent, offset_start);
    const float percent_passed_since_emission = std::fmin(percent_passed, offset_end) - lo;
    const float percent_between_emissions = (offset_end - offset_start) / particles_per_cycle;
    const float times_needed_to_emit = div_euclid_f32(percent_passed_since_emission, percent_between_emissions);
    times = f32_as_usize(times_needed_to_emit);
    const float next_last_emission_percent = lo + times_needed_to_emit * percent_between_emissions;
    next_last_emission = next_last_emission_percent * cycle_duration;
}

inline bool is_fifo(uint32_t v) { return variant_is_fifo(v); }
// A compacting ring compacts out of place inside its own ring (fw_kernels.cu: usable_capacity,
// live_first): it may only be half full, and after a frame its live particles start at
// head + count (a FIFO ring's: at head + dead).
inline uint64_t usable_capacity(uint32_t variant, uint64_t capacity) { return is_fifo(variant) ? capacity : capacity / 2; }
inline uint32_t live_first(uint32_t variant, const StreamState &s, uint32_t capacity) {
    return (uint32_t)(((uint64_t)s.head + (is_fifo(variant) ? s.dead : s.count)) % std::max(1u, capacity));
}
// the state to inject for `live` particles that sit at the start of the block
inline StreamState injected_state(uint32_t variant, uint32_t live, uint32_t capacity) {
    StreamState ns{};
    ns.count = live;
    ns.head = is_fifo(variant) || live == 0 ? 0u : (capacity - live % capacity) % capacity; // head + count == 0 (mod capacity)
    return ns;
}

inline uint32_t round_capacity(uint64_t want) {
    // round up so that freed blocks are reusable: multiples of 1024 up to 64 Ki, then 1/8-octave
    if (want < 1024) want = 1024;
    if (want <= 65536) return (uint32_t)((want + 1023) / 1024 * 1024);
    uint64_t p = 65536;
    while (p * 2 <= want) p *= 2;
    const uint64_t step = p / 8;
    uint64_t r = (want + step - 1) / step * step;
    if (r > 0xFFFFFF00ull) r = 0xFFFFFF00ull;
    return (uint32_t)r;
}
inline size_t block_bytes(uint32_t cap, uint32_t n_lea) { return (size_t)cap * (kBytesPerSlot + 4u * n_lea); }
inline StreamDesc block_desc(const Block &b, uint32_t variant, uint32_t flags, const Block *destroyed = nullptr) {
    StreamDesc d{};
    d.base = (uint8_t *)b.base;
    d.destroyed_base = destroyed ? (uint8_t *)destroyed->base : nullptr;
    d.capacity = b.capacity;
    d.variant = variant;
    d.n_lea = b.n_lea;
    d.flags = flags;
    return d;
}

int alloc_block(fw_context *ctx, uint32_t capacity, uint32_t n_lea, Block &out) {
    const size_t bytes = block_bytes(capacity, n_lea);
    auto it = ctx->block_cache.find(bytes);
    if (it != ctx->block_cache.end() && !it->second.empty()) {
        out.base = it->second.back();
        it->second.pop_back();
        out.capacity = capacity;
        out.n_lea = n_lea;
        return FW_OK;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        // drop the cache and retry once
        for (auto &kv : ctx->block_cache)
            for (void *q : kv.second) cudaFree(q);
        ctx->block_cache.clear();
        (void)cudaGetLastError();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "cudaMalloc of a %u-particle stream failed: %s", capacity, cudaGetErrorString(e));
    out.base = p;
    out.capacity = capacity;
    out.n_lea = n_lea;
    return FW_OK;
}
void release_block(fw_context *ctx, Block &b) {
    // all work is ordered on one CUDA stream, so a cached block can be handed out again at once
    if (b.base) ctx->block_cache[block_bytes(b.capacity, b.n_lea)].push_back(b.base);
    b.base = nullptr;
    b.capacity = 0;
}

template <typename T>
int grow_device_array(fw_context *ctx, T *&ptr, uint32_t old_n, uint32_t new_n) {
    T *np = nullptr;
    CU(ctx, cudaMalloc((void **)&np, sizeof(T) * (size_t)new_n));
    CU(ctx, cudaMemsetAsync(np, 0, sizeof(T) * (size_t)new_n, ctx->stream));
    if (ptr && old_n) CU(ctx, cudaMemcpyAsync(np, ptr, sizeof(T) * (size_t)old_n, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, sync_all(ctx));
    if (ptr) CU(ctx, cudaFree(ptr));
    ptr = np;
    return FW_OK;
}

int ensure_slots(fw_context *ctx, uint32_t need) {
    if (need <= ctx->slots_cap) return FW_OK;
    uint32_t ncap = std::max(1024u, ctx->slots_cap * 2);
    while (ncap < need) ncap *= 2;
    int rc;
    if ((rc = grow_device_array(ctx, ctx->d_descs, ctx->slots_cap, ncap))) return rc;
    for (int i = 0; i < 2; i++) {
        uint8_t *nb = nullptr;
        CU(ctx, cudaMalloc((void **)&nb, statebuf_bytes(ncap)));
        CU(ctx, cudaMemsetAsync(nb, 0, statebuf_bytes(ncap), ctx->stream));
        if (ctx->d_statebuf[i]) CU(ctx, cudaMemcpyAsync(nb, ctx->d_statebuf[i], statebuf_bytes(ctx->slots_cap), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(ctx, sync_all(ctx));
        if (ctx->d_statebuf[i]) CU(ctx, cudaFree(ctx->d_statebuf[i]));
        ctx->d_statebuf[i] = nb;
    }
    if ((rc = grow_device_array(ctx, ctx->d_settings, ctx->slots_cap, ncap))) return rc;
    if (ctx->d_tile_prefix) CU(ctx, cudaFree(ctx->d_tile_prefix));
    ctx->d_tile_prefix = nullptr;
    CU(ctx, cudaMalloc((void **)&ctx->d_tile_prefix, sizeof(uint32_t) * (size_t)kNumVariants * (ncap + 1)));
    ctx->h_descs.resize(ncap);
    ctx->slot_owner.resize(ncap, nullptr);
    ctx->slots_cap = ncap;
    topo_changed(ctx);
    return FW_OK;
}
int ensure_emitters(fw_context *ctx, uint32_t need) {
    if (need <= ctx->emitters_cap) return FW_OK;
    uint32_t ncap = std::max(1024u, ctx->emitters_cap * 2);
    while (ncap < need) ncap *= 2;
    int rc = grow_device_array(ctx, ctx->d_emitters, ctx->emitters_cap, ncap);
    if (rc) return rc;
    if ((rc = grow_device_array(ctx, ctx->d_nested_serial, ctx->emitters_cap, ncap))) return rc;
    ctx->h_emitters.resize(ncap);
    ctx->emitters_cap = ncap;
    topo_changed(ctx);
    return FW_OK;
}
int ensure_tiles(fw_context *ctx) {
    const uint64_t need = ctx->tiles_needed + 16;
    if (need <= ctx->tiles_cap) return FW_OK;
    uint64_t ncap = std::max<uint64_t>(4096, (uint64_t)ctx->tiles_cap * 2);
    while (ncap < need) ncap *= 2;
    if (ncap > 0xFFFFFFFFull) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "tile table too large");
    CU(ctx, sync_all(ctx));
    if (ctx->d_lookback) CU(ctx, cudaFree(ctx->d_lookback));
    ctx->d_lookback = nullptr;
    // two words per tile: [0, ncap) look-back status / precounted exclusive prefix, [ncap, 2 ncap)
    // precounted per-warp death counts (count_kernel)
    CU(ctx, cudaMalloc((void **)&ctx->d_lookback, sizeof(unsigned long long) * ncap * 2));
    CU(ctx, cudaMemsetAsync(ctx->d_lookback, 0, sizeof(unsigned long long) * ncap * 2, ctx->stream));
    ctx->tiles_cap = (uint32_t)ncap;
    topo_changed(ctx);
    return FW_OK;
}

inline uint32_t f32_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
// glam Quat::mul_quat with the identity on the left, in the kernels' fp32 operation order (this file
// is compiled -fmad=false / -ffp-contract=off like them): what src/core.rs:645-647 does to a rotation
// when the angular velocity is zero
inline void qmul_identity(const float b[4], float out[4]) {
    const float ax = 0.0f, ay = 0.0f, az = 0.0f, aw = 1.0f;
    out[0] = aw * b[0] + ax * b[3] + ay * b[2] - az * b[1];
    out[1] = aw * b[1] - ax * b[2] + ay * b[3] + az * b[0];
    out[2] = aw * b[2] + ax * b[1] - ay * b[0] + az * b[3];
    out[3] = aw * b[3] - ax * b[0] - ay * b[1] - az * b[2];
}
// What can be PROVED about a particle type at reset, and therefore need not be stored per particle
// (layout: fw_internal.h). Every proof is about the reference's arithmetic on these exact settings:
//  * FIFO (deaths are a prefix of the Vec): every particle gets the same lifetime (ages are monotone
//    in Vec order) and nothing else can kill a particle;
//  * static (no kVarRot): every emitter of the type draws angular velocity = direction * 0 = +-0
//    (RandVec3 with magnitude [0, 0], finite direction / spread) and angular_acceleration is +0 with a
//    finite drag: src/core.rs:648-650 then gives +0 after the first update, for ever (dt is finite
//    and >= 0, fw_frame checks), from_scaled_axis(0) is the identity (:645-647) and the rotation is
//    q = identity * initial_rotation, which must be the same for all emitters and a fixed point;
//  * constant base / emissive gradient, constant scale curve: :602-605, 652-655 return colors[0] /
//    initial_scale * values[0] on every frame (a particle is always updated in the frame it spawns);
//  * a handler for destroyed particles wants the previous frame's colours: such types keep everything.
void stream_proofs(const fw_particle_settings &ps, uint32_t type, const fw_emission_settings *es, uint32_t n_emitters,
                   uint32_t &variant, uint32_t &flags, DevParticleSettings &dev) {
    const bool collide = ps.collision.enabled != 0;
    const bool same_lifetime = ps.lifetime.min == ps.lifetime.max;
    const bool fifo = same_lifetime && !(collide && ps.collision.destroy_on_collision) && !ps.capture_destroyed;
    variant = (fifo ? 0u : (uint32_t)kVarCompact) | (collide ? (uint32_t)kVarCollide : 0u);
    flags = 0;
    if (ps.base_color.kind != FW_CURVE_CONSTANT) flags |= kStoreBase;
    if (ps.emissive_color.kind != FW_CURVE_CONSTANT) flags |= kStoreEmi;
    if (ps.scale_curve.kind != FW_CURVE_CONSTANT) flags |= kStoreScale;
    if (!same_lifetime) flags |= kStoreLife;
    dev.const_lifetime = 0.5f * (ps.lifetime.max - ps.lifetime.min) + ps.lifetime.min; // RandF32::generate with max == min
    bool is_static = f32_bits(ps.angular_acceleration[0]) == 0u && f32_bits(ps.angular_acceleration[1]) == 0u &&
                     f32_bits(ps.angular_acceleration[2]) == 0u && std::isfinite(ps.angular_drag);
    bool have_rotation = false;
    float rot[4] = {0.f, 0.f, 0.f, 1.f};
    for (uint32_t i = 0; i < n_emitters && is_static; i++) {
        if (es[i].particle_index != type) continue;
        const fw_rand_vec3 &av = es[i].initial_angular_velocity;
        is_static = av.magnitude.min == 0.0f && av.magnitude.max == 0.0f && std::isfinite(av.direction[0]) &&
                    std::isfinite(av.direction[1]) && std::isfinite(av.direction[2]) && std::isfinite(av.spread);
        float r1[4], r2[4];
        qmul_identity(es[i].initial_rotation, r1);
        qmul_identity(r1, r2);
        for (int k = 0; k < 4; k++) is_static = is_static && std::isfinite(r1[k]) && f32_bits(r1[k]) == f32_bits(r2[k]);
        if (have_rotation) {
            for (int k = 0; k < 4; k++) is_static = is_static && f32_bits(r1[k]) == f32_bits(rot[k]);
        } else {
            memcpy(rot, r1, sizeof(rot));
            have_rotation = true;
        }
    }
    if (!have_rotation) is_static = false; // no emitter feeds the type: nothing to prove from
    if (ps.capture_destroyed) {
        is_static = false;
        flags = kStoreAll;
    }
    memcpy(dev.const_rotation, rot, sizeof(rot));
    if (!is_static) variant |= kVarRot;
}
// host-written rows (fw_write_particles) against the proofs: which flags / variant bits they break
void check_rows_against_proofs(const Stream &st, const fw_particle_data *in, uint64_t n, uint32_t &variant, uint32_t &flags) {
    variant = st.variant;
    flags = st.flags;
    const DevParticleSettings &dev = st.dev;
    for (uint64_t i = 0; i < n; i++) {
        const fw_particle_data &r = in[i];
        if (!variant_rotates(variant)) {
            bool ok = r.angular_velocity[0] == 0.0f && r.angular_velocity[1] == 0.0f && r.angular_velocity[2] == 0.0f;
            for (int k = 0; k < 4; k++) ok = ok && f32_bits(r.rotation[k]) == f32_bits(dev.const_rotation[k]);
            if (!ok) variant |= kVarRot;
        }
        if (!(flags & kStoreBase) && memcmp(r.base_color, &dev.base_color.colors[0], 16) != 0) flags |= kStoreBase;
        if (!(flags & kStoreEmi) && memcmp(r.emissive_color, &dev.emissive_color.colors[0], 16) != 0) flags |= kStoreEmi;
        if (!(flags & kStoreScale) && f32_bits(r.scale) != f32_bits(r.initial_scale * dev.scale_curve.values[0])) flags |= kStoreScale;
        const bool same_life = f32_bits(r.lifetime) == f32_bits(dev.const_lifetime);
        if (!(flags & kStoreLife) && !same_life) flags |= kStoreLife;
        // FIFO: every lifetime is the type's constant and ages do not increase along the Vec
        if (variant_is_fifo(variant) && (!same_life || (i != 0 && r.age > in[i - 1].age))) variant |= kVarCompact;
    }
}

void fill_dev_settings(const fw_particle_settings &ps, DevParticleSettings &d) {
    memset(&d, 0, sizeof(d));
    d.scale_curve.kind = ps.scale_curve.kind;
    d.scale_curve.n = ps.scale_curve.n;
    memcpy(d.scale_curve.times, ps.scale_curve.times, sizeof(d.scale_curve.times));
    memcpy(d.scale_curve.values, ps.scale_curve.values, sizeof(d.scale_curve.values));
    const fw_gradient *src[2] = {&ps.base_color, &ps.emissive_color};
    DevGradient *dst[2] = {&d.base_color, &d.emissive_color};
    for (int k = 0; k < 2; k++) {
        dst[k]->kind = src[k]->kind;
        dst[k]->n = src[k]->n;
        memcpy(dst[k]->times, src[k]->times, sizeof(dst[k]->times));
        memcpy(dst[k]->colors, src[k]->colors, sizeof(dst[k]->colors));
    }
    memcpy(d.acceleration, ps.acceleration, sizeof(float) * 3);
    d.linear_drag = ps.linear_drag;
    memcpy(d.angular_acceleration, ps.angular_acceleration, sizeof(float) * 3);
    d.angular_drag = ps.angular_drag;
    d.collision = ps.collision;
    d.lifetime = ps.lifetime;
    d.initial_scale = ps.initial_scale;
    d.const_rotation[3] = 1.0f; // (stream_proofs fills the constants)
}

int validate_curve(fw_context *ctx, uint32_t kind, uint32_t n, const float *times, const char *what) {
    if (kind > FW_CURVE_UNEVEN) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "%s: unknown curve kind %u", what, kind);
    if (kind == FW_CURVE_CONSTANT) {
        if (n < 1) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "%s: Cannot create curve from 0 samples", what);
        return FW_OK;
    }
    if (n < 2 || n > FW_MAX_KNOTS) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "%s: %u samples (need 2..%u)", what, n, FW_MAX_KNOTS);
    if (kind == FW_CURVE_UNEVEN)
        for (uint32_t i = 0; i < n; i++) {
            if (!std::isfinite(times[i])) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "%s: non-finite sample time", what);
            if (i && !(times[i] > times[i - 1])) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "%s: sample times must be strictly increasing", what);
        }
    return FW_OK;
}

uint64_t estimate_capacity(const Spawner &sp, uint32_t type) {
    const fw_particle_settings &ps = sp.streams[type].ps;
    if (ps.capacity_hint) return ps.capacity_hint;
    double est = 0.0;
    const double life = std::max(0.0f, std::max(ps.lifetime.min, ps.lifetime.max));
    for (const Emitter &e : sp.emitters) {
        if (e.es.particle_index != type) continue;
        if (e.es.pacing_kind == FW_PACING_ONE_SHOT) est += (double)e.es.one_shot_count;
        else if (e.es.pacing_kind == FW_PACING_COUNT_OVER_DURATION && e.es.duration > 0.f && e.es.mode == FW_MODE_GLOBAL)
            est += (double)e.es.count / e.es.duration * (life + 0.05) * 1.02 + 64.0;
        else est += 1024.0;
    }
    if (est > 268435456.0) est = 268435456.0;
    return (uint64_t)est + 256;
}

void free_stream(fw_context *ctx, Stream &st) {
    if (!st.block.base) return; // never got a slot / block (failed reset)
    if (st.accounted) {
        ctx->tiles_needed -= max_tiles_of(st.block.capacity);
        ctx->variant_streams[st.variant]--;
        st.accounted = false;
    }
    release_block(ctx, st.block);
    release_block(ctx, st.destroyed);
    StreamDesc zero{};
    ctx->h_descs[st.slot] = zero;
    cudaMemcpyAsync(ctx->d_descs + st.slot, &zero, sizeof(zero), cudaMemcpyHostToDevice, ctx->stream);
    ctx->slot_owner[st.slot] = nullptr;
    ctx->free_slots.push_back(st.slot);
}
void free_spawner_resources(fw_context *ctx, Spawner &sp) {
    topo_changed(ctx);
    ctx->live_emitters -= (uint32_t)sp.emitters.size();
    for (Stream &st : sp.streams) free_stream(ctx, st);
    for (Emitter &e : sp.emitters)
        if (e.dev_idx != 0xFFFFFFFFu) ctx->free_emitters.push_back(e.dev_idx);
    sp.streams.clear();
    sp.emitters.clear();
}

int upload_desc(fw_context *ctx, const Stream &st) {
    const StreamDesc d = block_desc(st.block, st.variant, st.flags, st.destroyed.base ? &st.destroyed : nullptr);
    ctx->h_descs[st.slot] = d;
    CU(ctx, cudaMemcpyAsync(ctx->d_descs + st.slot, &d, sizeof(d), cudaMemcpyHostToDevice, ctx->stream));
    return FW_OK;
}

// wait for everything, read all stream states, make n_hi exact
int refresh_exact(fw_context *ctx) {
    // nothing was submitted since the last exact snapshot: it is still exact (lets a host loop
    // call fw_counts / fw_read_aabb per spawner without a sync + copy each time)
    if (ctx->snapshot_valid && ctx->snapshot.size() == ctx->n_slots) return FW_OK;
    // the last thing enqueued was a frame: its own asynchronous state readback (pinned) is the
    // exact state once that frame is done -- no second copy, no full-stream sync
    if (ctx->readback_is_current && ctx->frame_no > 0) {
        FrameSlot &ls = ctx->ring[(ctx->frame_no - 1) % kRing];
        if (ls.frame == ctx->frame_no && ls.states_slots >= ctx->n_slots) {
            CU(ctx, cudaEventSynchronize(ls.done));
            ctx->snapshot.assign(ls.states_host, ls.states_host + ctx->n_slots);
            for (FrameSlot &fs : ctx->ring) fs.in_flight = false; // stream order: older frames are done too
            for (uint32_t s = 0; s < ctx->n_slots; s++)
                if (Stream *st = ctx->slot_owner[s]) st->n_hi = ctx->snapshot[s].count - ctx->snapshot[s].dead;
            ctx->snapshot_valid = true;
            return FW_OK;
        }
    }
    CU(ctx, sync_all(ctx));
    for (FrameSlot &fs : ctx->ring) fs.in_flight = false;
    ctx->snapshot.resize(ctx->n_slots);
    if (ctx->n_slots)
        CU(ctx, cudaMemcpy(ctx->snapshot.data(), cur_states(ctx), sizeof(StreamState) * ctx->n_slots, cudaMemcpyDeviceToHost));
    for (uint32_t s = 0; s < ctx->n_slots; s++)
        if (Stream *st = ctx->slot_owner[s]) st->n_hi = ctx->snapshot[s].count - ctx->snapshot[s].dead;
    ctx->snapshot_valid = true;
    return FW_OK;
}

// consume finished asynchronous state readbacks to tighten the n_hi bounds without a sync
void poll_readbacks(fw_context *ctx) {
    uint64_t newest = 0;
    FrameSlot *best = nullptr;
    for (FrameSlot &fs : ctx->ring) {
        if (!fs.in_flight) continue;
        if (cudaEventQuery(fs.done) == cudaSuccess) {
            fs.in_flight = false;
            if (fs.plan_host) { // what the device flagged in that frame (reported by fw_sync / fw_poll_device_errors)
                ctx->device_error_flags |= fs.plan_host->error_flags;
                fs.plan_host->error_flags = 0;
            }
            if (fs.frame >= newest) { newest = fs.frame; best = &fs; }
        } else {
            (void)cudaGetLastError();
        }
    }
    if (!best) return;
    const uint32_t n = std::min(best->states_slots, ctx->n_slots);
    for (uint32_t s = 0; s < n; s++) {
        Stream *st = ctx->slot_owner[s];
        if (!st || st->born_frame > best->frame) continue;
        uint64_t bound = best->states_host[s].count - best->states_host[s].dead;
        for (const FrameSlot &g : ctx->ring)
            if (g.frame > best->frame && g.frame <= ctx->frame_no && s < g.spawn_per_slot.size()) bound += g.spawn_per_slot[s];
        if (bound < st->n_hi) st->n_hi = bound;
    }
}

int grow_stream(fw_context *ctx, Stream &st, uint64_t need) {
    // exact state is in ctx->snapshot (refresh_exact was just called)
    const StreamState s = ctx->snapshot[st.slot];
    const uint32_t live = s.count - s.dead;
    const uint32_t first = live_first(st.variant, s, st.block.capacity);
    const uint64_t slots = is_fifo(st.variant) ? need : 2 * need;
    const uint32_t ncap = round_capacity(std::max<uint64_t>(slots + slots / 4, (uint64_t)st.block.capacity * 2));
    if ((uint64_t)ncap < slots) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "stream would exceed 2^32 particles");
    Block nb;
    int rc = alloc_block(ctx, ncap, st.n_lea, nb);
    if (rc) return rc;
    if (st.destroyed.base) { // same capacity; its content only lives until the next frame
        release_block(ctx, st.destroyed);
        if ((rc = alloc_block(ctx, ncap, st.n_lea, st.destroyed))) return rc;
    }
    // unwrap the ring into the start of the new block
    ctx->readback_is_current = false;
    CU(ctx, launch_ring_copy(block_desc(st.block, st.variant, st.flags), first, live, block_desc(nb, st.variant, st.flags), ctx->stream));
    StreamState ns = s;
    const StreamState inj = injected_state(st.variant, live, ncap);
    ns.head = inj.head;
    ns.count = live;
    ns.dead = 0;
    CU(ctx, cudaMemcpyAsync(cur_states(ctx) + st.slot, &ns, sizeof(ns), cudaMemcpyHostToDevice, ctx->stream));
    ctx->snapshot[st.slot] = ns;
    ctx->tiles_needed -= max_tiles_of(st.block.capacity);
    release_block(ctx, st.block);
    st.block = nb;
    ctx->tiles_needed += max_tiles_of(ncap);
    return upload_desc(ctx, st);
}

int ensure_frame_slot(fw_context *ctx, FrameSlot &fs, size_t bytes) {
    if (bytes > fs.bytes) {
        size_t nb = std::max<size_t>(bytes * 2, 1 << 16);
        if (fs.host) CU(ctx, cudaFreeHost(fs.host));
        if (fs.dev) CU(ctx, cudaFree(fs.dev));
        fs.host = nullptr;
        fs.dev = nullptr;
        CU(ctx, cudaMallocHost((void **)&fs.host, nb));
        CU(ctx, cudaMalloc((void **)&fs.dev, nb));
        fs.bytes = nb;
        fs.graph_version = 0;
    }
    if (fs.states_slots < ctx->slots_cap || !fs.readback) {
        if (fs.readback) CU(ctx, cudaFreeHost(fs.readback));
        fs.readback = nullptr;
        CU(ctx, cudaMallocHost((void **)&fs.readback, statebuf_bytes(ctx->slots_cap)));
        memset(fs.readback, 0, statebuf_bytes(ctx->slots_cap));
        fs.plan_host = (PlanOut *)fs.readback;
        fs.states_host = (StreamState *)(fs.readback + sizeof(PlanOut));
        fs.states_slots = ctx->slots_cap;
        fs.graph_version = 0;
    }
    return FW_OK;
}

// account one completed frame: counts always, kernel times when it was a timed frame
void absorb_profile(fw_context *ctx, FrameSlot &fs) {
    if (!fs.pending_account) return;
    fs.pending_account = false;
    fw_frame_profile p{};
    if (fs.profiled) {
        cudaEventElapsedTime(&p.plan_ms, fs.ev[0], fs.ev[1]);
        cudaEventElapsedTime(&p.spawn_ms, fs.ev[1], fs.ev[2]);
        cudaEventElapsedTime(&p.update_ms, fs.ev[2], fs.ev[3]);
        cudaEventElapsedTime(&p.total_ms, fs.ev[0], fs.ev[3]);
        ctx->prof_timed_frames++;
        p.timed_frames = 1;
    }
    fs.profiled = false;
    p.kernel_launches = fs.launches;
    p.particles_spawned = fs.particles_spawned;
    p.h2d_bytes = fs.h2d_bytes;
    p.d2h_bytes = fs.d2h_bytes;
    p.particles_updated = fs.plan_host ? fs.plan_host->total_update : 0;
    ctx->prof_last = p;
    ctx->prof_sum.plan_ms += p.plan_ms;
    ctx->prof_sum.spawn_ms += p.spawn_ms;
    ctx->prof_sum.update_ms += p.update_ms;
    ctx->prof_sum.total_ms += p.total_ms;
    ctx->prof_sum.kernel_launches += p.kernel_launches;
    ctx->prof_sum.particles_spawned += p.particles_spawned;
    ctx->prof_sum.particles_updated += p.particles_updated;
    ctx->prof_sum.h2d_bytes += p.h2d_bytes;
    ctx->prof_sum.d2h_bytes += p.d2h_bytes;
    ctx->prof_sum.timed_frames = ctx->prof_timed_frames;
    ctx->prof_frames++;
}

// ParticleSpawnerData::active (reference src/core.rs:288-302); any_particles only matters for
// nested emitters
bool spawner_active(const Spawner &sp, bool any_particles) {
    bool enabled = false;
    for (const Emitter &e : sp.emitters) {
        if (e.emits_on_other_particles) enabled |= (e.enabled && any_particles);
        else enabled |= e.enabled;
    }
    return enabled;
}

// topology summary of the nested emitters: total and the number of phases a frame needs
void recount_nested(fw_context *ctx) {
    uint32_t total = 0, max_per = 0;
    for (auto &sp : ctx->spawners) {
        uint32_t n = 0;
        for (const Emitter &e : sp->emitters) n += e.emits_on_other_particles ? 1u : 0u;
        total += n;
        max_per = std::max(max_per, n);
    }
    ctx->live_nested = total;
    ctx->n_phases = 1 + max_per;
}

Spawner *find(fw_context *ctx, uint32_t key) {
    auto it = ctx->by_key.find(key);
    return it == ctx->by_key.end() ? nullptr : it->second;
}

// device staging buffer for ParticleData / ParticleInstance rows of one stream
int ensure_stage(fw_context *ctx, size_t bytes) {
    if (bytes <= ctx->stage_bytes) return FW_OK;
    CU(ctx, sync_all(ctx));
    if (ctx->d_stage) CU(ctx, cudaFree(ctx->d_stage));
    ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    const size_t nb = bytes + bytes / 4 + 4096;
    CU(ctx, cudaMalloc((void **)&ctx->d_stage, nb));
    ctx->stage_bytes = nb;
    return FW_OK;
}
// first live slot and live count of a stream from the exact snapshot
inline void live_range(const fw_context *ctx, const Stream &st, uint32_t &first, uint32_t &live) {
    const StreamState s = ctx->snapshot[st.slot];
    live = s.count - s.dead;
    first = live_first(st.variant, s, st.block.capacity);
}

inline float dec_f32(uint32_t u) {
    const uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

} // namespace

// =============================================================================== exports
extern "C" {

const char *fw_last_global_error(void) { return g_global_error.c_str(); }
const char *fw_last_error(const fw_context *ctx) { return ctx ? ctx->error.c_str() : g_global_error.c_str(); }
uint32_t fw_abi_version(void) { return FW_ABI_VERSION; }

uint32_t fw_abi_sizeof(const char *name) {
    if (!name) return 0;
#define FW_S(T) \
    if (!strcmp(name, #T)) return (uint32_t)sizeof(T);
#define FW_F(T, f)
#include "fw_abi_offsets.inc"
#undef FW_S
#undef FW_F
    return 0;
}

uint32_t fw_abi_offsetof(const char *struct_name, const char *field_name) {
    if (!struct_name || !field_name) return 0xFFFFFFFFu;
#define FW_S(T)
#define FW_F(T, f) \
    if (!strcmp(struct_name, #T) && !strcmp(field_name, #f)) return (uint32_t)offsetof(T, f);
#include "fw_abi_offsets.inc"
#undef FW_S
#undef FW_F
    return 0xFFFFFFFFu;
}

int fw_create(const fw_config *cfg, fw_context **out_ctx) {
    if (!cfg || !out_ctx) return fail(nullptr, FW_ERR_INVALID_ARGUMENT, "fw_create: null argument");
    *out_ctx = nullptr;
    if (cfg->abi_version != FW_ABI_VERSION) return fail(nullptr, FW_ERR_INVALID_ARGUMENT, "fw_create: ABI version %u, library is %u", cfg->abi_version, FW_ABI_VERSION);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, FW_ERR_NO_DEVICE, "fw_create: no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, FW_ERR_INVALID_ARGUMENT, "fw_create: device %d out of range (0..%d)", cfg->device, n_dev - 1);
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(nullptr, FW_ERR_NO_DEVICE, "fw_create: device %d is sm_%d%d; this build carries sm_100a code only", cfg->device, prop.major, prop.minor);
    // on any failure below fw_destroy releases whatever was created so far
    struct Guard {
        fw_context *c;
        ~Guard() { if (c) fw_destroy(c); }
    } ctx{new fw_context()};
    ctx.c->device = cfg->device;
    fw_context *c = ctx.c;
    c->concurrent_spawn = (cfg->flags & FW_FLAG_NO_CONCURRENT_SPAWN) == 0;
    c->seed = cfg->seed;
    c->flags = cfg->flags;
    c->use_graphs = (cfg->flags & FW_FLAG_NO_GRAPHS) == 0;
    c->profiling = (cfg->flags & FW_FLAG_PROFILE) != 0;
    CU(nullptr, cudaSetDevice(cfg->device));
    if (cfg->external_stream) {
        c->stream = (cudaStream_t)cfg->external_stream;
    } else {
        CU(nullptr, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->owns_stream = true;
    }
    CU(c, cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&c->rb_stream, cudaStreamNonBlocking));
    CU(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (FrameSlot &fs : c->ring) {
        CU(c, cudaEventCreateWithFlags(&fs.done, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&fs.ev_h2d, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&fs.ev_kernels, cudaEventDisableTiming));
        for (auto &ev : fs.ev) CU(c, cudaEventCreate(&ev));
    }
    int rc;
    if ((rc = ensure_slots(c, 1))) { g_global_error = c->error; return rc; }
    if ((rc = ensure_emitters(c, 1))) { g_global_error = c->error; return rc; }
    if ((rc = ensure_tiles(c))) { g_global_error = c->error; return rc; }
    CU(c, update_grid_size(c->device, c->grids, &c->team_size));
    *out_ctx = c;
    ctx.c = nullptr;
    return FW_OK;
}

static void gather_release(fw_context *ctx);
namespace {
void vmm_release(fw_context *ctx);
}
int fw_destroy(fw_context *ctx) {
    if (!ctx) return FW_OK;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    gather_release(ctx);
    for (auto &sp : ctx->spawners)
        for (Stream &st : sp->streams) {
            if (st.block.base) cudaFree(st.block.base);
            if (st.destroyed.base) cudaFree(st.destroyed.base);
        }
    for (auto &kv : ctx->block_cache)
        for (void *p : kv.second) cudaFree(p);
    for (FrameSlot &fs : ctx->ring) {
        if (fs.host) cudaFreeHost(fs.host);
        if (fs.dev) cudaFree(fs.dev);
        if (fs.readback) cudaFreeHost(fs.readback);
        if (fs.graph_exec) cudaGraphExecDestroy(fs.graph_exec);
        if (fs.done) cudaEventDestroy(fs.done);
        if (fs.ev_h2d) cudaEventDestroy(fs.ev_h2d);
        if (fs.ev_kernels) cudaEventDestroy(fs.ev_kernels);
        for (auto &ev : fs.ev)
            if (ev) cudaEventDestroy(ev);
    }
    cudaFree(ctx->d_descs);
    cudaFree(ctx->d_statebuf[0]);
    cudaFree(ctx->d_statebuf[1]);
    cudaFree(ctx->d_settings);
    cudaFree(ctx->d_emitters);
    cudaFree(ctx->d_colliders);
    cudaFree(ctx->d_broadphase);
    for (auto &h : ctx->h_col_stage)
        if (h) cudaFreeHost(h);
    for (auto &ev : ctx->col_stage_ev)
        if (ev) cudaEventDestroy(ev);
    cudaFree(ctx->d_tile_prefix);
    cudaFree(ctx->d_nested_scratch);
    cudaFree(ctx->d_nested_serial);
    cudaFree(ctx->d_nested_out);
    cudaFree(ctx->d_stage);
    cudaFree(ctx->d_lookback);
    cudaFree(ctx->d_pack);
    cudaFree(ctx->d_extract);
    if (ctx->xt_stream) cudaStreamSynchronize(ctx->xt_stream);
    for (auto &x : ctx->xt) {
        if (x.d_rows) cudaFree(x.d_rows);
        if (x.d_offsets) cudaFree(x.d_offsets);
        if (x.d_slots) cudaFree(x.d_slots);
        if (x.h_offsets) cudaFreeHost(x.h_offsets);
        if (x.packed) cudaEventDestroy(x.packed);
        if (x.landed) cudaEventDestroy(x.landed);
    }
    if (ctx->xt_stream) cudaStreamDestroy(ctx->xt_stream);
    vmm_release(ctx);
    for (auto &ev : ctx->user_events)
        if (ev) cudaEventDestroy(ev);
    if (ctx->h_pack) cudaFreeHost(ctx->h_pack);
    if (ctx->side_stream) {
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamDestroy(ctx->side_stream);
    }
    if (ctx->rb_stream) {
        cudaStreamSynchronize(ctx->rb_stream);
        cudaStreamDestroy(ctx->rb_stream);
    }
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FW_OK;
}

int fw_spawner_reset(fw_context *ctx, uint32_t key, const fw_particle_settings *ps, uint32_t n_types,
                     const fw_emission_settings *es, uint32_t n_emitters, uint32_t starts_enabled) {
    ENTER(ctx);
    if ((n_types && !ps) || (n_emitters && !es)) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_spawner_reset: null settings");
    // validate before touching any state (the reference panics at curve construction,
    // src/curve.rs:45,50,61,67; here the call fails and leaves the spawner as it was)
    for (uint32_t i = 0; i < n_types; i++) {
        int rc;
        if ((rc = validate_curve(ctx, ps[i].scale_curve.kind, ps[i].scale_curve.n, ps[i].scale_curve.times, "scale_curve"))) return rc;
        if ((rc = validate_curve(ctx, ps[i].base_color.kind, ps[i].base_color.n, ps[i].base_color.times, "base_color"))) return rc;
        if ((rc = validate_curve(ctx, ps[i].emissive_color.kind, ps[i].emissive_color.n, ps[i].emissive_color.times, "emissive_color"))) return rc;
    }
    for (uint32_t i = 0; i < n_emitters; i++) {
        if (es[i].particle_index >= n_types) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "emitter %u: particle_index %u out of range", i, es[i].particle_index);
        if (es[i].pacing_kind > FW_PACING_COUNT_OVER_DURATION) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "emitter %u: unknown pacing", i);
        if (es[i].shape_kind > FW_SHAPE_CIRCLE) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "emitter %u: unknown shape", i);
        if (es[i].mode > FW_MODE_NESTED) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "emitter %u: unknown emission mode", i);
        if (es[i].mode == FW_MODE_NESTED && es[i].target_particle_type >= n_types)
            return fail(ctx, FW_ERR_INVALID_ARGUMENT, "emitter %u: target_particle_type %u out of range", i, es[i].target_particle_type);
    }
    {
        uint32_t n_nested = 0;
        std::vector<uint32_t> lea_per_type(n_types, 0);
        for (uint32_t i = 0; i < n_emitters; i++)
            if (es[i].mode == FW_MODE_NESTED) {
                n_nested++;
                if (++lea_per_type[es[i].target_particle_type] > kMaxLea)
                    return fail(ctx, FW_ERR_UNSUPPORTED, "more than %u nested emitters target particle type %u", kMaxLea, es[i].target_particle_type);
            }
        if (n_nested + 1 > kMaxPhases) return fail(ctx, FW_ERR_UNSUPPORTED, "more than %u nested emitters in one spawner", kMaxPhases - 1);
    }
    Spawner *sp = find(ctx, key);
    if (!sp) {
        ctx->spawners.emplace_back(new Spawner());
        sp = ctx->spawners.back().get();
        sp->key = key;
        memset(&sp->input, 0, sizeof(sp->input));
        sp->input.rotation[3] = 1.0f;
        sp->input.modifier_scale = 1.0f;
        sp->input.modifier_speed = 1.0f;
        ctx->by_key[key] = sp;
    } else {
        free_spawner_resources(ctx, *sp); // data.particles = vec![Vec::new(); n]  (src/core.rs:360)
    }
    // everything below mutates the spawner; on a failure (out of memory, CUDA error) the
    // half-built spawner is removed again instead of being left inconsistent
    auto build = [&]() -> int {
        sp->initialized = true; // :361-363
        topo_changed(ctx);
        ctx->live_emitters += n_emitters;
        sp->emitters.resize(n_emitters);
        for (uint32_t i = 0; i < n_emitters; i++) { // :347-359
            Emitter &e = sp->emitters[i];
            e.es = es[i];
            e.last_emission = 0.f;
            e.time_passed_in_cycle = 0.f;
            e.enabled = starts_enabled != 0;
            e.emits_on_other_particles = es[i].mode == FW_MODE_NESTED;
            e.serial = 0;
            e.lea_index = 0;
            if (e.emits_on_other_particles)
                for (uint32_t k = 0; k < i; k++)
                    if (es[k].mode == FW_MODE_NESTED && es[k].target_particle_type == es[i].target_particle_type) e.lea_index++;
            if (!ctx->free_emitters.empty()) {
                e.dev_idx = ctx->free_emitters.back();
                ctx->free_emitters.pop_back();
            } else {
                int rc = ensure_emitters(ctx, ctx->n_emitters + 1);
                if (rc) return rc;
                e.dev_idx = ctx->n_emitters++;
            }
            CU(ctx, cudaMemcpyAsync(ctx->d_emitters + e.dev_idx, &es[i], sizeof(fw_emission_settings), cudaMemcpyHostToDevice, ctx->stream));
            CU(ctx, cudaMemsetAsync(ctx->d_nested_serial + e.dev_idx, 0, sizeof(unsigned long long), ctx->stream));
            ctx->h_emitters[e.dev_idx] = es[i];
        }
        sp->streams.resize(n_types);
        for (uint32_t t = 0; t < n_types; t++) {
            Stream &st = sp->streams[t];
            st.type = t;
            st.ps = ps[t];
            fill_dev_settings(ps[t], st.dev);
            stream_proofs(ps[t], t, es, n_emitters, st.variant, st.flags, st.dev);
            st.n_hi = 0;
            st.born_frame = ctx->frame_no + 1;
            st.injected = false;
            st.accounted = false;
            st.n_lea = 0;
            for (uint32_t i = 0; i < n_emitters; i++)
                if (es[i].mode == FW_MODE_NESTED && es[i].target_particle_type == t) st.n_lea++;
        }
        for (uint32_t t = 0; t < n_types; t++) {
            Stream &st = sp->streams[t];
            if (!ctx->free_slots.empty()) {
                st.slot = ctx->free_slots.back();
                ctx->free_slots.pop_back();
            } else {
                int rc = ensure_slots(ctx, ctx->n_slots + 1);
                if (rc) return rc;
                st.slot = ctx->n_slots++;
            }
            int rc = alloc_block(ctx, round_capacity(estimate_capacity(*sp, t) * (is_fifo(st.variant) ? 1 : 2)), st.n_lea, st.block);
            if (rc) {
                ctx->free_slots.push_back(st.slot);
                return rc;
            }
            ctx->tiles_needed += max_tiles_of(st.block.capacity);
            ctx->variant_streams[st.variant]++;
            st.accounted = true;
            ctx->slot_owner[st.slot] = &st;
            if (ps[t].capture_destroyed && (rc = alloc_block(ctx, st.block.capacity, st.n_lea, st.destroyed))) return rc;
            CU(ctx, cudaMemcpyAsync(ctx->d_settings + st.slot, &st.dev, sizeof(st.dev), cudaMemcpyHostToDevice, ctx->stream));
            CU(ctx, cudaMemsetAsync(cur_states(ctx) + st.slot, 0, sizeof(StreamState), ctx->stream));
            if ((rc = upload_desc(ctx, st))) return rc;
        }
        sp->finished_notified = false;
        sp->manual_queued_count = 0;

        return FW_OK;
    };
    {
        const int rc = build();
        if (rc != FW_OK) {
            const std::string msg = ctx->error;
            free_spawner_resources(ctx, *sp);
            ctx->by_key.erase(key);
            for (size_t i = 0; i < ctx->spawners.size(); i++)
                if (ctx->spawners[i].get() == sp) {
                    ctx->spawners.erase(ctx->spawners.begin() + (long)i);
                    break;
                }
            recount_nested(ctx);
            ctx->error = msg;
            return rc;
        }
    }
    recount_nested(ctx);
    return ensure_tiles(ctx);
}

int fw_spawner_remove(fw_context *ctx, uint32_t key) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_spawner_remove: unknown spawner %u", key);
    free_spawner_resources(ctx, *sp);
    ctx->by_key.erase(key);
    for (size_t i = 0; i < ctx->spawners.size(); i++)
        if (ctx->spawners[i].get() == sp) {
            ctx->spawners.erase(ctx->spawners.begin() + (long)i);
            break;
        }
    recount_nested(ctx);
    return FW_OK;
}

// Collider BVH in depth-first order: node k = (nodes[2k] = min.xyz | skip link, nodes[2k+1] =
// max.xyz | collider index or 0xFFFFFFFF for an inner node); the skip link is the first node
// after k's subtree (a leaf stores its layers bits there instead: its skip link is k + 1). A non-finite bound (NaN transform) becomes the whole line: it never culls.
static inline float bvh_sane(float v, float fallback) { return std::isfinite(v) ? v : fallback; }
static void build_bvh(std::vector<float4> &nodes, std::vector<uint32_t> &order, const std::vector<float> &lo,
                      const std::vector<float> &hi, const std::vector<uint32_t> &layers, uint32_t begin, uint32_t end) {
    float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    auto centre = [&](uint32_t i, int a) { return 0.5f * bvh_sane(lo[3 * i + a], -FLT_MAX) + 0.5f * bvh_sane(hi[3 * i + a], FLT_MAX); };
    for (uint32_t k = begin; k < end; k++) {
        const uint32_t i = order[k];
        for (int a = 0; a < 3; a++) {
            blo[a] = std::min(blo[a], bvh_sane(lo[3 * i + a], -FLT_MAX));
            bhi[a] = std::max(bhi[a], bvh_sane(hi[3 * i + a], FLT_MAX));
            clo[a] = std::min(clo[a], centre(i, a));
            chi[a] = std::max(chi[a], centre(i, a));
        }
    }
    const size_t me = nodes.size();
    nodes.push_back(make_float4(blo[0], blo[1], blo[2], 0.f));
    nodes.push_back(make_float4(bhi[0], bhi[1], bhi[2], 0.f));
    uint32_t leaf = 0xFFFFFFFFu;
    if (end - begin == 1) {
        leaf = order[begin];
    } else {
        int axis = 0;
        for (int a = 1; a < 3; a++)
            if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
        const uint32_t mid = begin + (end - begin) / 2;
        std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end, [&](uint32_t p, uint32_t q) {
            const float cp = centre(p, axis), cq = centre(q, axis);
            return cp < cq || (cp == cq && p < q);
        });
        build_bvh(nodes, order, lo, hi, layers, begin, mid);
        build_bvh(nodes, order, lo, hi, layers, mid, end);
    }
    const uint32_t skip = (uint32_t)(nodes.size() / 2);
    const uint32_t link = leaf == 0xFFFFFFFFu ? skip : layers[leaf]; // a leaf's skip link is k + 1
    memcpy(&nodes[me].w, &link, 4);
    memcpy(&nodes[me + 1].w, &leaf, 4);
}

// The whole broad phase of a collider set as one blob (layout: BroadPhaseHeader, fw_internal.h).
static void build_broadphase(const fw_collider *colliders, uint32_t n, std::vector<uint8_t> &blob) {
    // world AABB of every collider, inflated well beyond fp32 rounding of the exact ray tests
    std::vector<float> lo(3 * (size_t)n), hi(3 * (size_t)n);
    std::vector<uint32_t> layers(n);
    for (uint32_t i = 0; i < n; i++) {
        const fw_collider &c = colliders[i];
        layers[i] = c.layers;
        const double x = c.rotation[0], y = c.rotation[1], z = c.rotation[2], w = c.rotation[3];
        const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)},
                                {2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)},
                                {2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
        for (int a = 0; a < 3; a++) {
            // local half extents: cylinder / cone = (r, h, r), capsule = (r, h + r, r)
            const bool about_y = c.kind == FW_COLLIDER_CYLINDER || c.kind == FW_COLLIDER_CONE || c.kind == FW_COLLIDER_CAPSULE;
            const double hz = about_y ? c.half_extents[0] : c.half_extents[2];
            const double hy = c.kind == FW_COLLIDER_CAPSULE ? std::fabs((double)c.half_extents[1]) + std::fabs((double)c.half_extents[0])
                                                            : (double)c.half_extents[1];
            double ext = c.kind == FW_COLLIDER_SPHERE
                             ? std::fabs((double)c.half_extents[0])
                             : std::fabs(R[a][0] * c.half_extents[0]) + std::fabs(R[a][1] * hy) + std::fabs(R[a][2] * hz);
            const double margin = 1e-3 + 1e-3 * (std::fabs((double)c.translation[a]) + ext);
            // a non-finite bound (NaN transform) becomes the whole line: it never culls
            lo[3 * i + a] = bvh_sane((float)(c.translation[a] - ext - margin), -FLT_MAX);
            hi[3 * i + a] = bvh_sane((float)(c.translation[a] + ext + margin), FLT_MAX);
        }
    }
    // BVH: median split of the centroids along the widest axis, one collider per leaf, depth-first
    // order with skip links, so that the kernel walks it without a stack
    std::vector<float4> nodes;
    nodes.reserve(4 * (size_t)n);
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    build_bvh(nodes, order, lo, hi, layers, 0, n);

    // uniform grid over the colliders that are small against the scene; the others ("big": the
    // ground slab, anything unbounded) are tested for every ray
    float glo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ghi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    std::vector<double> diag(n);
    for (uint32_t i = 0; i < n; i++) {
        double d2 = 0;
        for (int a = 0; a < 3; a++) d2 += ((double)hi[3 * i + a] - lo[3 * i + a]) * ((double)hi[3 * i + a] - lo[3 * i + a]);
        diag[i] = std::sqrt(d2);
    }
    std::vector<double> sorted_diag(diag);
    std::sort(sorted_diag.begin(), sorted_diag.end());
    const double median_diag = sorted_diag[n / 2];
    std::vector<uint32_t> big, small;
    for (uint32_t i = 0; i < n; i++) {
        if (!(diag[i] <= 8.0 * median_diag) || !std::isfinite(diag[i])) big.push_back(i);
        else small.push_back(i);
    }
    for (uint32_t i : small)
        for (int a = 0; a < 3; a++) {
            glo[a] = std::min(glo[a], lo[3 * i + a]);
            ghi[a] = std::max(ghi[a], hi[3 * i + a]);
        }
    BroadPhaseHeader h{};
    std::vector<uint32_t> cell_start, items;
    bool use_grid = small.size() >= 8 && big.size() <= 16;
    float inv_cell[3] = {0, 0, 0};
    if (use_grid) {
        // cells about as large as a typical collider box, at most cells_cap of them
        double edge = std::max(median_diag / std::sqrt(3.0), 1e-6);
        const size_t cells_cap = broadphase_cells_cap(n), items_cap = broadphase_items_cap(n);
        for (int attempt = 0; attempt < 24; attempt++, edge *= 1.3) {
            size_t cells = 1;
            for (int a = 0; a < 3; a++) {
                const double ext = std::max((double)ghi[a] - glo[a], 1e-6);
                h.dim[a] = (uint32_t)std::min<double>(kGridMaxDim, std::max(1.0, std::ceil(ext / edge)));
                inv_cell[a] = (float)(h.dim[a] / ext);
                cells *= h.dim[a];
            }
            if (cells > cells_cap) continue;
            // count, then fill (CSR). Cell coordinates in fp32 with the kernel's own expression.
            auto cell_of = [&](float v, int a) {
                const float f = std::floor((v - glo[a]) * inv_cell[a]);
                return (int)std::min<float>(std::max<float>(f, 0.0f), (float)(h.dim[a] - 1));
            };
            cell_start.assign(cells + 1, 0);
            size_t total = 0;
            for (int pass = 0; pass < 2 && total <= items_cap; pass++) {
                if (pass == 1) {
                    uint32_t run = 0;
                    for (size_t c = 0; c < cells; c++) {
                        const uint32_t cnt = cell_start[c];
                        cell_start[c] = run;
                        run += cnt;
                    }
                    cell_start[cells] = run;
                    items.assign(run, 0);
                }
                std::vector<uint32_t> fill(pass == 1 ? cells : 0, 0);
                for (uint32_t i : small) {
                    int c0[3], c1[3];
                    for (int a = 0; a < 3; a++) {
                        c0[a] = cell_of(lo[3 * i + a], a);
                        c1[a] = cell_of(hi[3 * i + a], a);
                    }
                    for (int z = c0[2]; z <= c1[2]; z++)
                        for (int y = c0[1]; y <= c1[1]; y++)
                            for (int x = c0[0]; x <= c1[0]; x++) {
                                const size_t c = ((size_t)z * h.dim[1] + y) * h.dim[0] + x;
                                if (pass == 0) {
                                    cell_start[c]++;
                                    total++;
                                } else {
                                    items[cell_start[c] + fill[c]++] = i;
                                }
                            }
                }
            }
            if (total <= items_cap) break;
            cell_start.clear();
        }
        use_grid = !cell_start.empty();
    }
    if (!use_grid) {
        cell_start.assign(2, 0);
        items.clear();
        big.clear();
        h.dim[0] = h.dim[1] = h.dim[2] = 1;
    }
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    h.n_nodes = (uint32_t)(nodes.size() / 2);
    h.leaf_off = (uint32_t)sizeof(BroadPhaseHeader);
    h.nodes_off = (uint32_t)(h.leaf_off + 32 * (size_t)n);
    h.big_off = (uint32_t)(h.nodes_off + 16 * nodes.size());
    h.n_big = (uint32_t)big.size();
    h.cell_off = (uint32_t)align16(h.big_off + 4 * big.size());
    h.items_off = (uint32_t)align16(h.cell_off + 4 * cell_start.size());
    h.use_grid = use_grid ? 1u : 0u;
    for (int a = 0; a < 3; a++) {
        h.lo[a] = use_grid ? glo[a] : 0.f;
        h.inv_cell[a] = inv_cell[a];
    }
    blob.assign(align16(h.items_off + 4 * items.size()), 0);
    memcpy(blob.data(), &h, sizeof(h));
    float4 *leaf = (float4 *)(blob.data() + h.leaf_off);
    for (uint32_t i = 0; i < n; i++) {
        float lb;
        memcpy(&lb, &layers[i], 4);
        leaf[2 * i] = make_float4(lo[3 * i], lo[3 * i + 1], lo[3 * i + 2], lb);
        leaf[2 * i + 1] = make_float4(hi[3 * i], hi[3 * i + 1], hi[3 * i + 2], 0.f);
    }
    memcpy(blob.data() + h.nodes_off, nodes.data(), 16 * nodes.size());
    if (!big.empty()) memcpy(blob.data() + h.big_off, big.data(), 4 * big.size());
    memcpy(blob.data() + h.cell_off, cell_start.data(), 4 * cell_start.size());
    if (!items.empty()) memcpy(blob.data() + h.items_off, items.data(), 4 * items.size());
}

int fw_host_emission_count(float time_passed_in_cycle, float last_emission, float cycle_duration, float offset_start,
                           float offset_end, float particles_per_cycle, uint64_t *times, float *next_last_emission) {
    uint64_t n = 0;
    float next = 0.f;
    compute_emission_count(time_passed_in_cycle, last_emission, cycle_duration, offset_start, offset_end, particles_per_cycle, n, next);
    if (times) *times = n;
    if (next_last_emission) *next_last_emission = next;
    return FW_OK;
}

int fw_host_build_broadphase(const fw_collider *colliders, uint32_t n, void *out, uint64_t cap_bytes, uint64_t *n_bytes) {
    if (n && !colliders) return FW_ERR_INVALID_ARGUMENT;
    for (uint32_t i = 0; i < n; i++)
        if (colliders[i].kind > FW_COLLIDER_CAPSULE) return FW_ERR_UNSUPPORTED;
    std::vector<uint8_t> blob;
    if (n) build_broadphase(colliders, n, blob);
    if (n_bytes) *n_bytes = blob.size();
    if (blob.size() > broadphase_bytes(n)) return FW_ERR_INTERNAL; // the device buffer is sized by this bound
    if (!out || cap_bytes < blob.size()) return blob.empty() ? FW_OK : FW_ERR_BUFFER_TOO_SMALL;
    if (!blob.empty()) memcpy(out, blob.data(), blob.size());
    return FW_OK;
}

int fw_set_colliders(fw_context *ctx, const fw_collider *colliders, uint32_t n) {
    ENTER(ctx);
    if (n && !colliders) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_set_colliders: null");
    for (uint32_t i = 0; i < n; i++)
        if (colliders[i].kind > FW_COLLIDER_CAPSULE) return fail(ctx, FW_ERR_UNSUPPORTED, "collider %u: only cuboids, spheres, cylinders, cones and capsules are supported", i);
    // Same collider count as before (moving colliders, re-sent every physics step): same buffers,
    // same kernel arguments, the copies below are ordered on the context's stream between the
    // frames around them -- no synchronisation, captured frame graphs stay valid. A different
    // count reallocates.
    bool revolved = false;
    for (uint32_t i = 0; i < n; i++) revolved = revolved || colliders[i].kind >= FW_COLLIDER_CYLINDER;
    if (revolved != ctx->colliders_revolved) { // other kernels from the next frame on: captured graphs are stale
        ctx->colliders_revolved = revolved;
        topo_changed(ctx);
    }
    const bool same_shape = n == ctx->n_colliders && (n == 0 || (ctx->d_colliders && ctx->d_broadphase));
    if (!same_shape) {
        CU(ctx, sync_all(ctx));
        if (ctx->d_colliders) CU(ctx, cudaFree(ctx->d_colliders));
        ctx->d_colliders = nullptr;
        if (ctx->d_broadphase) CU(ctx, cudaFree(ctx->d_broadphase));
        ctx->d_broadphase = nullptr;
        ctx->n_colliders = n;
        topo_changed(ctx);
        if (n) {
            CU(ctx, cudaMalloc((void **)&ctx->d_colliders, sizeof(fw_collider) * n));
            CU(ctx, cudaMalloc((void **)&ctx->d_broadphase, broadphase_bytes(n)));
        }
    }
    if (n) {
        std::vector<uint8_t> blob;
        build_broadphase(colliders, n, blob);
        // both arrays through a pinned staging slot: the caller's array is free on return and the
        // host never waits for the stream (a pageable source would)
        const size_t col_bytes = sizeof(fw_collider) * n;
        const size_t blob_off = (col_bytes + 15) & ~(size_t)15;
        if (blob_off + blob.size() > ctx->col_stage_bytes) {
            CU(ctx, sync_all(ctx));
            for (auto &hp : ctx->h_col_stage) {
                if (hp) CU(ctx, cudaFreeHost(hp));
                hp = nullptr;
            }
            ctx->col_stage_bytes = blob_off + broadphase_bytes(n);
            for (auto &hp : ctx->h_col_stage) CU(ctx, cudaMallocHost((void **)&hp, ctx->col_stage_bytes));
        }
        const uint32_t k = ctx->col_stage_next++ % 4u;
        if (!ctx->col_stage_ev[k]) CU(ctx, cudaEventCreateWithFlags(&ctx->col_stage_ev[k], cudaEventDisableTiming));
        else CU(ctx, cudaEventSynchronize(ctx->col_stage_ev[k])); // the copy that last used this slot
        memcpy(ctx->h_col_stage[k], colliders, col_bytes);
        memcpy(ctx->h_col_stage[k] + blob_off, blob.data(), blob.size());
        CU(ctx, cudaMemcpyAsync(ctx->d_colliders, ctx->h_col_stage[k], col_bytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->d_broadphase, ctx->h_col_stage[k] + blob_off, blob.size(), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaEventRecord(ctx->col_stage_ev[k], ctx->stream));
    }
    return FW_OK;
}

int fw_frame(fw_context *ctx, float dt, const fw_spawner_frame_input *inputs, uint32_t n_inputs) {
    ENTER(ctx);
    if (n_inputs && !inputs) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_frame: null inputs");
    // dt is Res<Time>::delta_secs() (src/core.rs:413,594): a Duration as f32, finite and >= 0. The
    // per-stream constants of static streams rely on that (fw_internal.h), so anything else is refused.
    if (!std::isfinite(dt) || std::signbit(dt)) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_frame: dt = %g (must be finite and >= 0)", (double)dt);
    // resolve every key before touching any state (an unknown key fails the call and changes nothing).
    // A shim sends its spawners in the same order every tick -- usually creation order --, so the
    // spawner at the same position is tried before the hash map.
    std::vector<Spawner *> &targets = ctx->fr_targets;
    targets.resize(n_inputs);
    for (uint32_t k = 0; k < n_inputs; k++) {
        Spawner *sp = k < ctx->spawners.size() && ctx->spawners[k]->key == inputs[k].spawner_key ? ctx->spawners[k].get()
                                                                                              : find(ctx, inputs[k].spawner_key);
        if (!sp) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_frame: unknown spawner %u", inputs[k].spawner_key);
        targets[k] = sp;
    }
    for (uint32_t k = 0; k < n_inputs; k++) {
        Spawner *sp = targets[k];
        memcpy(sp->input.translation, inputs[k].origin_translation, sizeof(float) * 3);
        memcpy(sp->input.rotation, inputs[k].origin_rotation, sizeof(float) * 4);
        memcpy(sp->input.parent_velocity, inputs[k].parent_velocity, sizeof(float) * 3);
        sp->input.modifier_scale = inputs[k].modifier_scale;
        sp->input.modifier_speed = inputs[k].modifier_speed;
        sp->manual_queued_count += inputs[k].queue_particles; // src/core.rs:284-286
    }
    poll_readbacks(ctx);

    FrameSlot &fs = ctx->ring[ctx->frame_no % kRing];
    if (fs.in_flight) { // the host is kRing frames ahead: wait for that slot
        CU(ctx, cudaEventSynchronize(fs.done));
        fs.in_flight = false;
    }
    absorb_profile(ctx, fs);

    // ---- spawn_particles, host part (reference src/core.rs:377-428). The emitters of a spawner
    // are walked in order; its k-th Nested emitter closes phase k (see PhaseInfo).
    const uint32_t n_phases = ctx->n_phases;
    const uint32_t n_slots = ctx->n_slots;
    // Emission pacing is advanced while the frame is planned, but a later step can still fail (a
    // ring that cannot grow, a CUDA error): the emitters' and spawners' pacing state is then put back,
    // so that a failed fw_frame leaves everything as it was ("state untouched").
    struct UndoGuard {
        std::vector<fw_context::PacingUndo> &log;
        bool committed = false;
        ~UndoGuard() {
            if (committed) return;
            for (size_t k = log.size(); k-- > 0;) {
                const fw_context::PacingUndo &u = log[k];
                u.e->last_emission = u.last_emission;
                u.e->time_passed_in_cycle = u.time_passed_in_cycle;
                u.e->enabled = u.enabled;
                u.e->serial = u.serial;
                u.sp->manual_queued_count = u.manual_queued_count;
            }
        }
    } undo{ctx->fr_undo};
    undo.log.clear();
    // (scratch kept in the context: a steady-state frame allocates nothing)
    std::vector<SpawnCmd> *cmds = ctx->fr_cmds;
    std::vector<NestedCmd> *nested = ctx->fr_nested;
    for (uint32_t p = 0; p < kMaxPhases; p++) {
        cmds[p].clear();
        nested[p].clear();
    }
    std::vector<uint32_t> &spawn_per_slot = ctx->fr_spawn_per_slot; // [phase][slot]
    spawn_per_slot.assign((size_t)n_phases * std::max(1u, n_slots), 0u);
    uint64_t phase_total[kMaxPhases] = {0};
    std::vector<SpawnerInput> &sinputs = ctx->fr_inputs;
    sinputs.clear();
    for (auto &spp : ctx->spawners) {
        Spawner &sp = *spp;
        if (!spawner_active(sp, true)) continue; // :378 (a nested emitter without parents emits nothing anyway)
        int input_idx = -1;
        auto need_input = [&]() {
            if (input_idx < 0) {
                input_idx = (int)sinputs.size();
                sinputs.push_back(sp.input);
            }
            return (uint32_t)input_idx;
        };
        uint32_t phase = 0;
        for (uint32_t i = 0; i < sp.emitters.size(); i++) { // :386
            Emitter &e = sp.emitters[i];
            if (e.emits_on_other_particles) { // EmissionMode::Nested (:471-546)
                const uint32_t my_phase = phase++;
                if (!e.enabled) continue;                                        // :388-390
                if (e.es.pacing_kind != FW_PACING_COUNT_OVER_DURATION) continue; // :474-485 (warn_once! + skip)
                Stream &parent = sp.streams[e.es.target_particle_type];
                Stream &child = sp.streams[e.es.particle_index];
                NestedCmd c{};
                c.parent_stream = parent.slot;
                c.child_stream = child.slot;
                c.emitter_idx = e.dev_idx;
                c.emitter_local = i;
                c.spawner_key = sp.key;
                c.lea_index = e.lea_index;
                c.input_idx = need_input();
                nested[my_phase].push_back(c);
                continue;
            }
            if (!e.enabled) continue; // :388-390
            undo.log.push_back({&e, &sp, e.last_emission, e.time_passed_in_cycle, e.enabled, e.serial, sp.manual_queued_count});
            uint64_t n = 0;
            if (e.es.pacing_kind == FW_PACING_ONE_SHOT) { // :397-400
                e.enabled = false;
                n = e.es.one_shot_count;
            } else if (e.es.pacing_kind == FW_PACING_ON_DEMAND) { // :401-405
                n = sp.manual_queued_count;
                sp.manual_queued_count = 0;
            } else { // :406-427
                e.time_passed_in_cycle = rem_euclid_f32(e.time_passed_in_cycle + dt, e.es.duration);
                float next_last;
                compute_emission_count(e.time_passed_in_cycle, e.last_emission, e.es.duration, e.es.offset_start,
                                       e.es.offset_end, e.es.count, n, next_last);
                e.last_emission = next_last;
            }
            if (n == 0) continue;
            Stream &st = sp.streams[e.es.particle_index];
            if (n > 0xFFFFFF00ull - st.n_hi) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "spawner %u would exceed 2^32 particles in one stream", sp.key);
            uint32_t &slot_spawn = spawn_per_slot[(size_t)phase * n_slots + st.slot];
            SpawnCmd c{};
            c.stream = st.slot;
            c.emitter_idx = e.dev_idx;
            c.input_idx = need_input();
            c.count = (uint32_t)n;
            c.first = (uint32_t)phase_total[phase];
            c.dst_off = slot_spawn;
            c.spawner_key = sp.key;
            c.emitter_local = i;
            c.serial_base = e.serial;
            cmds[phase].push_back(c);
            e.serial += n;
            slot_spawn += (uint32_t)n;
            phase_total[phase] += n;
            if (phase_total[phase] > 0xFFFFFF00ull) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "more than 2^32 particles spawned in one frame");
        }
    }
    // ---- capacity planning: a ring is grown before it could overflow. Global spawn counts are
    // exact; what a Nested emitter will emit is only known on the device, so its child stream is
    // charged an upper bound (parents x per-parent cap; the count kernel enforces the cap and
    // raises an error flag if the reference would have emitted more). Bounds first, the exact
    // device state only if a bound does not fit.
    std::vector<uint64_t> &add = ctx->fr_add;
    add.assign(std::max(1u, n_slots), 0);
    // a parent's emission count covers the age it gained in the PREVIOUS frame's update
    // (compute_emission_count runs before this frame's update, src/core.rs:490-500)
    const float dt_bound = std::fmax(dt, ctx->prev_dt);
    auto plan_bounds = [&]() {
        std::fill(add.begin(), add.end(), 0);
        uint64_t scratch = 0;
        for (uint32_t p = 0; p < n_phases; p++) {
            for (uint32_t s = 0; s < n_slots; s++) add[s] += spawn_per_slot[(size_t)p * n_slots + s];
            for (NestedCmd &c : nested[p]) {
                const Stream *parent = ctx->slot_owner[c.parent_stream];
                const fw_emission_settings &es = ctx->h_emitters[c.emitter_idx];
                const float life_min = std::fmin(parent->ps.lifetime.min, parent->ps.lifetime.max);
                const double span = std::fmax((double)es.offset_end - (double)es.offset_start, 1e-9);
                double per = life_min > 0.f ? std::floor((double)es.count * ((double)dt_bound / life_min) / span) * 2.0 + 4.0
                                            : std::ceil((double)es.count) + 4.0;
                if (parent->injected) per = std::fmax(per, std::ceil((double)es.count) + 4.0);
                per = std::fmin(std::fmax(per, 1.0), 1048576.0);
                c.per_parent_cap = (uint32_t)per;
                const uint64_t parents = std::min<uint64_t>(parent->n_hi + add[c.parent_stream], 0xFFFFFF00ull);
                c.scratch_off = (uint32_t)scratch;
                scratch += parents + 1;
                add[c.child_stream] += parents * c.per_parent_cap;
            }
        }
        return scratch;
    };
    uint64_t scratch_need = plan_bounds();
    auto any_overflow = [&]() {
        for (uint32_t s = 0; s < n_slots; s++)
            if (add[s] && ctx->slot_owner[s] &&
                ctx->slot_owner[s]->n_hi + add[s] > usable_capacity(ctx->slot_owner[s]->variant, ctx->slot_owner[s]->block.capacity))
                return true;
        return false;
    };
    if (any_overflow()) {
        int rc = refresh_exact(ctx);
        if (rc) return rc;
        scratch_need = plan_bounds();
        for (uint32_t s = 0; s < n_slots; s++) {
            Stream *st = ctx->slot_owner[s];
            if (!st || !add[s]) continue;
            const uint64_t need = st->n_hi + add[s];
            if (need > 0xFFFFFF00ull) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "a stream would exceed 2^32 particles");
            if (need > usable_capacity(st->variant, st->block.capacity) && (rc = grow_stream(ctx, *st, need))) return rc;
        }
    }
    if (scratch_need > 0xFFFFFF00ull) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "nested emission scratch too large");
    if (scratch_need > ctx->nested_scratch_cap) {
        CU(ctx, sync_all(ctx));
        if (ctx->d_nested_scratch) CU(ctx, cudaFree(ctx->d_nested_scratch));
        ctx->d_nested_scratch = nullptr;
        ctx->nested_scratch_cap = scratch_need + scratch_need / 2 + 4096;
        CU(ctx, cudaMalloc((void **)&ctx->d_nested_scratch, sizeof(uint32_t) * ctx->nested_scratch_cap));
        topo_changed(ctx);
    }
    if (ctx->live_nested > ctx->nested_out_cap) {
        CU(ctx, sync_all(ctx));
        if (ctx->d_nested_out) CU(ctx, cudaFree(ctx->d_nested_out));
        ctx->d_nested_out = nullptr;
        ctx->nested_out_cap = std::max(64u, ctx->live_nested * 2);
        CU(ctx, cudaMalloc((void **)&ctx->d_nested_out, sizeof(NestedOut) * ctx->nested_out_cap));
        topo_changed(ctx);
    }
    {
        int rc = ensure_tiles(ctx);
        if (rc) return rc;
    }
    fs.spawn_per_slot.assign(n_slots, 0u);
    for (uint32_t s = 0; s < n_slots; s++)
        if (add[s] && ctx->slot_owner[s]) {
            ctx->slot_owner[s]->n_hi += add[s];
            fs.spawn_per_slot[s] = (uint32_t)std::min<uint64_t>(add[s], 0xFFFFFFFFull);
        }
    for (uint32_t s = 0; s < n_slots; s++)
        if (ctx->slot_owner[s]) ctx->slot_owner[s]->injected = false;

    // ---- parameter block of the frame: header | spawn_per_slot[phase][slot] | cmds | inputs |
    // nested cmds. Its layout depends only on the topology (stream slots, emitters, spawners,
    // phases), so that a frame is a fixed sequence of nodes that can be replayed as one CUDA graph.
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t cmds_cap = std::max<size_t>(1, ctx->live_emitters);
    const size_t inputs_cap = std::max<size_t>(1, ctx->spawners.size());
    const size_t nested_cap = std::max<size_t>(1, ctx->live_nested);
    const size_t off_spawn = align16(sizeof(FrameHeader));
    const size_t off_cmds = align16(off_spawn + sizeof(uint32_t) * n_phases * std::max(1u, n_slots));
    const size_t off_inputs = align16(off_cmds + sizeof(SpawnCmd) * cmds_cap);
    const size_t off_nested = align16(off_inputs + sizeof(SpawnerInput) * inputs_cap);
    const size_t off_prefix = align16(off_nested + sizeof(NestedCmd) * nested_cap);
    const size_t bytes = align16(off_prefix + sizeof(uint32_t) * kNumVariants * (n_slots + 1));
    {
        int rc = ensure_frame_slot(ctx, fs, bytes);
        if (rc) return rc;
    }
    FrameHeader *h = (FrameHeader *)fs.host;
    memset(h, 0, sizeof(*h));
    h->dt = dt;
    h->n_slots = n_slots;
    h->n_phases = n_phases;
    h->epoch = (uint32_t)((ctx->frame_no + 1) & 0x3FFFFFFFu);
    // fast path (no nested emitters): no plan kernel. The update tiles come from this host-built
    // table of upper bounds: ceil(n_hi / tile) per stream, n_hi >= the stream's real count.
    const bool derive = n_phases == 1;
    h->derive = derive ? 1u : 0u;
    // Concurrent spawn+step: when every stream is a FIFO ring (no compaction)
    // the spawn kernel also gives its new particles their first update, so it needs no ordering
    // against the update kernel (which then only covers older particles) and runs on a forked
    // branch. Timed (profiling) frames stay sequential so that kernels are timed in isolation.
    uint32_t n_fifo = 0, n_compacting = 0, n_fifo_collide = 0;
    for (uint32_t v = 0; v < kNumVariants; v++) {
        (variant_is_fifo(v) ? n_fifo : n_compacting) += ctx->variant_streams[v];
        if (variant_is_fifo(v) && variant_collides(v)) n_fifo_collide += ctx->variant_streams[v];
    }
    const bool step_in_spawn = derive && ctx->concurrent_spawn && !ctx->profiling && n_fifo > 0 && n_compacting == 0;
    const int spawn_collides = n_fifo_collide > 0 ? (ctx->colliders_revolved ? 2 : 1) : 0;
    h->step_in_spawn = step_in_spawn ? 1u : 0u;
    if (derive) {
        uint32_t *hp = (uint32_t *)(fs.host + off_prefix);
        uint32_t base = 0;
        for (uint32_t v = 0; v < kNumVariants; v++) {
            uint32_t run = 0;
            uint32_t *pv = hp + (size_t)v * (n_slots + 1);
            if (ctx->variant_streams[v]) {
                for (uint32_t s = 0; s < n_slots; s++) {
                    pv[s] = run;
                    const Stream *st = ctx->slot_owner[s];
                    if (st && st->variant == v) {
                        // upper bound of the particles the update kernel covers in this stream
                        const uint64_t covered = step_in_spawn ? st->n_hi - std::min<uint64_t>(st->n_hi, add[s]) : st->n_hi;
                        // (FIFO tiles are aligned to physical slots: up to 31 idle lanes in front, see update_kernel)
                        const uint64_t lead = 31;
                        if (covered) run += (uint32_t)((std::min<uint64_t>(covered, st->block.capacity) + lead + kTile - 1) / kTile);
                    }
                }
                pv[n_slots] = run;
            }
            h->host_n_tiles[v] = run;
            h->host_tile_base[v] = base;
            base += run;
        }
    }
    uint64_t total_spawn = 0;
    {
        SpawnCmd *hc = (SpawnCmd *)(fs.host + off_cmds);
        NestedCmd *hn = (NestedCmd *)(fs.host + off_nested);
        uint32_t nc = 0, nn = 0;
        for (uint32_t p = 0; p < n_phases; p++) {
            PhaseInfo &ph = h->phase[p];
            ph.cmd_begin = nc;
            for (const SpawnCmd &c : cmds[p]) hc[nc++] = c;
            ph.cmd_end = nc;
            ph.total_spawn = (uint32_t)phase_total[p];
            ph.nested_begin = nn;
            for (const NestedCmd &c : nested[p]) hn[nn++] = c;
            ph.nested_end = nn;
            total_spawn += phase_total[p];
        }
        h->n_cmds = nc;
        h->n_nested = nn;
    }
    if (n_slots) memcpy(fs.host + off_spawn, spawn_per_slot.data(), sizeof(uint32_t) * n_phases * n_slots);
    if (!sinputs.empty()) memcpy(fs.host + off_inputs, sinputs.data(), sizeof(SpawnerInput) * sinputs.size());

    DeviceTables t{};
    t.descs = ctx->d_descs;
    const int old_buf = cur_buf(ctx), new_buf = old_buf ^ 1; // frame f writes buffer f & 1
    t.states = states_of(ctx, new_buf);
    t.states_prev = states_of(ctx, old_buf);
    t.settings = ctx->d_settings;
    t.emitters = ctx->d_emitters;
    t.colliders = ctx->d_colliders;
    t.broadphase = ctx->d_broadphase;
    t.n_colliders = ctx->n_colliders;
    t.tile_prefix = ctx->d_tile_prefix;
    t.slots_cap = ctx->slots_cap;
    t.lookback_capacity = ctx->tiles_cap;
    t.plan = plan_of(ctx, new_buf);
    t.lookback = ctx->d_lookback;
    t.seed = ctx->seed;
    t.nested_scratch = ctx->d_nested_scratch;
    t.nested_serial = ctx->d_nested_serial;
    t.nested_out = ctx->d_nested_out;
    FrameDeviceInputs f{};
    f.header = (const FrameHeader *)fs.dev;
    f.spawn_per_slot = (const uint32_t *)(fs.dev + off_spawn);
    f.cmds = (const SpawnCmd *)(fs.dev + off_cmds);
    f.inputs = (const SpawnerInput *)(fs.dev + off_inputs);
    f.nested = (const NestedCmd *)(fs.dev + off_nested);
    f.host_tile_prefix = (const uint32_t *)(fs.dev + off_prefix);

    const bool prof = ctx->profiling;
    uint32_t variant_mask = 0;
    for (uint32_t v = 0; v < kNumVariants; v++)
        if (ctx->variant_streams[v]) variant_mask |= 1u << v;
    // nested emitters per phase is a property of the topology (enabled or not), so a replayed
    // graph launches the nested kernels of every phase that could have commands
    uint32_t nested_slots[kMaxPhases] = {0};
    for (auto &spp : ctx->spawners) {
        uint32_t p = 0;
        for (const Emitter &e : spp->emitters)
            if (e.emits_on_other_particles) nested_slots[p++]++;
    }
    uint32_t launches = 0;
    // the device work of one frame; `replay` = being captured into a graph, so nothing may
    // depend on this particular frame's counts
    auto enqueue = [&](bool replay) -> int {
        launches = 0;
        bool forked = false;
        CU(ctx, cudaMemsetAsync(ctx->d_statebuf[new_buf], 0, statebuf_bytes(n_slots), ctx->stream));
        if (prof) CU(ctx, cudaEventRecord(fs.ev[0], ctx->stream));
        const bool single = n_phases == 1;
        if (!derive) { // nested emitters: last frame's buffer -> this frame's buffer, phase 0 appended
            CU(ctx, launch_plan(t, f, variant_mask, kPlanDeaths | kPlanAppend | (single ? kPlanTiles : 0u), 0, ctx->stream));
            launches++;
        }
        if (prof) CU(ctx, cudaEventRecord(fs.ev[1], ctx->stream));
        for (uint32_t p = 0; p < n_phases; p++) {
            if (p > 0 && (replay || phase_total[p])) {
                CU(ctx, launch_plan(t, f, variant_mask, kPlanAppend, p, ctx->stream));
                launches++;
            }
            if (replay || phase_total[p]) {
                if (step_in_spawn) { // fork: the spawn+step kernel is independent of the update kernel
                    // (only the fork point here; the kernel itself is enqueued AFTER the update
                    // kernel, below: the persistent update grid fills every SM, so whichever of
                    // the two starts first runs alone, and the short spawn kernel belongs in the
                    // update kernel's tail, not in front of it -- C3 with graphs: 0.293 -> 0.27 ms)
                    CU(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
                    forked = true;
                } else {
                    CU(ctx, launch_spawn(t, f, p, replay ? 0xFFFFFFFFu : (uint32_t)phase_total[p], false, 0, ctx->stream));
                }
                launches++;
            }
            const uint32_t nn = replay ? nested_slots[p] : (uint32_t)nested[p].size();
            if (nn) {
                CU(ctx, launch_nested(t, f, p, nn, ctx->stream));
                launches += 3;
            }
        }
        if (!single) {
            CU(ctx, launch_plan(t, f, variant_mask, kPlanTiles, 0, ctx->stream));
            launches++;
        }
        if (prof) CU(ctx, cudaEventRecord(fs.ev[2], ctx->stream));
        for (uint32_t v : {(uint32_t)kCompact, (uint32_t)kCompact | (uint32_t)kVarRot}) {
            if (!ctx->variant_streams[v]) continue; // death counts + per-stream prefixes for the compacting update (timed with it)
            CU(ctx, launch_count_scan(t, f, v, n_slots, ctx->stream));
            launches += 2;
        }
        for (uint32_t v = 0; v < kNumVariants; v++) {
            if (!ctx->variant_streams[v]) continue;
            CU(ctx, launch_update(t, f, v, ctx->grids[v], ctx->team_size, ctx->colliders_revolved, ctx->stream));
            launches++;
        }
        if (forked) {
            CU(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
            CU(ctx, launch_spawn(t, f, 0, replay ? 0xFFFFFFFFu : (uint32_t)phase_total[0], true, spawn_collides, ctx->side_stream));
            CU(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
            CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0)); // join
        }
        if (prof) CU(ctx, cudaEventRecord(fs.ev[3], ctx->stream));
        return FW_OK;
    };
    // The parameter upload depends on nothing the GPU computes: it goes to the copy stream and
    // overlaps the previous frame's kernels. The kernels wait for it, and for the readback of
    // frame f-2, which read the state buffer this frame is about to zero and rewrite.
    CU(ctx, cudaMemcpyAsync(fs.dev, fs.host, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(ctx, cudaEventRecord(fs.ev_h2d, ctx->copy_stream));
    CU(ctx, cudaStreamWaitEvent(ctx->stream, fs.ev_h2d, 0));
    {
        const FrameSlot &two_back = ctx->ring[(ctx->frame_no + kRing - 2) % kRing];
        if (ctx->frame_no >= 2 && two_back.frame + 2 == ctx->frame_no + 1) CU(ctx, cudaStreamWaitEvent(ctx->stream, two_back.done, 0));
    }
    ctx->topo_stable_frames++;
    if (ctx->use_graphs && !prof && ctx->topo_stable_frames > kGraphWarmFrames) {
        if (!fs.graph_exec || fs.graph_version != ctx->topo_version) {
            if (fs.graph_exec) cudaGraphExecDestroy(fs.graph_exec);
            fs.graph_exec = nullptr;
            cudaGraph_t g = nullptr;
            int rc = FW_ERR_CUDA;
            cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                rc = enqueue(true);
                e = cudaStreamEndCapture(ctx->stream, &g);
            }
            cudaError_t ei = cudaErrorUnknown;
            if (rc == FW_OK && e == cudaSuccess) ei = cudaGraphInstantiate(&fs.graph_exec, g, 0);
            if (g) cudaGraphDestroy(g);
            if (ei != cudaSuccess) {
                // e.g. an external stream that cannot be captured: launch kernel by kernel from
                // now on instead of failing the frame
                (void)cudaGetLastError();
                fs.graph_exec = nullptr;
                ctx->use_graphs = false;
            } else {
                fs.graph_version = ctx->topo_version;
                fs.graph_launches = launches;
            }
        }
    }
    if (ctx->use_graphs && !prof && ctx->topo_stable_frames > kGraphWarmFrames && fs.graph_exec && fs.graph_version == ctx->topo_version) {
        launches = fs.graph_launches;
        CU(ctx, cudaGraphLaunch(fs.graph_exec, ctx->stream));
    } else {
        int rc = enqueue(false);
        if (rc) return rc;
    }
    // asynchronous readback of [PlanOut | stream states] (counts, AABBs, bounds for the next
    // frames) on its own stream (the next frame's kernels only read this buffer). Not the upload
    // stream: behind this copy, the next frame's parameter upload would wait for this frame's
    // kernels instead of overlapping them
    CU(ctx, cudaEventRecord(fs.ev_kernels, ctx->stream));
    CU(ctx, cudaStreamWaitEvent(ctx->rb_stream, fs.ev_kernels, 0));
    CU(ctx, cudaMemcpyAsync(fs.readback, ctx->d_statebuf[new_buf], statebuf_bytes(n_slots), cudaMemcpyDeviceToHost, ctx->rb_stream));
    CU(ctx, cudaEventRecord(fs.done, ctx->rb_stream));
    ctx->snapshot_valid = false;
    ctx->readback_is_current = true;
    ctx->frame_no++;
    ctx->prev_dt = dt;
    fs.in_flight = true;
    fs.frame = ctx->frame_no;
    fs.profiled = prof;
    fs.pending_account = true;
    fs.launches = launches;
    fs.particles_spawned = total_spawn;
    fs.h2d_bytes = bytes;
    fs.d2h_bytes = statebuf_bytes(n_slots);
    undo.committed = true;
    return FW_OK;
}

int fw_sync(fw_context *ctx) {
    ENTER(ctx);
    CU(ctx, sync_all(ctx));
    for (FrameSlot &fs : ctx->ring) {
        if (!fs.plan_host) continue;
        ctx->device_error_flags |= fs.plan_host->error_flags;
        fs.plan_host->error_flags = 0;
    }
    if (ctx->device_error_flags) {
        const uint32_t fl = ctx->device_error_flags;
        ctx->device_error_flags = 0;
        return fail(ctx, FW_ERR_INTERNAL, "device reported error flags 0x%x (1 = a ring overflowed and spawns were dropped, 2 = look-back table too small, 4 = a nested emitter exceeded its planned per-parent bound)", fl);
    }
    return FW_OK;
}

int fw_poll_device_errors(fw_context *ctx, uint32_t *flags) {
    ENTER(ctx);
    poll_readbacks(ctx); // (never waits)
    if (flags) *flags = ctx->device_error_flags;
    ctx->device_error_flags = 0;
    return FW_OK;
}

int fw_counts(fw_context *ctx, uint32_t key, uint32_t *out, uint32_t n_types) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_counts: unknown spawner %u", key);
    if (n_types < sp->streams.size() || !out) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_counts: need room for %zu counts", sp->streams.size());
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    for (size_t t = 0; t < sp->streams.size(); t++) out[t] = (uint32_t)sp->streams[t].n_hi;
    return FW_OK;
}

int fw_counts_all(fw_context *ctx, uint32_t *keys, uint32_t *types, uint32_t *counts, uint32_t cap, uint32_t *n_streams) {
    ENTER(ctx);
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    uint32_t n = 0;
    for (auto &sp : ctx->spawners) n += (uint32_t)sp->streams.size();
    if (n_streams) *n_streams = n;
    if (n > cap) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_counts_all: %u streams, room for %u", n, cap);
    uint32_t k = 0;
    for (auto &sp : ctx->spawners)
        for (Stream &st : sp->streams) {
            if (keys) keys[k] = sp->key;
            if (types) types[k] = st.type;
            if (counts) counts[k] = (uint32_t)st.n_hi;
            k++;
        }
    return FW_OK;
}

int fw_total_live(fw_context *ctx, uint64_t *out) {
    ENTER(ctx);
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    uint64_t n = 0;
    for (auto &sp : ctx->spawners)
        for (Stream &st : sp->streams) n += st.n_hi;
    if (out) *out = n;
    return FW_OK;
}

int fw_spawner_status_get(fw_context *ctx, uint32_t key, fw_spawner_status *out) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp || !out) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_spawner_status_get: unknown spawner %u", key);
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    uint64_t live = 0;
    for (Stream &st : sp->streams) live += st.n_hi;
    out->live_particles = live;
    out->all_empty = live == 0; // src/core.rs:679
    out->active = spawner_active(*sp, live != 0);
    out->finished = (out->all_empty && !out->active && sp->initialized && !sp->finished_notified) ? 1u : 0u; // :679-682
    out->finished_notified = sp->finished_notified;
    return FW_OK;
}

int fw_spawner_mark_finished_notified(fw_context *ctx, uint32_t key) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "unknown spawner %u", key);
    sp->finished_notified = true; // src/core.rs:685
    return FW_OK;
}

int fw_stream_layout_get(fw_context *ctx, uint32_t key, uint32_t type, fw_stream_layout *out) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp || type >= sp->streams.size() || !out) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_stream_layout_get: unknown spawner %u / type %u", key, type);
    const Stream &st = sp->streams[type];
    static_assert(FW_LAYOUT_COMPACTING == kVarCompact && FW_LAYOUT_COLLIDES == kVarCollide && FW_LAYOUT_ROTATES == kVarRot, "variant bits");
    static_assert(FW_STORE_BASE_COLOR == kStoreBase && FW_STORE_EMISSIVE_COLOR == kStoreEmi && FW_STORE_SCALE == kStoreScale &&
                      FW_STORE_LIFETIME == kStoreLife, "flag bits");
    const bool rot = variant_rotates(st.variant), compact = !variant_is_fifo(st.variant), life = (st.flags & kStoreLife) != 0;
    out->variant = st.variant;
    out->flags = st.flags;
    // position + age, velocity (+ angular_velocity.x | initial_scale): always; rotation 16, angular
    // velocity y,z 8, (lifetime, initial_scale) 8 for a rotating stream; (lifetime, age) 8 when a static
    // stream's lifetime varies
    out->bytes_read = 32u + (rot ? 32u : (life ? 8u : 0u));
    out->bytes_written = 32u + (rot ? 24u + (compact ? 8u : 0u) : (life ? 8u : 0u)) + ((st.flags & kStoreBase) ? 16u : 0u) +
                         ((st.flags & kStoreEmi) ? 16u : 0u) + ((st.flags & kStoreScale) ? 4u : 0u) + (compact ? 4u * st.n_lea : 0u);
    out->bytes_count_pass = (compact && !variant_collides(st.variant)) ? (rot ? 24u : (life ? 8u : 16u)) : 0u;
    out->capacity = st.block.capacity;
    return FW_OK;
}

int fw_read_particles(fw_context *ctx, uint32_t key, uint32_t type, fw_particle_data *out, uint64_t cap, uint64_t *n) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp || type >= sp->streams.size()) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_read_particles: unknown spawner %u / type %u", key, type);
    const Stream &st = sp->streams[type];
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    uint32_t first, live;
    live_range(ctx, st, first, live);
    if (n) *n = live;
    if (live > cap || (live && !out)) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_read_particles: %u particles, room for %llu", live, (unsigned long long)cap);
    if (!live) return FW_OK;
    if ((rc = ensure_stage(ctx, (size_t)live * sizeof(fw_particle_data)))) return rc;
    // pbr is a copy of the type's setting (src/core.rs:462)
    CU(ctx, launch_gather_particles(block_desc(st.block, st.variant, st.flags), ctx->d_settings + st.slot, first, live, st.ps.pbr,
                                    (fw_particle_data *)ctx->d_stage, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out, ctx->d_stage, (size_t)live * sizeof(fw_particle_data), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_all(ctx));
    return FW_OK;
}

int fw_write_particles(fw_context *ctx, uint32_t key, uint32_t type, const fw_particle_data *in, uint64_t n) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp || type >= sp->streams.size()) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_write_particles: unknown spawner %u / type %u", key, type);
    if (n && !in) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_write_particles: null");
    if (n > 0xFFFFFF00ull) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "too many particles");
    Stream &st = sp->streams[type];
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    // What the stream does not keep per particle was proved at reset from its settings (stream_proofs);
    // host-written rows that break a proof -- a lifetime other than the type's constant or ages that
    // increase along the Vec (FIFO: deaths are a prefix), a rotation / angular velocity / colour / scale
    // other than the constant -- turn that part of the state on for good.
    {
        uint32_t nv, nf;
        check_rows_against_proofs(st, in, n, nv, nf);
        if (nv != st.variant || nf != st.flags) {
            ctx->variant_streams[st.variant]--;
            st.variant = nv;
            st.flags = nf;
            ctx->variant_streams[st.variant]++;
            topo_changed(ctx);
            if ((rc = upload_desc(ctx, st))) return rc;
        }
    }
    if (n > usable_capacity(st.variant, st.block.capacity)) {
        ctx->snapshot[st.slot].count = 0;
        ctx->snapshot[st.slot].dead = 0;
        if ((rc = grow_stream(ctx, st, n))) return rc;
        if ((rc = ensure_tiles(ctx))) return rc;
    }
    if (n) {
        if ((rc = ensure_stage(ctx, (size_t)n * sizeof(fw_particle_data)))) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->d_stage, in, (size_t)n * sizeof(fw_particle_data), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, launch_scatter_particles(block_desc(st.block, st.variant, st.flags), (uint32_t)n, (const fw_particle_data *)ctx->d_stage, ctx->stream));
    }
    const StreamState ns = injected_state(st.variant, (uint32_t)n, st.block.capacity);
    CU(ctx, cudaMemcpyAsync(cur_states(ctx) + st.slot, &ns, sizeof(ns), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, sync_all(ctx));
    st.n_hi = n;
    st.born_frame = ctx->frame_no + 1;
    st.injected = true;
    ctx->snapshot_valid = false;
    ctx->readback_is_current = false;
    return FW_OK;
}

int fw_read_instances(fw_context *ctx, uint32_t key, uint32_t type, fw_particle_instance *out, uint64_t cap, uint64_t *n) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp || type >= sp->streams.size()) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_read_instances: unknown spawner %u / type %u", key, type);
    const Stream &st = sp->streams[type];
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    uint32_t first, live;
    live_range(ctx, st, first, live);
    if (n) *n = live;
    if (live > cap || (live && !out)) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_read_instances: %u rows, room for %llu", live, (unsigned long long)cap);
    if (!live) return FW_OK;
    // rows of this one stream: [n_rows, offset] words then the rows
    const size_t hdr = 64;
    if ((rc = ensure_stage(ctx, hdr + (size_t)live * 64))) return rc;
    DeviceTables t{};
    t.descs = ctx->d_descs;
    t.states = cur_states(ctx);
    t.settings = ctx->d_settings;
    CU(ctx, launch_pack_instances(t, st.slot, st.slot + 1, (float4 *)(ctx->d_stage + hdr), live, (unsigned long long *)ctx->d_stage, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out, ctx->d_stage + hdr, (size_t)live * 64, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_all(ctx));
    return FW_OK;
}

int fw_read_destroyed(fw_context *ctx, uint32_t key, uint32_t type, fw_particle_data *out, uint64_t cap, uint64_t *n) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp || type >= sp->streams.size()) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_read_destroyed: unknown spawner %u / type %u", key, type);
    const Stream &st = sp->streams[type];
    if (n) *n = 0;
    if (!st.destroyed.base) return FW_OK; // no handler registered: nothing is kept (src/core.rs:660-663)
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    const uint32_t dead = ctx->snapshot[st.slot].dead; // deaths of the last update, in Vec order
    if (n) *n = dead;
    if (dead > cap || (dead && !out)) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_read_destroyed: %u particles, room for %llu", dead, (unsigned long long)cap);
    if (!dead) return FW_OK;
    if ((rc = ensure_stage(ctx, (size_t)dead * sizeof(fw_particle_data)))) return rc;
    CU(ctx, launch_gather_particles(block_desc(st.destroyed, st.variant, st.flags), ctx->d_settings + st.slot, 0, dead, st.ps.pbr,
                                    (fw_particle_data *)ctx->d_stage, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out, ctx->d_stage, (size_t)dead * sizeof(fw_particle_data), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_all(ctx));
    return FW_OK;
}

int fw_read_aabb(fw_context *ctx, uint32_t key, float out_min[3], float out_max[3], uint32_t *empty) {
    ENTER(ctx);
    Spawner *sp = find(ctx, key);
    if (!sp) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_read_aabb: unknown spawner %u", key);
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    uint64_t live = 0;
    for (Stream &st : sp->streams) {
        const StreamState &s = ctx->snapshot[st.slot];
        if (s.count - s.dead == 0) continue;
        live += s.count - s.dead;
        for (int k = 0; k < 3; k++) {
            if (s.aabb_min_inv[k]) mn[k] = std::fmin(mn[k], dec_f32(~s.aabb_min_inv[k]));
            if (s.aabb_max[k]) mx[k] = std::fmax(mx[k], dec_f32(s.aabb_max[k]));
        }
    }
    if (out_min) memcpy(out_min, mn, sizeof(mn));
    if (out_max) memcpy(out_max, mx, sizeof(mx));
    if (empty) *empty = live == 0;
    return FW_OK;
}

int fw_pack_instances_device(fw_context *ctx, void *device_dst, uint64_t cap_rows, uint64_t *n_rows) {
    ENTER(ctx);
    if (!device_dst && cap_rows) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_pack_instances_device: null destination");
    if (ctx->pack_cap < ctx->n_slots + 2) {
        CU(ctx, sync_all(ctx));
        if (ctx->d_pack) CU(ctx, cudaFree(ctx->d_pack));
        ctx->d_pack = nullptr;
        ctx->pack_cap = std::max(1024u, (ctx->n_slots + 2) * 2);
        CU(ctx, cudaMalloc((void **)&ctx->d_pack, sizeof(unsigned long long) * ctx->pack_cap));
        if (!ctx->h_pack) CU(ctx, cudaMallocHost((void **)&ctx->h_pack, sizeof(unsigned long long)));
    }
    DeviceTables t{};
    t.descs = ctx->d_descs;
    t.states = cur_states(ctx);
    t.settings = ctx->d_settings;
    CU(ctx, launch_pack_instances(t, 0, ctx->n_slots, (float4 *)device_dst, cap_rows, ctx->d_pack, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->h_pack, ctx->d_pack, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_all(ctx));
    if (n_rows) *n_rows = *ctx->h_pack;
    if (*ctx->h_pack > cap_rows) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_pack_instances_device: %llu rows, room for %llu", *ctx->h_pack, (unsigned long long)cap_rows);
    return FW_OK;
}

int fw_extract_instances(fw_context *ctx, void *host_dst, uint64_t cap_rows, uint64_t *n_rows) {
    ENTER(ctx);
    if (!host_dst && cap_rows) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_extract_instances: null destination");
    uint64_t bound = 0; // host-side upper bound of the live rows, no sync needed
    for (auto &sp : ctx->spawners)
        for (Stream &st : sp->streams) bound += std::min<uint64_t>(st.n_hi, st.block.capacity);
    if (bound > ctx->extract_cap) {
        CU(ctx, sync_all(ctx));
        if (ctx->d_extract) CU(ctx, cudaFree(ctx->d_extract));
        ctx->d_extract = nullptr;
        ctx->extract_cap = bound + bound / 8 + 1024;
        CU(ctx, cudaMalloc((void **)&ctx->d_extract, ctx->extract_cap * 64));
    }
    uint64_t n = 0;
    int rc = fw_pack_instances_device(ctx, ctx->d_extract, ctx->extract_cap, &n);
    if (rc) return rc;
    if (n_rows) *n_rows = n;
    if (n > cap_rows) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_extract_instances: %llu rows, room for %llu", (unsigned long long)n, (unsigned long long)cap_rows);
    if (n) {
        CU(ctx, cudaMemcpyAsync(host_dst, ctx->d_extract, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, sync_all(ctx));
    }
    return FW_OK;
}

// ---- render hand-off that does not stall the simulation (reference consumer src/render.rs:439-461,
// upload :568-584): pack on the context's stream, copy on a stream of its own
int fw_extract_begin(fw_context *ctx, const uint32_t *spawner_keys, uint32_t n_keys, void *host_dst, uint64_t cap_rows) {
    ENTER(ctx);
    if ((!host_dst && cap_rows) || (n_keys && !spawner_keys)) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_extract_begin: null argument");
    if (ctx->xt_issued - ctx->xt_waited >= 2) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_extract_begin: two extracts are outstanding, call fw_extract_wait first");
    fw_context::ExtractSlot &x = ctx->xt[ctx->xt_issued & 1u];
    if (!ctx->xt_stream) CU(ctx, cudaStreamCreateWithFlags(&ctx->xt_stream, cudaStreamNonBlocking));
    if (!x.packed) {
        CU(ctx, cudaEventCreateWithFlags(&x.packed, cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&x.landed, cudaEventDisableTiming));
    }
    // which streams: every stream of the listed spawners (all spawners when no list), creation order
    x.slots.clear();
    uint64_t bound = 0; // host-side upper bound of the rows, no sync needed
    auto take = [&](Spawner &sp) {
        for (Stream &st : sp.streams) {
            x.slots.push_back(st.slot);
            bound += std::min<uint64_t>(st.n_hi, st.block.capacity);
        }
    };
    if (spawner_keys) {
        for (uint32_t k = 0; k < n_keys; k++) {
            Spawner *sp = find(ctx, spawner_keys[k]);
            if (!sp) return fail(ctx, FW_ERR_UNKNOWN_SPAWNER, "fw_extract_begin: unknown spawner %u", spawner_keys[k]);
            take(*sp);
        }
    } else {
        for (auto &sp : ctx->spawners) take(*sp);
    }
    if (bound > cap_rows) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_extract_begin: up to %llu rows, room for %llu", (unsigned long long)bound, (unsigned long long)cap_rows);
    if (bound > x.cap_rows) {
        CU(ctx, sync_all(ctx));
        CU(ctx, cudaStreamSynchronize(ctx->xt_stream));
        if (x.d_rows) CU(ctx, cudaFree(x.d_rows));
        x.d_rows = nullptr;
        x.cap_rows = bound + bound / 8 + 1024;
        CU(ctx, cudaMalloc((void **)&x.d_rows, x.cap_rows * 64));
    }
    if (x.slots.size() + 2 > x.offsets_cap) {
        CU(ctx, sync_all(ctx));
        CU(ctx, cudaStreamSynchronize(ctx->xt_stream));
        if (x.d_offsets) CU(ctx, cudaFree(x.d_offsets));
        if (x.h_offsets) CU(ctx, cudaFreeHost(x.h_offsets));
        x.d_offsets = nullptr;
        x.h_offsets = nullptr;
        if (x.d_slots) CU(ctx, cudaFree(x.d_slots));
        x.d_slots = nullptr;
        x.offsets_cap = (uint32_t)std::max<size_t>(1024, (x.slots.size() + 2) * 2);
        CU(ctx, cudaMalloc((void **)&x.d_offsets, sizeof(unsigned long long) * x.offsets_cap));
        CU(ctx, cudaMalloc((void **)&x.d_slots, sizeof(uint32_t) * x.offsets_cap));
        CU(ctx, cudaMallocHost((void **)&x.h_offsets, sizeof(unsigned long long) * x.offsets_cap));
    }
    if (!x.slots.empty()) CU(ctx, cudaMemcpyAsync(x.d_slots, x.slots.data(), sizeof(uint32_t) * x.slots.size(), cudaMemcpyHostToDevice, ctx->stream));
    {
        DeviceTables t{};
        t.descs = ctx->d_descs;
        t.states = cur_states(ctx);
        t.settings = ctx->d_settings;
        CU(ctx, launch_pack_listed(t, x.d_slots, (uint32_t)x.slots.size(), x.d_rows, x.cap_rows, x.d_offsets, ctx->stream));
    }
    CU(ctx, cudaEventRecord(x.packed, ctx->stream));
    // the copies: the row count is only known on the device, so the offsets come first and the rows
    // are copied up to the host-side bound (rows past the real count are never read by the caller)
    CU(ctx, cudaStreamWaitEvent(ctx->xt_stream, x.packed, 0));
    CU(ctx, cudaMemcpyAsync(x.h_offsets, x.d_offsets, sizeof(unsigned long long) * (x.slots.size() + 1), cudaMemcpyDeviceToHost, ctx->xt_stream));
    if (bound) CU(ctx, cudaMemcpyAsync(host_dst, x.d_rows, bound * 64, cudaMemcpyDeviceToHost, ctx->xt_stream));
    CU(ctx, cudaEventRecord(x.landed, ctx->xt_stream));
    // (this slot is packed again two extracts later; by then fw_extract_wait has waited for `landed`)
    x.host_dst = host_dst;
    x.host_cap_rows = cap_rows;
    x.pending = true;
    ctx->xt_issued++;
    return FW_OK;
}

int fw_extract_wait(fw_context *ctx, uint64_t *n_rows, uint64_t *stream_first_rows, uint32_t cap_streams, uint32_t *n_streams) {
    ENTER(ctx);
    if (ctx->xt_issued == ctx->xt_waited) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_extract_wait: no extract is outstanding");
    fw_context::ExtractSlot &x = ctx->xt[ctx->xt_waited & 1u];
    CU(ctx, cudaEventSynchronize(x.landed));
    x.pending = false;
    ctx->xt_waited++;
    if (n_rows) *n_rows = x.h_offsets[0];
    if (n_streams) *n_streams = (uint32_t)x.slots.size();
    if (stream_first_rows)
        for (uint32_t k = 0; k < cap_streams && k < x.slots.size(); k++) stream_first_rows[k] = x.h_offsets[1 + k];
    if (x.h_offsets[0] > x.host_cap_rows) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_extract_wait: %llu rows, room for %llu", x.h_offsets[0], (unsigned long long)x.host_cap_rows);
    return FW_OK;
}

// ---- zero-copy hand-off: the packed rows in a VMM allocation exported as a POSIX file descriptor.
// The driver entry points are fetched at run time (cudaGetDriverEntryPoint): the library does not link
// libcuda, so it still loads on a machine without a driver (the `-m "not gpu"` tests).
namespace {
struct VmmApi {
    // (signatures of cuda.h, spelled with plain types)
    int (*cuMemGetAllocationGranularity)(size_t *, const void *prop, int option) = nullptr;
    int (*cuMemCreate)(unsigned long long *handle, size_t size, const void *prop, unsigned long long flags) = nullptr;
    int (*cuMemAddressReserve)(unsigned long long *ptr, size_t size, size_t alignment, unsigned long long addr, unsigned long long flags) = nullptr;
    int (*cuMemMap)(unsigned long long ptr, size_t size, size_t offset, unsigned long long handle, unsigned long long flags) = nullptr;
    int (*cuMemSetAccess)(unsigned long long ptr, size_t size, const void *desc, size_t count) = nullptr;
    int (*cuMemExportToShareableHandle)(void *shareable, unsigned long long handle, int type, unsigned long long flags) = nullptr;
    int (*cuMemImportFromShareableHandle)(unsigned long long *handle, void *os_handle, int type) = nullptr;
    int (*cuMemUnmap)(unsigned long long ptr, size_t size) = nullptr;
    int (*cuMemRelease)(unsigned long long handle) = nullptr;
    int (*cuMemAddressFree)(unsigned long long ptr, size_t size) = nullptr;
    bool ok = false;
};
// CUmemAllocationProp / CUmemAccessDesc of cuda.h (CUDA 12): restated so that cuda.h is not needed
struct VmmLocation { int type; int id; };                       // CU_MEM_LOCATION_TYPE_DEVICE = 1
struct VmmAllocFlags { unsigned char compressionType, gpuDirectRDMACapable; unsigned short usage; unsigned char reserved[4]; };
struct VmmProp { int type; int requestedHandleTypes; VmmLocation location; void *win32HandleMetaData; VmmAllocFlags allocFlags; }; // type: PINNED = 1; handle type POSIX_FD = 1
struct VmmAccessDesc { VmmLocation location; int flags; };      // CU_MEM_ACCESS_FLAGS_PROT_READWRITE = 3
const VmmApi &vmm_api() {
    static VmmApi api = [] {
        VmmApi a;
        auto get = [](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        a.ok = get("cuMemGetAllocationGranularity", (void **)&a.cuMemGetAllocationGranularity) && get("cuMemCreate", (void **)&a.cuMemCreate) &&
               get("cuMemAddressReserve", (void **)&a.cuMemAddressReserve) && get("cuMemMap", (void **)&a.cuMemMap) &&
               get("cuMemSetAccess", (void **)&a.cuMemSetAccess) && get("cuMemExportToShareableHandle", (void **)&a.cuMemExportToShareableHandle) &&
               get("cuMemImportFromShareableHandle", (void **)&a.cuMemImportFromShareableHandle) && get("cuMemUnmap", (void **)&a.cuMemUnmap) &&
               get("cuMemRelease", (void **)&a.cuMemRelease) && get("cuMemAddressFree", (void **)&a.cuMemAddressFree);
        (void)cudaGetLastError();
        return a;
    }();
    return api;
}
VmmProp vmm_prop(int device) {
    VmmProp p;
    memset(&p, 0, sizeof(p));
    p.type = 1;                 // CU_MEM_ALLOCATION_TYPE_PINNED
    p.requestedHandleTypes = 1; // CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    p.location.type = 1;        // CU_MEM_LOCATION_TYPE_DEVICE
    p.location.id = device;
    return p;
}
void vmm_release(fw_context *ctx) {
    const VmmApi &a = vmm_api();
    if (ctx->vmm_ptr && a.ok) {
        a.cuMemUnmap(ctx->vmm_ptr, ctx->vmm_bytes);
        a.cuMemAddressFree(ctx->vmm_ptr, ctx->vmm_bytes);
        a.cuMemRelease(ctx->vmm_handle);
    }
    ctx->vmm_ptr = 0;
    ctx->vmm_handle = 0;
    ctx->vmm_bytes = 0;
}
} // namespace

int fw_export_instances_fd(fw_context *ctx, int32_t *fd, uint64_t *bytes, uint64_t *n_rows) {
    ENTER(ctx);
    if (!fd) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_export_instances_fd: null");
    const VmmApi &a = vmm_api();
    if (!a.ok) return fail(ctx, FW_ERR_UNSUPPORTED, "fw_export_instances_fd: the driver does not expose the virtual-memory API");
    int rc = refresh_exact(ctx);
    if (rc) return rc;
    uint64_t rows = 0;
    for (auto &sp : ctx->spawners)
        for (Stream &st : sp->streams) rows += st.n_hi;
    const VmmProp prop = vmm_prop(ctx->device);
    size_t gran = 0;
    if (a.cuMemGetAllocationGranularity(&gran, &prop, 0 /* MINIMUM */) != 0 || gran == 0) return fail(ctx, FW_ERR_CUDA, "cuMemGetAllocationGranularity failed");
    const size_t need = ((std::max<uint64_t>(rows, 1) * 64 + gran - 1) / gran) * gran;
    CU(ctx, sync_all(ctx));
    vmm_release(ctx);
    if (a.cuMemCreate(&ctx->vmm_handle, need, &prop, 0) != 0) return fail(ctx, FW_ERR_OUT_OF_MEMORY, "cuMemCreate(%zu bytes, POSIX fd handle) failed", need);
    if (a.cuMemAddressReserve(&ctx->vmm_ptr, need, 0, 0, 0) != 0) {
        a.cuMemRelease(ctx->vmm_handle);
        ctx->vmm_handle = 0;
        ctx->vmm_ptr = 0;
        return fail(ctx, FW_ERR_CUDA, "cuMemAddressReserve failed");
    }
    ctx->vmm_bytes = need;
    VmmAccessDesc acc;
    acc.location = prop.location;
    acc.flags = 3;
    if (a.cuMemMap(ctx->vmm_ptr, need, 0, ctx->vmm_handle, 0) != 0 || a.cuMemSetAccess(ctx->vmm_ptr, need, &acc, 1) != 0) {
        vmm_release(ctx);
        return fail(ctx, FW_ERR_CUDA, "cuMemMap / cuMemSetAccess failed");
    }
    uint64_t n = 0;
    rc = fw_pack_instances_device(ctx, (void *)(uintptr_t)ctx->vmm_ptr, need / 64, &n);
    if (rc) return rc;
    int os_fd = -1;
    if (a.cuMemExportToShareableHandle(&os_fd, ctx->vmm_handle, 1, 0) != 0 || os_fd < 0) return fail(ctx, FW_ERR_CUDA, "cuMemExportToShareableHandle failed");
    *fd = os_fd;
    if (bytes) *bytes = need;
    if (n_rows) *n_rows = n;
    return FW_OK;
}

int fw_import_instances_fd(int32_t device, int32_t fd, uint64_t bytes, uint64_t n_rows, void *host_dst) {
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, FW_ERR_NO_DEVICE, "fw_import_instances_fd: cudaSetDevice(%d) failed", device);
    cudaFree(nullptr); // make sure the primary context exists
    const VmmApi &a = vmm_api();
    if (!a.ok) return fail(nullptr, FW_ERR_UNSUPPORTED, "fw_import_instances_fd: the driver does not expose the virtual-memory API");
    if (n_rows * 64 > bytes || (n_rows && !host_dst)) return fail(nullptr, FW_ERR_INVALID_ARGUMENT, "fw_import_instances_fd: %llu rows do not fit %llu bytes", (unsigned long long)n_rows, (unsigned long long)bytes);
    unsigned long long handle = 0, ptr = 0;
    if (a.cuMemImportFromShareableHandle(&handle, (void *)(uintptr_t)fd, 1) != 0) return fail(nullptr, FW_ERR_CUDA, "cuMemImportFromShareableHandle failed");
    int rc = FW_OK;
    if (a.cuMemAddressReserve(&ptr, bytes, 0, 0, 0) != 0) {
        a.cuMemRelease(handle);
        return fail(nullptr, FW_ERR_CUDA, "cuMemAddressReserve failed");
    }
    VmmAccessDesc acc;
    acc.location.type = 1;
    acc.location.id = device;
    acc.flags = 3;
    if (a.cuMemMap(ptr, bytes, 0, handle, 0) != 0 || a.cuMemSetAccess(ptr, bytes, &acc, 1) != 0) {
        rc = fail(nullptr, FW_ERR_CUDA, "cuMemMap / cuMemSetAccess of the imported allocation failed");
    } else if (n_rows && cudaMemcpy(host_dst, (void *)(uintptr_t)ptr, n_rows * 64, cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = fail(nullptr, FW_ERR_CUDA, "copy from the imported allocation failed");
    }
    a.cuMemUnmap(ptr, bytes);
    a.cuMemAddressFree(ptr, bytes);
    a.cuMemRelease(handle);
    return rc;
}

// ---- multi-GPU render extract over NVLink peer memory (reference consumer
// src/render.rs:439-461 wants every instance row on the GPU that draws)
static void gather_release(fw_context *ctx) {
    for (uint32_t r = 0; r < kMaxGatherRanks; r++) {
        if (ctx->gather_ipc[r] && ctx->gather.base[r]) cudaIpcCloseMemHandle(ctx->gather.base[r]);
        ctx->gather_ipc[r] = false;
        ctx->gather.base[r] = nullptr;
    }
    if (ctx->d_gather) cudaFree(ctx->d_gather);
    ctx->d_gather = nullptr;
    if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
    ctx->h_gather = nullptr;
    ctx->gather_connected = false;
    ctx->gather.n_ranks = 0;
    (void)cudaGetLastError();
}

int fw_gather_create(fw_context *ctx, uint32_t n_ranks, uint32_t my_rank, uint64_t cap_rows_per_rank, fw_gather_handle *out) {
    ENTER(ctx);
    if (!out || n_ranks == 0 || n_ranks > kMaxGatherRanks || my_rank >= n_ranks || cap_rows_per_rank == 0)
        return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_gather_create: n_ranks %u (max %u), my_rank %u, cap %llu", n_ranks,
                    kMaxGatherRanks, my_rank, (unsigned long long)cap_rows_per_rank);
    CU(ctx, sync_all(ctx));
    gather_release(ctx);
    const size_t bytes = kGatherHeaderBytes + (size_t)n_ranks * cap_rows_per_rank * 64;
    CU(ctx, cudaMalloc((void **)&ctx->d_gather, bytes));
    CU(ctx, cudaMemset(ctx->d_gather, 0, kGatherHeaderBytes));
    CU(ctx, cudaMallocHost((void **)&ctx->h_gather, sizeof(GatherHeader)));
    ctx->gather.n_ranks = n_ranks;
    ctx->gather.my_rank = my_rank;
    ctx->gather.cap_rows_per_rank = cap_rows_per_rank;
    ctx->gather.base[my_rank] = ctx->d_gather;
    ctx->gather_epoch = 0;
    memset(out, 0, sizeof(*out));
    cudaIpcMemHandle_t h;
    CU(ctx, cudaIpcGetMemHandle(&h, ctx->d_gather));
    static_assert(sizeof(h) <= sizeof(out->ipc), "fw_gather_handle.ipc too small");
    memcpy(out->ipc, &h, sizeof(h));
    out->address = (uint64_t)(uintptr_t)ctx->d_gather;
    out->bytes = bytes;
    out->device = ctx->device;
    out->pid = (int32_t)getpid();
    return FW_OK;
}

int fw_gather_connect(fw_context *ctx, const fw_gather_handle *handles, uint32_t n_handles) {
    ENTER(ctx);
    if (!ctx->d_gather) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_gather_connect: call fw_gather_create first");
    if (!handles || n_handles != ctx->gather.n_ranks)
        return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_gather_connect: %u handles for %u ranks", n_handles, ctx->gather.n_ranks);
    const size_t bytes = kGatherHeaderBytes + (size_t)ctx->gather.n_ranks * ctx->gather.cap_rows_per_rank * 64;
    for (uint32_t r = 0; r < n_handles; r++) {
        if (r == ctx->gather.my_rank) continue;
        if (handles[r].bytes != bytes)
            return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_gather_connect: rank %u created its buffer with another geometry", r);
        if (handles[r].pid == (int32_t)getpid()) { // another context of this process: plain peer access
            if (handles[r].device != ctx->device) {
                int can = 0;
                CU(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, handles[r].device));
                if (!can) return fail(ctx, FW_ERR_UNSUPPORTED, "fw_gather_connect: device %d cannot access device %d", ctx->device, handles[r].device);
                cudaError_t e = cudaDeviceEnablePeerAccess(handles[r].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(ctx, e);
                (void)cudaGetLastError();
            }
            ctx->gather.base[r] = (uint8_t *)(uintptr_t)handles[r].address;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles[r].ipc, sizeof(h));
            void *p = nullptr;
            CU(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ctx->gather.base[r] = (uint8_t *)p;
            ctx->gather_ipc[r] = true;
        }
    }
    ctx->gather_connected = true;
    return FW_OK;
}

int fw_gather_instances(fw_context *ctx) {
    ENTER(ctx);
    if (!ctx->gather_connected) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_gather_instances: not connected (fw_gather_create / fw_gather_connect)");
    if (ctx->pack_cap < ctx->n_slots + 2) {
        CU(ctx, sync_all(ctx));
        if (ctx->d_pack) CU(ctx, cudaFree(ctx->d_pack));
        ctx->d_pack = nullptr;
        ctx->pack_cap = std::max(1024u, (ctx->n_slots + 2) * 2);
        CU(ctx, cudaMalloc((void **)&ctx->d_pack, sizeof(unsigned long long) * ctx->pack_cap));
    }
    if (!ctx->h_pack) CU(ctx, cudaMallocHost((void **)&ctx->h_pack, sizeof(unsigned long long)));
    const GatherPeers &g = ctx->gather;
    const unsigned long long epoch = ++ctx->gather_epoch;
    const unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
    DeviceTables t{};
    t.descs = ctx->d_descs;
    t.states = cur_states(ctx);
    t.settings = ctx->d_settings;
    PackDst dst{};
    dst.n = g.n_ranks;
    for (uint32_t r = 0; r < g.n_ranks; r++)
        dst.rows[r] = (float4 *)(g.base[r] + kGatherHeaderBytes + (size_t)g.my_rank * g.cap_rows_per_rank * 64);
    // every rank is done reading the previous epoch (stream order on each rank) before anyone
    // overwrites it; then the rows; then "landed" flags + counts, and wait for everybody's
    CU(ctx, launch_gather_signal(g, kGatherReady, epoch, nullptr, timeout_ns, ctx->stream));
    CU(ctx, launch_pack_instances(t, 0, ctx->n_slots, dst, g.cap_rows_per_rank, ctx->d_pack, ctx->stream));
    CU(ctx, launch_gather_signal(g, kGatherDone, epoch, ctx->d_pack, timeout_ns, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->h_pack, ctx->d_pack, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    return FW_OK;
}

int fw_gather_result(fw_context *ctx, void **device_rows, uint64_t *rows_per_rank, uint32_t n_ranks, uint64_t *region_stride_rows) {
    ENTER(ctx);
    if (!ctx->gather_connected || ctx->gather_epoch == 0) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_gather_result: no gather was issued");
    if (n_ranks < ctx->gather.n_ranks && rows_per_rank) return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_gather_result: room for %u ranks, %u needed", n_ranks, ctx->gather.n_ranks);
    CU(ctx, cudaMemcpyAsync(ctx->h_gather, ctx->d_gather, sizeof(GatherHeader), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_all(ctx));
    if (ctx->h_gather->error) {
        const unsigned long long who = ctx->h_gather->error - 1;
        CU(ctx, cudaMemset(&((GatherHeader *)ctx->d_gather)->error, 0, sizeof(unsigned long long)));
        return fail(ctx, FW_ERR_INTERNAL, "fw_gather: timed out waiting for rank %llu", who);
    }
    if (*ctx->h_pack > ctx->gather.cap_rows_per_rank)
        return fail(ctx, FW_ERR_BUFFER_TOO_SMALL, "fw_gather: this rank has %llu rows, its region holds %llu", *ctx->h_pack,
                    (unsigned long long)ctx->gather.cap_rows_per_rank);
    if (device_rows) *device_rows = ctx->d_gather + kGatherHeaderBytes;
    if (region_stride_rows) *region_stride_rows = ctx->gather.cap_rows_per_rank;
    if (rows_per_rank)
        for (uint32_t r = 0; r < ctx->gather.n_ranks; r++) rows_per_rank[r] = ctx->h_gather->rows[r];
    return FW_OK;
}

int fw_gather_destroy(fw_context *ctx) {
    ENTER(ctx);
    CU(ctx, sync_all(ctx));
    gather_release(ctx);
    return FW_OK;
}

int fw_device_sincos(fw_context *ctx, const float *x, uint64_t n, float *sin_out, float *cos_out) {
    ENTER(ctx);
    if (n && (!x || !sin_out || !cos_out)) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_device_sincos: null");
    if (!n) return FW_OK;
    int rc = ensure_stage(ctx, (size_t)n * 12);
    if (rc) return rc;
    float *dx = (float *)ctx->d_stage, *ds = dx + n, *dc = ds + n;
    CU(ctx, cudaMemcpyAsync(dx, x, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, launch_sincos(dx, n, ds, dc, ctx->stream));
    CU(ctx, cudaMemcpyAsync(sin_out, ds, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(cos_out, dc, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, sync_all(ctx));
    return FW_OK;
}

int fw_event_record(fw_context *ctx, uint32_t slot) {
    ENTER(ctx);
    if (slot >= 16) return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_event_record: slot %u out of range", slot);
    if (!ctx->user_events[slot]) CU(ctx, cudaEventCreate(&ctx->user_events[slot]));
    CU(ctx, cudaEventRecord(ctx->user_events[slot], ctx->stream));
    return FW_OK;
}
int fw_event_elapsed_ms(fw_context *ctx, uint32_t a, uint32_t b, float *out_ms) {
    ENTER(ctx);
    if (a >= 16 || b >= 16 || !ctx->user_events[a] || !ctx->user_events[b] || !out_ms)
        return fail(ctx, FW_ERR_INVALID_ARGUMENT, "fw_event_elapsed_ms: markers not recorded");
    CU(ctx, cudaEventSynchronize(ctx->user_events[b]));
    CU(ctx, cudaEventElapsedTime(out_ms, ctx->user_events[a], ctx->user_events[b]));
    return FW_OK;
}

int fw_set_profiling(fw_context *ctx, uint32_t on) {
    ENTER(ctx);
    ctx->profiling = on != 0;
    return FW_OK;
}

int fw_profile_last(fw_context *ctx, fw_frame_profile *out) {
    ENTER(ctx);
    CU(ctx, sync_all(ctx));
    // absorb in submission order
    for (uint32_t k = 0; k < kRing; k++) absorb_profile(ctx, ctx->ring[(ctx->frame_no + k) % kRing]);
    if (out) *out = ctx->prof_last;
    return FW_OK;
}
int fw_profile_sum(fw_context *ctx, fw_frame_profile *out, uint32_t *n_frames) {
    int rc = fw_profile_last(ctx, nullptr);
    if (rc) return rc;
    if (out) *out = ctx->prof_sum;
    if (n_frames) *n_frames = ctx->prof_frames;
    return FW_OK;
}
int fw_profile_reset(fw_context *ctx) {
    ENTER(ctx);
    int rc = fw_profile_last(ctx, nullptr);
    if (rc) return rc;
    memset(&ctx->prof_sum, 0, sizeof(ctx->prof_sum));
    memset(&ctx->prof_last, 0, sizeof(ctx->prof_last));
    ctx->prof_frames = 0;
    ctx->prof_timed_frames = 0;
    return FW_OK;
}

void *fw_stream_handle(fw_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

} // extern "C"
