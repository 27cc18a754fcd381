// fw_internal.h -- structures shared by the host side (fw_api.cu) and the kernels
// (fw_kernels.cu) of libfirework_b200.so. Not part of the public ABI.
//
// Device data layout. One *stream* = one (spawner, particle type) vector
// `data.particles[i]` (reference src/core.rs:274). A stream owns one device block of
// `capacity` particle slots holding eight structure-of-arrays packs. Which packs a frame touches
// depends on what fw_spawner_reset could PROVE about the stream (StreamDesc::flags, the kVarRot bit
// of its variant); a field that is provably the same for every particle of the stream, for ever, is
// a per-stream constant in DevParticleSettings and is neither loaded nor stored:
//
//   pack  offset   type    contents                                       update reads  writes
//   m0    0        float4  position.xyz, age                                   16         16
//   m2    32*cap   float4  velocity.xyz, w                                     16         16
//                          w = angular_velocity.x (rotating stream) | initial_scale (static one)
//   m1    16*cap   float4  rotation (x,y,z,w)             rotating streams     16         16
//   m3    48*cap   float2  angular_velocity.y, .z         rotating streams      8          8
//   k     56*cap   float2  rotating: lifetime, initial_scale                    8          - (moved by compaction)
//                          static + kStoreLife: lifetime, copy of age          (8)        (8)
//   o0    64*cap   float4  base_color                     kStoreBase            -         16
//   o1    80*cap   float4  emissive_color                 kStoreEmi             -         16
//   o2    96*cap   float   scale                          kStoreScale           -          4
//
//   generic stream (everything varies)          64 B read + 92 B written = 156 B  
//   examples/stress_test.rs (C2, C3; no angular motion, constant emissive / scale curve / lifetime):
//                                               32 B read + 48 B written =  80 B
//
// "Static" = no angular motion can ever occur (every emitter of the type draws a zero angular
// velocity, angular_acceleration is +0): rotation stays at the fixed point of
// `from_scaled_axis(0) * initial_rotation`, angular velocity at +0 (src/core.rs:645-650 evaluated on
// those inputs). Constant gradients / scale curve: src/core.rs:602-605,652-655 return the same value
// every frame. Host-written state (fw_write_particles) that breaks a proof turns the flag on.
//
// (An earlier layout kept the 64-byte ParticleInstance row of reference src/render.rs:95-103
// resident as AoS; ncu showed DRAM fetching the whole 64-byte row to read its 32-byte state
// half: +0.32 GB per 10 M particles, profiles/r1_a_*. Rows are now assembled by the
// pack/extract kernel, which every render hand-off needs anyway.)
//
// Each stream is a ring: logical particle i (the reference's Vec index) lives in slot
// (first + i) mod capacity, first = live_first(): head + dead for a FIFO ring, head + count for a
// compacting ring (which writes its survivors behind the particles a frame reads). Order inside the
// ring == the reference's Vec order (survivors keep their order, spawns are appended), so no
// per-particle serial is stored.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/firework_b200.h"

namespace fw {

#ifndef FW_TILE
#define FW_TILE 256
#endif
constexpr int kTile = FW_TILE; // particles per update tile == threads per CTA
constexpr int kUpdateThreads = FW_TILE;
constexpr uint32_t kBytesPerSlot = 100;

// variants of the update kernel = template instantiations, one launch and one tile table each.
// bit 0: compacting (else FIFO: constant lifetime, no destroy-on-collision -- deaths are a prefix of
//        the Vec; compacting = stable compaction out of place inside the ring, deaths precounted or,
//        with collisions, found by decoupled look-back)
// bit 1: sweeps its particles against the colliders
// bit 2: rotating (rotation / angular velocity are per-particle state; else per-stream constants)
enum Variant : uint32_t {
    kFifo = 0,
    kCompact = 1,
    kFifoCollide = 2,
    kCompactCollide = 3,
    kVarCompact = 1,
    kVarCollide = 2,
    kVarRot = 4,
    kNumVariants = 8
};
__host__ __device__ inline bool variant_is_fifo(uint32_t v) { return (v & kVarCompact) == 0u; }
__host__ __device__ inline bool variant_collides(uint32_t v) { return (v & kVarCollide) != 0u; }
__host__ __device__ inline bool variant_rotates(uint32_t v) { return (v & kVarRot) != 0u; }
// StreamDesc::flags: which derived / constant fields are per-particle state of this stream
constexpr uint32_t kStoreBase = 1u;  // base_color varies (gradient not constant, or host-written)
constexpr uint32_t kStoreEmi = 2u;   // emissive_color varies
constexpr uint32_t kStoreScale = 4u; // scale != initial_scale * constant
constexpr uint32_t kStoreLife = 8u;  // lifetime varies (static streams only; rotating ones always keep it in k)
constexpr uint32_t kStoreAll = 15u;

// device forms of fw_curve_f32 / fw_gradient with 16-byte aligned tables
struct alignas(16) DevCurve {
    float times[FW_MAX_KNOTS];
    float values[FW_MAX_KNOTS];
    uint32_t kind, n, pad[2];
};
struct alignas(16) DevGradient {
    float4 colors[FW_MAX_KNOTS];
    float times[FW_MAX_KNOTS];
    uint32_t kind, n, pad[2];
};

// per-stream settings; staged into shared memory once per tile with a bulk async copy (TMA
// unit), so sizeof must be a multiple of 16.
struct alignas(16) DevParticleSettings {
    DevCurve scale_curve;       // 144 B
    DevGradient base_color;     // 336 B
    DevGradient emissive_color; // 336 B
    float acceleration[3];
    float linear_drag;
    float angular_acceleration[3];
    float angular_drag;
    fw_collision_settings collision; // 56 B
    // spawn-only
    fw_rand_f32 lifetime;
    fw_rand_f32 initial_scale;
    // per-stream constants standing in for packs the stream does not keep (see the layout above)
    float const_lifetime;    // lifetime.generate() when min == max
    uint32_t pad[1];
    float const_rotation[4]; // fixed point of from_scaled_axis(0) * initial_rotation
};
static_assert(sizeof(DevParticleSettings) % 16 == 0, "bulk copy needs a 16-byte multiple");

constexpr uint32_t kMaxLea = 4;    // nested emitters that may target one particle type
constexpr uint32_t kMaxPhases = 8; // 1 + nested emitters per spawner

struct StreamDesc { // written by the host when a stream is created / grown / removed
    uint8_t *base;           // device block of capacity * (100 + 4*n_lea) bytes, 256-byte aligned
    uint8_t *destroyed_base; // same layout, holds the particles destroyed by the last update
                             // (only for types with a particles_destroyed handler), else null
    uint32_t capacity;       // multiple of 256; 0 = slot unused
    uint32_t variant;
    uint32_t n_lea;          // last_emitted_age arrays (one per nested emitter targeting this type)
    uint32_t flags;          // kStore*
};
// ParticleData.last_emitted_age[i] of reference src/core.rs:320, kept only for the emitters that
// read it (nested emitters whose target is this particle type): float[capacity] each, behind
// the eight packs
__host__ __device__ inline float *lea_array(uint8_t *base, uint32_t cap, uint32_t j) {
    return (float *)(base + (size_t)cap * (kBytesPerSlot + 4u * j));
}

// typed views of a stream block
struct StreamArrays {
    float4 *m0, *m1, *m2;
    float2 *m3, *k;
    float4 *o0, *o1;
    float *o2;
};
__host__ __device__ inline StreamArrays stream_arrays(uint8_t *base, uint32_t cap) {
    StreamArrays a;
    const size_t c = cap;
    a.m0 = (float4 *)(base);
    a.m1 = (float4 *)(base + 16 * c);
    a.m2 = (float4 *)(base + 32 * c);
    a.m3 = (float2 *)(base + 48 * c);
    a.k = (float2 *)(base + 56 * c);
    a.o0 = (float4 *)(base + 64 * c);
    a.o1 = (float4 *)(base + 80 * c);
    a.o2 = (float *)(base + 96 * c);
    return a;
}

// Broad phase of the collision sweep, one device blob rebuilt by fw_set_colliders:
//   leaf[2i], leaf[2i+1]  inflated world AABB of collider i: (min.xyz | layers bits), (max.xyz | -)
//   nodes                 BVH over those boxes in depth-first order, 2 float4 per node:
//                         (min.xyz | skip link, leaf: layers bits), (max.xyz | collider index or ~0)
//   big                   colliders too large for the grid: tested for every ray
//   cell_start, items     uniform grid over the other colliders (CSR): a ray segment no longer than a
//                         cell per axis looks at <= 8 cells instead of walking the BVH
struct BroadPhaseHeader {
    uint32_t n_nodes, nodes_off, leaf_off; // offsets in bytes from the blob's start
    uint32_t n_big, big_off;
    uint32_t use_grid, cell_off, items_off;
    uint32_t dim[3];
    float lo[3], inv_cell[3];
    uint32_t pad[3];
};
static_assert(sizeof(BroadPhaseHeader) == 80, "16-byte multiple: the arrays behind it hold float4");
constexpr uint32_t kGridMaxDim = 64;
// deterministic capacity of the blob for n colliders (same n => same buffer => same kernel arguments)
inline size_t broadphase_cells_cap(uint32_t n) { return (size_t)(8u * n < 64u ? 64u : (8u * n > kGridMaxDim * kGridMaxDim * kGridMaxDim ? kGridMaxDim * kGridMaxDim * kGridMaxDim : 8u * n)); }
inline size_t broadphase_items_cap(uint32_t n) { return 32u * (size_t)n; }
inline size_t broadphase_bytes(uint32_t n) {
    return sizeof(BroadPhaseHeader) + 32u * (size_t)n /*leaf*/ + 32u * (2u * (size_t)n) /*nodes*/ + 4u * (size_t)n /*big*/ +
           4u * (broadphase_cells_cap(n) + 1u) + 4u * broadphase_items_cap(n) + 64u;
}

// Stream states are double-buffered: frame f reads the buffer frame f-1 wrote and writes the
// other one (zeroed by a memset node first), so every kernel can DERIVE the state it needs --
// head after last frame's deaths, count after this frame's spawns -- functionally from the old
// buffer instead of waiting for a serial "plan" kernel.
struct StreamState { // mutated by kernels
    uint32_t head;
    uint32_t count;      // particles that entered this frame's update (live + spawned)
    uint32_t dead;       // deaths of this frame's update (atomic), applied by the next frame
    uint32_t spawn_base; // logical index of this frame's first spawned particle
    // per-stream AABB of position -/+ scale as order-preserving uint encodings, zero = empty:
    // aabb_min_inv = ~enc(min) (so that atomicMax keeps the minimum), aabb_max = enc(max)
    uint32_t aabb_min_inv[3];
    uint32_t aabb_max[3];
    uint32_t overflow; // spawns dropped because the ring was full (the host grows before that)
    uint32_t pad;
};

struct SpawnCmd { // one per (emitter, frame) with count > 0
    uint32_t stream;
    uint32_t emitter_idx;   // index into the device fw_emission_settings array
    uint32_t input_idx;     // index into the per-frame SpawnerInput array
    uint32_t count;
    uint32_t first;         // exclusive prefix of count over the commands of the frame
    uint32_t dst_off;       // offset inside the block appended to the stream this frame
    uint32_t spawner_key;   // RNG protocol
    uint32_t emitter_local; // emitter index inside its spawner (RNG protocol)
    uint64_t serial_base;   // first particle serial of this command
};

struct SpawnerInput {
    float translation[3];
    float rotation[4];
    float parent_velocity[3];
    float modifier_scale;
    float modifier_speed;
};

// One nested emitter instance of a frame (reference src/core.rs:471-546)
struct NestedCmd {
    uint32_t parent_stream;  // particles[target_particle_type]
    uint32_t child_stream;   // particles[particle_index]
    uint32_t emitter_idx;    // device emission settings index; also indexes nested_serial
    uint32_t emitter_local;  // RNG protocol
    uint32_t spawner_key;    // RNG protocol
    uint32_t lea_index;      // which last_emitted_age array of the parent stream
    uint32_t input_idx;      // SpawnerInput of the spawner (EffectModifier)
    uint32_t scratch_off;    // this command's per-parent counts inside nested_scratch
    uint32_t per_parent_cap; // the host's capacity planning assumed at most this many per parent
    uint32_t pad[3];
};
struct NestedOut { // device, per nested command of the frame
    uint32_t total;      // children emitted
    uint32_t spawn_base; // logical index of the first child in the child stream
    uint64_t serial_base;
};

// The reference walks a spawner's emitters in order, and Global and Nested emitters may feed
// the same particle type, so a frame is split into phases: phase p = the Global emitters that
// sit between the (p-1)-th and the p-th Nested emitter of their spawner, followed by the p-th
// Nested emitter of every spawner. Frames without nested emitters have exactly one phase.
struct PhaseInfo {
    uint32_t cmd_begin, cmd_end; // SpawnCmd range
    uint32_t total_spawn;        // particles of the Global commands of the phase
    uint32_t nested_begin, nested_end;
    uint32_t pad[3];
};
struct FrameHeader {
    float dt;
    uint32_t n_slots; // stream slots to scan
    uint32_t n_cmds;
    uint32_t n_phases;
    uint32_t epoch;
    uint32_t n_nested;
    // derive = 1: the fast path (no nested emitters). There is no plan kernel: spawn and update
    // derive head / counts from the previous state buffer and the update tiles come from a
    // host-built table of UPPER BOUNDS (tiles past a stream's real count exit at once).
    uint32_t derive;
    uint32_t step_in_spawn; // derive path only: spawn_kernel<STEP> runs concurrently with update_kernel
    uint32_t host_n_tiles[kNumVariants];
    uint32_t host_tile_base[kNumVariants];
    PhaseInfo phase[kMaxPhases];
};

struct PlanOut { // device, per frame; lives in front of the stream states of the same buffer
    uint32_t n_tiles[kNumVariants];
    uint32_t tile_base[kNumVariants]; // start of each variant inside the look-back array
    uint32_t error_flags;
    uint32_t total_update; // particles entering the update this frame
    uint32_t pad[14];
};
static_assert(sizeof(PlanOut) == 128, "state buffer layout: [PlanOut (128 B)][StreamState x slots]");

struct DeviceTables {
    const StreamDesc *descs;
    StreamState *states;            // this frame's buffer (written)
    const StreamState *states_prev; // last frame's buffer (read)
    const DevParticleSettings *settings; // indexed by stream slot
    const fw_emission_settings *emitters;
    const fw_collider *colliders;
    const uint8_t *broadphase; // BroadPhaseHeader + its arrays (see cast_ray in fw_math.cuh)
    uint32_t n_colliders;
    // tile_prefix[v * slots_cap + s] = number of update tiles of variant v in slots < s
    uint32_t *tile_prefix;
    uint32_t slots_cap;
    uint32_t lookback_capacity;
    PlanOut *plan;
    unsigned long long *lookback; // compact variants, two words per tile: [tile] look-back status or precounted
                                  // exclusive prefix, [lookback_capacity + tile] precounted per-warp counts
    uint64_t seed;
    // nested emission
    uint32_t *nested_scratch;          // per-parent emission counts -> exclusive offsets
    unsigned long long *nested_serial; // per device emitter: particles spawned so far
    NestedOut *nested_out;             // per nested command of the frame
};

struct FrameDeviceInputs {
    const FrameHeader *header;
    const uint32_t *host_tile_prefix; // derive path: [kNumVariants][n_slots + 1] upper-bound tiles
    const uint32_t *spawn_per_slot; // [n_phases][n_slots]
    const SpawnCmd *cmds;
    const SpawnerInput *inputs;
    const NestedCmd *nested;
};

// ---- multi-GPU render extract over peer memory: every rank owns a gather
// buffer [GatherHeader (4096 B) | n_ranks regions of cap_rows_per_rank 64-byte rows] that all
// ranks map (CUDA IPC or same-process peer access); rank r's pack kernel stores its rows into
// region r of EVERY buffer.
constexpr uint32_t kMaxGatherRanks = 16;
constexpr size_t kGatherHeaderBytes = 4096;
constexpr uint32_t kGatherReady = 0, kGatherDone = 1;
struct GatherHeader {                           // written by the peers, system-scope release/acquire
    unsigned long long rows[kMaxGatherRanks];   // rows[r]: rank r's row count of the current epoch
    unsigned long long ready[kMaxGatherRanks];  // ready[r] = e: rank r no longer reads epoch e-1
    unsigned long long done[kMaxGatherRanks];   // done[r] = e: rank r's rows of epoch e have landed
    unsigned long long error;                   // 1 + rank that timed out, 0 = none
};
static_assert(sizeof(GatherHeader) <= kGatherHeaderBytes, "gather header");
struct GatherPeers {
    uint8_t *base[kMaxGatherRanks]; // every rank's gather buffer as mapped in this process
    uint32_t n_ranks, my_rank;
    uint64_t cap_rows_per_rank;
};
struct PackDst {
    float4 *rows[kMaxGatherRanks];
    uint32_t n;
};

// launchers (fw_kernels.cu)
// plan: what = bit 0 apply last frame's deaths, bit 1 append the Global spawns of `phase`,
// bit 2 build the tile prefix tables
constexpr uint32_t kPlanDeaths = 1u, kPlanAppend = 2u, kPlanTiles = 4u;
cudaError_t launch_plan(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant_mask, uint32_t what,
                        uint32_t phase, cudaStream_t s);
// total_spawn: particles of the phase, or 0xFFFFFFFF = unknown at launch time (graph replay)
// step: the kernel also applies this frame's update to the particles it creates (see spawn_kernel);
// collide: ... including the collision sweep of the streams that have one
cudaError_t launch_spawn(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t phase, uint32_t total_spawn, bool step,
                         int collide /* 0 none, 1 cuboids / spheres, 2 + cylinders / cones */, cudaStream_t s);
// the nested emitters of a phase: count per parent, scan + append, spawn the children
cudaError_t launch_nested(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t phase, uint32_t n_cmds, cudaStream_t s);
// revolved: the collider set contains cylinders / cones (selects the kernel build that can test them)
cudaError_t launch_update(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant, int grid, int team_size, bool revolved,
                          cudaStream_t s);
// compaction without collisions (variants kCompact, kCompact | kVarRot): per-tile death counts and their per-stream exclusive prefixes (before launch_update)
cudaError_t launch_count_scan(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant, uint32_t n_slots, cudaStream_t s);
cudaError_t update_grid_size(int device, int *grids /*[kNumVariants]*/, int *team_size);
// live ParticleInstance rows of the streams [slot_begin, slot_end) -> contiguous 64-byte rows
cudaError_t launch_pack_instances(const DeviceTables &t, uint32_t slot_begin, uint32_t slot_end, float4 *dst,
                                  uint64_t cap_rows, unsigned long long *n_rows_and_offsets, cudaStream_t s);
cudaError_t launch_pack_instances(const DeviceTables &t, uint32_t slot_begin, uint32_t slot_end, const PackDst &dst,
                                  uint64_t cap_rows, unsigned long long *n_rows_and_offsets, cudaStream_t s);
// the same for the streams d_slot_list[0 .. n) (device array): render extract of a subset of the spawners
cudaError_t launch_pack_listed(const DeviceTables &t, const uint32_t *d_slot_list, uint32_t n, float4 *dst, uint64_t cap_rows,
                               unsigned long long *n_rows_and_offsets, cudaStream_t s);
cudaError_t launch_gather_signal(const GatherPeers &p, uint32_t which, unsigned long long epoch, const unsigned long long *rows_src,
                                 unsigned long long timeout_ns, cudaStream_t s);
// one stream <-> fw_particle_data rows (host mirror / fw_write_particles); d.base may be the stream's
// destroyed block; ps = the stream's device settings (constants of the packs it does not keep)
cudaError_t launch_gather_particles(const StreamDesc &d, const DevParticleSettings *ps, uint32_t first, uint32_t n, uint32_t pbr,
                                    fw_particle_data *dst, cudaStream_t s);
cudaError_t launch_scatter_particles(const StreamDesc &d, uint32_t n, const fw_particle_data *src, cudaStream_t s);
// include/fw_sincos.h evaluated on the device (parity hook)
cudaError_t launch_sincos(const float *x, uint64_t n, float *s, float *c, cudaStream_t st);
// ring -> linear copy into a bigger block (growth)
cudaError_t launch_ring_copy(const StreamDesc &src, uint32_t first, uint32_t n, const StreamDesc &dst, cudaStream_t s);

} // namespace fw
