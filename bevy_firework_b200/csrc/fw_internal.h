// fw_internal.h -- structures shared by the host side (fw_api.cu) and the kernels
// (fw_kernels.cu) of libfirework_b200.so. Not part of the public ABI.
//
// Device data layout (one *stream* = one (spawner, particle type) vector
// `data.particles[i]`, reference src/core.rs:274):
//
//   rows : float4[4*capacity]  the 64-byte ParticleInstance row of reference
//                              src/render.rs:95-103, AoS: [pos.xyz,scale][rot][base][emissive].
//                              It IS the live state for position/scale/rotation and at the same
//                              time the vertex-instance buffer a renderer consumes.
//   s0   : float4[capacity]    velocity.xyz, age
//   s1   : float4[capacity]    angular_velocity.xyz, lifetime
//   s2   : float [capacity]    initial_scale
//
// Each stream is a ring: logical particle i (the reference's Vec index) lives in slot
// (head + i) mod capacity. Order inside the ring == the reference's Vec order (survivors keep
// their order, spawns are appended), so no per-particle serial is stored.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/firework_b200.h"

namespace fw {

constexpr int kTile = 256;          // particles per update tile == threads per CTA
constexpr int kUpdateThreads = 256;

// variants of the update kernel (one tile table each)
enum Variant : uint32_t {
    kFifo = 0,          // constant lifetime, no destroy-on-collision: deaths are a prefix
    kCompact = 1,       // anything else: in-place stable compaction (decoupled look-back)
    kFifoCollide = 2,
    kCompactCollide = 3,
    kNumVariants = 4
};

// update-relevant part of fw_particle_settings; staged into shared memory once per tile with a
// bulk async copy, so sizeof must be a multiple of 16.
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

struct alignas(16) DevParticleSettings {
    DevCurve scale_curve;       // 144 B
    DevGradient base_color;     // 336 B
    DevGradient emissive_color; // 336 B
    float acceleration[3];
    float linear_drag;
    float angular_acceleration[3];
    float angular_drag;
    fw_collision_settings collision; // 20 B
    // spawn-only
    fw_rand_f32 lifetime;
    fw_rand_f32 initial_scale;
    uint32_t pad[3];
};
static_assert(sizeof(DevParticleSettings) % 16 == 0, "bulk copy needs a 16-byte multiple");

struct StreamDesc { // written by the host when a stream is created / grown / removed
    float4 *rows;
    float4 *s0;
    float4 *s1;
    float *s2;
    uint32_t capacity; // 0 = slot unused
    uint32_t settings_idx;
    uint32_t variant;
    uint32_t pad;
};

struct StreamState { // mutated by kernels
    uint32_t head;
    uint32_t count;      // live particles (after the plan kernel: including this frame's spawns)
    uint32_t dead;       // deaths of the last update, applied by the next plan kernel
    uint32_t spawn_base; // logical index of this frame's first spawned particle
    uint32_t aabb_min[3]; // order-preserving uint encoding of float
    uint32_t aabb_max[3];
    uint32_t overflow;   // spawns dropped because the ring was full (host grows before that)
    uint32_t pad;
};

struct TileEntry {
    uint32_t stream;
    uint32_t tile; // tile index inside the stream
};

struct SpawnCmd { // one per (emitter, frame) with count > 0
    uint32_t stream;
    uint32_t emitter_idx;  // index into the device fw_emission_settings array
    uint32_t input_idx;    // index into the per-frame SpawnerInput array
    uint32_t count;
    uint32_t first;        // exclusive prefix of count over the commands of the frame
    uint32_t dst_off;      // offset inside the block appended to the stream this frame
    uint32_t spawner_key;  // RNG protocol
    uint32_t emitter_local; // emitter index inside its spawner (RNG protocol)
    uint64_t serial_base;  // first particle serial of this command
};

struct SpawnerInput {
    float translation[3];
    float rotation[4];
    float parent_velocity[3];
    float modifier_scale;
    float modifier_speed;
};

struct FrameHeader {
    float dt;
    uint32_t n_slots;     // stream slots to scan
    uint32_t n_cmds;
    uint32_t total_spawn;
    uint32_t epoch;
    uint32_t pad[3];
};

struct PlanOut { // device, written by the plan kernel
    uint32_t n_tiles[kNumVariants];
    uint32_t tile_base[kNumVariants]; // start of each variant inside the tile table
    uint32_t error_flags;
    uint32_t total_update; // particles entering the update this frame
    uint32_t pad[2];
};

struct DeviceTables {
    const StreamDesc *descs;
    StreamState *states;
    const DevParticleSettings *settings;
    const fw_emission_settings *emitters;
    const fw_collider *colliders;
    uint32_t n_colliders;
    TileEntry *tiles;
    uint32_t tiles_capacity;
    PlanOut *plan;
    unsigned long long *lookback; // one status word per tile (compact variants)
    uint64_t seed;
};

struct FrameDeviceInputs {
    const FrameHeader *header;
    const uint32_t *spawn_per_slot;
    const SpawnCmd *cmds;
    const SpawnerInput *inputs;
};

// launchers (fw_kernels.cu)
cudaError_t launch_plan(const DeviceTables &t, const FrameDeviceInputs &f, cudaStream_t s);
cudaError_t launch_spawn(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t total_spawn,
                         cudaStream_t s);
cudaError_t launch_update(const DeviceTables &t, const FrameDeviceInputs &f, uint32_t variant,
                          int grid, cudaStream_t s);
cudaError_t update_grid_size(int device, int *grids /*[kNumVariants]*/);
cudaError_t launch_pack_instances(const DeviceTables &t, uint32_t n_slots, float4 *dst,
                                  uint64_t cap_rows, unsigned long long *n_rows, cudaStream_t s);

} // namespace fw
