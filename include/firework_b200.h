/*
 * firework_b200.h -- C ABI of the B200-native particle update path.
 *
 * This is the drop-in boundary for bevy_firework's per-frame simulation systems.
 * A Rust shim keeps the reference's public types (ParticleSystemPlugin, ParticleSpawner,
 * ParticleSettings, EmissionSettings, ParticleSpawnerData, ParticleData) and replaces the
 * bodies of three Bevy systems with calls into this library:
 *
 *   sync_spawner_data   (reference src/core.rs:343-365)  -> fw_spawner_reset
 *   spawn_particles     (reference src/core.rs:367-551)  \
 *   update_particles    (reference src/core.rs:577-670)  /-> fw_frame
 *   notify_finished_... (reference src/core.rs:674-688)  -> fw_spawner_status
 *   update_aabbs        (reference src/render.rs:677-703) -> fw_read_aabb
 *   extract -> ParticleInstance rows (reference src/render.rs:95-115, 368-461)
 *                                                         -> fw_read_instances / fw_pack_instances
 *
 * Rules of the ABI
 *   - plain C, plain-old-data only; no torch / C++ types cross it;
 *   - every call returns an int status (FW_OK == 0); the message of the last failure is
 *     fw_last_error(ctx) (or fw_last_global_error() when no context exists yet);
 *   - nothing unwinds across the boundary;
 *   - all device memory is owned by the opaque fw_context; callers pass and receive only
 *     caller-owned host buffers (the two *_device_* calls say so in their names);
 *   - any export may be called from any OS thread (the device is re-selected on entry),
 *     but a context is not re-entrant: one call at a time per context;
 *   - there is NO CPU fallback: without a usable CUDA device fw_create fails.
 *
 * All arithmetic is fp32; counts and indices are integers.
 */
#ifndef FIREWORK_B200_H
#define FIREWORK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FW_ABI_VERSION 2u

/* Maximum number of knots in one curve / gradient (the reference accepts any number of
 * samples, src/curve.rs:40-75; every shipped example uses <= 5). */
#define FW_MAX_KNOTS 32u

enum fw_status {
    FW_OK = 0,
    FW_ERR_INVALID_ARGUMENT = 1,
    FW_ERR_NO_DEVICE = 2,
    FW_ERR_CUDA = 3,
    FW_ERR_OUT_OF_MEMORY = 4,
    FW_ERR_UNKNOWN_SPAWNER = 5,
    FW_ERR_BUFFER_TOO_SMALL = 6,
    FW_ERR_UNSUPPORTED = 7,
    FW_ERR_INTERNAL = 8
};

/* ---- bevy_utilitarian 0.10.0 value generators (used at src/core.rs:102,107,155,157,161) */
typedef struct fw_rand_f32 {
    float min;
    float max;
} fw_rand_f32;

typedef struct fw_rand_vec3 {
    fw_rand_f32 magnitude;
    float direction[3];
    float spread; /* radians */
} fw_rand_vec3;

/* ---- FireworkCurve<f32> (src/curve.rs:8-75) and FireworkGradient<LinearRgba> (:171-239) */
enum fw_curve_kind {
    FW_CURVE_CONSTANT = 0, /* ConstantCurve: values[0]                                  */
    FW_CURVE_EVEN = 1,     /* SampleAuto / ColorSampleAuto: n >= 2 even samples on [0,1] */
    FW_CURVE_UNEVEN = 2    /* UnevenSampleAuto / ColorSampleUnevenAuto: n >= 2 (t,v)     */
};

typedef struct fw_curve_f32 {
    uint32_t kind;
    uint32_t n;
    float times[FW_MAX_KNOTS]; /* FW_CURVE_UNEVEN only; strictly increasing, finite */
    float values[FW_MAX_KNOTS];
} fw_curve_f32;

typedef struct fw_gradient {
    uint32_t kind;
    uint32_t n;
    float times[FW_MAX_KNOTS];
    float colors[FW_MAX_KNOTS][4]; /* LinearRgba r,g,b,a */
} fw_gradient;

/* ---- ParticleCollisionSettings (src/core.rs:240-248). `filter` is avian's SpatialQueryFilter
 * (:247, passed to cast_ray at :764): a layer mask and a set of excluded entities. */
#define FW_MAX_EXCLUDED 8u
#define FW_NO_KEY 0xFFFFFFFFu
typedef struct fw_collision_settings {
    uint32_t enabled; /* Option::is_some() */
    float restitution;
    float friction;
    uint32_t destroy_on_collision;
    uint32_t filter_mask; /* SpatialQueryFilter::mask: a collider is seen when (collider.layers & filter_mask) != 0 */
    uint32_t n_excluded;  /* SpatialQueryFilter::excluded_entities: colliders whose key is listed are skipped */
    uint32_t excluded_keys[FW_MAX_EXCLUDED];
} fw_collision_settings;

/* ---- ParticleSettings (src/core.rs:99-142), simulation-relevant fields only.
 * Texture handles, fade_edge, fade_scene and blend_mode are render-only and stay in Rust. */
typedef struct fw_particle_settings {
    fw_rand_f32 lifetime;
    fw_curve_f32 scale_curve;
    fw_rand_f32 initial_scale;
    float acceleration[3];
    float angular_acceleration[3];
    float linear_drag;
    float angular_drag;
    fw_gradient base_color;
    fw_gradient emissive_color;
    uint32_t pbr;
    fw_collision_settings collision;
    /* 1 when event_handlers.particles_destroyed is Some (src/core.rs:164-167, 660-667):
     * destroyed particles of the last frame are then kept for fw_read_destroyed. */
    uint32_t capture_destroyed;
    /* 0 = let the library size the stream; otherwise initial capacity in particles. */
    uint32_t capacity_hint;
} fw_particle_settings;

/* ---- EmissionPacing (src/core.rs:11-44), EmissionMode (:46-54), EmissionShape
 * (src/emission_shape.rs:6-16), EmissionSettings (src/core.rs:144-162) */
enum fw_pacing_kind {
    FW_PACING_ONE_SHOT = 0,
    FW_PACING_ON_DEMAND = 1,
    FW_PACING_COUNT_OVER_DURATION = 2
};
enum fw_emission_mode { FW_MODE_GLOBAL = 0, FW_MODE_NESTED = 1 };
enum fw_shape_kind { FW_SHAPE_POINT = 0, FW_SHAPE_SPHERE = 1, FW_SHAPE_CIRCLE = 2 };
enum fw_spawn_transform_mode { FW_TRANSFORM_GLOBAL = 0, FW_TRANSFORM_LOCAL = 1 };

typedef struct fw_emission_settings {
    uint32_t particle_index;
    uint32_t pacing_kind;
    uint64_t one_shot_count; /* EmissionPacing::OneShot(n) */
    float count;             /* CountOverDuration */
    float duration;
    float offset_start;
    float offset_end;
    uint32_t mode;
    uint32_t target_particle_type; /* EmissionMode::Nested */
    uint32_t shape_kind;
    float shape_radius;    /* Sphere(r) / Circle.radius */
    float shape_normal[3]; /* Circle.normal */
    fw_rand_vec3 initial_velocity;
    fw_rand_f32 initial_velocity_radial;
    uint32_t inherit_parent_velocity;
    float initial_rotation[4]; /* Quat x,y,z,w */
    fw_rand_vec3 initial_angular_velocity;
} fw_emission_settings;

/* ---- per-frame, per-spawner host inputs of spawn_particles (src/core.rs:367-376):
 * the spawn origin (GlobalTransform::compute_transform() or Transform according to
 * SpawnTransformMode -- the shim picks, :432-435), ParticleSpawnerData.parent_velocity
 * (:276, written by sync_parent_velocity :705-742), the EffectModifier (:323-336, default
 * 1/1) and the ParticleSpawnerData.manual_queued_count increment (:284-286). A spawner with
 * no entry in a frame keeps its previous origin/parent_velocity/modifier. */
typedef struct fw_spawner_frame_input {
    uint32_t spawner_key;
    float origin_translation[3];
    float origin_rotation[4]; /* Quat x,y,z,w */
    float parent_velocity[3];
    float modifier_scale;
    float modifier_speed;
    uint32_t queue_particles; /* added to manual_queued_count before spawning */
} fw_spawner_frame_input;

/* ---- ParticleData (src/core.rs:305-321) as a POD row; 104 bytes.
 * last_emitted_age (only read by Nested emission, :493,500) is kept device-side. */
typedef struct fw_particle_data {
    float position[3];
    float velocity[3];
    float rotation[4]; /* x,y,z,w */
    float angular_velocity[3];
    float initial_scale;
    float scale;
    float age;
    float lifetime;
    float base_color[4];
    float emissive_color[4];
    uint32_t pbr;
} fw_particle_data;

/* ---- ParticleInstance (src/render.rs:95-103): #[repr(C)] 64-byte vertex-instance row,
 * attributes 4 x Float32x4 at offsets 0/16/32/48 (src/render.rs:737-766). */
typedef struct fw_particle_instance {
    float position[3];
    float scale;
    float rotation[4];
    float base_color[4];
    float emissive_color[4];
} fw_particle_instance;

/* ---- static colliders seen by SpatialQuery::cast_ray (src/core.rs:756-765) */
enum fw_collider_kind {
    FW_COLLIDER_CUBOID = 0,   /* Collider::cuboid(x, y, z): half_extents = (x/2, y/2, z/2) */
    FW_COLLIDER_SPHERE = 1,   /* Collider::sphere(r): half_extents[0] = r */
    FW_COLLIDER_CYLINDER = 2, /* Collider::cylinder(r, height), axis +Y: half_extents = (r, height/2, -) */
    FW_COLLIDER_CONE = 3,     /* Collider::cone(r, height), apex at +Y: half_extents = (r, height/2, -)
                                 (examples/textures.rs:195,211 use both) */
    FW_COLLIDER_CAPSULE = 4   /* Collider::capsule(r, length), axis +Y: half_extents = (r, length/2, -); the segment
                                 from (0, -length/2, 0) to (0, length/2, 0) swept by a ball of radius r */
};

typedef struct fw_collider {
    uint32_t kind;
    uint32_t layers;       /* membership bits tested against fw_collision_settings.filter_mask */
    uint32_t key;          /* the collider's entity (any caller-chosen id, FW_NO_KEY = none): matched against
                              fw_collision_settings.excluded_keys */
    float half_extents[3]; /* see fw_collider_kind */
    float translation[3];
    float rotation[4]; /* Quat x,y,z,w */
} fw_collider;

/* ---- context */
typedef struct fw_config {
    uint32_t abi_version;  /* FW_ABI_VERSION */
    int32_t device;        /* CUDA ordinal */
    uint64_t seed;         /* Philox4x32-10 key of the spawn protocol */
    void *external_stream; /* cudaStream_t to launch on, or NULL: library-owned stream */
    uint32_t flags;        /* FW_FLAG_* */
    uint32_t reserved;
} fw_config;

#define FW_FLAG_PROFILE 1u   /* record CUDA events around every kernel of a frame */
#define FW_FLAG_NO_GRAPHS 2u /* always launch kernel by kernel (never replay frames as CUDA graphs) */
#define FW_FLAG_NO_CONCURRENT_SPAWN 4u /* always run the spawn kernel before the update kernel */

typedef struct fw_context fw_context;

typedef struct fw_spawner_status {
    uint32_t active;             /* ParticleSpawnerData::active(), src/core.rs:288-302 */
    uint32_t all_empty;          /* every particle vector empty, src/core.rs:679 */
    uint32_t finished;           /* condition of notify_finished_particle_spawners, :679-682 */
    uint32_t finished_notified;  /* latched by fw_spawner_mark_finished_notified (:685) */
    uint64_t live_particles;
} fw_spawner_status;

typedef struct fw_frame_profile {
    float plan_ms;
    float spawn_ms;
    float update_ms;
    float total_ms;
    uint32_t kernel_launches;
    uint32_t timed_frames; /* frames whose kernel times are in *_ms (profiling was on) */
    uint64_t particles_updated; /* live particles entering the update of that frame */
    uint64_t particles_spawned;
    uint64_t h2d_bytes; /* per-frame parameter block copied host -> device */
    uint64_t d2h_bytes; /* per-frame state readback copied device -> host */
} fw_frame_profile;

/* What the library keeps per particle for one stream (spawner, particle type). Fields that
 * fw_spawner_reset can prove constant over the stream's whole life from its settings -- no angular
 * motion (rotation, angular_velocity), a constant colour gradient or scale curve, one lifetime -- are
 * per-stream constants and cost no memory traffic; fw_read_* return them like any other field, and
 * fw_write_particles rows that contradict a proof make the field per-particle state again.
 * bytes_* = the algorithmic bytes one update moves per particle of this stream (SURVEY section 8d
 * restated per stream: 64 + 92 for a stream where nothing is provable). */
#define FW_LAYOUT_COMPACTING 1u /* variant bits */
#define FW_LAYOUT_COLLIDES 2u
#define FW_LAYOUT_ROTATES 4u
#define FW_STORE_BASE_COLOR 1u /* flags bits */
#define FW_STORE_EMISSIVE_COLOR 2u
#define FW_STORE_SCALE 4u
#define FW_STORE_LIFETIME 8u
typedef struct fw_stream_layout {
    uint32_t variant;
    uint32_t flags;
    uint32_t bytes_read;
    uint32_t bytes_written;
    uint32_t bytes_count_pass; /* compacting streams without collisions: read by the death-counting pass */
    uint32_t capacity;         /* slots of the stream's ring */
} fw_stream_layout;

const char *fw_last_global_error(void);
const char *fw_last_error(const fw_context *ctx);
uint32_t fw_abi_version(void);
/* sizeof() of a POD struct of this header by name ("fw_particle_settings", ...); 0 if unknown.
 * Lets a binding verify its layout against the compiled library. */
uint32_t fw_abi_sizeof(const char *struct_name);
/* offsetof(struct, field) by name, 0xFFFFFFFF if unknown: a binding in another language checks its
 * field order against the compiled library (rust/firework_b200_sys.rs, tests/test_abi.py) */
uint32_t fw_abi_offsetof(const char *struct_name, const char *field_name);

/* ---- host-side logic of the path, callable without a device (pinned by the `-m "not gpu"` tests) */
/* compute_emission_count (src/core.rs:553-575) exactly as fw_frame evaluates it on the host:
 * *times = particles to emit (`as usize`: NaN / negative -> 0), *next_last_emission = the emitter's
 * (or, for a nested emitter, the parent particle's) new last_emission */
int fw_host_emission_count(float time_passed_in_cycle, float last_emission, float cycle_duration,
                           float offset_start, float offset_end, float particles_per_cycle,
                           uint64_t *times, float *next_last_emission);
/* the broad phase fw_set_colliders builds for a collider set (what SpatialQuery::cast_ray's BVH
 * is to the reference, src/core.rs:756-765): inflated world AABBs, stackless BVH, uniform grid;
 * layout in DESIGN.md section 3 / BroadPhaseHeader. *n_bytes = size of the blob; it is written to
 * `out` if cap_bytes suffices (FW_ERR_BUFFER_TOO_SMALL otherwise; out may be NULL to ask). */
int fw_host_build_broadphase(const fw_collider *colliders, uint32_t n, void *out, uint64_t cap_bytes,
                             uint64_t *n_bytes);

/* include/fw_sincos.h as compiled into the kernels (-fmad=false): sine and cosine of n host floats,
 * evaluated ON THE DEVICE. A CPU replay that compiles the same header must get the same bits; the
 * parity tests check exactly that (every rotation and every spawned direction passes through it). */
int fw_device_sincos(fw_context *ctx, const float *x, uint64_t n, float *sin_out, float *cos_out);

int fw_create(const fw_config *cfg, fw_context **out_ctx);
int fw_destroy(fw_context *ctx);

/* sync_spawner_data for one changed / new spawner (src/core.rs:343-365): (re)create its
 * emitter state (last_emission 0, time_passed 0, enabled = starts_enabled) and DROP all of its
 * particles. spawner_key is any caller-chosen id (the shim uses Entity::to_bits() low word). */
int fw_spawner_reset(fw_context *ctx, uint32_t spawner_key,
                     const fw_particle_settings *particle_settings, uint32_t n_particle_types,
                     const fw_emission_settings *emission_settings, uint32_t n_emitters,
                     uint32_t starts_enabled);
/* entity despawned */
int fw_spawner_remove(fw_context *ctx, uint32_t spawner_key);

/* replace the collider set of the spatial query that particle_collision casts rays against
 * (SpatialQuery::cast_ray, src/core.rs:756-765): the world transforms of avian's colliders as of
 * the last physics step. Re-sending a set of the same size (moving colliders) is asynchronous --
 * an upload and a BVH rebuild ordered between the frames around it; another size reallocates. */
int fw_set_colliders(fw_context *ctx, const fw_collider *colliders, uint32_t n);

/* one tick of (spawn_particles ; update_particles) over every spawner of the context.
 * Asynchronous on the context's stream. dt = Res<Time>::delta_secs() (src/core.rs:413,594): finite
 * and >= 0, anything else is FW_ERR_INVALID_ARGUMENT. */
int fw_frame(fw_context *ctx, float dt, const fw_spawner_frame_input *inputs, uint32_t n_inputs);

/* wait for all queued frames; reports (once) what the device flagged since the last report */
int fw_sync(fw_context *ctx);
/* the same report without waiting: *flags = FW_DEVICE_* bits raised by frames that have already
 * completed (and not yet reported); returns FW_OK. A shim calls it once per tick after fw_frame, so a
 * dropped spawn is logged within a few frames instead of at the next synchronisation. */
#define FW_DEVICE_RING_OVERFLOW 1u   /* a ring was full: spawns were dropped (the host grows rings before that can happen) */
#define FW_DEVICE_LOOKBACK_SMALL 2u  /* look-back table too small */
#define FW_DEVICE_NESTED_CAP 4u      /* a parent wanted to emit more children than the planned per-parent bound */
int fw_poll_device_errors(fw_context *ctx, uint32_t *flags);

/* data.particles[i].len() for every particle type of a spawner (synchronises) */
int fw_counts(fw_context *ctx, uint32_t spawner_key, uint32_t *out_counts, uint32_t n_types);
/* same for every stream of the context in creation order: (spawner_key, type, count) triples
 * are returned as three parallel arrays; *n_streams receives the number written */
int fw_counts_all(fw_context *ctx, uint32_t *out_keys, uint32_t *out_types, uint32_t *out_counts,
                  uint32_t cap, uint32_t *n_streams);
int fw_spawner_status_get(fw_context *ctx, uint32_t spawner_key, fw_spawner_status *out);
/* data.finished_notified = true (src/core.rs:685), after the shim triggered the event */
int fw_spawner_mark_finished_notified(fw_context *ctx, uint32_t spawner_key);

int fw_stream_layout_get(fw_context *ctx, uint32_t spawner_key, uint32_t type, fw_stream_layout *out);

/* host mirror of data.particles[type] in the reference's Vec order (synchronises) */
int fw_read_particles(fw_context *ctx, uint32_t spawner_key, uint32_t type,
                      fw_particle_data *out, uint64_t cap, uint64_t *n);
/* replace data.particles[type] (the field is pub in the reference; also used by tests) */
int fw_write_particles(fw_context *ctx, uint32_t spawner_key, uint32_t type,
                       const fw_particle_data *in, uint64_t n);
/* ParticleInstance rows of one stream, in Vec order */
int fw_read_instances(fw_context *ctx, uint32_t spawner_key, uint32_t type,
                      fw_particle_instance *out, uint64_t cap, uint64_t *n);
/* particles destroyed by the last frame (only for types with capture_destroyed) */
int fw_read_destroyed(fw_context *ctx, uint32_t spawner_key, uint32_t type,
                      fw_particle_data *out, uint64_t cap, uint64_t *n);
/* min / max of position -/+ scale over all particle types of a spawner
 * (src/render.rs:681-692); *empty = 1 when it has no particles */
int fw_read_aabb(fw_context *ctx, uint32_t spawner_key, float out_min[3], float out_max[3],
                 uint32_t *empty);

/* gather the live ParticleInstance rows of every stream (creation order, Vec order inside a
 * stream) into ONE contiguous DEVICE buffer owned by the caller -- the per-GPU instance
 * buffer that a multi-GPU render extract all-gathers. Asynchronous on the context's stream;
 * *n_rows is valid after fw_sync. */
int fw_pack_instances_device(fw_context *ctx, void *device_dst, uint64_t cap_rows,
                             uint64_t *n_rows);
/* ---- multi-GPU render extract over NVLink peer memory. One context per
 * GPU (same process or one process per GPU); spawners are sharded across them and never interact
 * (src/core.rs:583-589), so the only exchange is the hand-off of the instance rows that
 * extract_firework_components (src/render.rs:439-461) wants on the GPU that draws. Every rank owns
 * a gather buffer [header | n_ranks regions of cap_rows_per_rank rows]; fw_gather_instances packs
 * this rank's live rows and STORES them into its region of every rank's buffer (peer-mapped over
 * NVLink / NVSwitch), bracketed by device-side ready / landed flags -- an all-gather-v with no
 * host synchronisation and no staging copy. */
#define FW_GATHER_MAX_RANKS 16
typedef struct fw_gather_handle { /* how another context / process maps a rank's gather buffer */
    uint8_t ipc[64];              /* cudaIpcMemHandle_t */
    uint64_t address;             /* device address in the creating process (same-process peers) */
    uint64_t bytes;
    int32_t device;               /* CUDA ordinal in the creating process */
    int32_t pid;
} fw_gather_handle;
/* allocate this rank's gather buffer; *out is what the other ranks pass to fw_gather_connect */
int fw_gather_create(fw_context *ctx, uint32_t n_ranks, uint32_t my_rank, uint64_t cap_rows_per_rank,
                     fw_gather_handle *out);
/* map every rank's buffer: handles[r] = rank r's fw_gather_create output (handles[my_rank] ignored) */
int fw_gather_connect(fw_context *ctx, const fw_gather_handle *handles, uint32_t n_handles);
/* collective: every rank calls it once per extract. Asynchronous on the context's stream. */
int fw_gather_instances(fw_context *ctx);
/* waits for the last fw_gather_instances; rank r's rows_per_rank[r] rows (same order as
 * fw_pack_instances_device) start at *device_rows + r * *region_stride_rows * 64 bytes */
int fw_gather_result(fw_context *ctx, void **device_rows, uint64_t *rows_per_rank, uint32_t n_ranks,
                     uint64_t *region_stride_rows);
int fw_gather_destroy(fw_context *ctx);

/* total live particles over the context (synchronises) */
int fw_total_live(fw_context *ctx, uint64_t *out);

/* profile of the most recent frame (FW_FLAG_PROFILE; synchronises) and the running sums since
 * the last fw_profile_reset */
/* Frame accounting. Counts (particles, launches, bytes) are accumulated for every frame;
 * kernel times only for frames submitted while profiling is on (FW_FLAG_PROFILE at creation or
 * fw_set_profiling): those frames are launched kernel by kernel with CUDA events around each
 * kernel, all other frames may be replayed as CUDA graphs. */
int fw_set_profiling(fw_context *ctx, uint32_t on);
int fw_profile_last(fw_context *ctx, fw_frame_profile *out);
int fw_profile_sum(fw_context *ctx, fw_frame_profile *out, uint32_t *n_frames);
int fw_profile_reset(fw_context *ctx);

/* render extract: the live ParticleInstance rows of every stream (same order as
 * fw_pack_instances_device) copied into ONE caller-owned HOST buffer (pinned memory makes it a
 * single DMA). Synchronises. */
int fw_extract_instances(fw_context *ctx, void *host_dst, uint64_t cap_rows, uint64_t *n_rows);

/* ---- render hand-off that does not stall the simulation (reference consumer: extract_firework_components,
 * src/render.rs:439-461, and the per-frame vertex-buffer upload, :568-584).
 * fw_extract_begin: pack the live ParticleInstance rows of the listed spawners (spawner_keys = NULL: of
 * every spawner; a renderer passes the spawners whose AABB -- fw_read_aabb -- survived its frustum
 * culling, src/render.rs:677-703) and start copying them into the caller's HOST buffer (pinned memory
 * makes it one DMA) on a stream of its own: the call returns at once, later fw_frame calls run while
 * the copy is in flight (two staging buffers alternate, so one extract may be outstanding while the
 * next one is packed). fw_extract_wait: wait for the oldest outstanding extract; *n_rows rows have
 * landed, stream_first_rows[k] = first row of the k-th listed stream (creation order, Vec order
 * inside a stream; cap_streams entries are written at most, *n_streams = how many there are). */
int fw_extract_begin(fw_context *ctx, const uint32_t *spawner_keys, uint32_t n_keys, void *host_dst, uint64_t cap_rows);
int fw_extract_wait(fw_context *ctx, uint64_t *n_rows, uint64_t *stream_first_rows, uint32_t cap_streams,
                    uint32_t *n_streams);
/* zero-copy hand-off to another API or process: pack the live rows of every stream into a device
 * allocation made with the CUDA virtual-memory API and export it as a POSIX file descriptor
 * (cuMemExportToShareableHandle) -- what VK_KHR_external_memory_fd / a wgpu external-memory import
 * takes. *bytes = size of the allocation, *n_rows = rows packed (64 bytes each, same order as
 * fw_pack_instances_device). The descriptor is the caller's to close; the allocation lives until the
 * next fw_export_instances_fd or fw_destroy. Synchronises. */
int fw_export_instances_fd(fw_context *ctx, int32_t *fd, uint64_t *bytes, uint64_t *n_rows);
/* the importing side of the above for a consumer without its own CUDA code (and for the tests): map
 * the descriptor on `device`, copy n_rows rows to host_dst, unmap. No context needed. */
int fw_import_instances_fd(int32_t device, int32_t fd, uint64_t bytes, uint64_t n_rows, void *host_dst);

/* CUDA-event stopwatch on the context's stream: record marker `slot` (0..15) now; elapsed
 * milliseconds between two recorded markers (synchronises on the later one). */
int fw_event_record(fw_context *ctx, uint32_t slot);
int fw_event_elapsed_ms(fw_context *ctx, uint32_t slot_begin, uint32_t slot_end, float *out_ms);

/* the stream every launch of the context goes to (cudaStream_t) */
void *fw_stream_handle(fw_context *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FIREWORK_B200_H */
