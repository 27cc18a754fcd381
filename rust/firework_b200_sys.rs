//! firework_b200_sys — raw FFI declarations of `include/firework_b200.h`.
//!
//! UNCOMPILED SOURCE: there is no Rust toolchain in the build image. The
//! `#[repr(C)]` layouts below restate the C header field by field. `tests/test_abi.py` parses THIS
//! file and the header mechanically -- every export (name, arity, argument types), every struct (field
//! names, order, types -> offsets by the C layout rules) -- and compares both with each other and with
//! the compiled library (`fw_abi_sizeof`, `fw_abi_offsetof`); `shim.rs::check_layouts()` asserts the
//! same from Rust at plugin start-up, so a drift fails loudly instead of corrupting memory.
//!
//! Link with `cargo:rustc-link-lib=dylib=firework_b200` (build.rs) and ship
//! `libfirework_b200.so` beside the executable.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const FW_ABI_VERSION: u32 = 2;
pub const FW_MAX_KNOTS: usize = 32;
pub const FW_MAX_EXCLUDED: usize = 8;
pub const FW_NO_KEY: u32 = 0xFFFF_FFFF;
pub const FW_FLAG_PROFILE: u32 = 1;
pub const FW_FLAG_NO_GRAPHS: u32 = 2;
pub const FW_FLAG_NO_CONCURRENT_SPAWN: u32 = 4;
pub const FW_LAYOUT_COMPACTING: u32 = 1;
pub const FW_LAYOUT_COLLIDES: u32 = 2;
pub const FW_LAYOUT_ROTATES: u32 = 4;
pub const FW_STORE_BASE_COLOR: u32 = 1;
pub const FW_STORE_EMISSIVE_COLOR: u32 = 2;
pub const FW_STORE_SCALE: u32 = 4;
pub const FW_STORE_LIFETIME: u32 = 8;
pub const FW_GATHER_MAX_RANKS: usize = 16;
pub const FW_DEVICE_RING_OVERFLOW: u32 = 1;
pub const FW_DEVICE_LOOKBACK_SMALL: u32 = 2;
pub const FW_DEVICE_NESTED_CAP: u32 = 4;

pub const FW_OK: c_int = 0;
pub const FW_ERR_INVALID_ARGUMENT: c_int = 1;
pub const FW_ERR_NO_DEVICE: c_int = 2;
pub const FW_ERR_CUDA: c_int = 3;
pub const FW_ERR_OUT_OF_MEMORY: c_int = 4;
pub const FW_ERR_UNKNOWN_SPAWNER: c_int = 5;
pub const FW_ERR_BUFFER_TOO_SMALL: c_int = 6;
pub const FW_ERR_UNSUPPORTED: c_int = 7;
pub const FW_ERR_INTERNAL: c_int = 8;

pub const FW_CURVE_CONSTANT: u32 = 0;
pub const FW_CURVE_EVEN: u32 = 1;
pub const FW_CURVE_UNEVEN: u32 = 2;
pub const FW_PACING_ONE_SHOT: u32 = 0;
pub const FW_PACING_ON_DEMAND: u32 = 1;
pub const FW_PACING_COUNT_OVER_DURATION: u32 = 2;
pub const FW_MODE_GLOBAL: u32 = 0;
pub const FW_MODE_NESTED: u32 = 1;
pub const FW_SHAPE_POINT: u32 = 0;
pub const FW_SHAPE_SPHERE: u32 = 1;
pub const FW_SHAPE_CIRCLE: u32 = 2;
pub const FW_COLLIDER_CUBOID: u32 = 0;
pub const FW_COLLIDER_SPHERE: u32 = 1;
pub const FW_COLLIDER_CYLINDER: u32 = 2; // half_extents = (radius, height / 2, -), axis +Y
pub const FW_COLLIDER_CONE: u32 = 3; // half_extents = (radius, height / 2, -), apex at +Y
pub const FW_COLLIDER_CAPSULE: u32 = 4; // half_extents = (radius, length / 2, -), axis +Y

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fw_rand_f32 {
    pub min: f32,
    pub max: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fw_rand_vec3 {
    pub magnitude: fw_rand_f32,
    pub direction: [f32; 3],
    pub spread: f32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_curve_f32 {
    pub kind: u32,
    pub n: u32,
    pub times: [f32; FW_MAX_KNOTS],
    pub values: [f32; FW_MAX_KNOTS],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_gradient {
    pub kind: u32,
    pub n: u32,
    pub times: [f32; FW_MAX_KNOTS],
    pub colors: [[f32; 4]; FW_MAX_KNOTS],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fw_collision_settings {
    pub enabled: u32,
    pub restitution: f32,
    pub friction: f32,
    pub destroy_on_collision: u32,
    pub filter_mask: u32,
    pub n_excluded: u32,
    pub excluded_keys: [u32; FW_MAX_EXCLUDED],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_particle_settings {
    pub lifetime: fw_rand_f32,
    pub scale_curve: fw_curve_f32,
    pub initial_scale: fw_rand_f32,
    pub acceleration: [f32; 3],
    pub angular_acceleration: [f32; 3],
    pub linear_drag: f32,
    pub angular_drag: f32,
    pub base_color: fw_gradient,
    pub emissive_color: fw_gradient,
    pub pbr: u32,
    pub collision: fw_collision_settings,
    pub capture_destroyed: u32,
    pub capacity_hint: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_emission_settings {
    pub particle_index: u32,
    pub pacing_kind: u32,
    pub one_shot_count: u64,
    pub count: f32,
    pub duration: f32,
    pub offset_start: f32,
    pub offset_end: f32,
    pub mode: u32,
    pub target_particle_type: u32,
    pub shape_kind: u32,
    pub shape_radius: f32,
    pub shape_normal: [f32; 3],
    pub initial_velocity: fw_rand_vec3,
    pub initial_velocity_radial: fw_rand_f32,
    pub inherit_parent_velocity: u32,
    pub initial_rotation: [f32; 4],
    pub initial_angular_velocity: fw_rand_vec3,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_spawner_frame_input {
    pub spawner_key: u32,
    pub origin_translation: [f32; 3],
    pub origin_rotation: [f32; 4],
    pub parent_velocity: [f32; 3],
    pub modifier_scale: f32,
    pub modifier_speed: f32,
    pub queue_particles: u32,
}

/// `ParticleData` (src/core.rs:305-321) as a POD row, 104 bytes.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_particle_data {
    pub position: [f32; 3],
    pub velocity: [f32; 3],
    pub rotation: [f32; 4],
    pub angular_velocity: [f32; 3],
    pub initial_scale: f32,
    pub scale: f32,
    pub age: f32,
    pub lifetime: f32,
    pub base_color: [f32; 4],
    pub emissive_color: [f32; 4],
    pub pbr: u32,
}

/// `ParticleInstance` (src/render.rs:95-103), 64 bytes — bytemuck-castable to the crate's own type.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_particle_instance {
    pub position: [f32; 3],
    pub scale: f32,
    pub rotation: [f32; 4],
    pub base_color: [f32; 4],
    pub emissive_color: [f32; 4],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_collider {
    pub kind: u32,
    pub layers: u32,
    pub key: u32,
    pub half_extents: [f32; 3],
    pub translation: [f32; 3],
    pub rotation: [f32; 4],
}

#[repr(C)]
pub struct fw_config {
    pub abi_version: u32,
    pub device: i32,
    pub seed: u64,
    pub external_stream: *mut c_void,
    pub flags: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fw_spawner_status {
    pub active: u32,
    pub all_empty: u32,
    pub finished: u32,
    pub finished_notified: u32,
    pub live_particles: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fw_frame_profile {
    pub plan_ms: f32,
    pub spawn_ms: f32,
    pub update_ms: f32,
    pub total_ms: f32,
    pub kernel_launches: u32,
    pub timed_frames: u32,
    pub particles_updated: u64,
    pub particles_spawned: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
}

/// what the library keeps per particle for one stream (`fw_stream_layout_get`)
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct fw_stream_layout {
    pub variant: u32,
    pub flags: u32,
    pub bytes_read: u32,
    pub bytes_written: u32,
    pub bytes_count_pass: u32,
    pub capacity: u32,
}

#[repr(C)]
pub struct fw_context {
    _private: [u8; 0],
}

/// how another context / process maps a rank's gather buffer (include/firework_b200.h)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct fw_gather_handle {
    pub ipc: [u8; 64],
    pub address: u64,
    pub bytes: u64,
    pub device: i32,
    pub pid: i32,
}
impl Default for fw_gather_handle {
    fn default() -> Self {
        Self { ipc: [0; 64], address: 0, bytes: 0, device: 0, pid: 0 }
    }
}

extern "C" {
    pub fn fw_last_global_error() -> *const c_char;
    pub fn fw_last_error(ctx: *const fw_context) -> *const c_char;
    pub fn fw_abi_version() -> u32;
    pub fn fw_abi_sizeof(struct_name: *const c_char) -> u32;
    pub fn fw_abi_offsetof(struct_name: *const c_char, field_name: *const c_char) -> u32;
    // host-side logic, callable without a device
    pub fn fw_host_emission_count(time_passed_in_cycle: f32, last_emission: f32, cycle_duration: f32, offset_start: f32,
        offset_end: f32, particles_per_cycle: f32, times: *mut u64, next_last_emission: *mut f32) -> c_int;
    pub fn fw_host_build_broadphase(colliders: *const fw_collider, n: u32, out: *mut c_void, cap_bytes: u64, n_bytes: *mut u64) -> c_int;
    pub fn fw_device_sincos(ctx: *mut fw_context, x: *const f32, n: u64, sin_out: *mut f32, cos_out: *mut f32) -> c_int;
    pub fn fw_create(cfg: *const fw_config, out_ctx: *mut *mut fw_context) -> c_int;
    pub fn fw_destroy(ctx: *mut fw_context) -> c_int;
    pub fn fw_spawner_reset(
        ctx: *mut fw_context,
        spawner_key: u32,
        particle_settings: *const fw_particle_settings,
        n_particle_types: u32,
        emission_settings: *const fw_emission_settings,
        n_emitters: u32,
        starts_enabled: u32,
    ) -> c_int;
    pub fn fw_spawner_remove(ctx: *mut fw_context, spawner_key: u32) -> c_int;
    pub fn fw_set_colliders(ctx: *mut fw_context, colliders: *const fw_collider, n: u32) -> c_int;
    pub fn fw_frame(ctx: *mut fw_context, dt: f32, inputs: *const fw_spawner_frame_input, n_inputs: u32) -> c_int;
    pub fn fw_sync(ctx: *mut fw_context) -> c_int;
    pub fn fw_poll_device_errors(ctx: *mut fw_context, flags: *mut u32) -> c_int;
    pub fn fw_counts(ctx: *mut fw_context, spawner_key: u32, out_counts: *mut u32, n_types: u32) -> c_int;
    pub fn fw_counts_all(ctx: *mut fw_context, out_keys: *mut u32, out_types: *mut u32, out_counts: *mut u32, cap: u32, n_streams: *mut u32) -> c_int;
    pub fn fw_spawner_status_get(ctx: *mut fw_context, spawner_key: u32, out: *mut fw_spawner_status) -> c_int;
    pub fn fw_spawner_mark_finished_notified(ctx: *mut fw_context, spawner_key: u32) -> c_int;
    pub fn fw_stream_layout_get(ctx: *mut fw_context, spawner_key: u32, ty: u32, out: *mut fw_stream_layout) -> c_int;
    pub fn fw_read_particles(ctx: *mut fw_context, spawner_key: u32, ty: u32, out: *mut fw_particle_data, cap: u64, n: *mut u64) -> c_int;
    pub fn fw_write_particles(ctx: *mut fw_context, spawner_key: u32, ty: u32, rows: *const fw_particle_data, n: u64) -> c_int;
    pub fn fw_read_instances(ctx: *mut fw_context, spawner_key: u32, ty: u32, out: *mut fw_particle_instance, cap: u64, n: *mut u64) -> c_int;
    pub fn fw_read_destroyed(ctx: *mut fw_context, spawner_key: u32, ty: u32, out: *mut fw_particle_data, cap: u64, n: *mut u64) -> c_int;
    pub fn fw_read_aabb(ctx: *mut fw_context, spawner_key: u32, out_min: *mut [f32; 3], out_max: *mut [f32; 3], empty: *mut u32) -> c_int;
    pub fn fw_extract_instances(ctx: *mut fw_context, host_dst: *mut c_void, cap_rows: u64, n_rows: *mut u64) -> c_int;
    pub fn fw_total_live(ctx: *mut fw_context, out: *mut u64) -> c_int;
    pub fn fw_pack_instances_device(ctx: *mut fw_context, device_dst: *mut c_void, cap_rows: u64, n_rows: *mut u64) -> c_int;
    // multi-GPU render extract over NVLink peer memory (one context per GPU)
    pub fn fw_gather_create(ctx: *mut fw_context, n_ranks: u32, my_rank: u32, cap_rows_per_rank: u64, out: *mut fw_gather_handle) -> c_int;
    pub fn fw_gather_connect(ctx: *mut fw_context, handles: *const fw_gather_handle, n_handles: u32) -> c_int;
    pub fn fw_gather_instances(ctx: *mut fw_context) -> c_int;
    pub fn fw_gather_result(ctx: *mut fw_context, device_rows: *mut *mut c_void, rows_per_rank: *mut u64, n_ranks: u32, region_stride_rows: *mut u64) -> c_int;
    pub fn fw_gather_destroy(ctx: *mut fw_context) -> c_int;
    // render hand-off without a host round trip / without blocking the simulation
    pub fn fw_extract_begin(ctx: *mut fw_context, spawner_keys: *const u32, n_keys: u32, host_dst: *mut c_void, cap_rows: u64) -> c_int;
    pub fn fw_extract_wait(ctx: *mut fw_context, n_rows: *mut u64, stream_first_rows: *mut u64, cap_streams: u32, n_streams: *mut u32) -> c_int;
    pub fn fw_export_instances_fd(ctx: *mut fw_context, fd: *mut i32, bytes: *mut u64, n_rows: *mut u64) -> c_int;
    pub fn fw_import_instances_fd(device: i32, fd: i32, bytes: u64, n_rows: u64, host_dst: *mut c_void) -> c_int;
    // frame accounting and timing
    pub fn fw_set_profiling(ctx: *mut fw_context, on: u32) -> c_int;
    pub fn fw_profile_last(ctx: *mut fw_context, out: *mut fw_frame_profile) -> c_int;
    pub fn fw_profile_sum(ctx: *mut fw_context, out: *mut fw_frame_profile, n_frames: *mut u32) -> c_int;
    pub fn fw_profile_reset(ctx: *mut fw_context) -> c_int;
    pub fn fw_event_record(ctx: *mut fw_context, slot: u32) -> c_int;
    pub fn fw_event_elapsed_ms(ctx: *mut fw_context, slot_begin: u32, slot_end: u32, out_ms: *mut f32) -> c_int;
    pub fn fw_stream_handle(ctx: *mut fw_context) -> *mut c_void;
}
