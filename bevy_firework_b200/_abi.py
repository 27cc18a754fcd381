"""ctypes mirror of ``include/firework_b200.h`` (POD layouts and enum values only).

Everything here is a byte-for-byte restatement of the C header; ``tests/test_abi.py`` checks the
struct sizes against the compiled library (``fw_abi_sizeof``) so the two cannot drift.
"""
from __future__ import annotations

import ctypes as C

FW_ABI_VERSION = 2
FW_MAX_KNOTS = 32
FW_MAX_EXCLUDED = 8
FW_NO_KEY = 0xFFFFFFFF

# enum fw_status
FW_OK = 0
FW_ERR_INVALID_ARGUMENT = 1
FW_ERR_NO_DEVICE = 2
FW_ERR_CUDA = 3
FW_ERR_OUT_OF_MEMORY = 4
FW_ERR_UNKNOWN_SPAWNER = 5
FW_ERR_BUFFER_TOO_SMALL = 6
FW_ERR_UNSUPPORTED = 7
FW_ERR_INTERNAL = 8

FW_CURVE_CONSTANT, FW_CURVE_EVEN, FW_CURVE_UNEVEN = 0, 1, 2
FW_PACING_ONE_SHOT, FW_PACING_ON_DEMAND, FW_PACING_COUNT_OVER_DURATION = 0, 1, 2
FW_MODE_GLOBAL, FW_MODE_NESTED = 0, 1
FW_SHAPE_POINT, FW_SHAPE_SPHERE, FW_SHAPE_CIRCLE = 0, 1, 2
FW_TRANSFORM_GLOBAL, FW_TRANSFORM_LOCAL = 0, 1
FW_COLLIDER_CUBOID, FW_COLLIDER_SPHERE, FW_COLLIDER_CYLINDER, FW_COLLIDER_CONE, FW_COLLIDER_CAPSULE = 0, 1, 2, 3, 4
FW_FLAG_PROFILE = 1
FW_FLAG_NO_GRAPHS = 2
FW_FLAG_NO_CONCURRENT_SPAWN = 4

f32 = C.c_float
u32 = C.c_uint32
u64 = C.c_uint64
i32 = C.c_int32


class fw_rand_f32(C.Structure):
    _fields_ = [("min", f32), ("max", f32)]


class fw_rand_vec3(C.Structure):
    _fields_ = [("magnitude", fw_rand_f32), ("direction", f32 * 3), ("spread", f32)]


class fw_curve_f32(C.Structure):
    _fields_ = [("kind", u32), ("n", u32), ("times", f32 * FW_MAX_KNOTS), ("values", f32 * FW_MAX_KNOTS)]


class fw_gradient(C.Structure):
    _fields_ = [("kind", u32), ("n", u32), ("times", f32 * FW_MAX_KNOTS),
                ("colors", (f32 * 4) * FW_MAX_KNOTS)]


class fw_collision_settings(C.Structure):
    _fields_ = [("enabled", u32), ("restitution", f32), ("friction", f32),
                ("destroy_on_collision", u32), ("filter_mask", u32),
                ("n_excluded", u32), ("excluded_keys", u32 * FW_MAX_EXCLUDED)]


class fw_particle_settings(C.Structure):
    _fields_ = [
        ("lifetime", fw_rand_f32),
        ("scale_curve", fw_curve_f32),
        ("initial_scale", fw_rand_f32),
        ("acceleration", f32 * 3),
        ("angular_acceleration", f32 * 3),
        ("linear_drag", f32),
        ("angular_drag", f32),
        ("base_color", fw_gradient),
        ("emissive_color", fw_gradient),
        ("pbr", u32),
        ("collision", fw_collision_settings),
        ("capture_destroyed", u32),
        ("capacity_hint", u32),
    ]


class fw_emission_settings(C.Structure):
    _fields_ = [
        ("particle_index", u32),
        ("pacing_kind", u32),
        ("one_shot_count", u64),
        ("count", f32),
        ("duration", f32),
        ("offset_start", f32),
        ("offset_end", f32),
        ("mode", u32),
        ("target_particle_type", u32),
        ("shape_kind", u32),
        ("shape_radius", f32),
        ("shape_normal", f32 * 3),
        ("initial_velocity", fw_rand_vec3),
        ("initial_velocity_radial", fw_rand_f32),
        ("inherit_parent_velocity", u32),
        ("initial_rotation", f32 * 4),
        ("initial_angular_velocity", fw_rand_vec3),
    ]


class fw_spawner_frame_input(C.Structure):
    _fields_ = [
        ("spawner_key", u32),
        ("origin_translation", f32 * 3),
        ("origin_rotation", f32 * 4),
        ("parent_velocity", f32 * 3),
        ("modifier_scale", f32),
        ("modifier_speed", f32),
        ("queue_particles", u32),
    ]


class fw_particle_data(C.Structure):
    _fields_ = [
        ("position", f32 * 3),
        ("velocity", f32 * 3),
        ("rotation", f32 * 4),
        ("angular_velocity", f32 * 3),
        ("initial_scale", f32),
        ("scale", f32),
        ("age", f32),
        ("lifetime", f32),
        ("base_color", f32 * 4),
        ("emissive_color", f32 * 4),
        ("pbr", u32),
    ]


class fw_particle_instance(C.Structure):
    _fields_ = [("position", f32 * 3), ("scale", f32), ("rotation", f32 * 4),
                ("base_color", f32 * 4), ("emissive_color", f32 * 4)]


class fw_collider(C.Structure):
    _fields_ = [("kind", u32), ("layers", u32), ("key", u32), ("half_extents", f32 * 3),
                ("translation", f32 * 3), ("rotation", f32 * 4)]


class fw_config(C.Structure):
    _fields_ = [("abi_version", u32), ("device", i32), ("seed", u64),
                ("external_stream", C.c_void_p), ("flags", u32), ("reserved", u32)]


class fw_spawner_status(C.Structure):
    _fields_ = [("active", u32), ("all_empty", u32), ("finished", u32),
                ("finished_notified", u32), ("live_particles", u64)]


class fw_frame_profile(C.Structure):
    _fields_ = [("plan_ms", f32), ("spawn_ms", f32), ("update_ms", f32), ("total_ms", f32),
                ("kernel_launches", u32), ("timed_frames", u32),
                ("particles_updated", u64), ("particles_spawned", u64),
                ("h2d_bytes", u64), ("d2h_bytes", u64)]


class fw_stream_layout(C.Structure):
    _fields_ = [("variant", u32), ("flags", u32), ("bytes_read", u32), ("bytes_written", u32),
                ("bytes_count_pass", u32), ("capacity", u32)]


class fw_gather_handle(C.Structure):
    _fields_ = [("ipc", C.c_uint8 * 64), ("address", u64), ("bytes", u64), ("device", C.c_int32), ("pid", C.c_int32)]


# numpy structured dtypes of the two row formats (for zero-copy readback)
def particle_data_dtype():
    import numpy as np

    return np.dtype([
        ("position", np.float32, 3), ("velocity", np.float32, 3), ("rotation", np.float32, 4),
        ("angular_velocity", np.float32, 3), ("initial_scale", np.float32), ("scale", np.float32),
        ("age", np.float32), ("lifetime", np.float32), ("base_color", np.float32, 4),
        ("emissive_color", np.float32, 4), ("pbr", np.uint32),
    ])


def particle_instance_dtype():
    import numpy as np

    return np.dtype([
        ("position", np.float32, 3), ("scale", np.float32), ("rotation", np.float32, 4),
        ("base_color", np.float32, 4), ("emissive_color", np.float32, 4),
    ])


FW_LAYOUT_COMPACTING, FW_LAYOUT_COLLIDES, FW_LAYOUT_ROTATES = 1, 2, 4
FW_STORE_BASE_COLOR, FW_STORE_EMISSIVE_COLOR, FW_STORE_SCALE, FW_STORE_LIFETIME = 1, 2, 4, 8

# every exported symbol of the C ABI: name -> (restype, argtypes)
P = C.POINTER
_ctx = C.c_void_p
EXPORTS = {
    "fw_last_global_error": (C.c_char_p, []),
    "fw_last_error": (C.c_char_p, [_ctx]),
    "fw_abi_version": (u32, []),
    "fw_abi_sizeof": (u32, [C.c_char_p]),
    "fw_abi_offsetof": (u32, [C.c_char_p, C.c_char_p]),
    "fw_host_emission_count": (C.c_int, [f32, f32, f32, f32, f32, f32, P(u64), P(f32)]),
    "fw_host_build_broadphase": (C.c_int, [P(fw_collider), u32, C.c_void_p, u64, P(u64)]),
    "fw_device_sincos": (C.c_int, [_ctx, C.c_void_p, u64, C.c_void_p, C.c_void_p]),
    "fw_create": (C.c_int, [P(fw_config), P(_ctx)]),
    "fw_destroy": (C.c_int, [_ctx]),
    "fw_spawner_reset": (C.c_int, [_ctx, u32, P(fw_particle_settings), u32,
                                   P(fw_emission_settings), u32, u32]),
    "fw_spawner_remove": (C.c_int, [_ctx, u32]),
    "fw_set_colliders": (C.c_int, [_ctx, P(fw_collider), u32]),
    "fw_frame": (C.c_int, [_ctx, f32, P(fw_spawner_frame_input), u32]),
    "fw_sync": (C.c_int, [_ctx]),
    "fw_poll_device_errors": (C.c_int, [_ctx, P(u32)]),
    "fw_counts": (C.c_int, [_ctx, u32, P(u32), u32]),
    "fw_counts_all": (C.c_int, [_ctx, P(u32), P(u32), P(u32), u32, P(u32)]),
    "fw_spawner_status_get": (C.c_int, [_ctx, u32, P(fw_spawner_status)]),
    "fw_spawner_mark_finished_notified": (C.c_int, [_ctx, u32]),
    "fw_stream_layout_get": (C.c_int, [_ctx, u32, u32, P(fw_stream_layout)]),
    "fw_read_particles": (C.c_int, [_ctx, u32, u32, C.c_void_p, u64, P(u64)]),
    "fw_write_particles": (C.c_int, [_ctx, u32, u32, C.c_void_p, u64]),
    "fw_read_instances": (C.c_int, [_ctx, u32, u32, C.c_void_p, u64, P(u64)]),
    "fw_read_destroyed": (C.c_int, [_ctx, u32, u32, C.c_void_p, u64, P(u64)]),
    "fw_read_aabb": (C.c_int, [_ctx, u32, P(f32 * 3), P(f32 * 3), P(u32)]),
    "fw_pack_instances_device": (C.c_int, [_ctx, C.c_void_p, u64, P(u64)]),
    "fw_gather_create": (C.c_int, [_ctx, u32, u32, u64, P(fw_gather_handle)]),
    "fw_gather_connect": (C.c_int, [_ctx, P(fw_gather_handle), u32]),
    "fw_gather_instances": (C.c_int, [_ctx]),
    "fw_gather_result": (C.c_int, [_ctx, P(C.c_void_p), P(u64), u32, P(u64)]),
    "fw_gather_destroy": (C.c_int, [_ctx]),
    "fw_total_live": (C.c_int, [_ctx, P(u64)]),
    "fw_set_profiling": (C.c_int, [_ctx, u32]),
    "fw_profile_last": (C.c_int, [_ctx, P(fw_frame_profile)]),
    "fw_profile_sum": (C.c_int, [_ctx, P(fw_frame_profile), P(u32)]),
    "fw_profile_reset": (C.c_int, [_ctx]),
    "fw_extract_instances": (C.c_int, [_ctx, C.c_void_p, u64, P(u64)]),
    "fw_extract_begin": (C.c_int, [_ctx, P(u32), u32, C.c_void_p, u64]),
    "fw_extract_wait": (C.c_int, [_ctx, P(u64), P(u64), u32, P(u32)]),
    "fw_export_instances_fd": (C.c_int, [_ctx, P(i32), P(u64), P(u64)]),
    "fw_import_instances_fd": (C.c_int, [i32, i32, u64, u64, C.c_void_p]),
    "fw_event_record": (C.c_int, [_ctx, u32]),
    "fw_event_elapsed_ms": (C.c_int, [_ctx, u32, u32, P(f32)]),
    "fw_stream_handle": (C.c_void_p, [_ctx]),
}

# struct name -> ctypes class, for the size cross-check against the compiled library
POD_TYPES = {
    "fw_rand_f32": fw_rand_f32, "fw_rand_vec3": fw_rand_vec3, "fw_curve_f32": fw_curve_f32,
    "fw_gradient": fw_gradient, "fw_collision_settings": fw_collision_settings,
    "fw_particle_settings": fw_particle_settings, "fw_emission_settings": fw_emission_settings,
    "fw_spawner_frame_input": fw_spawner_frame_input, "fw_particle_data": fw_particle_data,
    "fw_particle_instance": fw_particle_instance, "fw_collider": fw_collider,
    "fw_config": fw_config, "fw_spawner_status": fw_spawner_status,
    "fw_frame_profile": fw_frame_profile, "fw_gather_handle": fw_gather_handle,
    "fw_stream_layout": fw_stream_layout,
}
