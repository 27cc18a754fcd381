"""ctypes wrapper of the CPU ORACLE (``oracle/libfw_oracle.so``) -- test infrastructure.

Only ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and
``__graft_entry__.smoke()`` import this module. The product package never does.

``OracleWorld`` has the same method names as ``bevy_firework_b200._native.Engine`` so one test
body can drive both and compare.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from bevy_firework_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfw_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``oracle/Makefile``)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("fw_oracle.c", "fw_oracle.h", "../include/firework_b200.h", "../include/fw_sincos.h")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    P, f32, u32, u64, vp = C.POINTER, C.c_float, C.c_uint32, C.c_uint64, C.c_void_p
    sig = {
        "fwo_create": (vp, [u64]),
        "fwo_destroy": (None, [vp]),
        "fwo_spawner_reset": (C.c_int, [vp, u32, P(_abi.fw_particle_settings), u32,
                                        P(_abi.fw_emission_settings), u32, u32]),
        "fwo_spawner_remove": (C.c_int, [vp, u32]),
        "fwo_set_colliders": (None, [vp, P(_abi.fw_collider), u32]),
        "fwo_frame": (None, [vp, f32, P(_abi.fw_spawner_frame_input), u32, u32]),
        "fwo_spawn_only": (None, [vp, f32, P(_abi.fw_spawner_frame_input), u32]),
        "fwo_update_only": (None, [vp, f32, u32]),
        "fwo_count": (u64, [vp, u32, u32]),
        "fwo_total_live": (u64, [vp]),
        "fwo_read_particles": (C.c_int, [vp, u32, u32, vp, u64, P(u64)]),
        "fwo_write_particles": (C.c_int, [vp, u32, u32, vp, u64]),
        "fwo_read_destroyed": (C.c_int, [vp, u32, u32, vp, u64, P(u64)]),
        "fwo_status": (C.c_int, [vp, u32, P(_abi.fw_spawner_status)]),
        "fwo_mark_finished_notified": (C.c_int, [vp, u32]),
        "fwo_read_aabb": (C.c_int, [vp, u32, P(f32 * 3), P(f32 * 3), P(u32)]),
        "fwo_compute_emission_count": (None, [f32, f32, f32, f32, f32, f32, P(u64), P(f32)]),
        "fwo_sample_curve": (f32, [P(_abi.fw_curve_f32), f32]),
        "fwo_sample_gradient": (None, [P(_abi.fw_gradient), f32, P(f32 * 4)]),
        "fwo_philox4x32_10": (None, [P(u32 * 4), P(u32 * 2), P(u32 * 4)]),
        "fwo_uniform": (f32, [u64, u32, u32, u64, u32]),
        "fwo_generate_point": (None, [P(_abi.fw_emission_settings), f32, f32, f32, P(f32 * 3)]),
        "fwo_rand_vec3": (None, [P(_abi.fw_rand_vec3), f32, f32, f32, P(f32 * 3)]),
        "fwo_cast_ray": (C.c_int, [P(_abi.fw_collider), u32, u32, P(f32 * 3), P(f32 * 3), f32,
                                   P(f32), P(f32 * 3), P(u32)]),
        "fwo_cast_ray_culled": (C.c_int, [P(_abi.fw_collider), u32, u32, P(f32 * 3), P(f32 * 3), f32,
                                          P(f32), P(f32 * 3), P(u32)]),
        "fwo_set_cull": (None, [vp, C.c_int]),
        "fwo_particle_collision": (None, [P(_abi.fw_collider), u32, P(_abi.fw_collision_settings),
                                          P(f32 * 3), P(f32 * 3), f32, P(u32)]),
        "fwo_quat_from_scaled_axis": (None, [P(f32 * 3), P(f32 * 4)]),
        "fwo_quat_mul": (None, [P(f32 * 4), P(f32 * 4), P(f32 * 4)]),
        "fwo_quat_from_rotation_arc": (None, [P(f32 * 3), P(f32 * 3), P(f32 * 4)]),
        "fwo_quat_mul_vec3": (None, [P(f32 * 4), P(f32 * 3), P(f32 * 3)]),
        "fwo_vec3_normalize_or_zero": (None, [P(f32 * 3), P(f32 * 3)]),
        "fwo_vec3_project_onto": (None, [P(f32 * 3), P(f32 * 3), P(f32 * 3)]),
        "fwo_vec3_reject_from": (None, [P(f32 * 3), P(f32 * 3), P(f32 * 3)]),
        "fwo_pitch_yaw_to_unit_vec": (None, [f32, f32, P(f32 * 3)]),
        "fwo_sincosf": (None, [f32, P(f32), P(f32)]),
        "fwo_sincosf_array": (None, [vp, u64, vp, vp]),
        "fwo_rem_euclid": (f32, [f32, f32]),
        "fwo_div_euclid": (f32, [f32, f32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


# ---- pure helpers (known-answer tests)
def compute_emission_count(t, last, cycle, start, end, count):
    n, nl = C.c_uint64(), C.c_float()
    lib().fwo_compute_emission_count(t, last, cycle, start, end, count, C.byref(n), C.byref(nl))
    return n.value, nl.value


def sample_curve(curve_pod, t: float) -> float:
    return lib().fwo_sample_curve(C.byref(curve_pod), t)


def sample_gradient(grad_pod, t: float):
    out = (C.c_float * 4)()
    lib().fwo_sample_gradient(C.byref(grad_pod), t, C.byref(out))
    return tuple(out)


def philox4x32_10(ctr: Sequence[int], key: Sequence[int]):
    c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
    lib().fwo_philox4x32_10(C.byref(c), C.byref(k), C.byref(o))
    return tuple(o)


def sincosf(x):
    """include/fw_sincos.h as compiled into the oracle: scalar -> (sin, cos); float32 array -> two arrays"""
    if np.ndim(x) == 0:
        s, c = C.c_float(), C.c_float()
        lib().fwo_sincosf(float(x), C.byref(s), C.byref(c))
        return s.value, c.value
    xs = np.ascontiguousarray(x, dtype=np.float32)
    s, c = np.empty_like(xs), np.empty_like(xs)
    lib().fwo_sincosf_array(xs.ctypes.data, xs.size, s.ctypes.data, c.ctypes.data)
    return s, c


def uniform(seed, spawner_key, emitter, serial, draw) -> float:
    return lib().fwo_uniform(seed, spawner_key, emitter, serial, draw)


def generate_point(emission_pod, u0, u1, u2):
    out = (C.c_float * 3)()
    lib().fwo_generate_point(C.byref(emission_pod), u0, u1, u2, C.byref(out))
    return tuple(out)


def rand_vec3(rv_pod, ua, ur, um):
    out = (C.c_float * 3)()
    lib().fwo_rand_vec3(C.byref(rv_pod), ua, ur, um, C.byref(out))
    return tuple(out)


def _colliders_array(colliders):
    arr = (_abi.fw_collider * max(len(colliders), 1))()
    for i, c in enumerate(colliders):
        arr[i] = c
    return arr


def cast_ray(colliders, origin, direction, max_distance, filter_mask=0xFFFFFFFF, culled=False):
    """culled: through the test helper's conservative boxes (must equal the brute-force loop)"""
    arr = colliders if isinstance(colliders, C.Array) else _colliders_array(colliders)
    o, d = (C.c_float * 3)(*origin), (C.c_float * 3)(*direction)
    dist, nrm, idx = C.c_float(), (C.c_float * 3)(), C.c_uint32()
    fn = lib().fwo_cast_ray_culled if culled else lib().fwo_cast_ray
    hit = fn(arr, len(colliders), filter_mask, C.byref(o), C.byref(d), max_distance,
             C.byref(dist), C.byref(nrm), C.byref(idx))
    return (dist.value, tuple(nrm), idx.value) if hit else None


def particle_collision(colliders, cs_pod, pos, vel, delta):
    arr = _colliders_array(colliders)
    p, v, sd = (C.c_float * 3)(*pos), (C.c_float * 3)(*vel), C.c_uint32()
    lib().fwo_particle_collision(arr, len(colliders), C.byref(cs_pod), C.byref(p), C.byref(v), delta,
                                 C.byref(sd))
    return tuple(p), tuple(v), bool(sd.value)


def frame_inputs_array(inputs):
    arr = (_abi.fw_spawner_frame_input * max(len(inputs), 1))()
    for i, x in enumerate(inputs):
        arr[i] = x
    return arr


class OracleWorld:
    """The reference's schedule played on the CPU restatement."""

    def __init__(self, seed: int = 0x00F12E00, n_threads: int = 1, cull: bool = False):
        """cull: TEST HELPER for the full-size collision scenes (conservative boxes in front of the
        brute-force ray loop; same results). Never used by the timed CPU baseline."""
        self._L = lib()
        self._w = self._L.fwo_create(seed)
        if cull:
            self._L.fwo_set_cull(self._w, 1)
        self.n_threads = n_threads
        self._n_types = {}

    def close(self):
        if self._w:
            self._L.fwo_destroy(self._w)
            self._w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def spawner_reset(self, key, ps, n_types, es, n_emitters, starts_enabled=True):
        rc = self._L.fwo_spawner_reset(self._w, key, ps, n_types, es, n_emitters, 1 if starts_enabled else 0)
        assert rc == 0
        self._n_types[key] = n_types

    def spawner_remove(self, key):
        self._L.fwo_spawner_remove(self._w, key)
        self._n_types.pop(key, None)

    def set_colliders(self, colliders):
        self._L.fwo_set_colliders(self._w, _colliders_array(colliders), len(colliders))

    def frame(self, dt: float, inputs=()):
        arr = inputs if isinstance(inputs, C.Array) else frame_inputs_array(inputs)
        n = len(inputs)
        self._L.fwo_frame(self._w, dt, arr, n, self.n_threads)

    def spawn_only(self, dt: float, inputs=()):
        self._L.fwo_spawn_only(self._w, dt, frame_inputs_array(inputs), len(inputs))

    def update_only(self, dt: float):
        self._L.fwo_update_only(self._w, dt, self.n_threads)

    def sync(self):
        pass

    def counts(self, key, n_types=None) -> List[int]:
        n_types = self._n_types[key] if n_types is None else n_types
        return [int(self._L.fwo_count(self._w, key, t)) for t in range(n_types)]

    def total_live(self) -> int:
        return int(self._L.fwo_total_live(self._w))

    def _read(self, fn, key, type_):
        n = C.c_uint64()
        fn(self._w, key, type_, None, 0, C.byref(n))
        out = np.zeros(n.value, dtype=_abi.particle_data_dtype())
        if n.value:
            rc = fn(self._w, key, type_, out.ctypes.data, n.value, C.byref(n))
            assert rc == 0
        return out

    def read_particles(self, key, type_=0) -> np.ndarray:
        return self._read(self._L.fwo_read_particles, key, type_)

    def read_destroyed(self, key, type_=0) -> np.ndarray:
        return self._read(self._L.fwo_read_destroyed, key, type_)

    def write_particles(self, key, type_, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=_abi.particle_data_dtype())
        rc = self._L.fwo_write_particles(self._w, key, type_, rows.ctypes.data, len(rows))
        assert rc == 0

    def read_instances(self, key, type_=0) -> np.ndarray:
        """``From<&ParticleData> for ParticleInstance`` (ref src/render.rs:105-115)."""
        p = self.read_particles(key, type_)
        out = np.zeros(len(p), dtype=_abi.particle_instance_dtype())
        for f in ("position", "scale", "rotation", "base_color", "emissive_color"):
            out[f] = p[f]
        return out

    def status(self, key) -> _abi.fw_spawner_status:
        st = _abi.fw_spawner_status()
        rc = self._L.fwo_status(self._w, key, C.byref(st))
        assert rc == 0
        return st

    def mark_finished_notified(self, key):
        self._L.fwo_mark_finished_notified(self._w, key)

    def read_aabb(self, key):
        mn, mx, e = (C.c_float * 3)(), (C.c_float * 3)(), C.c_uint32()
        rc = self._L.fwo_read_aabb(self._w, key, C.byref(mn), C.byref(mx), C.byref(e))
        assert rc == 0
        return (None if e.value else (tuple(mn), tuple(mx)))
