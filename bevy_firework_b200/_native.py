"""ctypes binding of ``libfirework_b200.so`` (the C ABI of ``include/firework_b200.h``).

There is no fallback path: if the shared library is missing it is built with nvcc; if that is
impossible, or no sm_100 device is present, creating an ``Engine`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import _abi
from .build import LIB_PATH, build_native

_lib = None


class FireworkError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"firework_b200 error {code}: {message}")
        self.code = code
        self.message = message


def load_library(build_if_missing: bool = True):
    """dlopen the library and attach the prototypes of every symbol the header declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise FileNotFoundError(LIB_PATH)
        build_native()
    # FW_B200_LIB: a differently tuned build of the SAME library (kernel tuning experiments)
    L = C.CDLL(os.environ.get("FW_B200_LIB") or LIB_PATH)
    for name, (res, args) in _abi.EXPORTS.items():
        fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if L.fw_abi_version() != _abi.FW_ABI_VERSION:
        raise RuntimeError("libfirework_b200.so ABI version mismatch")
    _lib = L
    return L


def _frame_inputs_array(inputs):
    if isinstance(inputs, C.Array):
        return inputs, len(inputs)
    arr = (_abi.fw_spawner_frame_input * max(len(inputs), 1))()
    for i, x in enumerate(inputs):
        arr[i] = x
    return arr, len(inputs)


def frame_input(key: int, translation=(0.0, 0.0, 0.0), rotation=(0.0, 0.0, 0.0, 1.0),
                parent_velocity=(0.0, 0.0, 0.0), modifier_scale: float = 1.0,
                modifier_speed: float = 1.0, queue_particles: int = 0) -> _abi.fw_spawner_frame_input:
    x = _abi.fw_spawner_frame_input()
    x.spawner_key = key
    x.origin_translation[:] = [float(c) for c in translation]
    x.origin_rotation[:] = [float(c) for c in rotation]
    x.parent_velocity[:] = [float(c) for c in parent_velocity]
    x.modifier_scale = modifier_scale
    x.modifier_speed = modifier_speed
    x.queue_particles = int(queue_particles)
    return x


class Engine:
    """One ``fw_context``: all particle state of one GPU."""

    def __init__(self, device: int = 0, seed: int = 0x00F12E00, profile: bool = False,
                 external_stream: Optional[int] = None, graphs: bool = True, concurrent_spawn: bool = True):
        self._L = load_library()
        cfg = _abi.fw_config()
        cfg.abi_version = _abi.FW_ABI_VERSION
        cfg.device = device
        cfg.seed = seed
        cfg.external_stream = external_stream
        cfg.flags = ((_abi.FW_FLAG_PROFILE if profile else 0) | (0 if graphs else _abi.FW_FLAG_NO_GRAPHS)
                     | (0 if concurrent_spawn else _abi.FW_FLAG_NO_CONCURRENT_SPAWN))
        self._ctx = C.c_void_p()
        rc = self._L.fw_create(C.byref(cfg), C.byref(self._ctx))
        if rc != _abi.FW_OK:
            self._ctx = C.c_void_p()
            raise FireworkError(rc, self._L.fw_last_global_error().decode())
        self._n_types = {}
        self._n_streams = 0
        self._counts_buf = None
        self.device = device

    # -- plumbing
    def _check(self, rc: int):
        if rc != _abi.FW_OK:
            raise FireworkError(rc, self._L.fw_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.fw_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the ABI, one method per export
    def spawner_reset(self, key, ps, n_types, es, n_emitters, starts_enabled=True):
        self._check(self._L.fw_spawner_reset(self._ctx, key, ps, n_types, es, n_emitters,
                                             1 if starts_enabled else 0))
        self._n_streams += n_types - self._n_types.get(key, 0)
        self._n_types[key] = n_types

    def spawner_remove(self, key):
        self._check(self._L.fw_spawner_remove(self._ctx, key))
        self._n_streams -= self._n_types.pop(key, 0)

    def set_colliders(self, colliders: Sequence[_abi.fw_collider]):
        arr = (_abi.fw_collider * max(len(colliders), 1))()
        for i, c in enumerate(colliders):
            arr[i] = c
        self._check(self._L.fw_set_colliders(self._ctx, arr, len(colliders)))

    def frame(self, dt: float, inputs=()):
        arr, n = _frame_inputs_array(inputs)
        self._check(self._L.fw_frame(self._ctx, dt, arr, n))

    def sync(self):
        self._check(self._L.fw_sync(self._ctx))

    def poll_device_errors(self) -> int:
        """FW_DEVICE_* bits raised by frames that have completed (never waits)"""
        f = C.c_uint32()
        self._check(self._L.fw_poll_device_errors(self._ctx, C.byref(f)))
        return int(f.value)

    def counts(self, key, n_types=None) -> List[int]:
        n_types = self._n_types[key] if n_types is None else n_types
        out = (C.c_uint32 * max(n_types, 1))()
        self._check(self._L.fw_counts(self._ctx, key, out, n_types))
        return [int(out[i]) for i in range(n_types)]

    def counts_all(self):
        """(keys, types, counts) of every stream in creation order; the arrays are views of buffers
        this Engine reuses between calls (copy them to keep them)."""
        cap = self._n_streams
        buf = self._counts_buf
        if buf is None or buf[0] < cap:
            arrs = [np.zeros(max(cap, 1), dtype=np.uint32) for _ in range(3)]
            u32p = C.POINTER(C.c_uint32)
            buf = self._counts_buf = (max(cap, 1), arrs, [a.ctypes.data_as(u32p) for a in arrs], C.c_uint32())
        _, arrs, ptrs, n = buf
        self._check(self._L.fw_counts_all(self._ctx, ptrs[0], ptrs[1], ptrs[2], buf[0], C.byref(n)))
        k = n.value
        return arrs[0][:k], arrs[1][:k], arrs[2][:k]

    def total_live(self) -> int:
        n = C.c_uint64()
        self._check(self._L.fw_total_live(self._ctx, C.byref(n)))
        return int(n.value)

    def status(self, key) -> _abi.fw_spawner_status:
        st = _abi.fw_spawner_status()
        self._check(self._L.fw_spawner_status_get(self._ctx, key, C.byref(st)))
        return st

    def stream_layout(self, key, type_=0) -> _abi.fw_stream_layout:
        out = _abi.fw_stream_layout()
        self._check(self._L.fw_stream_layout_get(self._ctx, key, type_, C.byref(out)))
        return out

    def mark_finished_notified(self, key):
        self._check(self._L.fw_spawner_mark_finished_notified(self._ctx, key))

    def _read_rows(self, fn, dtype, key, type_):
        # first call sizes the buffer (FW_ERR_BUFFER_TOO_SMALL still reports the count)
        n = C.c_uint64()
        rc = fn(self._ctx, key, type_, None, 0, C.byref(n))
        if rc not in (_abi.FW_OK, _abi.FW_ERR_BUFFER_TOO_SMALL):
            self._check(rc)
        out = np.zeros(n.value, dtype=dtype)
        if n.value:
            self._check(fn(self._ctx, key, type_, out.ctypes.data, n.value, C.byref(n)))
        return out[: n.value]

    def read_particles(self, key, type_=0) -> np.ndarray:
        return self._read_rows(self._L.fw_read_particles, _abi.particle_data_dtype(), key, type_)

    def read_instances(self, key, type_=0) -> np.ndarray:
        return self._read_rows(self._L.fw_read_instances, _abi.particle_instance_dtype(), key, type_)

    def read_destroyed(self, key, type_=0) -> np.ndarray:
        return self._read_rows(self._L.fw_read_destroyed, _abi.particle_data_dtype(), key, type_)

    def write_particles(self, key, type_, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=_abi.particle_data_dtype())
        self._check(self._L.fw_write_particles(self._ctx, key, type_, rows.ctypes.data if len(rows) else None, len(rows)))

    def read_aabb(self, key):
        mn, mx, e = (C.c_float * 3)(), (C.c_float * 3)(), C.c_uint32()
        self._check(self._L.fw_read_aabb(self._ctx, key, C.byref(mn), C.byref(mx), C.byref(e)))
        return None if e.value else (tuple(mn), tuple(mx))

    def pack_instances_device(self, device_ptr: int, cap_rows: int) -> int:
        n = C.c_uint64()
        self._check(self._L.fw_pack_instances_device(self._ctx, device_ptr, cap_rows, C.byref(n)))
        return int(n.value)

    def extract_instances(self, host_ptr: int, cap_rows: int) -> int:
        n = C.c_uint64()
        self._check(self._L.fw_extract_instances(self._ctx, host_ptr, cap_rows, C.byref(n)))
        return int(n.value)

    def extract_begin(self, host_ptr: int, cap_rows: int, spawner_keys=None):
        """asynchronous render extract (rows of the listed spawners, or of all) into pinned host memory"""
        if spawner_keys is None:
            self._check(self._L.fw_extract_begin(self._ctx, None, 0, host_ptr, cap_rows))
        else:
            arr = (C.c_uint32 * max(len(spawner_keys), 1))(*spawner_keys)
            self._check(self._L.fw_extract_begin(self._ctx, arr, len(spawner_keys), host_ptr, cap_rows))

    def extract_wait(self, cap_streams: int = 0):
        """-> (rows landed, first row of every listed stream)"""
        n, ns = C.c_uint64(), C.c_uint32()
        firsts = (C.c_uint64 * max(cap_streams, 1))()
        self._check(self._L.fw_extract_wait(self._ctx, C.byref(n), firsts, cap_streams, C.byref(ns)))
        return int(n.value), [int(firsts[k]) for k in range(min(cap_streams, ns.value))]

    def export_instances_fd(self):
        """-> (fd, bytes, rows): the packed rows as a POSIX file descriptor of a CUDA VMM allocation"""
        fd, nbytes, rows = C.c_int32(-1), C.c_uint64(), C.c_uint64()
        self._check(self._L.fw_export_instances_fd(self._ctx, C.byref(fd), C.byref(nbytes), C.byref(rows)))
        return int(fd.value), int(nbytes.value), int(rows.value)

    # -- multi-GPU render extract over peer memory (fw_gather_*)
    def gather_create(self, n_ranks: int, my_rank: int, cap_rows_per_rank: int) -> bytes:
        """allocate this rank's gather buffer; returns its handle as bytes (send it to the peers)"""
        h = _abi.fw_gather_handle()
        self._check(self._L.fw_gather_create(self._ctx, n_ranks, my_rank, cap_rows_per_rank, C.byref(h)))
        return bytes(h)

    def gather_connect(self, handles: Sequence[bytes]):
        arr = (_abi.fw_gather_handle * len(handles))(*[_abi.fw_gather_handle.from_buffer_copy(b) for b in handles])
        self._check(self._L.fw_gather_connect(self._ctx, arr, len(handles)))

    def gather_instances(self):
        """collective, asynchronous: every rank must issue it before any rank waits for the result"""
        self._check(self._L.fw_gather_instances(self._ctx))

    def gather_result(self, n_ranks: int):
        """-> (device pointer of region 0, rows per rank, region stride in rows); synchronises"""
        ptr, stride = C.c_void_p(), C.c_uint64()
        rows = (C.c_uint64 * n_ranks)()
        self._check(self._L.fw_gather_result(self._ctx, C.byref(ptr), rows, n_ranks, C.byref(stride)))
        return int(ptr.value or 0), [int(r) for r in rows], int(stride.value)

    def gather_destroy(self):
        self._check(self._L.fw_gather_destroy(self._ctx))

    def device_sincos(self, x):
        """include/fw_sincos.h evaluated on the device -> (sin, cos) float32 arrays"""
        x = np.ascontiguousarray(x, dtype=np.float32)
        s, c = np.empty_like(x), np.empty_like(x)
        self._check(self._L.fw_device_sincos(self._ctx, x.ctypes.data, x.size, s.ctypes.data, c.ctypes.data))
        return s, c

    def event_record(self, slot: int):
        self._check(self._L.fw_event_record(self._ctx, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._check(self._L.fw_event_elapsed_ms(self._ctx, a, b, C.byref(ms)))
        return float(ms.value)

    def set_profiling(self, on: bool):
        self._check(self._L.fw_set_profiling(self._ctx, 1 if on else 0))

    def profile_last(self) -> _abi.fw_frame_profile:
        p = _abi.fw_frame_profile()
        self._check(self._L.fw_profile_last(self._ctx, C.byref(p)))
        return p

    def profile_sum(self):
        p, n = _abi.fw_frame_profile(), C.c_uint32()
        self._check(self._L.fw_profile_sum(self._ctx, C.byref(p), C.byref(n)))
        return p, int(n.value)

    def profile_reset(self):
        self._check(self._L.fw_profile_reset(self._ctx))

    @property
    def stream_handle(self) -> int:
        return int(self._L.fw_stream_handle(self._ctx) or 0)
