"""The render hand-off (reference consumer src/render.rs:439-461, upload :568-584):
asynchronous extract of a visibility-culled subset into pinned host memory while the simulation keeps
running, and the packed rows as a shareable POSIX file descriptor imported by ANOTHER PROCESS.
Plus SpatialQueryFilter::excluded_entities on the collision sweep (src/core.rs:247,764)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from bevy_firework_b200 import (EmissionPacing, EmissionSettings, ParticleCollisionSettings, ParticleSettings, ParticleSpawner,
                                RandF32, RandVec3, _abi)
from bevy_firework_b200._native import FireworkError, frame_input
from bevy_firework_b200.workloads import cuboid, grid_positions, stress_spawner
from _parity import assert_rows_match, reset_both

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = float(np.float32(1.0) / np.float32(60.0))


def _scene(engine, n=12, rate=3000.0):
    sp = stress_spawner(rate=rate)
    inputs = []
    for i, p in enumerate(grid_positions(n)):
        ps, nt, es, ne = sp.pods()
        engine.spawner_reset(50 + i, ps, nt, es, ne, True)
        inputs.append(frame_input(50 + i, p))
    for _ in range(40):
        engine.frame(DT, inputs)
    return inputs


def test_async_extract_matches_read_instances_while_frames_run(engine):
    import torch

    inputs = _scene(engine)
    keys = [50 + i for i in range(12)]
    cap = engine.total_live() + 12 * 256
    host = [torch.zeros((cap, 16), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    want = {k: engine.read_instances(k, 0).copy() for k in keys}   # the state the extract sees
    engine.extract_begin(host[0].data_ptr(), cap)                  # all spawners
    for _ in range(3):                                             # the simulation goes on underneath
        engine.frame(DT, inputs)
    subset = [keys[7], keys[2], keys[9]]                           # a culled subset, caller's order
    want2 = {k: engine.read_instances(k, 0).copy() for k in subset}
    engine.extract_begin(host[1].data_ptr(), cap, subset)          # second extract outstanding at the same time
    with pytest.raises(FireworkError):                             # a third one is refused, not queued
        engine.extract_begin(host[0].data_ptr(), cap)
    engine.frame(DT, inputs)
    n, firsts = engine.extract_wait(len(keys))
    rows = host[0].numpy().view(_abi.particle_instance_dtype()).reshape(-1)
    assert n == sum(len(want[k]) for k in keys) and len(firsts) == len(keys)
    for j, k in enumerate(keys):
        got = rows[firsts[j]: firsts[j] + len(want[k])]
        assert got.tobytes() == want[k].tobytes(), k
    n2, firsts2 = engine.extract_wait(len(subset))
    rows2 = host[1].numpy().view(_abi.particle_instance_dtype()).reshape(-1)
    assert n2 == sum(len(want2[k]) for k in subset)
    for j, k in enumerate(subset):
        assert rows2[firsts2[j]: firsts2[j] + len(want2[k])].tobytes() == want2[k].tobytes(), k
    with pytest.raises(FireworkError):
        engine.extract_wait()                                      # nothing outstanding
    with pytest.raises(FireworkError) as e:
        engine.extract_begin(host[0].data_ptr(), 10)               # too small: refused up front
    assert e.value.code == _abi.FW_ERR_BUFFER_TOO_SMALL
    with pytest.raises(FireworkError) as e:
        engine.extract_begin(host[0].data_ptr(), cap, [999999])
    assert e.value.code == _abi.FW_ERR_UNKNOWN_SPAWNER


_IMPORTER = r"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, {root!r})
from bevy_firework_b200._native import load_library
L = load_library()
fd, nbytes, rows = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
out = np.zeros((rows, 16), dtype=np.float32)
rc = L.fw_import_instances_fd(0, fd, nbytes, rows, out.ctypes.data)
assert rc == 0, (rc, L.fw_last_global_error())
sys.stdout.buffer.write(out.tobytes())
"""


def test_exported_fd_imported_by_another_process(engine):
    """cuMemExportToShareableHandle -> a child process with no fw_context maps the descriptor and reads
    the same rows (what a Vulkan / wgpu external-memory import would see)"""
    import torch

    _scene(engine, n=5, rate=2000.0)
    fd, nbytes, rows = engine.export_instances_fd()
    assert fd >= 0 and rows == engine.total_live() and nbytes >= rows * 64
    buf = torch.empty((rows + 16, 16), dtype=torch.float32, device="cuda:0")
    assert engine.pack_instances_device(buf.data_ptr(), buf.shape[0]) == rows
    want = buf[:rows].cpu().numpy().tobytes()
    try:
        out = subprocess.run([sys.executable, "-c", _IMPORTER.format(root=ROOT), str(fd), str(nbytes), str(rows)],
                             pass_fds=[fd], capture_output=True, timeout=180)
    finally:
        os.close(fd)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    assert out.stdout == want
    fd2, _, rows2 = engine.export_instances_fd()                   # a second export replaces the allocation
    os.close(fd2)
    assert rows2 == rows


def test_excluded_colliders_are_not_seen(engine, oracle):
    """two spawners over the same two cuboids; the second one's filter excludes the upper cuboid's key:
    its particles fall through it onto the ground, the first one's bounce on it -- both equal to the oracle"""
    def spawner(excluded):
        return ParticleSpawner(
            particle_settings=[ParticleSettings(lifetime=RandF32.constant(1.5), linear_drag=0.1,
                                                collision_settings=ParticleCollisionSettings(restitution=0.4, friction=0.2, excluded=excluded))],
            emission_settings=[EmissionSettings(emission_pacing=EmissionPacing.rate(1200.0),
                                                initial_velocity=RandVec3(RandF32(0.5, 1.5), (0.0, -1.0, 0.0), 0.3))])
    cols = [cuboid((20.0, 1.0, 20.0), (0.0, -0.5, 0.0), key=7), cuboid((4.0, 0.2, 4.0), (0.0, 1.0, 0.0), key=42)]
    w = oracle.OracleWorld()
    engine.set_colliders(cols)
    w.set_colliders(cols)
    reset_both(engine, w, 1, spawner(()))
    reset_both(engine, w, 2, spawner((42, 1234)))
    inp = [frame_input(1, (0.0, 2.0, 0.0)), frame_input(2, (0.0, 2.0, 0.0))]
    for k in range(80):
        engine.frame(DT, inp)
        w.frame(DT, inp)
    a, b = engine.read_particles(1, 0), engine.read_particles(2, 0)
    assert_rows_match(a, w.read_particles(1, 0), what="sees both")
    assert_rows_match(b, w.read_particles(2, 0), what="excludes the shelf")
    old_a, old_b = a[a["age"] > 0.9], b[b["age"] > 0.9]
    assert (old_a["position"][:, 1] > 1.05).all()       # resting on the shelf (top at y = 1.1)
    assert (old_b["position"][:, 1] < 0.5).all()        # fell through it onto the ground
    with pytest.raises(ValueError):
        ParticleCollisionSettings(restitution=0.1, friction=0.1, excluded=tuple(range(9))).__class__  # (validated in to_pod)
        ParticleSettings(collision_settings=ParticleCollisionSettings(restitution=0.1, friction=0.1, excluded=tuple(range(9)))).to_pod()
