"""Host-side logic of libfirework_b200.so that runs without a device (`fw_host_*` exports):

* the emission pacing `fw_frame` evaluates on the host (reference src/core.rs:553-575), pinned on
  the reference's own unit test G1 (src/core.rs:806-834) and bit-for-bit against the oracle;
* the collision broad phase `fw_set_colliders` uploads (inflated world AABBs, stackless BVH,
  uniform grid): the candidate enumeration of `cast_ray` (fw_math.cuh) is restated here in numpy
  float32 and must return exactly the colliders whose box overlaps the ray segment's box -- the
  property that makes the kernel's result identical to the oracle's test-every-collider loop.
"""
import ctypes as C
import struct

import numpy as np
import pytest

from bevy_firework_b200 import _abi
from bevy_firework_b200.build import build_native
from bevy_firework_b200.workloads import capsule, collision_scene_colliders, cone, cuboid, cylinder, sphere

f32 = np.float32


@pytest.fixture(scope="module")
def lib():
    build_native()
    from bevy_firework_b200._native import load_library

    return load_library()


def _emission_count(lib, t, last, cycle, start, end, count):
    n, nl = C.c_uint64(), C.c_float()
    assert lib.fw_host_emission_count(t, last, cycle, start, end, count, C.byref(n), C.byref(nl)) == _abi.FW_OK
    return int(n.value), float(nl.value)


def test_g1_on_the_library_host_path(lib):
    """src/core.rs:806-834 against the function fw_frame calls"""
    timestep, age = f32(0.016), f32(0.0)
    last, total = float(np.finfo(np.float32).min), 0
    while age <= f32(3.0):
        n, last = _emission_count(lib, float(age), last, 3.0, 0.0, 1.0, 23.0)
        total += n
        age = f32(age + timestep)
    assert total in (22, 23) and total == 22


def test_emission_count_matches_oracle_bitwise(lib):
    from oracle import oracle as O

    rng = np.random.default_rng(5)
    for _ in range(20000):
        cycle = float(f32(rng.uniform(0.05, 5.0)))
        t = float(f32(rng.uniform(0.0, cycle)))
        last = float(f32(rng.choice([rng.uniform(-cycle, cycle), np.finfo(np.float32).min, 0.0])))
        start = float(f32(rng.uniform(0.0, 0.5)))
        end = float(f32(rng.uniform(start, 1.0)))
        count = float(f32(rng.choice([rng.uniform(0.5, 100000.0), 0.0, 1.0])))
        got, want = _emission_count(lib, t, last, cycle, start, end, count), O.compute_emission_count(t, last, cycle, start, end, count)
        assert got[0] == want[0]
        assert struct.pack("f", got[1]) == struct.pack("f", want[1]) or (np.isnan(got[1]) and np.isnan(want[1]))


# ------------------------------------------------------------------------------ broad phase
HEADER = struct.Struct("<8I3I3f3f3I")  # BroadPhaseHeader (fw_internal.h), 80 bytes


class BroadPhase:
    def __init__(self, lib, colliders):
        n = len(colliders)
        arr = (_abi.fw_collider * max(n, 1))(*colliders)
        size = C.c_uint64()
        rc = lib.fw_host_build_broadphase(arr, n, None, 0, C.byref(size))
        assert rc in (_abi.FW_OK, _abi.FW_ERR_BUFFER_TOO_SMALL)
        buf = (C.c_uint8 * max(size.value, 1))()
        assert lib.fw_host_build_broadphase(arr, n, buf, size.value, C.byref(size)) == _abi.FW_OK
        b = bytes(buf)[: size.value]
        (self.n_nodes, nodes_off, leaf_off, self.n_big, big_off, self.use_grid, cell_off, items_off,
         dx, dy, dz, lx, ly, lz, ix, iy, iz, *_pad) = HEADER.unpack_from(b, 0)
        assert HEADER.size == 80
        self.dim = (dx, dy, dz)
        self.lo = np.array([lx, ly, lz], dtype=f32)
        self.inv = np.array([ix, iy, iz], dtype=f32)
        self.leaf = np.frombuffer(b, dtype=f32, count=8 * n, offset=leaf_off).reshape(n, 2, 4)
        self.leaf_u = np.frombuffer(b, dtype=np.uint32, count=8 * n, offset=leaf_off).reshape(n, 2, 4)
        self.nodes = np.frombuffer(b, dtype=f32, count=8 * self.n_nodes, offset=nodes_off).reshape(self.n_nodes, 2, 4)
        self.nodes_u = np.frombuffer(b, dtype=np.uint32, count=8 * self.n_nodes, offset=nodes_off).reshape(self.n_nodes, 2, 4)
        self.big = np.frombuffer(b, dtype=np.uint32, count=self.n_big, offset=big_off)
        n_cells = dx * dy * dz if self.use_grid else 1
        self.cell_start = np.frombuffer(b, dtype=np.uint32, count=n_cells + 1, offset=cell_off)
        self.items = np.frombuffer(b, dtype=np.uint32, count=int(self.cell_start[-1]), offset=items_off)
        self.n = n

    # -- the two enumerations of cast_ray (fw_math.cuh), restated
    def overlaps(self, i, slo, shi):
        lo, hi = self.leaf[i, 0, :3], self.leaf[i, 1, :3]
        return not ((shi < lo).any() or (slo > hi).any())

    def brute(self, slo, shi, mask=0xFFFFFFFF):
        return sorted(i for i in range(self.n) if self.overlaps(i, slo, shi) and (int(self.leaf_u[i, 0, 3]) & mask))

    def walk_bvh(self, slo, shi, mask=0xFFFFFFFF):
        out, k, steps = [], 0, 0
        while k < self.n_nodes:
            lo, hi = self.nodes[k, 0, :3], self.nodes[k, 1, :3]
            link, leaf = int(self.nodes_u[k, 0, 3]), int(self.nodes_u[k, 1, 3])
            disjoint = bool((shi < lo).any() or (slo > hi).any())
            inner = leaf == 0xFFFFFFFF
            k = link if (inner and disjoint) else k + 1
            if not inner and not disjoint and (link & mask):
                out.append(leaf)
            steps += 1
        return sorted(out), steps

    def grid_coord(self, x, a):
        return np.floor((f32(x) - self.lo[a]) * self.inv[a])  # float32 arithmetic, like the kernel

    def enumerate_grid(self, slo, shi, mask=0xFFFFFFFF):
        """-> candidate list, or None when the kernel would fall back to the BVH"""
        if not self.use_grid:
            return None
        r0, r1 = [], []
        for a in range(3):
            d = f32(self.dim[a])
            lo = min(max(self.grid_coord(slo[a], a), f32(-1.0)), d)
            hi = min(max(self.grid_coord(shi[a], a), f32(-1.0)), d)
            r0.append(max(int(lo), 0))
            r1.append(min(int(hi), self.dim[a] - 1))
        if any(r1[a] - r0[a] > 1 for a in range(3)):
            return None
        out = [int(i) for i in self.big if self.overlaps(int(i), slo, shi) and (int(self.leaf_u[int(i), 0, 3]) & mask)]
        if all(r1[a] >= r0[a] for a in range(3)):
            span = [(r1[a] - r0[a]) for a in range(3)]
            for sz in range(span[2] + 1):
                for sy in range(span[1] + 1):
                    for sx in range(span[0] + 1):
                        cell = ((r0[2] + sz) * self.dim[1] + (r0[1] + sy)) * self.dim[0] + (r0[0] + sx)
                        for i in self.items[self.cell_start[cell]: self.cell_start[cell + 1]]:
                            i = int(i)
                            if not self.overlaps(i, slo, shi) or not (int(self.leaf_u[i, 0, 3]) & mask):
                                continue
                            first = [1 if (span[a] and self.grid_coord(self.leaf[i, 0, a], a) > f32(r0[a])) else 0 for a in range(3)]
                            if first == [sx, sy, sz]:
                                out.append(i)
        return sorted(out)


def _scene(rng, n, with_big=True, with_nan=False):
    cols = [cuboid((40, 1, 40), (0, -0.5, 0))] if with_big else []
    while len(cols) < n:
        pos = rng.uniform(-8, 8, 3)
        pos[1] = rng.uniform(0, 6)
        layers = 1 if len(cols) % 3 else 2
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        kind = len(cols) % 5
        if kind == 4:
            cols.append(capsule(float(rng.uniform(0.1, 0.6)), float(rng.uniform(0.2, 1.5)), pos, tuple(q), layers=layers))
        elif kind == 1:
            cols.append(cuboid(rng.uniform(0.1, 1.6, 3), pos, tuple(q), layers=layers))
        elif kind == 2:
            cols.append(cylinder(float(rng.uniform(0.1, 0.8)), float(rng.uniform(0.2, 1.5)), pos, tuple(q), layers=layers))
        elif kind == 3:
            cols.append(cone(float(rng.uniform(0.1, 0.8)), float(rng.uniform(0.2, 1.5)), pos, tuple(q), layers=layers))
        else:
            cols.append(sphere(float(rng.uniform(0.1, 0.9)), pos, layers=layers))
    if with_nan:
        cols.append(cuboid((1, 1, 1), (float("nan"), 0.0, 0.0)))
    return cols


def test_broadphase_boxes_contain_the_colliders(lib):
    rng = np.random.default_rng(1)
    cols = _scene(rng, 200)
    bp = BroadPhase(lib, cols)
    for i, c in enumerate(cols):
        t = np.array(c.translation[:], dtype=np.float64)
        x, y, z, w = [float(v) for v in c.rotation[:]]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        he = np.array(c.half_extents[:], dtype=np.float64)
        if c.kind in (_abi.FW_COLLIDER_CYLINDER, _abi.FW_COLLIDER_CONE):
            he = np.array([he[0], he[1], he[0]])  # solid of revolution about +Y: (r, h, r)
        if c.kind == _abi.FW_COLLIDER_CAPSULE:
            he = np.array([he[0], he[1] + he[0], he[0]])  # segment half length + the ball
        ext = np.full(3, abs(he[0])) if c.kind == _abi.FW_COLLIDER_SPHERE else np.abs(R) @ np.abs(he)
        assert (bp.leaf[i, 0, :3] < t - ext).all() and (bp.leaf[i, 1, :3] > t + ext).all()
        assert int(bp.leaf_u[i, 0, 3]) == c.layers
    # the C5 scene: the ground slab is "big", everything else sits in the grid
    bp5 = BroadPhase(lib, collision_scene_colliders(256))
    assert bp5.use_grid and list(bp5.big) == [0] and bp5.n_nodes == 2 * 256 - 1


@pytest.mark.parametrize("seed,n,with_nan", [(2, 300, False), (3, 64, True), (4, 9, False), (5, 1, False)])
def test_broadphase_enumerations_equal_brute_force(lib, seed, n, with_nan):
    rng = np.random.default_rng(seed)
    bp = BroadPhase(lib, _scene(rng, n, with_big=n > 4, with_nan=with_nan))
    via_grid = via_bvh = 0
    for k in range(1500):
        o = rng.uniform(-9, 9, 3).astype(f32)
        o[1] = f32(rng.uniform(-1, 7))
        length = rng.choice([0.02, 0.15, 0.6, 3.0, 40.0])
        d = rng.normal(size=3)
        e = (o + f32(length) * (d / np.linalg.norm(d)).astype(f32)).astype(f32)
        slo, shi = np.minimum(o, e), np.maximum(o, e)
        mask = int(rng.choice([0xFFFFFFFF, 1, 2]))
        want = bp.brute(slo, shi, mask)
        got_bvh, _ = bp.walk_bvh(slo, shi, mask)
        assert got_bvh == want, (k, "bvh")
        got_grid = bp.enumerate_grid(slo, shi, mask)
        if got_grid is None:
            via_bvh += 1
        else:
            via_grid += 1
            assert got_grid == want, (k, "grid", got_grid, want)  # nothing missing, nothing twice
    if bp.use_grid:
        assert via_grid > 500 and via_bvh > 100


def test_bvh_walk_is_short_for_short_segments(lib):
    """the point of the structure: a short segment visits a small part of the tree"""
    rng = np.random.default_rng(9)
    bp = BroadPhase(lib, collision_scene_colliders(256))
    steps = []
    for _ in range(300):
        o = rng.uniform(-8, 8, 3).astype(f32)
        o[1] = f32(rng.uniform(0.5, 5))
        e = (o + rng.uniform(-0.08, 0.08, 3).astype(f32)).astype(f32)
        steps.append(bp.walk_bvh(np.minimum(o, e), np.maximum(o, e))[1])
    assert np.mean(steps) < 80 and bp.n_nodes == 511


def test_broadphase_never_drops_the_hit_the_exact_test_reports(lib):
    """what the kernels rely on (fw_math.cuh, cast_ray): whatever collider the brute-force loop over every
    collider reports as the closest hit is among the candidates of the grid walk, or of the BVH walk when the
    segment is long. Rays of a particle step, long rays, rays along cone slants and capsule / cylinder axes.
    (The random campaign found the one case where this failed: a false hit of the analytic cone at distance 0.)"""
    from oracle import oracle as O

    rng = np.random.default_rng(21)
    cols = _scene(rng, 120)
    arr = (_abi.fw_collider * len(cols))(*cols)
    bp = BroadPhase(lib, cols)
    n = int(__import__("os").environ.get("FW_BROADPHASE_RAYS", "4000"))
    hits = 0
    for i in range(n):
        c = cols[int(rng.integers(1, len(cols)))]
        tr = np.array(c.translation[:], dtype=np.float64)
        o = (tr + rng.uniform(-3, 3, 3)) if i % 3 else rng.uniform(-9, 9, 3) * (1.0, 0.4, 1.0) + (0.0, 1.0, 0.0)
        if i % 4 == 0 and c.kind in (_abi.FW_COLLIDER_CYLINDER, _abi.FW_COLLIDER_CONE, _abi.FW_COLLIDER_CAPSULE):
            x, y, z, w = [float(v) for v in c.rotation[:]]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            slope = float(c.half_extents[0]) / (2.0 * float(c.half_extents[1])) if c.kind == _abi.FW_COLLIDER_CONE else 0.0
            phi = rng.uniform(0, 2 * np.pi)
            d = R @ (np.array([slope * np.cos(phi), -1.0, slope * np.sin(phi)]) * rng.choice([-1.0, 1.0]))
        else:
            d = rng.normal(size=3)
        d = (d / np.linalg.norm(d)).astype(f32)
        o = o.astype(f32)
        md = f32(rng.choice([0.03, 0.15, 0.6, 25.0]))
        mask = int(rng.choice([0xFFFFFFFF, 1, 2]))
        hit = O.cast_ray(arr, tuple(float(v) for v in o), tuple(float(v) for v in d), float(md), filter_mask=mask)
        if hit is None:
            continue
        hits += 1
        e = (o + d * md).astype(f32)
        slo, shi = np.minimum(o, e), np.maximum(o, e)
        cand = bp.enumerate_grid(slo, shi, mask)
        if cand is None:
            cand = bp.walk_bvh(slo, shi, mask)[0]
        assert hit[2] in cand, (i, cols[hit[2]].kind, o, d, md, hit, cand)
    assert hits > n // 20


def test_fuzz_cast_ray_harness(lib, tmp_path):
    """scripts/probes/fuzz_cast_ray.c (the 1e9-ray campaign of profiles/r2/x_fuzz_cast_ray_1e9_rays.txt) on a
    few million rays: compiled from the oracle's source, boxes from the library's own broad-phase builder"""
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "fuzz_cast_ray")
    subprocess.check_call([gcc, "-O2", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-w",
                           "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "oracle"),
                           os.path.join(root, "scripts", "probes", "fuzz_cast_ray.c"), "-o", exe, "-lm", "-ldl"])
    from bevy_firework_b200._native import LIB_PATH

    out = subprocess.run([exe, LIB_PATH, "8", "400000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "broad-phase box: 0 |" in out.stdout and "culled helper: 0 |" in out.stdout


def test_kernel_cast_ray_source_on_the_host_equals_the_oracle(lib, tmp_path):
    """scripts/probes/host_cast_ray.cu compiles the KERNELS' cast_ray (fw_math.cuh: grid / BVH enumeration,
    candidate queue, exact tests, masks, exclusions, tie-break) for the host as a one-lane warp and compares
    it with the oracle's brute-force loop bit for bit (6.6e8 rays once: profiles/r2/x_host_cast_ray_6e8_rays.txt;
    here ~1.6 M). No GPU involved: it checks the source the GPU runs, not the GPU."""
    import os
    import shutil
    import subprocess

    from bevy_firework_b200._native import LIB_PATH
    from bevy_firework_b200.build import nvcc_path

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        nvcc = nvcc_path()
    except RuntimeError:
        pytest.skip("no nvcc")
    from oracle import oracle as O

    O.lib()  # builds oracle/libfw_oracle.so if needed
    exe = str(tmp_path / "host_cast_ray")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-w", "-Xcompiler", "-ffp-contract=off,-fno-fast-math,-fopenmp",
                           "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "bevy_firework_b200", "csrc"),
                           os.path.join(root, "scripts", "probes", "host_cast_ray.cu"), "-o", exe, "-ldl", "-lgomp"])
    out = subprocess.run([exe, LIB_PATH, os.path.join(root, "oracle", "libfw_oracle.so"), "16", "100000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "!= oracle brute force: 0" in out.stdout
    # the rest of the header the same way: particle_collision, curves, gradients, Philox, quaternions
    # (5.1e8 collisions once: profiles/r2/x_host_math_5e8_collisions.txt)
    exe2 = str(tmp_path / "host_math")
    # the static update kernel's own gradient sampler (knot interval carried from call to call) lives in
    # fw_kernels.cu: cut its text out, from its signature to the next template
    src = open(os.path.join(root, "bevy_firework_b200", "csrc", "fw_kernels.cu")).read()
    i = src.index("__device__ __forceinline__ float4 sample_gradient_hint")
    j = src.index("template <bool COMPACT>", i)
    (tmp_path / "sgh.inc").write_text(src[i:j])
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-w", "-DFW_HAVE_SGH", "-Xcompiler", "-ffp-contract=off,-fno-fast-math,-fopenmp",
                           "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "bevy_firework_b200", "csrc"),
                           "-I" + str(tmp_path),
                           os.path.join(root, "scripts", "probes", "host_math.cu"), "-o", exe2, "-ldl", "-lgomp"])
    out = subprocess.run([exe2, LIB_PATH, os.path.join(root, "oracle", "libfw_oracle.so"), "16", "50000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("mismatches 0") == 4


def test_warp_searches_restated():
    """the two 32-ary warp searches of the kernels (fw_kernels.cu: find_cmd_warp, find_tile_warp), restated lane
    by lane: every size from 1 to 130, around the powers of 32, past 32^3; streams without tiles share a prefix
    value with their successor and the LAST of them owns the tile"""
    def find_cmd_warp(first, begin, end, g):
        lo, n = begin, end - begin
        while n > 1:
            step = (n + 31) >> 5
            ok = [lane == 0 or (lo + lane * step < lo + n and first[lo + lane * step] <= g) for lane in range(32)]
            assert all(ok[i] or not ok[i + 1] for i in range(31))  # monotone: the ballot is a prefix of the lanes
            k = max(i for i in range(32) if ok[i])
            hi = lo + n
            lo += k * step
            n = min(step, hi - lo)
        return lo

    def find_tile_warp(prefix, n_slots, tile):
        lo, n = 0, n_slots
        while n > 1:
            step = (n + 31) // 32
            le = [lane * step < n and prefix[lo + lane * step] <= tile for lane in range(32)]
            k = max([i for i in range(32) if le[i]], default=0)
            lo += k * step
            n = min(step, n - k * step)
        return lo

    rng = np.random.default_rng(0)
    for n in list(range(1, 131)) + [255, 256, 257, 1023, 1024, 1025, 2048, 4097, 32768, 32769, 40000]:
        counts = rng.integers(1, 5, n)                      # every command has count > 0
        first = np.concatenate([[0], np.cumsum(counts)[:-1]])
        begin = int(rng.integers(0, 3))
        arr = np.concatenate([np.zeros(begin, dtype=np.int64), first])
        total = int(counts.sum())
        for g in (range(total) if total < 1500 else rng.integers(0, total, 1500)):
            assert find_cmd_warp(arr, begin, begin + n, int(g)) == begin + int(np.searchsorted(first, g, side="right") - 1)
        tiles = rng.integers(0, 4, n) * (rng.random(n) < 0.5)  # half the streams have no tile
        if tiles.sum() == 0:
            tiles[int(rng.integers(0, n))] = 2
        prefix = np.concatenate([[0], np.cumsum(tiles)]).astype(np.int64)
        total = int(prefix[-1])
        for t in (range(total) if total < 1500 else rng.integers(0, total, 1500)):
            s = find_tile_warp(prefix, n, int(t))
            assert prefix[s] <= t < prefix[s + 1]
