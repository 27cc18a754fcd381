//! gen_golden — emits `tests/golden/reference_vectors.json`: known answers of the third-party
//! arithmetic the reference's particle path calls, produced BY the pinned crates themselves.
//!
//! UNCOMPILED SOURCE: the build image has no Rust toolchain, so this program has
//! never run. It exists so that the "PARITY UNPINNED" rows of DESIGN.md section 4 can be closed
//! mechanically wherever cargo is available:
//!
//!     # inside a checkout of mbrea-c/bevy_firework @ 6eb47d9 (so Cargo.lock pins the versions)
//!     cp <this repo>/rust/gen_golden.rs examples/gen_golden.rs
//!     cargo run --release --example gen_golden --features physics_avian > reference_vectors.json
//!     cp reference_vectors.json <this repo>/tests/golden/
//!     cd <this repo> && python -m pytest tests/test_reference_vectors.py -q
//!
//! `tests/test_reference_vectors.py` skips while the file is absent and, once it exists, checks the
//! CPU oracle (and through the oracle the CUDA kernels, which are bit-equal to it) section by
//! section. Every section is a list of `{"in": ..., "out": ...}` records; floats are written as
//! their u32 bit patterns so that nothing is lost in decimal conversion.
//!
//! What each section pins (call sites in the reference, file:line relative to /root/reference):
//!   glam_*            glam 0.32.1 Quat / Vec3 (src/core.rs:440-448, 645-647, 776-784)
//!   pitch_yaw         bevy_utilitarian 0.10.0 PitchYaw::to_unit_vec (src/emission_shape.rs:28-30)
//!   rand_vec3         bevy_utilitarian RandVec3::generate — unseedable (thread-local rand 0.9):
//!                     20 000 samples per setting; the test recovers (polar angle, azimuth, magnitude)
//!                     and checks which distribution family they follow (src/core.rs:441,464-466)
//!   rand_f32          RandF32::generate samples (src/core.rs:443,451,455)
//!   curve_f32 / curve_rgba   bevy_math 0.19.0 EvenCore / UnevenCore through the crate's own
//!                     FireworkCurve / FireworkGradient::sample_clamped (src/curve.rs:8-75,79-164)
//!   cast_ray          parry3d 0.27.0 (through avian3d 0.7.0's Collider) cast_ray, solid = true, on
//!                     cuboid / sphere / cylinder / cone / capsule incl. origins inside, edge and corner hits
//!                     (src/core.rs:756-765)
//!   sin_cos           f32::sin_cos of the platform libm, against include/fw_sincos.h
//!   emission_count    the crate's own compute_emission_count on random inputs (src/core.rs:553-575;
//!                     needs `pub(crate)` -> `pub` or this file placed under src/ as a test)
use std::f32::consts::PI;

use avian3d::prelude::*;
use bevy::math::{Quat, Vec3};
use bevy::prelude::*;
use bevy_firework::curve::{FireworkCurve, FireworkGradient};
use bevy_utilitarian::prelude::*;
use serde_json::{json, Value};

/// splitmix64: inputs are generated here, so the file is reproducible
struct Sm(u64);
impl Sm {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E37_79B9_7F4A_7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
        z ^ (z >> 31)
    }
    fn unit(&mut self) -> f32 {
        (self.next() >> 40) as f32 / 16_777_216.0
    }
    fn range(&mut self, lo: f32, hi: f32) -> f32 {
        lo + (hi - lo) * self.unit()
    }
    fn vec3(&mut self, r: f32) -> Vec3 {
        Vec3::new(self.range(-r, r), self.range(-r, r), self.range(-r, r))
    }
    fn quat(&mut self) -> Quat {
        Quat::from_xyzw(self.range(-1., 1.), self.range(-1., 1.), self.range(-1., 1.), self.range(-1., 1.)).normalize()
    }
}

fn b(f: f32) -> Value {
    json!(f.to_bits())
}
fn bv(v: Vec3) -> Value {
    json!([v.x.to_bits(), v.y.to_bits(), v.z.to_bits()])
}
fn bq(q: Quat) -> Value {
    json!([q.x.to_bits(), q.y.to_bits(), q.z.to_bits(), q.w.to_bits()])
}

fn glam_sections(rng: &mut Sm, out: &mut serde_json::Map<String, Value>) {
    let mut scaled_axis = vec![];
    let mut arc = vec![];
    let mut mul_quat = vec![];
    let mut mul_vec3 = vec![];
    let mut norm_or_zero = vec![];
    let mut project = vec![];
    let mut reject = vec![];
    let mut rot_y = vec![];
    for i in 0..2000 {
        let v = if i % 50 == 0 { Vec3::ZERO } else { rng.vec3(if i % 7 == 0 { 1e-3 } else { 6.0 }) };
        scaled_axis.push(json!({"in": bv(v), "out": bq(Quat::from_scaled_axis(v))}));
        let (f, t) = (rng.vec3(1.0).normalize(), rng.vec3(1.0).normalize());
        let t = match i % 40 { 0 => f, 1 => -f, _ => t };
        arc.push(json!({"in": [bv(f), bv(t)], "out": bq(Quat::from_rotation_arc(f, t))}));
        let (p, q) = (rng.quat(), rng.quat());
        mul_quat.push(json!({"in": [bq(p), bq(q)], "out": bq(p * q)}));
        let w = rng.vec3(10.0);
        mul_vec3.push(json!({"in": [bq(p), bv(w)], "out": bv(p * w)}));
        let n = if i % 25 == 0 { Vec3::ZERO } else { rng.vec3(if i % 3 == 0 { 1e-20 } else { 5.0 }) };
        norm_or_zero.push(json!({"in": bv(n), "out": bv(n.normalize_or_zero())}));
        let (a, r) = (rng.vec3(8.0), rng.vec3(1.0).normalize());
        project.push(json!({"in": [bv(a), bv(r)], "out": bv(a.project_onto(r))}));
        reject.push(json!({"in": [bv(a), bv(r)], "out": bv(a.reject_from(r))}));
        let ang = rng.range(0.0, 2.0 * PI);
        rot_y.push(json!({"in": b(ang), "out": bq(Quat::from_rotation_y(ang))}));
    }
    out.insert("glam_from_scaled_axis".into(), json!(scaled_axis));
    out.insert("glam_from_rotation_arc".into(), json!(arc));
    out.insert("glam_mul_quat".into(), json!(mul_quat));
    out.insert("glam_mul_vec3".into(), json!(mul_vec3));
    out.insert("glam_normalize_or_zero".into(), json!(norm_or_zero));
    out.insert("glam_project_onto".into(), json!(project));
    out.insert("glam_reject_from".into(), json!(reject));
    out.insert("glam_from_rotation_y".into(), json!(rot_y));
}

fn utilitarian_sections(rng: &mut Sm, out: &mut serde_json::Map<String, Value>) {
    let mut py = vec![];
    for _ in 0..2000 {
        let (u, v) = (rng.range(0.0, 2.0 * PI), rng.range(0.0, PI));
        py.push(json!({"in": [b(u), b(v)], "out": bv(PitchYaw::new(u, v).to_unit_vec())}));
    }
    out.insert("pitch_yaw".into(), json!(py));
    // unseedable generators: raw samples per setting
    let settings = [
        (Vec3::Y, 30.0f32.to_radians(), 0.0f32, 10.0f32),
        (Vec3::new(1.0, 1.0, 0.0), 0.8, 6.0, 8.0),
        (Vec3::NEG_Z, 0.0, 2.0, 2.0),
        (Vec3::X, PI, 1.0, 1.0),
    ];
    let mut rv = vec![];
    for (dir, spread, lo, hi) in settings {
        let g = RandVec3 { magnitude: RandF32 { min: lo, max: hi }, direction: dir, spread };
        let samples: Vec<Value> = (0..20_000).map(|_| bv(g.generate())).collect();
        rv.push(json!({"in": {"direction": bv(dir), "spread": b(spread), "min": b(lo), "max": b(hi)}, "out": samples}));
    }
    out.insert("rand_vec3".into(), json!(rv));
    let mut rf = vec![];
    for (lo, hi) in [(0.02f32, 0.08f32), (0.5, 1.5), (3.0, 3.0)] {
        let g = RandF32 { min: lo, max: hi };
        let samples: Vec<Value> = (0..20_000).map(|_| b(g.generate())).collect();
        rf.push(json!({"in": [b(lo), b(hi)], "out": samples}));
    }
    out.insert("rand_f32".into(), json!(rf));
}

fn curve_sections(rng: &mut Sm, out: &mut serde_json::Map<String, Value>) {
    let ts: Vec<f32> = (0..400).map(|i| match i { 0 => -0.5, 1 => 0.0, 2 => 1.0, 3 => 1.5, 4 => 0.5, 5 => 0.7, _ => rng.range(-0.1, 1.1) }).collect();
    let mut f32_curves = vec![];
    let even = vec![1.0f32, 2.0, 0.5, 4.0];
    let uneven = vec![(0.0f32, 1.0f32), (0.1, 3.0), (0.7, 0.25), (1.0, 2.0)];
    let c_even = FireworkCurve::even_samples(even.clone());
    let c_uneven = FireworkCurve::uneven_samples(uneven.clone());
    let c_const = FireworkCurve::constant(1.7);
    for &t in &ts {
        f32_curves.push(json!({"in": {"kind": "even", "values": even.iter().map(|v| v.to_bits()).collect::<Vec<_>>(), "t": b(t)}, "out": b(c_even.sample_clamped(t))}));
        f32_curves.push(json!({"in": {"kind": "uneven", "knots": uneven.iter().map(|(a, v)| [a.to_bits(), v.to_bits()]).collect::<Vec<_>>(), "t": b(t)}, "out": b(c_uneven.sample_clamped(t))}));
        f32_curves.push(json!({"in": {"kind": "constant", "value": b(1.7), "t": b(t)}, "out": b(c_const.sample_clamped(t))}));
    }
    out.insert("curve_f32".into(), json!(f32_curves));
    // the stress_test.rs gradient (examples/stress_test.rs:100-106) and an even one
    let knots = vec![
        (0.0f32, LinearRgba::new(10.0, 7.0, 1.0, 1.0)),
        (0.7, LinearRgba::new(3.0, 1.0, 1.0, 1.0)),
        (0.8, LinearRgba::new(1.0, 0.3, 0.3, 1.0)),
        (0.9, LinearRgba::new(0.3, 0.3, 0.3, 1.0)),
        (1.0, LinearRgba::new(0.1, 0.1, 0.1, 0.0)),
    ];
    let g_uneven = FireworkGradient::uneven_samples(knots.clone());
    let evens = vec![LinearRgba::new(0.6, 0.3, 0.0, 0.0), LinearRgba::new(0.6, 0.3, 0.0, 0.35), LinearRgba::new(0.1, 0.2, 0.3, 1.0)];
    let g_even = FireworkGradient::even_samples(evens.clone());
    let rgba = |c: LinearRgba| json!([c.red.to_bits(), c.green.to_bits(), c.blue.to_bits(), c.alpha.to_bits()]);
    let mut rgba_curves = vec![];
    for &t in &ts {
        rgba_curves.push(json!({"in": {"kind": "uneven", "knots": knots.iter().map(|(a, c)| json!([a.to_bits(), rgba(*c)])).collect::<Vec<_>>(), "t": b(t)}, "out": rgba(g_uneven.sample_clamped(t))}));
        rgba_curves.push(json!({"in": {"kind": "even", "values": evens.iter().map(|c| rgba(*c)).collect::<Vec<_>>(), "t": b(t)}, "out": rgba(g_even.sample_clamped(t))}));
    }
    out.insert("curve_rgba".into(), json!(rgba_curves));
}

fn cast_ray_section(rng: &mut Sm, out: &mut serde_json::Map<String, Value>) {
    // shapes in their local frame and with a rigid transform, like avian hands them to parry
    let shapes: Vec<(&str, Collider, [f32; 3])> = vec![
        ("cuboid", Collider::cuboid(1.0, 2.0, 3.0), [0.5, 1.0, 1.5]),
        ("sphere", Collider::sphere(0.75), [0.75, 0.0, 0.0]),
        ("cylinder", Collider::cylinder(0.6, 2.0), [0.6, 1.0, 0.0]),
        ("cone", Collider::cone(0.5, 1.2), [0.5, 0.6, 0.0]),
        ("capsule", Collider::capsule(0.4, 1.5), [0.4, 0.75, 0.0]),
    ];
    let mut recs = vec![];
    for (name, col, he) in &shapes {
        for i in 0..1500 {
            let (tr, rot) = if i % 2 == 0 { (Vec3::ZERO, Quat::IDENTITY) } else { (rng.vec3(3.0), rng.quat()) };
            let origin = match i % 10 { 0 => tr + rng.vec3(0.2), _ => tr + rng.vec3(4.0) }; // some origins inside
            let mut dir = (tr + rng.vec3(0.8) - origin).normalize_or_zero();
            if i % 97 == 0 { dir = Vec3::new(1.0, 1.0, 0.0).normalize(); } // towards an edge when origin = (-1.5,-1.5,0) + tr
            let origin = if i % 97 == 0 { tr + rot * Vec3::new(-he[0] - 1.0, -he[1] - 1.0, 0.0) } else { origin };
            if dir == Vec3::ZERO { dir = Vec3::Y; }
            let max_distance = match i % 4 { 0 => 0.3, 1 => 2.0, _ => 50.0 };
            let hit = col.cast_ray(tr, rot, origin, dir, max_distance, true);
            let o = match hit { Some((d, n)) => json!({"distance": b(d), "normal": bv(n)}), None => Value::Null };
            recs.push(json!({"in": {"shape": name, "half_extents": [b(he[0]), b(he[1]), b(he[2])], "translation": bv(tr), "rotation": bq(rot),
                                     "origin": bv(origin), "direction": bv(dir), "max_distance": b(max_distance)}, "out": o}));
        }
    }
    out.insert("cast_ray".into(), json!(recs));
}

fn sin_cos_section(rng: &mut Sm, out: &mut serde_json::Map<String, Value>) {
    let mut recs = vec![];
    for i in 0..20_000 {
        let x = match i % 4 { 0 => rng.range(-PI, PI), 1 => rng.range(-100.0, 100.0), 2 => rng.range(-1e6, 1e6), _ => f32::from_bits((rng.next() >> 32) as u32) };
        if !x.is_finite() { continue; }
        let (s, c) = x.sin_cos();
        recs.push(json!({"in": b(x), "out": [b(s), b(c)]}));
    }
    out.insert("sin_cos".into(), json!(recs));
}

fn main() {
    let mut rng = Sm(0x00F1_2E00);
    let mut out = serde_json::Map::new();
    out.insert("meta".into(), json!({
        "generator": "rust/gen_golden.rs", "reference": "mbrea-c/bevy_firework @ 6eb47d9",
        "crates": {"glam": "0.32.1", "bevy_utilitarian": "0.10.0", "bevy_math": "0.19.0", "bevy_color": "0.19.0", "avian3d": "0.7.0", "parry3d": "0.27.0", "rand": "0.9.4"},
        "float_encoding": "u32 bit patterns", "target": std::env::consts::ARCH,
    }));
    glam_sections(&mut rng, &mut out);
    utilitarian_sections(&mut rng, &mut out);
    curve_sections(&mut rng, &mut out);
    cast_ray_section(&mut rng, &mut out);
    sin_cos_section(&mut rng, &mut out);
    println!("{}", serde_json::to_string(&Value::Object(out)).unwrap());
}
