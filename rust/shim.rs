//! shim.rs — the patch a bevy_firework maintainer adds to run the simulation on a B200.
//!
//! UNCOMPILED SOURCE (no Rust toolchain in the build image). It lives INSIDE the bevy_firework
//! crate (`src/gpu.rs`, feature `gpu_b200`), keeps every public type of `src/core.rs` unchanged
//! and replaces the bodies of three systems of the chain in `src/plugin.rs:46-60`:
//!
//!   sync_spawner_data  (src/core.rs:343-365)  -> gpu_sync_spawner_data   (fw_spawner_reset)
//!   spawn_particles    (src/core.rs:367-551)  \
//!   update_particles   (src/core.rs:577-670)  /-> gpu_frame              (ONE fw_frame call)
//!   notify_finished_particle_spawners (:674-688) -> gpu_notify_finished  (fw_spawner_status_get)
//!
//! `propagate_particle_spawner_modifier` (:690-703) and `sync_parent_velocity` (:705-742) stay as
//! they are: their results (EffectModifier, ParticleSpawnerData.parent_velocity) are inputs of
//! `fw_frame`. The render bridge (`src/render.rs`) asks `GpuParticles::instances(entity, type)`
//! instead of converting `ParticleData` rows (`:403`).
use crate::core::*;
use crate::curve::{FireworkCurve, FireworkGradient};
use crate::emission_shape::EmissionShape;
use crate::gpu_sys::*; // rust/firework_b200_sys.rs
use bevy::prelude::*;
use bevy::reflect::GetPath;
use std::ffi::{CStr, CString};

/// Owner of the `fw_context`. Held in a `ResMut`, which gives the "one call at a time per
/// context" guarantee the ABI asks for; the library re-selects the device on every entry, so
/// Bevy may run the systems on any worker thread.
#[derive(Resource)]
pub struct GpuParticles {
    ctx: *mut fw_context,
}
unsafe impl Send for GpuParticles {}
unsafe impl Sync for GpuParticles {}

impl GpuParticles {
    pub fn new(device: i32, seed: u64) -> Result<Self, String> {
        check_layouts()?;
        let cfg = fw_config { abi_version: FW_ABI_VERSION, device, seed, external_stream: std::ptr::null_mut(), flags: 0, reserved: 0 };
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { fw_create(&cfg, &mut ctx) };
        if rc != FW_OK {
            return Err(unsafe { CStr::from_ptr(fw_last_global_error()) }.to_string_lossy().into_owned());
        }
        Ok(Self { ctx })
    }
    fn check(&self, rc: i32, what: &str) -> bool {
        if rc != FW_OK {
            // error convention of SURVEY section 8b: log, leave state untouched
            let msg = unsafe { CStr::from_ptr(fw_last_error(self.ctx)) }.to_string_lossy();
            error!("firework_b200: {what}: {msg}");
        }
        rc == FW_OK
    }
    /// `data.particles[ty].len()` without copying rows (examples/stress_test.rs:197-199)
    pub fn counts(&self, entity: Entity, n_types: usize) -> Vec<u32> {
        let mut out = vec![0u32; n_types];
        self.check(unsafe { fw_counts(self.ctx, key(entity), out.as_mut_ptr(), n_types as u32) }, "fw_counts");
        out
    }
    /// lazily refreshed host mirror of `data.particles[ty]`
    pub fn particles(&self, entity: Entity, ty: usize) -> Vec<ParticleData> {
        let n = self.counts(entity, ty + 1)[ty] as usize;
        let mut rows: Vec<fw_particle_data> = Vec::with_capacity(n);
        let mut got = 0u64;
        if self.check(unsafe { fw_read_particles(self.ctx, key(entity), ty as u32, rows.as_mut_ptr(), n as u64, &mut got) }, "fw_read_particles") {
            unsafe { rows.set_len(got as usize) };
        }
        rows.iter().map(particle_from_pod).collect()
    }
    /// vertex-instance rows for `prepare_instance_buffers` (src/render.rs:568-584)
    pub fn instances(&self, entity: Entity, ty: usize, out: &mut Vec<fw_particle_instance>) {
        let n = self.counts(entity, ty + 1)[ty] as usize;
        out.clear();
        out.reserve(n);
        let mut got = 0u64;
        if self.check(unsafe { fw_read_instances(self.ctx, key(entity), ty as u32, out.as_mut_ptr(), n as u64, &mut got) }, "fw_read_instances") {
            unsafe { out.set_len(got as usize) };
        }
    }
}
impl Drop for GpuParticles {
    fn drop(&mut self) {
        unsafe { fw_destroy(self.ctx) };
    }
}

fn key(e: Entity) -> u32 {
    e.index() // unique among live entities; the generation is irrelevant to the library
}

/// sizes of the `#[repr(C)]` mirrors against the compiled library
fn check_layouts() -> Result<(), String> {
    macro_rules! chk {
        ($t:ty) => {{
            let name = CString::new(stringify!($t)).unwrap();
            let c = unsafe { fw_abi_sizeof(name.as_ptr()) } as usize;
            if c != std::mem::size_of::<$t>() {
                return Err(format!("layout drift: {} is {} bytes in C, {} in Rust", stringify!($t), c, std::mem::size_of::<$t>()));
            }
        }};
    }
    chk!(fw_rand_f32);
    chk!(fw_rand_vec3);
    chk!(fw_curve_f32);
    chk!(fw_gradient);
    chk!(fw_collision_settings);
    chk!(fw_particle_settings);
    chk!(fw_emission_settings);
    chk!(fw_spawner_frame_input);
    chk!(fw_particle_data);
    chk!(fw_particle_instance);
    chk!(fw_collider);
    chk!(fw_config);
    chk!(fw_spawner_status);
    chk!(fw_frame_profile);
    chk!(fw_stream_layout);
    chk!(fw_gather_handle);
    // field offsets of the structs whose fields are not all the same size (std::mem::offset_of!)
    macro_rules! off {
        ($t:ty, $f:ident) => {{
            let (sn, fnm) = (CString::new(stringify!($t)).unwrap(), CString::new(stringify!($f)).unwrap());
            let c = unsafe { fw_abi_offsetof(sn.as_ptr(), fnm.as_ptr()) } as usize;
            if c != std::mem::offset_of!($t, $f) {
                return Err(format!("layout drift: {}.{} at {} in C, {} in Rust", stringify!($t), stringify!($f), c, std::mem::offset_of!($t, $f)));
            }
        }};
    }
    off!(fw_emission_settings, one_shot_count);
    off!(fw_emission_settings, initial_angular_velocity);
    off!(fw_particle_settings, collision);
    off!(fw_particle_settings, capacity_hint);
    off!(fw_config, seed);
    off!(fw_config, external_stream);
    off!(fw_spawner_status, live_particles);
    off!(fw_frame_profile, particles_updated);
    off!(fw_collider, half_extents);
    off!(fw_gather_handle, address);
    Ok(())
}

// ------------------------------------------------------------------ settings -> POD
fn rand_f32(r: &RandF32) -> fw_rand_f32 {
    fw_rand_f32 { min: r.min, max: r.max }
}
fn rand_vec3(r: &RandVec3) -> fw_rand_vec3 {
    fw_rand_vec3 { magnitude: rand_f32(&r.magnitude), direction: r.direction.to_array(), spread: r.spread }
}
fn curve(c: &FireworkCurve<f32>) -> fw_curve_f32 {
    let mut p = fw_curve_f32 { kind: FW_CURVE_CONSTANT, n: 1, times: [0.; FW_MAX_KNOTS], values: [0.; FW_MAX_KNOTS] };
    match c {
        FireworkCurve::Constant(k) => p.values[0] = k.sample_unchecked(0.),
        // bevy_math keeps the cores crate-private; they are reachable through Reflect
        FireworkCurve::SampleAuto(k) => {
            let s = k.path::<Vec<f32>>(".core.samples").expect("EvenCore.samples");
            assert!(s.len() <= FW_MAX_KNOTS, "at most {FW_MAX_KNOTS} curve samples are supported");
            p.kind = FW_CURVE_EVEN;
            p.n = s.len() as u32;
            p.values[..s.len()].copy_from_slice(s);
        }
        FireworkCurve::UnevenSampleAuto(k) => {
            let t = k.path::<Vec<f32>>(".core.times").expect("UnevenCore.times");
            let s = k.path::<Vec<f32>>(".core.samples").expect("UnevenCore.samples");
            assert!(s.len() <= FW_MAX_KNOTS, "at most {FW_MAX_KNOTS} curve samples are supported");
            p.kind = FW_CURVE_UNEVEN;
            p.n = s.len() as u32;
            p.times[..t.len()].copy_from_slice(t);
            p.values[..s.len()].copy_from_slice(s);
        }
    }
    p
}
fn gradient(g: &FireworkGradient<LinearRgba>) -> fw_gradient {
    let mut p = fw_gradient { kind: FW_CURVE_CONSTANT, n: 1, times: [0.; FW_MAX_KNOTS], colors: [[0.; 4]; FW_MAX_KNOTS] };
    match g {
        FireworkGradient::Constant(k) => p.colors[0] = k.sample_unchecked(0.).to_f32_array(),
        // our own curve types (src/curve.rs:79-164): `core` is visible inside the crate
        FireworkGradient::ColorSampleAuto(k) => {
            p.kind = FW_CURVE_EVEN;
            p.n = k.core.samples.len() as u32;
            for (i, c) in k.core.samples.iter().enumerate() {
                p.colors[i] = c.to_f32_array();
            }
        }
        FireworkGradient::ColorSampleUnevenAuto(k) => {
            p.kind = FW_CURVE_UNEVEN;
            p.n = k.core.samples.len() as u32;
            for (i, (t, c)) in k.core.times.iter().zip(k.core.samples.iter()).enumerate() {
                p.times[i] = *t;
                p.colors[i] = c.to_f32_array();
            }
        }
    }
    p
}
fn particle_settings(s: &ParticleSettings) -> fw_particle_settings {
    fw_particle_settings {
        lifetime: rand_f32(&s.lifetime),
        scale_curve: curve(&s.scale_curve),
        initial_scale: rand_f32(&s.initial_scale),
        acceleration: s.acceleration.to_array(),
        angular_acceleration: s.angular_acceleration.to_array(),
        linear_drag: s.linear_drag,
        angular_drag: s.angular_drag,
        base_color: gradient(&s.base_color),
        emissive_color: gradient(&s.emissive_color),
        pbr: s.pbr as u32,
        #[cfg(feature = "physics_avian")]
        collision: s.collision_settings.as_ref().map_or(fw_collision_settings::default(), |c| fw_collision_settings {
            enabled: 1,
            restitution: c.restitution,
            friction: c.friction,
            destroy_on_collision: c.destroy_on_collision as u32,
            filter_mask: c.filter.mask.0, // LayerMask bits of the SpatialQueryFilter
            // SpatialQueryFilter::excluded_entities (src/core.rs:247,764): the colliders carry their
            // entity key (gpu_sync_colliders), the sweep skips the listed ones
            n_excluded: c.filter.excluded_entities.len().min(FW_MAX_EXCLUDED) as u32,
            excluded_keys: {
                let mut k = [FW_NO_KEY; FW_MAX_EXCLUDED];
                if c.filter.excluded_entities.len() > FW_MAX_EXCLUDED {
                    bevy::log::warn_once!("firework_b200: more than {} excluded entities in a SpatialQueryFilter, the rest is ignored", FW_MAX_EXCLUDED);
                }
                for (slot, e) in k.iter_mut().zip(c.filter.excluded_entities.iter()) {
                    *slot = key(*e);
                }
                k
            },
        }),
        #[cfg(not(feature = "physics_avian"))]
        collision: fw_collision_settings::default(),
        capture_destroyed: s.event_handlers.particles_destroyed.is_some() as u32,
        capacity_hint: 0,
    }
}
fn emission_settings(e: &EmissionSettings) -> fw_emission_settings {
    let (pacing_kind, one_shot_count, count, duration, offset_start, offset_end) = match e.emission_pacing {
        EmissionPacing::OneShot(n) => (FW_PACING_ONE_SHOT, n as u64, 0., 1., 0., 1.),
        EmissionPacing::OnDemand => (FW_PACING_ON_DEMAND, 0, 0., 1., 0., 1.),
        EmissionPacing::CountOverDuration { count, duration, offset_start, offset_end } => {
            (FW_PACING_COUNT_OVER_DURATION, 0, count, duration, offset_start, offset_end)
        }
    };
    let (mode, target_particle_type) = match e.emission_mode {
        EmissionMode::Global => (FW_MODE_GLOBAL, 0),
        EmissionMode::Nested { target_particle_type } => (FW_MODE_NESTED, target_particle_type as u32),
    };
    let (shape_kind, shape_radius, shape_normal) = match e.emission_shape {
        EmissionShape::Point => (FW_SHAPE_POINT, 0., [0., 1., 0.]),
        EmissionShape::Sphere(r) => (FW_SHAPE_SPHERE, r, [0., 1., 0.]),
        EmissionShape::Circle { normal, radius } => (FW_SHAPE_CIRCLE, radius, normal.to_array()),
    };
    fw_emission_settings {
        particle_index: e.particle_index as u32,
        pacing_kind, one_shot_count, count, duration, offset_start, offset_end,
        mode, target_particle_type, shape_kind, shape_radius, shape_normal,
        initial_velocity: rand_vec3(&e.initial_velocity),
        initial_velocity_radial: rand_f32(&e.initial_velocity_radial),
        inherit_parent_velocity: e.inherit_parent_velocity as u32,
        initial_rotation: e.initial_rotation.to_array(),
        initial_angular_velocity: rand_vec3(&e.initial_angular_velocity),
    }
}
fn particle_from_pod(p: &fw_particle_data) -> ParticleData {
    ParticleData {
        position: Vec3::from_array(p.position),
        velocity: Vec3::from_array(p.velocity),
        rotation: Quat::from_array(p.rotation),
        angular_velocity: Vec3::from_array(p.angular_velocity),
        initial_scale: p.initial_scale,
        scale: p.scale,
        age: p.age,
        lifetime: p.lifetime,
        base_color: LinearRgba::from_f32_array(p.base_color),
        emissive_color: LinearRgba::from_f32_array(p.emissive_color),
        pbr: p.pbr != 0,
        last_emitted_age: Vec::new(), // device-side state of nested emission
    }
}

// ------------------------------------------------------------------ the replaced systems
/// replaces `sync_spawner_data` (src/core.rs:343-365)
pub fn gpu_sync_spawner_data(
    gpu: ResMut<GpuParticles>,
    mut spawners: Query<(Entity, &ParticleSpawner, &mut ParticleSpawnerData), Changed<ParticleSpawner>>,
) {
    for (entity, settings, mut data) in &mut spawners {
        let ps: Vec<_> = settings.particle_settings.iter().map(particle_settings).collect();
        let es: Vec<_> = settings.emission_settings.iter().map(emission_settings).collect();
        let rc = unsafe {
            fw_spawner_reset(gpu.ctx, key(entity), ps.as_ptr(), ps.len() as u32, es.as_ptr(), es.len() as u32, settings.starts_enabled as u32)
        };
        if gpu.check(rc, "fw_spawner_reset") {
            data.particles = vec![Vec::new(); settings.particle_settings.len()]; // host mirror starts empty
            data.initialized = true;
        }
    }
}

/// replaces `spawn_particles` + `update_particles` (src/core.rs:367-670): one batched call
pub fn gpu_frame(
    gpu: ResMut<GpuParticles>,
    mut q: Query<(Entity, &Transform, &GlobalTransform, &ParticleSpawner, &mut ParticleSpawnerData, Option<&EffectModifier>)>,
    time: Res<Time>,
    mut commands: Commands,
    mut inputs: Local<Vec<fw_spawner_frame_input>>,
    mut destroyed: Local<Vec<fw_particle_data>>,
) {
    inputs.clear();
    for (entity, transform, global_transform, settings, mut data, modifier) in &mut q {
        let origin = match settings.spawn_transform_mode {
            SpawnTransformMode::Global => global_transform.compute_transform(), // src/core.rs:432-435
            SpawnTransformMode::Local => *transform,
        };
        let modifier = modifier.cloned().unwrap_or_default();
        inputs.push(fw_spawner_frame_input {
            spawner_key: key(entity),
            origin_translation: origin.translation.to_array(),
            origin_rotation: origin.rotation.to_array(),
            parent_velocity: data.parent_velocity.to_array(),
            modifier_scale: modifier.scale,
            modifier_speed: modifier.speed,
            queue_particles: std::mem::take(&mut data.manual_queued_count) as u32,
        });
    }
    let rc = unsafe { fw_frame(gpu.ctx, time.delta_secs(), inputs.as_ptr(), inputs.len() as u32) };
    if !gpu.check(rc, "fw_frame") {
        return;
    }
    // what frames that have completed meanwhile flagged on the device (never waits)
    let mut flags = 0u32;
    if unsafe { fw_poll_device_errors(gpu.ctx, &mut flags) } == FW_OK && flags != 0 {
        error!("firework_b200: device flags {flags:#x} (1 = a ring overflowed and spawns were dropped, 2 = look-back table too small, 4 = a nested emitter exceeded its planned per-parent bound)");
    }
    // particles_destroyed handlers (src/core.rs:660-667): the reference runs the handler system with
    // the Vec of particles the update just destroyed. Types with a handler were reset with
    // capture_destroyed = 1, so the library kept those rows (age already bumped, old colours).
    for (entity, _, _, settings, _, _) in &q {
        for (ty, ps) in settings.particle_settings.iter().enumerate() {
            let Some(handler) = ps.event_handlers.particles_destroyed else { continue };
            let mut n: u64 = 0;
            // first call sizes the buffer (FW_ERR_BUFFER_TOO_SMALL still reports the count)
            unsafe { fw_read_destroyed(gpu.ctx, key(entity), ty as u32, std::ptr::null_mut(), 0, &mut n) };
            if n == 0 {
                continue;
            }
            destroyed.clear();
            destroyed.resize(n as usize, unsafe { std::mem::zeroed() });
            let rc = unsafe { fw_read_destroyed(gpu.ctx, key(entity), ty as u32, destroyed.as_mut_ptr(), n, &mut n) };
            if gpu.check(rc, "fw_read_destroyed") {
                let rows: Vec<ParticleData> = destroyed[..n as usize].iter().map(particle_from_pod).collect();
                commands.run_system_with(handler, rows); // :665
            }
        }
    }
}

/// replaces `notify_finished_particle_spawners` (src/core.rs:674-688)
pub fn gpu_notify_finished(gpu: ResMut<GpuParticles>, mut commands: Commands, q: Query<(Entity, &ParticleSpawner), With<ParticleSpawnerData>>) {
    for (entity, _settings) in &q {
        // the reference's condition is only `all empty && !active()` (:679-682) -- a spawner that starts
        // disabled, or a OneShot burst with a Nested trail, finishes too -- so every spawner is asked;
        // fw_spawner_status_get is served from the frame's own state readback (one event wait per
        // frame, no copy per spawner)
        let mut st = fw_spawner_status::default();
        if gpu.check(unsafe { fw_spawner_status_get(gpu.ctx, key(entity), &mut st) }, "fw_spawner_status_get") && st.finished != 0 {
            commands.trigger(ParticleSpawnerFinished { entity });
            unsafe { fw_spawner_mark_finished_notified(gpu.ctx, key(entity)) };
        }
    }
}

/// entity despawned: drop its streams
pub fn gpu_forget_removed(gpu: ResMut<GpuParticles>, mut removed: RemovedComponents<ParticleSpawner>) {
    for entity in removed.read() {
        unsafe { fw_spawner_remove(gpu.ctx, key(entity)) };
    }
}

/// `SpatialQuery::cast_ray` (src/core.rs:756-765) sees avian's colliders as of the last physics
/// step; the library sees what this system sends. Same count as last time = an asynchronous
/// re-upload (moving colliders), so sending every frame is cheap. Shapes the library cannot test
/// (anything but cuboid / ball / cylinder / cone / capsule) are skipped with a warning, once.
#[cfg(feature = "physics_avian")]
pub fn gpu_sync_colliders(
    gpu: ResMut<GpuParticles>,
    q: Query<(Entity, &Collider, &GlobalTransform, Option<&CollisionLayers>)>,
    mut scratch: Local<Vec<fw_collider>>,
) {
    use avian3d::parry::shape::TypedShape;
    scratch.clear();
    for (entity, collider, transform, layers) in &q {
        let mut t = transform.compute_transform();
        let (kind, half_extents) = match collider.shape_scaled().as_typed_shape() {
            TypedShape::Cuboid(c) => (FW_COLLIDER_CUBOID, [c.half_extents.x, c.half_extents.y, c.half_extents.z]),
            TypedShape::Ball(b) => (FW_COLLIDER_SPHERE, [b.radius, 0.0, 0.0]),
            TypedShape::Cylinder(c) => (FW_COLLIDER_CYLINDER, [c.radius, c.half_height, 0.0]),
            TypedShape::Cone(c) => (FW_COLLIDER_CONE, [c.radius, c.half_height, 0.0]),
            TypedShape::Capsule(c) => {
                // parry's capsule is any segment a..b swept by a ball; the C ABI's is about +Y through the
                // collider's origin (what Collider::capsule(radius, length) builds): fold the rest into the pose
                let (a, b) = (Vec3::from(c.segment.a), Vec3::from(c.segment.b));
                let axis = b - a;
                if let Some(dir) = axis.try_normalize() {
                    t.translation += t.rotation * ((a + b) * 0.5);
                    t.rotation *= Quat::from_rotation_arc(Vec3::Y, dir);
                }
                (FW_COLLIDER_CAPSULE, [c.radius, axis.length() * 0.5, 0.0])
            }
            _ => {
                bevy::log::warn_once!("firework_b200: collider shape not supported by the GPU collision sweep, ignored");
                continue;
            }
        };
        scratch.push(fw_collider {
            kind,
            layers: layers.map(|l| l.memberships.0).unwrap_or(1),
            key: key(entity),
            half_extents,
            translation: t.translation.to_array(),
            rotation: t.rotation.to_array(),
        });
    }
    let rc = unsafe { fw_set_colliders(gpu.ctx, scratch.as_ptr(), scratch.len() as u32) };
    gpu.check(rc, "fw_set_colliders");
}

/// in `impl Plugin for ParticleSystemPlugin` (src/plugin.rs:35-61) the chain becomes:
pub fn add_gpu_systems(app: &mut App, schedule: impl bevy::ecs::schedule::ScheduleLabel + Clone) {
    app.insert_resource(GpuParticles::new(0, 0x00F1_2E00).expect("firework_b200: no B200 / library"));
    app.add_systems(
        schedule,
        (
            ApplyDeferred,
            propagate_particle_spawner_modifier,
            ApplyDeferred,
            gpu_forget_removed,
            gpu_sync_spawner_data,
            #[cfg(feature = "physics_avian")]
            sync_parent_velocity,
            #[cfg(feature = "physics_avian")]
            gpu_sync_colliders,
            gpu_frame,
            gpu_notify_finished,
        )
            .chain(),
    );
}
