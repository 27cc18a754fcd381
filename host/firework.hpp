// firework.hpp -- C++ host side above the C ABI (include/firework_b200.h).
//
// The reference's host code is Rust; no Rust toolchain exists in this image, so the host mirror
// of its plugin / component interface is written in C++ (header-only, C++17). Type and field
// names follow /root/reference/src/core.rs, src/curve.rs, src/emission_shape.rs, src/plugin.rs
// so that a program written against it reads like the reference's examples:
//
//     App app;
//     app.add_plugins(ParticleSystemPlugin{});
//     Entity e = app.spawn(sparks(), Transform::from_xyz(0.f, 0.1f, 0.f));
//     app.update(1.f / 60.f);
//     auto rows = app.data(e).particles(0);          // Vec<ParticleData>
//
// Nothing here computes particle state: every per-particle operation happens on the GPU behind
// fw_frame. Errors of the C ABI are thrown as firework::Error (the Rust shim logs them instead).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/firework_b200.h"

namespace firework {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("firework_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

using Vec3 = std::array<float, 3>;
using Quat = std::array<float, 4>; // x, y, z, w
using Entity = uint32_t;

// ---- bevy_utilitarian (src/core.rs:102,107,155,157,161)
struct RandF32 {
    float min = 0.f, max = 0.f;
    static RandF32 constant(float v) { return RandF32{v, v}; }
    fw_rand_f32 pod() const { return fw_rand_f32{min, max}; }
};
struct RandVec3 {
    RandF32 magnitude;
    Vec3 direction{0.f, 0.f, 0.f};
    float spread = 0.f;
    static RandVec3 constant(Vec3 v) {
        const float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        RandVec3 r;
        r.magnitude = RandF32::constant(len);
        if (len > 0.f) r.direction = Vec3{v[0] / len, v[1] / len, v[2] / len};
        return r;
    }
    fw_rand_vec3 pod() const {
        fw_rand_vec3 p{};
        p.magnitude = magnitude.pod();
        std::memcpy(p.direction, direction.data(), sizeof(p.direction));
        p.spread = spread;
        return p;
    }
};

// ---- src/curve.rs
struct LinearRgba {
    float red = 1.f, green = 1.f, blue = 1.f, alpha = 1.f;
    static LinearRgba WHITE() { return {1.f, 1.f, 1.f, 1.f}; }
    static LinearRgba BLACK() { return {0.f, 0.f, 0.f, 1.f}; }
};
struct FireworkCurve { // FireworkCurve<f32>, src/curve.rs:8-75
    fw_curve_f32 p{};
    static FireworkCurve constant(float v) {
        FireworkCurve c;
        c.p.kind = FW_CURVE_CONSTANT;
        c.p.n = 1;
        c.p.values[0] = v;
        return c;
    }
    static FireworkCurve even_samples(const std::vector<float> &s) {
        if (s.empty()) throw std::invalid_argument("Cannot create curve from 0 samples");
        if (s.size() == 1) return constant(s[0]);
        if (s.size() > FW_MAX_KNOTS) throw std::invalid_argument("too many curve samples");
        FireworkCurve c;
        c.p.kind = FW_CURVE_EVEN;
        c.p.n = (uint32_t)s.size();
        for (size_t i = 0; i < s.size(); i++) c.p.values[i] = s[i];
        return c;
    }
    static FireworkCurve uneven_samples(const std::vector<std::pair<float, float>> &s) {
        if (s.empty()) throw std::invalid_argument("Cannot create curve from 0 samples");
        if (s.size() == 1) return constant(s[0].second);
        if (s.size() > FW_MAX_KNOTS) throw std::invalid_argument("too many curve samples");
        FireworkCurve c;
        c.p.kind = FW_CURVE_UNEVEN;
        c.p.n = (uint32_t)s.size();
        for (size_t i = 0; i < s.size(); i++) {
            c.p.times[i] = s[i].first;
            c.p.values[i] = s[i].second;
        }
        return c;
    }
};
struct FireworkGradient { // FireworkGradient<LinearRgba>, src/curve.rs:171-239
    fw_gradient p{};
    static void put(fw_gradient &g, size_t i, const LinearRgba &c) {
        g.colors[i][0] = c.red;
        g.colors[i][1] = c.green;
        g.colors[i][2] = c.blue;
        g.colors[i][3] = c.alpha;
    }
    static FireworkGradient constant(const LinearRgba &v) {
        FireworkGradient g;
        g.p.kind = FW_CURVE_CONSTANT;
        g.p.n = 1;
        put(g.p, 0, v);
        return g;
    }
    static FireworkGradient even_samples(const std::vector<LinearRgba> &s) {
        if (s.empty()) throw std::invalid_argument("Cannot create curve from 0 samples");
        if (s.size() == 1) return constant(s[0]);
        if (s.size() > FW_MAX_KNOTS) throw std::invalid_argument("too many gradient samples");
        FireworkGradient g;
        g.p.kind = FW_CURVE_EVEN;
        g.p.n = (uint32_t)s.size();
        for (size_t i = 0; i < s.size(); i++) put(g.p, i, s[i]);
        return g;
    }
    static FireworkGradient uneven_samples(const std::vector<std::pair<float, LinearRgba>> &s) {
        if (s.empty()) throw std::invalid_argument("Cannot create curve from 0 samples");
        if (s.size() == 1) return constant(s[0].second);
        if (s.size() > FW_MAX_KNOTS) throw std::invalid_argument("too many gradient samples");
        FireworkGradient g;
        g.p.kind = FW_CURVE_UNEVEN;
        g.p.n = (uint32_t)s.size();
        for (size_t i = 0; i < s.size(); i++) {
            g.p.times[i] = s[i].first;
            put(g.p, i, s[i].second);
        }
        return g;
    }
};

// ---- src/emission_shape.rs:6-16
struct EmissionShape {
    uint32_t kind = FW_SHAPE_POINT;
    float radius = 0.f;
    Vec3 normal{0.f, 1.f, 0.f};
    static EmissionShape Point() { return {}; }
    static EmissionShape Sphere(float r) { return EmissionShape{FW_SHAPE_SPHERE, r, {0.f, 1.f, 0.f}}; }
    static EmissionShape Circle(Vec3 normal, float r) { return EmissionShape{FW_SHAPE_CIRCLE, r, normal}; }
};

// ---- src/core.rs:11-97
struct EmissionPacing {
    uint32_t kind = FW_PACING_COUNT_OVER_DURATION;
    uint64_t one_shot_count = 0;
    float count = 5.f, duration = 1.f, offset_start = 0.f, offset_end = 1.f;
    static EmissionPacing OneShot(uint64_t n) { return EmissionPacing{FW_PACING_ONE_SHOT, n, 0.f, 1.f, 0.f, 1.f}; }
    static EmissionPacing OnDemand() { return EmissionPacing{FW_PACING_ON_DEMAND, 0, 0.f, 1.f, 0.f, 1.f}; }
    static EmissionPacing CountOverDuration(float count, float duration, float start, float end) {
        return EmissionPacing{FW_PACING_COUNT_OVER_DURATION, 0, count, duration, start, end};
    }
    static EmissionPacing rate(float r) { return CountOverDuration(r, 1.f, 0.f, 1.f); } // :36-43
    bool is_one_shot() const { return kind == FW_PACING_ONE_SHOT; }
};
struct EmissionMode {
    uint32_t kind = FW_MODE_GLOBAL;
    uint32_t target_particle_type = 0;
    static EmissionMode Global() { return {}; }
    static EmissionMode Nested(uint32_t target) { return EmissionMode{FW_MODE_NESTED, target}; }
};
enum class SpawnTransformMode { Global, Local };
enum class BlendMode : uint32_t { Opaque = 0, Blend = 2, Premultiplied = 3, Add = 4, Multiply = 5 };

struct ParticleCollisionSettings { // src/core.rs:240-248
    float restitution = 0.f, friction = 0.f;
    bool destroy_on_collision = false;
    uint32_t filter = 0xFFFFFFFFu;
};
using ParticleData = fw_particle_data; // src/core.rs:305-321
struct ParticleEventHandlers {          // src/core.rs:164-167
    std::function<void(const std::vector<ParticleData> &)> particles_destroyed;
};

struct ParticleSettings { // src/core.rs:99-142, defaults :187-211
    RandF32 lifetime = RandF32::constant(5.f);
    FireworkCurve scale_curve = FireworkCurve::constant(1.f);
    RandF32 initial_scale = RandF32::constant(1.f);
    Vec3 acceleration{0.f, -9.81f, 0.f};
    Vec3 angular_acceleration{0.f, 0.f, 0.f};
    float linear_drag = 0.2f, angular_drag = 0.2f;
    FireworkGradient base_color = FireworkGradient::constant(LinearRgba::WHITE());
    FireworkGradient emissive_color = FireworkGradient::constant(LinearRgba::BLACK());
    float fade_edge = 0.7f, fade_scene = 1.f;
    BlendMode blend_mode = BlendMode::Blend;
    bool pbr = false;
    std::optional<ParticleCollisionSettings> collision_settings;
    ParticleEventHandlers event_handlers;
    fw_particle_settings pod() const {
        fw_particle_settings p{};
        p.lifetime = lifetime.pod();
        p.scale_curve = scale_curve.p;
        p.initial_scale = initial_scale.pod();
        std::memcpy(p.acceleration, acceleration.data(), 12);
        std::memcpy(p.angular_acceleration, angular_acceleration.data(), 12);
        p.linear_drag = linear_drag;
        p.angular_drag = angular_drag;
        p.base_color = base_color.p;
        p.emissive_color = emissive_color.p;
        p.pbr = pbr;
        if (collision_settings) {
            p.collision.enabled = 1;
            p.collision.restitution = collision_settings->restitution;
            p.collision.friction = collision_settings->friction;
            p.collision.destroy_on_collision = collision_settings->destroy_on_collision;
            p.collision.filter_mask = collision_settings->filter;
        }
        p.capture_destroyed = (bool)event_handlers.particles_destroyed;
        return p;
    }
};
struct EmissionSettings { // src/core.rs:144-162, defaults :213-227
    uint32_t particle_index = 0;
    EmissionPacing emission_pacing = EmissionPacing::rate(5.f);
    EmissionMode emission_mode;
    EmissionShape emission_shape;
    RandVec3 initial_velocity = RandVec3::constant({0.f, 0.f, 0.f});
    RandF32 initial_velocity_radial = RandF32::constant(0.f);
    bool inherit_parent_velocity = true;
    Quat initial_rotation{0.f, 0.f, 0.f, 1.f};
    RandVec3 initial_angular_velocity = RandVec3::constant({0.f, 0.f, 0.f});
    fw_emission_settings pod() const {
        fw_emission_settings p{};
        p.particle_index = particle_index;
        p.pacing_kind = emission_pacing.kind;
        p.one_shot_count = emission_pacing.one_shot_count;
        p.count = emission_pacing.count;
        p.duration = emission_pacing.duration;
        p.offset_start = emission_pacing.offset_start;
        p.offset_end = emission_pacing.offset_end;
        p.mode = emission_mode.kind;
        p.target_particle_type = emission_mode.target_particle_type;
        p.shape_kind = emission_shape.kind;
        p.shape_radius = emission_shape.radius;
        std::memcpy(p.shape_normal, emission_shape.normal.data(), 12);
        p.initial_velocity = initial_velocity.pod();
        p.initial_velocity_radial = initial_velocity_radial.pod();
        p.inherit_parent_velocity = inherit_parent_velocity;
        std::memcpy(p.initial_rotation, initial_rotation.data(), 16);
        p.initial_angular_velocity = initial_angular_velocity.pod();
        return p;
    }
};
struct ParticleSpawner { // src/core.rs:169-185, defaults :229-238
    std::vector<ParticleSettings> particle_settings{ParticleSettings{}};
    std::vector<EmissionSettings> emission_settings{EmissionSettings{}};
    bool starts_enabled = true;
    SpawnTransformMode spawn_transform_mode = SpawnTransformMode::Global;
};
struct EffectModifier { // src/core.rs:323-336
    float scale = 1.f, speed = 1.f;
};
struct ParticleSpawnerFinished { // src/core.rs:338-341
    Entity entity;
};

struct Transform {
    Vec3 translation{0.f, 0.f, 0.f};
    Quat rotation{0.f, 0.f, 0.f, 1.f};
    static Transform from_xyz(float x, float y, float z) { return Transform{{x, y, z}, {0.f, 0.f, 0.f, 1.f}}; }
    static Vec3 rotate(const Quat &q, const Vec3 &v) {
        const float x = q[0], y = q[1], z = q[2], w = q[3];
        const float b2 = x * x + y * y + z * z, d = v[0] * x + v[1] * y + v[2] * z;
        const Vec3 c{y * v[2] - z * v[1], z * v[0] - x * v[2], x * v[1] - y * v[0]};
        const float k = w * w - b2, m = 2.f * d, n = 2.f * w;
        return Vec3{v[0] * k + x * m + c[0] * n, v[1] * k + y * m + c[1] * n, v[2] * k + z * m + c[2] * n};
    }
    Transform mul_transform(const Transform &child) const {
        const Vec3 r = rotate(rotation, child.translation);
        const Quat &a = rotation, &b = child.rotation;
        return Transform{{translation[0] + r[0], translation[1] + r[1], translation[2] + r[2]},
                         {a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0],
                          a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3], a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]}};
    }
};

class App;
// ParticleSpawnerData (src/core.rs:269-303): a handle onto device state
class ParticleSpawnerData {
  public:
    bool initialized = true;
    Vec3 parent_velocity{0.f, 0.f, 0.f};
    uint64_t manual_queued_count = 0;
    void queue_particles(uint64_t n) { manual_queued_count += n; } // :284-286
    std::vector<uint32_t> counts() const;                          // data.particles[i].len()
    std::vector<ParticleData> particles(uint32_t type) const;      // lazily mirrored data.particles[type]
    std::vector<fw_particle_instance> instances(uint32_t type) const;
    bool active() const;                                           // :288-302

  private:
    friend class App;
    fw_context *ctx_ = nullptr;
    Entity key_ = 0;
    uint32_t n_types_ = 0;
};

struct ParticleSystemPlugin { // src/plugin.rs:22-32
    std::string update_schedule = "Update";
    int device = 0;
    uint64_t seed = 0x00F12E00ull;
};

class App {
  public:
    App() = default;
    App(const App &) = delete;
    ~App() {
        if (ctx_) fw_destroy(ctx_);
    }
    App &add_plugins(const ParticleSystemPlugin &p) { // Plugin::build, src/plugin.rs:35-61
        fw_config cfg{};
        cfg.abi_version = FW_ABI_VERSION;
        cfg.device = p.device;
        cfg.seed = p.seed;
        const int rc = fw_create(&cfg, &ctx_);
        if (rc != FW_OK) throw Error(rc, fw_last_global_error());
        return *this;
    }
    Entity spawn(std::optional<ParticleSpawner> spawner, Transform transform = {}, std::optional<Entity> parent = {},
                 std::optional<EffectModifier> modifier = {}) {
        const Entity id = next_id_++;
        Ent e;
        e.spawner = std::move(spawner);
        e.transform = transform;
        e.parent = parent;
        e.modifier = modifier;
        entities_.emplace(id, std::move(e));
        return id;
    }
    void despawn(Entity id) {
        auto it = entities_.find(id);
        if (it == entities_.end()) return;
        if (it->second.has_data) fw_spawner_remove(ctx_, id);
        entities_.erase(it);
    }
    void observe(Entity id, std::function<void(const ParticleSpawnerFinished &)> cb) { entities_.at(id).observers.push_back(std::move(cb)); }
    ParticleSpawner &spawner_mut(Entity id) { // Mut<ParticleSpawner>: marks the component changed (src/core.rs:344)
        Ent &e = entities_.at(id);
        e.changed = true;
        return *e.spawner;
    }
    Transform &transform_mut(Entity id) { return entities_.at(id).transform; }
    ParticleSpawnerData &data(Entity id) { return entities_.at(id).data; }
    fw_context *context() { return ctx_; }
    void set_colliders(const std::vector<fw_collider> &c) { check(fw_set_colliders(ctx_, c.data(), (uint32_t)c.size())); }

    // the system chain of src/plugin.rs:46-60
    void update(float dt) {
        // propagate_particle_spawner_modifier (src/core.rs:690-703)
        for (auto &kv : entities_)
            if (kv.second.modifier)
                for (auto &kv2 : entities_)
                    if (kv2.second.spawner && is_descendant(kv2.first, kv.first)) kv2.second.modifier = kv.second.modifier;
        // sync_spawner_data for Changed<ParticleSpawner> (:343-365)
        for (auto &kv : entities_) {
            Ent &e = kv.second;
            if (!e.spawner || !e.changed) continue;
            std::vector<fw_particle_settings> ps;
            std::vector<fw_emission_settings> es;
            for (const auto &s : e.spawner->particle_settings) ps.push_back(s.pod());
            for (const auto &s : e.spawner->emission_settings) es.push_back(s.pod());
            check(fw_spawner_reset(ctx_, kv.first, ps.data(), (uint32_t)ps.size(), es.data(), (uint32_t)es.size(), e.spawner->starts_enabled));
            e.data.ctx_ = ctx_;
            e.data.key_ = kv.first;
            e.data.n_types_ = (uint32_t)ps.size();
            e.has_data = true;
            e.changed = false;
        }
        // spawn_particles ; update_particles (:367-670) -> one batched call
        inputs_.clear();
        for (auto &kv : entities_) {
            Ent &e = kv.second;
            if (!e.spawner) continue;
            const Transform origin = e.spawner->spawn_transform_mode == SpawnTransformMode::Global ? global_transform(kv.first) : e.transform; // :432-435
            const EffectModifier m = e.modifier.value_or(EffectModifier{});
            fw_spawner_frame_input in{};
            in.spawner_key = kv.first;
            std::memcpy(in.origin_translation, origin.translation.data(), 12);
            std::memcpy(in.origin_rotation, origin.rotation.data(), 16);
            std::memcpy(in.parent_velocity, e.data.parent_velocity.data(), 12);
            in.modifier_scale = m.scale;
            in.modifier_speed = m.speed;
            in.queue_particles = (uint32_t)e.data.manual_queued_count;
            e.data.manual_queued_count = 0;
            inputs_.push_back(in);
        }
        check(fw_frame(ctx_, dt, inputs_.data(), (uint32_t)inputs_.size()));
        // particles_destroyed handlers (:660-667)
        for (auto &kv : entities_) {
            Ent &e = kv.second;
            if (!e.spawner) continue;
            for (uint32_t t = 0; t < e.spawner->particle_settings.size(); t++) {
                auto &h = e.spawner->particle_settings[t].event_handlers.particles_destroyed;
                if (!h) continue;
                uint64_t n = 0;
                fw_read_destroyed(ctx_, kv.first, t, nullptr, 0, &n);
                if (!n) continue;
                std::vector<ParticleData> rows(n);
                check(fw_read_destroyed(ctx_, kv.first, t, rows.data(), n, &n));
                h(rows);
            }
        }
        // notify_finished_particle_spawners (:674-688)
        std::vector<Entity> finished;
        for (auto &kv : entities_) {
            if (!kv.second.spawner || kv.second.observers.empty()) continue;
            fw_spawner_status st{};
            check(fw_spawner_status_get(ctx_, kv.first, &st));
            if (st.finished) {
                fw_spawner_mark_finished_notified(ctx_, kv.first);
                finished.push_back(kv.first);
            }
        }
        for (Entity id : finished) {
            auto it = entities_.find(id);
            if (it == entities_.end()) continue;
            auto cbs = it->second.observers; // a callback may despawn the entity
            for (auto &cb : cbs) cb(ParticleSpawnerFinished{id});
        }
    }
    void check(int rc) const {
        if (rc != FW_OK) throw Error(rc, fw_last_error(ctx_));
    }

  private:
    struct Ent {
        std::optional<ParticleSpawner> spawner;
        Transform transform;
        std::optional<Entity> parent;
        std::optional<EffectModifier> modifier;
        ParticleSpawnerData data;
        bool has_data = false, changed = true;
        std::vector<std::function<void(const ParticleSpawnerFinished &)>> observers;
    };
    bool is_descendant(Entity id, Entity ancestor) const {
        auto it = entities_.find(id);
        while (it != entities_.end() && it->second.parent) {
            if (*it->second.parent == ancestor) return true;
            it = entities_.find(*it->second.parent);
        }
        return false;
    }
    Transform global_transform(Entity id) const {
        const Ent &e = entities_.at(id);
        if (!e.parent || !entities_.count(*e.parent)) return e.transform;
        return global_transform(*e.parent).mul_transform(e.transform);
    }
    fw_context *ctx_ = nullptr;
    std::map<Entity, Ent> entities_;
    std::vector<fw_spawner_frame_input> inputs_;
    Entity next_id_ = 1;
};

inline std::vector<uint32_t> ParticleSpawnerData::counts() const {
    std::vector<uint32_t> out(n_types_);
    const int rc = fw_counts(ctx_, key_, out.data(), n_types_);
    if (rc != FW_OK) throw Error(rc, fw_last_error(ctx_));
    return out;
}
inline std::vector<ParticleData> ParticleSpawnerData::particles(uint32_t type) const {
    uint64_t n = 0;
    fw_read_particles(ctx_, key_, type, nullptr, 0, &n);
    std::vector<ParticleData> rows(n);
    if (n) {
        const int rc = fw_read_particles(ctx_, key_, type, rows.data(), n, &n);
        if (rc != FW_OK) throw Error(rc, fw_last_error(ctx_));
    }
    return rows;
}
inline std::vector<fw_particle_instance> ParticleSpawnerData::instances(uint32_t type) const {
    uint64_t n = 0;
    fw_read_instances(ctx_, key_, type, nullptr, 0, &n);
    std::vector<fw_particle_instance> rows(n);
    if (n) {
        const int rc = fw_read_instances(ctx_, key_, type, rows.data(), n, &n);
        if (rc != FW_OK) throw Error(rc, fw_last_error(ctx_));
    }
    return rows;
}
inline bool ParticleSpawnerData::active() const {
    fw_spawner_status st{};
    const int rc = fw_spawner_status_get(ctx_, key_, &st);
    if (rc != FW_OK) throw Error(rc, fw_last_error(ctx_));
    return st.active != 0;
}

} // namespace firework
