// sparks.cpp -- /root/reference/examples/sparks.rs:15-87 on the C++ host mirror: one spawner at
// (0, 0.1, 0), Circle{Y, 0.3} emitter, rate(1000), lifetime 0.75, the five-knot fire gradient.
// Prints the live count of every frame (the test compares them with the CPU oracle) and a few
// rows. usage: sparks [frames] [rate]
#include <cstdio>
#include <cstdlib>

#include "../firework.hpp"

using namespace firework;

static ParticleSpawner sparks(float rate) {
    ParticleSpawner s;
    ParticleSettings &p = s.particle_settings[0];
    p.lifetime = RandF32::constant(0.75f);
    p.initial_scale = RandF32{0.02f, 0.08f};
    p.scale_curve = FireworkCurve::constant(1.f);
    p.base_color = FireworkGradient::uneven_samples({{0.f, {150.f, 100.f, 15.f, 1.f}},
                                                     {0.7f, {3.f, 1.f, 1.f, 1.f}},
                                                     {0.8f, {1.f, 0.3f, 0.3f, 1.f}},
                                                     {0.9f, {0.3f, 0.3f, 0.3f, 1.f}},
                                                     {1.f, {0.1f, 0.1f, 0.1f, 0.f}}});
    p.linear_drag = 0.1f;
    p.pbr = false;
    EmissionSettings &e = s.emission_settings[0];
    e.emission_pacing = EmissionPacing::rate(rate);
    e.emission_shape = EmissionShape::Circle({0.f, 1.f, 0.f}, 0.3f);
    e.inherit_parent_velocity = true;
    e.initial_velocity = RandVec3{RandF32{0.f, 10.f}, {0.f, 1.f, 0.f}, 30.f / 180.f * 3.14159265358979323846f};
    return s;
}

int main(int argc, char **argv) {
    const int frames = argc > 1 ? std::atoi(argv[1]) : 120;
    const float rate = argc > 2 ? (float)std::atof(argv[2]) : 1000.f;
    try {
        App app;
        app.add_plugins(ParticleSystemPlugin{});
        const Entity e = app.spawn(sparks(rate), Transform::from_xyz(0.f, 0.1f, 0.f));
        std::printf("{\"entity\": %u, \"counts\": [", e);
        for (int k = 0; k < frames; k++) {
            app.update(1.f / 60.f);
            std::printf("%s%u", k ? ", " : "", app.data(e).counts()[0]);
        }
        const auto rows = app.data(e).particles(0);
        double sum_age = 0.0, sum_y = 0.0;
        for (const ParticleData &p : rows) {
            sum_age += p.age;
            sum_y += p.position[1];
        }
        std::printf("], \"live\": %zu, \"sum_age\": %.9g, \"sum_y\": %.9g, \"active\": %d}\n", rows.size(), sum_age, sum_y,
                    (int)app.data(e).active());
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "sparks: %s\n", ex.what());
        return 1;
    }
    return 0;
}
