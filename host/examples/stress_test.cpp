// stress_test.cpp -- /root/reference/examples/stress_test.rs:91-129 scaled the way BASELINE.json
// asks (C2: 64 spawners x rate 15625 => ~1 M particles; C3: 512 x 19531 => ~10 M), on the C++
// host mirror. Prints live particles (what the example's DebugInfo overlay shows, :186-201) and
// particles updated per second. usage: stress_test [spawners] [rate] [frames]
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../firework.hpp"

using namespace firework;

static ParticleSpawner stress(float rate) {
    ParticleSpawner s;
    ParticleSettings &p = s.particle_settings[0];
    p.lifetime = RandF32::constant(1.f);
    p.initial_scale = RandF32{0.02f, 0.08f};
    p.base_color = FireworkGradient::uneven_samples({{0.f, {10.f, 7.f, 1.f, 1.f}},
                                                     {0.7f, {3.f, 1.f, 1.f, 1.f}},
                                                     {0.8f, {1.f, 0.3f, 0.3f, 1.f}},
                                                     {0.9f, {0.3f, 0.3f, 0.3f, 1.f}},
                                                     {1.f, {0.1f, 0.1f, 0.1f, 0.f}}});
    p.linear_drag = 0.1f;
    EmissionSettings &e = s.emission_settings[0];
    e.emission_pacing = EmissionPacing::rate(rate);
    e.emission_shape = EmissionShape::Circle({0.f, 1.f, 0.f}, 0.3f);
    e.initial_velocity = RandVec3{RandF32{0.f, 10.f}, {0.f, 1.f, 0.f}, 30.f / 180.f * 3.14159265358979323846f};
    return s;
}

int main(int argc, char **argv) {
    const int n_spawners = argc > 1 ? std::atoi(argv[1]) : 64;
    const float rate = argc > 2 ? (float)std::atof(argv[2]) : 15625.f;
    const int frames = argc > 3 ? std::atoi(argv[3]) : 200;
    try {
        App app;
        app.add_plugins(ParticleSystemPlugin{});
        std::vector<Entity> ents;
        for (int i = 0; i < n_spawners; i++) ents.push_back(app.spawn(stress(rate), Transform::from_xyz(2.f * (i % 32), 0.1f, 2.f * (i / 32))));
        for (int k = 0; k < 64; k++) app.update(1.f / 60.f); // reach the stationary live count
        uint64_t live = 0;
        app.check(fw_total_live(app.context(), &live));
        app.check(fw_sync(app.context()));
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < frames; k++) app.update(1.f / 60.f);
        app.check(fw_sync(app.context()));
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t live2 = 0;
        app.check(fw_total_live(app.context(), &live2));
        std::printf("{\"spawners\": %d, \"live_particles\": %llu, \"frames\": %d, \"ms_per_frame\": %.4f, \"particles_per_s\": %.4g}\n", n_spawners,
                    (unsigned long long)live2, frames, s / frames * 1e3, 0.5 * (double)(live + live2) * frames / s);
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "stress_test: %s\n", ex.what());
        return 1;
    }
    return 0;
}
