/* sparks_c99.c -- /root/reference/examples/sparks.rs:15-87 through the bare C ABI of
 * include/firework_b200.h: plain C99, no C++ mirror, no Python. One spawner at (0, 0.1, 0), Circle{Y,
 * 0.3} emitter, rate(1000), lifetime 0.75 s, the five-knot fire gradient. Prints the live count of
 * every frame and a checksum of the last frame's rows as JSON (tests/test_gpu_cpp_host.py compares
 * them with the CPU oracle).   usage: sparks_c99 [frames] [rate] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/firework_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != FW_OK) {                                                              \
            fprintf(stderr, "sparks_c99: %s -> %d: %s\n", #call, rc_, ctx ? fw_last_error(ctx) : fw_last_global_error()); \
            if (ctx) fw_destroy(ctx);                                                    \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char **argv) {
    const int frames = argc > 1 ? atoi(argv[1]) : 120;
    const float rate = argc > 2 ? (float)atof(argv[2]) : 1000.0f;
    fw_context *ctx = NULL;
    fw_config cfg;
    fw_particle_settings ps;
    fw_emission_settings es;
    fw_spawner_frame_input in;
    fw_stream_layout lay;
    static const float knots_t[5] = {0.0f, 0.7f, 0.8f, 0.9f, 1.0f};
    static const float knots_c[5][4] = {{150.f, 100.f, 15.f, 1.f}, {3.f, 1.f, 1.f, 1.f}, {1.f, .3f, .3f, 1.f}, {.3f, .3f, .3f, 1.f}, {.1f, .1f, .1f, 0.f}};
    int k;

    if (fw_abi_sizeof("fw_particle_settings") != sizeof(fw_particle_settings) || fw_abi_version() != FW_ABI_VERSION) {
        fprintf(stderr, "sparks_c99: header and library disagree\n");
        return 1;
    }
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = FW_ABI_VERSION;
    cfg.device = 0;
    cfg.seed = 0x00F12E00u;
    CHECK(fw_create(&cfg, &ctx));

    /* ParticleSettings (examples/sparks.rs:49-66; the rest are the defaults of src/core.rs:187-211) */
    memset(&ps, 0, sizeof(ps));
    ps.lifetime.min = ps.lifetime.max = 0.75f;
    ps.initial_scale.min = 0.02f;
    ps.initial_scale.max = 0.08f;
    ps.scale_curve.kind = FW_CURVE_CONSTANT;
    ps.scale_curve.n = 1;
    ps.scale_curve.values[0] = 1.0f;
    ps.acceleration[1] = -9.81f;
    ps.linear_drag = 0.1f;
    ps.angular_drag = 0.2f;
    ps.base_color.kind = FW_CURVE_UNEVEN;
    ps.base_color.n = 5;
    memcpy(ps.base_color.times, knots_t, sizeof(knots_t));
    memcpy(ps.base_color.colors, knots_c, sizeof(knots_c));
    ps.emissive_color.kind = FW_CURVE_CONSTANT; /* LinearRgba::BLACK */
    ps.emissive_color.n = 1;
    ps.emissive_color.colors[0][3] = 1.0f;
    /* EmissionSettings (examples/sparks.rs:69-82) */
    memset(&es, 0, sizeof(es));
    es.pacing_kind = FW_PACING_COUNT_OVER_DURATION; /* EmissionPacing::rate(r) */
    es.count = rate;
    es.duration = 1.0f;
    es.offset_end = 1.0f;
    es.shape_kind = FW_SHAPE_CIRCLE;
    es.shape_radius = 0.3f;
    es.shape_normal[1] = 1.0f;
    es.initial_velocity.magnitude.max = 10.0f;
    es.initial_velocity.direction[1] = 1.0f;
    es.initial_velocity.spread = 30.0f / 180.0f * 3.14159265358979323846f;
    es.initial_angular_velocity.direction[1] = 1.0f;
    es.inherit_parent_velocity = 1;
    es.initial_rotation[3] = 1.0f;
    CHECK(fw_spawner_reset(ctx, 1u, &ps, 1u, &es, 1u, 1u));
    CHECK(fw_stream_layout_get(ctx, 1u, 0u, &lay));

    memset(&in, 0, sizeof(in));
    in.spawner_key = 1u;
    in.origin_translation[1] = 0.1f;
    in.origin_rotation[3] = 1.0f;
    in.modifier_scale = in.modifier_speed = 1.0f;
    printf("{\"bytes_per_particle\": %u, \"counts\": [", lay.bytes_read + lay.bytes_written);
    for (k = 0; k < frames; k++) {
        uint32_t count = 0;
        CHECK(fw_frame(ctx, 1.0f / 60.0f, &in, 1u));
        CHECK(fw_counts(ctx, 1u, &count, 1u));
        printf("%s%u", k ? ", " : "", count);
    }
    {
        uint64_t n = 0, i;
        double sum_age = 0.0, sum_y = 0.0;
        fw_particle_data *rows;
        fw_read_particles(ctx, 1u, 0u, NULL, 0u, &n); /* sizes the buffer */
        rows = (fw_particle_data *)malloc((n ? n : 1) * sizeof(*rows));
        CHECK(fw_read_particles(ctx, 1u, 0u, rows, n, &n));
        for (i = 0; i < n; i++) {
            sum_age += rows[i].age;
            sum_y += rows[i].position[1];
        }
        printf("], \"live\": %llu, \"sum_age\": %.9g, \"sum_y\": %.9g}\n", (unsigned long long)n, sum_age, sum_y);
        free(rows);
    }
    fw_destroy(ctx);
    return 0;
}
