/*
 * fw_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of bevy_firework's per-frame particle path, written from the
 * reference's Rust sources (cited per function in fw_oracle.c). Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it. The shipped
 * library (bevy_firework_b200/csrc) never includes, links or calls anything in oracle/.
 *
 * It shares ONLY the POD settings structs of the public ABI header, so that the same settings
 * bytes can be handed to the oracle and to the CUDA path, and include/fw_sincos.h: the sine / cosine
 * is defined by the library in IEEE operations (a platform libm is not a specification) and both
 * sides compile that one definition; its accuracy is pinned separately (tests/test_sincos.py,
 * scripts/sincos_exhaustive.c: correctly rounded against an 80-bit libm for every finite float).
 *
 * PARITY STATUS
 *   pinned by the reference's own tests: compute_emission_count (src/core.rs:806-834) and
 *     the even colour curve at its three knots (src/curve.rs:245-258);
 *   pinned only by the source text: update_particles, particle_collision, spawn_particles;
 *   PARITY UNPINNED (third-party crates absent from /root/reference, formulas restated from
 *     their published behaviour or DEFINED by this build -- see DESIGN.md section 4):
 *     bevy_utilitarian 0.10.0 RandF32/RandVec3/PitchYaw, rand 0.9.4 (unseedable; replaced by
 *     a Philox4x32-10 protocol), bevy_math 0.19.0 curve cores, bevy_color 0.19.0 Mix,
 *     glam 0.32.1 Quat/Vec3 scalar formulas, avian3d 0.7.0 / parry3d 0.27.0 cast_ray.
 */
#ifndef FW_ORACLE_H
#define FW_ORACLE_H

#include "../include/firework_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fwo_world fwo_world;

/* the library's sine / cosine (include/fw_sincos.h) as compiled into this oracle */
void fwo_sincosf(float x, float *s, float *c);
void fwo_sincosf_array(const float *x, uint64_t n, float *s, float *c);

fwo_world *fwo_create(uint64_t seed);
void fwo_destroy(fwo_world *w);

int fwo_spawner_reset(fwo_world *w, uint32_t key, const fw_particle_settings *ps, uint32_t n_types,
                      const fw_emission_settings *es, uint32_t n_emitters, uint32_t starts_enabled);
int fwo_spawner_remove(fwo_world *w, uint32_t key);
void fwo_set_colliders(fwo_world *w, const fw_collider *c, uint32_t n);
/* TEST HELPER for the full-size collision scenes: skip (with conservative boxes) colliders the ray
 * segment cannot reach; results equal the brute-force loop (checked in tests/test_oracle_golden.py).
 * Off by default; the CPU baseline that bench.py times never turns it on. */
void fwo_set_cull(fwo_world *w, int on);

/* spawn_particles then update_particles; n_threads tasks-per-spawner pool for the update
 * (mirrors Query::par_iter_mut, src/core.rs:583-585); spawn is sequential (:377). */
void fwo_frame(fwo_world *w, float dt, const fw_spawner_frame_input *in, uint32_t n_in,
               uint32_t n_threads);
/* only one of the two systems (unit tests) */
void fwo_spawn_only(fwo_world *w, float dt, const fw_spawner_frame_input *in, uint32_t n_in);
void fwo_update_only(fwo_world *w, float dt, uint32_t n_threads);

uint64_t fwo_count(const fwo_world *w, uint32_t key, uint32_t type);
uint64_t fwo_total_live(const fwo_world *w);
int fwo_read_particles(const fwo_world *w, uint32_t key, uint32_t type, fw_particle_data *out,
                       uint64_t cap, uint64_t *n);
int fwo_write_particles(fwo_world *w, uint32_t key, uint32_t type, const fw_particle_data *in,
                        uint64_t n);
int fwo_read_destroyed(const fwo_world *w, uint32_t key, uint32_t type, fw_particle_data *out,
                       uint64_t cap, uint64_t *n);
int fwo_status(fwo_world *w, uint32_t key, fw_spawner_status *out);
int fwo_mark_finished_notified(fwo_world *w, uint32_t key);
int fwo_read_aabb(const fwo_world *w, uint32_t key, float mn[3], float mx[3], uint32_t *empty);

/* pure pieces, exported for known-answer tests */
void fwo_compute_emission_count(float time_passed_in_cycle, float last_emission,
                                float cycle_duration, float offset_start, float offset_end,
                                float particles_per_cycle, uint64_t *n_emit, float *next_last);
float fwo_sample_curve(const fw_curve_f32 *c, float t);
void fwo_sample_gradient(const fw_gradient *g, float t, float out[4]);
void fwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float fwo_uniform(uint64_t seed, uint32_t spawner_key, uint32_t emitter, uint64_t serial,
                  uint32_t draw);
void fwo_generate_point(const fw_emission_settings *e, float u0, float u1, float u2, float out[3]);
void fwo_rand_vec3(const fw_rand_vec3 *r, float u_angle, float u_radius, float u_mag, float out[3]);
int fwo_cast_ray(const fw_collider *c, uint32_t n, uint32_t filter_mask, const float origin[3],
                 const float dir[3], float max_distance, float *distance, float normal[3],
                 uint32_t *index);
void fwo_particle_collision(const fw_collider *c, uint32_t n, const fw_collision_settings *cs,
                            float pos[3], float vel[3], float delta, uint32_t *should_destroy);
void fwo_quat_from_scaled_axis(const float v[3], float out[4]);
void fwo_quat_mul(const float a[4], const float b[4], float out[4]);
void fwo_quat_from_rotation_arc(const float from[3], const float to[3], float out[4]);
void fwo_quat_mul_vec3(const float q[4], const float v[3], float out[3]);
void fwo_vec3_normalize_or_zero(const float v[3], float out[3]);
void fwo_vec3_project_onto(const float a[3], const float b[3], float out[3]);
void fwo_vec3_reject_from(const float a[3], const float b[3], float out[3]);
void fwo_pitch_yaw_to_unit_vec(float u, float v, float out[3]);
float fwo_rem_euclid(float a, float b);
float fwo_div_euclid(float a, float b);

#ifdef __cplusplus
}
#endif
#endif
