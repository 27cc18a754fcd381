/* One-off campaign (CPU): hundreds of millions of rays against random colliders of every kind.
 * Property the kernels' broad phase relies on (fw_math.cuh, cast_ray): whichever collider the brute-force
 * loop of the oracle reports as the closest hit, the ray segment's box overlaps that collider's box as
 * fw_set_colliders builds it (taken from the real library: fw_host_build_broadphase). Also: brute force ==
 * the oracle's own culled helper. Rays: random, aimed, along axes / cone slants / faces, tangent, tiny
 * perturbations of those, origins on and near surfaces, particle-step and long distances.
 *   gcc -O2 -std=gnu11 -ffp-contract=off -fno-fast-math -fopenmp -Iinclude -Ioracle scripts/probes/fuzz_cast_ray.c -o scripts/probes/fuzz_cast_ray -lm -ldl
 *   scripts/probes/fuzz_cast_ray bevy_firework_b200/libfirework_b200.so <scenes> <rays per scene> */
#include "../../oracle/fw_oracle.c"
#include <dlfcn.h>
#include <stdio.h>

typedef int (*build_fn)(const fw_collider *, uint32_t, void *, uint64_t, uint64_t *);
static inline uint64_t rnd(uint64_t *s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return *s = x; }
static inline float uni(uint64_t *s, float a, float b) { return a + (b - a) * (float)((rnd(s) >> 40) * (1.0 / 16777216.0)); }
static void rquat(uint64_t *s, float q[4]) {
    float n = 0;
    do { n = 0; for (int i = 0; i < 4; i++) { q[i] = uni(s, -1, 1); n += q[i] * q[i]; } } while (n < 1e-3f || n > 1.0f);
    n = sqrtf(n);
    for (int i = 0; i < 4; i++) q[i] /= n;
}
int main(int argc, char **argv) {
    if (argc < 4) return 2;
    void *h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    build_fn build = (build_fn)dlsym(h, "fw_host_build_broadphase");
    const int scenes = atoi(argv[2]);
    const long rays = atol(argv[3]);
    enum { N = 24 };
    unsigned long long total = 0, hits = 0, bad_box = 0, bad_cull = 0, far_hits = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total, hits, bad_box, bad_cull, far_hits)
    for (int sc = 0; sc < scenes; sc++) {
        uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(sc + 1);
        fw_collider c[N];
        memset(c, 0, sizeof c);
        for (int i = 0; i < N; i++) {
            c[i].kind = (uint32_t)(i % 5);
            c[i].layers = 1;
            c[i].key = 0xFFFFFFFFu;
            c[i].half_extents[0] = uni(&s, 0.1f, 0.9f);
            c[i].half_extents[1] = uni(&s, 0.1f, 0.9f);
            c[i].half_extents[2] = uni(&s, 0.1f, 0.9f);
            for (int a = 0; a < 3; a++) c[i].translation[a] = uni(&s, -4, 4);
            rquat(&s, c[i].rotation);
            if (i % 7 == 0) { c[i].rotation[0] = c[i].rotation[1] = c[i].rotation[2] = 0; c[i].rotation[3] = 1; } /* axis aligned */
        }
        uint64_t nb = 0;
        build(c, N, NULL, 0, &nb);
        uint8_t *blob = (uint8_t *)malloc(nb);
        build(c, N, blob, nb, &nb);
        const uint32_t leaf_off = ((const uint32_t *)blob)[2];
        const float *leaf = (const float *)(blob + leaf_off);
        float boxes[6 * N];
        for (int i = 0; i < N; i++) collider_cull_box(&c[i], boxes + 6 * i);
        fw_collision_settings f;
        memset(&f, 0, sizeof f);
        f.filter_mask = 0xFFFFFFFFu;
        for (long r = 0; r < rays; r++) {
            const int t = (int)(rnd(&s) % N);
            q4 rot = {c[t].rotation[0], c[t].rotation[1], c[t].rotation[2], c[t].rotation[3]};
            v3 tr = v3_make(c[t].translation[0], c[t].translation[1], c[t].translation[2]);
            const float he0 = c[t].half_extents[0], he1 = c[t].half_extents[1];
            v3 ol = v3_make(uni(&s, -2, 2), uni(&s, -2, 2), uni(&s, -2, 2));
            const int mode = (int)(rnd(&s) % 8);
            v3 dl;
            if (mode == 0) dl = v3_make(uni(&s, -1, 1), uni(&s, -1, 1), uni(&s, -1, 1));
            else if (mode == 1) dl = v3_sub(v3_make(uni(&s, -0.3f, 0.3f), uni(&s, -0.3f, 0.3f), uni(&s, -0.3f, 0.3f)), ol); /* aimed */
            else if (mode == 2) dl = v3_make(0, rnd(&s) & 1 ? 1.0f : -1.0f, 0);                                      /* along the axis */
            else if (mode == 3 || mode == 4) {                                                                         /* along a cone's slant */
                const float slope = he0 / (2.0f * he1), phi = uni(&s, 0, 6.2831853f), sg = rnd(&s) & 1 ? 1.0f : -1.0f;
                dl = v3_make(sg * slope * cosf(phi), -sg, sg * slope * sinf(phi));
            } else if (mode == 5) { dl = v3_make(rnd(&s) & 1 ? 1.0f : -1.0f, 0, 0); }                                /* along a face / tangent plane */
            else if (mode == 6) { dl = v3_make(uni(&s, -1, 1), 0, uni(&s, -1, 1)); }                                  /* perpendicular to the axis */
            else { /* tangent to the ball of radius he0 around the local origin through ol */
                v3 a = v3_make(uni(&s, -1, 1), uni(&s, -1, 1), uni(&s, -1, 1));
                dl = v3_make(ol.y * a.z - ol.z * a.y, ol.z * a.x - ol.x * a.z, ol.x * a.y - ol.y * a.x);
            }
            if (rnd(&s) % 3 == 0) { dl.x += uni(&s, -1e-4f, 1e-4f); dl.y += uni(&s, -1e-4f, 1e-4f); dl.z += uni(&s, -1e-4f, 1e-4f); }
            float len = v3_length(dl);
            if (!(len > 1e-6f)) continue;
            dl = v3_div(dl, len);
            if (rnd(&s) % 5 == 0) { /* origin on (about) the surface: scale ol onto the bounding ball / a face */
                const float l = v3_length(ol);
                if (l > 1e-3f) ol = v3_mul(ol, he0 / l);
            }
            v3 o = v3_add(tr, q_mul_v3(rot, ol)), d = q_mul_v3(rot, dl);
            const float mds[5] = {0.02f, 0.1f, 0.4f, 2.0f, 30.0f};
            const float md = mds[rnd(&s) % 5];
            float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, dist = 0, nrm[3], dist2 = 0, nrm2[3];
            uint32_t idx = 0, idx2 = 0;
            total++;
            const int a = cast_ray_impl(c, N, NULL, &f, oo, dd, md, &dist, nrm, &idx);
            const int b = cast_ray_impl(c, N, boxes, &f, oo, dd, md, &dist2, nrm2, &idx2);
            if (a != b || (a && (idx != idx2 || memcmp(&dist, &dist2, 4) || memcmp(nrm, nrm2, 12)))) {
                bad_cull++;
                if (bad_cull < 4) fprintf(stderr, "CULL scene %d kind %u o %.9g %.9g %.9g d %.9g %.9g %.9g md %g -> %d %u %g | %d %u %g\n", sc, c[a ? idx : idx2].kind, oo[0], oo[1], oo[2], dd[0], dd[1], dd[2], md, a, idx, dist, b, idx2, dist2);
            }
            if (!a) continue;
            hits++;
            int disjoint = 0;
            for (int k = 0; k < 3; k++) {
                const float e = oo[k] + dd[k] * md;
                const float lo = fminf(oo[k], e), hi = fmaxf(oo[k], e);
                if (hi < leaf[8 * idx + k] || lo > leaf[8 * idx + 4 + k]) disjoint = 1;
            }
            if (disjoint) {
                bad_box++;
                if (bad_box < 4) fprintf(stderr, "BOX scene %d collider %u kind %u o %.9g %.9g %.9g d %.9g %.9g %.9g md %g toi %g\n", sc, idx, c[idx].kind, oo[0], oo[1], oo[2], dd[0], dd[1], dd[2], md, dist);
            }
            if (!(dist >= 0.0f) || dist > md) far_hits++;
        }
        free(blob);
    }
    printf("rays %llu hits %llu | hit outside the collider's broad-phase box: %llu | brute force != culled helper: %llu | toi outside [0, max]: %llu\n",
           total, hits, bad_box, bad_cull, far_hits);
    return (bad_box || bad_cull || far_hits) ? 1 : 0;
}
