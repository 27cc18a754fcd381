// One-off campaign (CPU): the KERNELS' cast_ray (fw_math.cuh: grid / BVH enumeration, candidate queue, exact
// tests, filter mask, exclusions, tie-break) compiled for the host as a one-lane warp, against the oracle's
// brute-force loop -- hit or not, distance bits, normal bits -- on hundreds of millions of rays.
//   nvcc -O2 -std=c++17 -Xcompiler -ffp-contract=off,-fno-fast-math,-fopenmp -Iinclude -Ibevy_firework_b200/csrc \
//        scripts/probes/host_cast_ray.cu -o scripts/probes/host_cast_ray -ldl -lgomp
//   scripts/probes/host_cast_ray bevy_firework_b200/libfirework_b200.so oracle/libfw_oracle.so <scenes> <rays per scene>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// ---- a one-lane warp on the host: everything declared __device__ below also exists for the host
#undef __device__
#define __device__ __location__(host) __location__(device)
static inline __host__ uint32_t emul_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#define __ldg(p) (*(p))
#define __float_as_uint(f) emul_f2u(f)
#define __any_sync(m, p) (p)
#define __reduce_max_sync(m, v) (v)
#define __syncwarp() ((void)0)
#define __umulhi(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#include "fw_math.cuh"

typedef int (*build_fn)(const fw_collider *, uint32_t, void *, uint64_t, uint64_t *);
typedef int (*oracle_fn)(const fw_collider *, uint32_t, uint32_t, const float *, const float *, float, float *, float *, uint32_t *);
static inline uint64_t rnd(uint64_t *s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return *s = x; }
static inline float uni(uint64_t *s, float a, float b) { return a + (b - a) * (float)((rnd(s) >> 40) * (1.0 / 16777216.0)); }

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    void *h = dlopen(argv[1], RTLD_NOW), *ho = dlopen(argv[2], RTLD_NOW);
    if (!h || !ho) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    build_fn build = (build_fn)dlsym(h, "fw_host_build_broadphase");
    oracle_fn oracle = (oracle_fn)dlsym(ho, "fwo_cast_ray");
    const int scenes = atoi(argv[3]);
    const long rays = atol(argv[4]);
    unsigned long long total = 0, hits = 0, bad = 0, grid_rays = 0, many = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total, hits, bad, grid_rays, many)
    for (int sc = 0; sc < scenes; sc++) {
        uint64_t s = 0xD1B54A32D192ED03ull * (uint64_t)(sc + 1);
        const int N = 8 + (int)(rnd(&s) % 120);
        fw_collider *c = (fw_collider *)calloc(N, sizeof(fw_collider));
        const float region = uni(&s, 1.5f, 8.0f); // dense scenes: many candidates per ray, several rounds of the queue
        for (int i = 0; i < N; i++) {
            c[i].kind = (uint32_t)(rnd(&s) % 5);
            c[i].layers = 1u + (uint32_t)(rnd(&s) % 3);
            c[i].key = 100u + (uint32_t)i;
            for (int a = 0; a < 3; a++) {
                c[i].half_extents[a] = uni(&s, 0.1f, 0.9f);
                c[i].translation[a] = uni(&s, -region, region);
            }
            float n = 0, q[4];
            do { n = 0; for (int k = 0; k < 4; k++) { q[k] = uni(&s, -1, 1); n += q[k] * q[k]; } } while (n < 1e-3f || n > 1.0f);
            n = sqrtf(n);
            for (int k = 0; k < 4; k++) c[i].rotation[k] = q[k] / n;
        }
        if (sc % 2 == 0) { // a ground slab: the "big" list
            c[0].kind = FW_COLLIDER_CUBOID;
            c[0].half_extents[0] = 40; c[0].half_extents[1] = 0.5f; c[0].half_extents[2] = 40;
            c[0].translation[0] = 0; c[0].translation[1] = -region - 0.5f; c[0].translation[2] = 0;
            c[0].rotation[0] = c[0].rotation[1] = c[0].rotation[2] = 0; c[0].rotation[3] = 1;
        }
        uint64_t nb = 0;
        build(c, N, NULL, 0, &nb);
        uint8_t *blob = (uint8_t *)malloc(nb);
        build(c, N, blob, nb, &nb);
        uint32_t *queue = (uint32_t *)malloc(sizeof(uint32_t) * fw::kCandQueue * fw::kUpdateThreads);
        for (long r = 0; r < rays; r++) {
            fw_collision_settings cs;
            memset(&cs, 0, sizeof cs);
            const uint32_t masks[4] = {0xFFFFFFFFu, 1u, 2u, 3u};
            cs.filter_mask = masks[rnd(&s) % 4];
            // (exclusions are compared separately below: the oracle export takes a mask only)
            float o[3], d[3];
            const int t = (int)(rnd(&s) % N);
            for (int a = 0; a < 3; a++) o[a] = (rnd(&s) % 3) ? c[t].translation[a] + uni(&s, -2, 2) : uni(&s, -region - 1, region + 1);
            float len = 0;
            for (int a = 0; a < 3; a++) { d[a] = (rnd(&s) % 4) ? uni(&s, -1, 1) : c[t].translation[a] - o[a]; len += d[a] * d[a]; }
            if (rnd(&s) % 16 == 0) { d[0] = 0; d[2] = 0; d[1] = d[1] < 0 ? -1.f : 1.f; len = 1; } // axis-parallel
            len = sqrtf(len);
            if (!(len > 1e-6f)) continue;
            for (int a = 0; a < 3; a++) d[a] /= len;
            const float mds[6] = {0.02f, 0.1f, 0.4f, 1.5f, 6.0f, 60.0f};
            const float md = mds[rnd(&s) % 6];
            float dist = 0, nrm[3] = {0, 0, 0};
            uint32_t idx = 0;
            const int a = oracle(c, (uint32_t)N, cs.filter_mask, o, d, md, &dist, nrm, &idx);
            float kd = 0;
            fw::V3 kn = fw::v3(0, 0, 0);
            const bool b = fw::cast_ray<true>(c, blob, cs, true, fw::v3(o[0], o[1], o[2]), fw::v3(d[0], d[1], d[2]), md, queue, kd, kn);
            total++;
            hits += a ? 1 : 0;
            const float kn3[3] = {kn.x, kn.y, kn.z};
            if ((a != 0) != b || (a && (memcmp(&dist, &kd, 4) || memcmp(nrm, kn3, 12)))) {
                bad++;
                if (bad < 6)
                    fprintf(stderr, "MISMATCH scene %d N %d mask %x o %.9g %.9g %.9g d %.9g %.9g %.9g md %g oracle %d idx %u kind %u toi %.9g | kernel %d toi %.9g\n",
                            sc, N, cs.filter_mask, o[0], o[1], o[2], d[0], d[1], d[2], md, a, idx, c[idx].kind, dist, (int)b, kd);
            }
            // with one exclusion: the kernel must then agree with the oracle run on the set without that collider
            if (a && (rnd(&s) % 8) == 0) {
                cs.n_excluded = 1;
                cs.excluded_keys[0] = c[idx].key;
                fw_collider *c2 = (fw_collider *)malloc(sizeof(fw_collider) * N);
                memcpy(c2, c, sizeof(fw_collider) * N);
                c2[idx].layers = 0; // invisible to every mask
                float dist2 = 0, nrm2[3] = {0, 0, 0};
                uint32_t idx2 = 0;
                const int a2 = oracle(c2, (uint32_t)N, cs.filter_mask, o, d, md, &dist2, nrm2, &idx2);
                const bool b2 = fw::cast_ray<true>(c, blob, cs, true, fw::v3(o[0], o[1], o[2]), fw::v3(d[0], d[1], d[2]), md, queue, kd, kn);
                const float k3[3] = {kn.x, kn.y, kn.z};
                if ((a2 != 0) != b2 || (a2 && (memcmp(&dist2, &kd, 4) || memcmp(nrm2, k3, 12)))) {
                    bad++;
                    if (bad < 6) fprintf(stderr, "MISMATCH (exclusion) scene %d idx %u -> oracle %d %.9g kernel %d %.9g\n", sc, idx, a2, dist2, (int)b2, kd);
                }
                free(c2);
                total++;
            }
        }
        free(queue);
        free(blob);
        free(c);
    }
    printf("rays %llu hits %llu | kernel cast_ray != oracle brute force: %llu\n", total, hits, bad);
    return bad ? 1 : 0;
}
