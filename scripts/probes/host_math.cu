// One-off campaign (CPU): the rest of the kernels' math header (fw_math.cuh) compiled for the host as a one-lane
// warp against the oracle's exports, bit for bit: particle_collision (up to 4 bounces: restitution, friction,
// destroy, the distance-0 push-out), sample_curve / sample_gradient (constant / even / uneven up to FW_MAX_KNOTS
// knots; t below 0, above 1, NaN, exactly on knots), Philox4x32-10, the glam helpers (quaternion from_scaled_axis / mul / from_rotation_arc / rotate, normalize_or_zero,
// project_onto, reject_from; parallel, antiparallel, tiny, zero and huge vectors).
//   nvcc -O2 -std=c++17 -Xcompiler -ffp-contract=off,-fno-fast-math,-fopenmp -Iinclude -Ibevy_firework_b200/csrc -Ioracle \
//        scripts/probes/host_math.cu -o scripts/probes/host_math -ldl -lgomp
//   scripts/probes/host_math bevy_firework_b200/libfirework_b200.so oracle/libfw_oracle.so <scenes> <cases per scene>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
#ifdef FW_HAVE_SGH
// the static update kernel's gradient sampler with a carried knot interval: its text, cut out of fw_kernels.cu by the
// caller (from its signature to the next template) into sgh.inc
namespace fw {
#include "sgh.inc"
}
#endif

typedef int (*build_fn)(const fw_collider *, uint32_t, void *, uint64_t, uint64_t *);
typedef void (*pc_fn)(const fw_collider *, uint32_t, const fw_collision_settings *, float *, float *, float, uint32_t *);
typedef float (*curve_fn)(const fw_curve_f32 *, float);
typedef void (*grad_fn)(const fw_gradient *, float, float *);
typedef void (*philox_fn)(const uint32_t *, const uint32_t *, uint32_t *);
typedef void (*qsa_fn)(const float *, float *);
typedef void (*qmul_fn)(const float *, const float *, float *);
typedef void (*v3v3_fn)(const float *, const float *, float *);
typedef void (*v3_fn)(const float *, float *);
static inline uint64_t rnd(uint64_t *s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return *s = x; }
static inline float uni(uint64_t *s, float a, float b) { return a + (b - a) * (float)((rnd(s) >> 40) * (1.0 / 16777216.0)); }
static inline bool same(const float *a, const float *b, int n) {
    for (int i = 0; i < n; i++)
        if (!(a[i] == b[i] || (a[i] != a[i] && b[i] != b[i]))) return false;
    return true;
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    void *h = dlopen(argv[1], RTLD_NOW), *ho = dlopen(argv[2], RTLD_NOW);
    if (!h || !ho) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    build_fn build = (build_fn)dlsym(h, "fw_host_build_broadphase");
    pc_fn o_pc = (pc_fn)dlsym(ho, "fwo_particle_collision");
    curve_fn o_curve = (curve_fn)dlsym(ho, "fwo_sample_curve");
    grad_fn o_grad = (grad_fn)dlsym(ho, "fwo_sample_gradient");
    philox_fn o_philox = (philox_fn)dlsym(ho, "fwo_philox4x32_10");
    qsa_fn o_qsa = (qsa_fn)dlsym(ho, "fwo_quat_from_scaled_axis");
    qmul_fn o_qmul = (qmul_fn)dlsym(ho, "fwo_quat_mul");
    v3v3_fn o_arc = (v3v3_fn)dlsym(ho, "fwo_quat_from_rotation_arc"), o_qrot = (v3v3_fn)dlsym(ho, "fwo_quat_mul_vec3");
    v3v3_fn o_proj = (v3v3_fn)dlsym(ho, "fwo_vec3_project_onto"), o_rej = (v3v3_fn)dlsym(ho, "fwo_vec3_reject_from");
    v3_fn o_norm = (v3_fn)dlsym(ho, "fwo_vec3_normalize_or_zero");
    const int scenes = atoi(argv[3]);
    const long cases = atol(argv[4]);
    unsigned long long n_pc = 0, n_hit = 0, bad_pc = 0, n_curve = 0, bad_curve = 0, n_grad = 0, bad_grad = 0, n_misc = 0, bad_misc = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : n_pc, n_hit, bad_pc, n_curve, bad_curve, n_grad, bad_grad, n_misc, bad_misc)
    for (int sc = 0; sc < scenes; sc++) {
        uint64_t s = 0xA0761D6478BD642Full * (uint64_t)(sc + 1);
        // ---- particle_collision in a dense scene
        const int N = 8 + (int)(rnd(&s) % 56);
        fw_collider *c = (fw_collider *)calloc(N, sizeof(fw_collider));
        const float region = uni(&s, 1.0f, 4.0f);
        for (int i = 0; i < N; i++) {
            c[i].kind = (uint32_t)(rnd(&s) % 5);
            c[i].layers = 1u + (uint32_t)(rnd(&s) % 3);
            c[i].key = 100u + (uint32_t)i;
            for (int a = 0; a < 3; a++) { c[i].half_extents[a] = uni(&s, 0.1f, 0.9f); c[i].translation[a] = uni(&s, -region, region); }
            float n = 0, q[4];
            do { n = 0; for (int k = 0; k < 4; k++) { q[k] = uni(&s, -1, 1); n += q[k] * q[k]; } } while (n < 1e-3f || n > 1.0f);
            n = sqrtf(n);
            for (int k = 0; k < 4; k++) c[i].rotation[k] = q[k] / n;
        }
        c[0].kind = FW_COLLIDER_CUBOID; // floor
        c[0].half_extents[0] = 30; c[0].half_extents[1] = 0.5f; c[0].half_extents[2] = 30;
        c[0].translation[0] = 0; c[0].translation[1] = -region - 0.5f; c[0].translation[2] = 0;
        c[0].rotation[0] = c[0].rotation[1] = c[0].rotation[2] = 0; c[0].rotation[3] = 1;
        uint64_t nb = 0;
        build(c, N, NULL, 0, &nb);
        uint8_t *blob = (uint8_t *)malloc(nb);
        build(c, N, blob, nb, &nb);
        uint32_t *queue = (uint32_t *)malloc(sizeof(uint32_t) * fw::kCandQueue * fw::kUpdateThreads);
        for (long r = 0; r < cases; r++) {
            fw_collision_settings cs;
            memset(&cs, 0, sizeof cs);
            const uint32_t masks[4] = {0xFFFFFFFFu, 1u, 2u, 3u};
            cs.enabled = 1;
            cs.filter_mask = masks[rnd(&s) % 4];
            cs.restitution = uni(&s, 0.0f, 1.0f);
            cs.friction = uni(&s, 0.0f, 0.8f);
            cs.destroy_on_collision = (rnd(&s) % 4) == 0;
            cs.n_excluded = (uint32_t)(rnd(&s) % 3);
            for (uint32_t x = 0; x < cs.n_excluded; x++) cs.excluded_keys[x] = 100u + (uint32_t)(rnd(&s) % N);
            float p[3], v[3];
            for (int a = 0; a < 3; a++) { p[a] = uni(&s, -region - 0.5f, region + 0.5f); v[a] = uni(&s, -12, 12); }
            if (rnd(&s) % 6 == 0) p[1] = -region + uni(&s, 0.0f, 0.01f);                 // resting on the floor
            if (rnd(&s) % 6 == 0) { v[0] *= 8; v[1] *= 8; v[2] *= 8; }                   // long segments: the BVH path
            if (rnd(&s) % 9 == 0) { v[0] = 0; v[2] = 0; }                                 // straight down / up
            if (rnd(&s) % 40 == 0) { v[0] = v[1] = v[2] = 0; }                            // at rest
            const float dts[5] = {1.0f / 60.0f, 1.0f / 144.0f, 1.0f / 30.0f, 0.1f, 0.0f};
            const float dt = dts[rnd(&s) % 5];
            float po[3] = {p[0], p[1], p[2]}, vo[3] = {v[0], v[1], v[2]};
            uint32_t sd_o = 0;
            o_pc(c, (uint32_t)N, &cs, po, vo, dt, &sd_o);
            fw::V3 pk = fw::v3(p[0], p[1], p[2]), vk = fw::v3(v[0], v[1], v[2]);
            bool sd_k = false;
            fw::particle_collision<true>(c, blob, cs, true, pk, vk, dt, queue, sd_k);
            const float pk3[3] = {pk.x, pk.y, pk.z}, vk3[3] = {vk.x, vk.y, vk.z};
            n_pc++;
            n_hit += (vo[0] != v[0] || vo[1] != v[1] || vo[2] != v[2] || sd_o) ? 1 : 0;
            if (!same(po, pk3, 3) || !same(vo, vk3, 3) || (sd_o != 0) != sd_k) {
                bad_pc++;
                if (bad_pc < 5)
                    fprintf(stderr, "COLLISION scene %d p %.9g %.9g %.9g v %.9g %.9g %.9g dt %g -> oracle p %.9g %.9g %.9g v %.9g %.9g %.9g d %u | kernel p %.9g %.9g %.9g v %.9g %.9g %.9g d %d\n",
                            sc, p[0], p[1], p[2], v[0], v[1], v[2], dt, po[0], po[1], po[2], vo[0], vo[1], vo[2], sd_o, pk.x, pk.y, pk.z, vk.x, vk.y, vk.z, (int)sd_k);
            }
        }
        free(queue);
        free(blob);
        free(c);
        // ---- curves and gradients
        for (long r = 0; r < cases / 8 + 1; r++) {
            fw_curve_f32 cv;
            fw_gradient g;
            memset(&cv, 0, sizeof cv);
            memset(&g, 0, sizeof g);
            fw::DevCurve dc;
            fw::DevGradient dg;
            memset(&dc, 0, sizeof dc);
            memset(&dg, 0, sizeof dg);
            cv.kind = g.kind = (uint32_t)(rnd(&s) % 3);
            const uint32_t n = cv.kind == FW_CURVE_CONSTANT ? 1u : 2u + (uint32_t)(rnd(&s) % (FW_MAX_KNOTS - 1));
            cv.n = g.n = n;
            float t0 = 0.0f;
            for (uint32_t i = 0; i < n; i++) {
                t0 = i == 0 ? uni(&s, -0.2f, 0.3f) : t0 + uni(&s, 1e-4f, 0.2f);
                cv.times[i] = g.times[i] = t0;
                cv.values[i] = uni(&s, -2, 3);
                for (int k = 0; k < 4; k++) g.colors[i][k] = uni(&s, 0, 4);
            }
            dc.kind = cv.kind; dc.n = n; dg.kind = g.kind; dg.n = n;
            for (uint32_t i = 0; i < n; i++) {
                dc.times[i] = cv.times[i]; dc.values[i] = cv.values[i];
                dg.times[i] = g.times[i]; dg.colors[i] = make_float4(g.colors[i][0], g.colors[i][1], g.colors[i][2], g.colors[i][3]);
            }
            for (int q = 0; q < 24; q++) {
                float t;
                const int m = (int)(rnd(&s) % 8);
                if (m == 0) t = cv.times[rnd(&s) % n];                                     // exactly on a knot
                else if (m == 1) t = (float)(rnd(&s) % n) / (float)(n > 1 ? n - 1 : 1);   // exactly on an even sample
                else if (m == 2) t = uni(&s, -1, 0);
                else if (m == 3) t = uni(&s, 1, 3);
                else if (m == 4) t = NAN;
                else t = uni(&s, 0, 1);
                const float a = o_curve(&cv, t), b = fw::sample_curve(dc, t);
                n_curve++;
                if (!same(&a, &b, 1)) { bad_curve++; if (bad_curve < 4) fprintf(stderr, "CURVE kind %u n %u t %.9g oracle %.9g kernel %.9g\n", cv.kind, n, t, a, b); }
                float ga[4];
                o_grad(&g, t, ga);
                const float4 gb = fw::sample_gradient(dg, t);
                const float gb4[4] = {gb.x, gb.y, gb.z, gb.w};
                n_grad++;
                if (!same(ga, gb4, 4)) { bad_grad++; if (bad_grad < 4) fprintf(stderr, "GRADIENT kind %u n %u t %.9g\n", g.kind, n, t); }
            }
#ifdef FW_HAVE_SGH
            {   // a run of ages as a warp lane sees them along a ring (slowly varying, with jumps, knots, tails, NaN),
                // the knot interval carried from call to call
                uint32_t hint = 0;
                float t = uni(&s, 0, 1);
                for (int q = 0; q < 48; q++) {
                    const int m = (int)(rnd(&s) % 10);
                    if (m == 0) t = g.times[rnd(&s) % n];
                    else if (m == 1) t = uni(&s, -0.5f, 1.5f);
                    else if (m == 2) t = NAN;
                    else if (m == 3) t = uni(&s, 0, 1);
                    else t = t + uni(&s, -0.02f, 0.03f);
                    float ga[4];
                    o_grad(&g, t, ga);
                    const float4 gb = fw::sample_gradient_hint(dg, t, hint);
                    const float gb4[4] = {gb.x, gb.y, gb.z, gb.w};
                    n_grad++;
                    if (!same(ga, gb4, 4)) { bad_grad++; if (bad_grad < 4) fprintf(stderr, "GRADIENT (hint) kind %u n %u t %.9g hint %u\n", g.kind, n, t, hint); }
                    if (t != t) t = uni(&s, 0, 1);
                }
            }
#endif
        }
        // ---- Philox, quaternions
        for (long r = 0; r < cases / 4 + 1; r++) {
            uint32_t ctr[4], key[2], out[4];
            for (int k = 0; k < 4; k++) ctr[k] = (uint32_t)rnd(&s);
            key[0] = (uint32_t)rnd(&s); key[1] = (uint32_t)rnd(&s);
            o_philox(ctr, key, out);
            const uint4 pk = fw::philox4x32_10(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]));
            n_misc++;
            if (pk.x != out[0] || pk.y != out[1] || pk.z != out[2] || pk.w != out[3]) bad_misc++;
            float v[3] = {uni(&s, -1, 1), uni(&s, -1, 1), uni(&s, -1, 1)}, qa[4], qb[4], qo[4];
            if (rnd(&s) % 10 == 0) v[0] = v[1] = v[2] = 0;
            if (rnd(&s) % 10 == 0) { v[0] *= 1e-4f; v[1] *= 1e-4f; v[2] *= 1e-4f; }
            o_qsa(v, qa);
            const fw::Q4 qk = fw::q_from_scaled_axis(fw::v3(v[0], v[1], v[2]));
            const float qk4[4] = {qk.x, qk.y, qk.z, qk.w};
            n_misc++;
            if (!same(qa, qk4, 4)) { bad_misc++; if (bad_misc < 4) fprintf(stderr, "QUAT from_scaled_axis %.9g %.9g %.9g\n", v[0], v[1], v[2]); }
            for (int k = 0; k < 4; k++) qb[k] = uni(&s, -1, 1);
            o_qmul(qa, qb, qo);
            const fw::Q4 qm = fw::qmul(fw::Q4{qa[0], qa[1], qa[2], qa[3]}, fw::Q4{qb[0], qb[1], qb[2], qb[3]});
            const float qm4[4] = {qm.x, qm.y, qm.z, qm.w};
            n_misc++;
            if (!same(qo, qm4, 4)) bad_misc++;
            // from_rotation_arc: random, parallel, antiparallel (the any_orthonormal branch) and almost so
            float fa[3] = {uni(&s, -1, 1), uni(&s, -1, 1), uni(&s, -1, 1)}, fb[3] = {uni(&s, -1, 1), uni(&s, -1, 1), uni(&s, -1, 1)};
            float la = sqrtf(fa[0] * fa[0] + fa[1] * fa[1] + fa[2] * fa[2]), lb = sqrtf(fb[0] * fb[0] + fb[1] * fb[1] + fb[2] * fb[2]);
            if (la > 1e-3f && lb > 1e-3f) {
                for (int k = 0; k < 3; k++) { fa[k] /= la; fb[k] /= lb; }
                const int m = (int)(rnd(&s) % 8);
                if (m == 0) { fa[0] = 0; fa[1] = 1; fa[2] = 0; }                                       // from = Y, as the spawn code calls it
                if (m == 1) for (int k = 0; k < 3; k++) fb[k] = fa[k];                                 // parallel
                if (m == 2) for (int k = 0; k < 3; k++) fb[k] = -fa[k];                                // antiparallel
                if (m == 3) for (int k = 0; k < 3; k++) fb[k] = -fa[k] + uni(&s, -3e-4f, 3e-4f);       // almost antiparallel
                if (m == 4) for (int k = 0; k < 3; k++) fb[k] = fa[k] + uni(&s, -3e-4f, 3e-4f);        // almost parallel
                float ao[4];
                o_arc(fa, fb, ao);
                const fw::Q4 ak = fw::q_from_rotation_arc(fw::v3(fa[0], fa[1], fa[2]), fw::v3(fb[0], fb[1], fb[2]));
                const float ak4[4] = {ak.x, ak.y, ak.z, ak.w};
                n_misc++;
                if (!same(ao, ak4, 4)) { bad_misc++; if (bad_misc < 4) fprintf(stderr, "ARC %.9g %.9g %.9g -> %.9g %.9g %.9g\n", fa[0], fa[1], fa[2], fb[0], fb[1], fb[2]); }
                float ro[3];
                o_qrot(ao, v, ro);
                const fw::V3 rk = fw::qrot(ak, fw::v3(v[0], v[1], v[2]));
                const float rk3[3] = {rk.x, rk.y, rk.z};
                n_misc++;
                if (!same(ro, rk3, 3)) bad_misc++;
            }
            // normalize_or_zero / project_onto / reject_from, tiny and zero vectors included
            float w[3] = {uni(&s, -3, 3), uni(&s, -3, 3), uni(&s, -3, 3)};
            const int mw = (int)(rnd(&s) % 6);
            if (mw == 0) { w[0] *= 1e-20f; w[1] *= 1e-20f; w[2] *= 1e-20f; }
            if (mw == 1) { w[0] = w[1] = w[2] = 0; }
            if (mw == 2) { w[0] *= 1e18f; w[1] *= 1e18f; w[2] *= 1e18f; }
            float no[3], po[3], jo[3];
            o_norm(w, no);
            o_proj(v, w, po);
            o_rej(v, w, jo);
            const fw::V3 nk = fw::normalize_or_zero(fw::v3(w[0], w[1], w[2]));
            const fw::V3 pjk = fw::project_onto(fw::v3(v[0], v[1], v[2]), fw::v3(w[0], w[1], w[2]));
            const fw::V3 jk = fw::reject_from(fw::v3(v[0], v[1], v[2]), fw::v3(w[0], w[1], w[2]));
            const float nk3[3] = {nk.x, nk.y, nk.z}, pk3[3] = {pjk.x, pjk.y, pjk.z}, jk3[3] = {jk.x, jk.y, jk.z};
            n_misc += 3;
            if (!same(no, nk3, 3)) { bad_misc++; if (bad_misc < 4) fprintf(stderr, "NORMALIZE %.9g %.9g %.9g\n", w[0], w[1], w[2]); }
            if (!same(po, pk3, 3)) { bad_misc++; if (bad_misc < 4) fprintf(stderr, "PROJECT onto %.9g %.9g %.9g\n", w[0], w[1], w[2]); }
            if (!same(jo, jk3, 3)) bad_misc++;
        }
    }
    printf("particle_collision %llu (changed or destroyed: %llu) mismatches %llu | sample_curve %llu mismatches %llu | sample_gradient %llu mismatches %llu | philox / quaternions %llu mismatches %llu\n",
           n_pc, n_hit, bad_pc, n_curve, bad_curve, n_grad, bad_grad, n_misc, bad_misc);
    return (bad_pc || bad_curve || bad_grad || bad_misc) ? 1 : 0;
}
