/* Exhaustive accuracy check of include/fw_sincos.h: every finite float against sinl / cosl
 * (x87 80-bit long double, 64-bit significand). Prints the worst error in ulps of the float
 * result and how many results are not the correctly rounded value.
 *   gcc -O2 -ffp-contract=off -pthread -o /tmp/sincos_exhaustive scripts/sincos_exhaustive.c -lm
 *   /tmp/sincos_exhaustive [threads] [stride]   (stride 1 = all 2^32 bit patterns) */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include "../include/fw_sincos.h"

typedef struct { uint32_t begin, end, stride; double worst_s, worst_c; float at_s, at_c; uint64_t n, misrounded_s, misrounded_c; } job;

static double ulp_err(float got, long double want) {
    float w = (float)want;
    if (got == w) { /* error relative to the ulp of the result */ }
    int e;
    frexpl(want, &e);
    long double ulp = ldexpl(1.0L, e - 24);
    if (fabsl(want) < 0x1p-126L) ulp = 0x1p-149L;
    return (double)(fabsl((long double)got - want) / ulp);
}
static void *run(void *arg) {
    job *j = (job *)arg;
    for (uint64_t u = j->begin; u < j->end; u += j->stride) {
        uint32_t bits = (uint32_t)u;
        float x;
        memcpy(&x, &bits, 4);
        if (!isfinite(x)) continue;
        float s, c;
        fw_sincosf(x, &s, &c);
        long double ws = sinl((long double)x), wc = cosl((long double)x);
        double es = ulp_err(s, ws), ec = ulp_err(c, wc);
        if (es > j->worst_s) { j->worst_s = es; j->at_s = x; }
        if (ec > j->worst_c) { j->worst_c = ec; j->at_c = x; }
        if (s != (float)ws) j->misrounded_s++;
        if (c != (float)wc) j->misrounded_c++;
        j->n++;
    }
    return NULL;
}
int main(int argc, char **argv) {
    int nt = argc > 1 ? atoi(argv[1]) : 8;
    uint32_t stride = argc > 2 ? (uint32_t)atoi(argv[2]) : 1;
    pthread_t th[64];
    job jobs[64];
    uint64_t span = (1ull << 32) / nt;
    for (int t = 0; t < nt; t++) {
        memset(&jobs[t], 0, sizeof(job));
        jobs[t].begin = (uint32_t)(span * t);
        jobs[t].end = t == nt - 1 ? 0xFFFFFFFFu : (uint32_t)(span * (t + 1));
        jobs[t].stride = stride;
        pthread_create(&th[t], NULL, run, &jobs[t]);
    }
    job tot;
    memset(&tot, 0, sizeof(tot));
    for (int t = 0; t < nt; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].worst_s > tot.worst_s) { tot.worst_s = jobs[t].worst_s; tot.at_s = jobs[t].at_s; }
        if (jobs[t].worst_c > tot.worst_c) { tot.worst_c = jobs[t].worst_c; tot.at_c = jobs[t].at_c; }
        tot.n += jobs[t].n; tot.misrounded_s += jobs[t].misrounded_s; tot.misrounded_c += jobs[t].misrounded_c;
    }
    printf("floats tested %llu\nsin: worst %.6f ulp at %a, not correctly rounded %llu\ncos: worst %.6f ulp at %a, not correctly rounded %llu\n",
           (unsigned long long)tot.n, tot.worst_s, tot.at_s, (unsigned long long)tot.misrounded_s, tot.worst_c, tot.at_c,
           (unsigned long long)tot.misrounded_c);
    return 0;
}
