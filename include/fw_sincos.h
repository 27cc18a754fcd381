/*
 * fw_sincos.h -- the sine / cosine of the particle path, as ONE definition compiled into both
 * sides of every parity comparison.
 *
 * Where the reference needs a sine or cosine it calls glam, which calls the platform libm
 * (f32::sin_cos): Quat::from_scaled_axis for the rotation step (reference src/core.rs:645-647),
 * Quat::from_rotation_y / PitchYaw::to_unit_vec in EmissionShape::generate_point
 * (src/emission_shape.rs:26-36), RandVec3::generate (bevy_utilitarian). A platform libm is not
 * a specification: glibc's, CUDA's and musl's sinf differ in the last ulp, and one ulp at spawn
 * is enough to flip a grazing ray cast a hundred frames later. So the library defines the
 * function it uses -- in terms of IEEE-754 operations only, which round identically on a CPU and
 * on sm_100a -- and the CUDA kernels (csrc/fw_math.cuh, built -fmad=false) and the CPU oracle
 * (oracle/fw_oracle.c, built -ffp-contract=off) both compile THIS file. Rotation, spawn and every
 * trajectory that follows are then bit-identical between the two.
 *
 *   fw_sincosf(x, &s, &c):
 *     |x| <= pi/4  r = x;
 *     |x| < 2^20   k = rint(x * 2/pi) and r = x - k*pi/2 in double (pi/2 = P1 + P2, P1 holds 33
 *                  bits, so k*P1 is exact);
 *     otherwise    Payne-Hanek: the 24-bit significand times 192 bits of 2/pi in integer
 *                  arithmetic, the two bits above the binary point are the quadrant, the next 64
 *                  the fraction -- exact for every finite float;
 *     then         sin r and cos r for |r| <= pi/4 as Taylor polynomials in double (degree 13 / 14,
 *                  truncation < 2^-37), rounded once to float, and the quadrant's symmetry.
 *   Result: within 0.5 ulp + 2^-12 ulp of the true value for every finite float; measured over all
 *   2^32 bit patterns against an 80-bit libm (scripts/sincos_exhaustive.c, profiles/r2/
 *   sincos_exhaustive.txt): 52 of 8.6e9 results are not the correctly rounded float, none is off by
 *   more than 0.5001 ulp. NaN for NaN / infinities; sin(-0) = -0.
 *
 * Nothing here depends on a rounding mode other than round-to-nearest-even or on a fused
 * multiply-add; do not build it with contraction enabled.
 */
#ifndef FW_SINCOS_H
#define FW_SINCOS_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FW_TRIG_FN __host__ __device__ inline
#define FW_TRIG_SLOW __host__ __device__ __noinline__
#else
#define FW_TRIG_FN static inline
#define FW_TRIG_SLOW static
#endif

FW_TRIG_FN uint32_t fw_trig_f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, sizeof(u));
    return u;
#endif
}

/* Payne-Hanek reduction for |x| >= 2^20 (finite): *r = x - q * pi/2 with |r| <= pi/4, returns q mod 4.
 * |x| = m * 2^(e-23), m the 24-bit significand. With T = the first 192 bits of 2/pi as an integer,
 * |x| * 2/pi = m*T * 2^(e-23-192); bits [130-s, 194-s) of the 216-bit product m*T (s = e - 23) are
 * two integer bits and 62 fraction bits of that number. */
FW_TRIG_SLOW int fw_trig_reduce_large(float x, double *r) {
    const uint32_t two_over_pi[6] = {0x3c439041u, 0xdb629599u, 0xf534ddc0u, 0xfc2757d1u, 0x4e441529u, 0xa2f9836eu}; /* least significant first */
    const uint32_t ix = fw_trig_f32_bits(x);
    const int s = (int)((ix >> 23) & 0xffu) - 127 - 23;
    const uint64_t m = (uint64_t)((ix & 0x007fffffu) | 0x00800000u);
    uint32_t p[7];
    uint64_t carry = 0;
    for (int i = 0; i < 6; i++) {
        carry += m * (uint64_t)two_over_pi[i];
        p[i] = (uint32_t)carry;
        carry >>= 32;
    }
    p[6] = (uint32_t)carry;
    const int b = 130 - s; /* s = -3 .. 104: b = 26 .. 133, word 0 .. 4 */
    const int word = b >> 5, bit = b & 31;
    /* 96 bits starting at limb `word`, selected without indexing by a run-time value (keeps the
     * limbs in registers on the device) */
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    for (int i = 0; i < 5; i++)
        if (word == i) {
            w0 = p[i];
            w1 = p[i + 1];
            w2 = p[i + 2];
        }
    uint64_t v = ((uint64_t)w1 << 32) | (uint64_t)w0;
    if (bit) v = (v >> bit) | ((uint64_t)w2 << (64 - bit));
    /* v = (|x| * 2/pi mod 4) in Q2.62; round the quadrant to nearest, keep the signed fraction */
    int q = (int)(v >> 62) + (int)((v >> 61) & 1u);
    const int64_t frac = (int64_t)(v << 2); /* Q0.64 in [-1/2, 1/2) */
    double rr = (double)frac * 0x1.921fb54442d18p-64; /* * pi/2 * 2^-64 */
    if (ix >> 31) {
        rr = -rr;
        q = -q;
    }
    *r = rr;
    return q & 3;
}

FW_TRIG_FN void fw_sincosf(float x, float *sin_out, float *cos_out) {
    const uint32_t ax = fw_trig_f32_bits(x) & 0x7fffffffu;
    if (ax >= 0x7f800000u) { /* NaN, +-inf */
        const float n = x - x;
        *sin_out = n;
        *cos_out = n;
        return;
    }
    double r;
    int q;
    if (ax <= 0x3f490fdbu) { /* |x| <= fl32(pi/4): nothing to reduce (and sin(-0) stays -0) */
        r = (double)x;
        q = 0;
    } else if (ax < 0x49800000u) { /* |x| < 2^20 */
        const double xd = (double)x;
        const double k = rint(xd * 0x1.45f306dc9c883p-1); /* 2/pi */
        r = (xd - k * 0x1.921fb544p+0) - k * 0x1.0b4611a626331p-34;
        q = (int)k & 3;
    } else {
        q = fw_trig_reduce_large(x, &r);
    }
    const double z = r * r;
    double ps = 0x1.6124613a86d09p-33;              /* 1/13! */
    ps = -0x1.ae64567f544e4p-26 + z * ps;           /* 1/11! */
    ps = 0x1.71de3a556c734p-19 + z * ps;            /* 1/9!  */
    ps = -0x1.a01a01a01a01ap-13 + z * ps;           /* 1/7!  */
    ps = 0x1.1111111111111p-7 + z * ps;             /* 1/5!  */
    ps = -0x1.5555555555555p-3 + z * ps;            /* 1/3!  */
    const double sn = r * (1.0 + z * ps); /* (a product keeps the sign of a zero argument) */
    double pc = -0x1.93974a8c07c9dp-37;             /* 1/14! */
    pc = 0x1.1eed8eff8d898p-29 + z * pc;            /* 1/12! */
    pc = -0x1.27e4fb7789f5cp-22 + z * pc;           /* 1/10! */
    pc = 0x1.a01a01a01a01ap-16 + z * pc;            /* 1/8!  */
    pc = -0x1.6c16c16c16c17p-10 + z * pc;           /* 1/6!  */
    pc = 0x1.5555555555555p-5 + z * pc;             /* 1/4!  */
    pc = -0.5 + z * pc;                             /* 1/2!  */
    const double cs = 1.0 + z * pc;
    const float s = (float)sn, c = (float)cs;
    switch (q) {
    case 0: *sin_out = s; *cos_out = c; break;
    case 1: *sin_out = c; *cos_out = -s; break;
    case 2: *sin_out = -s; *cos_out = -c; break;
    default: *sin_out = -c; *cos_out = s; break;
    }
}

#endif /* FW_SINCOS_H */
