// Philox4x32-10 counter RNG, addressed by (transition, purpose, a, b) and keyed by the
// chain's 64-bit seed.  The NumPy twin used by the CPU oracle is oracle/rng.py; the two
// must stay bit-identical (tests/test_hostsim.py, tests/test_gpu_parity.py check it).
//
// Replaces NumPy's global legacy RandomState in the reference
// (pymc3/step_methods/hmc/nuts.py:30-33,177,290,375; quadpotential.py:200-203;
// hmc.py:26-27,136): the reference consumes one stream in program order, which cannot be
// reproduced by a batched kernel, so randomness is addressed by counter instead.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

enum : uint32_t {
    B2_PURPOSE_MOMENTUM = 0,
    B2_PURPOSE_DIRECTION = 1,
    B2_PURPOSE_MERGE = 2,
    B2_PURPOSE_TOP = 3,
    B2_PURPOSE_HMC_JITTER = 4,
    B2_PURPOSE_HMC_ACCEPT = 5,
};

B2_HD uint32_t b2_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

B2_HD void b2_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                            uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = b2_mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = b2_mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two 32-bit words -> double in [0,1) with 53 random bits
B2_HD double b2_u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) / 9007199254740992.0;
}

B2_HD double b2_uniform(uint32_t k0, uint32_t k1, uint32_t t, uint32_t purpose, uint32_t a, uint32_t b) {
    uint32_t r[4];
    b2_philox4x32_10(t, purpose, a, b, k0, k1, r);
    return b2_u53(r[0], r[1]);
}

// standard normal for momentum component i of transition t: Box-Muller on the pair i>>1
B2_HD double b2_normal(uint32_t k0, uint32_t k1, uint32_t t, uint32_t i) {
    uint32_t r[4];
    b2_philox4x32_10(t, B2_PURPOSE_MOMENTUM, i >> 1, 0u, k0, k1, r);
    double u1 = 1.0 - b2_u53(r[0], r[1]);      // (0, 1]
    double u2 = b2_u53(r[2], r[3]);
    double rad = sqrt(-2.0 * log(u1));
    double ang = 6.283185307179586476925286766559 * u2;
    return rad * ((i & 1u) ? sin(ang) : cos(ang));
}

// Same stream, element type T: the fp32 production build does the Box-Muller transcendentals in float
// (double log/sqrt/sincos per momentum component dominated the transition-ending path of the lock-step
// advance kernel); the uniforms are the same 53-bit values, so fp32 and fp64 builds draw the same
// normals to ~1e-7.
template <typename T> struct B2Normal;
template <> struct B2Normal<double> {
    B2_HD static double draw(uint32_t k0, uint32_t k1, uint32_t t, uint32_t i) { return b2_normal(k0, k1, t, i); }
    // both normals of pair j (components 2j and 2j+1) from ONE Philox block and ONE Box-Muller transform
    B2_HD static void draw2(uint32_t k0, uint32_t k1, uint32_t t, uint32_t j, double& n0, double& n1) {
        uint32_t r[4];
        b2_philox4x32_10(t, B2_PURPOSE_MOMENTUM, j, 0u, k0, k1, r);
        const double u1 = 1.0 - b2_u53(r[0], r[1]);
        const double u2 = b2_u53(r[2], r[3]);
        const double rad = sqrt(-2.0 * log(u1));
        const double ang = 6.283185307179586476925286766559 * u2;
        n0 = rad * cos(ang); n1 = rad * sin(ang);
    }
};
template <> struct B2Normal<float> {
    B2_HD static float draw(uint32_t k0, uint32_t k1, uint32_t t, uint32_t i) {
        uint32_t r[4];
        b2_philox4x32_10(t, B2_PURPOSE_MOMENTUM, i >> 1, 0u, k0, k1, r);
        const float u1 = (float)(1.0 - b2_u53(r[0], r[1]));     // (0, 1]
        const float u2 = (float)b2_u53(r[2], r[3]);
        const float rad = sqrtf(-2.0f * logf(u1 > 1e-37f ? u1 : 1e-37f));
        const float ang = 6.2831853071795864f * u2;
        return rad * ((i & 1u) ? sinf(ang) : cosf(ang));
    }
    B2_HD static void draw2(uint32_t k0, uint32_t k1, uint32_t t, uint32_t j, float& n0, float& n1) {
        uint32_t r[4];
        b2_philox4x32_10(t, B2_PURPOSE_MOMENTUM, j, 0u, k0, k1, r);
        const float u1 = (float)(1.0 - b2_u53(r[0], r[1]));     // (0, 1]
        const float u2 = (float)b2_u53(r[2], r[3]);
        const float rad = sqrtf(-2.0f * logf(u1 > 1e-37f ? u1 : 1e-37f));
        const float ang = 6.2831853071795864f * u2;
        n0 = rad * cosf(ang); n1 = rad * sinf(ang);
    }
};
