// Philox4x32-10 counter-based generator (Salmon et al., SC'11) + Box-Muller.
// Replaces tf.random.normal in the update loop (models/dgp.py:210,212,240).  The stream is
// a pure function of (seed, chain, element/4, step, stream id) so that
//   - chains sharded over GPUs draw exactly what they would draw on one GPU, and
//   - data-parallel replicas draw identical noise without a broadcast.
#pragma once
#include <stdint.h>

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__host__ __device__ __forceinline__ uint64_t philox_key(uint64_t seed, uint64_t chain, uint32_t stream) {
    return seed + chain * 0x9E3779B97F4A7C15ull + (uint64_t)stream * 0xD1B54A32D192ED03ull;
}

#ifdef __CUDACC__
// two uniforms -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float* n0, float* n1) {
    const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 2.3283064365386963e-10f);   // (0, 1]
    const float u2 = (float)b * 2.3283064365386963e-10f;                                   // [0, 1]
    // MUFU-based log / sin / cos: abs error ~1e-6 on these ranges, far below the Monte-Carlo noise they feed
    const float r = sqrtf(-2.f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * (u2 - 0.5f), &s, &c);      // angle in [-pi, pi]; the half-turn shift
    s = -s; c = -c;                                            // is undone by the sign flip
    *n0 = r * c;
    *n1 = r * s;
}

// four N(0,1) draws for the 128-bit lane `idx4` of `chain` at `step`; stream 0 = eps,
// stream 1 = momentum resample.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t chain, uint64_t idx4,
                                                 uint64_t step, uint32_t stream) {
    const uint64_t key = philox_key(seed, chain, stream);
    uint32_t c[4] = {(uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)key, (uint32_t)(key >> 32));
    float4 o;
    box_muller(c[0], c[1], &o.x, &o.y);
    box_muller(c[2], c[3], &o.z, &o.w);
    return o;
}
#endif
