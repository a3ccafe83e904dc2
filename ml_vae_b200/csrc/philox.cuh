// Philox4x32-10 and the eps stream of the reparameterisation kernels.
// Host restatement: oracle/philox_ref.py (the stream definition lives there).
#pragma once
#include <stdint.h>

namespace mlvae {

struct PhiloxKey {
    uint32_t k0[10], k1[10];   // per-round keys; built on the HOST and passed by value where the kernel is issue bound (the
                               // rounds then take them straight from the constant bank: a key derived from a `seed` parameter is
                               // re-derived in uniform registers on every loop iteration, 18 issue slots per Philox call)
    __host__ __device__ __forceinline__ explicit PhiloxKey(uint64_t seed) {
        uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            k0[r] = a; k1[r] = b;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
    }
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKey &key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.k0[r], lo1, hi0 ^ c.w ^ key.k1[r], lo0);
    }
    return c;
}

// Box-Muller on one (a, b) pair of uint32:
//   u1 = fl32(fl32(a) + 1) * 2^-32 in (0, 1],   theta = fl32(int32(b)) * (pi * 2^-31) in [-pi, pi)
//   n0 = r cos(theta), n1 = r sin(theta), r = sqrt(-2 ln u1)
// MUFU-based (lg2 / sqrt / sin / cos approx): abs error < 2e-6, see tests/test_philox.py.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
    const float u1 = (__uint2float_rn(a) + 1.0f) * 2.3283064365386963e-10f;
    const float th = __int2float_rn((int)b) * 1.4629180792671596e-9f;
    // sqrt.approx (MUFU): the IEEE sqrtf costs ~10 extra instructions and a divergent slow path per call; u1 in (0, 1]
    // keeps the argument in [0, 44.4], sqrt.approx.ftz(+0) = +0
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
    float s, c;
    __sincosf(th, &s, &c);
    n0 = r * c;
    n1 = r * s;
}

// 4 consecutive normals of the stream: elements [4q, 4q+4).
__device__ __forceinline__ void philox_normal4(uint64_t q, uint64_t offset, const PhiloxKey &key, float out[4]) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
    box_muller(r.x, r.y, out[0], out[1]);
    box_muller(r.z, r.w, out[2], out[3]);
}

// ---- eps stream of the bf16 kernels ("v2"): EIGHT normals per Philox call ------------------------------------------------
// The float32 stream above spends one Philox4x32-10 call (~60 integer instructions) per 4 normals, which makes the bf16
// reparameterisation kernel instruction-issue bound at ~50 % of the HBM roofline (profiles/r01_reparam_ncu_raw.csv).  The bf16
// kernels therefore draw 16-bit uniforms: block q = element / 8, word j of the block gives the pair (2j, 2j+1):
//   u1 = (lo16 + 1) * 2^-16 in (0, 1],   theta = (hi16 - 32768) * (pi / 32768) in [-pi, pi)
//   n_even = r cos(theta), n_odd = r sin(theta), r = sqrt(-2 ln u1)        (|n| <= 4.71; 2^32 distinct pairs)
// Host restatement: oracle/philox_ref.py philox_normal_v2.
__device__ __forceinline__ void box_muller16(uint32_t w, float &n0, float &n1) {
    const float u1 = __uint2float_rn((w & 0xffffu) + 1u) * 1.52587890625e-05f;
    const float th = __int2float_rn((int)(w >> 16) - 32768) * 9.587379924285257e-05f;
    // -2 ln u1 = (-2 ln 2) lg2 u1; u1 >= 2^-16 is never denormal: the .ftz forms skip the range fix-ups (3 instructions each)
    float l2, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));
    float s, c;
    __sincosf(th, &s, &c);
    n0 = r * c;
    n1 = r * s;
}
// 8 consecutive normals of the v2 stream: elements [8q, 8q+8).
__device__ __forceinline__ void philox_normal8(uint64_t q, uint64_t offset, const PhiloxKey &key, float out[8]) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
    box_muller16(r.x, out[0], out[1]);
    box_muller16(r.y, out[2], out[3]);
    box_muller16(r.z, out[4], out[5]);
    box_muller16(r.w, out[6], out[7]);
}

}  // namespace mlvae
