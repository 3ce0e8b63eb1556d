// philox.cuh — Philox4x32-R counter-based generator (Salmon, Moraes, Dror, Shaw; SC'11).
//
// Replaces the reference's hidden-state libc stream (srand at tiny_mc.c:43, rand() at
// photon.c:21,37,38,46).  Key = 64-bit seed, counter = (photon index lo, hi, draw block, 0),
// so a photon's whole trajectory is a pure function of (seed, global photon index).
//
// The key schedule (key + r * Weyl) depends on the seed only; the host precomputes it
// (tmc_api.cu) and the kernel reads the round keys straight from the constant bank, so one
// round is 2 x IMAD.WIDE.U32 + 2 x LOP3 and nothing else.
#pragma once
#include <cstdint>

namespace tmc {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr int kMaxPhiloxRounds = 10;

struct PhiloxKeys {
    uint32_t k[2 * kMaxPhiloxRounds];  // k[2r], k[2r+1] = round-r key words
};

inline void philox_expand_key(uint64_t seed, PhiloxKeys* out)
{
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    for (int r = 0; r < kMaxPhiloxRounds; ++r) {
        out->k[2 * r] = k0;
        out->k[2 * r + 1] = k1;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
}

#ifdef __CUDACC__
// rounds FIRST .. LAST-1 of Philox4x32 on the state (c0, c1, c2, c3)
template <int FIRST, int LAST>
__device__ __forceinline__ void philox4x32_rounds(const PhiloxKeys& keys, uint32_t c0, uint32_t c1, uint32_t c2,
                                                  uint32_t c3, uint32_t (&out)[4])
{
#if defined(TMC_EXPERIMENT) && TMC_EXPERIMENT == 3   /* 3 rounds only (timing experiment) */
    constexpr int LAST_ = FIRST + 2;
#else
    constexpr int LAST_ = LAST;
#endif
#pragma unroll
    for (int r = FIRST; r < LAST_; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;  // IMAD.WIDE.U32
        const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ keys.k[2 * r];      // LOP3
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ keys.k[2 * r + 1];  // LOP3
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <int ROUNDS>
__device__ __forceinline__ void philox4x32(const PhiloxKeys& keys, uint32_t c0, uint32_t c1, uint32_t c2,
                                           uint32_t c3, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;  // IMAD.WIDE.U32
        const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ keys.k[2 * r];      // LOP3
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ keys.k[2 * r + 1];  // LOP3
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#endif

}  // namespace tmc
