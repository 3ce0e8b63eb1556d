// walk_kernel.cuh — the photon random walk as a persistent-thread sm_100a kernel.
//
// Replaces the reference hot path: photon() (reference photon.c:6-51) and the loop that
// drives it (reference tiny_mc.c:47-49).  One launch simulates a whole range of photons.
//
// Design:
//  * persistent threads: each thread owns a static strided list of photon indices and
//    regenerates the next photon in place at the end of the iteration in which the current
//    one is killed by roulette, so warps never wait for their longest-lived photon;
//  * random stream "tmc-stream-1": Philox4x32-R keyed by the seed, counter = (photon index,
//    draw block).  One Philox call yields the four words of TWO scatter events; the birth
//    block gives the photon's roulette fate word and its (isotropic) launch direction;
//  * weights are 32-bit fixed point, deposits are exact integers, tallies are u32 shared-
//    memory histograms privatised per block (overflow bin: per-thread registers), drained with
//    atomicExch every `flush_iters` iterations into u64 global tallies => the result is
//    independent of thread/block/GPU count and of atomic ordering (bit-reproducible);
//  * MUFU: lg2 (step), sqrt (radius), sqrt + sin + cos (direction) = 5 per event.
#pragma once
#include <cstdint>

#include "philox.cuh"

namespace tmc {

struct WalkArgs {
    PhiloxKeys keys;                // constant-bank round keys
    uint64_t first;                 // first global photon index of this launch
    uint64_t count;                 // photons in this launch
    unsigned long long* tallies;    // global u64[2*shells]: heat_fx | heat2_fx
    unsigned long long* counters;   // global u64[4]: events, photons, range flag, -
    float shells_per_mfp;           // reference photon.c:9
    uint32_t shells;                // SHELLS (reference params.h:5)
    uint32_t last_bits;             // 0x4B000000 + SHELLS-1 : clamp in the magic-number domain
    uint32_t weight_one;            // fixed-point 1.0
    uint32_t absorb_q32;            // round((1-albedo) * 2^32)  (reference photon.c:8,30)
    uint32_t heat2_rshift;          // deposit^2 >> heat2_rshift
    uint32_t heat2_half;            // rounding constant for that shift
    uint32_t roulette_thr;          // fixed-point 0.001 (reference photon.c:45)
    uint32_t flush_iters;           // iterations between drains of the shared histograms
};

constexpr uint32_t kMagicBits = 0x4B000000u;     // float 2^23
constexpr uint32_t kFateSurvive = 429496729u;    // floor(0.1 * 2^32): survive roulette iff fate < this
constexpr float kLn2 = 0.693147182464599609375f;
constexpr float kStepBias = 22.1807098388671875f;  // 32 * ln 2
constexpr float kAzimuthScale = 804.24774169921875f;  // float(256 * pi)

#ifdef __CUDACC__

__device__ __forceinline__ float mufu_lg2(float v)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float mufu_sqrt(float v)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float mufu_sin(float v)
{
    float r;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float mufu_cos(float v)
{
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// hi word of a*b + c (64-bit accumulate): one IMAD.WIDE.U32
__device__ __forceinline__ uint32_t mad_wide_hi(uint32_t a, uint32_t b, uint64_t c)
{
    return static_cast<uint32_t>((static_cast<uint64_t>(a) * b + c) >> 32);
}

struct Photon {
    float x, y, z;      // position, mean-free-path units (reference photon.c:12-14)
    float dx, dy, dz;   // direction cosines              (reference photon.c:15-17)
    uint32_t w;         // fixed-point weight             (reference photon.c:18); 0 <=> no live photon
    uint32_t fate;      // roulette fate word
};

// Loop-invariant constants that ptxas would otherwise re-materialise with a MOV per use.
__device__ __forceinline__ uint32_t pinned_u32(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ float pinned_f32(float v)
{
    float r;
    asm volatile("mov.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// New isotropic direction from one 32-bit word (replaces the rejection loop of
// reference photon.c:35-43): cos(theta) uniform from the top 23 bits, azimuth from the low 16.
__device__ __forceinline__ void spin(Photon& p, uint32_t wd, float az_scale)
{
    const float cf = __uint_as_float(__funnelshift_r(wd, 0x7Fu, 9));   // 1 + m*2^-23 in [1,2)
    const float ct = fmaf(cf, 2.0f, -3.0f);                              // [-1, 1)
    const float st = mufu_sqrt(fmaf(-ct, ct, 1.0f));
    const float af = __uint_as_float(__byte_perm(wd, 0x3F800000u, 0x7610));  // 1 + j*2^-23
    const float ang = fmaf(af, az_scale, -az_scale);                    // 2*pi*j/65536, exact FMA
    p.dx = ct;
    p.dy = st * mufu_cos(ang);
    p.dz = st * mufu_sin(ang);
}

// Tally one deposit: overflow bin SHELLS-1 -> per-thread registers, every other shell -> the
// block's shared-memory histograms (reference photon.c:27-31).  Straight-line, predicated.
__device__ __forceinline__ void tally(uint32_t sb, uint32_t last_bits, uint32_t addr, uint32_t addr2,
                                      uint32_t dep, uint32_t dep2, uint32_t& ov_heat, uint32_t& ov_heat2)
{
    asm volatile(
        "{\n"
        " .reg .pred ov;\n"
        " setp.eq.u32 ov, %2, %3;\n"
        " @ov add.u32 %0, %0, %6;\n"
        " @ov add.u32 %1, %1, %7;\n"
        " @!ov red.shared.add.u32 [%4], %6;\n"
        " @!ov red.shared.add.u32 [%5], %7;\n"
        "}\n"
        : "+r"(ov_heat), "+r"(ov_heat2)
        : "r"(sb), "r"(last_bits), "r"(addr), "r"(addr2), "r"(dep), "r"(dep2)
        : "memory");
}

template <int ROUNDS, int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) photon_walk_kernel(const __grid_constant__ WalkArgs a)
{
    extern __shared__ uint32_t bins[];   // heat_fx[shells] | heat2_fx[shells], u32, block-private
    const uint32_t tid = threadIdx.x;
    const uint32_t nbins = 2u * a.shells;
    for (uint32_t i = tid; i < nbins; i += BLOCK) bins[i] = 0u;
    __syncthreads();

    // shared-window byte address of heat_fx[shell] is (sb << 2) + addr_bias, sb = magic bits
    const uint32_t addr_bias = static_cast<uint32_t>(__cvta_generic_to_shared(bins)) - (kMagicBits << 2);
    const uint32_t heat2_off = a.shells * 4u;

    // static strided photon -> thread map: thread g owns first + g, first + g + stride, ...
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * BLOCK;
    const uint64_t gtid = static_cast<uint64_t>(blockIdx.x) * BLOCK + tid;
    uint32_t remaining = gtid < a.count ? static_cast<uint32_t>((a.count - gtid + stride - 1) / stride) : 0u;
    uint64_t idx = a.first + gtid - stride;   // the first regeneration steps onto first + gtid

    Photon p;
    p.x = p.y = p.z = 0.0f;
    p.dx = p.dy = p.dz = 0.0f;
    p.w = 0u;
    p.fate = 0u;
    uint32_t blk = 1u;

    uint32_t ov_heat = 0u, ov_heat2 = 0u;               // overflow bin SHELLS-1, per thread
    unsigned long long ov_heat64 = 0ull, ov_heat2_64 = 0ull;
    uint32_t n_events = 0u, range_flag = 0u;

    const uint32_t round_half = pinned_u32(0x80000000u);
    const uint32_t zero = pinned_u32(0u);
    const float neg_ln2 = pinned_f32(-kLn2);
    const float az_scale = pinned_f32(kAzimuthScale);

    // One scatter event: hop, drop, roulette (reference photon.c:21-32,45-49), branch-free.
    // A lane without a live photon has w == 0 and direction 0: it deposits 0 and stays put.
    auto scatter = [&](uint32_t ws) {
        // hop: xi = (ws + 0.5) / 2^32, t = -ln(xi) = 32 ln2 - ln2 * lg2(ws + 0.5)
        const float t = fmaf(mufu_lg2(__uint2float_rn(ws) + 0.5f), neg_ln2, kStepBias);
        p.x = fmaf(t, p.dx, p.x);
        p.y = fmaf(t, p.dy, p.y);
        p.z = fmaf(t, p.dz, p.z);
        // drop: shell = min(trunc(|r| * shells_per_mfp), SHELLS-1) without F2I: add 2^23 with
        // round-toward-zero, clamp the raw bits, the mantissa is the integer.
        const float rad = mufu_sqrt(fmaf(p.z, p.z, fmaf(p.y, p.y, p.x * p.x)));
        const uint32_t sb = min(__float_as_uint(__fmaf_rz(rad, a.shells_per_mfp, 8388608.0f)), a.last_bits);
        // deposit (1-albedo) * w, rounded: hi32(w * q32 + 2^31)
        const uint64_t acc = static_cast<uint64_t>(p.w) * a.absorb_q32 +
                             (static_cast<uint64_t>(zero) << 32 | round_half);
        const uint32_t dep = static_cast<uint32_t>(acc >> 32);
        const uint32_t dep2 = static_cast<uint32_t>(
            (static_cast<uint64_t>(dep) * dep + a.heat2_half) >> a.heat2_rshift);
        p.w -= dep;                                                               // w *= albedo
        const uint32_t addr = (sb << 2) + addr_bias;
        tally(sb, a.last_bits, addr, addr + heat2_off, dep, dep2, ov_heat, ov_heat2);
        n_events += (dep != 0u) ? 1u : 0u;
    };
    // roulette (reference photon.c:45-49).  The fate word is a uniform 32-bit integer drawn at
    // birth; surviving (prob 0.1) multiplies it by 10, which is again uniform.  Death: w = 0.
    auto roulette = [&]() {
        const bool play = p.w < a.roulette_thr;
        const bool survive = play && (p.fate < kFateSurvive);
        if (survive) p.fate *= 10u;
        if (play) p.w *= survive ? 10u : 0u;
    };

    bool more = remaining != 0u;
    while (more) {
        for (uint32_t it = 0; it < a.flush_iters; ++it) {
            if (p.w == 0u && remaining != 0u) {   // regenerate in place
                --remaining;
                idx += stride;
                blk = 0u;
                p.x = p.y = p.z = 0.0f;
                p.dx = p.dy = p.dz = 0.0f;
            }
            uint32_t r[4];
            philox4x32<ROUNDS>(a.keys, static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), blk, 0u, r);
            const bool born = (blk == 0u);
            // slot A.  In the birth block the photon still has w == 0 and no direction, so the
            // scatter is a no-op; then it gets its weight, fate word and launch direction.
            scatter(r[0]);
            if (born) {
                p.w = a.weight_one;
                p.fate = r[0];
            }
            roulette();
            spin(p, r[1], az_scale);
            // slot B
            scatter(r[2]);
            roulette();
            spin(p, r[3], az_scale);
            if (p.w == 0u) p.dx = p.dy = p.dz = 0.0f;   // killed: freeze until regenerated
            ++blk;
        }
        // Drain this thread's share of the block histogram (atomicExch: no barrier needed).
        for (uint32_t i = tid; i < nbins; i += BLOCK) {
            if (bins[i] != 0u) {
                const uint32_t v = atomicExch(&bins[i], 0u);
                range_flag |= v >> 31;
                atomicAdd(&a.tallies[i], static_cast<unsigned long long>(v));
            }
        }
        ov_heat64 += ov_heat;
        ov_heat2_64 += ov_heat2;
        ov_heat = ov_heat2 = 0u;
        more = (p.w != 0u) || (remaining != 0u);
    }

    // Final drain once every thread of the block is done.
    __syncthreads();
    for (uint32_t i = tid; i < nbins; i += BLOCK) {
        const uint32_t v = bins[i];
        if (v != 0u) {
            range_flag |= v >> 31;
            atomicAdd(&a.tallies[i], static_cast<unsigned long long>(v));
        }
    }
    // Per-thread overflow-bin accumulators and counters: warp reduce, one atomic per warp.
    unsigned long long ev = n_events, fl = range_flag;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ov_heat64 += __shfl_xor_sync(0xffffffffu, ov_heat64, o);
        ov_heat2_64 += __shfl_xor_sync(0xffffffffu, ov_heat2_64, o);
        ev += __shfl_xor_sync(0xffffffffu, ev, o);
        fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    }
    if ((tid & 31u) == 0u) {
        if (ov_heat64) atomicAdd(&a.tallies[a.shells - 1u], ov_heat64);
        if (ov_heat2_64) atomicAdd(&a.tallies[nbins - 1u], ov_heat2_64);
        atomicAdd(&a.counters[0], ev);
        if (fl) atomicOr(&a.counters[2], 1ull);
    }
    if (tid == 0u && blockIdx.x == 0u) atomicAdd(&a.counters[1], static_cast<unsigned long long>(a.count));
}

#endif  // __CUDACC__

}  // namespace tmc
