// walk_kernel.cuh — the photon random walk as a persistent-thread sm_100a kernel.
//
// Replaces the reference hot path: photon() (reference photon.c:6-51) and the loop that
// drives it (reference tiny_mc.c:47-49).  One launch simulates a whole range of photons.
//
// Design:
//  * persistent threads, TWO photons per thread: each thread owns two static strided lists of
//    photon indices and regenerates the next photon in place in the Philox block after the
//    one in which the current photon is killed by roulette, so warps never wait for their
//    longest-lived photon.  The two photons of a thread are processed as packed FP32 pairs
//    (Blackwell FFMA2 / FMUL2 / FADD2: two FMAs per issue slot) and give every thread two
//    independent dependency chains;
//  * random stream "tmc-stream-2": Philox4x32-R keyed by the seed, counter = (photon index,
//    block number).  One Philox call yields the four words of TWO scatter events; the birth
//    block gives the photon's roulette fate word, its (isotropic) launch direction and its
//    first event;
//  * weights are 32-bit fixed point, deposits are exact integers, tallies are u32 shared-
//    memory histograms privatised per block AND per lane ([shell][heat|heat2][lane]: every
//    lane of a warp owns its own bank, so an ATOMS.ADD is always one conflict-free wavefront,
//    also for the overflow bin that takes 20-67 % of the events), drained with atomicExch
//    every `flush_iters` iterations into u64 global tallies => the result is independent of
//    thread/block/GPU count and of atomic ordering (bit-reproducible).  Grids too fine for
//    per-lane copies (SHELLS > 760) use one u32 histogram per block plus 32 per-lane slots for
//    the overflow bin;
//  * MUFU: lg2 (step), sqrt (radius), sqrt (sin theta) = 3 per event; the azimuth (cos, sin)
//    comes from a 4096-entry table in shared memory (one LDS.64).
#pragma once
#include <cstdint>

#include "philox.cuh"

namespace tmc {

constexpr int kAzimuthBits = 12;
constexpr int kAzimuthEntries = 1 << kAzimuthBits;          // (cos, sin) pairs, 32 KB
constexpr uint32_t kAzimuthBytes = kAzimuthEntries * 8u;
constexpr uint32_t kLanePrivateMaxShells = 760u;            // 32 KB table + shells * 256 B <= 227 KB

struct WalkArgs {
    PhiloxKeys keys;                // constant-bank round keys
    uint64_t first;                 // first global photon index of this launch; the launch must not
    uint64_t count;                 // cross a multiple of 2^32, and count <= 2^30 (the host splits)
    unsigned long long* tallies;    // global u64[2*shells]: heat_fx | heat2_fx
    unsigned long long* counters;   // global u64[4]: events, photons, range flag, -
    const float2* azimuth;          // global (cos, sin)(2 pi i / 4096), i < 4096
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

// shared-memory bytes of one block
inline uint32_t walk_smem_bytes(uint32_t shells, bool lane_private)
{
    return kAzimuthBytes + (lane_private ? shells * 256u : 2u * (shells + 31u) * 4u);
}

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
// hi word of a*b + c (64-bit accumulate): one IMAD.WIDE.U32
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c)
{
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
// Loop-invariant 64-bit constant kept in a register pair (ptxas would otherwise turn the
// rounding add into a carry chain of two more instructions per use).
__device__ __forceinline__ uint64_t pinned_u64(uint64_t v)
{
    uint64_t r;
    asm volatile("mov.u64 %0, %1;" : "=l"(r) : "l"(v));
    return r;
}
// Roulette (reference photon.c:45-49) in five straight-line instructions.
__device__ __forceinline__ void roulette_one(uint32_t& w, uint32_t& fate, uint32_t thr)
{
    asm("{\n"
        " .reg .pred play, surv;\n"
        " .reg .u32 m;\n"
        " setp.lt.u32 play, %0, %2;\n"
        " setp.lt.and.u32 surv, %1, %3, play;\n"
        " selp.u32 m, 10, 0, surv;\n"
        " @play mul.lo.u32 %0, %0, m;\n"
        " @surv mul.lo.u32 %1, %1, 10;\n"
        "}\n"
        : "+r"(w), "+r"(fate)
        : "r"(thr), "n"(kFateSurvive));
}
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr)
{
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint32_t atom_shared_exch0(uint32_t addr)
{
    uint32_t r;
    asm volatile("atom.shared.exch.b32 %0, [%1], 0;" : "=r"(r) : "r"(addr) : "memory");
    return r;
}

// LANE_PRIVATE: bins[shell][kind][lane] (u32), kind 0 = heat, 1 = heat2; else heat[shells+31] | heat2[shells+31]
template <int ROUNDS, int BLOCK, int MIN_BLOCKS, bool LANE_PRIVATE>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) photon_walk_kernel(const __grid_constant__ WalkArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t bins_base = smem_base + kAzimuthBytes;
    const uint32_t plain_bins = a.shells + 31u;                              // per kind, plain layout
    const uint32_t nwords = LANE_PRIVATE ? a.shells * 64u : 2u * plain_bins;

    {   // stage the azimuth table and clear the histograms
        const float4* src = reinterpret_cast<const float4*>(a.azimuth);
        float4* dst = reinterpret_cast<float4*>(smem);
        for (uint32_t i = tid; i < kAzimuthEntries / 2; i += BLOCK) dst[i] = __ldg(src + i);
        uint32_t* bins = smem + kAzimuthBytes / 4u;
        for (uint32_t i = tid; i < nwords; i += BLOCK) bins[i] = 0u;
    }
    __syncthreads();

    // byte address of this lane's heat slot for magic-domain shell bits sb: (sb << SHIFT) + bias
    constexpr uint32_t SHIFT = LANE_PRIVATE ? 8u : 2u;
    const uint32_t addr_bias = bins_base + (LANE_PRIVATE ? lane * 4u : 0u) - (kMagicBits << SHIFT);
    const uint32_t heat2_off = LANE_PRIVATE ? 128u : plain_bins * 4u;
    // plain layout: lane l clamps to slot SHELLS-1+l, so the overflow bin never serialises a warp
    const uint32_t clamp_bits = LANE_PRIVATE ? a.last_bits : a.last_bits + lane;

    // static strided photon -> slot map: slot g owns photons first + g, first + g + stride, ...
    // rel = photon index - first, signed so that "before the first photon" is representable.
    const int32_t stride = static_cast<int32_t>(gridDim.x * (2u * BLOCK));
    const int32_t rel_limit = static_cast<int32_t>(a.count) - stride;     // a successor exists iff rel < rel_limit
    const uint32_t slot0 = (blockIdx.x * BLOCK + tid) * 2u;

    int32_t rel[2];
    uint32_t w[2], fate[2], blk[2];
    float2 px, py, pz, dx, dy, dz;      // .x = photon A, .y = photon B  (reference photon.c:12-17)
    px = py = pz = dx = dy = dz = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        rel[j] = static_cast<int32_t>(slot0 + j) - stride;   // the first regeneration steps onto slot0 + j
        w[j] = 0u;                                           // 0 <=> no live photon in this slot
        fate[j] = 0u;
        blk[j] = 1u;
    }
    uint32_t n_events = 0u, range_flag = 0u;

    // Philox round 0 with the counter folded in: c0 = first_lo + rel, c1 = first_hi, c3 = 0.
    //   M0 * c0 = M0 * rel + M0 * first_lo   (no wrap: the launch stays inside one 2^32 window)
    const uint32_t first_lo = static_cast<uint32_t>(a.first);
    const uint64_t m0_first = static_cast<uint64_t>(kPhiloxM0) * first_lo;
    const uint32_t c1k0 = static_cast<uint32_t>(a.first >> 32) ^ a.keys.k[0];

    const float2 half2 = make_float2(0.5f, 0.5f);
    const float2 negln2 = make_float2(-kLn2, -kLn2);
    const float2 bias2 = make_float2(kStepBias, kStepBias);
    const float2 spm2 = make_float2(a.shells_per_mfp, a.shells_per_mfp);
    const float2 magic2 = make_float2(8388608.0f, 8388608.0f);
    const float2 two2 = make_float2(2.0f, 2.0f), ntwo2 = make_float2(-2.0f, -2.0f);
    const float2 three2 = make_float2(3.0f, 3.0f), nthree2 = make_float2(-3.0f, -3.0f);
    const float2 one2 = make_float2(1.0f, 1.0f);
    const uint64_t round_half = pinned_u64(0x80000000ull);
    const uint64_t heat2_half = pinned_u64(a.heat2_half);

    // One scatter event for both photons of the thread: hop, drop (reference photon.c:21-32),
    // branch-free.  A slot without a live photon has w == 0: it deposits 0.
    auto scatter = [&](uint32_t wsA, uint32_t wsB) {
        n_events += min(w[0], 1u) + min(w[1], 1u);
        // hop: xi = (ws + 0.5) / 2^32, t = -ln(xi) = 32 ln2 - ln2 * lg2(ws + 0.5)
        float2 u = __fadd2_rn(make_float2(__uint2float_rn(wsA), __uint2float_rn(wsB)), half2);
        u.x = mufu_lg2(u.x);
        u.y = mufu_lg2(u.y);
        const float2 t = __ffma2_rn(u, negln2, bias2);
        px = __ffma2_rn(t, dx, px);
        py = __ffma2_rn(t, dy, py);
        pz = __ffma2_rn(t, dz, pz);
        // drop: shell = min(trunc(|r| * shells_per_mfp), SHELLS-1) without F2I: add 2^23 with
        // round-toward-zero, clamp the raw bits, the mantissa is the integer.
        float2 r2 = __ffma2_rn(pz, pz, __ffma2_rn(py, py, __fmul2_rn(px, px)));
        r2.x = mufu_sqrt(r2.x);
        r2.y = mufu_sqrt(r2.y);
        const float2 sbf = __ffma2_rz(r2, spm2, magic2);
        const uint32_t sbits[2] = { min(__float_as_uint(sbf.x), clamp_bits), min(__float_as_uint(sbf.y), clamp_bits) };
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            // deposit (1-albedo) * w, rounded: hi32(w * q32 + 2^31); its square, rescaled
            const uint32_t dep = static_cast<uint32_t>(mad_wide(w[j], a.absorb_q32, round_half) >> 32);
            const uint32_t dep2 = static_cast<uint32_t>(mad_wide(dep, dep, heat2_half) >> a.heat2_rshift);
            w[j] -= dep;                                                          // w *= albedo
            const uint32_t addr = (sbits[j] << SHIFT) + addr_bias;
            red_shared_add(addr, dep);
            red_shared_add(addr + heat2_off, dep2);
        }
    };
    // roulette (reference photon.c:45-49).  The fate word is a uniform 32-bit integer drawn at
    // birth; surviving (prob 0.1) multiplies it by 10, which is again uniform.  Death: w = 0.
    auto roulette = [&]() {
        roulette_one(w[0], fate[0], a.roulette_thr);
        roulette_one(w[1], fate[1], a.roulette_thr);
    };
    // New isotropic direction from one 32-bit word per photon (replaces the rejection loop of
    // reference photon.c:35-43): cos(theta) uniform from bits 9..31, azimuth from bits 3..14.
    auto spin = [&](uint32_t wdA, uint32_t wdB) {
        const float2 cf = make_float2(__uint_as_float(__funnelshift_r(wdA, 0x7Fu, 9)),    // 1 + m*2^-23 in [1,2)
                                      __uint_as_float(__funnelshift_r(wdB, 0x7Fu, 9)));
        const float2 ct = __ffma2_rn(cf, two2, nthree2);                                  // [-1, 1), exact
        const float2 nct = __ffma2_rn(cf, ntwo2, three2);                                 // -ct, exact
        float2 st = __ffma2_rn(ct, nct, one2);                                            // 1 - ct^2
        st.x = mufu_sqrt(st.x);
        st.y = mufu_sqrt(st.y);
        const float2 csA = lds_f32x2(smem_base + (wdA & ((kAzimuthEntries - 1u) << 3)));
        const float2 csB = lds_f32x2(smem_base + (wdB & ((kAzimuthEntries - 1u) << 3)));
        dx = ct;
        dy = make_float2(st.x * csA.x, st.y * csB.x);
        dz = make_float2(st.x * csA.y, st.y * csB.y);
    };
    // Drain the block histograms into the global u64 tallies (atomicExch: no barrier needed,
    // other warps keep adding).  Warp-uniform: every lane of the warp must be here.
    auto drain = [&]() {
        if constexpr (LANE_PRIVATE) {
            const uint32_t rows = 2u * a.shells;           // row = shell * 2 + kind, 32 lanes wide
            for (uint32_t r = tid >> 5; r < rows; r += BLOCK / 32) {
                const uint32_t v = atom_shared_exch0(bins_base + (r * 32u + lane) * 4u);
                range_flag |= v >> 31;
                if (__any_sync(0xffffffffu, v != 0u)) {
                    const uint32_t lo = __reduce_add_sync(0xffffffffu, v & 0xFFFFu);
                    const uint32_t hi = __reduce_add_sync(0xffffffffu, v >> 16);
                    if (lane == 0u)
                        atomicAdd(&a.tallies[(r & 1u) * a.shells + (r >> 1)], (static_cast<unsigned long long>(hi) << 16) + lo);
                }
            }
        } else {
            for (uint32_t i = tid; i < nwords; i += BLOCK) {
                const uint32_t v = atom_shared_exch0(bins_base + i * 4u);
                if (v != 0u) {
                    range_flag |= v >> 31;
                    const uint32_t kind = i >= plain_bins ? 1u : 0u;
                    const uint32_t s = min(i - kind * plain_bins, a.shells - 1u);
                    atomicAdd(&a.tallies[kind * a.shells + s], static_cast<unsigned long long>(v));
                }
            }
        }
    };

    bool more = true;
    while (more) {
        for (uint32_t it = 0; it < a.flush_iters; ++it) {
            uint32_t r[2][4];
            bool born[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                born[j] = (w[j] == 0u) && (rel[j] < rel_limit);   // regenerate in place
                if (born[j]) {
                    rel[j] += stride;
                    blk[j] = 0u;
                }
                const uint64_t p0 = mad_wide(static_cast<uint32_t>(rel[j]), kPhiloxM0, m0_first);
                const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * blk[j];
                philox4x32_rounds<1, ROUNDS>(a.keys, static_cast<uint32_t>(p1 >> 32) ^ c1k0, static_cast<uint32_t>(p1),
                                             static_cast<uint32_t>(p0 >> 32) ^ a.keys.k[1], static_cast<uint32_t>(p0), r[j]);
                ++blk[j];
            }
            // slot A.  In its birth block a photon still has w == 0, so the scatter deposits
            // nothing; then it gets its weight, its fate word, the origin, and (from the spin
            // of this slot) its launch direction.
            scatter(r[0][0], r[1][0]);
            if (born[0]) { w[0] = a.weight_one; fate[0] = r[0][0]; px.x = 0.0f; py.x = 0.0f; pz.x = 0.0f; }
            if (born[1]) { w[1] = a.weight_one; fate[1] = r[1][0]; px.y = 0.0f; py.y = 0.0f; pz.y = 0.0f; }
            roulette();
            spin(r[0][1], r[1][1]);
            // slot B
            scatter(r[0][2], r[1][2]);
            roulette();
            spin(r[0][3], r[1][3]);
        }
        drain();
        more = __any_sync(0xffffffffu, (w[0] | w[1]) != 0u || rel[0] < rel_limit || rel[1] < rel_limit);
    }

    // Final drain once every thread of the block is done.
    __syncthreads();
    drain();
    unsigned long long ev = n_events;
    uint32_t fl = range_flag;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ev += __shfl_xor_sync(0xffffffffu, ev, o);
        fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    }
    if (lane == 0u) {
        atomicAdd(&a.counters[0], ev);
        if (fl) atomicOr(&a.counters[2], 1ull);
    }
    if (tid == 0u && blockIdx.x == 0u) atomicAdd(&a.counters[1], static_cast<unsigned long long>(a.count));
}

#endif  // __CUDACC__

}  // namespace tmc
