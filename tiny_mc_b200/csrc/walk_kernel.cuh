// walk_kernel.cuh — the photon random walk as a persistent-warp sm_100a kernel.
//
// Replaces the reference hot path: photon() (reference photon.c:6-51) and the loop that
// drives it (reference tiny_mc.c:47-49).  One launch simulates a whole range of photons.
//
// Design:
//  * The weight of a photon is a pure function of how many events it has lived: every event
//    removes the fraction (1-albedo) (photon.c:30-32) and the roulette (photon.c:45-49) is the
//    only random influence, multiplying by exactly 10 when it is survived.  So all photons of
//    one GENERATION (= number of roulettes survived) carry the same weight at the same event
//    number, reach the roulette at the same event, and deposit the same amount per event.
//  * Persistent warps therefore walk COHORTS of 64 photons of one generation (two per lane,
//    two independent dependency chains per thread) in lock step: no lane ever waits for
//    another, the weight / deposit arithmetic is warp-uniform (the host tabulates it with the
//    exact integer recurrence; the kernel reads one broadcast table entry per event), and
//    nothing in the event loop branches per lane.  When a generation ends, every lane plays
//    roulette with its photon's fate word; the ~10 % survivors are parked in a per-warp queue
//    of the next generation (global scratch, a few accesses per cohort), the warp immediately
//    regenerates 64 fresh photons in place (or, when 64 survivors have accumulated, a full
//    cohort of them).  Generations beyond the third (1e-4 of the photons) continue in place
//    with the dead lanes masked.
//  * Random stream "tmc-stream-4": Philox4x32-R keyed by the seed, counter = (photon index,
//    block).  One Philox block = FOUR events, one 32-bit word each; event e of a photon is word
//    e % 4 of block e / 4, pseudo-event 0 is the roulette fate word.  A photon's trajectory
//    depends on (seed, photon index) only.
//  * Deposits are exact integers (32-bit fixed-point weights); tallies are u32 shared-memory
//    histograms privatised per block AND per lane ([shell][heat|heat2][lane]: every lane of a
//    warp owns its own bank, an ATOMS.ADD is always one conflict-free wavefront, also for the
//    overflow bin that takes 20-67 % of the events), drained with atomicExch into u64 global
//    tallies => the result does not depend on thread/block/GPU count or on atomic ordering.
//    Grids too fine for per-lane copies use one u32 histogram per block plus 32 per-lane
//    slots for the overflow bin.
//  * Per event and photon: 2 MUFU (lg2 for the step, sqrt for the radius), 10 FP32 operations
//    in 8 instructions (the (y, z) update is one FFMA2, xi and the shell number of the lane's two
//    photons come out of one FADD2 / FFMA2.RZ: Blackwell's packed f32x2, TMC_PACKED),
//    2 shared atomics, two conflict-free LDS.64 for the direction (polar and azimuth tables,
//    each replicated over 16 bank pairs so that lane l only ever reads bank pair l % 16) and a
//    quarter of a Philox block.
#pragma once
#include <cstdint>

#include "philox.cuh"

#ifndef TMC_GROUP
#define TMC_GROUP 4        /* events of a Philox block whose table look-ups are issued together (walk loop) */
#endif
#ifndef TMC_PPL
#define TMC_PPL 2          /* photons per lane (independent dependency chains per thread); a cohort is 32 * TMC_PPL photons */
#endif
#ifndef TMC_PACKED
#define TMC_PACKED 7       /* packed fp32 pairs (Blackwell f32x2) in the walk loop, a bit mask: 1 = FFMA2 for the (y, z) update of an
                              event, 2 = FADD2 for xi of two photons, 4 = FFMA2.RZ for the shell numbers of two photons */
#endif
#ifndef TMC_EXPERIMENT
#define TMC_EXPERIMENT 0   /* timing experiments (tools/experiments.sh); the product is built with 0 */
#endif

namespace tmc {

constexpr int kPacked = TMC_PACKED;
constexpr uint32_t kEventsPerBlock = 4u;                    // one 32-bit Philox word per scatter event
constexpr int kDirEntries = 256;                            // polar midpoints / azimuths (8 bits each)
// Direction table in shared memory: row k (256 B) = 16 copies of (-ln2 cos, -ln2 sin)(theta_k) followed
// by 16 copies of (cos, sin)(phi_k).  Lane l reads copy l % 16: an LDS.64 touches every bank pair once
// per half-warp, whatever the 32 random row numbers are (2 wavefronts, the minimum for 256 B).
constexpr uint32_t kDirTableBytes = kDirEntries * 256u;     // 64 KB
// Shared-memory map of a block, in ABSOLUTE addresses of the CTA's shared window, so that every
// LDS / ATOMS of the event loop carries its base as an immediate ([R + imm]) and the index register
// comes straight out of one PRMT: user shared memory starts at 0x400 (CUDA reserves the first KB),
// the table sits at 0x800, the histograms behind it.  The kernel checks the assumption at start-up.
constexpr uint32_t kSmemUserBase = 0x400u;
constexpr uint32_t kSmemTableAbs = 0x800u;
constexpr uint32_t kSmemBinsAbs = kSmemTableAbs + kDirTableBytes;
constexpr int kMaxGenerations = 24;                         // P(survive 24 roulettes) = 1e-24
constexpr uint32_t kQueueCap = 64u * TMC_PPL;               // entries per queued generation and warp: two cohorts
constexpr uint32_t kQueueFields = 5u;                       // x, y, z, photon offset, fate word
#ifndef TMC_QUEUED_GENS
#define TMC_QUEUED_GENS 3  /* generations whose survivors are parked until a full cohort has accumulated (1 .. TMC_QUEUED_GENS) */
#endif
constexpr uint32_t kQueuedGens = TMC_QUEUED_GENS;
constexpr uint32_t kQueueBytesPerWarp = kQueuedGens * kQueueFields * kQueueCap * 4u;
constexpr uint32_t kLanePrivateMaxShells = 512u;

struct GenPlan {
    uint32_t first_event;   // 1-based number of the generation's first event
    uint32_t n_events;      // events until the weight falls below the roulette threshold
    uint32_t w_start;       // fixed-point weight at the start of the generation
};

struct WalkArgs {
    PhiloxKeys keys;                // constant-bank round keys
    uint64_t first;                 // first global photon index of this launch; the launch must not
    uint64_t count;                 // cross a multiple of 2^32, and count <= 2^30 (the host splits)
    unsigned long long* tallies;    // global u64[2*shells]: heat_fx | heat2_fx
    unsigned long long* counters;   // global u64[4]: events, photons, range flag, shared-memory-map error
    const float2* directions;       // global [256] (-ln2 cos, -ln2 sin)(theta_k) | [256] (cos, sin)(phi_k)
    const uint2* deposits;          // global (deposit, rescaled deposit^2) of event e, e = 1 .. last event of gen[n_gen-1]
    uint32_t* queues;               // global scratch: kQueueBytesPerWarp per warp of the grid (survivor queues)
    float pos_scale;                // kappa = shells_per_mfp / SHELLS (photon.c:9): positions are kept in units of the grid
                                    // radius SHELLS / shells_per_mfp mean free paths, so that |r| = 1 is the outer edge of the
                                    // last shell and the clamp of photon.c:27-29 becomes the .sat of one FFMA
    float shell_scale;              // the largest float below SHELLS: trunc(|r| * shell_scale) <= SHELLS-1 for |r| <= 1
    float radial_step;              // -ln2 * kappa (reduced radial walk only)
    uint32_t shells;                // SHELLS (reference params.h:5)
    uint32_t last_bits;             // 0x4B000000 + SHELLS-1 : clamp in the magic-number domain
    uint32_t flush_blocks;          // Philox blocks (4 events) a warp walks between drains
    uint32_t check_shift;           // a drained word >= 2^check_shift raises the range flag (31)
    uint32_t n_gen;                 // generations in `gen` (a photon surviving them all is dropped)
    GenPlan gen[kMaxGenerations];
};

constexpr uint32_t kMagicBits = 0x4B000000u;     // float 2^23
constexpr uint32_t kFateSurvive = 429496729u;    // floor(0.1 * 2^32): survive roulette iff fate < this
constexpr float kLn2 = 0.693147182464599609375f;
constexpr float kOneMinusHalfUlp = 0.999999940395355224609375f;   // 1 - 2^-24

// shared-memory bytes of one block
inline uint32_t walk_smem_bytes(uint32_t shells, bool lane_private, uint32_t block_threads)
{
    (void)block_threads;
    return (kSmemTableAbs - kSmemUserBase) + kDirTableBytes + (lane_private ? shells * 256u : 2u * (shells + 31u) * 4u);
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
// a*b + c with a 64-bit accumulator: one IMAD.WIDE.U32
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c)
{
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
// (c0, c1) += s * (b.x, b.y) as ONE packed instruction (FFMA2 with a scalar first operand): two
// independent fp32 FMAs, each rounded exactly like fmaf
__device__ __forceinline__ void fma2_scalar(float s, float2 b, float& c0, float& c1)
{
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ra) : "f"(s));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(rc));
}
// (a0, a1) + (c, c) as one FADD2
__device__ __forceinline__ void add2_scalar(float a0, float a1, float c, float& r0, float& r1)
{
    unsigned long long ra, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(rc) : "f"(c));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(ra) : "l"(rc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(ra));
}
// the bits of (a0, a1) * s + m rounded toward zero, as one FFMA2.RZ
__device__ __forceinline__ void fma2_rz_bits(float a0, float a1, float s, float m, uint32_t& r0, uint32_t& r1)
{
    unsigned long long ra, rs, rm;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(rs) : "f"(s));
    asm("mov.b64 %0, {%1, %1};" : "=l"(rm) : "f"(m));
    asm("fma.rz.f32x2 %0, %0, %1, %2;" : "+l"(ra) : "l"(rs), "l"(rm));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r0), "=r"(r1) : "l"(ra));
}
// byte SEL of `word` into byte 1 of the result, byte 0 of `low` into byte 0, zeros above: the byte
// offset row * 256 + low of a direction-table entry in ONE instruction (PRMT)
template <int SEL>
__device__ __forceinline__ uint32_t row_offset(uint32_t word, uint32_t low)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(low), "n"(0x5504 | (SEL << 4)));
    return r;
}
// bytes 0-1 of `word` (a shell number) into bytes 1-2, byte 0 of `low` into byte 0: shell * 256 + low
__device__ __forceinline__ uint32_t shell_offset(uint32_t word, uint32_t low)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x5104;" : "=r"(r) : "r"(word), "r"(low));
    return r;
}
// The direction table is read-only after the block's first barrier: a pure function of the offset.
template <uint32_t BASE>
__device__ __forceinline__ float2 lds_f32x2(uint32_t off)
{
    float2 r;
    asm("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(r.x), "=f"(r.y) : "r"(off), "n"(BASE));
    return r;
}
// No "memory" clobber: the histograms are touched by these atomics and by drain_slice (an
// out-of-line call) only, so the compiler stays free to hoist the next event's table look-up
// and arithmetic above the atomics of this one.
template <uint32_t BASE>
__device__ __forceinline__ void red_shared_add(uint32_t off, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(off), "r"(v), "n"(BASE));
}
__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t r;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
    return r;
}

template <bool V>
struct BoolTag {
    static constexpr bool value = V;
};
template <int V>
struct IntTag {
    static constexpr int value = V;
};

// Move slice `slice` (of BLOCK/32) of the block's u32 histograms into the global u64 tallies.
// atomicExch: no barrier needed, the other warps keep adding meanwhile.  Called by a whole
// converged warp; returns bit 0 set when a word had come within a factor two of wrapping.
template <int BLOCK, bool LANE_PRIVATE>
__device__ __noinline__ uint32_t drain_slice(uint32_t* bins, unsigned long long* tallies, uint32_t shells, uint32_t slice,
                                             uint32_t check_shift)
{
    constexpr uint32_t WARPS = BLOCK / 32;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t flag = 0u;
    if constexpr (LANE_PRIVATE) {
        const uint32_t rows = 2u * shells;             // row = shell * 2 + kind, 32 lanes wide
        for (uint32_t row = slice; row < rows; row += WARPS) {
            const uint32_t v = atomicExch(&bins[row * 32u + lane], 0u);
            flag |= v >> check_shift;
            if (__any_sync(0xffffffffu, v != 0u)) {
                const uint32_t lo = __reduce_add_sync(0xffffffffu, v & 0xFFFFu);
                const uint32_t hi = __reduce_add_sync(0xffffffffu, v >> 16);
                if (lane == 0u)
                    atomicAdd(&tallies[(row & 1u) * shells + (row >> 1)], (static_cast<unsigned long long>(hi) << 16) + lo);
            }
        }
    } else {
        const uint32_t plain_bins = shells + 31u;
        for (uint32_t i = slice * 32u + lane; i < 2u * plain_bins; i += BLOCK) {
            const uint32_t v = atomicExch(&bins[i], 0u);
            if (v != 0u) {
                flag |= v >> check_shift;
                const uint32_t kind = i >= plain_bins ? 1u : 0u;
                const uint32_t s = min(i - kind * plain_bins, shells - 1u);
                atomicAdd(&tallies[kind * shells + s], static_cast<unsigned long long>(v));
            }
        }
    }
    return flag;
}

// LANE_PRIVATE: bins[shell][kind][lane] (u32), kind 0 = heat, 1 = heat2; else heat[shells+31] | heat2[shells+31]
// RADIAL: the reduced walk of SURVEY §8f rank 4 — only |r| is tallied and scattering is isotropic,
//         so r'^2 = r^2 + t^2 + 2 r t mu with mu ~ U[-1, 1] is distribution-identical and needs neither
//         a position vector nor an azimuth.  A separately selectable cross-check ("walk_mode" = 1):
//         NOT the path the north star names (it skips the position update and the direction
//         resampling), never used for the headline or the roofline figure.
// PPL:    photons per lane (independent dependency chains per thread); a cohort is 32 * PPL photons.
// SAT_PLAIN: the one-histogram layout with the saturating clamp of the lane-private layout instead of the integer clamp to
//         per-lane overflow slots (one VIMNMX per event less).  Every |r| >= 1 then lands in the ONE word of shell SHELLS-1, which
//         serialises the lanes that are out there: for grids no photon leaves in practice (the host decides, tmc_api.cu).
template <int ROUNDS, int BLOCK, int MIN_BLOCKS, bool LANE_PRIVATE, bool RADIAL = false, int PPL = TMC_PPL, bool SAT_PLAIN = false>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) photon_walk_kernel(const __grid_constant__ WalkArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];     // [gap | direction table 64 KB | histograms]
    __shared__ uint32_t drain_ticket;
    constexpr uint32_t WARPS = BLOCK / 32;
    constexpr uint32_t COHORT = 32u * PPL;
    constexpr bool SAT_CLAMP = LANE_PRIVATE || SAT_PLAIN;   // the clamp of photon.c:27-29 as the .sat of r^2's last FFMA
    static_assert(2u * COHORT <= kQueueCap, "a survivor queue must hold two cohorts");
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t wid = tid >> 5;
    // the table must start at kSmemTableAbs of the shared window (walk_smem_bytes() pays for the gap)
    const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    if (smem_base < kSmemUserBase || smem_base > kSmemTableAbs) {
        if (tid == 0u) atomicOr(&a.counters[3], 1ull);      // reported by the host as an internal error
        return;
    }
    uint32_t* const table = smem + (kSmemTableAbs - smem_base) / 4u;
    uint32_t* const bins = table + kDirTableBytes / 4u;
    const uint32_t plain_bins = a.shells + 31u;                              // per kind, plain layout
    const uint32_t nwords = LANE_PRIVATE ? a.shells * 64u : 2u * plain_bins;
    // this warp's survivor queues: [generation 1|2|3]{uint4 (x, y, z, photon offset)[kQueueCap], fate[kQueueCap]}, in global memory (a few
    // accesses per cohort; L2-resident), so that shared memory holds only the table and tallies
    uint32_t* const queue = a.queues + static_cast<size_t>(blockIdx.x * WARPS + wid) * (kQueueBytesPerWarp / 4u);

    {   // stage the direction table (16 copies of every entry) and clear the histograms
        float2* dst = reinterpret_cast<float2*>(table);
        for (uint32_t i = tid; i < kDirTableBytes / 8u; i += BLOCK) {
            float2 d = __ldg(a.directions + (i >> 5) + ((i & 16u) ? kDirEntries : 0));
            if (!(i & 16u)) {               // polar entries carry the step: scale them to grid-radius units
                d.x *= a.pos_scale;
                d.y *= a.pos_scale;
            }
            dst[i] = d;
        }
        for (uint32_t i = tid; i < nwords; i += BLOCK) bins[i] = 0u;
        if (tid == 0u) drain_ticket = 0u;
    }
    __syncthreads();

    // byte offset (from kSmemBinsAbs) of this lane's heat word for magic-domain shell bits sb:
    // lane-private: shell * 256 + lane * 4 by one PRMT; plain: (sb << 2) + bias
    const uint32_t lane_low = lane * 4u;
    const uint32_t plain_bias = 0u - (kMagicBits << 2);
    const uint32_t heat2_off = plain_bins * 4u;            // plain layout; lane-private: 128
    // plain layout: lane l clamps to slot SHELLS-1+l, so the overflow bin never serialises a warp
    const uint32_t clamp_bits = LANE_PRIVATE ? a.last_bits : a.last_bits + lane;
    // byte offsets of this lane's copies inside a direction-table row
    const uint32_t polar_low = (lane & 15u) * 8u, azimuth_low = polar_low + 128u;

    // Philox round 0 with the counter folded in: c0 = first_lo + rel, c1 = first_hi, c3 = 0.
    //   M0 * c0 = M0 * rel + M0 * first_lo   (no wrap: the launch stays inside one 2^32 window)
    const uint64_t m0_first = static_cast<uint64_t>(kPhiloxM0) * static_cast<uint32_t>(a.first);
    const uint32_t c1k0 = static_cast<uint32_t>(a.first >> 32) ^ a.keys.k[0];

    // PPL photons per lane: position (mean-free-path units, photon.c:12-14), photon offset from
    // a.first, roulette fate word, "this slot holds a photon"
    float px[PPL], py[PPL], pz[PPL];
    uint32_t rel[PPL], fate[PPL];
    bool act[PPL], surv[PPL];
    uint32_t r[PPL][4] = {};                // the Philox block of each photon whose events are running
#pragma unroll
    for (int j = 0; j < PPL; ++j) px[j] = py[j] = pz[j] = 0.0f;

    unsigned long long n_events = 0ull;     // warp-uniform
    uint32_t range_flag = 0u;
    uint32_t blocks_since_drain = 0u;       // warp-uniform

    // Drain one slice (1 / WARPS) of the block histograms into the global u64 tallies.  Slices
    // are handed out round-robin by a block-wide ticket, so every slice is drained once per
    // WARPS calls no matter which warps are still walking (a warp that has run out of photons
    // stops calling).  Out of line: the walk loop keeps its registers.
    auto drain = [&]() {
        uint32_t ticket = 0u;
        if (lane == 0u) ticket = atomicAdd(&drain_ticket, 1u);
        range_flag |= drain_slice<BLOCK, LANE_PRIVATE>(bins, a.tallies, a.shells,
                                                       __shfl_sync(0xffffffffu, ticket, 0) % WARPS, a.check_shift);
    };

    // drop (photon.c:26-29): shell = min(trunc(|r| * shells_per_mfp), SHELLS-1).  Positions are in units of the
    // grid radius, so |r|^2 saturated to 1 by the .sat of its last FFMA IS the clamp to the overflow bin (every
    // |r| >= 1 lands in shell SHELLS-1); trunc(|r| * shell_scale) without F2I: add 2^23 rounding toward zero, the
    // mantissa is the integer.  The one-histogram layout clamps the raw bits instead, to per-lane overflow slots
    // (unless SAT_PLAIN).
    auto radius_sq = [&](float x, float y, float z) {
        const float r2 = fmaf(z, z, fmaf(y, y, x * x));
        return SAT_CLAMP ? __saturatef(r2) : r2;
    };
    // spin + hop (photon.c:35-43 sampled directly, photon.c:22-24): L = log2(xi) <= 0, pol = kappa * (-ln2 cos, -ln2 sin)(theta),
    // azi = (cos, sin)(phi); the step t = -ln2 * L is folded into the polar entry
    auto move = [&](float L, float2 pol, float2 azi, float& x, float& y, float& z) {
        const float ts = L * pol.y;
        x = fmaf(L, pol.x, x);
        if constexpr ((kPacked & 1) != 0) {
            fma2_scalar(ts, azi, y, z);        // one FFMA2
        } else {
            y = fmaf(ts, azi.x, y);
            z = fmaf(ts, azi.y, z);
        }
    };
    auto shell_bits = [&](float rad) {
        const uint32_t sb = __float_as_uint(__fmaf_rz(rad, a.shell_scale, 8388608.0f));
        return SAT_CLAMP ? sb : min(sb, clamp_bits);
    };

    // The lane's PPL photons share packed instructions in pairs (2p, 2p+1); an odd last photon goes scalar.
    // hop (photon.c:21): L = log2(xi) <= 0 for the event word of every photon; the step t = -ln2 * L is folded into the polar table
    auto log2_xi = [&](const uint32_t (&w)[PPL], float (&L)[PPL]) {
#pragma unroll
        for (int j = 0; j < PPL; j += 2) {
            const float f0 = __uint_as_float(__funnelshift_r(w[j], 0x7Fu, 9));          // 0x3F800000 | m: 1 + m 2^-23
            if ((kPacked & 2) && j + 1 < PPL) {                                         // xi of both photons by one FADD2
                float xi0, xi1;
                add2_scalar(f0, __uint_as_float(__funnelshift_r(w[j + 1], 0x7Fu, 9)), -kOneMinusHalfUlp, xi0, xi1);
                L[j] = mufu_lg2(xi0);
                L[j + 1] = mufu_lg2(xi1);
            } else {
                L[j] = mufu_lg2(f0 - kOneMinusHalfUlp);
                if (j + 1 < PPL) L[j + 1] = mufu_lg2(__uint_as_float(__funnelshift_r(w[j + 1], 0x7Fu, 9)) - kOneMinusHalfUlp);
            }
        }
    };
    // drop (photon.c:26-29): byte offset of the lane's heat word for every photon's (clamped) radius
    auto tally_offsets = [&](const float (&rad)[PPL], uint32_t (&off)[PPL]) {
#pragma unroll
        for (int j = 0; j < PPL; j += 2) {
            uint32_t sb0, sb1 = 0u;
            if ((kPacked & 4) && j + 1 < PPL) {                                         // both shell numbers by one FFMA2.RZ
                fma2_rz_bits(rad[j], rad[j + 1], a.shell_scale, 8388608.0f, sb0, sb1);
                if constexpr (!SAT_CLAMP) {
                    sb0 = min(sb0, clamp_bits);
                    sb1 = min(sb1, clamp_bits);
                }
            } else {
                sb0 = shell_bits(rad[j]);
                if (j + 1 < PPL) sb1 = shell_bits(rad[j + 1]);
            }
            off[j] = LANE_PRIVATE ? shell_offset(sb0, lane_low) : (sb0 << 2) + plain_bias;
            if (j + 1 < PPL) off[j + 1] = LANE_PRIVATE ? shell_offset(sb1, lane_low) : (sb1 << 2) + plain_bias;
        }
    };
    // spin + hop of one photon for one event; returns the radius the tally sees
    auto walk_one = [&](float L, float2 pol, float2 azi, int j) {
        if constexpr (RADIAL) {
            // px holds the radius: r'^2 = r^2 + t^2 + 2 r (t mu), >= 0 up to rounding
            const float t = L * a.radial_step, tmu = L * pol.x;
            const float r2 = fmaf(px[j] + px[j], tmu, fmaf(t, t, px[j] * px[j]));
            px[j] = mufu_sqrt(fmaxf(r2, 0.0f));      // the state keeps the true radius,
            return fminf(px[j], 1.0f);               // the tally sees it clamped to the grid (photon.c:27-29)
        } else {
            move(L, pol, azi, px[j], py[j], pz[j]);
            return mufu_sqrt(radius_sq(px[j], py[j], pz[j]));
        }
    };

    // One scatter event (reference photon.c:21-43) from word S of the current Philox block for
    // every photon of the lane: spin, hop, drop.  `dep` / `dep2` are the warp-uniform deposit
    // (1-albedo) * w and its rescaled square.  Bits of the event word v:
    //    9..31  step:   xi = (m + 1/2) 2^-23, m = v >> 9 (midpoint rule; never 0 or 1)
    //    8..15  polar:  cos(theta) = (2k + 1)/256 - 1, sin(theta) from the table
    //    0..7   azimuth: phi = 2 pi k / 256, (cos, sin) from the table
    // (the polar index shares bits 9..15 with the step, where they only decide the step's last
    // seven mantissa bits: the step keeps a 23-bit marginal, the coupling is below 2^-16 in xi)
    auto event = [&](auto slot_tag, auto partial_tag, uint32_t dep, uint32_t dep2) {
        constexpr int S = decltype(slot_tag)::value;
        constexpr bool PARTIAL = decltype(partial_tag)::value;
        uint32_t w[PPL], off[PPL];
        float L[PPL], rad[PPL];
#pragma unroll
        for (int j = 0; j < PPL; ++j) w[j] = r[j][S];
        log2_xi(w, L);
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const float2 pol = lds_f32x2<kSmemTableAbs>(row_offset<1>(w[j], polar_low));
            float2 azi = pol;
            if constexpr (!RADIAL) azi = lds_f32x2<kSmemTableAbs>(row_offset<0>(w[j], azimuth_low));
            rad[j] = walk_one(L[j], pol, azi, j);
        }
        tally_offsets(rad, off);
#pragma unroll
        for (int j = 0; j < PPL; ++j)
            if (!PARTIAL || act[j]) {
#if TMC_EXPERIMENT == 1   /* one atomic instead of two (wrong tallies; timing experiment only) */
                red_shared_add<kSmemBinsAbs>(off[j], dep + dep2);
#else
                red_shared_add<kSmemBinsAbs>(off[j], dep);
                if constexpr (LANE_PRIVATE) red_shared_add<kSmemBinsAbs + 128u>(off[j], dep2);
                else red_shared_add<kSmemBinsAbs>(off[j] + heat2_off, dep2);
#endif
            }
    };

    // Walk generation g for the photons held by the warp, then play roulette (photon.c:45-49).
    auto phase = [&](auto partial_tag, uint32_t g) {
        constexpr bool PARTIAL = decltype(partial_tag)::value;
        const uint32_t ev_first = a.gen[g].first_event;
        const uint32_t ev_last = ev_first + a.gen[g].n_events - 1u;
        const uint32_t b_first = ev_first / kEventsPerBlock, b_last = ev_last / kEventsPerBlock;
        // The warp-uniform weight schedule: event e deposits (1-albedo) * w(e), rounded, and its
        // rescaled square; both come from a table the host computed with the exact integer
        // recurrence (tmc_api.cu: deposit_table), one broadcast load per event.
        const uint2* next_deposit = a.deposits + ev_first;
        uint32_t dep = 0u, dep2 = 0u;
        auto absorb = [&]() {
            const uint2 d = __ldg(next_deposit++);
            dep = d.x;
            dep2 = d.y;
        };
        // One Philox block per photon, counter = (photon index, b), into `out`.  The walk loop
        // draws block b+1 while the events of block b run: the Philox rounds (IMAD.WIDE on the
        // FMA-heavy pipe, LOP3) and the event arithmetic (FP32, MUFU, shared memory) use
        // different pipes and have no data dependence, so ptxas interleaves them.
        auto draw = [&](uint32_t b, uint32_t (&out)[PPL][4]) {
            const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * b;
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const uint64_t p0 = mad_wide(rel[j], kPhiloxM0, m0_first);
                philox4x32_rounds<1, ROUNDS>(a.keys, static_cast<uint32_t>(p1 >> 32) ^ c1k0, static_cast<uint32_t>(p1),
                                             static_cast<uint32_t>(p0 >> 32) ^ a.keys.k[1], static_cast<uint32_t>(p0), out[j]);
            }
        };
        auto take_block = [&](uint32_t (&cur)[PPL][4]) {
#pragma unroll
            for (int j = 0; j < PPL; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) r[j][k] = cur[j][k];
        };
        // A full block: its four events for every photon of the lane, software-pipelined in SOURCE order.
        // ptxas may not move a shared-memory load (direction table) above an earlier shared-memory atomic
        // (tallies): it cannot know that the two never alias.  Written event by event, every table look-up
        // would therefore wait for the previous event's atomics, and the events of a warp would run as one
        // serial chain (SHF -> FADD -> MUFU -> FMUL -> FFMA x5 -> MUFU -> FFMA.RZ -> VIMNMX -> PRMT -> ATOMS,
        // ~60 cycles of fixed latencies each).  So: all look-ups and logarithms of a group of TMC_GROUP events
        // first (independent of each other), then the position chain and the radii, then the atomics; the
        // next group's look-ups are issued BEFORE this group's atomics.
        auto four_events = [&](uint32_t (&cur)[PPL][4]) {
            take_block(cur);
            constexpr int G = TMC_GROUP;                     // events per group: 1, 2 or 4
            float L[G][PPL];
            float2 pol[G][PPL];
            [[maybe_unused]] float2 azi[G][PPL];
            uint32_t off[G][PPL];
            auto look_up = [&](auto first_tag) {
                constexpr int S0 = decltype(first_tag)::value;
#pragma unroll
                for (int e = 0; e < G; ++e) {
                    uint32_t w[PPL];
#pragma unroll
                    for (int j = 0; j < PPL; ++j) w[j] = r[j][S0 + e];
                    log2_xi(w, L[e]);
#pragma unroll
                    for (int j = 0; j < PPL; ++j) {
                        pol[e][j] = lds_f32x2<kSmemTableAbs>(row_offset<1>(w[j], polar_low));
                        if constexpr (!RADIAL) azi[e][j] = lds_f32x2<kSmemTableAbs>(row_offset<0>(w[j], azimuth_low));
                    }
                }
            };
            auto walk_group = [&]() {
#pragma unroll
                for (int e = 0; e < G; ++e) {
                    float rad[PPL];
#pragma unroll
                    for (int j = 0; j < PPL; ++j) rad[j] = walk_one(L[e][j], pol[e][j], RADIAL ? pol[e][j] : azi[e][j], j);
                    tally_offsets(rad, off[e]);
                }
            };
            auto tally_group = [&]() {
#pragma unroll
                for (int e = 0; e < G; ++e) {
                    absorb();
#pragma unroll
                    for (int j = 0; j < PPL; ++j) {
                        if (!PARTIAL || act[j]) {
                            red_shared_add<kSmemBinsAbs>(off[e][j], dep);
                            if constexpr (LANE_PRIVATE) red_shared_add<kSmemBinsAbs + 128u>(off[e][j], dep2);
                            else red_shared_add<kSmemBinsAbs>(off[e][j] + heat2_off, dep2);
                        }
                    }
                }
            };
            look_up(IntTag<0>{});
            if constexpr (G == 4) {
                walk_group();
                // the four deposits of a full block are 32 contiguous, 32-byte aligned bytes of the table
                const uint4 d01 = __ldg(reinterpret_cast<const uint4*>(next_deposit));
                const uint4 d23 = __ldg(reinterpret_cast<const uint4*>(next_deposit) + 1);
                next_deposit += 4;
                const uint32_t dp[4] = { d01.x, d01.z, d23.x, d23.z }, dp2[4] = { d01.y, d01.w, d23.y, d23.w };
#pragma unroll
                for (int e = 0; e < G; ++e)
#pragma unroll
                    for (int j = 0; j < PPL; ++j)
                        if (!PARTIAL || act[j]) {
                            red_shared_add<kSmemBinsAbs>(off[e][j], dp[e]);
                            if constexpr (LANE_PRIVATE) red_shared_add<kSmemBinsAbs + 128u>(off[e][j], dp2[e]);
                            else red_shared_add<kSmemBinsAbs>(off[e][j] + heat2_off, dp2[e]);
                        }
            } else if constexpr (G == 2) {
                walk_group();
                uint32_t off0[G][PPL];
#pragma unroll
                for (int e = 0; e < G; ++e)
#pragma unroll
                    for (int j = 0; j < PPL; ++j) off0[e][j] = off[e][j];
                look_up(IntTag<2>{});                    // before the atomics of events 0-1
#pragma unroll
                for (int e = 0; e < G; ++e) {
                    absorb();
#pragma unroll
                    for (int j = 0; j < PPL; ++j)
                        if (!PARTIAL || act[j]) {
                            red_shared_add<kSmemBinsAbs>(off0[e][j], dep);
                            if constexpr (LANE_PRIVATE) red_shared_add<kSmemBinsAbs + 128u>(off0[e][j], dep2);
                            else red_shared_add<kSmemBinsAbs>(off0[e][j] + heat2_off, dep2);
                        }
                }
                walk_group();
                tally_group();
            } else {
                static_assert(G == 2 || G == 4, "TMC_GROUP must be 2 or 4");
            }
        };
        auto maybe_drain = [&](uint32_t blocks) {
            blocks_since_drain += blocks;
            if (blocks_since_drain >= a.flush_blocks) {
                drain();
                blocks_since_drain = 0u;
            }
        };
        // first / last block of the generation (in `cur`): only words s_lo .. s_hi belong to it
        auto ragged_block = [&](uint32_t b, uint32_t (&cur)[PPL][4], uint32_t s_lo, uint32_t s_hi) {
            take_block(cur);
            if (g == 0u && b == 0u) {            // pseudo-event 0 of a fresh photon: its fate word
#pragma unroll
                for (int j = 0; j < PPL; ++j) fate[j] = r[j][0];
            }                                    // (later generations that start inside block 0 carry their fate)
            if (s_lo == 0u) { absorb(); event(IntTag<0>{}, partial_tag, dep, dep2); }
            if (s_lo <= 1u && s_hi >= 1u) { absorb(); event(IntTag<1>{}, partial_tag, dep, dep2); }
            if (s_lo <= 2u && s_hi >= 2u) { absorb(); event(IntTag<2>{}, partial_tag, dep, dep2); }
            if (s_hi == 3u) { absorb(); event(IntTag<3>{}, partial_tag, dep, dep2); }
            maybe_drain(1u);
        };
        const uint32_t s_first = ev_first - kEventsPerBlock * b_first, s_last = ev_last - kEventsPerBlock * b_last;
        uint32_t ra[PPL][4], rb[PPL][4];        // ping-pong: the block being walked / the next one
        draw(b_first, ra);
        if (b_first == b_last) {
            ragged_block(b_first, ra, s_first, s_last);
        } else {
            draw(b_first + 1u, rb);
            ragged_block(b_first, ra, s_first, kEventsPerBlock - 1u);
            // full blocks b_first+1 .. b_last-1, the current one in rb.  Chunks of at most
            // flush_blocks blocks run without a call, so the Philox keys stay in uniform registers.
            uint32_t b = b_first + 1u;
            while (b < b_last) {
                uint32_t chunk = min(b_last - b, a.flush_blocks - min(blocks_since_drain, a.flush_blocks - 1u));
                const uint32_t done = chunk;
                for (; chunk >= 2u; chunk -= 2u, b += 2u) {
                    draw(b + 1u, ra);
                    four_events(rb);
                    draw(b + 2u, rb);
                    four_events(ra);
                }
                if (chunk) {
                    draw(b + 1u, ra);
                    four_events(rb);
#pragma unroll
                    for (int j = 0; j < PPL; ++j)
#pragma unroll
                        for (int k = 0; k < 4; ++k) rb[j][k] = ra[j][k];
                    ++b;
                }
                maybe_drain(done);
            }
            ragged_block(b_last, rb, 0u, s_last);
        }
        uint32_t n_act = COHORT;
        if constexpr (PARTIAL) {
            n_act = 0u;
#pragma unroll
            for (int j = 0; j < PPL; ++j) n_act += __popc(__ballot_sync(0xffffffffu, act[j]));
        }
        n_events += static_cast<unsigned long long>(n_act) * a.gen[g].n_events;
        // roulette: the fate word is a uniform 32-bit integer; surviving (probability 0.1)
        // multiplies it by 10, which is uniform again.  The x10 weight boost is in gen[g+1].
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            surv[j] = (!PARTIAL || act[j]) && fate[j] < kFateSurvive;
            if (surv[j]) fate[j] *= 10u;
        }
    };

    // static cohort -> warp map: warp W owns cohorts W, W + total_warps, ... of COHORT photons each
    const uint32_t total_warps = gridDim.x * WARPS;
    const uint32_t count = static_cast<uint32_t>(a.count);
    const uint32_t n_cohorts = (count + COHORT - 1u) / COHORT;
    uint32_t cohort = blockIdx.x * WARPS + wid;
    uint32_t nq[kQueuedGens] = {};          // fill of this warp's survivor queues (generations 1 .. kQueuedGens)

    for (;;) {
        // a full cohort of parked survivors (deepest generation first), else fresh photons, else the leftovers
        // (shallowest first: their survivors feed the deeper queues)
        static_assert(kQueuedGens == 2u || kQueuedGens == 3u, "TMC_QUEUED_GENS must be 2 or 3");
        uint32_t g, take;
        if (kQueuedGens >= 3u && nq[kQueuedGens - 1u] >= COHORT) { g = 3u; take = COHORT; }
        else if (nq[1] >= COHORT) { g = 2u; take = COHORT; }
        else if (nq[0] >= COHORT) { g = 1u; take = COHORT; }
        else if (cohort < n_cohorts) { g = 0u; take = min(COHORT, count - cohort * COHORT); }
        else if (nq[0] > 0u) { g = 1u; take = nq[0]; }
        else if (nq[1] > 0u) { g = 2u; take = nq[1]; }
        else if (kQueuedGens >= 3u && nq[kQueuedGens - 1u] > 0u) { g = 3u; take = nq[kQueuedGens - 1u]; }
        else break;

        if (g == 0u) {                      // regenerate: a cohort of fresh photons at the origin
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                rel[j] = cohort * COHORT + lane * PPL + j;
                px[j] = py[j] = pz[j] = 0.0f;
                fate[j] = 0u;
                act[j] = lane * PPL + j < take;
            }
            cohort += total_warps;
        } else {                            // a cohort of parked survivors of generation g
            // one generation's queue: uint4 (x, y, z, photon offset)[kQueueCap], then the fate words
            const uint32_t* q = queue + (g - 1u) * (kQueueFields * kQueueCap);
            const uint32_t start = nq[g - 1u] - take;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const uint32_t e = lane * PPL + j;
                act[j] = e < take;
                const uint32_t at = act[j] ? start + e : 0u;
                const uint4 e4 = reinterpret_cast<const uint4*>(q)[at];
                px[j] = __uint_as_float(e4.x);
                py[j] = __uint_as_float(e4.y);
                pz[j] = __uint_as_float(e4.z);
                rel[j] = e4.w;
                fate[j] = q[4u * kQueueCap + at];
            }
            __syncwarp();
            nq[g - 1u] = start;
        }

        bool partial = take < COHORT;
        for (;;) {
            if (partial) phase(BoolTag<true>{}, g);
            else phase(BoolTag<false>{}, g);
            if (g + 1u >= a.n_gen) break;
            if (g + 1u <= kQueuedGens) {    // park the survivors for a later full cohort
                uint32_t* q = queue + g * (kQueueFields * kQueueCap);
#pragma unroll
                for (int j = 0; j < PPL; ++j) {
                    const uint32_t m = __ballot_sync(0xffffffffu, surv[j]);
                    const uint32_t at = nq[g] + __popc(m & lanemask_lt());
                    if (surv[j]) {
                        reinterpret_cast<uint4*>(q)[at] = make_uint4(__float_as_uint(px[j]), __float_as_uint(py[j]), __float_as_uint(pz[j]), rel[j]);
                        q[4u * kQueueCap + at] = fate[j];
                    }
                    nq[g] += __popc(m);
                }
                __syncwarp();
                break;
            }
            // deeper generations (1e-4 of the photons): the survivors continue in place
            bool any = false;
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                any = any || surv[j];
                act[j] = surv[j];
            }
            if (!__any_sync(0xffffffffu, any)) break;
            partial = true;
            ++g;
        }
    }

    // Final drain once every warp of the block is done.
    __syncthreads();
    range_flag |= drain_slice<BLOCK, LANE_PRIVATE>(bins, a.tallies, a.shells, wid, a.check_shift);
    uint32_t fl = range_flag;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    if (lane == 0u) {
        atomicAdd(&a.counters[0], n_events);
        if (fl != 0u) atomicOr(&a.counters[2], 1ull);
    }
    if (tid == 0u && blockIdx.x == 0u) atomicAdd(&a.counters[1], static_cast<unsigned long long>(a.count));
}

#endif  // __CUDACC__

}  // namespace tmc
