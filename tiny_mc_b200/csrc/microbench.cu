// microbench.cu — issue-rate micro-benchmarks for the pipes the photon walk lives on.
//
// MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only; this path is bound by the FP32/INT
// issue rate, the MUFU (XU) pipe and the shared-memory atomic unit.  This program
// measures those denominators on the actual B200: warp-instructions per clock per SM for each
// SASS opcode class the kernel uses, for the mixes it issues (FMA pipe + ALU pipe), for a whole
// Philox4x32-10 block, and for ATOMS.ADD under the address patterns of the three tally regimes.
//
// Output: one JSON object per line on stdout:
//   {"test": "...", "warp_inst_per_clk_per_sm": x, "lanes_per_clk_per_sm": 32x, "sm_mhz": f, ...}
// Cycle counts come from clock64() inside the kernel (mean over blocks), so they are
// independent of DVFS; the MHz figure is cycles / CUDA-event time.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                           \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                    \
            exit(2);                                                                    \
        }                                                                               \
    } while (0)

constexpr int kBlock = 256;
constexpr int kIters = 16384;  // outer loop trips (long enough that launch latency and clock ramp vanish)
constexpr int kChains = 8;     // independent dependency chains per thread

struct Out {
    long long cycles;
    unsigned sink;
};

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

// ---- generic harness: BODY is a statement list executed kIters times -----------------------
#define BENCH_KERNEL(NAME, DECLS, BODY, SINK)                                            \
    __global__ void __launch_bounds__(kBlock) NAME(Out* out, unsigned seed)              \
    {                                                                                    \
        DECLS;                                                                           \
        __syncthreads();                                                                 \
        const long long t0 = clock64();                                                  \
        _Pragma("unroll 1") for (int it = 0; it < kIters; ++it) { BODY; }                \
        const long long t1 = clock64();                                                  \
        __syncthreads();                                                                 \
        if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;                          \
        if ((SINK) == 0x12345u) out[blockIdx.x].sink = 1;                                \
    }

#define F_DECL float f0 = seed * 1e-9f + threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7; const float ca = 1.0001f + seed * 1e-12f, cb = 0.5f
#define U_DECL unsigned u0 = seed + threadIdx.x, u1 = u0 * 3 + 1, u2 = u0 * 5 + 2, u3 = u0 * 7 + 3, u4 = u0 * 11 + 4, u5 = u0 * 13 + 5, u6 = u0 * 17 + 6, u7 = u0 * 19 + 7; const unsigned ka = seed | 0xD2511F53u, kb = seed ^ 0x9E3779B9u
#define F_SINK __float_as_uint(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7)
#define U_SINK (u0 ^ u1 ^ u2 ^ u3 ^ u4 ^ u5 ^ u6 ^ u7)

#define FFMA_(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f##i) : "f"(ca), "f"(cb));
BENCH_KERNEL(k_ffma, F_DECL, REP8(FFMA_) REP8(FFMA_), F_SINK)

#define FMUL_(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f##i) : "f"(ca));
BENCH_KERNEL(k_fmul, F_DECL, REP8(FMUL_) REP8(FMUL_), F_SINK)

#define FADD_(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f##i) : "f"(cb));
BENCH_KERNEL(k_fadd, F_DECL, REP8(FADD_) REP8(FADD_), F_SINK)

// packed FP32x2 (Blackwell): two FMAs per lane per instruction
#define FFMA2_DECL F_DECL; unsigned long long p0, p1, p2, p3, pc, pd; \
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(f0), "f"(f1)); \
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(f2), "f"(f3)); \
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(f4), "f"(f5)); \
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(f6), "f"(f7)); \
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(ca)); \
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pd) : "f"(cb))
#define FFMA2_(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p##i) : "l"(pc), "l"(pd));
#define REP4(X) X(0) X(1) X(2) X(3)
BENCH_KERNEL(k_ffma2, FFMA2_DECL, REP4(FFMA2_) REP4(FFMA2_) REP4(FFMA2_) REP4(FFMA2_), (unsigned)(p0 ^ p1 ^ p2 ^ p3))

#define IMADW_DECL U_DECL; unsigned long long w0 = u0, w1 = u1, w2 = u2, w3 = u3, w4 = u4, w5 = u5, w6 = u6, w7 = u7
#define IMADW_(i) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(w##i) : "r"(ka));
BENCH_KERNEL(k_imad_wide, IMADW_DECL, REP8(IMADW_) REP8(IMADW_), (unsigned)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7))

#define IMAD_(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u##i) : "r"(ka), "r"(kb));
BENCH_KERNEL(k_imad, U_DECL, REP8(IMAD_) REP8(IMAD_), U_SINK)

#define IMADHI_(i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(u##i) : "r"(ka), "r"(kb));
BENCH_KERNEL(k_imad_hi, U_DECL, REP8(IMADHI_) REP8(IMADHI_), U_SINK)

#define LOP3_(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u##i) : "r"(ka), "r"(kb));
BENCH_KERNEL(k_lop3, U_DECL, REP8(LOP3_) REP8(LOP3_), U_SINK)

#define IADD3_(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(u##i) : "r"(ka));
BENCH_KERNEL(k_iadd, U_DECL, REP8(IADD3_) REP8(IADD3_), U_SINK)

#define SHF_(i) asm volatile("shf.r.wrap.b32 %0, %0, %1, 9;" : "+r"(u##i) : "r"(ka));
BENCH_KERNEL(k_shf, U_DECL, REP8(SHF_) REP8(SHF_), U_SINK)

#define PRMT_(i) asm volatile("prmt.b32 %0, %0, %1, 0x7610;" : "+r"(u##i) : "r"(ka));
BENCH_KERNEL(k_prmt, U_DECL, REP8(PRMT_) REP8(PRMT_), U_SINK)

#define ISETP_(i) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, 1; }" : "+r"(u##i) : "r"(ka));
BENCH_KERNEL(k_isetp_padd, U_DECL, REP8(ISETP_) REP8(ISETP_), U_SINK)

#define I2FP_(i) asm volatile("{ .reg .f32 t; cvt.rn.f32.u32 t, %0; mov.b32 %0, t; }" : "+r"(u##i));
BENCH_KERNEL(k_i2fp, U_DECL, REP8(I2FP_) REP8(I2FP_), U_SINK)

#define F2I_(i) asm volatile("{ .reg .u32 t; cvt.rzi.u32.f32 t, %0; mov.b32 %0, t; }" : "+f"(f##i));
BENCH_KERNEL(k_f2i, F_DECL, REP8(F2I_) REP8(F2I_), F_SINK)

#define LG2_(i) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f##i));
BENCH_KERNEL(k_mufu_lg2, F_DECL, REP8(LG2_) REP8(LG2_), F_SINK)
#define SQRT_(i) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f##i));
BENCH_KERNEL(k_mufu_sqrt, F_DECL, REP8(SQRT_) REP8(SQRT_), F_SINK)
#define RSQ_(i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f##i));
BENCH_KERNEL(k_mufu_rsq, F_DECL, REP8(RSQ_) REP8(RSQ_), F_SINK)
#define SIN_(i) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f##i));
BENCH_KERNEL(k_sin_approx, F_DECL, REP8(SIN_) REP8(SIN_), F_SINK)
#define EX2_(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f##i));
BENCH_KERNEL(k_mufu_ex2, F_DECL, REP8(EX2_) REP8(EX2_), F_SINK)

// mixes: FMA pipe + ALU pipe, FMA pipe + MUFU
#define MIX_DECL F_DECL; U_DECL
#define MIX_FL_(i) FFMA_(i) LOP3_(i)
BENCH_KERNEL(k_mix_ffma_lop3, MIX_DECL, REP8(MIX_FL_), F_SINK ^ U_SINK)
#define MIX_IL_DECL IMADW_DECL
#define MIX_IL_(i) IMADW_(i) LOP3_(i)
BENCH_KERNEL(k_mix_imadw_lop3, MIX_IL_DECL, REP8(MIX_IL_), U_SINK ^ (unsigned)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7))
#define MIX_FI_DECL F_DECL; IMADW_DECL
#define MIX_FI_(i) FFMA_(i) IMADW_(i)
BENCH_KERNEL(k_mix_ffma_imadw, MIX_FI_DECL, REP8(MIX_FI_), F_SINK ^ (unsigned)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7))
// packed FP32 against the integer multiplier: do FFMA2 and IMAD.WIDE share the FMA-heavy pipe?
#define MIX_F2I_DECL FFMA2_DECL; unsigned long long w0 = u0x, w1 = w0 + 1, w2 = w0 + 2, w3 = w0 + 3; const unsigned ka = seed | 0xD2511F53u
#define IMADW4_(i) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(w##i) : "r"(ka));
#define MIX_F2I_(i) FFMA2_(i) IMADW4_(i)
__global__ void __launch_bounds__(kBlock) k_mix_ffma2_imadw(Out* out, unsigned seed)
{
    const unsigned u0x = seed + threadIdx.x;
    MIX_F2I_DECL;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) { REP4(MIX_F2I_) REP4(MIX_F2I_) }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
    if ((unsigned)(p0 ^ p1 ^ p2 ^ p3 ^ w0 ^ w1 ^ w2 ^ w3) == 0x12345u) out[blockIdx.x].sink = 1;
}
// scalar FMUL against the integer multiplier (FMUL may use the FMA-lite pipe)
#define MIX_MI_(i) FMUL_(i) IMADW_(i)
BENCH_KERNEL(k_mix_fmul_imadw, MIX_FI_DECL, REP8(MIX_MI_), F_SINK ^ (unsigned)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7))
// 2 scalar FFMA per IMAD.WIDE
#define MIX_2FI_(i) FFMA_(i) FFMA_(i) IMADW_(i)
BENCH_KERNEL(k_mix_2ffma_imadw, MIX_FI_DECL, REP8(MIX_2FI_), F_SINK ^ (unsigned)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7))
// FFMA2 + LOP3
#define MIX_F2L_DECL FFMA2_DECL; U_DECL
#define MIX_F2L_(i) FFMA2_(i) LOP3_(i)
BENCH_KERNEL(k_mix_ffma2_lop3, MIX_F2L_DECL, REP4(MIX_F2L_) REP4(MIX_F2L_), (unsigned)(p0 ^ p1 ^ p2 ^ p3) ^ U_SINK)
// Does IMAD.WIDE.U32 take one dispatch slot or two?  Philox-shaped true wide products (both halves live:
// x = hi(x * M) ^ lo(x * M) ^ k), each group balanced over the three pipes: 4 FMA-heavy cycles (one
// IMAD.WIDE, or two 32-bit IMADs in the reference kernel), 4 ALU cycles (2 LOP3), 4 FMA-lite cycles (2 FMUL).
#define WIDE_DECL F_DECL; U_DECL
#define WIDE_STEP_(i) asm volatile("{ .reg .u64 p; .reg .u32 lo, hi; mul.wide.u32 p, %0, 0xD2511F53; mov.b64 {lo, hi}, p; lop3.b32 %0, lo, hi, %1, 0x96; }" : "+r"(u##i) : "r"(kb));
#define NARROW_STEP_(i) asm volatile("{ .reg .u32 a, b; mad.lo.u32 a, %0, 0xD2511F53, %1; mad.lo.u32 b, %0, 0xCD9E8D57, %1; lop3.b32 %0, a, b, %1, 0x96; }" : "+r"(u##i) : "r"(kb));
#define EXTRA_LOP_(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u##i) : "r"(ka), "r"(kb));
#define WIDE_GROUP_(i) WIDE_STEP_(i) EXTRA_LOP_(i) FMUL_(i) FMUL_(i)
#define NARROW_GROUP_(i) NARROW_STEP_(i) EXTRA_LOP_(i) FMUL_(i) FMUL_(i)
BENCH_KERNEL(k_dispatch_imad_wide_group, WIDE_DECL, REP8(WIDE_GROUP_), F_SINK ^ U_SINK)
BENCH_KERNEL(k_dispatch_two_imad_group, WIDE_DECL, REP8(NARROW_GROUP_), F_SINK ^ U_SINK)
// the same with nothing but the wide products and their LOP3 (heavy-pipe bound: 4 cycles each)
BENCH_KERNEL(k_imad_wide_live, U_DECL, REP8(WIDE_STEP_) REP8(WIDE_STEP_), U_SINK)

// 7 FFMA : 1 MUFU
#define MIX_FM_(i) FFMA_(i) FFMA_(i)
BENCH_KERNEL(k_mix_14ffma_2mufu, F_DECL, REP8(MIX_FM_) LG2_(0) SQRT_(1), F_SINK)
// FFMA : LOP3 : MUFU = 8 : 8 : 2 (roughly the walk's shape)
BENCH_KERNEL(k_mix_8ffma_8lop3_2mufu, MIX_DECL, REP8(MIX_FL_) LG2_(0) SQRT_(1), F_SINK ^ U_SINK)

// whole Philox4x32-10 blocks, 2 independent counters per thread
__device__ __forceinline__ void philox10(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ (k0 + r * 0x9E3779B9u);
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ (k1 + r * 0xBB67AE85u);
        c1 = (unsigned)p1; c3 = (unsigned)p0; c0 = n0; c2 = n2;
    }
}
__global__ void __launch_bounds__(kBlock) k_philox10(Out* out, unsigned seed)
{
    unsigned a0 = threadIdx.x, a1 = seed, a2 = 1, a3 = 0, b0 = threadIdx.x + 7777, b1 = seed, b2 = 2, b3 = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        philox10(a0, a1, a2, a3, seed, seed ^ 0x55u);
        philox10(b0, b1, b2, b3, seed, seed ^ 0x55u);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
    if ((a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3) == 0x12345u) out[blockIdx.x].sink = 1;
}

// shared-memory atomics: MODE 0 conflict-free (bank = lane), 1 random over `nb` bins,
// 2 one address per warp-instruction (all 32 lanes collide), 3 random, 20 % of lanes on one hot bin
template <int MODE, bool RETURN>
__global__ void __launch_bounds__(kBlock) k_atoms(Out* out, unsigned seed, int nb)
{
    extern __shared__ unsigned bins[];
    for (int i = threadIdx.x; i < nb + 32; i += kBlock) bins[i] = 0;
    unsigned s = seed * 2654435761u + (blockIdx.x * kBlock + threadIdx.x) * 40503u + 1u;
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s = s * 1664525u + 1013904223u;
            unsigned idx;
            if (MODE == 0) idx = (threadIdx.x & 31) + ((s >> 27) & ~31u) % (unsigned)(nb > 32 ? nb - 32 : 1);
            else if (MODE == 1) idx = __umulhi(s, (unsigned)nb);
            else if (MODE == 2) idx = (unsigned)(it & 63);
            else idx = (s & 0xF0000000u) < 0x30000000u ? (unsigned)nb - 1 : __umulhi(s, (unsigned)nb);
            if (RETURN) acc += atomicAdd(&bins[idx], s >> 12);
            else asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bins[idx])), "r"(s >> 12) : "memory");
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
    if ((acc ^ bins[threadIdx.x % nb]) == 0x12345u) out[blockIdx.x].sink = 1;
}
// same address arithmetic without the atomic, to subtract its cost
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_atoms_baseline(Out* out, unsigned seed, int nb)
{
    unsigned s = seed * 2654435761u + (blockIdx.x * kBlock + threadIdx.x) * 40503u + 1u;
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s = s * 1664525u + 1013904223u;
            unsigned idx = (MODE == 1) ? __umulhi(s, (unsigned)nb) : (unsigned)(it & 63);
            acc += idx ^ (s >> 12);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
    if (acc == 0x12345u) out[blockIdx.x].sink = 1;
}

// The walk's own instruction mix with perfect instruction-level parallelism: per "event" 1 SHF, 1 FADD, 2 MUFU,
// 3 PRMT, 2 conflict-free LDS.64, 2 FMUL, 6 FFMA, 2 lane-private RED.shared and a quarter of a Philox block
// (4.5 live IMAD.WIDE + 5 LOP3) = 28.5 instructions, issued for eight independent chains at a time (every
// instruction's seven neighbours are independent of it), 16 events per loop trip.  No generations, roulette,
// queues, drains or ragged blocks: what the SM can issue of THIS mix, the ceiling the walk kernel is compared with.
#define WM_PRMT_A_(i) asm volatile("prmt.b32 %0, %1, %2, 0x5514;" : "=r"(a##i) : "r"(u##i), "r"(lane8));
#define WM_PRMT_B_(i) asm volatile("prmt.b32 %0, %1, %2, 0x5504;" : "=r"(b##i) : "r"(u##i), "r"(lane8 + 128u));
#define WM_PRMT_C_(i) asm volatile("prmt.b32 %0, %1, %2, 0x5504;" : "=r"(c##i) : "r"(u##i), "r"(lane4));
#define WM_LDS_A_(i) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+0x800];" : "=f"(g##i), "=f"(h##i) : "r"(a##i));
#define WM_LDS_B_(i) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+0x800];" : "=f"(p##i), "=f"(q##i) : "r"(b##i));
#define WM_FFMA_G_(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f##i) : "f"(g##i), "f"(h##i));
#define WM_FFMA_P_(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f##i) : "f"(p##i), "f"(q##i));
#define WM_RED_(i) asm volatile("red.shared.add.u32 [%0+0x10800], %1;" ::"r"(c##i), "r"(u##i));
#define WM_RED2_(i) asm volatile("red.shared.add.u32 [%0+0x10880], %1;" ::"r"(c##i), "r"(u##i));
#define WM_SHFF_(i) asm volatile("{ .reg .u32 t; shf.r.wrap.b32 t, %1, 0x7f, 9; mov.b32 %0, t; }" : "=f"(e##i) : "r"(u##i));
#define WM_FADD_E_(i) asm volatile("add.rn.f32 %0, %0, 0fBF7FFFFF;" : "+f"(e##i));
#define WM_LG2_E_(i) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(e##i));
#define WM_FMUL_E_(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f##i) : "f"(e##i));
#define WM_SQRT_(i) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f##i));
// the packed forms kernel v9 uses (Blackwell f32x2): the (y, z) update of one chain as one FFMA2 with a scalar first
// operand, xi and the shell number of two chains as one FADD2 / FFMA2.RZ
#define WM_FFMA2_YZ_(i) asm volatile("{ .reg .b64 a, b, c; mov.b64 a, {%2, %2}; mov.b64 b, {%3, %4}; mov.b64 c, {%0, %1}; fma.rn.f32x2 c, a, b, c; mov.b64 {%0, %1}, c; }" : "+f"(y##i), "+f"(z##i) : "f"(f##i), "f"(p##i), "f"(q##i));
#define WM_FADD2_E_(i, j) asm volatile("{ .reg .b64 a, c; mov.b64 a, {%0, %1}; mov.b64 c, {%2, %2}; add.rn.f32x2 a, a, c; mov.b64 {%0, %1}, a; }" : "+f"(e##i), "+f"(e##j) : "f"(-0.99999994f));
#define WM_FFMA2_RZ_(i, j) asm volatile("{ .reg .b64 a, b, c; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; mov.b64 c, {%3, %3}; fma.rz.f32x2 a, a, b, c; mov.b64 {%0, %1}, a; }" : "+f"(f##i), "+f"(f##j) : "f"(ca), "f"(cb));
#define PAIRS4(X) X(0, 1) X(2, 3) X(4, 5) X(6, 7)
// DROP selects what is left out, to see what each instruction class costs inside the mix (6 = nothing left out, packed forms):
// 0 nothing, 1 the three PRMT and the SHF (ALU pipe), 2 the Philox quarter block, 3 the shared-memory instructions,
// 4 the two MUFU, 5 the FP32 arithmetic
template <int DROP>
__global__ void __launch_bounds__(512, 1) k_walk_mix(Out* out, unsigned seed)
{
    extern __shared__ __align__(16) unsigned wm_smem[];      // [pad to 0x800 | 64 KB table | 64 KB bins: 256 shells x 2 x 32 lanes]
    const unsigned base = (unsigned)__cvta_generic_to_shared(wm_smem);
    for (unsigned i = threadIdx.x; i < (0x800u - base + 0x10000u + 0x10000u) / 4u; i += blockDim.x) wm_smem[i] = 0x3F800000u;
    F_DECL;
    U_DECL;
    const unsigned lane8 = (threadIdx.x & 15u) * 8u, lane4 = (threadIdx.x & 31u) * 4u;
    unsigned a0 = lane8, a1 = lane8, a2 = lane8, a3 = lane8, a4 = lane8, a5 = lane8, a6 = lane8, a7 = lane8;
    unsigned b0 = lane8 + 128u, b1 = b0, b2 = b0, b3 = b0, b4 = b0, b5 = b0, b6 = b0, b7 = b0;
    unsigned c0 = lane4, c1 = lane4, c2 = lane4, c3 = lane4, c4 = lane4, c5 = lane4, c6 = lane4, c7 = lane4;
    float g0 = ca, g1 = ca, g2 = ca, g3 = ca, g4 = ca, g5 = ca, g6 = ca, g7 = ca, h0 = cb, h1 = cb, h2 = cb, h3 = cb, h4 = cb, h5 = cb, h6 = cb, h7 = cb;
    float p0 = ca, p1 = ca, p2 = ca, p3 = ca, p4 = ca, p5 = ca, p6 = ca, p7 = ca, q0 = cb, q1 = cb, q2 = cb, q3 = cb, q4 = cb, q5 = cb, q6 = cb, q7 = cb;
    float e0 = 1.5f, e1 = 1.5f, e2 = 1.5f, e3 = 1.5f, e4 = 1.5f, e5 = 1.5f, e6 = 1.5f, e7 = 1.5f;
    float y0 = cb, y1 = cb, y2 = cb, y3 = cb, y4 = cb, y5 = cb, y6 = cb, y7 = cb, z0 = ca, z1 = ca, z2 = ca, z3 = ca, z4 = ca, z5 = ca, z6 = ca, z7 = ca;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters / 4; ++it) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {      // 8 events each; 4 wide steps in the first half, 5 in the second
            if (DROP == 6) {                        // kernel v9's mix: 26.5 instructions per event
                REP8(WM_SHFF_) PAIRS4(WM_FADD2_E_) REP8(WM_LG2_E_) REP8(WM_PRMT_A_) REP8(WM_PRMT_B_) REP8(WM_LDS_A_) REP8(WM_LDS_B_)
                REP8(WM_FMUL_E_) REP8(WM_FFMA_G_) REP8(WM_FFMA2_YZ_) REP8(FMUL_) REP8(FFMA_) REP8(FFMA_) REP8(WM_SQRT_) PAIRS4(WM_FFMA2_RZ_)
                REP8(WM_PRMT_C_) REP8(WM_RED_) REP8(WM_RED2_)
                REP8(WIDE_STEP_) REP8(WIDE_STEP_) REP8(WIDE_STEP_) REP8(WIDE_STEP_)
                if (half == 0) { REP8(EXTRA_LOP_) } else { REP8(WIDE_STEP_) }
                continue;
            }
            if (DROP != 1) { REP8(WM_SHFF_) }
            if (DROP != 5) { REP8(WM_FADD_E_) }
            if (DROP != 4) { REP8(WM_LG2_E_) }
            if (DROP != 1) { REP8(WM_PRMT_A_) REP8(WM_PRMT_B_) }
            if (DROP != 3) { REP8(WM_LDS_A_) REP8(WM_LDS_B_) }
            if (DROP != 5) { REP8(WM_FMUL_E_) REP8(WM_FFMA_G_) REP8(WM_FFMA_P_) REP8(FFMA_) REP8(FMUL_) REP8(FFMA_) REP8(FFMA_) }
            if (DROP != 4) { REP8(WM_SQRT_) }
            if (DROP != 5) { REP8(FFMA_) }
            if (DROP != 1) { REP8(WM_PRMT_C_) }
            if (DROP != 3) { REP8(WM_RED_) REP8(WM_RED2_) }
            if (DROP != 2) {
                REP8(WIDE_STEP_) REP8(WIDE_STEP_) REP8(WIDE_STEP_) REP8(WIDE_STEP_)
                if (half == 0) { REP8(EXTRA_LOP_) } else { REP8(WIDE_STEP_) }
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
    if ((F_SINK ^ U_SINK ^ a0 ^ b0 ^ c0 ^ __float_as_uint(e0 + g0 + h0 + p0 + q0 + y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7 + z0 + z1 + z2 + z3 + z4 + z5 + z6 + z7)) == 0x12345u) out[blockIdx.x].sink = 1;
}

// The same mix at twice the occupancy: four chains per thread (so that 64 registers suffice), two 512-thread blocks
// per SM = 32 warps per SM, bins of 128 shells (97 KB per block).  Does the ceiling of the mix rise with more warps?
__global__ void __launch_bounds__(512, 2) k_walk_mix_32warps(Out* out, unsigned seed)
{
    extern __shared__ __align__(16) unsigned wm_smem[];      // [pad to 0x800 | 64 KB table | 32 KB bins: 128 shells x 2 x 32 lanes]
    const unsigned base = (unsigned)__cvta_generic_to_shared(wm_smem);
    for (unsigned i = threadIdx.x; i < (0x800u - base + 0x10000u + 0x8000u) / 4u; i += blockDim.x) wm_smem[i] = 0x3F800000u;
    float f0 = seed * 1e-9f + threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    const float ca = 1.0001f + seed * 1e-12f, cb = 0.5f;
    unsigned u0 = seed + threadIdx.x, u1 = u0 * 3 + 1, u2 = u0 * 5 + 2, u3 = u0 * 7 + 3;
    const unsigned ka = seed | 0xD2511F53u, kb = seed ^ 0x9E3779B9u;
    const unsigned lane8 = (threadIdx.x & 15u) * 8u, lane4 = (threadIdx.x & 31u) * 4u;
    unsigned a0 = lane8, a1 = lane8, a2 = lane8, a3 = lane8, b0 = lane8 + 128u, b1 = b0, b2 = b0, b3 = b0, c0 = lane4, c1 = lane4, c2 = lane4, c3 = lane4;
    float g0 = ca, g1 = ca, g2 = ca, g3 = ca, h0 = cb, h1 = cb, h2 = cb, h3 = cb, p0 = ca, p1 = ca, p2 = ca, p3 = ca, q0 = cb, q1 = cb, q2 = cb, q3 = cb;
    float e0 = 1.5f, e1 = 1.5f, e2 = 1.5f, e3 = 1.5f;
    (void)ka;
#define WM_PRMT_C7_(i) asm volatile("{ .reg .u32 t; and.b32 t, %1, 0x7f; prmt.b32 %0, t, %2, 0x5504; }" : "=r"(c##i) : "r"(u##i), "r"(lane4));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters / 4; ++it) {
#pragma unroll
        for (int quarter = 0; quarter < 4; ++quarter) {      // 4 events each; 4 or 5 wide steps alternately
            REP4(WM_SHFF_) REP4(WM_FADD_E_) REP4(WM_LG2_E_) REP4(WM_PRMT_A_) REP4(WM_PRMT_B_) REP4(WM_LDS_A_) REP4(WM_LDS_B_)
            REP4(WM_FMUL_E_) REP4(WM_FFMA_G_) REP4(WM_FFMA_P_) REP4(FFMA_) REP4(FMUL_) REP4(FFMA_) REP4(FFMA_) REP4(WM_SQRT_) REP4(FFMA_)
            REP4(WM_PRMT_C7_) REP4(WM_RED_) REP4(WM_RED2_)
            REP4(WIDE_STEP_) REP4(WIDE_STEP_) REP4(WIDE_STEP_) REP4(WIDE_STEP_)
            if (quarter & 1) { REP4(WIDE_STEP_) }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
    if ((__float_as_uint(f0 + f1 + f2 + f3 + e0 + g0 + h0 + p0 + q0) ^ u0 ^ u1 ^ u2 ^ u3 ^ a0 ^ b0 ^ c0) == 0x12345u) out[blockIdx.x].sink = 1;
}

struct Result {
    double cycles, ms;
};

template <typename F>
Result run(F launch, Out* d_out, int grid)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<Out> h(grid);
    Result best{1e300, 0};
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(h.data(), d_out, grid * sizeof(Out), cudaMemcpyDeviceToHost));
        double c = 0;
        for (auto& o : h) c += (double)o.cycles;
        c /= grid;
        if (rep > 0 && c < best.cycles) best = {c, ms};
    }
    return best;
}

int main(int argc, char** argv)
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    int blocks_per_sm = argc > 1 ? atoi(argv[1]) : 4;   // 4 x 256 threads = 32 warps / SM
    const int grid = sms * blocks_per_sm;
    Out* d_out;
    CK(cudaMalloc(&d_out, grid * sizeof(Out)));
    CK(cudaMemset(d_out, 0, grid * sizeof(Out)));
    const double warps_per_sm = blocks_per_sm * kBlock / 32.0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"warps_per_sm\": %.0f, \"clock_rate_khz\": %d}\n",
           prop.name, sms, prop.major, prop.minor, warps_per_sm, prop.clockRate);

#define REPORT(NAME, INST_PER_ITER, LAUNCH)                                                        \
    do {                                                                                           \
        Result r = run([&] { LAUNCH; }, d_out, grid);                                              \
        const double winst = (double)(INST_PER_ITER) * kIters * warps_per_sm;                      \
        printf("{\"test\": \"%s\", \"warp_inst_per_clk_per_sm\": %.3f, \"lanes_per_clk_per_sm\": %.1f, " \
               "\"cycles\": %.0f, \"ms\": %.4f, \"sm_mhz\": %.0f}\n",                             \
               NAME, winst / r.cycles, 32.0 * winst / r.cycles, r.cycles, r.ms, r.cycles / r.ms * 1e-3); \
        fflush(stdout);                                                                            \
    } while (0)

#define SIMPLE(K, N) REPORT(#K, N, (K<<<grid, kBlock>>>(d_out, 1u)))
    SIMPLE(k_ffma, 16);
    SIMPLE(k_fmul, 16);
    SIMPLE(k_fadd, 16);
    SIMPLE(k_ffma2, 16);
    SIMPLE(k_imad_wide, 16);
    SIMPLE(k_imad, 16);
    SIMPLE(k_imad_hi, 16);
    SIMPLE(k_lop3, 16);
    SIMPLE(k_iadd, 16);
    SIMPLE(k_shf, 16);
    SIMPLE(k_prmt, 16);
    SIMPLE(k_isetp_padd, 32);
    SIMPLE(k_i2fp, 16);
    SIMPLE(k_f2i, 16);
    SIMPLE(k_mufu_lg2, 16);
    SIMPLE(k_mufu_sqrt, 16);
    SIMPLE(k_mufu_rsq, 16);
    SIMPLE(k_sin_approx, 32);   // FMUL.RZ + MUFU.SIN
    SIMPLE(k_mufu_ex2, 16);
    SIMPLE(k_mix_ffma_lop3, 16);
    SIMPLE(k_mix_imadw_lop3, 16);
    SIMPLE(k_mix_ffma_imadw, 16);
    SIMPLE(k_mix_ffma2_imadw, 16);
    SIMPLE(k_mix_fmul_imadw, 16);
    SIMPLE(k_mix_2ffma_imadw, 24);
    SIMPLE(k_mix_ffma2_lop3, 16);
    SIMPLE(k_imad_wide_live, 32);              // 16 x (IMAD.WIDE + LOP3)
    SIMPLE(k_dispatch_imad_wide_group, 40);    // 8 x (IMAD.WIDE + 2 LOP3 + 2 FMUL)
    SIMPLE(k_dispatch_two_imad_group, 48);     // 8 x (2 IMAD + 2 LOP3 + 2 FMUL)
    SIMPLE(k_mix_14ffma_2mufu, 18);
    SIMPLE(k_mix_8ffma_8lop3_2mufu, 18);
    SIMPLE(k_philox10, 80);     // 2 x (20 IMAD.WIDE + 20 LOP3), key schedule folded by ptxas
    {   // the walk's instruction mix, 16 warps per SM like the shipped kernel (one 512-thread block per SM, 130 KB)
        const size_t wm_bytes = 0x800 + 0x10000 + 0x10000;
        auto mix = [&](const char* name, auto kernel, double inst_per_trip) {
            CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wm_bytes));
            Result r = run([&] { kernel<<<sms, 512, wm_bytes>>>(d_out, 1u); }, d_out, sms);
            const double winst = inst_per_trip * (kIters / 4) * 16.0;
            printf("{\"test\": \"%s\", \"warp_inst_per_clk_per_sm\": %.3f, \"cycles\": %.0f, \"ms\": %.4f, \"sm_mhz\": %.0f, \"warps_per_sm\": 16, "
                   "\"instructions_per_event\": %.2f, \"cycles_per_event_per_smsp\": %.2f}\n",
                   name, winst / r.cycles, r.cycles, r.ms, r.cycles / r.ms * 1e-3, inst_per_trip / 16.0, r.cycles / ((kIters / 4) * 16.0 * 4.0));
            fflush(stdout);
        };
        // per trip of 16 events: 19 walk instructions + a quarter Philox block (4.5 IMAD.WIDE + 5 LOP3) each = 456
        mix("k_walk_mix", k_walk_mix<0>, 456.0);
        mix("k_walk_mix_no_prmt_shf", k_walk_mix<1>, 456.0 - 64.0);
        mix("k_walk_mix_no_philox", k_walk_mix<2>, 456.0 - 152.0);
        mix("k_walk_mix_no_shared", k_walk_mix<3>, 456.0 - 64.0);
        mix("k_walk_mix_no_mufu", k_walk_mix<4>, 456.0 - 32.0);
        mix("k_walk_mix_no_fp32", k_walk_mix<5>, 456.0 - 144.0);
        mix("k_walk_mix_packed", k_walk_mix<6>, 456.0 - 32.0);
        {   // 32 warps per SM: two blocks per SM, four chains per thread (one extra LOP3 per event for the 7-bit shell mask)
            const size_t b2 = 0x800 + 0x10000 + 0x8000;
            CK(cudaFuncSetAttribute(k_walk_mix_32warps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b2));
            Result r = run([&] { k_walk_mix_32warps<<<2 * sms, 512, b2>>>(d_out, 1u); }, d_out, sms);
            printf("{\"test\": \"k_walk_mix_32warps\", \"cycles\": %.0f, \"ms\": %.4f, \"warps_per_sm\": 32, \"cycles_per_event_per_smsp\": %.2f}\n",
                   r.cycles, r.ms, r.cycles / ((kIters / 4) * 16.0 * 8.0));
            fflush(stdout);
        }
    }
    const int nbs[] = {101, 1024, 8192};
    for (int nb : nbs) {
        const size_t sm = (nb + 32) * sizeof(unsigned);
        char name[96];
        snprintf(name, sizeof name, "atoms_red_conflictfree_nb%d", nb);
        REPORT(name, 8, (k_atoms<0, false><<<grid, kBlock, sm>>>(d_out, 1u, nb)));
        snprintf(name, sizeof name, "atoms_red_random_nb%d", nb);
        REPORT(name, 8, (k_atoms<1, false><<<grid, kBlock, sm>>>(d_out, 1u, nb)));
        snprintf(name, sizeof name, "atoms_ret_random_nb%d", nb);
        REPORT(name, 8, (k_atoms<1, true><<<grid, kBlock, sm>>>(d_out, 1u, nb)));
        snprintf(name, sizeof name, "atoms_red_sameaddr_nb%d", nb);
        REPORT(name, 8, (k_atoms<2, false><<<grid, kBlock, sm>>>(d_out, 1u, nb)));
        snprintf(name, sizeof name, "atoms_red_random_hot20_nb%d", nb);
        REPORT(name, 8, (k_atoms<3, false><<<grid, kBlock, sm>>>(d_out, 1u, nb)));
        snprintf(name, sizeof name, "atoms_baseline_random_nb%d", nb);
        REPORT(name, 8, (k_atoms_baseline<1><<<grid, kBlock>>>(d_out, 1u, nb)));
    }
    REPORT("atoms_baseline_sameaddr", 8, (k_atoms_baseline<2><<<grid, kBlock>>>(d_out, 1u, 101)));
    CK(cudaFree(d_out));
    return 0;
}
