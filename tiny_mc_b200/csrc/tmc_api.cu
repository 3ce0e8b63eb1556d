// tmc_api.cu — C-ABI entry points of libtinymc_b200.so (declared in include/tiny_mc_b200.h).
//
// Host side of the drop-in boundary: turns the reference's call site
//     for (i = 0; i < PHOTONS; ++i) photon(heat, heat2);        (reference tiny_mc.c:47-49)
// into one launch per GPU of tmc::photon_walk_kernel, one reduce of the 2*SHELLS+4 tally
// words (NCCL over NVLink when more than one GPU is driven by this process), one small D2H
// copy, and a += into the caller's float arrays (reference photon.c:30-31 semantics).
//
// No CPU fallback: without a usable sm_100 device every compute entry point fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; the library itself is dlopen()ed on first multi-GPU init

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tiny_mc_b200.h"
#include "walk_kernel.cuh"

namespace {

using tmc::WalkArgs;

struct Plan {
    float albedo;
    float shells_per_mfp;
    tmc_scales sc;
    uint32_t weight_one;
    uint32_t heat2_half;
    uint32_t n_gen;
    tmc::GenPlan gen[tmc::kMaxGenerations];   // the deterministic weight schedule, per generation
};

struct Device {
    int id = -1;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;        // odd slots of a batched call: the tail of one launch overlaps the next
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    unsigned long long* d_buf = nullptr;   // u64[2*shells+4]
    size_t buf_words = 0;
    ncclComm_t comm = nullptr;
};

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct Options {
    int philox_rounds = 10;
    int block_threads = 0;   // 0 = auto
    int blocks_per_sm = 0;   // 0 = occupancy
    int flush_iters = 0;     // 0 = auto
    int nccl_reduce = 1;
    int tally_layout = 0;    // 0 = auto, 1 = plain per-block histogram (per-lane overflow slots), 2 = lane-private,
                             // 3 = plain with the saturating clamp (one word for the overflow shell)
    int tally_check_bits = 31;   // a drained u32 word >= 2^bits triggers the retry (tests lower it)
    int walk_mode = 0;           // 0 = the 3-D walk (the product), 1 = the reduced radial walk (cross-check only)
    int batch_streams = 2;       // streams per device the launches of a batched call alternate between (1 or 2)
    int batch_capacity = 1;      // tally slots tmc_prepare sizes the device and pinned buffers for
};

struct Lib {
    bool inited = false;
    std::vector<Device> devs;
    NcclApi nccl;
    Options opt;
    std::string err = "";
    tmc_run_info info{};
    unsigned long long* h_pinned = nullptr;
    size_t h_words = 0;
} g;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g.err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(TMC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t r_ = (expr);                                                                   \
        if (r_ != ncclSuccess)                                                                      \
            return fail(TMC_ERR_NCCL, "%s failed: %s", #expr, g.nccl.GetErrorString ? g.nccl.GetErrorString(r_) : "?"); \
    } while (0)

uint32_t ceil_log2_u64(uint64_t v)
{
    uint32_t b = 0;
    while (b < 63 && (1ull << b) < v) ++b;
    return b;
}

// Optical constants (reference photon.c:8-9) and the fixed-point plan.
int make_plan(const tmc_params* p, Plan* pl)
{
    if (!p) return fail(TMC_ERR_BAD_ARG, "params is NULL");
    if (p->shells < 1u || p->shells > (1u << 22)) return fail(TMC_ERR_BAD_ARG, "SHELLS=%u out of range [1, 2^22]", p->shells);
    if (!(p->mu_a > 0.0f) || !(p->mu_s >= 0.0f) || !(p->microns_per_shell > 0.0f))
        return fail(TMC_ERR_BAD_ARG, "need MU_A > 0, MU_S >= 0, MICRONS_PER_SHELL > 0");
    pl->albedo = p->mu_s / (p->mu_s + p->mu_a);
    pl->shells_per_mfp = static_cast<float>(1e4 / static_cast<double>(p->microns_per_shell) /
                                            static_cast<double>(p->mu_a + p->mu_s));
    const double absorb = 1.0 - static_cast<double>(pl->albedo);
    int hs = 18 - static_cast<int>(std::ceil(std::log2(absorb)));   // largest deposit in [2^17, 2^18)
    if (hs > 30) hs = 30;
    if (hs < 14) hs = 14;
    pl->sc.heat_shift = static_cast<uint32_t>(hs);
    pl->weight_one = 1u << hs;
    double q = std::floor(absorb * 4294967296.0 + 0.5);
    if (q > 4294967295.0) q = 4294967295.0;
    if (q < 1.0) q = 1.0;
    pl->sc.absorb_q32 = static_cast<uint32_t>(q);
    const uint64_t dep_max = (static_cast<uint64_t>(pl->weight_one) * pl->sc.absorb_q32) >> 32;
    const uint32_t bits = ceil_log2_u64(dep_max + 1);
    pl->sc.heat2_rshift = (2 * bits > 18) ? 2 * bits - 18 : 0;
    pl->heat2_half = pl->sc.heat2_rshift ? (1u << (pl->sc.heat2_rshift - 1)) : 0u;
    pl->sc.roulette_thr = static_cast<uint32_t>(std::floor(0.001 * pl->weight_one + 0.5));
    // Generation g = number of roulettes survived.  Every photon of a generation starts it with
    // the same weight and needs the same number of events to fall below the threshold
    // (reference photon.c:32,45-48: w *= albedo per event, x10 per survived roulette).
    uint32_t w = pl->weight_one, e = 1;
    pl->n_gen = tmc::kMaxGenerations;
    for (uint32_t g = 0; g < pl->n_gen; ++g) {
        uint32_t k = 0;
        pl->gen[g].first_event = e;
        pl->gen[g].w_start = w;
        do {
            w -= static_cast<uint32_t>((static_cast<uint64_t>(w) * pl->sc.absorb_q32 + 0x80000000ull) >> 32);
            ++k;
        } while (w >= pl->sc.roulette_thr && k < (1u << 22));
        if (w >= pl->sc.roulette_thr)   // the deposit table would exceed ~300 MB (absorbed fraction per event below ~2e-6)
            return fail(TMC_ERR_BAD_ARG, "MU_A / (MU_A + MU_S) = %g is too small: a photon needs > 2^22 events per generation", absorb);
        pl->gen[g].n_events = k;
        e += k;
        w *= 10u;
    }
    return TMC_OK;
}

// (deposit, rescaled deposit^2) of every event number of the deterministic weight schedule
// (Plan::gen), computed with the exact integer recurrence and uploaded once per device and
// optics.  Tables are never overwritten (kernels on other streams may still read them).
struct DepositTable {
    int device;
    uint32_t absorb_q32, weight_one, heat2_rshift, roulette_thr;
    uint2* d_table;
};
std::vector<DepositTable> g_deposit_tables;

int deposit_table(int device, const Plan& pl, const uint2** out)
{
    for (const DepositTable& t : g_deposit_tables)
        if (t.device == device && t.absorb_q32 == pl.sc.absorb_q32 && t.weight_one == pl.weight_one &&
            t.heat2_rshift == pl.sc.heat2_rshift && t.roulette_thr == pl.sc.roulette_thr) {
            *out = t.d_table;
            return TMC_OK;
        }
    // bounded cache (parameter sweeps): beyond kMaxTablesPerDevice optics the oldest table of this device
    // goes, after every kernel that may still read it has finished
    constexpr size_t kMaxTablesPerDevice = 16;
    size_t mine = 0;
    for (const DepositTable& t : g_deposit_tables) mine += t.device == device;
    if (mine >= kMaxTablesPerDevice)
        for (size_t i = 0; i < g_deposit_tables.size(); ++i)
            if (g_deposit_tables[i].device == device) {
                CUDA_TRY(cudaDeviceSynchronize());
                CUDA_TRY(cudaFree(g_deposit_tables[i].d_table));
                g_deposit_tables.erase(g_deposit_tables.begin() + static_cast<long>(i));
                break;
            }
    const tmc::GenPlan& last = pl.gen[pl.n_gen - 1];
    std::vector<uint2> h(static_cast<size_t>(last.first_event) + last.n_events, make_uint2(0u, 0u));
    for (uint32_t g = 0; g < pl.n_gen; ++g) {
        uint32_t w = pl.gen[g].w_start;
        for (uint32_t k = 0; k < pl.gen[g].n_events; ++k) {
            const uint32_t dep = static_cast<uint32_t>((static_cast<uint64_t>(w) * pl.sc.absorb_q32 + 0x80000000ull) >> 32);
            const uint32_t dep2 = static_cast<uint32_t>((static_cast<uint64_t>(dep) * dep + pl.heat2_half) >> pl.sc.heat2_rshift);
            h[pl.gen[g].first_event + k] = make_uint2(dep, dep2);
            w -= dep;
        }
    }
    uint2* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(uint2)));
    CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());   // see directions_table(): the kernel's stream is not ordered after this copy
    g_deposit_tables.push_back(DepositTable{ device, pl.sc.absorb_q32, pl.weight_one, pl.sc.heat2_rshift, pl.sc.roulette_thr, d });
    *out = d;
    return TMC_OK;
}

using KernelFn = void (*)(const WalkArgs);
constexpr int kPhotonsPerLane = TMC_PPL;   // photons per lane of every compiled kernel variant

// default block shapes (measured best, profiles/): lane-private tallies / one histogram per block
#ifndef TMC_DEFAULT_BLOCK_PRIVATE
#define TMC_DEFAULT_BLOCK_PRIVATE 512
#endif
#ifndef TMC_DEFAULT_BLOCK_PLAIN
#define TMC_DEFAULT_BLOCK_PLAIN 512
#endif

// Block shapes: threads per block (two photons per thread) x the residency the register
// budget is compiled for.  A block needs 64 KB (direction table) + its tallies (SHELLS * 256 B
// lane-private, else (SHELLS + 31) * 8 B) of shared memory, so at most 3 blocks fit an SM.
template <int ROUNDS, bool LANE_PRIVATE>
KernelFn kernel_for_block(int block, int per_sm)
{
#ifdef TMC_ONLY_BLOCK   /* timing-experiment builds (tools/experiments.sh): one block shape, one block per SM */
    return block == TMC_ONLY_BLOCK && per_sm == 1 ? tmc::photon_walk_kernel<ROUNDS, TMC_ONLY_BLOCK, 1, LANE_PRIVATE, false, kPhotonsPerLane> : nullptr;
#else
    // (threads per block, blocks per SM the register budget is compiled for)
    switch (block * 8 + per_sm) {
    case 128 * 8 + 3: return tmc::photon_walk_kernel<ROUNDS, 128, 3, LANE_PRIVATE, false, kPhotonsPerLane>;   // 168 registers
    case 256 * 8 + 2: return tmc::photon_walk_kernel<ROUNDS, 256, 2, LANE_PRIVATE, false, kPhotonsPerLane>;   // 128
    case 256 * 8 + 3: return tmc::photon_walk_kernel<ROUNDS, 256, 3, LANE_PRIVATE, false, kPhotonsPerLane>;   //  80
    case 512 * 8 + 1: return tmc::photon_walk_kernel<ROUNDS, 512, 1, LANE_PRIVATE, false, kPhotonsPerLane>;   // 128
    case 768 * 8 + 1: return tmc::photon_walk_kernel<ROUNDS, 768, 1, LANE_PRIVATE, false, kPhotonsPerLane>;   //  80
    case 1024 * 8 + 1: return tmc::photon_walk_kernel<ROUNDS, 1024, 1, LANE_PRIVATE, false, kPhotonsPerLane>; //  64
    default: return nullptr;
    }
#endif
}

// the reduced radial walk ("walk_mode" = 1) is compiled for the default shapes only
template <int ROUNDS>
KernelFn radial_kernel(int block, int per_sm, bool lane_private)
{
    if (lane_private && block == TMC_DEFAULT_BLOCK_PRIVATE && per_sm == 1) return tmc::photon_walk_kernel<ROUNDS, TMC_DEFAULT_BLOCK_PRIVATE, 1, true, true, kPhotonsPerLane>;
    if (!lane_private && block == TMC_DEFAULT_BLOCK_PLAIN && per_sm == 1) return tmc::photon_walk_kernel<ROUNDS, TMC_DEFAULT_BLOCK_PLAIN, 1, false, true, kPhotonsPerLane>;
    return nullptr;
}

// the one-histogram layout with the saturating clamp (walk_kernel.cuh: SAT_PLAIN) is compiled for the default shape only
template <int ROUNDS>
KernelFn sat_plain_kernel(int block, int per_sm)
{
    if (block == TMC_DEFAULT_BLOCK_PLAIN && per_sm == 1) return tmc::photon_walk_kernel<ROUNDS, TMC_DEFAULT_BLOCK_PLAIN, 1, false, false, kPhotonsPerLane, true>;
    return nullptr;
}

KernelFn pick_kernel(int rounds, int block, int per_sm, bool lane_private, bool radial, bool sat_plain = false)
{
    if (sat_plain && !lane_private && !radial) {
        KernelFn fn = rounds == 10 ? sat_plain_kernel<10>(block, per_sm) : rounds == 7 ? sat_plain_kernel<7>(block, per_sm) : nullptr;
        if (fn) return fn;      // else: the per-lane-slot variant of that shape (same tallies)
    }
    if (radial) return rounds == 10 ? radial_kernel<10>(block, per_sm, lane_private) : rounds == 7 ? radial_kernel<7>(block, per_sm, lane_private) : nullptr;
    switch (rounds) {
    case 7: return lane_private ? kernel_for_block<7, true>(block, per_sm) : kernel_for_block<7, false>(block, per_sm);
    case 10: return lane_private ? kernel_for_block<10, true>(block, per_sm) : kernel_for_block<10, false>(block, per_sm);
    default: return nullptr;
    }
}

// the register budgets (blocks per SM) compiled for each block size, best first
int default_blocks_per_sm(int block)
{
    switch (block) {
    case 128: return 3;
    case 256: return 2;
    default: return 1;
    }
}

// The direction table (walk_kernel.cuh: event()): 256 polar entries (-ln2 cos, -ln2 sin)(theta_k) with
// cos(theta_k) = (2k + 1)/256 - 1 (midpoints: odd moments vanish, E[cos^2] = 1/3 (1 - 2^-16)), then 256
// azimuth entries (cos, sin)(2 pi k / 256); computed in double, rounded once, uploaded once per
// device; every block stages 16 copies of it into shared memory.  -ln2 turns log2(xi) into the step.
const float2* g_directions[64] = {};

// Survivor-queue scratch (walk_kernel.cuh): kQueueBytesPerWarp per warp of the largest grid, one
// buffer per device and stream (two launches on one stream are ordered, so they can share it).
struct QueueScratch {
    int device;
    cudaStream_t stream;
    uint32_t* d_buf;
    size_t bytes;
};
std::vector<QueueScratch> g_queue_scratch;

int queue_scratch(int device, cudaStream_t stream, size_t bytes, uint32_t** out)
{
    for (QueueScratch& q : g_queue_scratch)
        if (q.device == device && q.stream == stream) {
            if (q.bytes < bytes) {
                CUDA_TRY(cudaStreamSynchronize(stream));
                CUDA_TRY(cudaFree(q.d_buf));
                q.d_buf = nullptr;
                q.bytes = 0;
                CUDA_TRY(cudaMalloc(&q.d_buf, bytes));
                q.bytes = bytes;
            }
            *out = q.d_buf;
            return TMC_OK;
        }
    // bounded: a host that keeps creating and destroying streams must not grow this list for ever
    // (a destroyed stream's handle may be reused by a later stream, which then simply inherits the buffer)
    constexpr size_t kMaxScratchPerDevice = 32;
    size_t mine = 0;
    for (const QueueScratch& q : g_queue_scratch) mine += q.device == device;
    if (mine >= kMaxScratchPerDevice)
        for (size_t i = 0; i < g_queue_scratch.size(); ++i)
            if (g_queue_scratch[i].device == device) {
                CUDA_TRY(cudaDeviceSynchronize());
                CUDA_TRY(cudaFree(g_queue_scratch[i].d_buf));
                g_queue_scratch.erase(g_queue_scratch.begin() + static_cast<long>(i));
                break;
            }
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, bytes));
    g_queue_scratch.push_back(QueueScratch{ device, stream, d, bytes });
    *out = d;
    return TMC_OK;
}

int directions_table(int device, const float2** out)
{
    if (device < 0 || device >= 64) return fail(TMC_ERR_BAD_ARG, "device %d out of range", device);
    if (!g_directions[device]) {
        std::vector<float2> h(2 * tmc::kDirEntries);
        const double ln2 = 0.693147180559945309417232;
        for (int k = 0; k < tmc::kDirEntries; ++k) {
            const double ct = (2.0 * k + 1.0) / tmc::kDirEntries - 1.0;
            const double st = std::sqrt(1.0 - ct * ct);
            h[k] = make_float2(static_cast<float>(-ln2 * ct), static_cast<float>(-ln2 * st));
            const double phi = 6.283185307179586476925 * static_cast<double>(k) / tmc::kDirEntries;
            h[tmc::kDirEntries + k] = make_float2(static_cast<float>(std::cos(phi)), static_cast<float>(std::sin(phi)));
        }
        float2* d = nullptr;
        CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(float2)));
        CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
        // a pageable cudaMemcpy may return once the data is staged; the walk runs on non-blocking
        // streams that are not ordered after the legacy stream, so wait for the DMA here (one-off)
        CUDA_TRY(cudaDeviceSynchronize());
        g_directions[device] = d;
    }
    *out = g_directions[device];
    return TMC_OK;
}

struct LaunchCfg {
    KernelFn fn;
    int block;
    int grid;
    int full_grid;   // the grid a launch that fills the device would use (>= grid)
    size_t smem;
    uint32_t flush_iters;
};

// Per-device and per-kernel facts are looked up once: cudaGetDeviceProperties and
// cudaFuncSetAttribute cost milliseconds, and the device-resident entry point is called per step.
struct DeviceFacts {
    bool known = false;
    int rc = TMC_OK;
    int sms = 0;
};
DeviceFacts g_facts[64];

struct KernelFacts {
    KernelFn fn;
    int device;
    int block;
    size_t smem;
    int per_sm;
};
std::vector<KernelFacts> g_kernel_facts;

int check_device_arch(int dev);

int device_facts(int dev, int* sms)
{
    if (dev < 0 || dev >= 64) return fail(TMC_ERR_NO_DEVICE, "CUDA device %d out of range", dev);
    DeviceFacts& f = g_facts[dev];
    if (!f.known) {
        f.rc = check_device_arch(dev);
        if (f.rc == TMC_OK) CUDA_TRY(cudaDeviceGetAttribute(&f.sms, cudaDevAttrMultiProcessorCount, dev));
        f.known = (f.rc == TMC_OK);
        if (f.rc) return f.rc;
    }
    *sms = f.sms;
    return TMC_OK;
}

int kernel_occupancy(KernelFn fn, int block, size_t smem, int* per_sm)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    for (const KernelFacts& k : g_kernel_facts)
        if (k.fn == fn && k.device == dev && k.block == block && k.smem == smem) {
            *per_sm = k.per_sm;
            return TMC_OK;
        }
    // allow the largest block this device offers once and for all: the attribute is per function, and
    // setting it to each request's size would LOWER it again for a later, larger request
    int optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes fattr;
    CUDA_TRY(cudaFuncGetAttributes(&fattr, fn));
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - static_cast<int>(fattr.sharedSizeBytes)));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, fn, block, smem));
    if (g_kernel_facts.size() >= 1024) g_kernel_facts.clear();    // plain numbers, recomputed on demand (SHELLS sweeps)
    g_kernel_facts.push_back(KernelFacts{ fn, dev, block, smem, *per_sm });
    return TMC_OK;
}

double shells_per_mfp_of(const tmc_params* p)
{
    return 1e4 / static_cast<double>(p->microns_per_shell) / static_cast<double>(p->mu_a + p->mu_s);
}

double g_launch_us = 0.0;   // host time inside cudaLaunchKernel since the last trace line (TMC_TRACE)

// An upper bound on the share of scatter events beyond the grid radius R = SHELLS / shells_per_mfp mean free paths.
// A photon's n-th position is a sum of n isotropic steps of Exp(1) length; one coordinate X of it has the moment
// generating function (atanh(s) / s)^n, so P(|r| > R) <= 6 min_s exp(-s R / sqrt(3)) (atanh(s) / s)^n (Chernoff, three
// coordinates, two signs), with n = the events of generations 0 .. 5 (deeper ones hold 1e-6 of the photons).
double overflow_bound(const Plan& pl, uint32_t shells)
{
    const double radius = static_cast<double>(shells) / static_cast<double>(pl.shells_per_mfp);
    double n = 0.0;
    for (uint32_t i = 0; i < pl.n_gen && i < 6u; ++i) n += static_cast<double>(pl.gen[i].n_events);
    double best = 1.0;
    for (double s = 0.05; s < 0.96; s += 0.05) {
        const double b = 6.0 * std::exp(-s * radius / std::sqrt(3.0) + n * std::log(std::atanh(s) / s));
        if (b < best) best = b;
    }
    return best;
}

int configure_launch(const tmc_params* p, const Plan& pl, int device_sms, uint64_t count, uint32_t flush_override, LaunchCfg* cfg)
{
    bool lane_private = p->shells <= tmc::kLanePrivateMaxShells;
    if (g.opt.tally_layout == 1 || g.opt.tally_layout == 3) lane_private = false;
    // one histogram per block: the integer clamp to per-lane overflow slots costs an instruction per event; where no
    // photon leaves the grid in practice the saturating clamp (all of them into the ONE word of the last shell) is free
    const bool sat_plain = !lane_private && (g.opt.tally_layout == 3 || (g.opt.tally_layout == 0 && overflow_bound(pl, p->shells) < 1e-6));
    if (g.opt.tally_layout == 2 && !lane_private)
        return fail(TMC_ERR_BAD_ARG, "SHELLS=%u is too large for lane-private tallies (max %u)", p->shells, tmc::kLanePrivateMaxShells);
    int block = g.opt.block_threads;
    if (block == 0) block = lane_private ? TMC_DEFAULT_BLOCK_PRIVATE : TMC_DEFAULT_BLOCK_PLAIN;
    const size_t smem = tmc::walk_smem_bytes(p->shells, lane_private, static_cast<uint32_t>(block));
    if (smem > 227u * 1024u)
        return fail(TMC_ERR_BAD_ARG, "SHELLS=%u with %d-thread blocks needs %zu B of shared memory per block (> 227 KB)", p->shells, block, smem);
    // register budget: the variant compiled for `want` resident blocks (fewer if shared memory says so)
    int want = g.opt.blocks_per_sm > 0 ? g.opt.blocks_per_sm : default_blocks_per_sm(block);
    const int smem_fit = static_cast<int>((227u * 1024u) / (smem + 1024u));
    if (want > smem_fit) want = smem_fit;
    KernelFn fn = nullptr;
    for (int c = want; c >= 1 && !fn; --c) fn = pick_kernel(g.opt.philox_rounds, block, c, lane_private, g.opt.walk_mode == 1, sat_plain);   // largest budget <= want
    for (int c = want + 1; c <= 3 && !fn; ++c) fn = pick_kernel(g.opt.philox_rounds, block, c, lane_private, g.opt.walk_mode == 1, sat_plain);   // else the next one
    if (!fn)
        return fail(TMC_ERR_BAD_ARG, "no kernel for philox_rounds=%d block_threads=%d blocks_per_sm=%d walk_mode=%d", g.opt.philox_rounds,
                    block, g.opt.blocks_per_sm, g.opt.walk_mode);
    int per_sm = 0;
    int orc = kernel_occupancy(fn, block, smem, &per_sm);
    if (orc) return orc;
    if (per_sm < 1) return fail(TMC_ERR_CUDA, "kernel does not fit on an SM (block=%d smem=%zu)", block, smem);
    if (g.opt.blocks_per_sm > 0 && g.opt.blocks_per_sm < per_sm) per_sm = g.opt.blocks_per_sm;
    uint64_t grid = static_cast<uint64_t>(device_sms) * per_sm;
    cfg->full_grid = static_cast<int>(grid);
    const uint64_t warps = static_cast<uint64_t>(block) / 32u;
    const uint64_t cohort = 32ull * kPhotonsPerLane;
    const uint64_t needed = ((count + cohort - 1) / cohort + warps - 1) / warps;   // cohorts of photons, one per warp
    if (needed < grid) grid = needed ? needed : 1;
    uint32_t flush = flush_override ? flush_override : static_cast<uint32_t>(g.opt.flush_iters);
    if (flush == 0) {
        // Philox blocks (kEventsPerBlock events for each of a warp's photons) between two drains of
        // the u32 block histograms.  d0 = the largest amount one event adds to a word.
        const uint64_t dep0 = (static_cast<uint64_t>(pl.weight_one) * pl.sc.absorb_q32 + 0x80000000ull) >> 32;
        const uint64_t dep20 = (dep0 * dep0 + pl.heat2_half) >> pl.sc.heat2_rshift;
        const double d0 = static_cast<double>(dep0 > dep20 ? dep0 : dep20);
        // A slice is drained once per `warps` drain calls of the block.  Between two drains of one slice
        // every warp walks at most 2 * flush blocks (its own calls may fall anywhere in the interval),
        // i.e. ev = 2 * flush * kEventsPerBlock events per photon slot.
        const double ev_per_flush = 2.0 * tmc::kEventsPerBlock;
        const double ppl = static_cast<double>(kPhotonsPerLane);
        if (lane_private) {
            // A (shell, lane) word only ever sees the ppl photon slots per warp of its own lane.  Two
            // hard bounds on what they can add to it between two drains:
            //  (a) every event deposits at most d0:  warps * ppl * ev_per_flush * flush * d0 < 2^32;
            //  (b) a photon deposits at most its whole weight 2^heat_shift in its life, and at most
            //      ceil(ev_per_flush * flush / K0) + 1 lives per photon slot touch the interval.
            // Either suffices, so the larger interval is taken; the 2^31 check stays as a tripwire.
            const double by_event = 4294967296.0 / (ev_per_flush * ppl * static_cast<double>(warps) * (d0 + 1.0));
            const double lives = 4294967296.0 / (ppl * static_cast<double>(warps) * static_cast<double>(pl.weight_one)) - 1.0;
            const double by_life = lives >= 2.0 ? (std::floor(lives) - 1.0) * static_cast<double>(pl.gen[0].n_events) / ev_per_flush : 0.0;
            double blocks = by_event > by_life ? by_event : by_life;
            if (blocks > 512.0) blocks = 512.0;
            flush = blocks < 1.0 ? 1u : static_cast<uint32_t>(blocks);
        } else {
            // One histogram per block, every photon of the block can hit the same word.  The busiest
            // shell takes at most the first-collision share 1 - exp(-1/shells_per_mfp) <= 1/shells_per_mfp
            // of a block's events in expectation (x2 margin, and the target is 2^31, half of what would
            // wrap; the 2^31 check catches the rest: tmc_photons* repeat the range with a shorter interval,
            // callers of tmc_photons_device must test the flag word, e.g. with tmc_device_tallies_check).
            double share = 2.0 / static_cast<double>(shells_per_mfp_of(p));
            if (share > 1.0 || g.opt.tally_layout == 3) share = 1.0;   // forced single overflow word: it may take every event
            const double blocks = 2147483648.0 / (32.0 * ppl * static_cast<double>(warps) * ev_per_flush * share * (d0 + 1.0));
            flush = blocks > 64.0 ? 64u : (blocks < 1.0 ? 1u : static_cast<uint32_t>(blocks));
        }
        if (flush < 1u) flush = 1u;
    }
    if (flush > 4096u) flush = 4096u;
    cfg->fn = fn;
    cfg->block = block;
    cfg->grid = static_cast<int>(grid);
    cfg->smem = smem;
    cfg->flush_iters = flush;
    return TMC_OK;
}

// Enqueue the walk over [first, first + count) on `stream` of the current device.  The kernel
// addresses photons with 32-bit offsets inside one 2^32 window of the global index space, so
// the range is cut at multiples of 2^32 and into pieces of at most 2^30 photons.
int enqueue_walk(const tmc_params* p, const Plan& pl, uint64_t seed, uint64_t first, uint64_t count,
                 int device_sms, uint32_t flush_override, unsigned long long* d_buf, cudaStream_t stream, LaunchCfg* first_cfg)
{
    WalkArgs a;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = directions_table(dev, &a.directions);
    if (rc) return rc;
    tmc::philox_expand_key(seed, &a.keys);
    a.tallies = d_buf;
    a.counters = d_buf + 2ull * p->shells;
    // positions in units of the grid radius SHELLS / shells_per_mfp (walk_kernel.cuh: radius_sq / shell_bits)
    a.pos_scale = static_cast<float>(static_cast<double>(pl.shells_per_mfp) / static_cast<double>(p->shells));
    a.shell_scale = std::nextafterf(static_cast<float>(p->shells), 0.0f);
    a.radial_step = static_cast<float>(-0.693147180559945309417232 * static_cast<double>(a.pos_scale));
    a.shells = p->shells;
    a.last_bits = tmc::kMagicBits + p->shells - 1u;
    rc = deposit_table(dev, pl, &a.deposits);
    if (rc) return rc;
    a.n_gen = pl.n_gen;
    for (uint32_t i = 0; i < pl.n_gen; ++i) a.gen[i] = pl.gen[i];
    bool have_first = false;
    while (count > 0) {
        const uint64_t window_left = (1ull << 32) - (first & 0xFFFFFFFFull);
        uint64_t n = count < (1ull << 30) ? count : (1ull << 30);
        if (n > window_left) n = window_left;
        LaunchCfg cfg{};
        rc = configure_launch(p, pl, device_sms, n, flush_override, &cfg);
        if (rc) return rc;
        if (!have_first && first_cfg) *first_cfg = cfg;
        have_first = true;
        // sized for a full grid of this block shape even when this launch is small (tmc_prepare's is): growing the
        // buffer later would cost a stream synchronisation and a cudaFree / cudaMalloc inside a timed call
        rc = queue_scratch(dev, stream, static_cast<size_t>(cfg.full_grid) * (cfg.block / 32) * tmc::kQueueBytesPerWarp, &a.queues);
        if (rc) return rc;
        a.first = first;
        a.count = n;
        a.flush_blocks = cfg.flush_iters;
        a.check_shift = static_cast<uint32_t>(g.opt.tally_check_bits);
        void* params[] = { &a };
        const auto tl0 = std::chrono::steady_clock::now();
        CUDA_TRY(cudaLaunchKernel(reinterpret_cast<const void*>(cfg.fn), dim3(cfg.grid), dim3(cfg.block), params, cfg.smem, stream));
        g_launch_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tl0).count();
        g.info.gpu_launches += 1;
        first += n;
        count -= n;
    }
    return TMC_OK;
}

int check_device_arch(int dev)
{
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(TMC_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev, prop.name, prop.major, prop.minor);
    return TMC_OK;
}

int load_nccl()
{
    if (g.nccl.handle) return TMC_OK;
    // NCCL prints its banner ("NCCL version ...", NCCL_DEBUG=VERSION/INFO) on stdout by default; stdout
    // belongs to the host program's printout (reference tiny_mc.c:37-66), so send it to stderr.
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) {
        g.nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g.nccl.handle) break;
    }
    if (!g.nccl.handle) return fail(TMC_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LOAD(sym)                                                                     \
    g.nccl.sym = reinterpret_cast<decltype(g.nccl.sym)>(dlsym(g.nccl.handle, "nccl" #sym)); \
    if (!g.nccl.sym) return fail(TMC_ERR_NCCL, "libnccl lacks nccl" #sym)
    LOAD(CommInitAll);
    LOAD(CommDestroy);
    LOAD(Reduce);
    LOAD(GroupStart);
    LOAD(GroupEnd);
    LOAD(GetErrorString);
#undef LOAD
    return TMC_OK;
}

int ensure_buffers(size_t words)
{
    for (Device& d : g.devs) {
        if (d.buf_words >= words) continue;
        CUDA_TRY(cudaSetDevice(d.id));
        if (d.d_buf) CUDA_TRY(cudaFree(d.d_buf));
        d.d_buf = nullptr;
        CUDA_TRY(cudaMalloc(&d.d_buf, words * sizeof(unsigned long long)));
        d.buf_words = words;
    }
    const size_t host_words = words * g.devs.size();
    if (g.h_words < host_words) {
        if (g.h_pinned) CUDA_TRY(cudaFreeHost(g.h_pinned));
        g.h_pinned = nullptr;
        CUDA_TRY(cudaMallocHost(&g.h_pinned, host_words * sizeof(unsigned long long)));
        g.h_words = host_words;
    }
    return TMC_OK;
}

// One photon range with its own tally words ("slot"): a whole call, or one batch of a batched call.
struct Slot {
    uint64_t first, count;
};

bool trace_enabled()
{
    static const bool on = std::getenv("TMC_TRACE") != nullptr;   // per-phase host timings on stderr
    return on;
}

// One pass: every slot's range is sharded over the devices, EVERYTHING is enqueued before anything is
// waited for (all launches of all devices, then the one reduce, then the one copy), and only then
// does the host block.  On success `out` (host, slots * (2*shells+4) words) holds the summed buffers.
int run_slots(const tmc_params* p, const Plan& pl, uint64_t seed, const std::vector<Slot>& slots,
              uint32_t flush_override, std::vector<unsigned long long>& out, double* kernel_ms)
{
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    const size_t words = 2ull * p->shells + 4ull;
    const size_t total = words * slots.size();
    const int ng = static_cast<int>(g.devs.size());
    const size_t capacity = words * static_cast<size_t>(g.opt.batch_capacity);      // growing later costs a (pinned) re-allocation
    int rc = ensure_buffers(total > capacity ? total : capacity);
    if (rc) return rc;
    LaunchCfg cfg0{};
    // Slot-major: every device gets its first launch before any device gets its second.  With more than one slot
    // the launches alternate between two streams per device, so that the blocks of the next launch fill the SMs the
    // previous one's tail has already left (the walk is one persistent block per SM with a static share of work).
    const bool two_streams = slots.size() > 1 && g.opt.batch_streams == 2;
    for (int i = 0; i < ng; ++i) {
        Device& d = g.devs[i];
        CUDA_TRY(cudaSetDevice(d.id));
        CUDA_TRY(cudaMemsetAsync(d.d_buf, 0, total * sizeof(unsigned long long), d.stream));
        CUDA_TRY(cudaEventRecord(d.ev0, d.stream));
        if (two_streams) CUDA_TRY(cudaStreamWaitEvent(d.stream2, d.ev0, 0));
    }
    for (size_t k = 0; k < slots.size(); ++k)
        for (int i = 0; i < ng; ++i) {
            Device& d = g.devs[i];
            const uint64_t count = slots[k].count;
            const uint64_t lo = slots[k].first + count / ng * i + (static_cast<uint64_t>(i) < count % ng ? i : count % ng);
            const uint64_t n = count / ng + (static_cast<uint64_t>(i) < count % ng ? 1 : 0);
            if (n == 0) continue;
            CUDA_TRY(cudaSetDevice(d.id));
            rc = enqueue_walk(p, pl, seed, lo, n, d.sms, flush_override, d.d_buf + k * words, (two_streams && (k & 1u)) ? d.stream2 : d.stream,
                              (i == 0 && k == 0) ? &cfg0 : nullptr);
            if (rc) return rc;
        }
    for (int i = 0; i < ng; ++i) {
        Device& d = g.devs[i];
        CUDA_TRY(cudaSetDevice(d.id));
        if (two_streams) {
            CUDA_TRY(cudaEventRecord(d.ev2, d.stream2));
            CUDA_TRY(cudaStreamWaitEvent(d.stream, d.ev2, 0));
        }
        CUDA_TRY(cudaEventRecord(d.ev1, d.stream));
    }
    const auto t1 = clk::now();
    const bool use_nccl = ng > 1 && g.opt.nccl_reduce && g.devs[0].comm;
    if (use_nccl) {
        // The single collective of the path: sum heat|heat2|counters (u64, exact) onto device 0.
        NCCL_TRY(g.nccl.GroupStart());
        for (int i = 0; i < ng; ++i) {
            Device& d = g.devs[i];
            NCCL_TRY(g.nccl.Reduce(d.d_buf, d.d_buf, total, ncclUint64, ncclSum, 0, d.comm, d.stream));
        }
        NCCL_TRY(g.nccl.GroupEnd());
    }
    const int n_read = use_nccl ? 1 : ng;
    for (int i = 0; i < n_read; ++i) {
        Device& d = g.devs[i];
        CUDA_TRY(cudaSetDevice(d.id));
        CUDA_TRY(cudaMemcpyAsync(g.h_pinned + i * total, d.d_buf, total * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream));
    }
    const auto t2 = clk::now();
    // device 0 last: with NCCL its stream ends with the reduce that needs every other device's kernel
    double worst = 0.0;
    for (int i = ng - 1; i >= 0; --i) {
        Device& d = g.devs[i];
        CUDA_TRY(cudaStreamSynchronize(d.stream));
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        if (ms > worst) worst = ms;
    }
    const auto t3 = clk::now();
    out.assign(total, 0ull);
    for (int i = 0; i < n_read; ++i)
        for (size_t w = 0; w < total; ++w) out[w] += g.h_pinned[i * total + w];
    *kernel_ms += worst;
    g.info.blocks_per_gpu = static_cast<uint32_t>(cfg0.grid);
    g.info.threads_per_block = static_cast<uint32_t>(cfg0.block);
    g.info.flush_iters = cfg0.flush_iters;
    g.info.smem_bytes = static_cast<uint32_t>(cfg0.smem);
    if (trace_enabled()) {
        auto us = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        std::fprintf(stderr, "tmc trace: %d gpus, %zu slots: enqueue %.0f us (of which cudaLaunchKernel %.0f us), reduce+copy enqueue %.0f us, wait %.0f us (kernels %.0f us), sum %.0f us\n",
                     ng, slots.size(), us(t0, t1), g_launch_us, us(t1, t2), us(t2, t3), worst * 1e3, us(t3, clk::now()));
        g_launch_us = 0.0;
    }
    return TMC_OK;
}

int run_range(const tmc_params* p, const Plan& pl, uint64_t seed, uint64_t first, uint64_t count,
              uint32_t flush_override, std::vector<unsigned long long>& out, double* kernel_ms)
{
    return run_slots(p, pl, seed, std::vector<Slot>{ Slot{ first, count } }, flush_override, out, kernel_ms);
}

// The counter words behind a slot's tallies: range flag and internal error -> status code.
int check_counters(const tmc_params* p, const unsigned long long* slot_words)
{
    if (slot_words[2ull * p->shells + 3] != 0ull)
        return fail(TMC_ERR_CUDA, "internal: the block's shared-memory window does not start where walk_kernel.cuh assumes");
    if (slot_words[2ull * p->shells + 2] != 0ull) return TMC_ERR_TALLY_RANGE;
    return TMC_OK;
}

// Whole call: chunk so that no u64 tally can overflow, retry with a shorter drain interval if
// the 2^31 range check fired, add the exact totals into the caller's u64 arrays.
int run_fx(const tmc_params* p, uint64_t seed, uint64_t first, uint64_t n, uint64_t* heat_fx, uint64_t* heat2_fx)
{
    if (!g.inited) return fail(TMC_ERR_NO_DEVICE, "tmc_init has not been called (or found no sm_100 GPU)");
    if (!heat_fx || !heat2_fx) return fail(TMC_ERR_BAD_ARG, "tally pointer is NULL");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    g.info = tmc_run_info{};
    g.info.n_gpus = static_cast<uint32_t>(g.devs.size());
    g.info.philox_rounds = static_cast<uint32_t>(g.opt.philox_rounds);
    double kernel_ms = 0.0;
    const uint64_t max_chunk = 1ull << (62 - pl.sc.heat_shift);   // chunk * 2^heat_shift < 2^62
    std::vector<unsigned long long> sum;
    uint64_t done = 0;
    while (done < n) {
        const uint64_t todo = (n - done < max_chunk) ? n - done : max_chunk;
        uint32_t flush_override = 0;
        for (int attempt = 0;; ++attempt) {
            rc = run_range(p, pl, seed, first + done, todo, flush_override, sum, &kernel_ms);
            if (rc) return rc;
            const int st = check_counters(p, sum.data());
            if (st == TMC_OK) break;
            if (st != TMC_ERR_TALLY_RANGE) return st;
            if (attempt == 3 || g.info.flush_iters <= 1u)
                return fail(TMC_ERR_TALLY_RANGE, "a shared tally exceeded 2^31 within %u iterations", g.info.flush_iters);
            flush_override = g.info.flush_iters / 8u ? g.info.flush_iters / 8u : 1u;
            g.info.retries += 1;
        }
        for (uint32_t s = 0; s < p->shells; ++s) {
            heat_fx[s] += sum[s];
            heat2_fx[s] += sum[p->shells + s];
        }
        g.info.events += sum[2ull * p->shells];
        g.info.photons += sum[2ull * p->shells + 1];
        done += todo;
    }
    g.info.kernel_ms = kernel_ms;
    g.info.call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (g.info.photons != n) return fail(TMC_ERR_CUDA, "internal: simulated %llu photons, expected %llu", (unsigned long long)g.info.photons, (unsigned long long)n);
    return TMC_OK;
}

// n_batches consecutive sub-ranges of [first, first + n), each into its OWN tally arrays, in one pass:
// every batch is sharded over all devices, all launches are enqueued back to back, one reduce and one
// copy bring all batches home (the per-call costs are paid once, not n_batches times).
int run_fx_batches(const tmc_params* p, uint64_t seed, uint64_t first, uint64_t n, uint32_t n_batches, uint64_t* heat_fx, uint64_t* heat2_fx)
{
    if (!g.inited) return fail(TMC_ERR_NO_DEVICE, "tmc_init has not been called (or found no sm_100 GPU)");
    if (!heat_fx || !heat2_fx) return fail(TMC_ERR_BAD_ARG, "tally pointer is NULL");
    if (n_batches < 1u || n_batches > 4096u) return fail(TMC_ERR_BAD_ARG, "n_batches must be 1..4096");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc) return rc;
    if ((n + n_batches - 1) / n_batches > (1ull << (62 - pl.sc.heat_shift)))
        return fail(TMC_ERR_BAD_ARG, "batches of more than 2^%u photons would overflow the 64-bit tallies; use more batches", 62 - pl.sc.heat_shift);
    const auto t0 = std::chrono::steady_clock::now();
    g.info = tmc_run_info{};
    g.info.n_gpus = static_cast<uint32_t>(g.devs.size());
    g.info.philox_rounds = static_cast<uint32_t>(g.opt.philox_rounds);
    std::vector<Slot> slots(n_batches);
    uint64_t lo = first;
    for (uint32_t b = 0; b < n_batches; ++b) {        // the same split as shards.py / run_slots
        const uint64_t cnt = n / n_batches + (b < n % n_batches ? 1 : 0);
        slots[b] = Slot{ lo, cnt };
        lo += cnt;
    }
    const size_t words = 2ull * p->shells + 4ull;
    std::vector<unsigned long long> sum;
    double kernel_ms = 0.0;
    uint32_t flush_override = 0;
    for (int attempt = 0;; ++attempt) {
        rc = run_slots(p, pl, seed, slots, flush_override, sum, &kernel_ms);
        if (rc) return rc;
        int st = TMC_OK;
        for (uint32_t b = 0; b < n_batches && st == TMC_OK; ++b) st = check_counters(p, sum.data() + b * words);
        if (st == TMC_OK) break;
        if (st != TMC_ERR_TALLY_RANGE) return st;
        if (attempt == 3 || g.info.flush_iters <= 1u)
            return fail(TMC_ERR_TALLY_RANGE, "a shared tally exceeded 2^31 within %u iterations", g.info.flush_iters);
        flush_override = g.info.flush_iters / 8u ? g.info.flush_iters / 8u : 1u;
        g.info.retries += 1;
    }
    for (uint32_t b = 0; b < n_batches; ++b) {
        const unsigned long long* w = sum.data() + b * words;
        for (uint32_t sh = 0; sh < p->shells; ++sh) {
            heat_fx[static_cast<size_t>(b) * p->shells + sh] += w[sh];
            heat2_fx[static_cast<size_t>(b) * p->shells + sh] += w[p->shells + sh];
        }
        g.info.events += w[2ull * p->shells];
        g.info.photons += w[2ull * p->shells + 1];
    }
    g.info.kernel_ms = kernel_ms;
    g.info.call_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (g.info.photons != n) return fail(TMC_ERR_CUDA, "internal: simulated %llu photons, expected %llu", (unsigned long long)g.info.photons, (unsigned long long)n);
    return TMC_OK;
}

void accumulate_float(const tmc_params* p, const Plan& pl, const uint64_t* heat_fx, const uint64_t* heat2_fx,
                      float* heats, float* heats_squared)
{
    const double s1 = std::ldexp(1.0, -static_cast<int>(pl.sc.heat_shift));
    const double s2 = std::ldexp(1.0, static_cast<int>(pl.sc.heat2_rshift) - 2 * static_cast<int>(pl.sc.heat_shift));
    for (uint32_t s = 0; s < p->shells; ++s) {
        heats[s] += static_cast<float>(static_cast<double>(heat_fx[s]) * s1);
        heats_squared[s] += static_cast<float>(static_cast<double>(heat2_fx[s]) * s2);
    }
}

}  // namespace

extern "C" {

int tmc_abi_version(void) { return TMC_ABI_VERSION; }
const char* tmc_version(void) { return "tiny_mc_b200 0.4 (sm_100a, stream tmc-stream-4)"; }
const char* tmc_last_error(void) { return g.err.c_str(); }
int tmc_device_count(void) { return g.inited ? static_cast<int>(g.devs.size()) : 0; }

int tmc_finalize(void)
{
    for (Device& d : g.devs) {
        cudaSetDevice(d.id);
        if (d.comm && g.nccl.CommDestroy) g.nccl.CommDestroy(d.comm);
        if (d.d_buf) cudaFree(d.d_buf);
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.ev2) cudaEventDestroy(d.ev2);
        if (d.stream2) cudaStreamDestroy(d.stream2);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    g.devs.clear();
    // cached tables and scratch (recreated lazily by the next call on that device)
    for (int d = 0; d < 64; ++d)
        if (g_directions[d]) {
            cudaSetDevice(d);
            cudaDeviceSynchronize();
            cudaFree(const_cast<float2*>(g_directions[d]));
            g_directions[d] = nullptr;
        }
    for (DepositTable& t : g_deposit_tables) {
        cudaSetDevice(t.device);
        cudaFree(t.d_table);
    }
    g_deposit_tables.clear();
    for (QueueScratch& q : g_queue_scratch) {
        cudaSetDevice(q.device);
        cudaFree(q.d_buf);
    }
    g_queue_scratch.clear();
    if (g.h_pinned) cudaFreeHost(g.h_pinned);
    g.h_pinned = nullptr;
    g.h_words = 0;
    g.inited = false;
    return TMC_OK;
}

static int init_devices(int n_gpus);

int tmc_init(int n_gpus)
{
    tmc_finalize();                     // also releases whatever a failed earlier attempt left behind
    const int rc = init_devices(n_gpus);
    if (rc) {
        const std::string why = g.err;
        tmc_finalize();                 // streams, events and communicators created before the failure
        g.err = why;
    }
    return rc;
}

static int init_devices(int n_gpus)
{
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible < 1)
        return fail(TMC_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (n_gpus <= 0) n_gpus = visible;
    if (n_gpus > visible) return fail(TMC_ERR_NO_DEVICE, "asked for %d GPUs, %d visible", n_gpus, visible);
    g.devs.resize(n_gpus);
    for (int i = 0; i < n_gpus; ++i) {
        Device& d = g.devs[i];
        d = Device{};
        d.id = i;
        int rc = check_device_arch(i);
        if (rc) return rc;
        CUDA_TRY(cudaSetDevice(i));
        CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, i));
        CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d.stream2, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&d.ev0));
        CUDA_TRY(cudaEventCreate(&d.ev1));
        CUDA_TRY(cudaEventCreateWithFlags(&d.ev2, cudaEventDisableTiming));
    }
    if (n_gpus > 1 && g.opt.nccl_reduce) {
        int rc = load_nccl();
        if (rc) return rc;
        std::vector<ncclComm_t> comms(n_gpus);
        std::vector<int> ids(n_gpus);
        for (int i = 0; i < n_gpus; ++i) ids[i] = i;
        NCCL_TRY(g.nccl.CommInitAll(comms.data(), n_gpus, ids.data()));
        for (int i = 0; i < n_gpus; ++i) g.devs[i].comm = comms[i];
    }
    g.inited = true;
    return TMC_OK;
}

int tmc_prepare(const tmc_params* p)
{
    if (!g.inited) return fail(TMC_ERR_NO_DEVICE, "tmc_init has not been called (or found no sm_100 GPU)");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc) return rc;
    const tmc_run_info keep = g.info;
    std::vector<unsigned long long> sum;
    double ms = 0.0;
    // two slots: both streams of every device get their tables, kernel attributes and scratch (result discarded)
    const uint64_t n = 64ull * g.devs.size();
    rc = run_slots(p, pl, 0x7072657061726521ull, std::vector<Slot>{ Slot{ 0, n }, Slot{ n, n } }, 0, sum, &ms);
    g.info = keep;
    return rc;
}

int tmc_set_option(const char* name, long long value)
{
    if (!name) return fail(TMC_ERR_BAD_ARG, "option name is NULL");
    const std::string n(name);
    if (n == "philox_rounds") {
        if (value == 0) value = 10;
        if (value != 7 && value != 10) return fail(TMC_ERR_BAD_ARG, "philox_rounds must be 10 (default) or 7");
        g.opt.philox_rounds = static_cast<int>(value);
    } else if (n == "block_threads") {
        if (value != 0 && value != 128 && value != 256 && value != 512 && value != 768 && value != 1024)
            return fail(TMC_ERR_BAD_ARG, "block_threads must be 0, 128, 256, 512, 768 or 1024");
        g.opt.block_threads = static_cast<int>(value);
    } else if (n == "blocks_per_sm") {
        if (value < 0 || value > 3) return fail(TMC_ERR_BAD_ARG, "blocks_per_sm must be 0..3");
        g.opt.blocks_per_sm = static_cast<int>(value);
    } else if (n == "flush_iters") {
        if (value < 0 || value > 4096) return fail(TMC_ERR_BAD_ARG, "flush_iters must be 0..4096");
        g.opt.flush_iters = static_cast<int>(value);
    } else if (n == "nccl_reduce") {
        g.opt.nccl_reduce = value ? 1 : 0;
    } else if (n == "tally_check_bits") {
        if (value == 0) value = 31;
        if (value < 8 || value > 31) return fail(TMC_ERR_BAD_ARG, "tally_check_bits must be 8..31");
        g.opt.tally_check_bits = static_cast<int>(value);
    } else if (n == "walk_mode") {
        if (value != 0 && value != 1) return fail(TMC_ERR_BAD_ARG, "walk_mode must be 0 (3-D walk) or 1 (radial cross-check)");
        g.opt.walk_mode = static_cast<int>(value);
    } else if (n == "batch_streams") {
        if (value == 0) value = 2;
        if (value != 1 && value != 2) return fail(TMC_ERR_BAD_ARG, "batch_streams must be 1 or 2");
        g.opt.batch_streams = static_cast<int>(value);
    } else if (n == "batch_capacity") {
        if (value == 0) value = 1;
        if (value < 1 || value > 4096) return fail(TMC_ERR_BAD_ARG, "batch_capacity must be 1..4096");
        g.opt.batch_capacity = static_cast<int>(value);
    } else if (n == "tally_layout") {
        if (value < 0 || value > 3)
            return fail(TMC_ERR_BAD_ARG, "tally_layout must be 0 (auto), 1 (plain, per-lane overflow slots), 2 (lane-private) or 3 (plain, saturating clamp)");
        g.opt.tally_layout = static_cast<int>(value);
    } else {
        return fail(TMC_ERR_BAD_ARG, "unknown option '%s'", name);
    }
    return TMC_OK;
}

int tmc_fx_scales(const tmc_params* p, tmc_scales* out)
{
    if (!out) return fail(TMC_ERR_BAD_ARG, "out is NULL");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc) return rc;
    *out = pl.sc;
    return TMC_OK;
}

int tmc_generation_plan(const tmc_params* p, uint32_t max_gen, uint32_t* first_event, uint32_t* n_events, uint32_t* w_start)
{
    if (!first_event || !n_events || !w_start) return -fail(TMC_ERR_BAD_ARG, "NULL pointer");
    Plan pl;
    const int rc = make_plan(p, &pl);
    if (rc) return -rc;
    const uint32_t n = max_gen < pl.n_gen ? max_gen : pl.n_gen;
    for (uint32_t g = 0; g < n; ++g) {
        first_event[g] = pl.gen[g].first_event;
        n_events[g] = pl.gen[g].n_events;
        w_start[g] = pl.gen[g].w_start;
    }
    return static_cast<int>(n);
}

int tmc_fx_accumulate(const tmc_params* p, const uint64_t* heat_fx, const uint64_t* heat2_fx, float* heats, float* heats_squared)
{
    if (!heat_fx || !heat2_fx || !heats || !heats_squared) return fail(TMC_ERR_BAD_ARG, "NULL pointer");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc) return rc;
    accumulate_float(p, pl, heat_fx, heat2_fx, heats, heats_squared);
    return TMC_OK;
}

int tmc_photons_fx(const tmc_params* p, uint64_t seed, uint64_t first_photon, uint64_t n_photons, uint64_t* heat_fx, uint64_t* heat2_fx)
{
    return run_fx(p, seed, first_photon, n_photons, heat_fx, heat2_fx);
}

int tmc_photons_fx_batches(const tmc_params* p, uint64_t seed, uint64_t first_photon, uint64_t n_photons, uint32_t n_batches,
                           uint64_t* heat_fx, uint64_t* heat2_fx)
{
    return run_fx_batches(p, seed, first_photon, n_photons, n_batches, heat_fx, heat2_fx);
}

int tmc_photons(const tmc_params* p, uint64_t seed, uint64_t first_photon, uint64_t n_photons, float* heats, float* heats_squared)
{
    if (!heats || !heats_squared) return fail(TMC_ERR_BAD_ARG, "tally pointer is NULL");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc) return rc;
    std::vector<uint64_t> fx(2ull * p->shells, 0ull);
    rc = run_fx(p, seed, first_photon, n_photons, fx.data(), fx.data() + p->shells);
    if (rc) return rc;
    accumulate_float(p, pl, fx.data(), fx.data() + p->shells, heats, heats_squared);
    return TMC_OK;
}

int tmc_photons_device(const tmc_params* p, uint64_t seed, uint64_t first_photon, uint64_t n_photons, int device, void* d_tallies, void* cuda_stream)
{
    if (!d_tallies) return fail(TMC_ERR_BAD_ARG, "d_tallies is NULL");
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc) return rc;
    if (n_photons > (1ull << (62 - pl.sc.heat_shift)))
        return fail(TMC_ERR_BAD_ARG, "n_photons too large for one device call with heat_shift=%u; split the range", pl.sc.heat_shift);
    int sms = 0;
    rc = device_facts(device, &sms);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    if (n_photons == 0) return TMC_OK;
    LaunchCfg cfg{};
    g.info.gpu_launches = 0;
    rc = enqueue_walk(p, pl, seed, first_photon, n_photons, sms, 0, static_cast<unsigned long long*>(d_tallies),
                      static_cast<cudaStream_t>(cuda_stream), &cfg);
    g.info.blocks_per_gpu = static_cast<uint32_t>(cfg.grid);
    g.info.threads_per_block = static_cast<uint32_t>(cfg.block);
    g.info.flush_iters = cfg.flush_iters;
    g.info.smem_bytes = static_cast<uint32_t>(cfg.smem);
    g.info.philox_rounds = static_cast<uint32_t>(g.opt.philox_rounds);
    return rc;
}

int tmc_device_tallies_check(const tmc_params* p, int device, const void* d_tallies, void* cuda_stream)
{
    if (!p || !d_tallies) return fail(TMC_ERR_BAD_ARG, "NULL pointer");
    unsigned long long counters[4] = { 0ull, 0ull, 0ull, 0ull };
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMemcpyAsync(counters, static_cast<const unsigned long long*>(d_tallies) + 2ull * p->shells, sizeof counters,
                             cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(cuda_stream)));
    CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    if (counters[3] != 0ull)
        return fail(TMC_ERR_CUDA, "internal: the block's shared-memory window does not start where walk_kernel.cuh assumes");
    if (counters[2] != 0ull)
        return fail(TMC_ERR_TALLY_RANGE, "a shared tally exceeded its 32-bit range check: the tallies in this buffer are invalid; "
                                         "zero the buffer and repeat the range with a smaller \"flush_iters\" option");
    return TMC_OK;
}

int tmc_last_run_info(tmc_run_info* out)
{
    if (!out) return fail(TMC_ERR_BAD_ARG, "out is NULL");
    *out = g.info;
    return TMC_OK;
}

}  // extern "C"
