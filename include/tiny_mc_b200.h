/* tiny_mc_b200.h — C ABI of the B200-native photon random-walk library.
 *
 * This is the drop-in boundary for ONE hot path of computacionparalela/tiny_mc:
 *
 *     void photon(float *heats, float *heats_squared);            (reference photon.h:3)
 *     for (i = 0; i < PHOTONS; ++i) photon(heat, heat2);          (reference tiny_mc.c:47-49)
 *
 * The reference simulates one photon packet per call on the CPU with libc rand().  This
 * library runs the same walk (reference photon.c:20-50) for a whole RANGE of photons in one
 * call on one or more B200 GPUs, with a counter-based Philox4x32 stream keyed by the global
 * photon index, and ADDS the result into the same caller-owned `float[SHELLS]` tallies.
 *
 * Plain C11, plain pointers and sizes; no CUDA or torch types.  There is no CPU fallback:
 * every entry point that needs a GPU fails with TMC_ERR_NO_DEVICE when none is usable.
 * Not re-entrant (neither is the reference: global rand() state, non-atomic +=).
 */
#ifndef TINY_MC_B200_H
#define TINY_MC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMC_ABI_VERSION 1

/* Status codes (0 = success).  The reference has no error path at all (photon() is void). */
enum {
    TMC_OK = 0,
    TMC_ERR_NO_DEVICE = 1,   /* no usable CUDA device / wrong architecture / not initialised */
    TMC_ERR_BAD_ARG = 2,     /* NULL pointer, SHELLS == 0, MU_A <= 0, n too large ...         */
    TMC_ERR_CUDA = 3,        /* a CUDA runtime call failed (see tmc_last_error)              */
    TMC_ERR_NCCL = 4,        /* NCCL could not be loaded or a collective failed               */
    TMC_ERR_TALLY_RANGE = 5  /* a privatised tally came too close to its 32-bit range         */
};

/* Run-time form of the reference's compile-time configuration macros
 * (reference params.h:5-23).  include/params.h keeps the macros themselves. */
typedef struct tmc_params {
    uint32_t shells;           /* SHELLS: number of radial bins, last one is the overflow bin */
    float mu_a;                /* MU_A: absorption coefficient [1/cm], must be > 0            */
    float mu_s;                /* MU_S: (reduced) scattering coefficient [1/cm]               */
    float microns_per_shell;   /* MICRONS_PER_SHELL: shell thickness [um]                     */
} tmc_params;

/* Fixed-point representation of the tallies (exact, order-independent accumulation):
 *   heat  [s] = heat_fx [s] / 2^heat_shift
 *   heat2 [s] = heat2_fx[s] * 2^heat2_rshift / 2^(2*heat_shift)                              */
typedef struct tmc_scales {
    uint32_t heat_shift;
    uint32_t heat2_rshift;
    uint32_t absorb_q32;       /* round((1-albedo) * 2^32)                */
    uint32_t roulette_thr;     /* fixed-point weight below which roulette is played */
} tmc_scales;

/* What the last tmc_photons* call did (for benchmarks and tests). */
typedef struct tmc_run_info {
    uint64_t photons;          /* photons simulated                                          */
    uint64_t events;           /* scatter events = iterations of reference photon.c:20-50    */
    double kernel_ms;          /* device time of the walk kernel(s), CUDA events, max over GPUs */
    double call_ms;            /* host wall time of the whole call                            */
    uint32_t n_gpus;
    uint32_t gpu_launches;     /* kernels of this library launched by the call, all GPUs      */
    uint32_t blocks_per_gpu;
    uint32_t threads_per_block;
    uint32_t philox_rounds;
    uint32_t flush_iters;
    uint32_t smem_bytes;
    uint32_t retries;          /* relaunches after TMC_ERR_TALLY_RANGE was detected           */
} tmc_run_info;

/* Library / device management.  n_gpus <= 0 selects every visible device.
 * Creates streams and tally buffers; with n_gpus > 1 also one NCCL communicator per device
 * (single process, ncclCommInitAll).  One-off cost, excluded from every timing.            */
int tmc_init(int n_gpus);
/* Optional, after tmc_init: pay the remaining one-off costs for `p` now (device buffers, the
 * azimuth and deposit tables, kernel attributes, the first NCCL collective) by walking 64
 * photons per device into a scratch tally, so that the first timed tmc_photons* call measures
 * the walk.  The reference has no counterpart (its photon() has no set-up).                   */
int tmc_prepare(const tmc_params* p);
int tmc_finalize(void);
int tmc_device_count(void);            /* devices in use after tmc_init, else 0 */
const char* tmc_last_error(void);
const char* tmc_version(void);
int tmc_abi_version(void);

/* Tunables: "philox_rounds" (10 = default, or 7), "block_threads" (128..1024), "blocks_per_sm"
 * (1..4: residency, and with it the register budget of the kernel variant),
 * "flush_iters", "nccl_reduce" (1 = NCCL, 0 = host-side sum; default 1), "tally_layout"
 * (0 = auto, 1 = one histogram per block with per-lane slots for the overflow shell, 2 = one histogram per lane,
 * 3 = one per block with a single overflow word: auto picks it for grids no photon leaves), "tally_check_bits" (31; tests lower
 * it to exercise the TMC_ERR_TALLY_RANGE retry), "walk_mode" (0 = the 3-D walk of reference
 * photon.c:20-50; 1 = a reduced radial walk, r'^2 = r^2 + t^2 + 2 r t mu, distribution-identical for
 * this isotropic problem: a cross-check, never the benchmarked path; default block shape only),
 * "batch_streams" (2 = the launches of tmc_photons_fx_batches alternate between two streams per device so
 * that one launch's tail overlaps the next, 1 = one stream), "batch_capacity" (tally slots tmc_prepare
 * sizes the buffers for, so that a later batched call allocates nothing; default 1).
 * 0 restores the default.                                                                     */
int tmc_set_option(const char* name, long long value);

/* The batched form of the reference call site tiny_mc.c:47-49:
 * simulate photons first_photon .. first_photon + n_photons - 1 of the stream `seed`
 * (the reference seeds with srand(SEED), tiny_mc.c:43) and ADD absorbed weight and its
 * per-event square into heats[SHELLS] / heats_squared[SHELLS] exactly as photon() does
 * (reference photon.c:30-31).  The arrays are caller-owned and are never zeroed.
 * The result does not depend on the number of GPUs, blocks or threads.                      */
int tmc_photons(const tmc_params* p, uint64_t seed, uint64_t first_photon, uint64_t n_photons,
                float* heats, float* heats_squared);

/* Same, but adds the exact fixed-point tallies into caller-owned uint64_t[SHELLS] arrays
 * (bit-reproducible across 1/2/4/8 GPUs and across any split of the photon range).         */
int tmc_photons_fx(const tmc_params* p, uint64_t seed, uint64_t first_photon, uint64_t n_photons,
                   uint64_t* heat_fx, uint64_t* heat2_fx);

/* n_batches consecutive sub-ranges of [first_photon, first_photon + n_photons) in ONE pass, batch b
 * (photons first + b*n/n_batches ..., the remainder spread over the first batches) ADDED into
 * heat_fx[b*SHELLS .. ] / heat2_fx[b*SHELLS ..]: the batch-means form of the driver loop
 * (reference tiny_mc.c:47-49), for a standard error the reference's per-event Error column
 * (tiny_mc.c:64) cannot give.  Every batch is sharded over all GPUs, all kernels are enqueued back
 * to back, one reduce and one copy bring all batches home.  The sum over b is bit-identical to
 * tmc_photons_fx of the whole range.  1 <= n_batches <= 4096.                                   */
int tmc_photons_fx_batches(const tmc_params* p, uint64_t seed, uint64_t first_photon, uint64_t n_photons,
                           uint32_t n_batches, uint64_t* heat_fx, uint64_t* heat2_fx);

/* Device-resident, asynchronous form for one-process-per-GPU hosts (e.g. torchrun ranks):
 * enqueue the walk on `cuda_stream` (a cudaStream_t, NULL = default stream) of CUDA device
 * `device`, ADDING into the DEVICE buffer d_tallies = uint64_t[2*SHELLS + 4] laid out as
 *   [0, SHELLS)            heat_fx
 *   [SHELLS, 2*SHELLS)     heat2_fx
 *   [2*SHELLS + 0..3]      events, photons, tally-range flag (non-zero = invalid), reserved
 * which the caller zeroes once.  No synchronisation and no collective: the caller sums that
 * one buffer across ranks itself (a single NCCL all-reduce of 2*SHELLS+4 int64 words).
 * Works without tmc_init.
 * MANDATORY before the tallies are used: word 2*SHELLS+2 (summed over ranks) must be 0.  The
 * privatised u32 tallies are drained at an interval that provably cannot wrap for SHELLS <= 512
 * and that rests on a 4x statistical margin plus a 2^31 tripwire for larger grids; this entry
 * point cannot repeat a range by itself the way tmc_photons* do, it only raises that word.
 * tmc_device_tallies_check() does the test for hosts that do not read the buffer themselves.  */
int tmc_photons_device(const tmc_params* p, uint64_t seed, uint64_t first_photon, uint64_t n_photons,
                       int device, void* d_tallies, void* cuda_stream);

/* Synchronise `cuda_stream`, read the four counter words of the device buffer d_tallies and return
 * TMC_OK, or TMC_ERR_TALLY_RANGE when the range flag is set (the tallies of this buffer must then be
 * discarded: zero it and repeat with tmc_set_option("flush_iters", smaller)).                    */
int tmc_device_tallies_check(const tmc_params* p, int device, const void* d_tallies, void* cuda_stream);

/* Fixed-point scales used for `p` (a pure function of the optics). */
int tmc_fx_scales(const tmc_params* p, tmc_scales* out);

/* The deterministic weight schedule behind the kernel: generation g (= number
 * of roulettes survived, reference photon.c:45-48) starts at scatter event first_event[g]
 * (1-based), lasts n_events[g] events and starts with fixed-point weight w_start[g].
 * Fills up to max_gen entries, returns the number filled (or a negative status).  Pure host
 * arithmetic: works without a GPU.                                                           */
int tmc_generation_plan(const tmc_params* p, uint32_t max_gen, uint32_t* first_event, uint32_t* n_events,
                        uint32_t* w_start);

/* Convert / accumulate fixed-point tallies into the reference's float arrays (+=). */
int tmc_fx_accumulate(const tmc_params* p, const uint64_t* heat_fx, const uint64_t* heat2_fx,
                      float* heats, float* heats_squared);

int tmc_last_run_info(tmc_run_info* out);

#ifdef __cplusplus
}
#endif
#endif /* TINY_MC_B200_H */
