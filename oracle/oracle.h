/* oracle/oracle.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * C interface of the CPU oracle: a restatement of the reference hot path
 * (/root/reference/photon.c:6-51, driven as in /root/reference/tiny_mc.c:43-49)
 * plus a CPU replay of the product's Philox stream.  Never linked into the product.
 */
#ifndef TMC_ORACLE_H
#define TMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Run-time form of the reference's compile-time macros (params.h:5-23). */
typedef struct orc_optics {
    uint32_t shells;            /* SHELLS            params.h:5-7   */
    float mu_a;                 /* MU_A   [1/cm]     params.h:13-15 */
    float mu_s;                 /* MU_S   [1/cm]     params.h:17-19 */
    float microns_per_shell;    /* MICRONS_PER_SHELL params.h:21-23 */
} orc_optics;

/* Signature of the reference entry point (photon.h:3). */
typedef void (*orc_photon_fn)(float* heats, float* heats_squared);

/* ---- photon_port.c : restatement of photon.c with the libc rand() stream ---- */

/* One photon packet, accumulating (+=) into caller-owned float[shells] tallies.
 * Consumes libc rand() exactly as the reference does, so after the same srand()
 * it is bit-identical to the reference object code.  Returns the number of
 * scatter events (iterations of photon.c:20-50). */
uint32_t orc_photon(const orc_optics* o, float* heats, float* heats_squared);

/* Seed the uniform source used by orc_photon (photon_port.c explains why there are two). */
enum { ORC_RNG_LIBC = 0, ORC_RNG_XOSHIRO = 1, ORC_RNG_PCG = 2 };
void orc_seed(int kind, unsigned seed);

/* Per-photon event counter of the last orc_run_* call (sum over photons). */

/* ---- harness.c : batch runner ---- */

/* Seed libc rand() with `seed` (tiny_mc.c:43), then simulate n_photons photons.
 *   fn != NULL : call the given reference-ABI function (e.g. photon() from oracle/_ref);
 *   fn == NULL : call orc_photon(o, ...).
 * chunk == 0  : accumulate like the reference driver does — one long float[shells]
 *               accumulation (tiny_mc.c:26-27,47-49); the float bits are also returned
 *               through heat_f/heat2_f when those are non-NULL.
 * chunk  > 0  : fresh float arrays every `chunk` photons, summed into double (H6).
 * rng_kind: ORC_RNG_LIBC (the reference's stream) or ORC_RNG_XOSHIRO (port only).
 * heat/heat2 are double[shells] outputs (overwritten).  Returns total events when the
 * port ran (0 for fn != NULL). */
uint64_t orc_run_batch(const orc_optics* o, orc_photon_fn fn, int rng_kind, unsigned seed, uint64_t n_photons,
                       uint32_t chunk, double* heat, double* heat2, float* heat_f, float* heat2_f);

/* ---- stream_replay.c : CPU replay of the product's stream "tmc-stream-4" ---- */

typedef struct orc_fx_scales {
    uint32_t weight_one;     /* fixed-point value of weight 1.0 (= 2^heat_shift)        */
    uint32_t heat_shift;     /* heat_fx  = deposit * 2^heat_shift                      */
    uint32_t heat2_rshift;   /* heat2_fx = (deposit_fx^2 + half) >> heat2_rshift       */
    uint32_t absorb_q32;     /* round((1-albedo) * 2^32)                                */
    uint32_t roulette_thr;   /* weight_fx below which roulette is played               */
} orc_fx_scales;

void orc_philox4x32(uint32_t rounds, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Fixed-point plan shared with the product (restated from DESIGN.md §4, not imported). */
void orc_fx_plan(const orc_optics* o, orc_fx_scales* s);

/* Deterministic weight schedule: generation g (= roulettes survived) starts at event
 * first_event[g] (1-based) with weight w_start[g] and lasts n_events[g] events. */
uint32_t orc_generation_plan(const orc_optics* o, uint32_t max_gen, uint32_t* first_event, uint32_t* n_events,
                             uint32_t* w_start);

/* word -> step length [mfp] (bits 9..31), its uniform variate xi, and word -> cos(theta) (bits 8..15) */
float orc_step_of_word(uint32_t v);
double orc_xi_of_word(uint32_t v);
void orc_step_moments(double out[2]);   /* exact enumeration of the 2^23 step values: E[t], E[t^2] */
float orc_costheta_of_word(uint32_t v);

/* Replay photons [first, first+n) of stream `seed`; ADD into u64 heat_fx/heat2_fx[shells].
 * Returns number of scatter events. */
uint64_t orc_replay(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n,
                    uint64_t* heat_fx, uint64_t* heat2_fx);
/* mode 0 = the 3-D walk (orc_replay), 1 = the product's reduced radial cross-check walk. */
uint64_t orc_replay_mode(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n, int mode,
                         uint64_t* heat_fx, uint64_t* heat2_fx);

/* The 3-D replay that also accumulates per_photon_sq[s] += X_s^2, X_s = one photon's total deposit in shell s
 * (weight units): the per-PHOTON second moment behind a correct standard error of heat[s] (tiny_mc.c:64 means it,
 * photon.c:31 accumulates per EVENT instead). */
uint64_t orc_replay_per_photon(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n,
                               uint64_t* heat_fx, uint64_t* heat2_fx, double* per_photon_sq);

#ifdef __cplusplus
}
#endif
#endif
