/* tiny_mc.c — the `headless` benchmark driver on top of the B200 library.
 *
 * Same role, configuration macros and printout as the reference driver
 * (reference tiny_mc.c:34-69); the per-photon loop `for (i < PHOTONS) photon(heat, heat2)`
 * (reference tiny_mc.c:47-49) becomes ONE call of tmc_photons().  Host code stays plain C11.
 *
 * Environment: TMC_GPUS=<n> selects how many GPUs to use (default: all visible); TMC_NCCL=0 sums the
 * per-GPU tally words on the host instead of with ncclReduce; TMC_TRACE=1 prints per-phase host timings;
 * TMC_JSON=<path> additionally writes the exact tallies in machine-readable form (SURVEY §8f
 * rank 1: the float printout of the reference loses digits at large PHOTONS; the reference's
 * own stdout contract is untouched).
 */
#include "params.h"
#include "report.h"
#include "tiny_mc_b200.h"
#include "wtime.h"

#include <assert.h>
#include <inttypes.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

/* caller-owned tallies, zero-initialised by static storage like reference tiny_mc.c:26-27 */
static float heat[SHELLS];
static float heat2[SHELLS];
/* the same tallies as exact fixed-point integers (bit-identical for any GPU count) */
static uint64_t heat_fx[SHELLS];
static uint64_t heat2_fx[SHELLS];

/* Batch-means standard error of the per-photon mean heat of every shell (filled when TMC_JSON is set):
 * the estimator the reference's Error column (tiny_mc.c:64) is not - that one sums squares per EVENT, so it
 * under-estimates the spread of the per-PHOTON shell sums and is NaN-prone.  The variance of
 * TMC_BATCHES batch means estimates Var(mean heat) without bias whatever the correlation between the events of
 * one photon; its own relative precision is 1/sqrt(2 (TMC_BATCHES-1)) = 9 %.  One library call walks all
 * batches (tmc_photons_fx_batches); their sum is bit-identical to the unsplit run. */
#define TMC_BATCHES 64
static double heat_stderr[SHELLS];
static int have_stderr = 0;

static int walk_in_batches(const tmc_params* params, uint64_t seed, uint64_t photons)
{
    uint64_t* b_heat = calloc((size_t)TMC_BATCHES * SHELLS, sizeof(uint64_t));
    uint64_t* b_heat2 = calloc((size_t)TMC_BATCHES * SHELLS, sizeof(uint64_t));
    if (!b_heat || !b_heat2) return TMC_ERR_BAD_ARG;
    const int rc = tmc_photons_fx_batches(params, seed, 0, photons, TMC_BATCHES, b_heat, b_heat2);
    if (rc == TMC_OK) {
        for (unsigned i = 0; i < SHELLS; ++i) {
            double sum = 0.0, sum_sq = 0.0;
            for (unsigned b = 0; b < TMC_BATCHES; ++b) {
                const uint64_t n = photons / TMC_BATCHES + ((uint64_t)b < photons % TMC_BATCHES ? 1 : 0);
                const double m = (double)b_heat[(size_t)b * SHELLS + i] / (double)n;
                heat_fx[i] += b_heat[(size_t)b * SHELLS + i];
                heat2_fx[i] += b_heat2[(size_t)b * SHELLS + i];
                sum += m;
                sum_sq += m * m;
            }
            const double mean = sum / TMC_BATCHES, var = (sum_sq / TMC_BATCHES - mean * mean) * TMC_BATCHES / (TMC_BATCHES - 1.0);
            heat_stderr[i] = sqrt((var > 0.0 ? var : 0.0) / TMC_BATCHES);
        }
        have_stderr = 1;
    }
    free(b_heat);
    free(b_heat2);
    return rc;
}

static void write_json(const char* path, const tmc_params* params, uint64_t seed, uint64_t photons, double seconds)
{
    tmc_scales sc;
    tmc_run_info info;
    FILE* f = fopen(path, "w");
    if (!f || tmc_fx_scales(params, &sc) != TMC_OK || tmc_last_run_info(&info) != TMC_OK) {
        fprintf(stderr, "tiny_mc_b200: cannot write %s\n", path);
        if (f) fclose(f);
        return;
    }
    const double s1 = ldexp(1.0, -(int)sc.heat_shift), s2 = ldexp(1.0, (int)sc.heat2_rshift - 2 * (int)sc.heat_shift);
    fprintf(f, "{\"shells\": %u, \"mu_a\": %.9g, \"mu_s\": %.9g, \"microns_per_shell\": %.9g, \"photons\": %" PRIu64
               ", \"seed\": %" PRIu64 ", \"seconds\": %.9g, \"events\": %" PRIu64 ", \"n_gpus\": %u, \"kernel_ms\": %.6f,\n",
            params->shells, params->mu_a, params->mu_s, params->microns_per_shell, photons, seed, seconds, info.events,
            info.n_gpus, info.kernel_ms);
    fprintf(f, " \"heat_shift\": %u, \"heat2_rshift\": %u,\n \"heat\": [", sc.heat_shift, sc.heat2_rshift);
    for (unsigned i = 0; i < SHELLS; ++i) fprintf(f, "%s%.17g", i ? ", " : "", (double)heat_fx[i] * s1);
    fprintf(f, "],\n \"heat2\": [");
    for (unsigned i = 0; i < SHELLS; ++i) fprintf(f, "%s%.17g", i ? ", " : "", (double)heat2_fx[i] * s2);
    if (have_stderr) {
        fprintf(f, "],\n \"stderr_batches\": %d,\n \"heat_per_photon_stderr_fx_units\": [", TMC_BATCHES);
        for (unsigned i = 0; i < SHELLS; ++i) fprintf(f, "%s%.9g", i ? ", " : "", heat_stderr[i]);
    }
    fprintf(f, "],\n \"heat_fx\": [");
    for (unsigned i = 0; i < SHELLS; ++i) fprintf(f, "%s%" PRIu64, i ? ", " : "", heat_fx[i]);
    fprintf(f, "]}\n");
    fclose(f);
}

int main(void)
{
    const uint64_t photons = (uint64_t)(PHOTONS);
    const tmc_params params = { SHELLS, MU_A, MU_S, (float)(MICRONS_PER_SHELL) };

    tmc_report_heading(stdout, "B200 version (tiny_mc_b200: sm_100a persistent-thread walk, Philox4x32 per photon)",
                       MU_S, MU_A, photons);

    if (getenv("TMC_NCCL")) tmc_set_option("nccl_reduce", atoi(getenv("TMC_NCCL")));   /* 0: sum the per-GPU words on the host */
    if (getenv("TMC_BATCH_STREAMS")) tmc_set_option("batch_streams", atoi(getenv("TMC_BATCH_STREAMS")));
    const char* env = getenv("TMC_GPUS");
    if (tmc_init(env ? atoi(env) : 0) != TMC_OK) { /* one-off, outside the timed region */
        fprintf(stderr, "tiny_mc_b200: %s\n", tmc_last_error());
        return 1;
    }

    if (getenv("TMC_JSON")) tmc_set_option("batch_capacity", TMC_BATCHES);   /* buffers for all batches, sized by tmc_prepare */
    if (tmc_prepare(&params) != TMC_OK) { /* tables, buffers, first collective: also outside the timed region */
        fprintf(stderr, "tiny_mc_b200: %s\n", tmc_last_error());
        return 1;
    }

    const uint64_t seed = (uint64_t)(SEED); /* role of srand(SEED), reference tiny_mc.c:43 */
    const double start = wtime();
    int rc = (getenv("TMC_JSON") && photons >= 64u * TMC_BATCHES) ? walk_in_batches(&params, seed, photons)
                                                                     : tmc_photons_fx(&params, seed, 0, photons, heat_fx, heat2_fx);
    if (rc == TMC_OK) rc = tmc_fx_accumulate(&params, heat_fx, heat2_fx, heat, heat2);   /* the += of photon.c:30-31 */
    const double end = wtime();
    if (rc != TMC_OK) {
        fprintf(stderr, "tiny_mc_b200: %s\n", tmc_last_error());
        return 1;
    }
    assert(start <= end);

    tmc_report_timing(stdout, end - start, photons);
    tmc_report_table(stdout, SHELLS, (float)(MICRONS_PER_SHELL), photons, heat, heat2);
    if (getenv("TMC_JSON")) write_json(getenv("TMC_JSON"), &params, seed, photons, end - start);
    tmc_finalize();
    return 0;
}
