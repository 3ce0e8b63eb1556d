/* tiny_mc.c — the `headless` benchmark driver on top of the B200 library.
 *
 * Same role, configuration macros and printout as the reference driver
 * (reference tiny_mc.c:34-69); the per-photon loop `for (i < PHOTONS) photon(heat, heat2)`
 * (reference tiny_mc.c:47-49) becomes ONE call of tmc_photons().  Host code stays plain C11.
 *
 * Environment: TMC_GPUS=<n> selects how many GPUs to use (default: all visible).
 */
#include "params.h"
#include "report.h"
#include "tiny_mc_b200.h"
#include "wtime.h"

#include <assert.h>
#include <stdio.h>
#include <stdlib.h>

/* caller-owned tallies, zero-initialised by static storage like reference tiny_mc.c:26-27 */
static float heat[SHELLS];
static float heat2[SHELLS];

int main(void)
{
    const uint64_t photons = (uint64_t)(PHOTONS);
    const tmc_params params = { SHELLS, MU_A, MU_S, (float)(MICRONS_PER_SHELL) };

    tmc_report_heading(stdout, "B200 version (tiny_mc_b200: sm_100a persistent-thread walk, Philox4x32 per photon)",
                       MU_S, MU_A, photons);

    const char* env = getenv("TMC_GPUS");
    if (tmc_init(env ? atoi(env) : 0) != TMC_OK) { /* one-off, outside the timed region */
        fprintf(stderr, "tiny_mc_b200: %s\n", tmc_last_error());
        return 1;
    }

    const uint64_t seed = (uint64_t)(SEED); /* role of srand(SEED), reference tiny_mc.c:43 */
    const double start = wtime();
    const int rc = tmc_photons(&params, seed, 0, photons, heat, heat2);
    const double end = wtime();
    if (rc != TMC_OK) {
        fprintf(stderr, "tiny_mc_b200: %s\n", tmc_last_error());
        return 1;
    }
    assert(start <= end);

    tmc_report_timing(stdout, end - start, photons);
    tmc_report_table(stdout, SHELLS, (float)(MICRONS_PER_SHELL), photons, heat, heat2);
    tmc_finalize();
    return 0;
}
