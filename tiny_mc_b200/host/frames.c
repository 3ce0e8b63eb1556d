/* frames.c — the incremental use of the tally contract, as the reference viewer does it
 * (reference cg_mc.c:71-87: up to MAX_PHOTONS_PER_FRAME calls of photon() per frame into the
 * same running `heats`, until PHOTON_CAP photons are spent; the display reads `heats` between
 * frames), without the OpenGL part: one batched call per frame, continuing the photon index.
 *
 * Prints, per frame, the photons spent so far and the running total absorbed weight; the last
 * line is the same "# extra" line the headless driver prints.  Because any split of a photon
 * range gives the same fixed-point tallies, the result after the last frame is exactly the
 * one-shot result.
 */
#include "params.h"
#include "tiny_mc_b200.h"

#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>

#ifndef PHOTON_CAP
#define PHOTON_CAP (1 << 16) /* reference cg_mc.c:10 */
#endif
#ifndef PHOTONS_PER_FRAME
#define PHOTONS_PER_FRAME 4096 /* the reference spends 20 per frame on one CPU core (cg_mc.c:11) */
#endif

static uint64_t heat_fx[SHELLS], heat2_fx[SHELLS]; /* running tallies, caller-owned (cg_mc.c:13-14) */

int main(void)
{
    const tmc_params params = { SHELLS, MU_A, MU_S, (float)(MICRONS_PER_SHELL) };
    tmc_scales sc;
    if (tmc_init(1) != TMC_OK || tmc_prepare(&params) != TMC_OK || tmc_fx_scales(&params, &sc) != TMC_OK) {
        fprintf(stderr, "frames: %s\n", tmc_last_error());
        return 1;
    }
    const uint64_t seed = (uint64_t)(SEED);
    uint64_t spent = 0;
    for (unsigned frame = 0; spent < (uint64_t)(PHOTON_CAP); ++frame) { /* cg_mc.c:73,79-84 */
        uint64_t n = (uint64_t)(PHOTON_CAP)-spent;
        if (n > PHOTONS_PER_FRAME) n = PHOTONS_PER_FRAME;
        if (tmc_photons_fx(&params, seed, spent, n, heat_fx, heat2_fx) != TMC_OK) {
            fprintf(stderr, "frames: %s\n", tmc_last_error());
            return 1;
        }
        spent += n;
        uint64_t total = 0;
        for (unsigned i = 0; i < SHELLS; ++i) total += heat_fx[i];
        printf("frame %u\tphotons %" PRIu64 "\tabsorbed/photon %.6f\n", frame, spent,
               (double)total / (double)((uint64_t)1 << sc.heat_shift) / (double)spent);
    }
    printf("# extra\t%12.5f\n", (double)heat_fx[SHELLS - 1] / (double)((uint64_t)1 << sc.heat_shift) / (double)spent);
    /* checksum of the running tallies, for comparison with a one-shot run */
    uint64_t sum = 0;
    for (unsigned i = 0; i < SHELLS; ++i) sum = sum * 1000003u + heat_fx[i] + 31u * heat2_fx[i];
    printf("# checksum\t%" PRIu64 "\n", sum);
    tmc_finalize();
    return 0;
}
