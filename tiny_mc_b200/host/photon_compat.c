/* photon_compat.c — `void photon(float*, float*)` with the reference's exact signature
 * (reference photon.h:3, photon.c:6), for callers that still simulate one packet per call
 * (e.g. the viewer loop, reference cg_mc.c:79-84).  Each call is one GPU launch of one
 * photon: correct, source compatible, and slow by construction — batch with tmc_photons().
 * Compiled with the params.h macros, like the reference's photon.c.
 */
#include "photon.h"

#include "params.h"
#include "tiny_mc_b200.h"

#include <stdio.h>
#include <stdlib.h>

static unsigned long long stream_seed = 0;
static unsigned long long next_photon = 0;
static int ready = 0;

void photon_seed(unsigned long long seed)
{
    stream_seed = seed;
    next_photon = 0;
}

void photon(float* heats, float* heats_squared)
{
    static const tmc_params params = { SHELLS, MU_A, MU_S, (float)(MICRONS_PER_SHELL) };
    if (!ready) {
        if (tmc_init(1) != TMC_OK) {
            fprintf(stderr, "photon(): %s\n", tmc_last_error());
            abort(); /* the reference signature has no error channel, and there is no CPU fallback */
        }
        ready = 1;
    }
    if (tmc_photons(&params, stream_seed, next_photon++, 1, heats, heats_squared) != TMC_OK) {
        fprintf(stderr, "photon(): %s\n", tmc_last_error());
        abort();
    }
}
