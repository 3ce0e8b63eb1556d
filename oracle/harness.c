/* oracle/harness.c — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Batch runner around the reference entry point `photon()` (photon.h:3) or its port.
 * It does what the reference driver does — srand(SEED) once, then a plain loop of
 * photon() calls (tiny_mc.c:43,47-49) — but can hand photon() FRESH float tallies every
 * `chunk` photons and sum those chunks in double, because one long float accumulation as
 * in tiny_mc.c:26-27 loses 1.2e-3 of the absorbed weight at 2^20 photons.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

uint64_t orc_run_batch(const orc_optics* o, orc_photon_fn fn, int rng_kind, unsigned seed, uint64_t n_photons,
                       uint32_t chunk, double* heat, double* heat2, float* heat_f, float* heat2_f)
{
    const uint32_t shells = o->shells;
    float* h = calloc(shells, sizeof(float));
    float* h2 = calloc(shells, sizeof(float));
    uint64_t events = 0;
    memset(heat, 0, shells * sizeof(double));
    memset(heat2, 0, shells * sizeof(double));

    orc_seed(fn ? ORC_RNG_LIBC : rng_kind, seed); /* srand(SEED), tiny_mc.c:43 */
    uint64_t done = 0;
    while (done < n_photons) {
        uint64_t todo = n_photons - done;
        if (chunk && todo > chunk)
            todo = chunk;
        for (uint64_t i = 0; i < todo; ++i) { /* tiny_mc.c:47-49 */
            if (fn)
                fn(h, h2);
            else
                events += orc_photon(o, h, h2);
        }
        done += todo;
        if (chunk) {
            for (uint32_t s = 0; s < shells; ++s) {
                heat[s] += (double)h[s];
                heat2[s] += (double)h2[s];
                h[s] = 0.0f;
                h2[s] = 0.0f;
            }
        }
    }
    if (!chunk) {
        for (uint32_t s = 0; s < shells; ++s) {
            heat[s] = (double)h[s];
            heat2[s] = (double)h2[s];
        }
        if (heat_f) memcpy(heat_f, h, shells * sizeof(float));
        if (heat2_f) memcpy(heat2_f, h2, shells * sizeof(float));
    }
    free(h);
    free(h2);
    return events;
}
