/* oracle/photon_port.c — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Plain-C restatement of the reference hot path, one photon packet per call:
 *   /root/reference/photon.c:6-51   (photon(): hop / drop / spin / roulette)
 * with the compile-time macros of /root/reference/params.h:5-23 turned into run-time
 * fields.  It consumes libc rand() in the reference's order, and evaluates every float
 * expression with the reference's operand order and precision, so that after the same
 * srand() it reproduces the reference object code bit for bit (pinned against
 * oracle/_ref and tests/golden by tests/test_oracle_pinned.py).
 *
 * Build with -std=c11 (ISO mode => -ffp-contract=off, as the reference Makefile:5 does)
 * so that no FMA contraction changes the float results.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>

/* ---- source of the 31-bit integers the reference takes from libc rand() -----------------
 * ORC_RNG_LIBC    : libc rand() itself — the reference's stream, bit for bit (default).
 * ORC_RNG_XOSHIRO : xoshiro256** >> 33, same 31-bit range.  Exists because glibc's rand()
 *                   is the additive-feedback generator r[i] = r[i-3] + r[i-31]; its 3-point
 *                   correlation measurably biases THIS walk (inner shells -0.3 %, outer shells
 *                   +1.5 %, > 6 sigma at 4e6 photons; DESIGN.md §7, tests/test_oracle_pinned.py).
 * ORC_RNG_PCG     : PCG32 >> 1 (pcg31.c), the generator oracle/_ref/libphoton_pcg_*.so binds the
 *                   UNMODIFIED photon.c to: port on PCG == reference on PCG, bit for bit.
 *                   The walk code below is byte-identical for all sources. */
static uint64_t xo[4];
static inline uint64_t rotl64(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
static int xoshiro31(void)
{
    const uint64_t out = rotl64(xo[1] * 5u, 7) * 9u;
    const uint64_t t = xo[1] << 17;
    xo[2] ^= xo[0];
    xo[3] ^= xo[1];
    xo[1] ^= xo[2];
    xo[0] ^= xo[3];
    xo[2] ^= t;
    xo[3] = rotl64(xo[3], 45);
    return (int)(out >> 33);
}
int pcg31(void);               /* pcg31.c: the generator the unmodified reference is ALSO compiled against */
void pcg31_seed(uint64_t seed);
static int (*draw31)(void) = rand;

void orc_seed(int kind, unsigned seed)
{
    if (kind == ORC_RNG_XOSHIRO) {
        uint64_t z = (uint64_t)seed * 0x9E3779B97F4A7C15ull + 1u; /* splitmix64 expansion */
        for (int i = 0; i < 4; ++i) {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t v = z;
            v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
            v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
            xo[i] = v ^ (v >> 31);
        }
        draw31 = xoshiro31;
    } else if (kind == ORC_RNG_PCG) {
        pcg31_seed(seed);
        draw31 = pcg31;
    } else {
        srand(seed); /* tiny_mc.c:43 */
        draw31 = rand;
    }
}

/* photon.c:21,46 — `rand() / (float)RAND_MAX`; (float)RAND_MAX == 2^31 on glibc. */
static inline float unit_uniform(void)
{
    return draw31() / (float)RAND_MAX;
}

/* photon.c:37-38 — `2.0f * rand() / (float)RAND_MAX - 1.0f` (multiply first, then divide). */
static inline float symmetric_uniform(void)
{
    return 2.0f * draw31() / (float)RAND_MAX - 1.0f;
}

uint32_t orc_photon(const orc_optics* o, float* heats, float* heats_squared)
{
    /* photon.c:8 — float expression. */
    const float albedo = o->mu_s / (o->mu_s + o->mu_a);
    /* photon.c:9 — `1e4` is a double literal: the quotient is formed in double and then
     * rounded to float; the (MU_A + MU_S) sum itself is a float sum. */
    const float shells_per_mfp = (float)(1e4 / (double)o->microns_per_shell / (double)(o->mu_a + o->mu_s));
    const unsigned last_shell = o->shells - 1u;

    /* photon.c:12-18 — position in mean-free-path units, initial direction +z, weight 1. */
    float pos[3] = { 0.0f, 0.0f, 0.0f };
    float dir[3] = { 0.0f, 0.0f, 1.0f };
    float weight = 1.0f;
    uint32_t events = 0;

    for (;;) {
        ++events;

        /* hop — photon.c:21-24 */
        const float step = -logf(unit_uniform());
        pos[0] += step * dir[0];
        pos[1] += step * dir[1];
        pos[2] += step * dir[2];

        /* drop — photon.c:26-32: truncate radius to a shell, clamp to the overflow bin,
         * deposit (1-albedo)*w and its square (per EVENT), then attenuate. */
        unsigned shell = sqrtf(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]) * shells_per_mfp;
        if (shell > last_shell)
            shell = last_shell;
        heats[shell] += (1.0f - albedo) * weight;
        heats_squared[shell] += (1.0f - albedo) * (1.0f - albedo) * weight * weight;
        weight *= albedo;

        /* spin — photon.c:35-43: Marsaglia rejection on the unit disc. */
        float a, b, t;
        do {
            a = symmetric_uniform();
            b = symmetric_uniform();
            t = a * a + b * b;
        } while (1.0f < t);
        dir[0] = 2.0f * t - 1.0f;
        dir[1] = a * sqrtf((1.0f - dir[0] * dir[0]) / t);
        dir[2] = b * sqrtf((1.0f - dir[0] * dir[0]) / t);

        /* roulette — photon.c:45-49: survive with probability 0.1, boosted by 1/0.1f. */
        if (weight < 0.001f) {
            if (unit_uniform() > 0.1f)
                break;
            weight /= 0.1f;
        }
    }
    return events;
}
