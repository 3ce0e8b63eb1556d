/* oracle/stream_replay.c — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU replay of the PRODUCT's random stream and fixed-point tally arithmetic
 * ("tmc-stream-2"), restated independently in scalar C with libm
 * (log2f / sqrtf, cos / sin in double for the azimuth table) standing in for the GPU's MUFU
 * approximations.
 *
 * Stream layout: photon i draws Philox4x32-R(counter = (i_lo, i_hi, block, 0), key = seed).
 *   block 0 : word 0 = roulette fate word, word 1 = launch direction, words 2-3 = first event;
 *   block b : words 0-1 = one event (step word, direction word), words 2-3 = the next event.
 *
 * What it pins (tests/test_gpu_parity.py):
 *   - integer-exact: scatter events per photon, roulette fates, every deposit value and
 *     therefore sum_s heat_fx[s] and sum_s heat2_fx[s] (pure u32/u64 arithmetic);
 *   - near-exact: per-shell tallies (a shell index can flip only when |r|*shells_per_mfp
 *     is within MUFU error of an integer).
 *
 * The walk itself is the reference's (photon.c:20-50): hop -log(xi) (photon.c:21-24), shell
 * index with truncation and clamp (photon.c:26-29), deposit (1-albedo)*w and its square
 * per event (photon.c:30-31), w *= albedo (photon.c:32), isotropic new direction
 * (photon.c:35-43, sampled directly instead of by rejection), roulette with survival
 * probability 0.1 and x10 boost below w = 0.001 (photon.c:45-49).
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

/* Philox4x32 (Salmon et al., SC'11): multipliers and Weyl key increments. */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32(uint32_t rounds, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t ceil_log2_u64(uint64_t v)
{
    uint32_t b = 0;
    while (b < 63 && ((uint64_t)1 << b) < v)
        ++b;
    return b;
}

void orc_fx_plan(const orc_optics* o, orc_fx_scales* s)
{
    const float albedo = o->mu_s / (o->mu_s + o->mu_a);
    const double absorb = 1.0 - (double)albedo;
    /* weight 1.0 -> 2^heat_shift, chosen so that the largest deposit is in [2^20, 2^21). */
    int hs = 21 - (int)ceil(log2(absorb));
    if (hs > 30) hs = 30;
    if (hs < 16) hs = 16;
    s->heat_shift = (uint32_t)hs;
    s->weight_one = 1u << hs;
    double q = floor(absorb * 4294967296.0 + 0.5);
    if (q > 4294967295.0) q = 4294967295.0;
    if (q < 1.0) q = 1.0;
    s->absorb_q32 = (uint32_t)q;
    const uint64_t dep_max = ((uint64_t)s->weight_one * s->absorb_q32) >> 32;
    const uint32_t bits = ceil_log2_u64(dep_max + 1);
    s->heat2_rshift = (2 * bits > 22) ? 2 * bits - 22 : 0;
    s->roulette_thr = (uint32_t)floor(0.001 * (double)s->weight_one + 0.5);
}

typedef union { uint32_t u; float f; } bits32;

/* New isotropic direction from one word (replaces the rejection loop of photon.c:35-43):
 * cos(theta) uniform from bits 9..31, azimuth index from bits 3..14 into a 4096-entry table of
 * (float)cos, (float)sin of 2 pi i / 4096 evaluated in double. */
static void spin_direction(uint32_t wd, float* dx, float* dy, float* dz)
{
    bits32 cb;
    cb.u = (wd >> 9) | 0x3F800000u;                   /* 1 + m * 2^-23 in [1, 2) */
    const float ct = fmaf(cb.f, 2.0f, -3.0f);         /* exact */
    const float nct = fmaf(cb.f, -2.0f, 3.0f);        /* exact */
    const float st = sqrtf(fmaf(ct, nct, 1.0f));
    const double phi = 6.283185307179586476925 * (double)((wd >> 3) & 4095u) / 4096.0;
    *dx = ct;
    *dy = st * (float)cos(phi);
    *dz = st * (float)sin(phi);
}

uint64_t orc_replay(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n,
                    uint64_t* heat_fx, uint64_t* heat2_fx)
{
    orc_fx_scales s;
    orc_fx_plan(o, &s);
    const float spm = (float)(1e4 / (double)o->microns_per_shell / (double)(o->mu_a + o->mu_s));
    const uint32_t last = o->shells - 1u;
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    const uint64_t half = s.heat2_rshift ? ((uint64_t)1 << (s.heat2_rshift - 1)) : 0;
    const float LN2 = 0.693147182464599609375f;          /* float(ln 2)      */
    const float STEP_BIAS = 22.1807098388671875f;        /* float(32 * ln 2) */
    const uint32_t FATE_SURVIVE = 429496729u;            /* floor(0.1 * 2^32) */
    uint64_t events = 0;

    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t p = first + i;
        float x = 0.0f, y = 0.0f, z = 0.0f;
        float dx = 0.0f, dy = 0.0f, dz = 0.0f;
        uint32_t w = s.weight_one;
        uint32_t fate = 0;
        int alive = 1;
        for (uint32_t blk = 0; alive; ++blk) {
            const uint32_t ctr[4] = { (uint32_t)p, (uint32_t)(p >> 32), blk, 0u };
            uint32_t r[4];
            orc_philox4x32(rounds, ctr, key, r);
            int slot = 0;
            if (blk == 0) { /* birth block: word 0 = roulette fate, word 1 = isotropic launch direction */
                fate = r[0];
                spin_direction(r[1], &dx, &dy, &dz);
                slot = 1;
            }
            for (; slot < 2 && alive; ++slot) {
                const uint32_t ws = r[2 * slot], wd = r[2 * slot + 1];
                ++events;
                /* hop: xi = (ws + 0.5) / 2^32 in float; t = -ln(xi) */
                const float fxi = (float)ws + 0.5f;
                const float t = fmaf(log2f(fxi), -LN2, STEP_BIAS);
                x = fmaf(t, dx, x);
                y = fmaf(t, dy, y);
                z = fmaf(t, dz, z);
                /* drop */
                const float r2 = fmaf(z, z, fmaf(y, y, x * x));
                const float rad = sqrtf(r2);
                double sf = floor((double)rad * (double)spm);
                uint32_t shell = (sf >= (double)last) ? last : (uint32_t)sf;
                /* deposit = round(w * (1-albedo)): one 32x32+64 multiply-add, high word */
                const uint32_t dep = (uint32_t)(((uint64_t)w * s.absorb_q32 + 0x80000000ull) >> 32);
                w -= dep;
                heat_fx[shell] += dep;
                heat2_fx[shell] += ((uint64_t)dep * dep + half) >> s.heat2_rshift;
                /* roulette */
                if (w < s.roulette_thr) {
                    if (fate < FATE_SURVIVE) {
                        fate *= 10u;
                        w *= 10u;
                    } else {
                        alive = 0;
                    }
                }
                spin_direction(wd, &dx, &dy, &dz);
            }
        }
    }
    return events;
}
