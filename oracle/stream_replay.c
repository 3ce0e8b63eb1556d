/* oracle/stream_replay.c — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU replay of the PRODUCT's random stream and fixed-point tally arithmetic
 * ("tmc-stream-3"), restated independently in scalar C with libm
 * (log2f / sqrtf, cos / sin in double for the azimuth table) standing in for the GPU's MUFU
 * approximations.
 *
 * Stream layout: photon i draws Philox4x32-R(counter = (i_lo, i_hi, block, 0), key = seed);
 * one block = four words = THREE scatter events of 42 bits each.  Event e >= 1 of a photon is
 * slot e % 3 of block e / 3 and uses word[slot] plus bits [10*slot, 10*slot+10) of word[3];
 * pseudo-event 0 (block 0, slot 0) is the photon's roulette fate word.  Inside an event word v:
 *   bits 10..31  step:      xi = 2 * (1.5 - f), f = 1 + (v >> 10) * 2^-23      (22 bits)
 *   bits  1..9   cos theta: (2k + 1) / 512 - 1                                 ( 9 bits, midpoints)
 *   word[3] slice: azimuth index into a 1024-entry (cos, sin) table            (10 bits)
 * The direction is drawn at the START of an event (spin, then hop), so no direction is carried
 * from one event to the next; the first event's direction is the isotropic launch direction.
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
    /* weight 1.0 -> 2^heat_shift, chosen so that the largest deposit is in [2^17, 2^18). */
    int hs = 18 - (int)ceil(log2(absorb));
    if (hs > 30) hs = 30;
    if (hs < 14) hs = 14;
    s->heat_shift = (uint32_t)hs;
    s->weight_one = 1u << hs;
    double q = floor(absorb * 4294967296.0 + 0.5);
    if (q > 4294967295.0) q = 4294967295.0;
    if (q < 1.0) q = 1.0;
    s->absorb_q32 = (uint32_t)q;
    const uint64_t dep_max = ((uint64_t)s->weight_one * s->absorb_q32) >> 32;
    const uint32_t bits = ceil_log2_u64(dep_max + 1);
    s->heat2_rshift = (2 * bits > 18) ? 2 * bits - 18 : 0;
    s->roulette_thr = (uint32_t)floor(0.001 * (double)s->weight_one + 0.5);
}

typedef union { uint32_t u; float f; } bits32;

#define AZIMUTH_ENTRIES 1024u

/* Deterministic weight schedule: every photon of generation g (= number of
 * roulettes survived) starts it with the same weight and needs the same number of events
 * to fall below the roulette threshold.  Restated here independently of the product. */
uint32_t orc_generation_plan(const orc_optics* o, uint32_t max_gen, uint32_t* first_event, uint32_t* n_events,
                             uint32_t* w_start)
{
    orc_fx_scales s;
    orc_fx_plan(o, &s);
    uint32_t w = s.weight_one, e = 1, g = 0;
    for (; g < max_gen; ++g) {
        uint32_t k = 0;
        first_event[g] = e;
        w_start[g] = w;
        do {
            w -= (uint32_t)(((uint64_t)w * s.absorb_q32 + 0x80000000ull) >> 32);
            ++k;
        } while (w >= s.roulette_thr && k < 0x400000u);
        n_events[g] = k;
        e += k;
        w *= 10u;
    }
    return g;
}

/* The two word -> variate mappings of the stream, exported for edge-case tests. */
float orc_step_of_word(uint32_t v)
{
    bits32 fb;
    fb.u = 0x3F800000u | (v >> 10);
    return fmaf(log2f(1.5f - fb.f), -0.693147182464599609375f, -0.693147182464599609375f);
}

float orc_costheta_of_word(uint32_t v)
{
    return fmaf((float)(8388608u + ((v & 0x3FEu) | 1u)), 0.001953125f, -16385.0f);
}

uint64_t orc_replay(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n,
                    uint64_t* heat_fx, uint64_t* heat2_fx)
{
    return orc_replay_mode(o, rounds, seed, first, n, 0, heat_fx, heat2_fx);
}

/* mode 0: the 3-D walk; mode 1: the product's reduced radial cross-check walk ("walk_mode" = 1):
 * r'^2 = r^2 + t^2 + (t r)(2 mu), mu = cos between r and the new direction, uniform on [-1, 1]
 * because scattering is isotropic (photon.c:35-43) — same step, same mu bits, no azimuth. */
uint64_t orc_replay_mode(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n, int mode,
                         uint64_t* heat_fx, uint64_t* heat2_fx)
{
    orc_fx_scales s;
    orc_fx_plan(o, &s);
    const float spm = (float)(1e4 / (double)o->microns_per_shell / (double)(o->mu_a + o->mu_s));
    const uint32_t last = o->shells - 1u;
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    const uint64_t half = s.heat2_rshift ? ((uint64_t)1 << (s.heat2_rshift - 1)) : 0;
    const float LN2 = 0.693147182464599609375f;          /* float(ln 2)       */
    const uint32_t FATE_SURVIVE = 429496729u;            /* floor(0.1 * 2^32) */
    static float az_cos[AZIMUTH_ENTRIES], az_sin[AZIMUTH_ENTRIES];
    static int az_ready = 0;
    if (!az_ready) {
        for (uint32_t i = 0; i < AZIMUTH_ENTRIES; ++i) {
            const double phi = 6.283185307179586476925 * (double)i / (double)AZIMUTH_ENTRIES;
            az_cos[i] = (float)cos(phi);
            az_sin[i] = (float)sin(phi);
        }
        az_ready = 1;
    }
    uint64_t events = 0;

    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t p = first + i;
        float x = 0.0f, y = 0.0f, z = 0.0f;   /* photon.c:12-14 */
        uint32_t w = s.weight_one;            /* photon.c:18    */
        uint32_t r[4] = { 0, 0, 0, 0 };
        uint32_t have_blk = 0xFFFFFFFFu, fate = 0;
        for (uint32_t e = 0;; ++e) {
            /* pseudo-event 0 is the fate word; event e >= 1 is slot e % 3 of block e / 3 */
            const uint32_t blk = e / 3u, slot = e % 3u;
            if (blk != have_blk) {
                const uint32_t ctr[4] = { (uint32_t)p, (uint32_t)(p >> 32), blk, 0u };
                orc_philox4x32(rounds, ctr, key, r);
                have_blk = blk;
            }
            if (e == 0) {
                fate = r[0];
                continue;
            }
            const uint32_t v = r[slot], az = (r[3] >> (10u * slot)) & 1023u;
            ++events;
            /* spin (photon.c:35-43, sampled directly, BEFORE the hop so that no direction is
             * carried between events): cos(theta) = (2k+1)/512 - 1 from bits 1..9, azimuth
             * from a 1024-entry table indexed by 10 bits of word 3 */
            /* hop (photon.c:21-24): xi = 2 * (1.5 - f), f = 1 + (v >> 10) * 2^-23 */
            bits32 fb;
            fb.u = 0x3F800000u | (v >> 10);
            const float t = fmaf(log2f(1.5f - fb.f), -LN2, -LN2);
            float rad;
            if (mode == 1) { /* reduced radial walk: x holds the radius, 2 mu = (2k+1)/256 - 2 */
                const float mu2 = fmaf((float)(8388608u + ((v & 0x3FEu) | 1u)), 0.00390625f, -32770.0f);
                const float r2 = fmaf(t * x, mu2, fmaf(t, t, x * x));
                rad = sqrtf(r2 > 0.0f ? r2 : 0.0f);
                x = rad;
            } else {
                const float ct = fmaf((float)(8388608u + ((v & 0x3FEu) | 1u)), 0.001953125f, -16385.0f);
                const float st = sqrtf(fmaf(-ct, ct, 1.0f));
                const float ts = t * st;
                x = fmaf(t, ct, x);
                y = fmaf(ts, az_cos[az], y);
                z = fmaf(ts, az_sin[az], z);
                /* drop (photon.c:26-32) */
                rad = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
            }
            const double sf = floor((double)rad * (double)spm);
            const uint32_t shell = (sf >= (double)last) ? last : (uint32_t)sf;
            const uint32_t dep = (uint32_t)(((uint64_t)w * s.absorb_q32 + 0x80000000ull) >> 32);
            w -= dep;
            heat_fx[shell] += dep;
            heat2_fx[shell] += ((uint64_t)dep * dep + half) >> s.heat2_rshift;
            /* roulette (photon.c:45-49) */
            if (w < s.roulette_thr) {
                if (fate >= FATE_SURVIVE)
                    break;
                fate *= 10u;
                w *= 10u;
            }
        }
    }
    return events;
}
