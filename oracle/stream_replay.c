/* oracle/stream_replay.c — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU replay of the PRODUCT's random stream and fixed-point tally arithmetic
 * ("tmc-stream-4"), restated independently in scalar C with libm
 * (log2f / sqrtf, cos / sin / sqrt in double for the direction table) standing in for the GPU's
 * MUFU approximations.
 *
 * Stream layout: photon i draws Philox4x32-R(counter = (i_lo, i_hi, block, 0), key = seed);
 * one block = four words = FOUR scatter events, one word each.  Event e >= 1 of a photon is
 * word e % 4 of block e / 4; pseudo-event 0 (block 0, word 0) is the photon's roulette fate
 * word.  Inside an event word v:
 *   bits  9..31  step:      xi = (m + 1/2) * 2^-23, m = v >> 9                 (23 bits, midpoints)
 *   bits  8..15  cos theta: (2k + 1) / 256 - 1                                 ( 8 bits, midpoints)
 *   bits  0..7   azimuth:   phi = 2 pi k / 256                                 ( 8 bits)
 * (the polar index shares bits 9..15 with the step's seven lowest mantissa bits.)
 * The direction table holds float(-ln2 cos theta), float(-ln2 sin theta), float(cos phi),
 * float(sin phi), computed in double; the step enters as L = log2(xi) <= 0.
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
#include <stdlib.h>
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

#define DIR_ENTRIES 256u

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
static const float ONE_MINUS_HALF_ULP = 0.999999940395355224609375f;   /* 1 - 2^-24 */

/* L = log2(xi), xi = (m + 1/2) 2^-23, m = v >> 9: exact in float (Sterbenz) */
static float log2_xi_of_word(uint32_t v)
{
    bits32 b;
    b.u = 0x3F800000u | (v >> 9);
    return log2f(b.f - ONE_MINUS_HALF_ULP);
}

float orc_step_of_word(uint32_t v)
{
    return -0.693147182464599609375f * log2_xi_of_word(v);
}

double orc_xi_of_word(uint32_t v)
{
    return ((double)(v >> 9) + 0.5) / 8388608.0;
}

/* E[t] and E[t^2] over ALL 2^23 step mantissas, through the float path the replay uses */
void orc_step_moments(double out[2])
{
    double m1 = 0.0, m2 = 0.0;
    for (uint32_t m = 0; m < (1u << 23); ++m) {
        const double t = (double)orc_step_of_word(m << 9);
        m1 += t;
        m2 += t * t;
    }
    out[0] = m1 / 8388608.0;
    out[1] = m2 / 8388608.0;
}

float orc_costheta_of_word(uint32_t v)
{
    return (float)((2.0 * (double)((v >> 8) & 255u) + 1.0) / 256.0 - 1.0);
}

uint64_t orc_replay(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n,
                    uint64_t* heat_fx, uint64_t* heat2_fx)
{
    return orc_replay_mode(o, rounds, seed, first, n, 0, heat_fx, heat2_fx);
}

/* mode 0: the 3-D walk; mode 1: the product's reduced radial cross-check walk ("walk_mode" = 1):
 * r'^2 = r^2 + t^2 + (t r)(2 mu), mu = cos between r and the new direction, uniform on [-1, 1]
 * because scattering is isotropic (photon.c:35-43) — same step, same mu bits, no azimuth. */
static uint64_t replay_core(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n, int mode,
                            uint64_t* heat_fx, uint64_t* heat2_fx, double* per_photon_sq);

uint64_t orc_replay_mode(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n, int mode,
                         uint64_t* heat_fx, uint64_t* heat2_fx)
{
    return replay_core(o, rounds, seed, first, n, mode, heat_fx, heat2_fx, 0);
}

/* The estimator the reference's Error column means but does not compute (tiny_mc.c:64 squares per EVENT,
 * photon.c:31): per_photon_sq[s] += X_s^2 for every photon, X_s = the photon's TOTAL deposit in shell s
 * (weight units).  Var(mean heat[s]) = (sum X_s^2 / N - mean^2) / N.  Checks the batch-means standard
 * error the product reports (TMC_JSON, tmc_photons_fx_batches). */
uint64_t orc_replay_per_photon(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n,
                               uint64_t* heat_fx, uint64_t* heat2_fx, double* per_photon_sq)
{
    return replay_core(o, rounds, seed, first, n, 0, heat_fx, heat2_fx, per_photon_sq);
}

static uint64_t replay_core(const orc_optics* o, uint32_t rounds, uint64_t seed, uint64_t first, uint64_t n, int mode,
                            uint64_t* heat_fx, uint64_t* heat2_fx, double* per_photon_sq)
{
    orc_fx_scales s;
    orc_fx_plan(o, &s);
    const float spm = (float)(1e4 / (double)o->microns_per_shell / (double)(o->mu_a + o->mu_s));   /* photon.c:9 */
    const uint32_t last = o->shells - 1u;
    /* positions in units of the grid radius SHELLS / shells_per_mfp: |r| = 1 is the outer edge of the last
     * shell, so saturating |r|^2 to 1 is the clamp of photon.c:27-29 and shell = trunc(|r| * shell_scale) with
     * shell_scale the largest float below SHELLS */
    const float kappa = (float)((double)spm / (double)o->shells);
    const float shell_scale = nextafterf((float)o->shells, 0.0f);
    const float radial_step = (float)(-0.693147180559945309417232 * (double)kappa);
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    const uint64_t half = s.heat2_rshift ? ((uint64_t)1 << (s.heat2_rshift - 1)) : 0;
    const uint32_t FATE_SURVIVE = 429496729u;            /* floor(0.1 * 2^32) */
    static float pol_c[DIR_ENTRIES], pol_s[DIR_ENTRIES], az_cos[DIR_ENTRIES], az_sin[DIR_ENTRIES];
    static int dir_ready = 0;
    if (!dir_ready) {
        const double ln2 = 0.693147180559945309417232;
        for (uint32_t k = 0; k < DIR_ENTRIES; ++k) {
            const double ct = (2.0 * (double)k + 1.0) / (double)DIR_ENTRIES - 1.0;
            const double st = sqrt(1.0 - ct * ct);
            const double phi = 6.283185307179586476925 * (double)k / (double)DIR_ENTRIES;
            pol_c[k] = (float)(-ln2 * ct);
            pol_s[k] = (float)(-ln2 * st);
            az_cos[k] = (float)cos(phi);
            az_sin[k] = (float)sin(phi);
        }
        dir_ready = 1;
    }
    uint64_t events = 0;
    /* per-photon shell sums (only when per_photon_sq is asked for): own[s] and the list of touched shells */
    uint64_t* own = per_photon_sq ? calloc(o->shells, sizeof(uint64_t)) : 0;
    uint32_t* touched = per_photon_sq ? malloc(o->shells * sizeof(uint32_t)) : 0;
    const double unit = 1.0 / (double)s.weight_one;

    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t p = first + i;
        uint32_t n_touched = 0;
        float x = 0.0f, y = 0.0f, z = 0.0f;   /* photon.c:12-14 */
        uint32_t w = s.weight_one;            /* photon.c:18    */
        uint32_t r[4] = { 0, 0, 0, 0 };
        uint32_t have_blk = 0xFFFFFFFFu, fate = 0;
        for (uint32_t e = 0;; ++e) {
            /* pseudo-event 0 is the fate word; event e >= 1 is word e % 4 of block e / 4 */
            const uint32_t blk = e / 4u, slot = e % 4u;
            if (blk != have_blk) {
                const uint32_t ctr[4] = { (uint32_t)p, (uint32_t)(p >> 32), blk, 0u };
                orc_philox4x32(rounds, ctr, key, r);
                have_blk = blk;
            }
            if (e == 0) {
                fate = r[0];
                continue;
            }
            const uint32_t v = r[slot], kp = (v >> 8) & 255u, ka = v & 255u;
            ++events;
            /* hop (photon.c:21-24): L = log2(xi), t = -ln2 L; -ln2 is folded into the polar table */
            const float L = log2_xi_of_word(v);
            float rad;
            const float pc = pol_c[kp] * kappa, ps = pol_s[kp] * kappa;   /* the step in grid-radius units */
            float r2;
            if (mode == 1) { /* reduced radial walk: x holds the radius, mu = cos(theta_k) */
                const float t = L * radial_step, tmu = L * pc;
                r2 = fmaf(x + x, tmu, fmaf(t, t, x * x));
            } else {
                /* spin (photon.c:35-43, sampled directly, BEFORE the hop so that no direction is
                 * carried between events) and move (photon.c:22-24) */
                const float ts = L * ps;
                x = fmaf(L, pc, x);
                y = fmaf(ts, az_cos[ka], y);
                z = fmaf(ts, az_sin[ka], z);
                r2 = fmaf(z, z, fmaf(y, y, x * x));
            }
            /* drop (photon.c:26-32): the tally sees |r| clamped to the grid radius, the walk does not */
            if (mode == 1) {
                x = sqrtf(r2 > 0.0f ? r2 : 0.0f);
                rad = x > 1.0f ? 1.0f : x;
            } else {
                rad = sqrtf(r2 > 1.0f ? 1.0f : r2);
            }
            const double sf = floor((double)rad * (double)shell_scale);
            const uint32_t shell = (sf >= (double)last) ? last : (uint32_t)sf;
            const uint32_t dep = (uint32_t)(((uint64_t)w * s.absorb_q32 + 0x80000000ull) >> 32);
            w -= dep;
            heat_fx[shell] += dep;
            heat2_fx[shell] += ((uint64_t)dep * dep + half) >> s.heat2_rshift;
            if (own) {
                if (own[shell] == 0 && dep != 0) touched[n_touched++] = shell;
                own[shell] += dep;
            }
            /* roulette (photon.c:45-49) */
            if (w < s.roulette_thr) {
                if (fate >= FATE_SURVIVE)
                    break;
                fate *= 10u;
                w *= 10u;
            }
        }
        for (uint32_t k = 0; k < n_touched; ++k) {
            const double x = (double)own[touched[k]] * unit;
            per_photon_sq[touched[k]] += x * x;
            own[touched[k]] = 0;
        }
    }
    free(own);
    free(touched);
    return events;
}
