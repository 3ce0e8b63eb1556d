/* oracle/pcg31.c — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * A drop-in for libc rand() under the UNMODIFIED reference sources: oracle/Makefile compiles
 * /root/reference/photon.c with `-Drand=pcg31`, so every `rand()` of photon.c:21,37,38,46 calls
 * this function instead of glibc's additive-feedback generator; nothing else of the reference
 * changes (RAND_MAX stays 2^31 - 1, the float conversions stay the reference's).  PCG32
 * (O'Neill 2014: 64-bit LCG state, XSH-RR output permutation), top 31 bits of the 32-bit output.
 * A second sound generator beside the xoshiro256** port (photon_port.c): two unrelated
 * generators that agree with each other while libc rand() disagrees with both is the evidence
 * behind DESIGN.md §8's finding that glibc rand() biases this walk.
 */
#include <stdint.h>

static uint64_t pcg_state = 0x853c49e6748fea9bull;
static uint64_t pcg_inc = 0xda3e39cb94b95bdbull;

int pcg31(void)
{
    const uint64_t old = pcg_state;
    pcg_state = old * 6364136223846793005ull + pcg_inc;
    const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = (uint32_t)(old >> 59u);
    const uint32_t out = (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
    return (int)(out >> 1);
}

/* pcg32_srandom_r(initstate = seed, initseq = seed) */
void pcg31_seed(uint64_t seed)
{
    pcg_state = 0u;
    pcg_inc = (seed << 1u) | 1u;
    (void)pcg31();
    pcg_state += seed * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull;
    (void)pcg31();
}
