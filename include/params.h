/* params.h — compile-time configuration of the photon walk, macro-compatible with the
 * reference's params.h (reference params.h:5-27): the same names, the same defaults and the
 * same override mechanism (-DNAME=value on the compiler command line).  The C host program
 * (tiny_mc_b200/host/tiny_mc.c) is compiled against these macros exactly like the
 * reference's tiny_mc.c and hands them to the CUDA library at run time (tmc_params).
 *
 * One widening: PHOTONS may exceed INT_MAX here (write -DPHOTONS=4294967296ULL); the
 * reference's `unsigned int` loop cannot express that.
 */
#ifndef TMC_PARAMS_H
#define TMC_PARAMS_H

#include <time.h> /* time(), for the default SEED */

#ifndef SHELLS
#define SHELLS 101 /* number of radial bins; the last one is the overflow bin */
#endif

#ifndef PHOTONS
#define PHOTONS 32768 /* photon packets to simulate */
#endif

#ifndef MU_A
#define MU_A 2.0f /* absorption coefficient [1/cm], must be non-zero */
#endif

#ifndef MU_S
#define MU_S 20.0f /* reduced scattering coefficient [1/cm] */
#endif

#ifndef MICRONS_PER_SHELL
#define MICRONS_PER_SHELL 50 /* shell thickness [um] */
#endif

#ifndef SEED
#define SEED (time(NULL)) /* stream seed */
#endif

#endif /* TMC_PARAMS_H */
