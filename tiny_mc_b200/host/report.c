/* report.c — prints what the reference's `headless` prints (reference tiny_mc.c:37-40,55-66).
 *
 * The table arithmetic follows the reference's C promotions exactly:
 *   t      : float  <- 4.0f * M_PI(double) * powf(MPS, 3.0f) * PHOTONS / 1e12   (tiny_mc.c:60)
 *   heat   : float / float, then / double (i*i + i + 1.0/3.0)                  (tiny_mc.c:63)
 *   error  : sqrt(double <- float expr) / t / float (i*i + i + 1.0f/3.0f)      (tiny_mc.c:64)
 * PHOTONS is 64-bit here; the field widths of the reference are kept.
 */
#define _XOPEN_SOURCE 500 /* M_PI under -std=c11, as reference tiny_mc.c:8 */
#include "report.h"

#include <inttypes.h>
#include <math.h>

void tmc_report_heading(FILE* out, const char* backend_line, float mu_s, float mu_a, uint64_t photons)
{
    fprintf(out, "# %s\n# %s\n# %s\n", "Tiny Monte Carlo by Scott Prahl (http://omlc.ogi.edu)",
            "1 W Point Source Heating in Infinite Isotropic Scattering Medium", backend_line);
    fprintf(out, "# Scattering = %8.3f/cm\n", mu_s);
    fprintf(out, "# Absorption = %8.3f/cm\n", mu_a);
    fprintf(out, "# Photons    = %8" PRIu64 "\n#\n", photons);
}

void tmc_report_timing(FILE* out, double elapsed_s, uint64_t photons)
{
    fprintf(out, "# %lf seconds\n", elapsed_s);
    fprintf(out, "# %lf K photons per second\n", 1e-3 * (double)photons / elapsed_s);
}

void tmc_report_table(FILE* out, uint32_t shells, float microns_per_shell, uint64_t photons,
                      const float* heat, const float* heat2)
{
    fprintf(out, "# Radius\tHeat\n");
    fprintf(out, "# [microns]\t[W/cm^3]\tError\n");
    const float t = 4.0f * M_PI * powf(microns_per_shell, 3.0f) * photons / 1e12;
    const float n = (float)photons;
    for (unsigned int i = 0; i + 1 < shells; ++i) {
        fprintf(out, "%6.0f\t%12.5f\t%12.5f\n", i * microns_per_shell,
                heat[i] / t / (i * i + i + 1.0 / 3.0),
                sqrt(heat2[i] - heat[i] * heat[i] / n) / t / (i * i + i + 1.0f / 3.0f));
    }
    fprintf(out, "# extra\t%12.5f\n", heat[shells - 1] / n);
}
