/* report.h — the reference's stdout contract (reference tiny_mc.c:37-40,55-66). */
#ifndef TMC_REPORT_H
#define TMC_REPORT_H

#include <stdint.h>
#include <stdio.h>

/* Heading block: three title lines, optics, photon count (reference tiny_mc.c:37-40). */
void tmc_report_heading(FILE* out, const char* backend_line, float mu_s, float mu_a, uint64_t photons);

/* Timing lines (reference tiny_mc.c:55-56). */
void tmc_report_timing(FILE* out, double elapsed_s, uint64_t photons);

/* Radial table + "extra" line (reference tiny_mc.c:58-66), with the reference's mixed
 * float/double arithmetic reproduced operand by operand so the printed digits match. */
void tmc_report_table(FILE* out, uint32_t shells, float microns_per_shell, uint64_t photons,
                      const float* heat, const float* heat2);

#endif
