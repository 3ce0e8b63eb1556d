/* photon.h — source compatibility with the reference entry point (reference photon.h:3). */
#ifndef TMC_PHOTON_H
#define TMC_PHOTON_H
void photon(float* heats, float* heats_squared);
/* role of srand(): choose the stream and restart the photon counter */
void photon_seed(unsigned long long seed);
#endif
