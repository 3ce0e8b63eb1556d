/* wtime.c — host timer around the library call (role of reference wtime.c:6-12). */
#define _POSIX_C_SOURCE 199309L
#include "wtime.h"

#include <time.h>

double wtime(void)
{
    struct timespec now;
    clock_gettime(CLOCK_MONOTONIC_RAW, &now);
    return (double)now.tv_sec + 1e-9 * (double)now.tv_nsec;
}
