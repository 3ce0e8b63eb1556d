/* wtime.h — monotonic wall clock in seconds (same contract as reference wtime.h:2). */
#ifndef TMC_WTIME_H
#define TMC_WTIME_H
double wtime(void);
#endif
