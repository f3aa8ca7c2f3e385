/*
 * oracle/fftw3_shim/fftw3.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A from-scratch stand-in for the 5 functions, 2 types and 2 macros of the FFTW3 API
 * that the reference's src/spectrum.c uses (spectrum.c:4 include, :10-12 types,
 * :40-42 fftw_malloc / fftw_plan_dft_1d, :21 fftw_execute, :103-105 destroy / free).
 * FFTW3 itself is absent from this image (system -lfftw3 in the reference Makefile:21,
 * unpinned, un-vendored), so the unmodified spectrum.c is compiled against this header
 * and linked with fftw3_shim.c, which forwards to oracle/fft_f64.c.
 * It is NOT FFTW: any CPU spectrum timing built on it says "stand-in FFT".
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct orc_fftw_plan_s* fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

void* fftw_malloc(size_t n);
void fftw_free(void* p);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif
#endif
