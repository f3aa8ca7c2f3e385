/*
 * oracle/fft_f64.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Double-precision unnormalised forward DFT  X[k] = sum_j x[j] * exp(-2*pi*i*j*k/N)
 * for power-of-two N.  This is the stand-in for the one third-party routine on the
 * reference's hot path: FFTW3's fftw_execute() on a fftw_plan_dft_1d(N, in, out,
 * FFTW_FORWARD, FFTW_ESTIMATE) plan (reference call sites: src/spectrum.c:21 and
 * src/spectrum.c:42).  FFTW3 is linked by the reference as the system -lfftw3
 * (Makefile:21), is not vendored, has no pinned version and is absent from this image,
 * so its published definition (the FFTW manual's "What FFTW Really Computes", forward
 * transform, no normalisation) is restated here and checked against numpy.fft.fft
 * (pocketfft, f64) and a naive long-double DFT in tests/test_oracle.py.
 */
#ifndef ORACLE_FFT_F64_H
#define ORACLE_FFT_F64_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_fft_plan orc_fft_plan;

/* N must be a power of two >= 1.  Returns NULL otherwise. */
orc_fft_plan* orc_fft_plan_create(int N);
void orc_fft_plan_destroy(orc_fft_plan* p);
int orc_fft_plan_size(const orc_fft_plan* p);

/* in/out: N interleaved (re, im) doubles.  Out of place; `in` is preserved. */
void orc_fft_execute(orc_fft_plan* p, const double* in, double* out);

/* O(N^2) long-double DFT, any N; the KAT the fast transform is pinned against. */
void orc_dft_naive(int N, const double* in, double* out);

#ifdef __cplusplus
}
#endif
#endif
