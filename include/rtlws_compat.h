/*
 * rtlws_compat.h -- the DSP interface of sjappig/rtl-ws as exported by libb200sdr.so.
 *
 * This header is written for this repository; it declares, with identical names, argument
 * order, types and return conventions, the symbols that the reference declares in
 *   src/common_sp.h:7-20     sample types (2-byte cmplx_u8, 8-byte cmplx_s32 union)
 *   src/spectrum.h:9-17      spectrum_*
 *   src/resample.h:6-17      HALF_BAND_N, struct cic_delay_line, cic_decimate, halfband_decimate
 *   src/rf_decimator.h:6-21  rf_decimator_callback, rf_decimator_*
 * so that the reference's own cbb_main.c, audio_main.c and main.c compile against either
 * set of headers and link against this library in place of spectrum.o, resample.o and
 * rf_decimator.o.  If the reference's headers are on the include path
 * first, their include guards make the duplicate blocks below drop out.
 */
#ifndef RTLWS_COMPAT_H
#define RTLWS_COMPAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- sample types (common_sp.h:7-20) ---- */
#ifndef COMMON_SP_H
#define COMMON_SP_H
typedef struct
{
    uint8_t re;
    uint8_t im;
} cmplx_u8;                    /* wire sample of the dongle: offset binary, 128 == 0 */

typedef union
{
    int64_t bulk;              /* both halves as one word */
    struct p
    {
        int32_t re;
        int32_t im;
    } p;
} cmplx_s32;
#endif /* COMMON_SP_H */

/* ---- power spectrum (spectrum.h:9-17) ---- */
#ifndef SPECTRUM_H
#define SPECTRUM_H
struct spectrum;

/* N-point plan; NULL if N is not a power of two in [16, 65536] or no CUDA device answers */
struct spectrum* spectrum_alloc(int N);

/* Each call transforms ONE frame of exactly N samples and adds its fftshifted power to the
 * caller's double[N] (read-modify-write; the caller zeroes it).  Returns 0, -1 when
 * len != N, -2 on a CUDA failure. */
int spectrum_add_cmplx_u8(struct spectrum* s, const cmplx_u8* src, double* power_spectrum, int len);
int spectrum_add_cmplx_s32(struct spectrum* s, const cmplx_s32* src, double* power_spectrum, int len);
int spectrum_add_real_f32(struct spectrum* s, const float* src, double* power_spectrum, int len);

void spectrum_free(struct spectrum* s);
#endif /* SPECTRUM_H */

/* ---- decimators (resample.h:6-17) ---- */
#ifndef RESAMPLE_H
#define RESAMPLE_H
#define HALF_BAND_N 11

struct cic_delay_line
{
    cmplx_s32 integrator_prev_out;
    cmplx_s32 comb_prev_in;
};

/* 0 ok; -1 when dst_len * R != src_len; -2 on a CUDA failure */
int cic_decimate(int R, const cmplx_u8* src, int src_len, cmplx_s32* dst, int dst_len, struct cic_delay_line* delay);

/* input holds 2 * output_len floats; delay holds HALF_BAND_N - 1 floats, in and out */
void halfband_decimate(const float* input, float* output, int output_len, float* delay);
#endif /* RESAMPLE_H */

/* ---- RF decimator (rf_decimator.h:6-21) ---- */
#ifndef RF_DECIMATOR_H
#define RF_DECIMATOR_H
typedef void (*rf_decimator_callback)(const cmplx_s32*, int);

struct rf_decimator;

struct rf_decimator* rf_decimator_alloc();
void rf_decimator_add_callback(struct rf_decimator* d, rf_decimator_callback callback);
/* 0 ok, -1 for non-positive arguments (or a CUDA allocation failure) */
int rf_decimator_set_parameters(struct rf_decimator* d, double sample_rate, int down_factor);
/* 0 ok, -1 if parameters were never set, -2 on a CUDA failure.  Callbacks run on the calling
 * thread with a pointer into the decimator's pinned host buffer, valid until they return. */
int rf_decimator_decimate_cmplx_u8(struct rf_decimator* d, const cmplx_u8* complex_signal, int len);
void rf_decimator_remove_callbacks(struct rf_decimator* d);
void rf_decimator_free(struct rf_decimator* d);
#endif /* RF_DECIMATOR_H */

#ifdef __cplusplus
}
#endif
#endif /* RTLWS_COMPAT_H */
