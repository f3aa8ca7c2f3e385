/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the IQ hot path of sjappig/rtl-ws, with all state made
 * explicit so that many streams can be checked in one process.  Every function cites
 * the reference file:line it follows (paths are relative to the reference's src/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call this; the product (libb200sdr.so) never does.
 *
 * Pinning: the reference has no tests or golden vectors ("parity unpinned" by the
 * reference itself), so this port is pinned against the UNMODIFIED reference sources
 * compiled in place (oracle/_ref/libref_rtlws.so, built by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py, and against the golden vectors that build emitted
 * (tests/golden/, made by tests/golden/make_golden.py).
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_HALF_BAND_N 11          /* resample.h:6 */

/* ---- spectrum.c --------------------------------------------------------------- */

typedef struct orc_spectrum orc_spectrum;

orc_spectrum* orc_spectrum_alloc(int N);                                  /* spectrum.c:37-45 */
void orc_spectrum_free(orc_spectrum* s);                                  /* spectrum.c:101-107 */
/* optional per-sample window, multiplied in after the unpack (an extension; NULL = the
 * reference's rectangular behaviour) */
void orc_spectrum_set_window(orc_spectrum* s, const double* window);
int orc_spectrum_add_cmplx_u8(orc_spectrum* s, const uint8_t* iq, double* ps, int len);   /* spectrum.c:47-63 */
int orc_spectrum_add_cmplx_s32(orc_spectrum* s, const int32_t* iq, double* ps, int len);  /* spectrum.c:65-81 */
int orc_spectrum_add_real_f32(orc_spectrum* s, const float* x, double* ps, int len);      /* spectrum.c:83-99 */

/* Batched driver in the shape of cbb_main.c:48-59: for every output row r, zero the row,
 * then add K frames starting at sample r*row_hop + j*hop.  rows: [n_rows][N] doubles. */
int orc_spectrum_rows_cmplx_u8(orc_spectrum* s, const uint8_t* iq, int64_t n_samples,
                               int hop, int K, int64_t row_hop, double* rows, int64_t n_rows);

/* cbb_main.c:106-135: linear gain with INTEGER gain_db/10, 10*log10(|g*P/count|),
 * truncation toward zero, clamp to [0,255].  db_float (nullable) receives the value
 * before the (int) cast. */
void orc_db_payload(const double* ps, int n, int count, int gain_db, uint8_t* out, double* db_float);

/* ---- resample.c --------------------------------------------------------------- */

typedef struct
{
    int32_t integrator_prev_out[2];   /* resample.h:8-12; (re, im) */
    int32_t comb_prev_in[2];
} orc_cic_state;

int orc_cic_decimate(int R, const uint8_t* src, int src_len, int32_t* dst, int dst_len,
                     orc_cic_state* delay);                               /* resample.c:6-45 */
void orc_halfband_decimate(const float* input, float* output, int output_len,
                           float* delay);                                 /* resample.c:47-67 */

/* ---- common_sp.h -------------------------------------------------------------- */

float orc_atan2_approx(float y, float x);                                 /* common_sp.h:40-76 */

/* ---- audio_main.c:74-145 with the function-local statics made explicit --------- */

typedef struct
{
    float prev_sample;                       /* audio_main.c:79 */
    float delay_line_1[ORC_HALF_BAND_N - 1]; /* audio_main.c:77 */
    float delay_line_2[ORC_HALF_BAND_N - 1]; /* audio_main.c:78 */
} orc_fm_state;

/* signal: len interleaved (re, im) int32.  demod: len floats (scratch / inspectable),
 * work: len/2 floats, audio: len/4 floats. */
void orc_fm_demodulate(const int32_t* signal, int len, orc_fm_state* st,
                       float* demod, float* work, float* audio);

/* ---- rf_decimator.c + audio_main.c chained, one explicit-state stream ----------- */

typedef struct orc_chain orc_chain;

/* rf_decimator.c:53-78: out_len = (int)((fs/R)*100/1000), in_len = out_len*R.
 * Returns NULL for non-positive arguments (rf_decimator.c:55-58,77). */
orc_chain* orc_chain_create(double sample_rate, int down_factor);
void orc_chain_free(orc_chain* c);
int orc_chain_block_in(const orc_chain* c);     /* input_signal_len  */
int orc_chain_block_out(const orc_chain* c);    /* resampled_signal_len */
/* rf_decimator.c:80-119 re-blocking, then CIC + FM demodulation of every completed block.
 * Appends to the caller's buffers; *n_dec / *n_audio are running element counts
 * (complex decimated samples / audio floats).  dec may be NULL. */
int orc_chain_push(orc_chain* c, const uint8_t* iq, int len,
                   int32_t* dec, int64_t* n_dec, int64_t dec_cap,
                   float* audio, int64_t* n_audio, int64_t audio_cap);

/* ---- extensions without a reference counterpart (csrc/audio_post.cu); definitions, not ports ---- */
#define ORC_RESAMPLE_TAPS 240
void orc_deemphasis(const float* x, int n, float alpha, float* state, float* y);
float orc_deemphasis_alpha(double rate_hz, double tau_s);
void orc_resample_taps(float* h);
int orc_resample_15_16(const float* x, int n, float* hist, float* y);

#ifdef __cplusplus
}
#endif
#endif
