/*
 * oracle/ref_harness_audio.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Drives the UNMODIFIED reference FM branch: rf_decimator.c -> resample.c (cic_decimate)
 * -> audio_main.c (audio_fm_demodulator, halfband_decimate x2).  audio_main.c keeps its
 * buffer pool in file-level statics, so it is textually #included here (from where it
 * lies under the reference tree, via -I) instead of being compiled as its own object;
 * no reference source is copied into this repository.
 *
 * audio_fm_demodulator's discriminator / filter state lives in function-local statics
 * (audio_main.c:77-79) that cannot be reset: a caller that wants a fresh stream loads a
 * fresh copy of the shared library (tests do exactly that).
 */
#include <stdint.h>
#include <string.h>

#include "audio_main.c"          /* resolved through -I<reference>/src */
#include "rf_decimator.h"
#include "resample.h"

static int32_t* g_dec_out = NULL;
static int64_t g_dec_cap = 0;
static int64_t g_dec_n = 0;
static float* g_audio_out = NULL;
static int64_t g_audio_cap = 0;
static int64_t g_audio_n = 0;
static int g_overflow = 0;

/* first callback: keep a copy of what cic_decimate produced (rf_decimator.c:31-36) */
void ref_cbb_capture_dec(const cmplx_s32* signal, int len)
{
    if (g_dec_out == NULL)
        return;
    if (g_dec_n + len > g_dec_cap)
    {
        g_overflow = 1;
        return;
    }
    memcpy(g_dec_out + 2 * g_dec_n, signal, sizeof(cmplx_s32) * (size_t) len);
    g_dec_n += len;
}

/* third callback: move every finished audio buffer out of the pool at BUFFER level
 * (audio_get_audio_payload's drain duplicates / skips samples, audio_main.c:53-63) */
void ref_cbb_drain_audio(const cmplx_s32* signal, int len)
{
    (void) signal;
    (void) len;
    pthread_mutex_lock(&audio_mutex);
    while (list_length(full_audio_buffers) > 0)
    {
        float* data = (float*) list_peek(full_audio_buffers);
        if (g_audio_out != NULL)
        {
            if (g_audio_n + audio_buffer_len > g_audio_cap)
                g_overflow = 1;
            else
            {
                memcpy(g_audio_out + g_audio_n, data, sizeof(float) * (size_t) audio_buffer_len);
                g_audio_n += audio_buffer_len;
            }
        }
        list_poll_to_list(full_audio_buffers, used_audio_buffers);
    }
    pthread_mutex_unlock(&audio_mutex);
}

static struct rf_decimator* g_decim = NULL;

/* main.c:197-205 order: audio_init, decimator alloc + parameters, callback registration */
int ref_fm_open(double sample_rate, int down_factor)
{
    if (g_decim != NULL)
        return -10;
    audio_init();
    g_decim = rf_decimator_alloc();
    if (rf_decimator_set_parameters(g_decim, sample_rate, down_factor))
        return -1;
    rf_decimator_add_callback(g_decim, ref_cbb_capture_dec);
    rf_decimator_add_callback(g_decim, audio_fm_demodulator);
    rf_decimator_add_callback(g_decim, ref_cbb_drain_audio);
    return 0;
}

void ref_fm_set_outputs(int32_t* dec, int64_t dec_cap, float* audio, int64_t audio_cap)
{
    g_dec_out = dec;
    g_dec_cap = dec_cap;
    g_dec_n = 0;
    g_audio_out = audio;
    g_audio_cap = audio_cap;
    g_audio_n = 0;
    g_overflow = 0;
}

/* feed the stream in `chunk`-sample pieces, as signal_source.c:29-35 would */
int ref_fm_push(const uint8_t* iq, int64_t n_samples, int chunk)
{
    int64_t pos = 0;
    while (pos < n_samples)
    {
        int len = (int) ((n_samples - pos) < chunk ? (n_samples - pos) : chunk);
        int r = rf_decimator_decimate_cmplx_u8(g_decim, (const cmplx_u8*) (iq + 2 * pos), len);
        if (r)
            return r;
        pos += len;
    }
    return g_overflow ? -3 : 0;
}

int64_t ref_fm_n_decimated(void) { return g_dec_n; }
int64_t ref_fm_n_audio(void) { return g_audio_n; }

void ref_fm_close(void)
{
    if (g_decim != NULL)
        rf_decimator_free(g_decim);
    g_decim = NULL;
    audio_close();
}

/* common_sp.h:40-76 is static inline; give it an address */
float ref_atan2_approx(float y, float x)
{
    return atan2_approx(y, x);
}

/* audio_fm_demodulator on an explicit block, output at buffer level (for unit pins) */
int ref_fm_demodulate_block(const int32_t* signal, int len, float* audio_out)
{
    float* data;
    int n;
    if (full_audio_buffers == NULL)
        audio_init();
    audio_fm_demodulator((const cmplx_s32*) signal, len);
    pthread_mutex_lock(&audio_mutex);
    data = (float*) list_peek(full_audio_buffers);
    n = audio_buffer_len;
    if (data == NULL)
    {
        pthread_mutex_unlock(&audio_mutex);
        return -1;
    }
    memcpy(audio_out, data, sizeof(float) * (size_t) n);
    list_poll_to_list(full_audio_buffers, used_audio_buffers);
    pthread_mutex_unlock(&audio_mutex);
    return n;
}

/* the discriminator output after the limiter (audio_main.c:110-131), before the filters */
int ref_fm_copy_demod(float* out, int max_len)
{
    int n = demod_buffer_len < max_len ? demod_buffer_len : max_len;
    memcpy(out, demod_buffer, sizeof(float) * (size_t) n);
    return n;
}
