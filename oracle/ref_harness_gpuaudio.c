/*
 * oracle/ref_harness_gpuaudio.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The counterpart of ref_harness_audio.c for the drop-in build in which audio_main.o itself is
 * replaced: audio_init / audio_fm_demodulator / audio_get_audio_payload / audio_close come from the
 * product's libb200audio.so (the reference's audio_main.h interface over the GPU demodulator), so the
 * unmodified cbb_main.c + main.c (ref_harness_cbb.c, ref_harness_ws.c) register and drain a GPU
 * demodulator exactly as main.c:197,205 and main.c:86-110 do.  Exports the same ref_fm_* / ref_cbb_*
 * symbols as ref_harness_audio.c; finished audio leaves the pool at buffer level through
 * b200_audio_take_buffer (the product's own extension) because the pool's lists are not visible here.
 */
#include <stdint.h>
#include <string.h>

#include "audio_main.h"          /* the reference's own header, via -I<reference>/src */
#include "rf_decimator.h"
#include "resample.h"

extern int b200_audio_take_buffer(float* out, int max_floats);
extern int b200_audio_buffer_len(void);
extern void b200_audio_reset_stream(void);

static int32_t* g_dec_out = NULL;
static int64_t g_dec_cap = 0;
static int64_t g_dec_n = 0;
static float* g_audio_out = NULL;
static int64_t g_audio_cap = 0;
static int64_t g_audio_n = 0;
static int g_overflow = 0;

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

void ref_cbb_drain_audio(const cmplx_s32* signal, int len)
{
    (void) signal;
    (void) len;
    for (;;)
    {
        const int n = b200_audio_buffer_len();
        int got;
        if (n <= 0)
            return;
        if (g_audio_out == NULL || g_audio_n + n > g_audio_cap)
        {
            float scratch[16384];
            got = b200_audio_take_buffer(scratch, 16384);
            if (got < 0)
                return;
            if (g_audio_out != NULL)
                g_overflow = 1;
            continue;
        }
        got = b200_audio_take_buffer(g_audio_out + g_audio_n, (int) (g_audio_cap - g_audio_n));
        if (got < 0)
            return;
        g_audio_n += got;
    }
}

static struct rf_decimator* g_decim = NULL;

int ref_fm_open(double sample_rate, int down_factor)
{
    if (g_decim != NULL)
        return -10;
    audio_init();
    b200_audio_reset_stream();          /* the library (and its state) is shared by every harness copy */
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
    b200_audio_reset_stream();
}

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

float ref_atan2_approx(float y, float x)
{
    return atan2_approx(y, x);
}

int ref_fm_demodulate_block(const int32_t* signal, int len, float* audio_out)
{
    audio_init();
    audio_fm_demodulator((const cmplx_s32*) signal, len);
    return b200_audio_take_buffer(audio_out, len / 4);
}

int ref_fm_copy_demod(float* out, int max_len)
{
    (void) out;
    (void) max_len;
    return 0;                           /* the discriminator output stays on the device in this build */
}
