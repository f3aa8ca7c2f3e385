/*
 * oracle/ref_harness_cbb.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Runs the UNMODIFIED reference driver end to end without a dongle:
 *   signal_source.c worker thread -> rtl_async_callback -> {log_data_rate, decimate,
 *   estimate_spectrum} (cbb_main.c:86-88) -> spectrum.c / rf_decimator.c / audio_main.c
 * with three seams replaced by this file:
 *   - rtl_sensor.h is implemented here as a synthetic sensor whose rtl_read_async replays
 *     a caller-supplied capture in librtlsdr's default 262144-byte buffers
 *     (rtl_sensor.c:146-153 is the seam; 0,0 => 15 x 262144 B is librtlsdr's default);
 *   - common.h's timestamp() is a virtual clock advanced by the replayed sample count, so
 *     the 250 ms spectrum cadence (cbb_main.c:46) is deterministic;
 *   - the consumer (main.c:74-137, libwebsockets) is replaced by a poll after every buffer.
 * cbb_main.c is #included (from the reference tree, via -I) to read its file statics.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "cbb_main.c"            /* resolved through -I<reference>/src */
#include "audio_main.h"

/* ---- virtual clock (replaces common.c) ---- */
static volatile uint64_t g_virtual_ms = 1000000;
uint64_t timestamp()
{
    return g_virtual_ms;
}

/* ---- synthetic sensor (replaces rtl_sensor.c) ---- */
struct rtl_dev
{
    uint32_t f;
    uint32_t fs;
    double gain;
};

#define SYNTH_USB_BUF_BYTES 262144

static const uint8_t* g_capture = NULL;
static int64_t g_capture_bytes = 0;
static pthread_mutex_t g_go_mutex = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_go_cond = PTHREAD_COND_INITIALIZER;
static int g_go = 0;
static int g_done = 0;
static volatile int g_cancel = 0;

/* outputs collected by the poll */
static uint8_t* g_payload_out = NULL;     /* [max_spectra][FFT_POINTS] */
static double* g_power_out = NULL;        /* [max_spectra][FFT_POINTS] */
static int* g_count_out = NULL;           /* [max_spectra] */
static int g_max_spectra = 0;
static int g_n_spectra = 0;
static int g_gain_db = 0;
/* optional replacement for the consumer poll (ref_harness_ws.c drives main.c's callback from it);
 * with a hook set the audio stays in audio_main.c's pool for the hook to fetch */
static void (*g_consumer_hook)(void) = NULL;
void ref_cbb_set_consumer(void (*hook)(void)) { g_consumer_hook = hook; }

int rtl_init(struct rtl_dev** dev, int dev_index)
{
    (void) dev_index;
    *dev = (struct rtl_dev*) calloc(1, sizeof(struct rtl_dev));
    (*dev)->fs = 2048000;      /* rtl_sensor.c:12 */
    (*dev)->f = 100000000;     /* rtl_sensor.c:13 */
    (*dev)->gain = 25.4;       /* rtl_sensor.c:14 */
    return 0;
}
int rtl_set_frequency(struct rtl_dev* dev, uint32_t f) { dev->f = f; return 0; }
int rtl_set_sample_rate(struct rtl_dev* dev, uint32_t fs) { dev->fs = fs; return 0; }
int rtl_set_gain(struct rtl_dev* dev, double gain) { dev->gain = gain; return 0; }
uint32_t rtl_freq(const struct rtl_dev* dev) { return dev->f; }
uint32_t rtl_sample_rate(const struct rtl_dev* dev) { return dev->fs; }
double rtl_gain(const struct rtl_dev* dev) { return dev->gain; }
void rtl_cancel(struct rtl_dev* dev) { (void) dev; g_cancel = 1; }
void rtl_close(struct rtl_dev* dev) { free(dev); }

static void poll_consumer(void)
{
    if (g_consumer_hook != NULL)
    {
        g_consumer_hook();
        return;
    }
    /* main.c:77-84: if a new spectrum is there, fetch the payload */
    if (cbb_new_spectrum_available() && g_n_spectra < g_max_spectra)
    {
        char buf[8192];
        int len;
        /* the doubles behind the payload, read before the getter clears the flag */
        pthread_mutex_lock(&spectrum_mutex);
        memcpy(g_power_out + (size_t) g_n_spectra * FFT_POINTS, power_spectrum_transfer, sizeof(double) * FFT_POINTS);
        g_count_out[g_n_spectra] = spectrum_averaging_count;
        pthread_mutex_unlock(&spectrum_mutex);
        len = cbb_get_spectrum_payload(buf, (int) sizeof(buf), g_gain_db);
        if (len == FFT_POINTS)
            memcpy(g_payload_out + (size_t) g_n_spectra * FFT_POINTS, buf, FFT_POINTS);
        g_n_spectra++;
    }
}

int rtl_read_async(struct rtl_dev* dev, void (*callback)(unsigned char*, uint32_t, void*), void* user)
{
    int64_t pos = 0;
    unsigned char* usb_buf = (unsigned char*) malloc(SYNTH_USB_BUF_BYTES);

    pthread_mutex_lock(&g_go_mutex);
    while (!g_go)
        pthread_cond_wait(&g_go_cond, &g_go_mutex);
    pthread_mutex_unlock(&g_go_mutex);

    while (!g_cancel && pos + SYNTH_USB_BUF_BYTES <= g_capture_bytes)
    {
        memcpy(usb_buf, g_capture + pos, SYNTH_USB_BUF_BYTES);
        pos += SYNTH_USB_BUF_BYTES;
        /* the buffer holds SYNTH_USB_BUF_BYTES/2 samples taken at dev->fs */
        g_virtual_ms += (uint64_t) (SYNTH_USB_BUF_BYTES / 2) * 1000u / dev->fs;
        callback(usb_buf, SYNTH_USB_BUF_BYTES, user);
        poll_consumer();
    }
    free(usb_buf);

    pthread_mutex_lock(&g_go_mutex);
    g_done = 1;
    pthread_cond_broadcast(&g_go_cond);
    pthread_mutex_unlock(&g_go_mutex);
    return 0;
}

/* ---- audio capture: rf_decimator callback appended after audio_fm_demodulator ---- */
extern void ref_fm_set_outputs(int32_t* dec, int64_t dec_cap, float* audio, int64_t audio_cap);
extern int64_t ref_fm_n_audio(void);
extern int64_t ref_fm_n_decimated(void);
extern void ref_cbb_drain_audio(const cmplx_s32* signal, int len);
extern void ref_cbb_capture_dec(const cmplx_s32* signal, int len);

/*
 * Replays `capture` (interleaved u8 IQ; only whole 262144-byte buffers are delivered)
 * through the reference's own init order (main.c:197-205) and returns the number of
 * spectra the consumer poll collected.
 */
int ref_cbb_run(const uint8_t* capture, int64_t capture_bytes, int gain_db,
                uint8_t* payload_out, double* power_out, int* count_out, int max_spectra,
                int32_t* dec_out, int64_t dec_cap, float* audio_out, int64_t audio_cap)
{
    g_capture = capture;
    g_capture_bytes = capture_bytes;
    g_payload_out = payload_out;
    g_power_out = power_out;
    g_count_out = count_out;
    g_max_spectra = max_spectra;
    g_n_spectra = 0;
    g_gain_db = gain_db;
    g_go = 0;
    g_done = 0;
    g_cancel = 0;

    ref_fm_set_outputs(dec_out, dec_cap, audio_out, audio_cap);

    audio_init();                                        /* main.c:197 */
    cbb_init(192000);                                    /* main.c:199, DECIMATED_TARGET_BW_HZ main.c:23 */
    rf_decimator_add_callback(cbb_rf_decimator(), ref_cbb_capture_dec);
    rf_decimator_add_callback(cbb_rf_decimator(), audio_fm_demodulator);   /* main.c:205 */
    if (g_consumer_hook == NULL)
        rf_decimator_add_callback(cbb_rf_decimator(), ref_cbb_drain_audio);

    pthread_mutex_lock(&g_go_mutex);
    g_go = 1;
    pthread_cond_broadcast(&g_go_cond);
    while (!g_done)
        pthread_cond_wait(&g_go_cond, &g_go_mutex);
    pthread_mutex_unlock(&g_go_mutex);

    cbb_close();
    audio_close();
    return g_n_spectra;
}
