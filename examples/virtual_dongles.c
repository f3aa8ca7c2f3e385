/*
 * examples/virtual_dongles.c -- a C host on top of the two libraries, no Python, no libwebsockets:
 *
 *   N virtual dongles (libb200replay.so: the reference's rtl_sensor.h over a capture)
 *     -> one reader thread each, exactly like signal_source.c's worker (signal_source.c:37-55):
 *        rtl_read_async(dev, callback, ctx), 262144-byte buffers
 *     -> the callback body is b200_stream_push (libb200sdr.so): H2D, fused chain kernel, D2H, asynchronously
 *     -> sinks on the reader thread: the 6-frame averaged payload bytes of cbb_main.c:121-130 (computed on the
 *        GPU) become spectrum messages, audio is packed into the websocket messages of main.c:86-110
 *        (b200_wire_*), and both are counted / checksummed instead of being sent.
 *
 * Each dongle gets its own b200_stream (the push API is single-producer, as the reference's source is).
 *
 *   gcc -O2 -pthread -o examples/virtual_dongles examples/virtual_dongles.c -Iinclude \
 *       -Lrtl-ws_b200 -lb200sdr -lb200replay -Wl,-rpath,$PWD/rtl-ws_b200 -lm
 *   examples/virtual_dongles [n_dongles=8] [seconds_of_iq=2] [realtime=0]
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200sdr.h"
#include "rtl_sensor_replay.h"

#define FS 2048000
#define BATCH 204800                   /* rf_decimator.c:65-66: 100 ms blocks */

struct dongle {
    int index;
    struct rtl_dev* dev;
    b200_stream* gpu;
    uint8_t* capture;
    int64_t capture_bytes;
    /* what the sinks produced */
    int64_t spectrum_messages, spectrum_bytes, audio_messages, audio_bytes, audio_samples;
    double audio_energy;
    float pending_audio[B200_WIRE_AUDIO_FRAGMENTS * B200_WIRE_AUDIO_FRAGMENT_BYTES / 4];
    int n_pending;
    uint8_t message[B200_WIRE_SPECTRUM_HEADER_MAX + 1024];
};

/* FM broadcast-like capture: carrier at 0 Hz, 1 kHz + 5 kHz message, 25 kHz deviation, amplitude 100 */
static void synth_fm(uint8_t* iq, int64_t n, unsigned seed)
{
    double phase = 0.0;
    unsigned x = seed * 2654435761u + 1u;
    for (int64_t i = 0; i < n; ++i) {
        const double t = (double) i / FS;
        const double msg = 0.5 * (sin(2 * M_PI * (1000.0 + 37.0 * (seed % 16)) * t) + sin(2 * M_PI * 5000.0 * t));
        phase += 2 * M_PI * 25000.0 / FS * msg;
        x = x * 1664525u + 1013904223u;
        const double nr = ((x >> 8) & 0xffff) / 65536.0 - 0.5;
        x = x * 1664525u + 1013904223u;
        const double ni = ((x >> 8) & 0xffff) / 65536.0 - 0.5;
        double re = 127.5 + 100.0 * cos(phase) + 4.0 * nr, im = 127.5 + 100.0 * sin(phase) + 4.0 * ni;
        re = re < 0 ? 0 : (re > 255 ? 255 : re);
        im = im < 0 ? 0 : (im > 255 ? 255 : im);
        iq[2 * i] = (uint8_t) lrint(re);
        iq[2 * i + 1] = (uint8_t) lrint(im);
    }
}

/* payload sink: the 6-frame average of cbb_main.c:48-59,121-130, computed on the GPU -> one websocket message */
static void on_payload(void* user, int stream, int64_t first_frame, int k, const uint8_t* payload)
{
    struct dongle* d = (struct dongle*) user;
    (void) stream;
    (void) first_frame;
    (void) k;
    const int n = b200_wire_spectrum_message(d->message, (int) sizeof(d->message), rtl_freq(d->dev), rtl_sample_rate(d->dev), 0,
                                             payload, 1024);
    if (n > 0) {
        d->spectrum_messages++;
        d->spectrum_bytes += n;
    }
}

/* audio sink: 4096 floats per websocket message, eight writes each (main.c:86-110) */
static void on_audio(void* user, int stream, int64_t first_sample, int n, const float* audio)
{
    struct dongle* d = (struct dongle*) user;
    (void) stream;
    (void) first_sample;
    for (int i = 0; i < n; ++i) {
        d->audio_energy += (double) audio[i] * audio[i];
        d->pending_audio[d->n_pending++] = audio[i];
        if (d->n_pending == (int) (sizeof(d->pending_audio) / sizeof(float))) {
            for (int f = 0; f < B200_WIRE_AUDIO_FRAGMENTS; ++f) {
                int32_t off, len, flags;
                b200_wire_audio_fragment(f, &off, &len, &flags);      /* lws_write(wsi, msg + off, len, mode(flags)) */
                d->audio_bytes += len;
            }
            d->audio_messages++;
            d->n_pending = 0;
        }
    }
    d->audio_samples += n;
}

/* the rtl_read_async callback: signal_source.c:29-35 hands the buffer on as cmplx_u8 samples */
static void on_buffer(unsigned char* buf, uint32_t len, void* ctx)
{
    struct dongle* d = (struct dongle*) ctx;
    if (b200_stream_push(d->gpu, 0, buf, (int) (len / 2)) != B200_OK) fprintf(stderr, "push: %s\n", b200_last_error());
}

static void* reader(void* arg)
{
    struct dongle* d = (struct dongle*) arg;
    b200_init(0);                                       /* the CUDA device of this thread */
    rtl_read_async(d->dev, on_buffer, d);
    b200_stream_flush(d->gpu);
    return NULL;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 8;
    const double seconds = argc > 2 ? atof(argv[2]) : 2.0;
    const int realtime = argc > 3 ? atoi(argv[3]) : 0;
    const int64_t n_buffers = (int64_t) (seconds * FS * 2 / B200_REPLAY_BUFFER_BYTES);
    if (n < 1 || n > B200_REPLAY_MAX_DEVICES || n_buffers < 1) return 2;
    if (b200_init(0) != B200_OK) {
        fprintf(stderr, "b200_init: %s\n", b200_last_error());
        return 1;
    }
    struct dongle* ds = (struct dongle*) calloc((size_t) n, sizeof(struct dongle));
    pthread_t* th = (pthread_t*) calloc((size_t) n, sizeof(pthread_t));
    for (int i = 0; i < n; ++i) {
        struct dongle* d = &ds[i];
        d->index = i;
        d->capture_bytes = n_buffers * B200_REPLAY_BUFFER_BYTES;
        d->capture = (uint8_t*) malloc((size_t) d->capture_bytes);
        synth_fm(d->capture, d->capture_bytes / 2, 1000u + (unsigned) i);
        b200_replay_set_capture(i, d->capture, d->capture_bytes, 1, realtime);
        if (rtl_init(&d->dev, i) != 0) return 1;
        d->gpu = b200_stream_create(1, BATCH, 0);
        if (d->gpu == NULL) {
            fprintf(stderr, "b200_stream_create: %s\n", b200_last_error());
            return 1;
        }
        b200_stream_set_sinks(d->gpu, NULL, on_audio, d);              /* no per-frame rows: they stay on the device */
        b200_stream_set_payload_sink(d->gpu, 6, on_payload);          /* FFT_AVERAGE, cbb_main.c:18 */
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < n; ++i) pthread_create(&th[i], NULL, reader, &ds[i]);
    for (int i = 0; i < n; ++i) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    int64_t sm = 0, sb = 0, am = 0, ab = 0, as = 0;
    double energy = 0.0;
    for (int i = 0; i < n; ++i) {
        sm += ds[i].spectrum_messages; sb += ds[i].spectrum_bytes; am += ds[i].audio_messages; ab += ds[i].audio_bytes;
        as += ds[i].audio_samples; energy += ds[i].audio_energy;
    }
    const double samples = (double) n * n_buffers * (B200_REPLAY_BUFFER_BYTES / 2);
    printf("{\"dongles\": %d, \"iq_samples\": %.0f, \"seconds\": %.3f, \"msamples_per_s\": %.1f, \"realtime_factor\": %.1f, "
           "\"spectrum_messages\": %lld, \"spectrum_bytes\": %lld, \"audio_messages\": %lld, \"audio_bytes\": %lld, "
           "\"audio_samples\": %lld, \"audio_rms\": %.4f, \"kernel_launches\": %llu}\n",
           n, samples, dt, samples / dt / 1e6, samples / dt / ((double) n * FS), (long long) sm, (long long) sb, (long long) am,
           (long long) ab, (long long) as, as ? sqrt(energy / (double) as) : 0.0, (unsigned long long) b200_launch_count());
    for (int i = 0; i < n; ++i) {
        b200_stream_destroy(ds[i].gpu);
        rtl_close(ds[i].dev);
        free(ds[i].capture);
    }
    free(ds);
    free(th);
    return 0;
}
