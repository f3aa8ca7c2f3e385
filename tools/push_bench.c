/* tools/push_bench.c -- the ingest path at the reference's callback granularity: n_streams virtual dongles,
 * every "callback" hands b200_stream_push one librtlsdr buffer (262144 bytes = 131072 cmplx_u8, signal_source.c:31),
 * round robin over the streams, one producer thread (as signal_source.c has).  Sinks only count.
 *
 *   gcc -O2 -o tools/bin/push_bench tools/push_bench.c -Iinclude -Lrtl-ws_b200 -lb200sdr -Wl,-rpath,$PWD/rtl-ws_b200
 *   tools/bin/push_bench [n_streams=256] [buffers_per_stream=64] [batch_samples=204800] [products=0]
 *   products 0: per-frame dB rows + audio (4.1 bytes per sample back over PCIe);
 *            1: what the reference sends its client -- 6-frame averaged payload bytes + audio (0.1 bytes per sample)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200sdr.h"

static int64_t g_frames = 0, g_audio = 0, g_payloads = 0;
static void on_spectra(void* u, int s, int64_t first, int n, const float* db) { (void) u; (void) s; (void) first; (void) db; g_frames += n; }
static void on_audio(void* u, int s, int64_t first, int n, const float* a) { (void) u; (void) s; (void) first; (void) a; g_audio += n; }

static void on_payload(void* u, int s, int64_t first, int k, const uint8_t* p) { (void) u; (void) s; (void) first; (void) k; (void) p; g_payloads += 1; }

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char** argv)
{
    const int n_streams = argc > 1 ? atoi(argv[1]) : 256;
    const int n_buffers = argc > 2 ? atoi(argv[2]) : 64;
    const int64_t batch = argc > 3 ? atoll(argv[3]) : 204800;
    const int products = argc > 4 ? atoi(argv[4]) : 0;
    const int BUF = 262144;
    if (b200_init(0)) { fprintf(stderr, "b200_init: %s\n", b200_last_error()); return 1; }
    b200_stream* st = b200_stream_create(n_streams, batch, 0);
    if (!st) { fprintf(stderr, "b200_stream_create: %s\n", b200_last_error()); return 1; }
    b200_stream_set_sinks(st, products == 0 ? on_spectra : NULL, on_audio, NULL);
    if (products == 1) b200_stream_set_payload_sink(st, 6, on_payload);
    /* 16 distinct pseudo-random buffers, reused: the path does not depend on the data */
    uint8_t* caps = malloc((size_t) 16 * BUF);
    uint32_t x = 12345;
    for (size_t i = 0; i < (size_t) 16 * BUF; ++i) { x = x * 1664525u + 1013904223u; caps[i] = (uint8_t) (x >> 24); }
    for (int pass = 0; pass < 2; ++pass) {                     /* pass 0 warms up */
        const int nb = pass == 0 ? 4 : n_buffers;
        g_frames = g_audio = g_payloads = 0;
        const double t0 = now();
        for (int b = 0; b < nb; ++b)
            for (int s = 0; s < n_streams; ++s)
                if (b200_stream_push(st, s, caps + (size_t) ((b + s) & 15) * BUF, BUF / 2)) {
                    fprintf(stderr, "push: %s\n", b200_last_error());
                    return 1;
                }
        b200_stream_flush(st);
        const double dt = now() - t0;
        if (pass == 1) {
            const double samples = (double) n_streams * nb * (BUF / 2);
            printf("{\"n_streams\": %d, \"buffers_per_stream\": %d, \"batch_samples\": %lld, \"seconds\": %.4f, "
                   "\"msamples_per_s\": %.1f, \"frames_delivered\": %lld, \"payloads_delivered\": %lld, \"audio_delivered\": %lld, \"launches\": %llu}\n",
                   n_streams, nb, (long long) batch, dt, samples / dt / 1e6, (long long) g_frames, (long long) g_payloads, (long long) g_audio,
                   (unsigned long long) b200_launch_count());
        }
    }
    b200_stream_destroy(st);
    free(caps);
    return 0;
}
