/*
 * oracle/ref_bench.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Times the reference's own CPU implementation of the hot path (the unmodified
 * spectrum.c / rf_decimator.c / resample.c / audio_main.c inside libref_rtlws.so) on the
 * host cores.  The reference is single-threaded per stream and audio_main.c keeps its
 * demodulator state in statics, so parallelism is one PROCESS per worker, each taking
 * whole streams (streams are independent).
 *
 *   ref_bench <iq_file> <n_streams> <samples_per_stream> <n_workers> <mode> [unique]
 *       mode: chain | spectrum | fm | products
 *       unique: the file holds this many distinct captures (default n_streams); stream s reads
 *               capture s mod unique, so a long run does not need a multi-gigabyte file
 *
 * Per stream, in the reference's own granularity:
 *   - the stream arrives in 131072-sample source buffers (signal_source.c:29-31);
 *   - spectrum: EVERY 1024-sample frame goes through spectrum_add_cmplx_u8 into a zeroed
 *     double[1024] (cbb_main.c:50-54 with one frame per estimate) and the 10*log10
 *     epilogue of cbb_main.c:125 is applied per bin and stored as float (no truncation);
 *   - fm: rf_decimator_decimate_cmplx_u8 -> cic_decimate -> audio_fm_demodulator;
 *   - products: what the reference actually hands its clients -- fm as above plus, once per
 *     PRODUCT_PERIOD samples (250 ms at 2.048 MS/s: cbb_main.c:16 SPECTRUM_EST_MS), the
 *     FFT_AVERAGE = 6 frame average and its payload bytes (cbb_main.c:40-70, 106-135).
 * The FFT inside spectrum.c is the stand-in of oracle/fftw3_shim (FFTW3 is not installed).
 *
 * Prints one JSON object on stdout.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>

#include "spectrum.h"

extern int ref_fm_open(double sample_rate, int down_factor);
extern void ref_fm_set_outputs(int32_t* dec, int64_t dec_cap, float* audio, int64_t audio_cap);
extern int ref_fm_push(const uint8_t* iq, int64_t n_samples, int chunk);
extern int64_t ref_fm_n_audio(void);

#define FFT_POINTS 1024
#define SOURCE_BUF_SAMPLES 131072
#define PRODUCT_PERIOD 512000          /* 250 ms of IQ at rtl_sensor.c:12's 2.048 MS/s */
#define FFT_AVERAGE 6                  /* cbb_main.c:18 */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

struct shared_ctl
{
    volatile int ready;
    volatile int go;
    double checksum[1024];
};

static void run_stream(const uint8_t* iq, int64_t n, int do_spec, int do_fm,
                       struct spectrum* spect, float* db_out, float* audio_out, int64_t audio_cap,
                       double* checksum)
{
    int64_t pos = 0;
    int64_t next_product = 0;
    double ps[FFT_POINTS];
    if (do_fm)
        ref_fm_set_outputs(NULL, 0, audio_out, audio_cap);
    while (pos < n)
    {
        int len = (int) ((n - pos) < SOURCE_BUF_SAMPLES ? (n - pos) : SOURCE_BUF_SAMPLES);
        if (do_fm)
            ref_fm_push(iq + 2 * pos, len, len);
        if (do_spec == 2)
        {
            /* cbb_main.c:48-59 at the first frames of every period, then cbb_main.c:121-130 */
            for (; next_product + FFT_AVERAGE * FFT_POINTS <= pos + len; next_product += PRODUCT_PERIOD)
            {
                const int64_t p0 = next_product;
                int f, i;
                unsigned char* payload = (unsigned char*) db_out;
                memset(ps, 0, sizeof(ps));
                for (f = 0; f < FFT_AVERAGE; f++)
                    spectrum_add_cmplx_u8(spect, (const cmplx_u8*) (iq + 2 * (p0 + (int64_t) f * FFT_POINTS)), ps, FFT_POINTS);
                for (i = 0; i < FFT_POINTS; i++)
                {
                    int m = (int) (10 * log10(fabs(ps[i] / FFT_AVERAGE)));
                    payload[i] = (unsigned char) (m < 0 ? 0 : (m > 255 ? 255 : m));
                }
                *checksum += payload[17];
            }
        }
        else if (do_spec)
        {
            int frames = len / FFT_POINTS;
            int f, i;
            for (f = 0; f < frames; f++)
            {
                float* row = db_out + (size_t) ((pos / FFT_POINTS + f) % 64) * FFT_POINTS;
                memset(ps, 0, sizeof(ps));
                spectrum_add_cmplx_u8(spect, (const cmplx_u8*) (iq + 2 * (pos + (int64_t) f * FFT_POINTS)), ps, FFT_POINTS);
                for (i = 0; i < FFT_POINTS; i++)
                    row[i] = (float) (10 * log10(fabs(ps[i])));
                *checksum += row[17];
            }
        }
        pos += len;
    }
    if (do_fm && ref_fm_n_audio() > 0)
        *checksum += audio_out[ref_fm_n_audio() - 1];
}

int main(int argc, char** argv)
{
    const char* path;
    int n_streams, n_workers, w, do_spec, do_fm, unique;
    int64_t per_stream;
    int fd;
    struct stat st;
    const uint8_t* iq;
    struct shared_ctl* ctl;
    pid_t* pids;
    double t0, t1, checksum = 0;

    if (argc < 6)
    {
        fprintf(stderr, "usage: %s <iq_file> <n_streams> <samples_per_stream> <n_workers> <chain|spectrum|fm|products> [unique]\n", argv[0]);
        return 2;
    }
    path = argv[1];
    n_streams = atoi(argv[2]);
    per_stream = atoll(argv[3]);
    n_workers = atoi(argv[4]);
    do_spec = strcmp(argv[5], "fm") != 0;
    if (strcmp(argv[5], "products") == 0) do_spec = 2;
    do_fm = strcmp(argv[5], "spectrum") != 0;
    unique = argc > 6 ? atoi(argv[6]) : n_streams;
    if (unique < 1 || unique > n_streams) unique = n_streams;
    if (n_workers < 1) n_workers = 1;
    if (n_workers > 1024) n_workers = 1024;
    if (n_workers > n_streams) n_workers = n_streams;

    fd = open(path, O_RDONLY);
    if (fd < 0 || fstat(fd, &st) != 0 || st.st_size < (off_t) (2 * per_stream * unique))
    {
        fprintf(stderr, "ref_bench: cannot use %s\n", path);
        return 2;
    }
    iq = (const uint8_t*) mmap(NULL, (size_t) st.st_size, PROT_READ, MAP_SHARED | MAP_POPULATE, fd, 0);
    ctl = (struct shared_ctl*) mmap(NULL, sizeof(*ctl), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (iq == MAP_FAILED || ctl == MAP_FAILED)
        return 2;
    memset((void*) ctl, 0, sizeof(*ctl));
    pids = (pid_t*) calloc((size_t) n_workers, sizeof(pid_t));

    for (w = 0; w < n_workers; w++)
    {
        pids[w] = fork();
        if (pids[w] == 0)
        {
            struct spectrum* spect = spectrum_alloc(FFT_POINTS);
            float* db_out = (float*) malloc(sizeof(float) * 64 * FFT_POINTS);
            int64_t audio_cap = per_stream / 40 + 8192;
            float* audio_out = (float*) malloc(sizeof(float) * (size_t) audio_cap);
            double cs = 0;
            int s;
            volatile uint8_t sink = 0;
            int64_t k;
            if (do_fm)
                ref_fm_open(2048000.0, 10);          /* rtl_sensor.c:12, cbb_main.c:80 */
            /* touch this worker's input once so page faults are outside the timed region */
            for (s = w; s < n_streams; s += n_workers)
                for (k = 0; k < 2 * per_stream; k += 4096)
                    sink ^= iq[(size_t) (s % unique) * 2 * per_stream + k];
            __sync_fetch_and_add(&ctl->ready, 1);
            while (!ctl->go)
                usleep(50);
            for (s = w; s < n_streams; s += n_workers)
                run_stream(iq + (size_t) (s % unique) * 2 * per_stream, per_stream, do_spec, do_fm,
                           spect, db_out, audio_out, audio_cap, &cs);
            ctl->checksum[w] = cs;
            _exit(0);
        }
    }
    while (ctl->ready < n_workers)
        usleep(100);
    t0 = now_s();
    ctl->go = 1;
    for (w = 0; w < n_workers; w++)
    {
        int status = 0;
        waitpid(pids[w], &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0)
        {
            fprintf(stderr, "ref_bench: worker %d failed\n", w);
            return 3;
        }
    }
    t1 = now_s();
    for (w = 0; w < n_workers; w++)
        checksum += ctl->checksum[w];

    printf("{\"mode\": \"%s\", \"n_streams\": %d, \"samples_per_stream\": %lld, \"workers\": %d, "
           "\"seconds\": %.6f, \"msamples_per_s\": %.3f, \"checksum\": %.6g, "
           "\"fft\": \"stand-in f64 radix-4 Stockham (FFTW3 absent)\"}\n",
           argv[5], n_streams, (long long) per_stream, n_workers, t1 - t0,
           1e-6 * (double) n_streams * (double) per_stream / (t1 - t0), checksum);
    return 0;
}
