/*
 * multi_gpu_gather.c -- a plain-C host that shards dongle streams over every GPU of the box and gathers
 * what the reference would send its UI clients (the 6-frame averaged payload bytes of cbb_main.c:121-130,
 * main.c:80-84) on GPU 0, all through the C ABI of libb200sdr.so: no CUDA headers, no NCCL headers.
 *
 *   gcc -O2 -Iinclude examples/multi_gpu_gather.c -Lrtl-ws_b200 -lb200sdr -Wl,-rpath,rtl-ws_b200 -o multi_gpu_gather
 *   ./multi_gpu_gather [n_gpus] [n_streams]
 *
 * Stream s lives on GPU s mod G (b200_shard_stream); each GPU runs b200_session_chain on its own streams
 * (host buffers in, PCIe copies inside) and leaves its payload rows on the device; one call,
 * b200_comm_gather_rows_all, moves them to GPU 0 in global stream order over NCCL.  The program checks
 * the gathered rows against the same rows computed by a single GPU.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200sdr.h"

#define BATCH 20480        /* 4 tiles of 5120 samples */
#define MAX_GPUS 8

/* device memory without CUDA headers: the session owns the IQ; the payload rows come from a tiny helper
 * plan executed on host-visible pinned memory is not possible, so the example keeps to the C ABI and uses
 * b200_host_alloc'ed staging plus the stream API's payload sink for the single-GPU reference rows */
static uint8_t g_ref_rows[1024][1024];
static void payload_sink(void* user, int stream, int64_t first_frame, int k, const uint8_t* payload)
{
    (void) user;
    (void) k;
    if (first_frame == 0 && stream < 1024) memcpy(g_ref_rows[stream], payload, 1024);
}

int main(int argc, char** argv)
{
    const int n_gpus = argc > 1 ? atoi(argv[1]) : 2;
    const int n_streams = argc > 2 ? atoi(argv[2]) : 16;
    int s, g;
    if (n_gpus < 1 || n_gpus > MAX_GPUS || n_streams < 1 || n_streams > 1024) {
        fprintf(stderr, "usage: %s [n_gpus <= %d] [n_streams <= 1024]\n", argv[0], MAX_GPUS);
        return 2;
    }
    printf("sharding %d streams over %d GPUs:", n_streams, n_gpus);
    for (g = 0; g < n_gpus; ++g) printf(" gpu%d=%d", g, b200_shard_count(n_streams, n_gpus, g));
    printf("\n");

    /* synthetic IQ, one batch per stream (a tone whose frequency depends on the stream) */
    uint8_t* iq = (uint8_t*) malloc((size_t) n_streams * BATCH * 2);
    for (s = 0; s < n_streams; ++s)
        for (int i = 0; i < BATCH; ++i) {
            const int ph = (i * (s + 1) * 37) & 1023;
            iq[((size_t) s * BATCH + i) * 2 + 0] = (uint8_t) (128 + (ph < 512 ? 60 : -60));
            iq[((size_t) s * BATCH + i) * 2 + 1] = (uint8_t) (128 + (((ph + 256) & 1023) < 512 ? 60 : -60));
        }

    /* single-GPU reference rows through the push API's payload sink */
    if (b200_init(0) != B200_OK) {
        fprintf(stderr, "no CUDA device: %s\n", b200_last_error());
        return 1;
    }
    b200_stream* st = b200_stream_create(n_streams, BATCH, 0);
    if (!st) {
        fprintf(stderr, "%s\n", b200_last_error());
        return 1;
    }
    b200_stream_set_payload_sink(st, 6, payload_sink);
    for (s = 0; s < n_streams; ++s) b200_stream_push(st, s, iq + (size_t) s * BATCH * 2, BATCH);
    b200_stream_flush(st);
    b200_stream_destroy(st);

    /* the sharded run: one multi-GPU job object does the bookkeeping of rows and buffers */
    b200_multi* job = b200_multi_create(n_gpus, n_streams, BATCH, 0, 6);
    if (!job) {
        fprintf(stderr, "b200_multi_create: %s\n", b200_last_error());
        return 1;
    }
    uint8_t* rows = (uint8_t*) malloc((size_t) n_streams * 1024);
    if (b200_multi_chain(job, iq, BATCH, NULL, NULL, rows) != B200_OK) {
        fprintf(stderr, "b200_multi_chain: %s\n", b200_last_error());
        return 1;
    }
    int bad = 0;
    for (s = 0; s < n_streams; ++s)
        if (memcmp(rows + (size_t) s * 1024, g_ref_rows[s], 1024) != 0) ++bad;
    printf("gathered %d payload rows on gpu0 over NCCL %d: %s\n", n_streams, b200_comm_nccl_version(),
           bad ? "MISMATCH" : "all equal to the single-GPU rows");
    b200_multi_destroy(job);
    free(rows);
    free(iq);
    return bad ? 1 : 0;
}
