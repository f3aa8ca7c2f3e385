// stream.cu -- b200_stream_*: the push-style entry point a signal_source callback calls
// (signal_source.h:7 `void (*)(const cmplx_u8*, int)`, fan-out at signal_source.c:29-35).
//
// Each of n independent dongle streams owns two pinned host slots of `batch_samples`
// (the "pinned host rings" of the north star).  b200_stream_push copies the callback's
// borrowed buffer (librtlsdr recycles it) into the current slot; a full slot is submitted
// without blocking: cudaMemcpyAsync to the stream's device row [history | batch], the fused
// chain kernel for that row, the history carry, and asynchronous copies of the dB rows and
// the audio back into pinned result buffers -- all on one of a small pool of CUDA streams,
// so batches of different dongles overlap.  Dongles that run in step (the usual case: same sample rate, same
// buffer size) fill their batches at the same time, so full slots are submitted in RUNS: when every stream
// of a group of up to 16 consecutive streams has its batch ready, the group goes out as one 2-D copy, one
// kernel launch over the group's rows and 2-D copies back -- seven CUDA calls per group instead of per stream.
// A stream that gets ahead of its group (its second slot fills) is submitted on its own, and poll / flush
// submit whatever is ready.  Results are handed to the caller's sinks from
// b200_stream_push / _poll / _flush on the calling thread (no hidden threads), in submission
// order per stream.  Re-blocking is aligned to stream start like rf_decimator.c:88-115, so
// results do not depend on how the source chunks the samples.
#include <string.h>

#include <vector>

#include "b200_common.cuh"

using namespace b200;

namespace b200 {
int launch_fm_history_carry(uint8_t* iq, int64_t stride, int n_streams, int64_t n_samples, int R, cudaStream_t stream);
int launch_fm_history_reset(uint8_t* iq, int64_t stride, int n_streams, int R, cudaStream_t stream);
int fm_history_samples(int R);
}  // namespace b200

namespace {
constexpr int ST_R = 10;
constexpr int ST_SLOTS = 2;
constexpr int ST_LANES = 4;
constexpr int ST_GROUP = 16;       // streams submitted together when all of them are ready
}  // namespace

struct b200_stream {
    int n_streams;
    int64_t batch;                 // samples per batch (multiple of 5120)
    int gain_db;
    int R;                         // down factor of the FM branch
    int hist;                      // samples
    int64_t row_bytes;             // device row pitch: history + batch
    uint8_t* d_rows;               // [n_streams][ST_SLOTS? no: one row per stream][row_bytes]
    float* d_db;                   // [n_streams][batch]
    float* d_audio;                // [n_streams][batch / (4R)]
    uint8_t* h_in;                 // pinned [n_streams][ST_SLOTS][2 * batch]
    float* h_db;                   // pinned [n_streams][ST_SLOTS][batch]
    float* h_audio;                // pinned [n_streams][ST_SLOTS][batch / (4R)]
    uint8_t* d_payload;            // [n_streams][1024]: K-frame averaged payload bytes of the batch start
    uint8_t* h_payload;            // pinned [n_streams][ST_SLOTS][1024]
    int payload_K;                 // 0 = no payload sink
    b200_payload_sink payload_sink;
    cudaStream_t lanes[ST_LANES];
    struct PerStream {
        int64_t fill;              // samples in the current slot
        int slot;                  // slot being filled
        int64_t batches_submitted;
        int64_t batches_delivered;
        cudaEvent_t done[ST_SLOTS];
        bool in_flight[ST_SLOTS];
        bool ready[ST_SLOTS];      // full, not yet submitted
        bool has_db[ST_SLOTS];     // the submitted slot computed and copied per-frame dB rows
    };
    std::vector<PerStream> st;
    b200_spectrum_sink spectrum_sink;
    b200_audio_sink audio_sink;
    void* user;
};

static void stream_deliver(b200_stream* s, int i, bool block)
{
    b200_stream::PerStream& p = s->st[i];
    while (p.batches_delivered < p.batches_submitted) {
        const int slot = (int) (p.batches_delivered % ST_SLOTS);
        if (!p.in_flight[slot]) break;
        if (block) {
            if (cudaEventSynchronize(p.done[slot]) != cudaSuccess) break;
        } else if (cudaEventQuery(p.done[slot]) != cudaSuccess) {
            break;
        }
        const int64_t b = p.batches_delivered;
        const size_t off = ((size_t) i * ST_SLOTS + slot);
        if (s->payload_sink && s->payload_K > 0)
            s->payload_sink(s->user, i, b * (s->batch / 1024), s->payload_K, s->h_payload + off * 1024);
        // a sink installed after this batch was submitted without one gets nothing for it (there are no rows)
        if (s->spectrum_sink && p.has_db[slot])
            s->spectrum_sink(s->user, i, b * (s->batch / 1024), (int) (s->batch / 1024), s->h_db + off * (size_t) s->batch);
        if (s->audio_sink)
            s->audio_sink(s->user, i, b * (s->batch / (4 * s->R)), (int) (s->batch / (4 * s->R)), s->h_audio + off * (size_t) (s->batch / (4 * s->R)));
        p.in_flight[slot] = false;
        ++p.batches_delivered;
    }
}

// Submit the ready slot `slot` of streams [a, b) as one unit.  All of them run on the lane of their group, so
// the batches of one stream stay ordered (the history carry of batch n precedes the copy of batch n + 1).
static int stream_submit_run(b200_stream* s, int a, int b, int slot)
{
    const int n = b - a;
    const size_t batch = (size_t) s->batch;
    const size_t n_audio = batch / (size_t) (4 * s->R);
    cudaStream_t lane = s->lanes[(a / ST_GROUP) % ST_LANES];
    const size_t off = (size_t) a * ST_SLOTS + (size_t) slot;            // first stream's slot; next stream: + ST_SLOTS
    uint8_t* d_batch = s->d_rows + (size_t) a * (size_t) s->row_bytes + 2 * (size_t) s->hist;
    // per-frame dB rows exist only once a spectrum sink has been installed (b200_stream_set_sinks allocates them)
    const bool want_db = s->spectrum_sink != nullptr && s->d_db != nullptr && s->h_db != nullptr;
    float* d_db = want_db ? s->d_db + (size_t) a * batch : nullptr;
    float* d_audio = s->d_audio + (size_t) a * n_audio;
    uint8_t* d_pay = s->d_payload + (size_t) a * 1024;
    const bool want_payload = s->payload_sink != nullptr && s->payload_K > 0;
    B200_CUDA_TRY(cudaMemcpy2DAsync(d_batch, (size_t) s->row_bytes, s->h_in + off * 2 * batch, ST_SLOTS * 2 * batch, 2 * batch,
                                    (size_t) n, cudaMemcpyHostToDevice, lane));
    // without a per-frame sink the dB rows are not computed at all: the FM kernel and the K-frame average only
    int rc = b200_chain_exec_r(d_batch, s->row_bytes, n, s->batch, s->R, s->gain_db, d_db, d_audio,
                               (int64_t) n_audio, want_payload ? d_pay : nullptr, want_payload ? s->payload_K : 0, lane);
    if (rc) return rc;
    rc = launch_fm_history_carry(d_batch, s->row_bytes, n, s->batch, s->R, lane);
    if (rc) return rc;
    // the per-frame dB rows cross PCIe only if somebody listens (4 bytes per sample; the payload is 1 KB per batch)
    if (want_db)
        B200_CUDA_TRY(cudaMemcpy2DAsync(s->h_db + off * batch, ST_SLOTS * batch * sizeof(float), d_db, batch * sizeof(float),
                                        batch * sizeof(float), (size_t) n, cudaMemcpyDeviceToHost, lane));
    B200_CUDA_TRY(cudaMemcpy2DAsync(s->h_audio + off * n_audio, ST_SLOTS * n_audio * sizeof(float), d_audio,
                                    n_audio * sizeof(float), n_audio * sizeof(float), (size_t) n, cudaMemcpyDeviceToHost, lane));
    if (want_payload)
        B200_CUDA_TRY(cudaMemcpy2DAsync(s->h_payload + off * 1024, ST_SLOTS * 1024, d_pay, 1024, 1024, (size_t) n,
                                        cudaMemcpyDeviceToHost, lane));
    for (int i = a; i < b; ++i) {
        b200_stream::PerStream& p = s->st[i];
        B200_CUDA_TRY(cudaEventRecord(p.done[slot], lane));
        p.ready[slot] = false;
        p.in_flight[slot] = true;
        p.has_db[slot] = want_db;
        ++p.batches_submitted;
    }
    return B200_OK;
}

// the slot stream i would submit next (its oldest full one), or -1
static int stream_ready_slot(const b200_stream* s, int i)
{
    const b200_stream::PerStream& p = s->st[i];
    const int slot = (int) (p.batches_submitted % ST_SLOTS);
    return p.ready[slot] ? slot : -1;
}

// submit the whole group of stream i if every member is ready with the same slot; returns 1 if it went out
static int stream_try_group(b200_stream* s, int i)
{
    const int a = (i / ST_GROUP) * ST_GROUP;
    const int b = a + ST_GROUP < s->n_streams ? a + ST_GROUP : s->n_streams;
    const int slot = stream_ready_slot(s, a);
    if (slot < 0) return 0;
    for (int k = a + 1; k < b; ++k)
        if (stream_ready_slot(s, k) != slot) return 0;
    const int rc = stream_submit_run(s, a, b, slot);
    return rc ? rc : 1;
}

// submit everything that is ready, in maximal runs of consecutive streams of one group with the same slot
static int stream_submit_ready(b200_stream* s)
{
    int i = 0;
    while (i < s->n_streams) {
        const int slot = stream_ready_slot(s, i);
        if (slot < 0) {
            ++i;
            continue;
        }
        const int group_end = (i / ST_GROUP + 1) * ST_GROUP < s->n_streams ? (i / ST_GROUP + 1) * ST_GROUP : s->n_streams;
        int j = i + 1;
        while (j < group_end && stream_ready_slot(s, j) == slot) ++j;
        const int rc = stream_submit_run(s, i, j, slot);
        if (rc) return rc;
        // the same stream may hold a second full slot: look at it again
    }
    return B200_OK;
}

extern "C" {

b200_stream* b200_stream_create(int n_streams, int64_t batch_samples, int gain_db)
{
    return b200_stream_create_r(n_streams, batch_samples, gain_db, ST_R);
}

b200_stream* b200_stream_create_r(int n_streams, int64_t batch_samples, int gain_db, int R)
{
    const int64_t tile = b200_chain_tile_samples(R);
    if (tile < 0) {
        set_error("stream: down factor %d outside [1, 256]", R);
        return nullptr;
    }
    if (n_streams < 1 || batch_samples < tile || batch_samples % tile != 0 || batch_samples < fm_history_samples(R)) {
        // (the FM state is the last 32 R input samples: at R > 32 one tile is shorter than that)
        set_error("stream: batch_samples must be a positive multiple of %lld and at least %d", (long long) tile,
                  fm_history_samples(R));
        return nullptr;
    }
    b200_stream* s = new b200_stream();
    s->n_streams = n_streams;
    s->batch = batch_samples;
    s->gain_db = gain_db;
    s->R = R;
    s->hist = fm_history_samples(R);
    s->row_bytes = 2 * ((int64_t) s->hist + batch_samples);
    s->d_rows = nullptr;
    s->d_db = nullptr;
    s->d_audio = nullptr;
    s->h_in = nullptr;
    s->h_db = nullptr;
    s->h_audio = nullptr;
    s->d_payload = nullptr;
    s->h_payload = nullptr;
    s->payload_K = 0;
    s->payload_sink = nullptr;
    s->spectrum_sink = nullptr;
    s->audio_sink = nullptr;
    s->user = nullptr;
    for (int i = 0; i < ST_LANES; ++i) s->lanes[i] = nullptr;
    const size_t ns = (size_t) n_streams;
    bool ok = cudaMalloc(&s->d_rows, ns * (size_t) s->row_bytes) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_audio, sizeof(float) * ns * (size_t) (batch_samples / (4 * R))) == cudaSuccess;
    ok = ok && cudaHostAlloc(&s->h_in, ns * ST_SLOTS * 2 * (size_t) batch_samples, cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc(&s->h_audio, sizeof(float) * ns * ST_SLOTS * (size_t) (batch_samples / (4 * R)), cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_payload, ns * 1024) == cudaSuccess;
    ok = ok && cudaHostAlloc(&s->h_payload, ns * ST_SLOTS * 1024, cudaHostAllocDefault) == cudaSuccess;
    for (int i = 0; ok && i < ST_LANES; ++i) ok = cudaStreamCreateWithFlags(&s->lanes[i], cudaStreamNonBlocking) == cudaSuccess;
    s->st.resize(ns);
    for (size_t i = 0; i < ns; ++i) {
        s->st[i].fill = 0;
        s->st[i].slot = 0;
        s->st[i].batches_submitted = 0;
        s->st[i].batches_delivered = 0;
        for (int k = 0; k < ST_SLOTS; ++k) {
            s->st[i].done[k] = nullptr;
            s->st[i].in_flight[k] = false;
            s->st[i].ready[k] = false;
            s->st[i].has_db[k] = false;
            if (ok) ok = cudaEventCreateWithFlags(&s->st[i].done[k], cudaEventDisableTiming) == cudaSuccess;
        }
    }
    if (ok)
        ok = launch_fm_history_reset(s->d_rows + 2 * (size_t) s->hist, s->row_bytes, n_streams, R, s->lanes[0]) == B200_OK &&
             cudaStreamSynchronize(s->lanes[0]) == cudaSuccess;
    if (!ok) {
        set_error("stream: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        b200_stream_destroy(s);
        return nullptr;
    }
    return s;
}

void b200_stream_destroy(b200_stream* s)
{
    if (s == nullptr) return;
    for (int i = 0; i < ST_LANES; ++i)
        if (s->lanes[i]) {
            cudaStreamSynchronize(s->lanes[i]);
            cudaStreamDestroy(s->lanes[i]);
        }
    for (auto& p : s->st)
        for (int k = 0; k < ST_SLOTS; ++k)
            if (p.done[k]) cudaEventDestroy(p.done[k]);
    if (s->d_rows) cudaFree(s->d_rows);
    if (s->d_db) cudaFree(s->d_db);
    if (s->d_audio) cudaFree(s->d_audio);
    if (s->h_in) cudaFreeHost(s->h_in);
    if (s->h_db) cudaFreeHost(s->h_db);
    if (s->h_audio) cudaFreeHost(s->h_audio);
    if (s->d_payload) cudaFree(s->d_payload);
    if (s->h_payload) cudaFreeHost(s->h_payload);
    delete s;
}

void b200_stream_set_sinks(b200_stream* s, b200_spectrum_sink spectrum_sink, b200_audio_sink audio_sink, void* user)
{
    if (s == nullptr) return;
    // The per-frame dB rows (4 bytes per sample on the device, twice that pinned on the host: 420 MB for 256 streams
    // of 204800 samples) exist only for callers that listen to them: allocated when the first spectrum sink arrives.
    if (spectrum_sink != nullptr && s->d_db == nullptr) {
        const size_t ns = (size_t) s->n_streams;
        bool ok = cudaMalloc(&s->d_db, sizeof(float) * ns * (size_t) s->batch) == cudaSuccess;
        ok = ok && cudaHostAlloc(&s->h_db, sizeof(float) * ns * ST_SLOTS * (size_t) s->batch, cudaHostAllocDefault) == cudaSuccess;
        if (!ok) {
            set_error("stream: dB row buffers: %s", cudaGetErrorString(cudaGetLastError()));
            if (s->d_db) cudaFree(s->d_db);
            s->d_db = nullptr;
            s->h_db = nullptr;
            spectrum_sink = nullptr;      // reported through b200_last_error(); audio and payload keep working
        }
    }
    s->spectrum_sink = spectrum_sink;
    s->audio_sink = audio_sink;
    s->user = user;
}

int b200_stream_set_payload_sink(b200_stream* s, int K, b200_payload_sink sink)
{
    if (s == nullptr || K < 0 || (int64_t) K * 1024 > s->batch) {
        set_error("stream payload sink: K = %d frames do not fit a batch", K);
        return B200_ERR_ARG;
    }
    s->payload_K = sink ? K : 0;
    s->payload_sink = sink;
    return B200_OK;
}

int b200_stream_push(b200_stream* s, int stream, const uint8_t* samples, int len)
{
    if (s == nullptr || stream < 0 || stream >= s->n_streams || len < 0 || (len > 0 && samples == nullptr)) {
        set_error("stream push: bad arguments");
        return B200_ERR_ARG;
    }
    b200_stream::PerStream& p = s->st[stream];
    const uint8_t* src = samples;
    int64_t remaining = len;
    while (remaining > 0) {
        // the slot about to be written must be free: a batch still waiting for its group goes out on its own,
        // a batch in flight is waited for and delivered (two batches per stream at most)
        if (p.fill == 0 && p.ready[p.slot]) {
            const int rc = stream_submit_run(s, stream, stream + 1, p.slot);
            if (rc) return rc;
        }
        if (p.fill == 0 && p.in_flight[p.slot]) stream_deliver(s, stream, true);
        const int64_t room = s->batch - p.fill;
        const int64_t n = remaining < room ? remaining : room;
        uint8_t* dst = s->h_in + ((size_t) stream * ST_SLOTS + p.slot) * 2 * (size_t) s->batch + 2 * (size_t) p.fill;
        memcpy(dst, src, 2 * (size_t) n);
        src += 2 * n;
        p.fill += n;
        remaining -= n;
        if (p.fill == s->batch) {
            p.ready[p.slot] = true;
            p.slot = (p.slot + 1) % ST_SLOTS;
            p.fill = 0;
            const int rc = stream_try_group(s, stream);
            if (rc < 0) return rc;
        }
    }
    stream_deliver(s, stream, false);
    return B200_OK;
}

int b200_stream_poll(b200_stream* s)
{
    if (s == nullptr) return B200_ERR_ARG;
    if (int rc = stream_submit_ready(s)) return rc;         // batches waiting for a slower member of their group
    for (int i = 0; i < s->n_streams; ++i) stream_deliver(s, i, false);
    return B200_OK;
}

int b200_stream_flush(b200_stream* s)
{
    if (s == nullptr) return B200_ERR_ARG;
    if (int rc = stream_submit_ready(s)) return rc;
    for (int i = 0; i < s->n_streams; ++i) stream_deliver(s, i, true);
    return B200_OK;
}

int64_t b200_stream_pending_samples(const b200_stream* s, int stream)
{
    if (s == nullptr || stream < 0 || stream >= s->n_streams) return B200_ERR_ARG;
    return s->st[stream].fill;
}

}  // extern "C"
