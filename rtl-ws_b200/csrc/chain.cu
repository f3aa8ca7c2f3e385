// chain.cu -- the full spectrum + FM chain entry points (b200sdr.h): device-resident
// b200_chain_exec and the host-buffer session that wraps it with pipelined PCIe copies.
#include <mutex>
#include <map>

#include "b200_common.cuh"
#include "fm_kernels.cuh"
#include "spectrum_kernels.cuh"

namespace b200 {
int launch_chain_fused(const uint8_t* d_iq, int64_t stride, int n_streams, int64_t n_samples, float db_offset,
                       const float2* twiddle, float* d_db, float* d_audio, int64_t audio_stride, cudaStream_t stream);
int launch_fm_history_carry(uint8_t* iq, int64_t stride, int n_streams, int64_t n_samples, int R, cudaStream_t stream);
int launch_fm_history_reset(uint8_t* iq, int64_t stride, int n_streams, int R, cudaStream_t stream);
int fm_history_samples(int R);
}  // namespace b200

using namespace b200;

namespace {

constexpr int CHAIN_N = 1024;      // cbb_main.c:17 FFT_POINTS
constexpr int CHAIN_R = 10;        // cbb_main.c:80 at rtl_sensor.c:12's 2.048 MS/s and main.c:23's 192 kHz
constexpr int CHAIN_TILE = 5120;   // lcm(1024, 4 * R): 5 frames = 128 audio samples

int64_t gcd64(int64_t a, int64_t b)
{
    while (b) {
        const int64_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

std::mutex g_plan_mutex;
// (device, K, gain_db) -> plan; plans are tiny (a twiddle table) and live for the process
std::map<std::tuple<int, int, int>, b200_spectrum_plan*> g_plans;

b200_spectrum_plan* cached_plan(int K, int gain_db)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    // only the integer quotient gain_db / 10 reaches the arithmetic (cbb_main.c:112): one plan per decade,
    // so a client sweeping `spectrumgain` does not grow the cache
    gain_db = (gain_db / 10) * 10;
    auto key = std::make_tuple(dev, K, gain_db);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) return it->second;
    b200_spectrum_plan* pl = b200_spectrum_plan_create(CHAIN_N, CHAIN_N, K, 0, B200_WINDOW_RECT, gain_db);
    if (pl != nullptr) g_plans[key] = pl;
    return pl;
}

}  // namespace

// layout of b200_spectrum_plan is private to capi.cu; the fused launcher needs two fields
extern "C" const void* b200_spectrum_plan_twiddle_(const b200_spectrum_plan* plan);
extern "C" float b200_spectrum_plan_db_offset_(const b200_spectrum_plan* plan);

extern "C" {

int64_t b200_chain_tile_samples(int R)
{
    if (R < 1 || R > 256) return B200_ERR_ARG;
    const int64_t a = CHAIN_N, b = 4 * (int64_t) R;
    return a / gcd64(a, b) * b;                 // lcm(1024, 4R): whole frames and whole audio samples (a multiple of 8)
}

int b200_chain_exec(const uint8_t* d_iq, int64_t stream_stride_bytes, int n_streams, int64_t n_samples, int gain_db,
                    float* d_db, float* d_audio, int64_t audio_stride, uint8_t* d_avg_u8, int K_avg,
                    void* cuda_stream)
{
    return b200_chain_exec_r(d_iq, stream_stride_bytes, n_streams, n_samples, CHAIN_R, gain_db, d_db, d_audio,
                             audio_stride, d_avg_u8, K_avg, cuda_stream);
}

int b200_chain_exec_r(const uint8_t* d_iq, int64_t stream_stride_bytes, int n_streams, int64_t n_samples, int R,
                      int gain_db, float* d_db, float* d_audio, int64_t audio_stride, uint8_t* d_avg_u8, int K_avg,
                      void* cuda_stream)
{
    const int64_t tile = b200_chain_tile_samples(R);
    if (tile < 0) {
        set_error("chain exec: down factor %d outside [1, 256]", R);
        return B200_ERR_ARG;
    }
    if (d_iq == nullptr || n_streams < 0 || n_samples < 0 || n_samples % tile != 0) {
        set_error("chain exec: n_samples %lld must be a non-negative multiple of %lld", (long long) n_samples, (long long) tile);
        return B200_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_iq) & 15) != 0 || (stream_stride_bytes & 15) != 0) {
        set_error("chain exec: IQ pointer and stream stride must be 16-byte aligned");
        return B200_ERR_ALIGN;
    }
    if (n_streams == 0 || n_samples == 0) return B200_OK;
    if (d_audio != nullptr && n_samples < fm_history_samples(R)) {
        // the carried state is the last 32 R input samples: a batch must hold at least that many
        set_error("chain exec: a batch of %lld samples is shorter than the FM history (%d samples at R = %d)",
                  (long long) n_samples, fm_history_samples(R), R);
        return B200_ERR_ARG;
    }
    if (d_avg_u8 != nullptr && (K_avg < 1 || (int64_t) K_avg * CHAIN_N > n_samples)) {     // before anything is launched
        set_error("chain exec: K_avg = %d does not fit the batch", K_avg);
        return B200_ERR_ARG;
    }
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    b200_spectrum_plan* pl = cached_plan(1, gain_db);
    if (pl == nullptr) return B200_ERR_CUDA;

    int rc;
    if (d_db != nullptr && d_audio != nullptr && R == CHAIN_R) {      // one pass over the IQ (chain_fused.cu)
        rc = launch_chain_fused(d_iq, stream_stride_bytes, n_streams, n_samples, b200_spectrum_plan_db_offset_(pl),
                                reinterpret_cast<const float2*>(b200_spectrum_plan_twiddle_(pl)), d_db, d_audio,
                                audio_stride, stream);
        if (rc) return rc;
    } else {
        if (d_db != nullptr) {
            rc = b200_spectrum_exec(pl, d_iq, stream_stride_bytes, n_streams, n_samples / CHAIN_N, d_db, nullptr,
                                    nullptr, cuda_stream);
            if (rc) return rc;
        }
        if (d_audio != nullptr) {
            rc = b200_fm_exec(d_iq, stream_stride_bytes, n_streams, n_samples, R, d_audio, audio_stride, nullptr,
                              0, cuda_stream);
            if (rc) return rc;
        }
    }
    if (d_avg_u8 != nullptr) {
        b200_spectrum_plan* avg = cached_plan(K_avg, gain_db);
        if (avg == nullptr) return B200_ERR_CUDA;
        rc = b200_spectrum_exec(avg, d_iq, stream_stride_bytes, n_streams, 1, nullptr, nullptr, d_avg_u8, cuda_stream);
        if (rc) return rc;
    }
    return B200_OK;
}

}  // extern "C"

// ---- host-buffer session -----------------------------------------------------------------

struct b200_session {
    int n_streams;
    int64_t max_samples;
    int R;                     // down factor of the FM branch
    int64_t tile;              // lcm(1024, 4R)
    int hist_samples;
    int64_t stride_bytes;      // per-stream row pitch in the device ring: history + batch
    uint8_t* d_ring;           // [n_streams][stride_bytes]
    float* d_db;               // [n_streams][max_samples]            (1024 bins per 1024 samples)
    float* d_audio;            // [n_streams][max_samples / (4R)]
    uint8_t* d_avg;            // [n_streams][1024] payload rows, allocated on first use (b200_session_products)
    static constexpr int LANES = 4;
    cudaStream_t streams[LANES];
    int device;
};

namespace b200 {

// Submit one batch of every stream of the session without waiting: rows of the host arrays may be strided
// (pitches in bytes for h_iq, in floats for h_db / h_audio), which is how a multi-GPU host hands each device
// its streams s mod G of one global array.  d_avg_u8 (device, nullable): [n_streams][1024] payload rows.
int session_chain_submit(b200_session* s, const uint8_t* h_iq, int64_t iq_pitch_bytes, int64_t n_samples, int gain_db,
                         float* h_db, int64_t db_pitch, float* h_audio, int64_t audio_pitch, uint8_t* d_avg_u8, int K_avg)
{
    if (s == nullptr || h_iq == nullptr || n_samples < 0 || n_samples > s->max_samples || n_samples % s->tile != 0) {
        set_error("session chain: n_samples %lld must be a multiple of %lld and at most %lld", (long long) n_samples,
                  s ? (long long) s->tile : 0ll, s ? (long long) s->max_samples : 0ll);
        return B200_ERR_ARG;
    }
    if (n_samples == 0) return B200_OK;
    const int lanes = b200_session::LANES;
    // stream groups small enough to pipeline H2D / kernels / D2H across the lanes
    int group = (s->n_streams + 4 * lanes - 1) / (4 * lanes);
    if (group < 1) group = 1;
    const int64_t n_audio = n_samples / (4 * s->R);
    int lane = 0;
    for (int s0 = 0; s0 < s->n_streams; s0 += group, lane = (lane + 1) % lanes) {
        const int ns = (s->n_streams - s0) < group ? (s->n_streams - s0) : group;
        cudaStream_t st = s->streams[lane];
        uint8_t* d_batch = s->d_ring + (int64_t) s0 * s->stride_bytes + 2 * (int64_t) s->hist_samples;
        float* d_db = s->d_db + (size_t) s0 * (size_t) n_samples;
        float* d_audio = s->d_audio + (size_t) s0 * (size_t) n_audio;
        B200_CUDA_TRY(cudaMemcpy2DAsync(d_batch, (size_t) s->stride_bytes, h_iq + (size_t) s0 * (size_t) iq_pitch_bytes,
                                        (size_t) iq_pitch_bytes, (size_t) (2 * n_samples), (size_t) ns,
                                        cudaMemcpyHostToDevice, st));
        const int rc = b200_chain_exec_r(d_batch, s->stride_bytes, ns, n_samples, s->R, gain_db, h_db ? d_db : nullptr,
                                         h_audio ? d_audio : nullptr, n_audio, d_avg_u8 ? d_avg_u8 + (size_t) s0 * 1024 : nullptr,
                                         K_avg, st);
        if (rc) return rc;
        if (h_db)
            B200_CUDA_TRY(cudaMemcpy2DAsync(h_db + (size_t) s0 * (size_t) db_pitch, sizeof(float) * (size_t) db_pitch, d_db,
                                            sizeof(float) * (size_t) n_samples, sizeof(float) * (size_t) n_samples,
                                            (size_t) ns, cudaMemcpyDeviceToHost, st));
        if (h_audio)
            B200_CUDA_TRY(cudaMemcpy2DAsync(h_audio + (size_t) s0 * (size_t) audio_pitch, sizeof(float) * (size_t) audio_pitch,
                                            d_audio, sizeof(float) * (size_t) n_audio, sizeof(float) * (size_t) n_audio,
                                            (size_t) ns, cudaMemcpyDeviceToHost, st));
        const int rc2 = launch_fm_history_carry(d_batch, s->stride_bytes, ns, n_samples, s->R, st);
        if (rc2) return rc2;
    }
    return B200_OK;
}

int session_wait(b200_session* s)
{
    for (int i = 0; i < b200_session::LANES; ++i) B200_CUDA_TRY(cudaStreamSynchronize(s->streams[i]));
    return B200_OK;
}

}  // namespace b200

extern "C" {

b200_session* b200_session_create(int n_streams, int64_t max_samples_per_batch)
{
    return b200_session_create_r(n_streams, max_samples_per_batch, CHAIN_R);
}

b200_session* b200_session_create_r(int n_streams, int64_t max_samples_per_batch, int R)
{
    const int64_t tile = b200_chain_tile_samples(R);
    if (tile < 0) {
        set_error("session: down factor %d outside [1, 256]", R);
        return nullptr;
    }
    if (n_streams < 1 || max_samples_per_batch < tile || max_samples_per_batch % tile != 0 ||
        max_samples_per_batch < fm_history_samples(R)) {
        set_error("session: max_samples_per_batch must be a positive multiple of %lld and at least %d", (long long) tile,
                  fm_history_samples(R));
        return nullptr;
    }
    b200_session* s = new b200_session();
    memset(s, 0, sizeof(*s));
    s->n_streams = n_streams;
    s->max_samples = max_samples_per_batch;
    s->R = R;
    s->tile = tile;
    s->hist_samples = fm_history_samples(R);
    s->stride_bytes = 2 * ((int64_t) s->hist_samples + max_samples_per_batch);
    bool ok = cudaGetDevice(&s->device) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_ring, (size_t) n_streams * (size_t) s->stride_bytes) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_db, sizeof(float) * (size_t) n_streams * (size_t) max_samples_per_batch) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_audio, sizeof(float) * (size_t) n_streams * (size_t) (max_samples_per_batch / (4 * R))) ==
                   cudaSuccess;
    for (int i = 0; ok && i < b200_session::LANES; ++i)
        ok = cudaStreamCreateWithFlags(&s->streams[i], cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        set_error("session: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        b200_session_destroy(s);
        return nullptr;
    }
    b200_session_reset(s);
    return s;
}

void b200_session_destroy(b200_session* s)
{
    if (s == nullptr) return;
    for (int i = 0; i < b200_session::LANES; ++i)
        if (s->streams[i]) cudaStreamDestroy(s->streams[i]);
    if (s->d_ring) cudaFree(s->d_ring);
    if (s->d_db) cudaFree(s->d_db);
    if (s->d_audio) cudaFree(s->d_audio);
    if (s->d_avg) cudaFree(s->d_avg);
    delete s;
}

void b200_session_reset(b200_session* s)
{
    if (s == nullptr) return;
    launch_fm_history_reset(s->d_ring + 2 * (int64_t) s->hist_samples, s->stride_bytes, s->n_streams, s->R,
                            s->streams[0]);
    cudaStreamSynchronize(s->streams[0]);
}

int b200_session_chain(b200_session* s, const uint8_t* h_iq, int64_t n_samples, int gain_db, float* h_db,
                       float* h_audio)
{
    if (s == nullptr) {
        set_error("session chain: null session");
        return B200_ERR_ARG;
    }
    const int rc = b200::session_chain_submit(s, h_iq, 2 * n_samples, n_samples, gain_db, h_db, n_samples, h_audio,
                                              n_samples / (4 * s->R), nullptr, 0);
    if (rc) return rc;
    return b200::session_wait(s);
}

int b200_session_products(b200_session* s, const uint8_t* h_iq, int64_t n_samples, int gain_db, int K_avg,
                          float* h_audio, uint8_t* h_avg_u8)
{
    if (s == nullptr || h_avg_u8 == nullptr || K_avg < 1) {
        set_error("session products: bad arguments");
        return B200_ERR_ARG;
    }
    if (s->d_avg == nullptr) B200_CUDA_TRY(cudaMalloc((void**) &s->d_avg, (size_t) s->n_streams * 1024));
    int rc = b200::session_chain_submit(s, h_iq, 2 * n_samples, n_samples, gain_db, nullptr, 0, h_audio,
                                        n_samples / (4 * s->R), s->d_avg, K_avg);
    if (rc) return rc;
    rc = b200::session_wait(s);
    if (rc) return rc;
    B200_CUDA_TRY(cudaMemcpy(h_avg_u8, s->d_avg, (size_t) s->n_streams * 1024, cudaMemcpyDeviceToHost));
    return B200_OK;
}

}  // extern "C"
