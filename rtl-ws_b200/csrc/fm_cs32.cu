// fm_cs32.cu -- the FM demodulator behind the reference's own callback signature (sm_100a).
//
// audio_fm_demodulator (audio_main.c:74-145, declared audio_main.h:12) is an rf_decimator_callback:
// it receives DECIMATED cmplx_s32 blocks and keeps three pieces of state in function statics
// (audio_main.c:77-79): the previous phase and the two 10-float half-band delay lines.  This file
// is that function's arithmetic with the state made explicit, so that
//   - an unmodified main.c:205 can register a GPU demodulator (libb200audio.so, audio_compat.cpp), and
//   - a caller that already holds decimated samples on the device (b200_fm_exec's d_decimated, or its
//     own CIC) can run discriminator + limiter + both half-bands without going back to u8 IQ.
//
//   common_sp.h:40-76    atan2_approx, evaluated in the reference's own order (quotient first, then the
//                        rational form, +-M_PI added in double): the inputs here are arbitrary int32,
//                        not the small CIC sums the fused kernels can treat exactly
//   audio_main.c:110-131 first difference (no unwrap), +-1 hard limiter
//   resample.c:47-67     halfband_decimate twice (audio_main.c:133,139), delay lines refreshed from the
//                        last ten inputs of each stage
//   audio_main.c:137-143 stage 2 runs only while the pool has a free buffer; when it does not, the
//                        block is dropped and delay_line_2 is NOT advanced (B200_FM_SKIP_STAGE2)
//
// One CTA per tile of 256 audio samples (1024 decimated samples); a tile recomputes its own halo
// (31 phases -> 30 discriminator outputs -> 10 first-stage outputs) from the input, the first tile of
// a block takes it from the carried state instead.  8 B in + 1 B out per decimated sample: HBM-bound.
#include <string.h>

#include "b200_common.cuh"
#include "fm_kernels.cuh"

namespace b200 {

namespace {

constexpr int CS_TILE_AUDIO = 256;
constexpr int CS_TILE_DEC = 4 * CS_TILE_AUDIO;      // 1024 decimated samples
constexpr int CS_THREADS = 256;
constexpr int CS_DEMOD_HALO = 30;
constexpr int CS_WORK_HALO = 10;

// layout of one stream's state, in floats (include/b200sdr.h documents it)
constexpr int ST_PREV = 0;          // previous phase (audio_main.c:79 prev_sample)
constexpr int ST_DELAY1 = 1;        // 10 floats (audio_main.c:77)
constexpr int ST_DELAY2 = 11;       // 10 floats (audio_main.c:78)
constexpr int ST_NEXT = 24;         // staging copy of [0..20] written by the block's last tile
constexpr int ST_FLOATS = B200_FM_STATE_FLOATS;
static_assert(ST_NEXT + 21 <= ST_FLOATS, "state staging area");

// common_sp.h:40-76 in its own order of operations
__device__ __forceinline__ float atan2_approx_seq(float y, float x)
{
    const float pi_by_2 = 1.57079632679489661923f;
    const double pi_d = 3.14159265358979323846;
    if (x == 0.0f) return y > 0.0f ? pi_by_2 : (y == 0.0f ? 0.0f : -pi_by_2);
    const float z = __fdiv_rn(y, x);
    if (fabsf(z) < 1.0f) {
        const float a = __fdiv_rn(z, 1.0f + 0.28f * z * z);
        if (x < 0.0f) return (float) (y < 0.0f ? (double) a - pi_d : (double) a + pi_d);
        return a;
    }
    const float a = pi_by_2 - __fdiv_rn(z, z * z + 0.28f);
    return y < 0.0f ? (float) ((double) a - pi_d) : a;
}

struct Cs32Params {
    const int2* dec;              // [n_streams][dec_stride] (re, im)
    int64_t dec_stride;           // complex samples
    int64_t n_dec;                // per stream, multiple of 4
    float* state;                 // [n_streams][ST_FLOATS]
    float* audio;                 // [n_streams][audio_stride], n_dec / 4 per stream; null with skip_stage2
    int64_t audio_stride;
    float* demod;                 // nullable: limiter output, [n_streams][demod_stride]
    int64_t demod_stride;
    float* phase;                 // nullable: atan2_approx output, [n_streams][phase_stride]
    int64_t phase_stride;
    int skip_stage2;
};

__global__ void __launch_bounds__(CS_THREADS) fm_cs32_kernel(const Cs32Params p)
{
    __shared__ float ph[CS_TILE_DEC + CS_DEMOD_HALO + 1];        // phase[g0 - 31 + i]
    __shared__ float dm[CS_TILE_DEC + CS_DEMOD_HALO];            // demod[g0 - 30 + i]
    __shared__ float wk[CS_TILE_DEC / 2 + CS_WORK_HALO];         // work[w0 - 10 + i]
    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const int64_t tile = blockIdx.x;
    const int64_t g0 = tile * CS_TILE_DEC;                       // first decimated sample of the tile
    const int n_here = (int) min((int64_t) CS_TILE_DEC, p.n_dec - g0);     // multiple of 4, > 0
    const int2* in = p.dec + (int64_t) s * p.dec_stride;
    float* st = p.state + (int64_t) s * ST_FLOATS;
    const bool first = tile == 0;
    const bool last = g0 + n_here == p.n_dec;

    // phases of the tile and of its 31-sample halo (the first tile has no input halo: state instead)
    for (int i = tid; i < n_here + CS_DEMOD_HALO + 1; i += CS_THREADS) {
        const int64_t g = g0 - (CS_DEMOD_HALO + 1) + i;
        float v = 0.0f;
        if (g >= 0) {
            const int2 c = in[g];
            v = atan2_approx_seq((float) c.y, (float) c.x);
            if (p.phase != nullptr && i > CS_DEMOD_HALO) p.phase[(int64_t) s * p.phase_stride + g] = v;
        } else if (g == -1) {
            v = st[ST_PREV];
        }
        ph[i] = v;
    }
    __syncthreads();
    // discriminator + limiter (audio_main.c:114-130)
    for (int i = tid; i < n_here + CS_DEMOD_HALO; i += CS_THREADS) {
        const int64_t g = g0 - CS_DEMOD_HALO + i;
        float v;
        if (g >= 0) {
            v = fm_limit(ph[i + 1], ph[i]);
            if (p.demod != nullptr && i >= CS_DEMOD_HALO) p.demod[(int64_t) s * p.demod_stride + g] = v;
        } else {
            v = g >= -10 ? st[ST_DELAY1 + 10 + (int) g] : 0.0f;       // delay_line_1 = the last ten demod values
        }
        dm[i] = v;
    }
    __syncthreads();
    // first half-band (audio_main.c:133): work[m] from demod[2m - 10 .. 2m]
    const int n_work = n_here / 2;
    for (int i = tid; i < n_work + CS_WORK_HALO; i += CS_THREADS) {
        const int64_t m = g0 / 2 - CS_WORK_HALO + i;
        float v;
        if (m >= 0) {
            const float* x = dm + CS_DEMOD_HALO + 2 * (i - CS_WORK_HALO);     // demod[2m]
            v = halfband_taps(x[0], x[-2], x[-4], x[-5], x[-6], x[-8], x[-10]);
        } else {
            v = st[ST_DELAY2 + 10 + (int) m];                                 // delay_line_2 = the last ten work values
        }
        wk[i] = v;
    }
    __syncthreads();
    // second half-band (audio_main.c:139)
    if (!p.skip_stage2) {
        const int n_audio = n_here / 4;
        float* out = p.audio + (int64_t) s * p.audio_stride + g0 / 4;
        for (int n = tid; n < n_audio; n += CS_THREADS) {
            const float* x = wk + CS_WORK_HALO + 2 * n;
            out[n] = halfband_taps(x[0], x[-2], x[-4], x[-5], x[-6], x[-8], x[-10]);
        }
    }
    // the block's last tile stages the next state (committed by fm_cs32_commit_kernel: the first
    // tile of this launch may still be reading the current one)
    if (last && tid < 21) {
        float v;
        if (tid == 0) {
            v = ph[n_here + CS_DEMOD_HALO];                                   // phase of the last sample
        } else if (tid <= 10) {
            // last ten demod values; a block shorter than ten keeps the newest old ones in front
            const int j = n_here - 10 + (tid - 1);                            // index relative to g0
            v = dm[CS_DEMOD_HALO + j];                                        // j >= -6: inside the halo
        } else {
            const int j = n_work - 10 + (tid - 11);
            v = p.skip_stage2 ? st[ST_DELAY2 + (tid - 11)] : wk[CS_WORK_HALO + j];
        }
        st[ST_NEXT + tid] = v;
    }
}

__global__ void fm_cs32_commit_kernel(float* state, int n_streams)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_streams * 21) return;
    float* st = state + (int64_t) (i / 21) * ST_FLOATS;
    st[i % 21] = st[ST_NEXT + i % 21];
}

// diagnostic: the three device forms of atan2_approx on arbitrary integer pairs (tests pin each of them
// to the reference's grid: tests/golden/atan2.npz)
__global__ void atan2_forms_kernel(const int2* __restrict__ yx, int n, float* __restrict__ out, int which)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float y = (float) yx[i].x, x = (float) yx[i].y;
    out[i] = which == 0 ? atan2_approx_dev(y, x) : (which == 1 ? atan2_approx_dev2(y, x) : atan2_approx_seq(y, x));
}

}  // namespace

int launch_fm_cs32(const int32_t* d_dec, int64_t dec_stride, int n_streams, int64_t n_dec, float* d_state,
                   float* d_audio, int64_t audio_stride, float* d_demod, int64_t demod_stride, float* d_phase,
                   int64_t phase_stride, int skip_stage2, cudaStream_t stream)
{
    if (n_streams == 0 || n_dec == 0) return B200_OK;
    const int64_t tiles = (n_dec + CS_TILE_DEC - 1) / CS_TILE_DEC;
    if (tiles >= (1ll << 31) || n_streams > 65535) {
        set_error("fm cs32: %lld tiles x %d streams exceed one launch", (long long) tiles, n_streams);
        return B200_ERR_ARG;
    }
    Cs32Params p;
    p.dec = reinterpret_cast<const int2*>(d_dec);
    p.dec_stride = dec_stride;
    p.n_dec = n_dec;
    p.state = d_state;
    p.audio = d_audio;
    p.audio_stride = audio_stride;
    p.demod = d_demod;
    p.demod_stride = demod_stride;
    p.phase = d_phase;
    p.phase_stride = phase_stride;
    p.skip_stage2 = skip_stage2;
    fm_cs32_kernel<<<dim3((unsigned) tiles, (unsigned) n_streams), CS_THREADS, 0, stream>>>(p);
    B200_LAUNCH_CHECK();
    fm_cs32_commit_kernel<<<(n_streams * 21 + 255) / 256, 256, 0, stream>>>(d_state, n_streams);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace b200

using namespace b200;

// ---- C ABI (include/b200sdr.h) -----------------------------------------------------------------

extern "C" {

int b200_fm_exec_cs32(const int32_t* d_decimated, int64_t dec_stride, int n_streams, int64_t n_decimated,
                      float* d_state, float* d_audio, int64_t audio_stride, float* d_demod, int64_t demod_stride,
                      float* d_phase, int64_t phase_stride, int flags, void* cuda_stream)
{
    const int skip = (flags & B200_FM_SKIP_STAGE2) != 0;
    if (d_decimated == nullptr || d_state == nullptr || n_streams < 0 || n_decimated < 0 || (!skip && d_audio == nullptr)) {
        set_error("fm exec cs32: bad arguments");
        return B200_ERR_ARG;
    }
    if (n_decimated % 4 != 0) {             // audio_main.c:90,100: len/2 and len/4 samples come out
        set_error("fm exec cs32: n_decimated = %lld is not a multiple of 4", (long long) n_decimated);
        return B200_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_decimated) & 7) != 0) {
        set_error("fm exec cs32: the cmplx_s32 input must be 8-byte aligned");
        return B200_ERR_ALIGN;
    }
    return launch_fm_cs32(d_decimated, dec_stride, n_streams, n_decimated, d_state, d_audio, audio_stride, d_demod,
                          demod_stride, d_phase, phase_stride, skip, reinterpret_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"

extern "C" int b200_debug_atan2(const int32_t* d_yx, int n, float* d_out, int which, void* cuda_stream)
{
    if (d_yx == nullptr || d_out == nullptr || n < 0 || which < 0 || which > 2) {
        set_error("debug atan2: bad arguments");
        return B200_ERR_ARG;
    }
    if (n == 0) return B200_OK;
    atan2_forms_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
        reinterpret_cast<const int2*>(d_yx), n, d_out, which);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// ---- host-buffer demodulator: what an rf_decimator_callback calls -----------------------------------

struct b200_fm_demod {
    int capacity;              // decimated samples per block the buffers hold
    int2* d_in;
    float* d_state;
    float* d_audio;
    float* d_demod;
    int2* h_in;                // pinned staging (the callback's pointer is borrowed, rf_decimator.c:31-36)
    float* h_out;              // pinned: [audio capacity/4 | demod capacity]
    cudaStream_t stream;
    int device;
};

static void fm_demod_release(b200_fm_demod* d)
{
    if (d->d_in) cudaFree(d->d_in);
    if (d->d_audio) cudaFree(d->d_audio);
    if (d->d_demod) cudaFree(d->d_demod);
    if (d->h_in) cudaFreeHost(d->h_in);
    if (d->h_out) cudaFreeHost(d->h_out);
    d->d_in = nullptr;
    d->d_audio = d->d_demod = nullptr;
    d->h_in = nullptr;
    d->h_out = nullptr;
    d->capacity = 0;
}

static bool fm_demod_reserve(b200_fm_demod* d, int len)
{
    if (len <= d->capacity) return true;
    fm_demod_release(d);
    bool ok = cudaMalloc((void**) &d->d_in, (size_t) len * 8) == cudaSuccess;
    ok = ok && cudaMalloc((void**) &d->d_audio, (size_t) len) == cudaSuccess;          // len / 4 floats
    ok = ok && cudaMalloc((void**) &d->d_demod, (size_t) len * 4) == cudaSuccess;
    ok = ok && cudaHostAlloc((void**) &d->h_in, (size_t) len * 8, cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc((void**) &d->h_out, (size_t) len * 5, cudaHostAllocDefault) == cudaSuccess;
    if (ok) d->capacity = len;
    return ok;
}

extern "C" {

b200_fm_demod* b200_fm_demod_create(void)
{
    b200_fm_demod* d = new b200_fm_demod();
    memset(d, 0, sizeof(*d));
    bool ok = cudaGetDevice(&d->device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc((void**) &d->d_state, sizeof(float) * B200_FM_STATE_FLOATS) == cudaSuccess;
    ok = ok && cudaMemset(d->d_state, 0, sizeof(float) * B200_FM_STATE_FLOATS) == cudaSuccess;     // audio_main.c:77-79
    if (!ok) {
        set_error("fm demod: %s", cudaGetErrorString(cudaGetLastError()));
        b200_fm_demod_destroy(d);
        return nullptr;
    }
    return d;
}

void b200_fm_demod_destroy(b200_fm_demod* d)
{
    if (d == nullptr) return;
    fm_demod_release(d);
    if (d->d_state) cudaFree(d->d_state);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

int b200_fm_demod_reset(b200_fm_demod* d)
{
    if (d == nullptr) return B200_ERR_ARG;
    B200_CUDA_TRY(cudaMemsetAsync(d->d_state, 0, sizeof(float) * B200_FM_STATE_FLOATS, d->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(d->stream));
    return B200_OK;
}

int b200_fm_demod_block(b200_fm_demod* d, const int32_t* h_signal, int len, float* h_audio, float* h_demod)
{
    if (d == nullptr || h_signal == nullptr || len < 0 || len % 4 != 0) {
        set_error("fm demod block: len = %d must be a non-negative multiple of 4", len);
        return B200_ERR_ARG;
    }
    if (len == 0) return B200_OK;
    if (!fm_demod_reserve(d, len)) {
        set_error("fm demod block: %s", cudaGetErrorString(cudaGetLastError()));
        return B200_ERR_CUDA;
    }
    memcpy(d->h_in, h_signal, (size_t) len * 8);
    B200_CUDA_TRY(cudaMemcpyAsync(d->d_in, d->h_in, (size_t) len * 8, cudaMemcpyHostToDevice, d->stream));
    const int rc = launch_fm_cs32(reinterpret_cast<const int32_t*>(d->d_in), len, 1, len, d->d_state, d->d_audio, len / 4,
                                  h_demod ? d->d_demod : nullptr, len, nullptr, 0, h_audio == nullptr, d->stream);
    if (rc) return rc;
    float* h_a = d->h_out;
    float* h_d = d->h_out + len / 4;
    if (h_audio) B200_CUDA_TRY(cudaMemcpyAsync(h_a, d->d_audio, (size_t) len, cudaMemcpyDeviceToHost, d->stream));
    if (h_demod) B200_CUDA_TRY(cudaMemcpyAsync(h_d, d->d_demod, (size_t) len * 4, cudaMemcpyDeviceToHost, d->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(d->stream));
    if (h_audio) memcpy(h_audio, h_a, (size_t) len);
    if (h_demod) memcpy(h_demod, h_d, (size_t) len * 4);
    return B200_OK;
}

}  // extern "C"
