// wire.cu -- the reference's websocket wire formats (main.c:74-111), emitted for many streams
// at once so that what comes back over PCIe is ready for lws_write():
//   spectrum message  "t s;f %u;b %u;s %d;d" + n_bins payload bytes         main.c:80-84
//   audio message     "FF;t a;d" + 8 fragments of 2048 bytes of float32     main.c:86-110
// The payload bytes are whatever the spectrum kernels wrote (cbb_main.c:125-128 arithmetic),
// the floats are the FM branch's output; nothing is recomputed here, this is layout only.
// The consumer that parses these is rtl_ui.js:98-141.
#include <string.h>

#include <mutex>
#include <vector>

#include "b200_common.cuh"

namespace b200 {

namespace {

constexpr int AUDIO_HDR = 8;
constexpr int AUDIO_FLOATS = B200_WIRE_AUDIO_FRAGMENTS * (B200_WIRE_AUDIO_FRAGMENT_BYTES / 4);   // 4096

// audio_get_audio_payload (audio_main.c:40-72) walks pool buffers of `buffer_len` floats with a
// static cursor.  When the cursor has reached the end of the head buffer, the call that notices
// it moves that buffer to the used list and resets the cursor -- but keeps copying from the
// pointer it peeked BEFORE the move (audio_main.c:50-62).  Net effect on the wire: the first
// `call` samples (512 = 2048 bytes / sizeof(float), main.c:99) of every buffer after the first
// are replaced by the first `call` samples of the buffer before it.
__host__ __device__ __forceinline__ int64_t drain_index(int64_t w, int buffer_len, int call)
{
    if (w < buffer_len) return w;
    return (w % buffer_len) < call ? w - buffer_len : w;
}

// One CTA per stream.  Words of the message are produced 4 bytes at a time: header bytes come
// from the staged header block, payload bytes from two aligned payload words merged by PRMT
// (the payload starts at byte `hlen`, which is not a multiple of 4).
__global__ void wire_spectrum_kernel(const uint8_t* __restrict__ payload, int64_t payload_stride, int n_bins,
                                     const uint8_t* __restrict__ headers, const int32_t* __restrict__ hlens,
                                     uint8_t* __restrict__ msgs, int64_t msg_stride)
{
    const int s = blockIdx.x;
    const int hlen = hlens[s];
    const uint8_t* hdr = headers + (size_t) s * B200_WIRE_SPECTRUM_HEADER_MAX;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(payload + (int64_t) s * payload_stride);
    uint32_t* dst = reinterpret_cast<uint32_t*>(msgs + (int64_t) s * msg_stride);
    const int total = hlen + n_bins;
    const int n_words = (total + 3) / 4;
    const int src_words = n_bins / 4;
    for (int w = threadIdx.x; w < n_words; w += blockDim.x) {
        const int byte0 = 4 * w;
        uint32_t out = 0;
        if (byte0 + 4 <= hlen) {
            out = *reinterpret_cast<const uint32_t*>(hdr + byte0);
        } else if (byte0 >= hlen) {
            const int p = byte0 - hlen;                     // payload byte offset of this word
            const int pw = p >> 2;
            const uint32_t lo = src[pw];
            const uint32_t hi = pw + 1 < src_words ? src[pw + 1] : 0u;
            out = __funnelshift_r(lo, hi, 8 * (p & 3));
            const int valid = n_bins - p;                   // bytes of this word that exist
            if (valid < 4) out &= (1u << (8 * valid)) - 1u;
        } else {                                            // the word straddles header and payload
            const int nh = hlen - byte0;                    // 1..3 header bytes
            uint32_t h = 0;
            for (int i = 0; i < nh; ++i) h |= (uint32_t) hdr[byte0 + i] << (8 * i);
            out = h | (src[0] << (8 * nh));
        }
        dst[w] = out;
    }
}

// One thread per float2 of message body.  first_wire_sample and the drain boundaries are even,
// so a pair never straddles a remapped region.
__global__ void wire_audio_kernel(const float* __restrict__ audio, int64_t audio_stride, int64_t first_wire_sample,
                                  int n_messages, int drain, int buffer_len, uint8_t* __restrict__ msgs,
                                  int64_t msg_stride, int64_t message_pitch)
{
    const int s = blockIdx.y;
    const int m = blockIdx.x;
    uint8_t* dst = msgs + (int64_t) s * msg_stride + (int64_t) m * message_pitch;
    if (threadIdx.x < 2) {
        // "FF;t a;d" as two little-endian words
        reinterpret_cast<uint32_t*>(dst)[threadIdx.x] = threadIdx.x == 0 ? 0x743B4646u : 0x643B6120u;
    }
    const float* src = audio + (int64_t) s * audio_stride;
    const int64_t w0 = first_wire_sample + (int64_t) m * AUDIO_FLOATS;
    for (int i = 2 * threadIdx.x; i < AUDIO_FLOATS; i += 2 * blockDim.x) {
        const int64_t w = w0 + i;
        const int64_t j = drain ? drain_index(w, buffer_len, B200_WIRE_AUDIO_FRAGMENT_BYTES / 4) : w;
        const float2 v = *reinterpret_cast<const float2*>(src + j);
        *reinterpret_cast<float2*>(dst + AUDIO_HDR + 4 * (size_t) i) = v;
    }
}

// pinned + device staging for the formatted headers, grown on demand, per calling thread
struct HeaderStage {
    uint8_t* h = nullptr;
    uint8_t* d = nullptr;
    int32_t* h_len = nullptr;
    int32_t* d_len = nullptr;
    int cap = 0;
    int dev = -1;
    cudaEvent_t done = nullptr;
    ~HeaderStage()
    {
        // the context may already be gone at thread / process exit: ignore errors
        if (h) cudaFreeHost(h);
        if (h_len) cudaFreeHost(h_len);
        if (d) cudaFree(d);
        if (d_len) cudaFree(d_len);
        if (done) cudaEventDestroy(done);
        (void) cudaGetLastError();
    }
};

int stage_reserve(HeaderStage& st, int n)
{
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (st.dev == dev && st.cap >= n) return B200_OK;
    if (st.h) cudaFreeHost(st.h);
    if (st.h_len) cudaFreeHost(st.h_len);
    if (st.d) cudaFree(st.d);
    if (st.d_len) cudaFree(st.d_len);
    if (st.done && st.dev != dev) {          // an event belongs to the device it was created on
        cudaEventDestroy(st.done);
        st.done = nullptr;
    }
    st.h = st.d = nullptr;
    st.h_len = st.d_len = nullptr;
    st.cap = 0;
    const int cap = n < 256 ? 256 : n;
    B200_CUDA_TRY(cudaMallocHost(&st.h, (size_t) cap * B200_WIRE_SPECTRUM_HEADER_MAX));
    B200_CUDA_TRY(cudaMallocHost(&st.h_len, (size_t) cap * sizeof(int32_t)));
    B200_CUDA_TRY(cudaMalloc(&st.d, (size_t) cap * B200_WIRE_SPECTRUM_HEADER_MAX));
    B200_CUDA_TRY(cudaMalloc(&st.d_len, (size_t) cap * sizeof(int32_t)));
    if (!st.done) B200_CUDA_TRY(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    st.cap = cap;
    st.dev = dev;
    return B200_OK;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_wire_spectrum_header(char* dst, int dst_len, uint32_t freq_hz, uint32_t sample_rate_hz, int gain_db)
{
    char tmp[B200_WIRE_SPECTRUM_HEADER_MAX];
    // main.c:81, verbatim format.  The widest case is 6 + 10 + 3 + 10 + 3 + 11 + 2 = 45 characters.
    const int n = snprintf(tmp, sizeof(tmp), "t s;f %u;b %u;s %d;d", freq_hz, sample_rate_hz, gain_db);
    if (n < 0 || n >= (int) sizeof(tmp) || dst == nullptr || n > dst_len) {
        set_error("b200_wire_spectrum_header: %d bytes do not fit in %d", n, dst_len);
        return B200_ERR_ARG;
    }
    memcpy(dst, tmp, (size_t) n);
    return n;
}

int b200_wire_spectrum_message(uint8_t* dst, int dst_len, uint32_t freq_hz, uint32_t sample_rate_hz, int gain_db,
                               const uint8_t* payload, int n_bins)
{
    if (payload == nullptr || n_bins < 0) {
        set_error("b200_wire_spectrum_message: bad payload");
        return B200_ERR_ARG;
    }
    const int n = b200_wire_spectrum_header(reinterpret_cast<char*>(dst), dst_len, freq_hz, sample_rate_hz, gain_db);
    if (n < 0) return n;
    if (n + n_bins > dst_len) {
        set_error("b200_wire_spectrum_message: %d bytes do not fit in %d", n + n_bins, dst_len);
        return B200_ERR_ARG;
    }
    memcpy(dst + n, payload, (size_t) n_bins);
    return n + n_bins;
}

int b200_wire_spectrum_messages(const uint8_t* d_payload, int64_t payload_stride, int n_streams, int n_bins,
                                const uint32_t* freq_hz, const uint32_t* sample_rate_hz, const int32_t* gain_db,
                                uint8_t* d_msgs, int64_t msg_stride, int32_t* lens, void* cuda_stream)
{
    if (n_streams < 0 || n_bins <= 0 || (n_bins & 3) || d_payload == nullptr || d_msgs == nullptr ||
        freq_hz == nullptr || sample_rate_hz == nullptr || gain_db == nullptr) {
        set_error("b200_wire_spectrum_messages: bad arguments (n_bins must be a positive multiple of 4)");
        return B200_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_payload) & 3) || (payload_stride & 3) || (reinterpret_cast<uintptr_t>(d_msgs) & 3) ||
        (msg_stride & 3)) {
        set_error("b200_wire_spectrum_messages: payload / message pointers and strides must be multiples of 4 bytes");
        return B200_ERR_ALIGN;
    }
    if (msg_stride < (int64_t) B200_WIRE_SPECTRUM_HEADER_MAX + n_bins) {
        set_error("b200_wire_spectrum_messages: msg_stride %lld < %d", (long long) msg_stride,
                  B200_WIRE_SPECTRUM_HEADER_MAX + n_bins);
        return B200_ERR_ARG;
    }
    if (n_streams == 0) return B200_OK;
    static thread_local HeaderStage st;
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    if (st.done && st.cap > 0) B200_CUDA_TRY(cudaEventSynchronize(st.done));   // previous call's upload has left the pinned block
    if (int rc = stage_reserve(st, n_streams)) return rc;
    memset(st.h, 0, (size_t) n_streams * B200_WIRE_SPECTRUM_HEADER_MAX);
    for (int s = 0; s < n_streams; ++s) {
        const int n = b200_wire_spectrum_header(reinterpret_cast<char*>(st.h) + (size_t) s * B200_WIRE_SPECTRUM_HEADER_MAX,
                                                B200_WIRE_SPECTRUM_HEADER_MAX, freq_hz[s], sample_rate_hz[s], gain_db[s]);
        if (n < 0) return n;
        st.h_len[s] = n;
        if (lens) lens[s] = n + n_bins;
    }
    B200_CUDA_TRY(cudaMemcpyAsync(st.d, st.h, (size_t) n_streams * B200_WIRE_SPECTRUM_HEADER_MAX, cudaMemcpyHostToDevice, stream));
    B200_CUDA_TRY(cudaMemcpyAsync(st.d_len, st.h_len, (size_t) n_streams * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    B200_CUDA_TRY(cudaEventRecord(st.done, stream));
    wire_spectrum_kernel<<<n_streams, 128, 0, stream>>>(d_payload, payload_stride, n_bins, st.d, st.d_len, d_msgs, msg_stride);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int64_t b200_wire_reference_drain_index(int64_t wire_sample, int buffer_len)
{
    if (wire_sample < 0 || buffer_len <= 0) return -1;
    return drain_index(wire_sample, buffer_len, B200_WIRE_AUDIO_FRAGMENT_BYTES / 4);
}

int b200_wire_audio_fragment(int index, int32_t* offset, int32_t* len, int32_t* flags)
{
    if (index < 0 || index >= B200_WIRE_AUDIO_FRAGMENTS) {
        set_error("b200_wire_audio_fragment: index %d", index);
        return B200_ERR_ARG;
    }
    // main.c:89-110: the first write carries the 8-byte header and opens a binary message, the next six
    // continue it, the eighth closes it (no NO_FIN)
    const int32_t off = index == 0 ? 0 : AUDIO_HDR + index * B200_WIRE_AUDIO_FRAGMENT_BYTES;
    const int32_t n = B200_WIRE_AUDIO_FRAGMENT_BYTES + (index == 0 ? AUDIO_HDR : 0);
    int32_t f = index == 0 ? B200_WIRE_BINARY : B200_WIRE_CONTINUATION;
    if (index < B200_WIRE_AUDIO_FRAGMENTS - 1) f |= B200_WIRE_NO_FIN;
    if (offset) *offset = off;
    if (len) *len = n;
    if (flags) *flags = f;
    return B200_OK;
}

int b200_wire_audio_messages(const float* d_audio, int64_t audio_stride, int n_streams, int64_t first_wire_sample,
                             int n_messages, int flags, int buffer_len, uint8_t* d_msgs, int64_t msg_stride,
                             void* cuda_stream)
{
    const int drain = (flags & B200_WIRE_REFERENCE_DRAIN) != 0;
    if (n_streams < 0 || n_messages < 0 || d_audio == nullptr || d_msgs == nullptr || first_wire_sample < 0 ||
        (first_wire_sample & 1) || (drain && (buffer_len < B200_WIRE_AUDIO_FRAGMENT_BYTES / 4 || (buffer_len & 1)))) {
        set_error("b200_wire_audio_messages: bad arguments (first_wire_sample and buffer_len must be even)");
        return B200_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_audio) & 7) || (audio_stride & 1) || (reinterpret_cast<uintptr_t>(d_msgs) & 7) ||
        (msg_stride & 7)) {
        set_error("b200_wire_audio_messages: audio / message pointers and strides must be multiples of 8 bytes");
        return B200_ERR_ALIGN;
    }
    // drain_index() models the reference's cursor for pool buffers that a whole number of 512-sample calls fills
    // (5120 at R = 10, rf_decimator.c:65-66 / audio_main.c:90); at other lengths the call that reaches the buffer end
    // copies a partial chunk, which is not modelled -- refuse instead of emitting a stream that differs from the reference
    if (drain && buffer_len % (B200_WIRE_AUDIO_FRAGMENT_BYTES / 4) != 0) {
        set_error("b200_wire_audio_messages: B200_WIRE_REFERENCE_DRAIN needs buffer_len to be a multiple of %d (got %d)",
                  B200_WIRE_AUDIO_FRAGMENT_BYTES / 4, buffer_len);
        return B200_ERR_ARG;
    }
    // a row of d_audio holds at most audio_stride floats: the messages may not read past it
    if (first_wire_sample + (int64_t) n_messages * AUDIO_FLOATS > audio_stride) {
        set_error("b200_wire_audio_messages: wire samples %lld .. %lld lie outside an audio row of %lld floats",
                  (long long) first_wire_sample, (long long) (first_wire_sample + (int64_t) n_messages * AUDIO_FLOATS),
                  (long long) audio_stride);
        return B200_ERR_ARG;
    }
    if (msg_stride < (int64_t) n_messages * B200_WIRE_AUDIO_MESSAGE_BYTES) {
        set_error("b200_wire_audio_messages: msg_stride %lld < %lld", (long long) msg_stride,
                  (long long) n_messages * B200_WIRE_AUDIO_MESSAGE_BYTES);
        return B200_ERR_ARG;
    }
    if (n_streams == 0 || n_messages == 0) return B200_OK;
    if (n_streams > 65535) {
        set_error("b200_wire_audio_messages: at most 65535 streams per call");
        return B200_ERR_ARG;
    }
    wire_audio_kernel<<<dim3((unsigned) n_messages, (unsigned) n_streams), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_audio, audio_stride, first_wire_sample, n_messages, drain, buffer_len, d_msgs, msg_stride,
        B200_WIRE_AUDIO_MESSAGE_BYTES);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

}  // extern "C"
