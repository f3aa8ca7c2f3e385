// audio_compat.cpp -- libb200audio.so: the reference's audio_main.h interface over the GPU demodulator.
//
// Links in place of audio_main.o: an unmodified main.c calls audio_init() (main.c:197), registers
// audio_fm_demodulator as the decimator's callback (main.c:205) and drains the pool from its
// writeable callback through audio_new_audio_available / audio_get_audio_payload (main.c:86-110).
// Host code only (no CUDA headers): the arithmetic of audio_main.c:110-139 -- atan2_approx, first
// difference, limiter, both half-band decimators -- runs in libb200sdr.so's fm_cs32_kernel through
// b200_fm_demod_block; what stays here is the plumbing of audio_main.c:25-72,83-108,135-144:
//
//   * a pool of AUDIO_BUFFER_POOL = 50 buffers of len / 4 floats, rebuilt (and everything queued
//     dropped) whenever the callback's block length changes (audio_main.c:83-108);
//   * a block whose turn finds no free buffer is dropped, and the second half-band's delay line is
//     not advanced for it (audio_main.c:137-143: stage 2 sits inside the `if`);
//   * audio_get_audio_payload keeps copying from the buffer it peeked BEFORE it moved that buffer back
//     to the free list (audio_main.c:49-63), so at every buffer boundary the next call starts with the
//     old buffer's samples again.  Reproduced deliberately: it is what the reference puts on the wire
//     (tests/golden/ws_stream.npz), and b200_audio_take_buffer below is the clean buffer-level exit.
//
// The demodulator state (previous phase, two delay lines) lives for the life of the process, like the
// function statics of audio_main.c:77-79: audio_close() / audio_init() do not reset it.
#include <pthread.h>
#include <stdio.h>
#include <string.h>

#include <deque>
#include <vector>

#include "../../include/b200sdr.h"
#include "../../include/rtlws_audio_compat.h"

namespace {

constexpr int AUDIO_BUFFER_POOL = 50;                // audio_main.c:11

struct Pool {
    pthread_mutex_t mutex;
    bool mutex_ready = false;
    std::deque<float*> full;                         // finished buffers, oldest first
    std::deque<float*> used;                         // free buffers
    std::vector<float*> owned;
    int buffer_len = 0;                              // floats per buffer (audio_main.c:100)
    int block_len = 0;                               // demod_buffer_len (audio_main.c:83)
    int cur_idx = 0;                                 // audio_get_audio_payload's static cursor (audio_main.c:42)
    b200_fm_demod* demod = nullptr;
    long long dropped = 0;
};

Pool g_pool;

void free_buffers(Pool& p)
{
    for (float* b : p.owned) delete[] b;
    p.owned.clear();
    p.full.clear();
    p.used.clear();
}

}  // namespace

extern "C" {

void audio_init()
{
    Pool& p = g_pool;
    if (!p.mutex_ready) {
        pthread_mutex_init(&p.mutex, NULL);
        p.mutex_ready = true;
    }
    if (p.demod == nullptr) {
        p.demod = b200_fm_demod_create();
        if (p.demod == nullptr) fprintf(stderr, "libb200audio [E] audio_init: %s\n", b200_last_error());
    }
}

int audio_new_audio_available()
{
    return !g_pool.full.empty();                     // audio_main.c:37 reads the list length without the lock too
}

int audio_get_audio_payload(char* buf, int buf_len)
{
    Pool& p = g_pool;
    float* data = nullptr;
    int copied = 0;
    int room = buf_len / (int) sizeof(float);
    do {
        pthread_mutex_lock(&p.mutex);
        data = p.full.empty() ? nullptr : p.full.front();
        if (data != nullptr) {
            if (p.cur_idx >= p.buffer_len) {
                // the finished buffer goes back to the free list, but `data` still points at it
                p.used.push_back(p.full.front());
                p.full.pop_front();
                p.cur_idx = 0;
            }
            int n = p.buffer_len - p.cur_idx;
            if (n > room) n = room;
            if (n > 0) {
                // audio_main.c:59 indexes the char buffer with the SAMPLE count: the offset is in bytes
                memcpy(buf + copied, data + p.cur_idx, (size_t) n * sizeof(float));
                p.cur_idx += n;
                copied += n;
                room -= n;
            }
        }
        pthread_mutex_unlock(&p.mutex);
    } while (data != nullptr && room > 0);
    return copied * (int) sizeof(float);
}

void audio_fm_demodulator(const cmplx_s32* signal, int len)
{
    Pool& p = g_pool;
    if (p.demod == nullptr || len <= 0) return;
    if (p.block_len != len) {                        // audio_main.c:83-108
        pthread_mutex_lock(&p.mutex);
        free_buffers(p);
        p.block_len = len;
        p.buffer_len = (len / 2) / 2;
        for (int i = 0; i < AUDIO_BUFFER_POOL; ++i) {
            float* b = new float[p.buffer_len > 0 ? p.buffer_len : 1]();
            p.owned.push_back(b);
            p.used.push_back(b);
        }
        pthread_mutex_unlock(&p.mutex);
    }
    // The kernel runs outside the lock (the reference computes the discriminator and the first half-band
    // outside it as well, audio_main.c:110-133); only this thread ever takes buffers from `used`, so a
    // buffer seen free here is still free when the block comes back.
    pthread_mutex_lock(&p.mutex);
    float* dst = p.used.empty() ? nullptr : p.used.front();
    pthread_mutex_unlock(&p.mutex);
    const int rc = b200_fm_demod_block(p.demod, reinterpret_cast<const int32_t*>(signal), len & ~3, dst, nullptr);
    if (rc != B200_OK) {
        fprintf(stderr, "libb200audio [E] audio_fm_demodulator: %s\n", b200_last_error());
        return;
    }
    if (dst == nullptr) {
        ++p.dropped;
        return;
    }
    pthread_mutex_lock(&p.mutex);
    p.used.pop_front();
    p.full.push_back(dst);
    pthread_mutex_unlock(&p.mutex);
}

void audio_close()
{
    Pool& p = g_pool;
    if (p.mutex_ready) pthread_mutex_lock(&p.mutex);
    free_buffers(p);
    p.block_len = 0;
    p.buffer_len = 0;
    if (p.mutex_ready) pthread_mutex_unlock(&p.mutex);
}

/* ---- extensions (not in audio_main.h) ---- */

int b200_audio_take_buffer(float* out, int max_floats)
{
    Pool& p = g_pool;
    int n = -1;
    pthread_mutex_lock(&p.mutex);
    if (!p.full.empty() && p.buffer_len <= max_floats) {
        n = p.buffer_len;
        memcpy(out, p.full.front(), (size_t) n * sizeof(float));
        p.used.push_back(p.full.front());
        p.full.pop_front();
    }
    pthread_mutex_unlock(&p.mutex);
    return n;
}

int b200_audio_buffer_len(void)
{
    return g_pool.buffer_len;
}

long long b200_audio_dropped_blocks(void)
{
    return g_pool.dropped;
}

void b200_audio_reset_stream(void)
{
    Pool& p = g_pool;
    if (p.demod != nullptr) b200_fm_demod_reset(p.demod);
    if (p.mutex_ready) pthread_mutex_lock(&p.mutex);
    while (!p.full.empty()) {
        p.used.push_back(p.full.front());
        p.full.pop_front();
    }
    p.cur_idx = 0;
    p.dropped = 0;
    if (p.mutex_ready) pthread_mutex_unlock(&p.mutex);
}

}  // extern "C"
