// replay_sensor.cpp -- libb200replay.so: rtl_sensor.h's ten functions over a replayed capture
// (include/rtl_sensor_replay.h).  Host code only.
//
// What it stands in for: rtl_sensor.c with REAL_SENSOR forwards to librtlsdr -- rtlsdr_open,
// rtlsdr_set_sample_rate / _center_freq / _tuner_gain, rtlsdr_read_async(dev, cb, user, 0, 0) --
// whose asynchronous reader fills a ring of 15 buffers of 262144 bytes from USB and calls `cb`
// once per full buffer on the calling thread until rtlsdr_cancel_async.  Here the "USB" is a
// capture in host memory, copied buffer by buffer into the same kind of ring (the callback may
// keep no pointer past its return: the next buffers overwrite the ring, as with the dongle).
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <mutex>

#include "../../include/rtl_sensor_replay.h"

namespace {

struct Capture {
    const uint8_t* iq = nullptr;
    int64_t n_bytes = 0;
    int loops = 1;
    int realtime = 0;
    std::atomic<int64_t> delivered{0};
    std::atomic<int> gate_closed{0};
};

std::mutex g_mutex;                               // guards the capture table's plain fields
Capture g_captures[B200_REPLAY_MAX_DEVICES];

void sleep_until(const timespec& t0, double seconds)
{
    timespec t = t0;
    const int64_t ns = (int64_t) (seconds * 1e9);
    t.tv_sec += ns / 1000000000ll;
    t.tv_nsec += ns % 1000000000ll;
    if (t.tv_nsec >= 1000000000l) {
        t.tv_nsec -= 1000000000l;
        t.tv_sec += 1;
    }
    while (clock_nanosleep(CLOCK_MONOTONIC, TIMER_ABSTIME, &t, nullptr) != 0) {
    }
}

}  // namespace

struct rtl_dev {
    int index;
    uint32_t f;
    uint32_t fs;
    double gain;
    std::atomic<int> cancel;
    uint8_t* ring;                                // B200_REPLAY_BUFFERS x B200_REPLAY_BUFFER_BYTES
};

extern "C" {

int rtl_init(struct rtl_dev** dev, int dev_index)
{
    if (dev == nullptr) return -1;
    *dev = nullptr;
    if (dev_index < 0 || dev_index >= B200_REPLAY_MAX_DEVICES) return -1;
    rtl_dev* d = new rtl_dev();
    d->index = dev_index;
    d->fs = 2048000;            // rtl_sensor.c:12
    d->f = 100000000;           // rtl_sensor.c:13
    d->gain = 25.4;             // rtl_sensor.c:14
    d->cancel.store(0);
    d->ring = nullptr;
    if (posix_memalign(reinterpret_cast<void**>(&d->ring), 4096, (size_t) B200_REPLAY_BUFFERS * B200_REPLAY_BUFFER_BYTES) != 0) {
        delete d;
        return -1;
    }
    *dev = d;
    return 0;
}

int rtl_set_frequency(struct rtl_dev* dev, uint32_t f)
{
    dev->f = f;
    return 0;
}

int rtl_set_sample_rate(struct rtl_dev* dev, uint32_t fs)
{
    if (fs == 0) return -1;
    dev->fs = fs;
    return 0;
}

int rtl_set_gain(struct rtl_dev* dev, double gain)
{
    dev->gain = gain;
    return 0;
}

uint32_t rtl_freq(const struct rtl_dev* dev) { return dev->f; }
uint32_t rtl_sample_rate(const struct rtl_dev* dev) { return dev->fs; }
double rtl_gain(const struct rtl_dev* dev) { return dev->gain; }

int rtl_read_async(struct rtl_dev* dev, void (*callback)(unsigned char*, uint32_t, void*), void* user)
{
    if (dev == nullptr || callback == nullptr) return -1;
    const uint8_t* iq;
    int64_t n_bytes;
    int loops, realtime;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        const Capture& c = g_captures[dev->index];
        iq = c.iq;
        n_bytes = c.n_bytes;
        loops = c.loops;
        realtime = c.realtime;
    }
    if (iq == nullptr || n_bytes < B200_REPLAY_BUFFER_BYTES) return 0;      // nothing to deliver: like the stub
    const int64_t per_pass = n_bytes / B200_REPLAY_BUFFER_BYTES;
    // a closed gate holds the first buffer back (signal_source_start spawns the reader before cbb_main.c:87-88
    // has registered its callbacks; a dongle's first buffer is 64 ms away, a replay's is not)
    while (g_captures[dev->index].gate_closed.load(std::memory_order_acquire) && !dev->cancel.load(std::memory_order_acquire)) {
        timespec nap = {0, 1000000};
        nanosleep(&nap, nullptr);
    }
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int64_t n = 0;
    for (int pass = 0; (loops <= 0 || pass < loops) && !dev->cancel.load(std::memory_order_acquire); ++pass) {
        for (int64_t b = 0; b < per_pass && !dev->cancel.load(std::memory_order_acquire); ++b, ++n) {
            uint8_t* buf = dev->ring + (size_t) (n % B200_REPLAY_BUFFERS) * B200_REPLAY_BUFFER_BYTES;
            memcpy(buf, iq + b * B200_REPLAY_BUFFER_BYTES, B200_REPLAY_BUFFER_BYTES);
            if (realtime)      // the buffer is complete when its last sample has been taken
                sleep_until(t0, (double) (n + 1) * (B200_REPLAY_BUFFER_BYTES / 2) / (double) dev->fs);
            callback(buf, B200_REPLAY_BUFFER_BYTES, user);
            g_captures[dev->index].delivered.fetch_add(B200_REPLAY_BUFFER_BYTES, std::memory_order_relaxed);
        }
    }
    return 0;
}

void rtl_cancel(struct rtl_dev* dev)
{
    if (dev != nullptr) dev->cancel.store(1, std::memory_order_release);
}

void rtl_close(struct rtl_dev* dev)
{
    if (dev == nullptr) return;
    free(dev->ring);
    delete dev;
}

int b200_replay_set_capture(int dev_index, const uint8_t* iq, int64_t n_bytes, int loops, int realtime)
{
    if (dev_index < 0 || dev_index >= B200_REPLAY_MAX_DEVICES || n_bytes < 0 || (iq == nullptr && n_bytes > 0)) return -1;
    std::lock_guard<std::mutex> lock(g_mutex);
    Capture& c = g_captures[dev_index];
    c.iq = iq;
    c.n_bytes = n_bytes;
    c.loops = loops;
    c.realtime = realtime;
    c.delivered.store(0);
    return 0;
}

int b200_replay_gate(int dev_index, int open)
{
    if (dev_index < 0 || dev_index >= B200_REPLAY_MAX_DEVICES) return -1;
    g_captures[dev_index].gate_closed.store(open ? 0 : 1, std::memory_order_release);
    return 0;
}

int64_t b200_replay_delivered_bytes(int dev_index)
{
    if (dev_index < 0 || dev_index >= B200_REPLAY_MAX_DEVICES) return -1;
    return g_captures[dev_index].delivered.load();
}

}  // extern "C"
