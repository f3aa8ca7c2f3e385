/*
 * rtl_sensor_replay.h -- libb200replay.so: the reference's sensor interface (src/rtl_sensor.h:9-27)
 * implemented over a replayed capture instead of a dongle, for any number of virtual dongles.
 *
 * The reference talks to the RTL-SDR through ten functions (rtl_sensor.h) and, built without
 * REAL_SENSOR, its own rtl_sensor.c is an empty stub whose rtl_read_async returns at once
 * (rtl_sensor.c:146-153).  This library exports the same ten symbols; linked in place of
 * rtl_sensor.o, an unmodified signal_source.c / cbb_main.c / main.c sees a sensor that delivers
 * a caller-supplied capture exactly as librtlsdr's asynchronous reader does with the 0, 0
 * defaults the reference passes (rtl_sensor.c:149): full buffers of 262144 bytes out of a ring of
 * 15, one callback per buffer on the thread that called rtl_read_async, until rtl_cancel or the end
 * of the capture.  It is host code only (no CUDA); it lives in its own library so that
 * libb200sdr.so does not export rtl_* names.
 *
 * Defaults follow rtl_sensor.c:12-14: 2 048 000 S/s, 100 MHz, gain 25.4.
 */
#ifndef RTL_SENSOR_REPLAY_H
#define RTL_SENSOR_REPLAY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the reference's interface, same names and meanings (rtl_sensor.h:9-27) ---- */
struct rtl_dev;
int rtl_init(struct rtl_dev** dev, int dev_index);
int rtl_set_frequency(struct rtl_dev* dev, uint32_t f);
int rtl_set_sample_rate(struct rtl_dev* dev, uint32_t fs);
int rtl_set_gain(struct rtl_dev* dev, double gain);
uint32_t rtl_freq(const struct rtl_dev* dev);
uint32_t rtl_sample_rate(const struct rtl_dev* dev);
double rtl_gain(const struct rtl_dev* dev);
int rtl_read_async(struct rtl_dev* dev, void (*callback)(unsigned char*, uint32_t, void*), void* user);
void rtl_cancel(struct rtl_dev* dev);
void rtl_close(struct rtl_dev* dev);

/* ---- what a virtual dongle replays ---- */
#define B200_REPLAY_MAX_DEVICES 1024
#define B200_REPLAY_BUFFER_BYTES 262144   /* librtlsdr's default buffer: 131072 cmplx_u8 per callback */
#define B200_REPLAY_BUFFERS 15            /* librtlsdr's default ring */

/* Capture of virtual dongle `dev_index`: n_bytes of interleaved u8 IQ at `iq` (borrowed: it must stay
 * valid while a reader runs).  loops: how many passes over it (<= 0: until rtl_cancel).  realtime != 0:
 * buffers are paced at the device's sample rate (one per 64 ms at 2.048 MS/s), else back to back.
 * Only whole buffers are delivered; a tail shorter than one buffer is dropped, as a dongle would never
 * deliver it.  Returns 0, or -1 for a bad index / size. */
int b200_replay_set_capture(int dev_index, const uint8_t* iq, int64_t n_bytes, int loops, int realtime);
/* open = 0: readers of `dev_index` wait before their first buffer until the gate is opened again (open != 0).
 * signal_source_start (signal_source.c:57-70) spawns the reader before its caller has registered any
 * callback (cbb_main.c:85-88); a replay that must not lose its first buffers closes the gate around that. */
int b200_replay_gate(int dev_index, int open);
/* Bytes handed to callbacks so far by readers of `dev_index` (all readers, since the last set_capture). */
int64_t b200_replay_delivered_bytes(int dev_index);

#ifdef __cplusplus
}
#endif
#endif
