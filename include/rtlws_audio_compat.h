/*
 * rtlws_audio_compat.h -- the audio interface of sjappig/rtl-ws as exported by libb200audio.so.
 *
 * Declares, with identical names, argument order, types and return conventions, what the
 * reference declares in src/audio_main.h:6-14.  libb200audio.so links in place of audio_main.o:
 * an unmodified main.c then registers a GPU demodulator at main.c:205 and drains it through the
 * same two calls at main.c:86-110.  The arithmetic of audio_main.c:110-139 runs
 * in libb200sdr.so (b200_fm_demod_block, csrc/fm_cs32.cu); this library is host plumbing only.
 * If the reference's audio_main.h is on the include path first, its guard makes the block drop out.
 */
#ifndef RTLWS_AUDIO_COMPAT_H
#define RTLWS_AUDIO_COMPAT_H

#include "rtlws_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef AUDIO_MAIN_H
#define AUDIO_MAIN_H
void audio_init();
int audio_new_audio_available();
/* copies up to buf_len bytes of float32 audio; returns the bytes copied.  Keeps the reference's drain
 * order at pool-buffer boundaries (audio_main.c:49-63), see csrc/audio_compat.cpp. */
int audio_get_audio_payload(char* buf, int buf_len);
/* an rf_decimator_callback: len decimated samples in, len / 4 audio samples into the pool */
void audio_fm_demodulator(const cmplx_s32* signal, int len);
void audio_close();
#endif /* AUDIO_MAIN_H */

/* ---- extensions ---- */
/* Oldest finished pool buffer as a whole (no drain quirk); returns its length in floats, -1 if none
 * is ready or it does not fit. */
int b200_audio_take_buffer(float* out, int max_floats);
int b200_audio_buffer_len(void);
/* blocks dropped because all 50 pool buffers were full (audio_main.c:137) */
long long b200_audio_dropped_blocks(void);
/* back to stream start: zero phase and delay lines (the reference's statics cannot be reset), empty pool */
void b200_audio_reset_stream(void);

#ifdef __cplusplus
}
#endif
#endif /* RTLWS_AUDIO_COMPAT_H */
