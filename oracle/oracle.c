/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
 *
 * Plain-C restatement of the rtl-ws IQ hot path.  Compile WITHOUT -ffast-math so the
 * float operation order below is the one executed (the reference itself is built with
 * -O3 -ffast-math, Makefile:4; the difference is bounded by a few f32 ulps and is
 * measured in tests/test_oracle_vs_ref.py).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "oracle.h"
#include "fft_f64.h"

/* =============================== spectrum.c ==================================== */

struct orc_spectrum
{
    int N;
    orc_fft_plan* plan;       /* stands for the fftw_plan of spectrum.c:42 */
    double* in;               /* N (re, im) pairs; spectrum.c:40 */
    double* out;              /* spectrum.c:41 */
    double* window;           /* extension; NULL = rectangular (the reference) */
};

orc_spectrum* orc_spectrum_alloc(int N)
{
    orc_spectrum* s = (orc_spectrum*) calloc(1, sizeof(*s));
    s->plan = orc_fft_plan_create(N);
    if (s->plan == NULL)
    {
        free(s);
        return NULL;
    }
    s->N = N;
    s->in = (double*) malloc(sizeof(double) * 2 * (size_t) N);
    s->out = (double*) malloc(sizeof(double) * 2 * (size_t) N);
    return s;
}

void orc_spectrum_free(orc_spectrum* s)
{
    if (s == NULL)
        return;
    orc_fft_plan_destroy(s->plan);
    free(s->in);
    free(s->out);
    free(s->window);
    free(s);
}

void orc_spectrum_set_window(orc_spectrum* s, const double* window)
{
    free(s->window);
    s->window = NULL;
    if (window != NULL)
    {
        s->window = (double*) malloc(sizeof(double) * (size_t) s->N);
        memcpy(s->window, window, sizeof(double) * (size_t) s->N);
    }
}

/* spectrum.c:15-35: transform, then walk the OUTPUT positions i = 0..len-1 reading bin
 * (N/2 + i) mod len -- i.e. an fftshift -- accumulating |X|^2 into the caller's array.
 * The one position that maps to bin 0 (i = N/2) does not get |X[0]|^2: it gets the value
 * just accumulated into its left neighbour (spectrum.c:30-33). */
static void accumulate_power(orc_spectrum* s, double* ps, int len)
{
    const int half = s->N / 2;
    int i;
    if (s->window != NULL)
    {
        for (i = 0; i < s->N; i++)
        {
            s->in[2 * i] *= s->window[i];
            s->in[2 * i + 1] *= s->window[i];
        }
    }
    orc_fft_execute(s->plan, s->in, s->out);
    for (i = 0; i < len; i++)
    {
        const int bin = (half + i) % len;
        if (bin > 0)
        {
            const double re = s->out[2 * bin];
            const double im = s->out[2 * bin + 1];
            ps[i] += (re * re + im * im);
        }
        else
        {
            ps[i] += ps[i - 1];
        }
    }
}

int orc_spectrum_add_cmplx_u8(orc_spectrum* s, const uint8_t* iq, double* ps, int len)
{
    int i;
    if (len != s->N)                      /* spectrum.c:51-52 */
        return -1;
    for (i = 0; i < s->N; i++)            /* spectrum.c:54-58 */
    {
        s->in[2 * i] = (((double) iq[2 * i]) - 128) / 128;
        s->in[2 * i + 1] = (((double) iq[2 * i + 1]) - 128) / 128;
    }
    accumulate_power(s, ps, len);
    return 0;
}

int orc_spectrum_add_cmplx_s32(orc_spectrum* s, const int32_t* iq, double* ps, int len)
{
    int i;
    if (len != s->N)                      /* spectrum.c:69-70 */
        return -1;
    for (i = 0; i < s->N; i++)            /* spectrum.c:72-76 */
    {
        s->in[2 * i] = ((double) iq[2 * i]) / 128;
        s->in[2 * i + 1] = ((double) iq[2 * i + 1]) / 128;
    }
    accumulate_power(s, ps, len);
    return 0;
}

int orc_spectrum_add_real_f32(orc_spectrum* s, const float* x, double* ps, int len)
{
    int i;
    if (len != s->N)                      /* spectrum.c:87-88 */
        return -1;
    for (i = 0; i < s->N; i++)            /* spectrum.c:90-94 */
    {
        s->in[2 * i] = x[i];
        s->in[2 * i + 1] = 0;
    }
    accumulate_power(s, ps, len);
    return 0;
}

int orc_spectrum_rows_cmplx_u8(orc_spectrum* s, const uint8_t* iq, int64_t n_samples,
                               int hop, int K, int64_t row_hop, double* rows, int64_t n_rows)
{
    int64_t r;
    int j;
    const int N = s->N;
    for (r = 0; r < n_rows; r++)
    {
        double* ps = rows + r * N;
        memset(ps, 0, sizeof(double) * (size_t) N);            /* cbb_main.c:50 */
        for (j = 0; j < K; j++)                                  /* cbb_main.c:52-59 */
        {
            const int64_t start = r * row_hop + (int64_t) j * hop;
            if (start + N > n_samples)
                return -1;
            if (orc_spectrum_add_cmplx_u8(s, iq + 2 * start, ps, N))
                return -1;
        }
    }
    return 0;
}

void orc_db_payload(const double* ps, int n, int count, int gain_db, uint8_t* out, double* db_float)
{
    /* cbb_main.c:112 -- gain_db/10 is an INTEGER division before pow() */
    const double linear_energy_gain = pow(10, gain_db / 10);
    int idx;
    for (idx = 0; idx < n; idx++)
    {
        /* cbb_main.c:125-128 */
        const double v = 10 * log10(fabs(linear_energy_gain * ps[idx] / count));
        int m;
        if (db_float != NULL)
            db_float[idx] = v;
        /* (int) of -inf / NaN / out-of-range is undefined in C; x86 yields INT_MIN, which the
         * clamp turns into 0.  Made explicit so the oracle is portable. */
        if (!(v > -2147483648.0))
            m = 0;
        else if (v >= 2147483647.0)
            m = 255;
        else
            m = (int) v;
        m = m >= 0 ? m : 0;
        m = m <= 255 ? m : 255;
        if (out != NULL)
            out[idx] = (uint8_t) m;
    }
}

/* =============================== resample.c ==================================== */

int orc_cic_decimate(int R, const uint8_t* src, int src_len, int32_t* dst, int dst_len,
                     orc_cic_state* delay)
{
    /* int32 wrap-around is part of the contract; do it in uint32 to stay defined */
    uint32_t integ_re = (uint32_t) delay->integrator_prev_out[0];
    uint32_t integ_im = (uint32_t) delay->integrator_prev_out[1];
    uint32_t comb_re = (uint32_t) delay->comb_prev_in[0];
    uint32_t comb_im = (uint32_t) delay->comb_prev_in[1];
    int n, m = 0;

    if (dst_len * R != src_len)           /* resample.c:18-19 */
        return -1;

    for (n = 0; n < src_len; n++)
    {
        /* integrator y(n) = y(n-1) + (x(n) - 128)   resample.c:23-25 */
        integ_re += (uint32_t) ((int32_t) src[2 * n] - 128);
        integ_im += (uint32_t) ((int32_t) src[2 * n + 1] - 128);
        if (((n + 1) % R) == 0)           /* resample.c:28 */
        {
            if (m >= dst_len)             /* resample.c:31-34 */
                return -2;
            /* comb y(m) = x(m) - x(m-1)   resample.c:35-36 */
            dst[2 * m] = (int32_t) (integ_re - comb_re);
            dst[2 * m + 1] = (int32_t) (integ_im - comb_im);
            comb_re = integ_re;
            comb_im = integ_im;
            m++;
        }
    }
    delay->integrator_prev_out[0] = (int32_t) integ_re;   /* resample.c:42-43 */
    delay->integrator_prev_out[1] = (int32_t) integ_im;
    delay->comb_prev_in[0] = (int32_t) comb_re;
    delay->comb_prev_in[1] = (int32_t) comb_im;
    return 0;
}

/* resample.c:4 */
static const float orc_half_band_kernel[ORC_HALF_BAND_N] = {
    0.01824f, 0.0f, -0.11614f, 0.0f, 0.34790f, 0.5f, 0.34790f, 0.0f, -0.11614f, 0.0f, 0.01824f
};

static inline float hb_tap_input(const float* input, const float* delay, int idx)
{
    /* resample.c:56,62: negative indices read the tail of the previous call's input */
    return idx >= 0 ? input[idx] : delay[(ORC_HALF_BAND_N - 1) + idx];
}

void orc_halfband_decimate(const float* input, float* output, int output_len, float* delay)
{
    int n, k;
    for (n = 0; n < output_len; n++)
    {
        /* centre tap first (resample.c:55-56), then the even taps in ascending k (:59-63) */
        float acc = orc_half_band_kernel[ORC_HALF_BAND_N / 2] *
                    hb_tap_input(input, delay, 2 * n - ORC_HALF_BAND_N / 2);
        for (k = 0; k < ORC_HALF_BAND_N; k += 2)
            acc += orc_half_band_kernel[k] * hb_tap_input(input, delay, 2 * n - k);
        output[n] = acc;
    }
    /* resample.c:66 */
    memcpy(delay, &input[2 * output_len - (ORC_HALF_BAND_N - 1)], (ORC_HALF_BAND_N - 1) * sizeof(float));
}

/* =============================== common_sp.h =================================== */

float orc_atan2_approx(float y, float x)
{
    const float pi_by_2 = (float) (M_PI / 2);          /* common_sp.h:43 */
    float atan_v;
    float z;

    if (x == 0)                                         /* common_sp.h:47-56 */
    {
        if (y > 0.0f)
            return pi_by_2;
        if (y == 0)
            return 0;
        return -pi_by_2;
    }
    z = y / x;                                          /* common_sp.h:57 */
    if (fabs(z) < 1.0f)                                 /* common_sp.h:58 */
    {
        atan_v = z / (1.0f + 0.28f * z * z);            /* common_sp.h:60 */
        if (x < 0)
        {
            /* common_sp.h:63-66: M_PI is a double, so the add happens in double and is
             * rounded to float on return */
            if (y < 0.0f)
                return (float) (atan_v - M_PI);
            return (float) (atan_v + M_PI);
        }
    }
    else
    {
        atan_v = pi_by_2 - z / (z * z + 0.28f);         /* common_sp.h:71 */
        if (y < 0.0f)
            return (float) (atan_v - M_PI);             /* common_sp.h:72-73 */
    }
    return atan_v;
}

/* =============================== audio_main.c ================================== */

void orc_fm_demodulate(const int32_t* signal, int len, orc_fm_state* st,
                       float* demod, float* work, float* audio)
{
    const float scale = 1;                              /* audio_main.c:76 */
    int i;
    for (i = 0; i < len; i++)                           /* audio_main.c:110-131 */
    {
        float temp;
        demod[i] = orc_atan2_approx((float) signal[2 * i + 1], (float) signal[2 * i]);
        temp = demod[i];
        demod[i] -= st->prev_sample;                    /* first difference, no unwrap */
        st->prev_sample = temp;
        if (demod[i] > scale)                           /* hard limiter */
            demod[i] = 1;
        else if (demod[i] < -scale)
            demod[i] = -1;
        else
            demod[i] /= scale;
    }
    orc_halfband_decimate(demod, work, len / 2, st->delay_line_1);      /* audio_main.c:133 */
    /* audio_main.c:137-139: the reference runs stage 2 only while its 50-buffer pool has a
     * free buffer; the oracle always has room (the harness drains every block). */
    orc_halfband_decimate(work, audio, len / 4, st->delay_line_2);
}

/* =============================== rf_decimator.c ================================ */

struct orc_chain
{
    double sample_rate;
    int down_factor;
    uint8_t* input_signal;        /* interleaved u8 IQ, input_signal_len samples */
    int input_signal_len;
    int surplus;
    int32_t* resampled_signal;
    int resampled_signal_len;
    orc_cic_state delay;
    orc_fm_state fm;
    float* demod;
    float* work;
    float* audio;
};

orc_chain* orc_chain_create(double sample_rate, int down_factor)
{
    orc_chain* c;
    if (!(sample_rate > 0 && down_factor > 0))          /* rf_decimator.c:58 */
        return NULL;
    c = (orc_chain*) calloc(1, sizeof(*c));
    c->sample_rate = sample_rate;
    c->down_factor = down_factor;
    /* rf_decimator.c:65-66, INTERNAL_BUF_LEN_MS == 100 */
    c->resampled_signal_len = (int) ((sample_rate / down_factor) * 100 / 1000);
    c->input_signal_len = c->resampled_signal_len * down_factor;
    c->input_signal = (uint8_t*) malloc(2 * (size_t) c->input_signal_len + 2);
    c->resampled_signal = (int32_t*) malloc(2 * sizeof(int32_t) * (size_t) c->resampled_signal_len + 8);
    c->demod = (float*) malloc(sizeof(float) * (size_t) c->resampled_signal_len + 4);
    c->work = (float*) malloc(sizeof(float) * (size_t) (c->resampled_signal_len / 2) + 4);
    c->audio = (float*) malloc(sizeof(float) * (size_t) (c->resampled_signal_len / 4) + 4);
    return c;
}

void orc_chain_free(orc_chain* c)
{
    if (c == NULL)
        return;
    free(c->input_signal);
    free(c->resampled_signal);
    free(c->demod);
    free(c->work);
    free(c->audio);
    free(c);
}

int orc_chain_block_in(const orc_chain* c)
{
    return c->input_signal_len;
}

int orc_chain_block_out(const orc_chain* c)
{
    return c->resampled_signal_len;
}

int orc_chain_push(orc_chain* c, const uint8_t* iq, int len,
                   int32_t* dec, int64_t* n_dec, int64_t dec_cap,
                   float* audio, int64_t* n_audio, int64_t audio_cap)
{
    int current = 0;
    int remaining = len;
    int block_size = c->input_signal_len - c->surplus;   /* rf_decimator.c:88 */

    while (remaining >= block_size)                      /* rf_decimator.c:93 */
    {
        const int n_out = c->resampled_signal_len;
        memcpy(c->input_signal + 2 * (size_t) c->surplus, iq + 2 * (size_t) current, 2 * (size_t) block_size);
        remaining -= block_size;
        current += block_size;

        if (orc_cic_decimate(c->down_factor, c->input_signal, c->input_signal_len,
                             c->resampled_signal, n_out, &c->delay))      /* rf_decimator.c:99 */
            return -2;

        if (dec != NULL)
        {
            if (*n_dec + n_out > dec_cap)
                return -3;
            memcpy(dec + 2 * (*n_dec), c->resampled_signal, 2 * sizeof(int32_t) * (size_t) n_out);
            *n_dec += n_out;
        }
        /* rf_decimator.c:105 -> callback audio_fm_demodulator (main.c:205) */
        orc_fm_demodulate(c->resampled_signal, n_out, &c->fm, c->demod, c->work, c->audio);
        if (audio != NULL)
        {
            if (*n_audio + n_out / 4 > audio_cap)
                return -3;
            memcpy(audio + *n_audio, c->audio, sizeof(float) * (size_t) (n_out / 4));
            *n_audio += n_out / 4;
        }

        c->surplus = 0;                                  /* rf_decimator.c:107-108 */
        block_size = c->input_signal_len;
    }
    if (remaining > 0)                                   /* rf_decimator.c:111-115 */
    {
        memcpy(c->input_signal + 2 * (size_t) c->surplus, iq + 2 * (size_t) (len - remaining), 2 * (size_t) remaining);
        c->surplus += remaining;
    }
    return 0;
}

/* ============ extensions that are NOT in the reference (product: csrc/audio_post.cu) ============
 *
 * The reference's chain ends at fs/(4R) = 51.2 kS/s without de-emphasis (audio_main.c:133-139,
 * rf_decimator.c:65-66); its UI asks for a 48 kHz AudioContext (resources/rtl_ui.js:79-82).
 * BASELINE.json's north star names "de-emphasis and resampler ... 48 kHz audio", so the product
 * offers both as opt-in extensions.  There is no reference code to follow; what follows is the
 * DEFINITION the product's kernels are checked against, written as the plain sequential loops. */

/* single-pole de-emphasis: y[n] = y[n-1] + alpha * (x[n] - y[n-1]); *state = y[-1] in, y[n-1] out */
void orc_deemphasis(const float* x, int n, float alpha, float* state, float* y)
{
    float prev = *state;
    int i;
    for (i = 0; i < n; ++i)
    {
        prev = prev + alpha * (x[i] - prev);
        y[i] = prev;
    }
    *state = prev;
}

/* alpha = 1 - exp(-1 / (rate * tau)), evaluated in double and rounded once */
float orc_deemphasis_alpha(double rate_hz, double tau_s)
{
    return (float) (1.0 - exp(-1.0 / (rate_hz * tau_s)));
}

/* the 15/16 resampler's prototype: 240-tap Blackman-windowed sinc, cutoff 1/32 cycles per
 * interpolated sample (24 kHz at 15 * 51.2 kHz), scaled to a DC gain of 15 */
void orc_resample_taps(float* h)
{
    const double pi = 3.14159265358979323846;
    const double fc = 1.0 / 32.0;
    const double mid = (ORC_RESAMPLE_TAPS - 1) / 2.0;
    double w[ORC_RESAMPLE_TAPS];
    double sum = 0.0;
    int k;
    for (k = 0; k < ORC_RESAMPLE_TAPS; ++k)
    {
        const double t = (double) k - mid;
        const double a = 2.0 * pi * fc * t;
        const double sinc = t == 0.0 ? 1.0 : sin(a) / a;
        const double win = 0.42 - 0.5 * cos(2.0 * pi * (double) k / (ORC_RESAMPLE_TAPS - 1))
                           + 0.08 * cos(4.0 * pi * (double) k / (ORC_RESAMPLE_TAPS - 1));
        w[k] = sinc * win;
        sum += w[k];
    }
    for (k = 0; k < ORC_RESAMPLE_TAPS; ++k)
        h[k] = (float) (w[k] * 15.0 / sum);
}

/* y[m] = sum_{t=0}^{15} h[r + 15 t] * x[q - t], q = floor(16 m / 15), r = 16 m mod 15; x[-15..-1]
 * comes from hist[0..14] (zero at stream start), which is then refreshed with the last 15 inputs.
 * n must be a multiple of 16; writes n / 16 * 15 outputs. */
int orc_resample_15_16(const float* x, int n, float* hist, float* y)
{
    float h[ORC_RESAMPLE_TAPS];
    float ext[15];
    int m, t, n_out;
    if (n % 16 != 0)
        return -1;
    orc_resample_taps(h);
    n_out = n / 16 * 15;
    for (m = 0; m < n_out; ++m)
    {
        const int q = (16 * m) / 15;
        const int r = 16 * m - 15 * q;
        float acc = 0.0f;
        for (t = 0; t < 16; ++t)
        {
            const int idx = q - t;
            const float v = idx >= 0 ? x[idx] : hist[15 + idx];
            acc += h[r + 15 * t] * v;
        }
        y[m] = acc;
    }
    for (t = 0; t < 15; ++t)
    {
        const int idx = n - 15 + t;
        ext[t] = idx >= 0 ? x[idx] : hist[15 + idx];
    }
    memcpy(hist, ext, sizeof(ext));
    return n_out;
}
