/*
 * oracle/fft_f64.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see fft_f64.h).
 *
 * Stockham autosort FFT, radix 4 with one radix-2 clean-up stage when log2(N) is odd.
 * Out-of-place ping-pong between two work arrays, twiddles tabulated once per plan
 * with cosl/sinl.  Plain C, no SIMD intrinsics: the compiler vectorises the inner
 * q-loops of the later stages.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fft_f64.h"

typedef struct { double re, im; } cplx;

struct orc_fft_plan
{
    int N;
    cplx* w;      /* w[k] = exp(-2*pi*i*k/N), k = 0..N-1 */
    cplx* work0;
    cplx* work1;
};

orc_fft_plan* orc_fft_plan_create(int N)
{
    orc_fft_plan* p;
    int k;
    if (N < 1 || (N & (N - 1)) != 0)
        return NULL;
    p = (orc_fft_plan*) calloc(1, sizeof(*p));
    p->N = N;
    p->w = (cplx*) malloc(sizeof(cplx) * (size_t) N);
    p->work0 = (cplx*) malloc(sizeof(cplx) * (size_t) N);
    p->work1 = (cplx*) malloc(sizeof(cplx) * (size_t) N);
    for (k = 0; k < N; k++)
    {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double) k / (long double) N;
        p->w[k].re = (double) cosl(a);
        p->w[k].im = (double) sinl(a);
    }
    return p;
}

void orc_fft_plan_destroy(orc_fft_plan* p)
{
    if (p == NULL)
        return;
    free(p->w);
    free(p->work0);
    free(p->work1);
    free(p);
}

int orc_fft_plan_size(const orc_fft_plan* p)
{
    return p->N;
}

/* One radix-4 Stockham stage: n = current sub-transform length, s = stride. */
static void stage4(const orc_fft_plan* pl, int n, int s, const cplx* restrict x, cplx* restrict y)
{
    const int n1 = n / 4;
    const int n2 = n / 2;
    const int n3 = n1 + n2;
    const int tstep = pl->N / n;
    int p, q;
    for (p = 0; p < n1; p++)
    {
        const cplx w1 = pl->w[p * tstep];
        const cplx w2 = pl->w[2 * p * tstep];
        const cplx w3 = pl->w[3 * p * tstep];
        const cplx* xa = x + (size_t) s * (p);
        const cplx* xb = x + (size_t) s * (p + n1);
        const cplx* xc = x + (size_t) s * (p + n2);
        const cplx* xd = x + (size_t) s * (p + n3);
        cplx* y0 = y + (size_t) s * (4 * p + 0);
        cplx* y1 = y + (size_t) s * (4 * p + 1);
        cplx* y2 = y + (size_t) s * (4 * p + 2);
        cplx* y3 = y + (size_t) s * (4 * p + 3);
        for (q = 0; q < s; q++)
        {
            const double apc_r = xa[q].re + xc[q].re, apc_i = xa[q].im + xc[q].im;
            const double amc_r = xa[q].re - xc[q].re, amc_i = xa[q].im - xc[q].im;
            const double bpd_r = xb[q].re + xd[q].re, bpd_i = xb[q].im + xd[q].im;
            /* j*(b-d) */
            const double jbmd_r = -(xb[q].im - xd[q].im), jbmd_i = (xb[q].re - xd[q].re);
            const double t1_r = amc_r - jbmd_r, t1_i = amc_i - jbmd_i;
            const double t2_r = apc_r - bpd_r, t2_i = apc_i - bpd_i;
            const double t3_r = amc_r + jbmd_r, t3_i = amc_i + jbmd_i;
            y0[q].re = apc_r + bpd_r;
            y0[q].im = apc_i + bpd_i;
            y1[q].re = t1_r * w1.re - t1_i * w1.im;
            y1[q].im = t1_r * w1.im + t1_i * w1.re;
            y2[q].re = t2_r * w2.re - t2_i * w2.im;
            y2[q].im = t2_r * w2.im + t2_i * w2.re;
            y3[q].re = t3_r * w3.re - t3_i * w3.im;
            y3[q].im = t3_r * w3.im + t3_i * w3.re;
        }
    }
}

/* Final radix-2 stage (n == 2): twiddle is 1. */
static void stage2_last(int s, const cplx* restrict x, cplx* restrict y)
{
    int q;
    for (q = 0; q < s; q++)
    {
        const cplx a = x[q];
        const cplx b = x[q + s];
        y[q].re = a.re + b.re;
        y[q].im = a.im + b.im;
        y[q + s].re = a.re - b.re;
        y[q + s].im = a.im - b.im;
    }
}

void orc_fft_execute(orc_fft_plan* p, const double* in, double* out)
{
    int n = p->N;
    int s = 1;
    const cplx* src = (const cplx*) in;
    cplx* bufs[2];
    int which = 0;
    bufs[0] = p->work0;
    bufs[1] = p->work1;

    if (n == 1)
    {
        out[0] = in[0];
        out[1] = in[1];
        return;
    }
    while (n >= 4)
    {
        stage4(p, n, s, src, bufs[which]);
        src = bufs[which];
        which ^= 1;
        n /= 4;
        s *= 4;
    }
    if (n == 2)
    {
        stage2_last(s, src, bufs[which]);
        src = bufs[which];
    }
    memcpy(out, src, sizeof(cplx) * (size_t) p->N);
}

void orc_dft_naive(int N, const double* in, double* out)
{
    int k, j;
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    for (k = 0; k < N; k++)
    {
        long double sr = 0, si = 0;
        for (j = 0; j < N; j++)
        {
            /* reduce j*k mod N before forming the angle to keep it exact */
            long long jk = ((long long) j * (long long) k) % (long long) N;
            long double a = -two_pi * (long double) jk / (long double) N;
            long double c = cosl(a), sn = sinl(a);
            sr += (long double) in[2 * j] * c - (long double) in[2 * j + 1] * sn;
            si += (long double) in[2 * j] * sn + (long double) in[2 * j + 1] * c;
        }
        out[2 * k] = (double) sr;
        out[2 * k + 1] = (double) si;
    }
}
