/*
 * oracle/fftw3_shim/fftw3_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see fftw3.h).
 * Only the forward, power-of-two, out-of-place plan that spectrum.c:42 creates is
 * supported; anything else returns a NULL plan.
 */
#include <stdlib.h>
#include "fftw3.h"
#include "../fft_f64.h"

struct orc_fftw_plan_s
{
    orc_fft_plan* plan;
    fftw_complex* in;
    fftw_complex* out;
};

void* fftw_malloc(size_t n)
{
    void* p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64) != 0)
        return NULL;
    return p;
}

void fftw_free(void* p)
{
    free(p);
}

fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int sign, unsigned flags)
{
    struct orc_fftw_plan_s* p;
    (void) flags;
    if (sign != FFTW_FORWARD || in == out)
        return NULL;
    p = (struct orc_fftw_plan_s*) calloc(1, sizeof(*p));
    p->plan = orc_fft_plan_create(n);
    if (p->plan == NULL)
    {
        free(p);
        return NULL;
    }
    p->in = in;
    p->out = out;
    return p;
}

void fftw_execute(const fftw_plan p)
{
    orc_fft_execute(p->plan, (const double*) p->in, (double*) p->out);
}

void fftw_destroy_plan(fftw_plan p)
{
    if (p == NULL)
        return;
    orc_fft_plan_destroy(p->plan);
    free(p);
}
