// fm_kernels.cuh -- device pieces of the FM branch shared by the standalone FM kernel, the
// compat kernels and the fused chain kernel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// resample.c:4 -- the 11-tap half-band kernel; odd taps other than the centre are zero
#define B200_HB0 0.01824f
#define B200_HB2 (-0.11614f)
#define B200_HB4 0.34790f
#define B200_HB5 0.5f

// common_sp.h:40-76 on the CIC output (integers |v| <= 128*R).  The branch |y/x| < 1 is taken
// on the integers (|y| < |x|), which is what the correctly rounded float quotient decides for
// these magnitudes, so the quotient itself may carry a couple of ulps without ever flipping
// the (discontinuous, 0.0083 rad) branch.
__device__ __forceinline__ float atan2_approx_dev(int yi, int xi)
{
    const float pi = 3.14159265358979323846f;
    const float pi_by_2 = 1.57079632679489661923f;
    if (xi == 0) return yi > 0 ? pi_by_2 : (yi == 0 ? 0.0f : -pi_by_2);
    const float y = (float) yi;
    const float x = (float) xi;
    const float z = __fdiv_rn(y, x);
    const int ay = yi < 0 ? -yi : yi;
    const int ax = xi < 0 ? -xi : xi;
    if (ay < ax) {
        const float a = __fdiv_rn(z, fmaf(0.28f * z, z, 1.0f));
        if (xi < 0) return yi < 0 ? a - pi : a + pi;
        return a;
    }
    const float a = pi_by_2 - __fdiv_rn(z, fmaf(z, z, 0.28f));
    return yi < 0 ? a - pi : a;
}

// audio_main.c:117-130: first difference (no unwrap) then the +-1 hard limiter
__device__ __forceinline__ float fm_limit(float cur, float prev)
{
    const float d = cur - prev;
    return fminf(fmaxf(d, -1.0f), 1.0f);
}

// resample.c:53-64 with x(k) = input[2n - k]: centre tap first, then k = 0, 2, ..., 10
__device__ __forceinline__ float halfband_taps(float x0, float x2, float x4, float x5, float x6, float x8,
                                               float x10)
{
    float acc = B200_HB5 * x5;
    acc = fmaf(B200_HB0, x0, acc);
    acc = fmaf(B200_HB2, x2, acc);
    acc = fmaf(B200_HB4, x4, acc);
    acc = fmaf(B200_HB4, x6, acc);
    acc = fmaf(B200_HB2, x8, acc);
    acc = fmaf(B200_HB0, x10, acc);
    return acc;
}

struct FmParams {
    const uint8_t* iq;            // stream 0, first sample of the BATCH (history lies before it)
    int64_t stream_stride_bytes;
    int n_streams;
    int64_t n_samples;            // per stream, multiple of 4*R and of 8
    int R;
    float* audio;                 // [n_streams][n_samples / (4R)]
    int64_t audio_stride;         // floats
    int32_t* decimated;           // nullable, [n_streams][n_samples / R] (re, im)
    int64_t dec_stride;           // complex samples
};

}  // namespace b200
