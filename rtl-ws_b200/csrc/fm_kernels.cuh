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

// ---- CIC boxcar (resample.c:21-40) -----------------------------------------------------------
// Sum of R u8 samples minus 128*R per component, exact.  The accumulator starts at the bit
// pattern of the float 1.5 * 2^23 minus 128*R, so after the integer adds it IS the float
// 1.5 * 2^23 + s: one FADD turns it into float(s) without an int->float conversion.
constexpr uint32_t CIC_MAGIC_BITS = 0x4B400000u;       // 12582912.0f
constexpr float CIC_MAGIC = 12582912.0f;

// R = 10: five aligned 32-bit words = ten samples; dp4a sums bytes 0,2 (re) and 1,3 (im)
__device__ __forceinline__ void cic10_sum(const uint32_t* w, uint32_t& ure, uint32_t& uim)
{
    ure = CIC_MAGIC_BITS - 1280u;
    uim = CIC_MAGIC_BITS - 1280u;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t v = w[k];
        ure = __dp4a(v, 0x00010001u, ure);
        uim = __dp4a(v, 0x01000100u, uim);
    }
}

// even R: R/2 aligned 32-bit words (a decimated sample starts on a word boundary when 2R is a multiple of 4)
template <int R>
__device__ __forceinline__ void cic_even_sum(const uint32_t* w, uint32_t& ure, uint32_t& uim)
{
    static_assert(R % 2 == 0, "word-wise CIC sum needs an even down factor");
    ure = CIC_MAGIC_BITS - 128u * (uint32_t) R;
    uim = CIC_MAGIC_BITS - 128u * (uint32_t) R;
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
        const uint32_t v = w[k];
        ure = __dp4a(v, 0x00010001u, ure);
        uim = __dp4a(v, 0x01000100u, uim);
    }
}

// odd R: R aligned 16-bit loads (a decimated sample starts on a 2-byte boundary only)
template <int R>
__device__ __forceinline__ void cic_odd_sum(const uint16_t* h, uint32_t& ure, uint32_t& uim)
{
    ure = CIC_MAGIC_BITS - 128u * (uint32_t) R;
    uim = CIC_MAGIC_BITS - 128u * (uint32_t) R;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const uint32_t v = h[k];
        ure += v & 0xffu;
        uim += v >> 8;
    }
}

// any R, sample-wise
__device__ __forceinline__ void cic_sum(const uint16_t* h, int R, uint32_t& ure, uint32_t& uim)
{
    ure = CIC_MAGIC_BITS - 128u * (uint32_t) R;
    uim = CIC_MAGIC_BITS - 128u * (uint32_t) R;
    for (int k = 0; k < R; ++k) {
        const uint32_t v = h[k];
        ure += v & 0xffu;
        uim += v >> 8;
    }
}

// ---- atan2_approx (common_sp.h:40-76) on the CIC output -----------------------------------------
// x, y hold integers (|v| <= 128*R).  The reference computes z = y/x and then
//   |z| < 1 : z / (1 + 0.28 z^2)            = x*y / (x^2 + 0.28 y^2)
//   else    : pi/2 - z / (z^2 + 0.28)       = pi/2 - x*y / (y^2 + 0.28 x^2)
// i.e. one quotient num/den with den = big^2 + 0.28 * small^2; the right-hand forms are used
// here (one reciprocal instead of two divisions; x*y, x^2, y^2 are exact for |v| < 4096).
// The branch (discontinuous by 0.0083 rad at |y| == |x|) is decided on the exact magnitudes,
// which is what the correctly rounded quotient decides for integers of this size, so the
// approximate reciprocal (2 ulp) can never flip it.  Quadrant fix-ups as in common_sp.h:61-74;
// x == 0 falls out of the second form (num = 0), x == y == 0 is forced to 0 (common_sp.h:52-53).
// Deviation from the reference sequence: a few f32 ulps (measured in tests: < 1e-6 rad).
__device__ __forceinline__ float atan2_approx_dev(float y, float x)
{
    const float pi = 3.14159265358979323846f;
    const float pi_by_2 = 1.57079632679489661923f;
    const float x2 = x * x;
    const float y2 = y * y;
    const bool lt = fabsf(y) < fabsf(x);
    const float den = fmaf(0.28f, lt ? y2 : x2, lt ? x2 : y2);
    // den is an exact non-negative integer-valued float (0 or >= 1): no denormal handling needed
    float rden;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
    const float q = (x * y) * rden;
    float r = lt ? (q + (x < 0.0f ? copysignf(pi, y) : 0.0f)) : (copysignf(pi_by_2, y) - q);
    return den == 0.0f ? 0.0f : r;
}

// The same function with the quadrant logic folded into one signed offset and one FMA:
//   r = copysign(off, y) + (lt ? q : -q),   off = lt ? (x < 0 ? pi : 0) : pi/2
// (both branches of the form above are "signed offset +- q"); the +-q sign rides on the product
// x * (lt ? y : -y).  Five instructions fewer per sample than atan2_approx_dev; same branch decision,
// same special cases, last-ulp rounding differs (one rounding instead of two after the quotient).
__device__ __forceinline__ float atan2_approx_dev2(float y, float x)
{
    const float pi = 3.14159265358979323846f;
    const float pi_by_2 = 1.57079632679489661923f;
    const float x2 = x * x;
    const float y2 = y * y;
    const bool lt = fabsf(y) < fabsf(x);
    const float den = fmaf(0.28f, lt ? y2 : x2, lt ? x2 : y2);
    float rden;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
    const float num = x * (lt ? y : -y);
    const float off = lt ? (x < 0.0f ? pi : 0.0f) : pi_by_2;
    const float r = fmaf(num, rden, copysignf(off, y));
    return den == 0.0f ? 0.0f : r;
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

// One half-band output from an 8-byte aligned pointer to the sample feeding tap k = 10 (the
// oldest): taps sit at +10, +8, +6, +5, +4, +2, +0 from there.  Five 64-bit loads and one 32-bit
// load instead of seven scalar ones, and lanes that walk the array in steps of two floats read
// contiguous words (no 2-way bank conflict).
__device__ __forceinline__ float halfband_from(const float* oldest)
{
    const float2 a = *reinterpret_cast<const float2*>(oldest);          // k = 10, 9
    const float2 b = *reinterpret_cast<const float2*>(oldest + 2);      // k = 8, 7
    const float2 c = *reinterpret_cast<const float2*>(oldest + 4);      // k = 6, 5
    const float2 d = *reinterpret_cast<const float2*>(oldest + 6);      // k = 4, 3
    const float2 e = *reinterpret_cast<const float2*>(oldest + 8);      // k = 2, 1
    const float f = oldest[10];                                         // k = 0
    return halfband_taps(f, e.x, d.x, c.y, c.x, b.x, a.x);
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
