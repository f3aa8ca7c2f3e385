// fft_regs.cuh -- fully unrolled in-register forward FFTs (radix-2 DIF, sizes 2..32).
//
// One thread owns R complex points in registers.  All twiddles inside a size-R transform
// are compile-time constants (multiples of 2*pi/32), so after unrolling every butterfly
// is FADD/FMUL/FFMA with immediate operands.  Output is left in bit-reversed register
// order; callers permute by renaming registers (bitrev<R>() is constexpr), which is free.
//
// Sign convention: X[k] = sum_j x[j] exp(-2*pi*i*j*k/R)  -- the FFTW_FORWARD transform the
// reference plans at spectrum.c:42.
#pragma once

#include <cuda_runtime.h>

namespace b200 {

// cos(2*pi*q/32), q = 0..8 (first octant + quadrant end)
__host__ __device__ constexpr float cos32_table(int q)
{
    return q == 0   ? 1.0f
           : q == 1 ? 0.98078528040323044913f
           : q == 2 ? 0.92387953251128675613f
           : q == 3 ? 0.83146961230254523708f
           : q == 4 ? 0.70710678118654752440f
           : q == 5 ? 0.55557023301960222474f
           : q == 6 ? 0.38268343236508977173f
           : q == 7 ? 0.19509032201612826785f
                    : 0.0f;
}

// cos / sin of 2*pi*q/32 for any integer q
__host__ __device__ constexpr float cos32(int q)
{
    q &= 31;
    if (q > 16) q = 32 - q;                 // cos is even about pi
    return q <= 8 ? cos32_table(q) : -cos32_table(16 - q);
}
__host__ __device__ constexpr float sin32(int q)
{
    return cos32(q - 8);
}

template <int R>
__host__ __device__ constexpr int bitrev(int i)
{
    int r = 0;
    for (int b = 1; b < R; b <<= 1) {
        r = (r << 1) | (i & 1);
        i >>= 1;
    }
    return r;
}

// t * exp(-2*pi*i*q/32), 0 <= q < 16; q is a compile-time constant after unrolling
__device__ __forceinline__ float2 mul_w32(float2 t, int q)
{
    const float h = 0.70710678118654752440f;
    if (q == 0) return t;
    if (q == 8) return make_float2(t.y, -t.x);
    if (q == 4) return make_float2((t.x + t.y) * h, (t.y - t.x) * h);
    if (q == 12) return make_float2((t.y - t.x) * h, -(t.x + t.y) * h);
    const float c = cos32(q);
    const float s = sin32(q);
    return make_float2(fmaf(t.y, s, t.x * c), fmaf(-t.x, s, t.y * c));
}

// general complex multiply a * w
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    return make_float2(fmaf(-a.y, w.y, a.x * w.x), fmaf(a.x, w.y, a.y * w.x));
}

// In-place forward DIF FFT of R points held in registers; result index bitrev<R>(p) is in a[p].
template <int R>
__device__ __forceinline__ void fft_dif(float2 (&a)[R])
{
    static_assert(R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "register FFT sizes");
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int g = 0; g < R; g += 2 * half) {
#pragma unroll
            for (int k = 0; k < half; ++k) {
                const int i = g + k;
                const int j = i + half;
                const float2 u = a[i];
                const float2 v = a[j];
                a[i] = make_float2(u.x + v.x, u.y + v.y);
                a[j] = mul_w32(make_float2(u.x - v.x, u.y - v.y), k * (16 / half));
            }
        }
    }
}

}  // namespace b200
