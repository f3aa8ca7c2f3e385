// tools/ubench2.cu -- can the scheduler issue other work in the second cycle of a packed FFMA2?
// Loop body: 8 independent FFMA2 + M independent integer (ALU-pipe) ops / shared loads per thread.
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 2048

template <int MODE>
__global__ void k(float* out, float a0, unsigned long long* cycles)
{
    __shared__ float sm[1024];
    sm[threadIdx.x & 1023] = a0;
    __syncthreads();
    unsigned long long xp[8];
    unsigned int y[16];
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %1};" : "=l"(xp[i]) : "f"(a0 + i + threadIdx.x));
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = threadIdx.x * 7 + i;
    unsigned long long bp, cp;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bp) : "f"(1.0001f));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cp) : "f"(a0));
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE != 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xp[i]) : "l"(bp), "l"(cp));
            if (MODE == 1 || MODE == 5) {   // +8 integer ops
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[i + 8]), "r"(it));
            } else if (MODE == 2) {         // +16 integer ops
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[i + 8]), "r"(it));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i + 8]) : "r"(y[i]), "r"(it));
            } else if (MODE == 3) {         // +8 shared loads
                acc += sm[(threadIdx.x + i * 32 + it) & 1023];
            } else if (MODE == 4) {         // +8 scalar FADD (same pipe)
                asm volatile("add.f32 %0, %0, %1;" : "+f"(acc) : "f"(a0));
            }
        }
    }
    const long long t1 = clock64();
    float s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(xp[i]));
        s += lo + hi + (float) y[i] + (float) y[i + 8];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long) (t1 - t0);
}

template <int MODE>
void run(const char* name, int threads)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    unsigned long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * threads);
    cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    k<MODE><<<sms, threads>>>(out, 1.0f, cyc);
    k<MODE><<<sms, threads>>>(out, 1.0f, cyc);
    cudaDeviceSynchronize();
    unsigned long long h[1024];
    cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double) h[i];
    mean /= sms;
    const double warps_per_smsp = threads / 32.0 / 4.0;
    printf("%-28s threads/SM=%4d  cycles/iter/warp-on-SMSP = %6.2f   (per SMSP: %6.2f cycles per 8 FFMA2-groups)  %s\n", name, threads,
           mean / ITERS, mean / ITERS / warps_per_smsp, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int threads : {128, 256, 512}) {
        run<0>("8 FFMA2", threads);
        run<5>("8 LOP3 only", threads);
        run<1>("8 FFMA2 + 8 LOP3", threads);
        run<2>("8 FFMA2 + 16 LOP3", threads);
        run<3>("8 FFMA2 + 8 LDS + 8 FADD", threads);
        run<4>("8 FFMA2 + 8 FADD", threads);
    }
    return 0;
}
