// tools/ubench.cu -- issue-rate microbenchmarks that calibrate the FFT kernel design on B200:
// scalar FFMA / FADD vs the packed f32x2 forms (sm_100 PTX), per SM per clock.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void k(float* out, float a0, float b0, unsigned long long* cycles)
{
    float x[ILP * 2];
#pragma unroll
    for (int i = 0; i < ILP * 2; ++i) x[i] = a0 + i + threadIdx.x;
    unsigned long long xp[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xp[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
    unsigned long long bp, cp;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bp) : "f"(b0));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cp) : "f"(a0));
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) {          // scalar FFMA x2
                x[2 * i] = fmaf(x[2 * i], b0, a0);
                x[2 * i + 1] = fmaf(x[2 * i + 1], b0, a0);
            } else if (MODE == 1) {   // packed FFMA2
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xp[i]) : "l"(bp), "l"(cp));
            } else if (MODE == 2) {   // scalar FADD x2
                x[2 * i] = x[2 * i] + b0;
                x[2 * i + 1] = x[2 * i + 1] + b0;
            } else if (MODE == 3) {   // packed FADD2
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(xp[i]) : "l"(bp));
            } else if (MODE == 4) {   // packed FMUL2
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(xp[i]) : "l"(bp));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP * 2; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(xp[i]));
        s += lo + hi;
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
    k<MODE><<<sms, threads>>>(out, 1.0f, 1.0001f, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, threads>>>(out, 1.0f, 1.0001f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[1024];
    cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double) h[i];
    mean /= sms;
    // lane-ops: each iteration does ILP*2 float results per thread
    const double lane_ops = (double) ITERS * ILP * 2 * threads;
    printf("%-14s threads/SM=%4d  cycles=%9.0f  float-results/clk/SM=%7.1f  (%.3f ms, err=%s)\n", name, threads, mean,
           lane_ops / mean, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("FFMA scalar", threads);
        run<1>("FFMA2 packed", threads);
        run<2>("FADD scalar", threads);
        run<3>("FADD2 packed", threads);
        run<4>("FMUL2 packed", threads);
    }
    return 0;
}
