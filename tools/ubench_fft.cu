// tools/ubench_fft.cu -- the per-warp 1024-point transform in isolation (frame bytes resident in
// shared memory, no global traffic): cycles per frame per scheduler at 1..4 warps per SMSP.
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
#include <math.h>
#include "../rtl-ws_b200/csrc/fft1024_warp.cuh"

using namespace b200;

template <int PART>
__global__ void __launch_bounds__(512, 1) k(const float2* twiddle, float* out, int iters, unsigned long long* cycles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* frame = smem + warp * (2048 + FFT1024_XCH_BYTES);
    float2* xch = reinterpret_cast<float2*>(frame + 2048);
    for (int i = lane; i < 512; i += 32) reinterpret_cast<uint32_t*>(frame)[i] = 0x80808080u ^ (i * 2654435761u + warp);
    __syncwarp();
    float2 tw[32];
    fft1024_load_twiddles(twiddle, lane, tw);
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        c64 a[32];
        fft1024_load<false>(reinterpret_cast<const uint16_t*>(frame), nullptr, lane, a);
        float pw[32];
        if (PART == 0) {            // load only
#pragma unroll
            for (int q = 0; q < 32; ++q) { float re, im; cunpack(a[q], re, im); pw[q] = re + im; }
        } else {
            fft1024_core<true>(a, tw, xch, lane, pw);
        }
        if (PART >= 2) {            // + dB epilogue
#pragma unroll
            for (int q = 0; q < 32; ++q) pw[q] = fmaf(DB_PER_LOG2, lg2_ftz(pw[q]), 1.0f);
        }
        if (PART == 3) {            // + 128-byte coalesced streaming stores (4 KB per frame, small L2-resident target)
            float* dst = out + ((blockIdx.x * 16 + warp) * 4 + (it & 3)) * 1024 + lane;
#pragma unroll
            for (int q = 0; q < 32; ++q) __stcs(dst + fft1024_col(q), pw[q]);
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) acc += pw[q];
        if (acc == 12345.678f) reinterpret_cast<uint32_t*>(frame)[lane] ^= 1;   // keep the loop honest
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long) (t1 - t0);
}

template <int PART>
void run(const char* name, const float2* d_tw, int warps)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    unsigned long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 16 * 4 * 1024 + 4096);
    cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    const int smem = warps * (2048 + FFT1024_XCH_BYTES);
    cudaFuncSetAttribute(k<PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 200;
    k<PART><<<sms, warps * 32, smem>>>(d_tw, out, iters, cyc);
    k<PART><<<sms, warps * 32, smem>>>(d_tw, out, iters, cyc);
    cudaDeviceSynchronize();
    unsigned long long h[1024];
    cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double) h[i];
    mean /= sms;
    printf("%-22s warps/SM=%2d  cycles/frame/warp=%7.0f  cycles/frame/SMSP=%7.0f  -> %6.0f Gsamples/s @1.965GHz  %s\n", name,
           warps, mean / iters, mean / iters / (warps / 4.0), 1024.0 * warps / (mean / iters) * 148 * 1.965,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    std::vector<float2> tw(1024);
    for (int i = 0; i < 1024; ++i) tw[i] = make_float2((float) cos(-2 * M_PI * i / 1024), (float) sin(-2 * M_PI * i / 1024));
    float2* d_tw;
    cudaMalloc(&d_tw, sizeof(float2) * 1024);
    cudaMemcpy(d_tw, tw.data(), sizeof(float2) * 1024, cudaMemcpyHostToDevice);
    for (int warps : {4, 8, 12, 16}) {
        run<0>("load+unpack only", d_tw, warps);
        run<1>("load+fft+power", d_tw, warps);
        run<2>("load+fft+power+dB", d_tw, warps);
        run<3>("load+fft+power+dB+STG", d_tw, warps);
    }
    return 0;
}
