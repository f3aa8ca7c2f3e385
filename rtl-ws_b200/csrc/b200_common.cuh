// b200_common.cuh -- shared device helpers (sm_100a): mbarrier + 1-D bulk TMA copies,
// error plumbing.  No CPU fallback anywhere: a failed CUDA call surfaces as B200_ERR_CUDA.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/b200sdr.h"

namespace b200 {

// ---- host-side error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
int sm_count();

#define B200_CUDA_TRY(expr)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            ::b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),  \
                              __FILE__, __LINE__);                                       \
            return B200_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

// Opt a kernel into more than 48 KB of dynamic shared memory, once per (kernel, device): the
// attribute belongs to the device's copy of the function, and a process may drive several GPUs.
int ensure_dynamic_smem(const void* kernel, int bytes);
// Resident CTAs per SM for (kernel, block, smem), cached per device like the attribute.
int cached_occupancy(const void* kernel, int block_threads, int smem_bytes);

#define B200_LAUNCH_CHECK()                                                              \
    do {                                                                                 \
        ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            ::b200::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                              __FILE__, __LINE__);                                       \
            return B200_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

// ---- device: mbarrier + bulk async copy (TMA, 1-D) -------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make mbarrier initialisation visible to the async proxy before the first TMA uses it
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// order this thread's prior generic-proxy smem accesses before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy through the TMA unit (SASS: UBLKCP); completion is signalled
// on `bar` as `bytes` transaction bytes.  dst, src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// drop a 128-byte line from L2 without writing it back (its contents are dead)
__device__ __forceinline__ void l2_discard_128(const void* ptr)
{
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(ptr) : "memory");
}

// log2 on the MUFU unit, flush-to-zero form: no denormal pre-scaling instructions.  Inputs
// here are sums of squares of integers (exact zero or >= 1), never denormal.  lg2(0) = -inf.
__device__ __forceinline__ float lg2_ftz(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#endif  // __CUDACC__

}  // namespace b200
