// spectrum_mx1024.cu -- batched power spectra for N = M * 1024, M in {2, 4, 8} (sm_100a):
// the frame lengths between the reference's 1024 and what still fits one CTA's shared memory
// (4096 is BASELINE config 2's second size).
//
// Decimation in time by M:  X[k + 1024 q] = sum_{r<M} W_M^(r q) * ( W_N^(r k) * F_r[k] ),
// F_r = the 1024-point transform of the polyphase branch x[M m + r].  A CTA is M warps:
//   1. the frame (2N bytes) arrives by one TMA bulk copy (two-deep ring);
//   2. warp r runs the 32x32 register transform of fft1024_warp.cuh on branch r (stride-M u16
//      reads from shared memory), multiplies by W_N^(r k) and parks Z_r[k] in its own exchange
//      tile (the tile is free again after the transform's transpose);
//   3. block barrier; every warp takes 1024/M values of k, reads Z_0..Z_{M-1}[k], runs the
//      M-point butterfly in registers and gets bins k, k + 1024, ... : |X|^2, K-frame
//      accumulation with the cumulative DC-position patch (spectrum.c:30-33), dB / power / u8.
// Arithmetic per frame as in spectrum1024.cu (spectrum.c:15-58, cbb_main.c:112-128).
#include "b200_common.cuh"
#include "fft1024_warp.cuh"
#include "spectrum_kernels.cuh"

namespace b200 {

namespace {

template <int M>
struct MxCfg {
    static constexpr int N = M * 1024;
    static constexpr int FRAME_BYTES = 2 * N;
    static constexpr int THREADS = M * 32;
    static constexpr int KPW = 32 / M;                       // k-groups of 32 per warp in the combine
    static constexpr int SMEM = 2 * FRAME_BYTES + M * FFT1024_XCH_BYTES + 64;
};

template <int M, bool WINDOW, bool MULTI>
__global__ void __launch_bounds__(MxCfg<M>::THREADS, (12 / M) > 0 ? (12 / M) : 1) spectrum_mx1024_kernel(const SpecParams p)
{
    using C = MxCfg<M>;
    constexpr int N = C::N;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int r = tid >> 5;                                   // polyphase branch of this warp
    uint8_t* ring = smem;
    uint8_t* xch_base = smem + 2 * C::FRAME_BYTES;
    float2* xch = reinterpret_cast<float2*>(xch_base + r * FFT1024_XCH_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch_base + M * FFT1024_XCH_BYTES);
    float* dc_slot = reinterpret_cast<float*>(bars + 2);      // bin N-1's accumulated weight, warp M-1 -> warp 0

    const uint32_t total_items = (uint32_t) p.n_streams * (uint32_t) p.n_rows;
    const uint32_t n_rows = (uint32_t) p.n_rows;
    const int K = MULTI ? p.K : 1;
    if (blockIdx.x >= total_items) return;
    const uint32_t n_items = (total_items - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint32_t n_frames = n_items * (uint32_t) K;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto frame_src = [&](uint32_t f) -> const uint8_t* {
        const uint32_t item = blockIdx.x + (f / (uint32_t) K) * gridDim.x;
        const uint32_t j = f % (uint32_t) K;
        const uint32_t s = item / n_rows;
        const uint32_t row = item - s * n_rows;
        return p.iq + (int64_t) s * p.stream_stride_bytes + 2 * ((int64_t) row * p.row_hop + (int64_t) j * p.hop);
    };
    if (tid == 0) {
        for (uint32_t f = 0; f < 2 && f < n_frames; ++f) {
            mbar_arrive_expect_tx(&bars[f], C::FRAME_BYTES);
            tma_load_1d(ring + f * C::FRAME_BYTES, frame_src(f), C::FRAME_BYTES, &bars[f]);
        }
    }

    float2 tw[32];
    fft1024_load_twiddles(p.twiddle, lane, tw);               // W_1024^(lane*k1) = table[M * ...]: see launcher
    const float dboff = p.db_offset - 16.0f * DB_PER_LOG2;

    float acc[MULTI ? 32 : 1];
    if (MULTI) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
    }
    float dcacc = 0.0f;

    // one output value -> the requested arrays (display order: fftshift of spectrum.c:25)
    auto emit = [&](size_t row_base, int bin, float pw) {
        const int col = (bin + N / 2) & (N - 1);
        const float db = fmaf(DB_PER_LOG2, lg2_ftz(pw), dboff);
        if (p.db) __stcs(p.db + row_base + col, db);
        if (p.power) __stcs(p.power + row_base + col, pw * FFT1024_POWER_SCALE);
        if (p.db_u8) {
            int m = __float2int_rz(db);
            m = m < 0 ? 0 : (m > 255 ? 255 : m);
            p.db_u8[row_base + col] = (uint8_t) m;
        }
    };

    uint32_t f = 0;
    for (uint32_t it = 0; it < n_items; ++it) {
        const uint32_t item = blockIdx.x + it * gridDim.x;
        for (int j = 0; j < K; ++j, ++f) {
            const int st = f & 1;
            mbar_wait(&bars[st], (f >> 1) & 1);
            const uint16_t* in16 = reinterpret_cast<const uint16_t*>(ring + st * C::FRAME_BYTES);

            // ---- branch r: samples M*(32*n1 + lane) + r ----
            c64 a[32];
            {
                const c64 bias1 = cpack(8421376.0f, 8421376.0f);           // 2^23 + 256 * 128
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const int idx = M * (32 * n1 + lane) + r;
                    const uint32_t v = in16[idx];
                    const int q = bitrev<32>(n1);
                    a[q] = cpack(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7504)),
                                 __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7514)));
                    if (WINDOW) {
                        const float w = __ldg(&p.window[idx]);
                        a[q] = cmul2(csub(a[q], bias1), cpack(w, w));
                    }
                }
            }
            __syncthreads();                      // the stage is consumed by all branches
            if (tid == 0 && f + 2 < n_frames) {
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(&bars[st], C::FRAME_BYTES);
                tma_load_1d(ring + st * C::FRAME_BYTES, frame_src(f + 2), C::FRAME_BYTES, &bars[st]);
            }

            c64 b[32];
            fft1024_transform<!WINDOW>(a, tw, xch, lane, b);
            // ---- Z_r[k] = W_N^(r k) F_r[k], k = lane + 32 k2, parked in this warp's tile as [k2][lane] ----
            __syncwarp();                         // every lane has finished reading the transpose
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) {
                c64 z = b[k2];
                if (r > 0) {
                    const float2 w = __ldg(&p.twiddle_n[(r * (lane + 32 * k2)) & (N - 1)]);
                    z = cmul(z, w.x, w.y);
                }
                reinterpret_cast<c64*>(xch)[k2 * 32 + lane] = z;
            }
            __syncthreads();

            // ---- combine: this warp owns k = r * (1024 / M) + 32 g + lane, g < 32 / M ----
            const size_t row_base = (size_t) item * N;
#pragma unroll
            for (int g = 0; g < C::KPW; ++g) {
                const int krow = r * C::KPW + g;                         // k >> 5
                c64 z[M];
#pragma unroll
                for (int rr = 0; rr < M; ++rr)
                    z[bitrev<M>(rr)] = reinterpret_cast<const c64*>(xch_base + rr * FFT1024_XCH_BYTES)[krow * 32 + lane];
                fft_dit_small<M>(z);
#pragma unroll
                for (int q = 0; q < M; ++q) {
                    float re, im;
                    cunpack(z[q], re, im);
                    const float pw = fmaf(re, re, im * im);
                    const int bin = krow * 32 + lane + 1024 * q;
                    if (MULTI) {
                        acc[g * M + q] += pw;
                    } else if (bin != 0) {
                        emit(row_base, bin, pw);                          // bin 0 waits for bin N-1 (below)
                    }
                    // bin N-1: k = 1023 (warp M-1, last group, lane 31), q = M-1
                    if (g == C::KPW - 1 && q == M - 1) dcacc = fmaf((float) (K - j), pw, dcacc);
                }
            }
            if (j == K - 1 && r == M - 1 && lane == 31) *dc_slot = dcacc;
            __syncthreads();                      // tiles are rewritten by the next frame; dc_slot is visible
            if (!MULTI) {
                // spectrum.c:30-33: the DC position repeats its left neighbour (bin N-1)
                if (tid == 0) emit(row_base, 0, *dc_slot);
                dcacc = 0.0f;
            }
        }

        if (MULTI) {
            // ---- row epilogue: bin = k + 1024 q ----
            const float dc = *dc_slot;
            const size_t row_base = (size_t) item * N;
#pragma unroll
            for (int g = 0; g < C::KPW; ++g) {
#pragma unroll
                for (int q = 0; q < M; ++q) {
                    const int bin = (r * C::KPW + g) * 32 + lane + 1024 * q;
                    emit(row_base, bin, bin == 0 ? dc : acc[g * M + q]);
                    acc[g * M + q] = 0.0f;
                }
            }
            dcacc = 0.0f;
        }
    }
}

template <int M>
int launch_m(const SpecParams& p, cudaStream_t stream)
{
    using C = MxCfg<M>;
    const uint64_t total = (uint64_t) p.n_streams * (uint64_t) p.n_rows;
    if (total >= (1ull << 31)) {
        set_error("spectrum: n_streams * n_rows = %llu exceeds 2^31 - 1 rows per launch", (unsigned long long) total);
        return B200_ERR_ARG;
    }
    auto kern = p.K > 1 ? (p.window ? spectrum_mx1024_kernel<M, true, true> : spectrum_mx1024_kernel<M, false, true>)
                        : (p.window ? spectrum_mx1024_kernel<M, true, false> : spectrum_mx1024_kernel<M, false, false>);
    if (int rc = ensure_dynamic_smem((const void*) kern, C::SMEM)) return rc;
    const int ctas_per_sm = cached_occupancy((const void*) kern, C::THREADS, C::SMEM);
    uint64_t grid = (uint64_t) sm_count() * (uint64_t) ctas_per_sm;
    if (grid > total) grid = total;
    kern<<<(unsigned) grid, C::THREADS, C::SMEM, stream>>>(p);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace

// N = 2048, 4096 or 8192, cmplx_u8 input.  p.twiddle must be the 1024-point table and
// p.twiddle_n the N-point table (capi.cu keeps both in the plan).
int launch_spectrum_mx1024(const SpecParams& p, int N, cudaStream_t stream)
{
    if ((uint64_t) p.n_streams * (uint64_t) p.n_rows == 0) return B200_OK;
    switch (N) {
        case 2048: return launch_m<2>(p, stream);
        case 4096: return launch_m<4>(p, stream);
        case 8192: return launch_m<8>(p, stream);
        default:
            set_error("spectrum_mx1024: N = %d", N);
            return B200_ERR_ARG;
    }
}

}  // namespace b200
