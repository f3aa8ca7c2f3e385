// spectrum64k_cluster.cu -- 65536-point power spectra (BASELINE config 4: wideband spectrogram, Hann
// window, 50 % overlap) on a CLUSTER of four CTAs whose shared memories hold the whole intermediate
// array: nothing but the IQ bytes and the finished rows ever touches L2 / HBM (sm_100a).
//
// Same four-step factorisation as spectrum64k.cu (which stays for K > 1 rows):
//   65536 = 64 x 1024,  X[k + 1024 q] = sum_{r<64} W_64^(r q) * ( W_N^(r k) * F_r[k] ),  F_r = FFT_1024( x[64 m + r] )
// but the 512 KB of Z[r][k] = W_N^(r k) F_r[k] no longer go through a global scratch.  The four CTAs of a
// cluster take one frame together:
//   phase 1  CTA c owns the polyphase branches r = 16 c .. 16 c + 15, i.e. bytes 32 c .. 32 c + 31 of every
//            128-byte row of the frame (1024 rows): 32 KB of input per CTA, copied asynchronously (cp.async,
//            word-swizzled so that the stride-32-byte reads of a branch are conflict-free).  Warp w runs the
//            32x32 register transform (fft1024_warp.cuh) on branches 2 w and 2 w + 1, multiplies by W_N^(r k)
//            and PUSHES Z[r][k] straight from its registers into the shared memory of the CTA that owns
//            column k (k / 256) with st.async, whose completion bytes are counted by that CTA's mbarrier:
//            no staging, no cluster-wide barrier, 96 of every 128 KB cross the SM-to-SM network.
//   phase 2  CTA d holds Z[0..63][256 d .. 256 d + 255] (128 KB); thread t loads column 256 d + t (64 LDS.64),
//            hands the buffer back to the four producers (remote mbarrier arrive) and runs the 64-point
//            transform in registers, |X|^2, dB / power / u8, 128-byte coalesced streaming stores.
// The next frame's input is fetched while phase 2 runs; with 50 % overlap only the new half is fetched (the two
// 16 KB half buffers alternate roles).  Arithmetic per frame is that of spectrum64k.cu instruction for
// instruction (spectrum.c:15-58, cbb_main.c:112-128), so the two kernels agree bit for bit.
#include <type_traits>
#include <utility>

#include "b200_common.cuh"
#include "fft1024_warp.cuh"
#include "spectrum_kernels.cuh"

namespace b200 {

namespace {

constexpr int N64K = 65536;
constexpr int CL = 4;                                  // CTAs per cluster
constexpr int C64_THREADS = 384;                       // 8 producer + 4 consumer warps (see the kernel)
constexpr int C64_WARPS = 8;                           // warps with an exchange tile (the producers)
constexpr int C64_Z_BYTES = 64 * 256 * 8;              // Z[64][256] complex
constexpr int C64_HALF_BYTES = 512 * 32;               // half a frame's rows, this CTA's 32 bytes of each
constexpr int C64_XCH_BYTES = 32 * 32 * 8;             // per warp, unpadded, 16-byte granules XOR-swizzled
constexpr int C64_OFF_IN = C64_Z_BYTES;
constexpr int C64_OFF_XCH = C64_OFF_IN + 2 * C64_HALF_BYTES;
constexpr int C64_OFF_BAR = C64_OFF_XCH + C64_WARPS * C64_XCH_BYTES;
constexpr int C64_SMEM = C64_OFF_BAR + 64;
static_assert(C64_SMEM <= 232448, "more shared memory than a CTA may have on sm_100");

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_count_x()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local_smem_addr` in the shared memory of CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// 8-byte store into a (possibly remote) CTA's shared memory; the destination's mbarrier counts its bytes
// (OFF is a compile-time byte offset folded into the instruction: no address arithmetic per store)
template <int OFF>
__device__ __forceinline__ void st_async_b64(uint32_t dst_cluster_addr, c64 v, uint32_t bar_cluster_addr)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0+%3], %1, [%2];" ::"r"(dst_cluster_addr),
                 "l"(v), "r"(bar_cluster_addr), "n"(OFF)
                 : "memory");
}
// f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>): a loop whose index is a constant expression
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}
// Arrive on a (possibly remote) CTA's mbarrier WITHOUT release semantics: a release at cluster scope is a
// MEMBAR.ALL.GPU in SASS, which waits for every global store the warp has in flight (64 spectrum stores per run:
// measured, 24 % of the consumers' stall samples).  The callers order what has to be ordered by a data dependency.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst_smem, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
// this thread's earlier cp.async copies arrive on `bar` when they have landed (the count is not raised)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp roles.  Phase 2 needs 64 complex values per thread (200+ registers), phase 1 about 150, and eight warps of
// the larger size are all an SM holds -- two per scheduler, not enough to hide the latencies of either phase (the
// symmetric first version of this kernel: 33 % of issue slots used, 98 Gsamples/s).  So the roles are split over
// twelve warps: warps 0-7 PRODUCE (phase 1, two branches per warp and frame), warps 8-11 CONSUME (phase 2, two runs
// of 32 columns per warp and frame) and fetch the input.  Producers work on frame f + 1 while consumers finish
// frame f.  With the inter-pass twiddles read from an L1-resident table instead of 62 registers, either role fits
// the 168 registers a 384-thread CTA leaves per thread (ptxas: 142), so no setmaxnreg is needed.
constexpr int C64_PRODUCERS = 8;
constexpr int C64_CONSUMERS = 4;

// Where the two halves of a frame's rows live.  Both roles replay the same sequence of frame addresses, so they
// agree on buffers and barrier parities without talking to each other.
struct InputState {
    int lo_buf = 0;                     // half buffer with rows 0..511 of the current frame
    const uint8_t* held_hi = nullptr;   // global rows behind the other half buffer
    uint32_t n_fetch0 = 0, n_fetch1 = 0;
    // advance to the frame at `nf`; returns true if only the new second half has to be fetched (into `fetch_buf`)
    __device__ __forceinline__ bool step(const uint8_t* nf, int& fetch_buf)
    {
        const bool half = (nf == held_hi);          // hop = N/2: the old second half is the new first half
        if (half) {
            fetch_buf = lo_buf;
            if (lo_buf) ++n_fetch1;
            else ++n_fetch0;
            lo_buf ^= 1;
        } else {
            fetch_buf = -1;
            ++n_fetch0;
            ++n_fetch1;
            lo_buf = 0;
        }
        held_hi = nf + N64K;
        return half;
    }
};

template <bool WINDOW>
__global__ void __launch_bounds__(C64_THREADS, 1) spectrum64k_cluster_kernel(const SpecParams p, const Spec64kExtra x)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();

    const c64* Zl = reinterpret_cast<const c64*>(smem);                       // Z[64][256], this CTA's columns
    uint8_t* inbuf = smem + C64_OFF_IN;                                       // two half buffers [512 rows][8 words]
    uint64_t* in_full = reinterpret_cast<uint64_t*>(smem + C64_OFF_BAR);      // [2], one per half buffer
    uint64_t* in_empty = in_full + 2;
    uint64_t* z_full = in_full + 3;
    uint64_t* z_free = in_full + 4;                                           // [2]: columns of consumer run 0 / run 1

    if (tid == 0) {
        mbar_init(&in_full[0], 32 * C64_CONSUMERS);
        mbar_init(&in_full[1], 32 * C64_CONSUMERS);
        mbar_init(in_empty, C64_PRODUCERS);
        mbar_init(z_full, 1);
        mbar_init(&z_free[0], CL * C64_CONSUMERS);
        mbar_init(&z_free[1], CL * C64_CONSUMERS);
        fence_mbar_init();
    }
    __syncthreads();

    // a cluster takes a CONTIGUOUS run of (stream, row) items, so that with hop = N/2 consecutive frames
    // share half their rows
    const int64_t total = (int64_t) p.n_streams * p.n_rows;
    const int64_t per_cluster = (total + cluster_count_x() - 1) / cluster_count_x();
    const int64_t item_begin = (int64_t) cluster_id_x() * per_cluster;
    const int64_t item_end = item_begin + per_cluster < total ? item_begin + per_cluster : total;
    const uint32_t n_mine = item_begin < item_end ? (uint32_t) (item_end - item_begin) : 0;

    if (tid == 0 && n_mine > 0) mbar_arrive_expect_tx(z_full, C64_Z_BYTES);
    cluster_sync_all();          // every CTA's barriers exist before anything remote is sent

    auto frame_of = [&](uint32_t it) {
        const int64_t item = item_begin + it;
        const int64_t s = item / p.n_rows;
        const int64_t row = item - s * p.n_rows;
        return p.iq + s * p.stream_stride_bytes + 2 * row * p.row_hop;
    };
    InputState in;

    if (warp < C64_PRODUCERS) {
        // ================================ producers: phase 1 ================================
        uint8_t* xch = smem + C64_OFF_XCH + warp * C64_XCH_BYTES;
        uint32_t z_dst[CL], zfull_dst[CL];
#pragma unroll
        for (int d = 0; d < CL; ++d) {
            z_dst[d] = map_to_rank(smem_u32(smem), d) + lane * 8;
            zfull_dst[d] = map_to_rank(smem_u32(z_full), d);
        }
        const int sw_in = (lane >> 2) & 7;
        // the eight XOR patterns of the tile swizzle as pointers, so that every tile access is base + immediate
        uint8_t* xw[8];
        const uint8_t* xr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            xw[j] = xch + ((((lane >> 1) ^ j) << 4) | ((lane & 1) << 3));
            xr[j] = xch + 256 * lane + ((j ^ (lane & 7)) << 4);
        }
        const float2* tw_lane = x.twiddle_32x32 + lane;       // [n2][k1] W_1024^(n2 k1), this lane's column (L1-resident)

        for (uint32_t it = 0; it < n_mine; ++it) {
            int unused;
            in.step(frame_of(it), unused);
            const int lo_buf = in.lo_buf, hi_buf = lo_buf ^ 1;
            mbar_wait(&in_full[0], (in.n_fetch0 - 1) & 1);
            mbar_wait(&in_full[1], (in.n_fetch1 - 1) & 1);
            c64 held[16];                                       // branch A's odd-k2 half, pushed after branch B
#pragma unroll 1
            for (int jb = 0; jb < 2; ++jb) {
                const int rl = 2 * warp + jb;                   // branch within this CTA
                const int r = 16 * (int) rank + rl;
                c64 a[32];
                {
                    // Hann window of sample 64*(32*n1 + lane) + r without a table:
                    //   w = 1/2 - 1/2 cos(2 pi n1 / 32 + phi),  phi = 2 pi (64 lane + r) / 65536
                    float cphi = 1.0f, sphi = 0.0f;
                    if (WINDOW) sincospif((float) (64 * lane + r) * (1.0f / 32768.0f), &sphi, &cphi);
                    const c64 bias1 = cpack(8421376.0f, 8421376.0f);        // 2^23 + 256 * 128
                    const int col = (((rl >> 1) ^ sw_in) << 2) + ((rl & 1) << 1);
                    const uint8_t* lo = inbuf + lo_buf * C64_HALF_BYTES + 32 * lane + col;
                    const uint8_t* hi = inbuf + hi_buf * C64_HALF_BYTES + 32 * lane + col;
#pragma unroll
                    for (int n1 = 0; n1 < 32; ++n1) {
                        const uint32_t v = *reinterpret_cast<const uint16_t*>((n1 < 16 ? lo : hi) + 1024 * (n1 & 15));
                        const int q = bitrev<32>(n1);
                        a[q] = cpack(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7504)),
                                     __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7514)));
                        if (WINDOW) {
                            const float w = fmaf(0.5f * sin32(n1), sphi, fmaf(-0.5f * cos32(n1), cphi, 0.5f));
                            a[q] = cmul2(csub(a[q], bias1), cpack(w, w));
                        }
                    }
                }
                if (jb == 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(in_empty);       // this warp is done with the frame's input bytes
                }
                fft1024_pass1<!WINDOW>(a);
                // transpose through the warp's tile: element [k1][n2] sits in 16-byte granule (n2 >> 1) ^ (k1 & 7) of row k1
                __syncwarp();
#pragma unroll
                for (int k1 = 0; k1 < 32; ++k1) {
                    float re, im;
                    cunpack(a[k1], re, im);
                    *reinterpret_cast<float2*>(xw[k1 & 7] + 256 * k1) = make_float2(re, im);
                }
                __syncwarp();
                c64 b[32];
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xr[m & 7] + 128 * (m >> 3));
                    b[bitrev<32>(2 * m)] = v.x;
                    b[bitrev<32>(2 * m + 1)] = v.y;
                }
                fft_dit32_pretwiddled_ldg(b, tw_lane);
                // W_N^(r k) F_r[k]: the table loads of half a row are issued together, ahead of their products (the
                // st.async below are volatile: inside the push loop every load would be waited for on its own --
                // measured, 56 % of all stall samples).  ld.global.cg: 128 KB per frame must not sweep the small L1.
                // (Moving this product to the consumers, fused into their first butterfly stage, was measured: 94 vs
                // 111 Gsamples/s -- 64 just-in-time L2 loads per column are worse than 32 batched ones per branch.)
                const float2* twr = x.twiddle_rk + r * 1024 + lane;
                if (r > 0) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float2 w[16];
#pragma unroll
                        for (int k2 = 0; k2 < 16; ++k2) w[k2] = __ldcg(twr + 32 * (16 * h + k2));
#pragma unroll
                        for (int k2 = 0; k2 < 16; ++k2) b[16 * h + k2] = cmul(b[16 * h + k2], w[k2].x, w[k2].y);
                    }
                }
                // Z hand-back is split by consumer run: columns with even k2 belong to run 0 of some consumer warp
                // (read at the very start of its frame), odd k2 to run 1 (read half a frame later).  Branch A pushes
                // its even half at once and keeps the odd half in registers until branch B is done, so the wait for
                // the late hand-back never sits between a producer and its next transform.  The consumers' arrives
                // are release.cluster; nothing they WROTE is read here (what follows are stores that cannot start
                // before the wait has succeeded), so the waits need no cluster-scope acquire -- which ptxas implements
                // as CCTL.IVALL after every poll: measured, 27 % of all stall samples and an L1 that never keeps the
                // twiddle table.
                // destination of element k2 of branch rr: CTA k2 >> 3, Z[rr][(k2 & 7) * 32 + lane]; per branch one base per
                // destination, the rest is an immediate
                uint32_t zrow[CL];
#pragma unroll
                for (int d = 0; d < CL; ++d) zrow[d] = z_dst[d] + (uint32_t) (r * 2048);
                if (jb == 0) {
                    if (it > 0) mbar_wait(&z_free[0], (it - 1) & 1);
                    static_for<16>([&](auto i) {
                        constexpr int k2 = 2 * decltype(i)::value;
                        st_async_b64<(k2 & 7) * 256>(zrow[k2 >> 3], b[k2], zfull_dst[k2 >> 3]);
                    });
#pragma unroll
                    for (int k2 = 1; k2 < 32; k2 += 2) held[k2 >> 1] = b[k2];
                } else {
                    if (it > 0) mbar_wait(&z_free[1], (it - 1) & 1);
                    static_for<16>([&](auto i) {
                        constexpr int k2 = 2 * decltype(i)::value + 1;
                        st_async_b64<(k2 & 7) * 256 - 2048>(zrow[k2 >> 3], held[k2 >> 1], zfull_dst[k2 >> 3]);     // branch r - 1
                    });
                    static_for<32>([&](auto i) {
                        constexpr int k2 = decltype(i)::value;
                        st_async_b64<(k2 & 7) * 256>(zrow[k2 >> 3], b[k2], zfull_dst[k2 >> 3]);
                    });
                }
            }
        }
    } else {
        // ================================ consumers: input fetch + phase 2 ================================
        const int ct = tid - 32 * C64_PRODUCERS;             // 0..127
        const int cw = ct >> 5;
        uint32_t zfree_dst[CL];                               // z_free[0] of every CTA of the cluster (z_free[1] is 8 bytes on)
#pragma unroll
        for (int d = 0; d < CL; ++d) zfree_dst[d] = map_to_rank(smem_u32(z_free), d);
        // copy half a frame (512 rows starting at `src_rows`, this CTA's 32 bytes of each) into half buffer hb: word w
        // of row m lands at word 8 m + (w ^ ((m >> 2) & 7)).  Thread ct takes word ct & 7 of rows (ct >> 3) + 16 i, whose
        // swizzle is cw for even i and cw + 4 for odd i.
        const uint32_t in_dst_even = smem_u32(inbuf) + 32 * (ct >> 3) + 4 * ((ct & 7) ^ cw);
        const uint32_t in_dst_odd = smem_u32(inbuf) + 32 * (ct >> 3) + 4 * ((ct & 7) ^ cw ^ 4);
        const int in_src0 = 128 * (ct >> 3) + 32 * (int) rank + 4 * (ct & 7);
        auto fetch_half = [&](const uint8_t* src_rows, int hb) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                cp_async_4(((i & 1) ? in_dst_odd : in_dst_even) + hb * C64_HALF_BYTES + 512 * i, src_rows + in_src0 + 2048 * i);
            cp_async_arrive(&in_full[hb]);
        };
        // fetch frame g (g >= 1: once the producers have taken frame g - 1 out of the buffers)
        uint32_t fetched = 0;                                 // frames fetched so far
        auto fetch_frame = [&](bool blocking) {
            const uint32_t g = fetched;
            if (g >= n_mine) return;
            if (g > 0) {
                if (blocking) mbar_wait(in_empty, (g - 1) & 1);
                else if (!mbar_try_wait(in_empty, (g - 1) & 1)) return;
            }
            const uint8_t* nf = frame_of(g);
            int fb;
            if (in.step(nf, fb)) {
                fetch_half(nf + N64K, fb);
            } else {
                fetch_half(nf, 0);
                fetch_half(nf + N64K, 1);
            }
            ++fetched;
        };
        fetch_frame(true);
        fetch_frame(true);

        const float dboff = p.db_offset - 16.0f * DB_PER_LOG2;
        float* const out_db = p.db;                          // kernel parameters, read once
        float* const out_power = p.power;
        uint8_t* const out_u8 = p.db_u8;
        for (uint32_t it = 0; it < n_mine; ++it) {
            const size_t row_base = (size_t) (item_begin + it) * N64K;
            mbar_wait(z_full, it & 1);
            if (ct == 0 && it + 1 < n_mine) mbar_arrive_expect_tx(z_full, C64_Z_BYTES);
#pragma unroll 1
            for (int jc = 0; jc < 2; ++jc) {
                const int kl = 64 * cw + 32 * jc + lane;      // column within this CTA
                c64 z[64];
#pragma unroll
                for (int r = 0; r < 64; ++r) z[bitrev<64>(r)] = Zl[r * 256 + kl];
                fft_dit64(z);
                // This warp no longer needs its columns of run jc: tell the four CTAs' producers.  The hand-back only
                // has to keep their next stores behind OUR LOADS (write after read), and loads are over once their
                // values exist: every output of the transform depends on all 64 loaded values, and the barrier address
                // below depends on an output (the comparison is never true: 0x7fc12345 is a NaN no sum produces), so
                // the arrive cannot issue before the last load has returned -- without any fence.
                {
                    float x0r, x0i;
                    cunpack(z[0], x0r, x0i);
                    const uint32_t dep = __float_as_uint(x0r) == 0x7fc12345u ? 4u : 0u;
                    const uint32_t all = __reduce_or_sync(0xffffffffu, dep);     // every lane's loads, not just lane 0's
                    if (lane == 0) {
#pragma unroll
                        for (int d = 0; d < CL; ++d) mbar_arrive_remote_relaxed((zfree_dst[d] + 8 * jc) ^ all);
                    }
                }
                const int k = 256 * (int) rank + kl;
                float pw[64];
#pragma unroll
                for (int q = 0; q < 64; ++q) {
                    float re, im;
                    cunpack(z[q], re, im);
                    pw[q] = fmaf(re, re, im * im);
                }
                // fftshift (spectrum.c:25): bin k + 1024 q shows at column 1024 ((q + 32) & 63) + k -- a compile-time
                // offset from one pointer per output array.  Bin 0 is not stored (the DC position shows bin N-1's value,
                // spectrum.c:30-33 with K = 1): the thread that owns bin N-1 writes it.
                const bool owns_first = (k == 0), owns_last = (k == 1023);
                if (out_db != nullptr) {
                    float* o = out_db + row_base + k;
#pragma unroll
                    for (int q = 0; q < 64; ++q) {
                        const float db = fmaf(DB_PER_LOG2, lg2_ftz(pw[q]), dboff);
                        if (q != 0 || !owns_first) __stcs(o + 1024 * ((q + 32) & 63), db);
                        if (q == 63 && owns_last) __stcs(out_db + row_base + N64K / 2, db);
                    }
                }
                if (out_power != nullptr) {
                    float* o = out_power + row_base + k;
#pragma unroll
                    for (int q = 0; q < 64; ++q) {
                        const float v = pw[q] * FFT1024_POWER_SCALE;
                        if (q != 0 || !owns_first) __stcs(o + 1024 * ((q + 32) & 63), v);
                        if (q == 63 && owns_last) __stcs(out_power + row_base + N64K / 2, v);
                    }
                }
                if (out_u8 != nullptr) {
                    uint8_t* o = out_u8 + row_base + k;
#pragma unroll
                    for (int q = 0; q < 64; ++q) {
                        int m = __float2int_rz(fmaf(DB_PER_LOG2, lg2_ftz(pw[q]), dboff));
                        m = m < 0 ? 0 : (m > 255 ? 255 : m);
                        if (q != 0 || !owns_first) o[1024 * ((q + 32) & 63)] = (uint8_t) m;
                        if (q == 63 && owns_last) out_u8[row_base + N64K / 2] = (uint8_t) m;
                    }
                }
                // the input of frame it + 2: as soon as the producers have emptied the buffers, at the latest now
                if (fetched == it + 2) fetch_frame(jc == 1);
            }
        }
    }
    cluster_sync_all();          // nobody leaves while a neighbour may still signal its barriers
}

}  // namespace

// Largest number of four-CTA clusters the device runs at once for this kernel (0: cluster launch unavailable).
static int cluster_capacity(const void* kern)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * 64, 1, 1);
    cfg.blockDim = dim3(C64_THREADS, 1, 1);
    cfg.dynamicSmemBytes = C64_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// K = 1 rows only; returns B200_ERR_ARG (without an error message) if the plan needs the scratch kernel instead
int launch_spectrum64k_cluster(const SpecParams& p, const Spec64kExtra& x, cudaStream_t stream)
{
    const int64_t total = (int64_t) p.n_streams * p.n_rows;
    if (total == 0) return B200_OK;
    auto kern = p.window ? spectrum64k_cluster_kernel<true> : spectrum64k_cluster_kernel<false>;
    if (int rc = ensure_dynamic_smem((const void*) kern, C64_SMEM)) return rc;
    static int capacity[64][2];
    static bool known[64][2];
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    const int wi = p.window ? 1 : 0;
    if (dev < 64 && !known[dev][wi]) {
        capacity[dev][wi] = cluster_capacity((const void*) kern);
        known[dev][wi] = true;
    }
    int64_t clusters = dev < 64 ? capacity[dev][wi] : cluster_capacity((const void*) kern);
    if (clusters <= 0) {
        set_error("spectrum 65536: the device does not co-schedule clusters of %d CTAs with %d bytes of shared memory", CL,
                  C64_SMEM);
        return B200_ERR_CUDA;
    }
    if (clusters > total) clusters = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) (CL * clusters), 1, 1);
    cfg.blockDim = dim3(C64_THREADS, 1, 1);
    cfg.dynamicSmemBytes = C64_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p, x));
    B200_LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace b200
