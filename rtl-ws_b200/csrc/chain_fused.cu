// chain_fused.cu -- spectrum + FM branch on the same IQ bytes.
// First cut: the two specialised kernels back to back on one stream (the IQ is read twice,
// the second time largely from L2 for batches that fit).  The single-pass fused kernel
// replaces the body of launch_chain_fused without touching its callers.
#include "b200_common.cuh"
#include "fm_kernels.cuh"
#include "spectrum_kernels.cuh"

namespace b200 {

int launch_spectrum1024(const SpecParams& p, cudaStream_t stream);
int launch_fm_chain(const FmParams& p, cudaStream_t stream);

int launch_chain_fused(const uint8_t* d_iq, int64_t stride, int n_streams, int64_t n_samples, float db_offset,
                       const float2* twiddle, float* d_db, float* d_audio, int64_t audio_stride, cudaStream_t stream)
{
    SpecParams sp;
    sp.iq = d_iq;
    sp.stream_stride_bytes = stride;
    sp.n_streams = n_streams;
    sp.n_rows = n_samples / 1024;
    sp.hop = 1024;
    sp.K = 1;
    sp.row_hop = 1024;
    sp.db = d_db;
    sp.power = nullptr;
    sp.db_u8 = nullptr;
    sp.db_offset = db_offset;
    sp.twiddle = twiddle;
    sp.window = nullptr;
    int rc = launch_spectrum1024(sp, stream);
    if (rc) return rc;
    FmParams fp;
    fp.iq = d_iq;
    fp.stream_stride_bytes = stride;
    fp.n_streams = n_streams;
    fp.n_samples = n_samples;
    fp.R = 10;
    fp.audio = d_audio;
    fp.audio_stride = audio_stride;
    fp.decimated = nullptr;
    fp.dec_stride = 0;
    return launch_fm_chain(fp, stream);
}

}  // namespace b200
