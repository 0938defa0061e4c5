// Launchers of the non-GEMM preprocess / decode kernels (frontend.cu).
#pragma once
#include "kernels.h"

namespace vv {

void launch_mel(const int16_t* audio, int64_t n, float target_rms, float* scale_tmp, const float* hann,
                const float2* tw, const float* fb, int n_mel, float clamp_min, int frames, float* mel_out,
                cudaStream_t st);
void launch_text_gather(const int32_t* ids, const int32_t* row_pos, const uint8_t* row_mask, const float* embed,
                        const float* pos_table, int pos_len, int rows, int td, float* out, cudaStream_t st);
void launch_dwconv_rows(const float* x, const int32_t* row_pos, const int32_t* row_len, const float* w,
                        const float* b, int rows, int C, int K, float* out, cudaStream_t st);
void launch_grn(const float* h, const int32_t* seq_off, const int32_t* seq_len, const int32_t* row_seq, int n_seq,
                int max_len, int rows, int C, const float* g, const float* b, float* gx2, float* nx, bf16* out,
                cudaStream_t st);
void launch_cat_cond(const float* mel, const float* tx, const uint8_t* row_mask, int rows, int R, int n_mel, int td,
                     int ld, bf16* out, float* out_f32, cudaStream_t st);
void launch_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t key, cudaStream_t st);
void launch_noise_to_bf16(const float* noise, int rows_u, int n_mel, bf16* nb0, bf16* nb1, int ld, cudaStream_t st);
void launch_voc_im2col(const float* mel, const int32_t* src_row, const int32_t* row_pos, const int32_t* row_len,
                       int rows, int n_mel, int K, int ld, bf16* out, cudaStream_t st);
// all chunks of a batch in one launch: chunk c owns head rows [dec_off[c], dec_off[c] + dec_len[c]) and writes
// (dec_len[c] - 1) * 256 samples at pcm + pcm_off[c]; max_frames = max over dec_len (host copy)
void launch_istft_ola(const float* head, int ld_head, const float* hann, const float2* tw, float mag_clip,
                      const int32_t* dec_off, const int32_t* dec_len, const int64_t* pcm_off, int n_chunks,
                      int max_frames, float pcm_scale, int16_t* pcm, cudaStream_t st);
// chunks: device array of n {const int16_t* src; int64_t len; int64_t out_off} records (frontend.cu XfChunk);
// clipped / has / ratio: n-element scratch; nf: cross-fade samples (every chunk must hold >= 2 * nf samples)
void launch_crossfade(const void* chunks, int n, int64_t max_len, int nf, const double* fade_out, const double* fade_in,
                      int* clipped, int* has, float* ratio, int16_t* out, cudaStream_t st);
void launch_permute_conv_w(const float* w, int dim, int cg, int taps, bf16* out, cudaStream_t st);
void launch_permute_embed_w(const float* w, int vd, int n_mel, int K, int ld, bf16* out, cudaStream_t st);

}  // namespace vv
