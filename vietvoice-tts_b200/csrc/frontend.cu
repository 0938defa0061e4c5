// Preprocess-graph and decode-graph kernels that are not GEMMs: log-mel front-end (smem FFT), text embedding
// gather, depthwise conv over rows, GRN, conditioning concat, Philox noise, Vocos im2col, iSTFT (smem inverse FFT,
// overlap-add, int16 pack).  All are bandwidth/latency bound and < 0.2 % of the path's FLOPs.
//
// Replaces the non-GEMM nodes of `preprocess.onnx` and `decode.onnx`
// (/root/reference/vietvoicetts/core/tts_engine.py:133-146, 176-187).
#include "frontend.h"
#include "ptx.cuh"

namespace vv {

// ------------------------------------------------------------------------------------------------ FFT-1024
// radix-2 DIT on 1024 complex points in shared memory, 256 threads; input already bit-reversed.
__device__ __forceinline__ void fft1024(float2* s, const float2* tw, bool inverse, int tid) {
#pragma unroll 1
  for (int len = 2, shift = 9; len <= 1024; len <<= 1, --shift) {
    const int half = len >> 1;
#pragma unroll
    for (int b = tid; b < 512; b += 256) {
      const int grp = b / half, j = b - grp * half;
      const int i0 = grp * len + j, i1 = i0 + half;
      float2 w = tw[j << shift];
      if (inverse) w.y = -w.y;
      const float2 a = s[i0], c = s[i1];
      const float2 t = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
      s[i0] = make_float2(a.x + t.x, a.y + t.y);
      s[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ mel front-end
__global__ void __launch_bounds__(1024) rms_scale_kernel(const int16_t* __restrict__ audio, int64_t n,
                                                         float target_rms, float* __restrict__ scale_out) {
  __shared__ float red[32];
  float s = 0.f;
  // eight samples per 16-byte load (the PCM buffer comes from the allocator: 256-byte aligned), scalar tail
  const int64_t n8 = ((reinterpret_cast<uintptr_t>(audio) & 15) == 0) ? n / 8 : 0;
  const uint4* a8 = reinterpret_cast<const uint4*>(audio);
  for (int64_t i = threadIdx.x; i < n8; i += blockDim.x) {
    const uint4 u = __ldg(a8 + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = (float)(int16_t)(w[k] & 0xffffu) * (1.0f / 32768.0f);
      const float hi = (float)(int16_t)(w[k] >> 16) * (1.0f / 32768.0f);
      s += lo * lo;
      s += hi * hi;
    }
  }
  for (int64_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
    const float v = (float)audio[i] * (1.0f / 32768.0f);
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      const float rms = sqrtf(s / (float)n);
      *scale_out = (rms < target_rms && rms > 0.f) ? target_rms / rms : 1.0f;
    }
  }
}

// one block per frame: windowed frame -> |rFFT| -> mel filterbank -> log
__global__ void __launch_bounds__(256)
mel_kernel(const int16_t* __restrict__ audio, int64_t n, const float* __restrict__ scale,
           const float* __restrict__ hann, const float2* __restrict__ tw_g, const float* __restrict__ fb, int n_mel,
           float clamp_min, int max_frames, float* __restrict__ mel_out) {
  __shared__ float2 s[1024];
  __shared__ float2 tw[512];
  __shared__ float mag[520];
  const int f = blockIdx.x;
  if (f >= max_frames) return;
  const int tid = threadIdx.x;
  for (int i = tid; i < 512; i += 256) tw[i] = tw_g[i];
  const float sc = *scale * (1.0f / 32768.0f);
  for (int i = tid; i < 1024; i += 256) {
    int64_t idx = (int64_t)f * 256 - 512 + i;
    if (idx < 0) idx = -idx;
    if (idx >= n) idx = 2 * (n - 1) - idx;
    const float v = (float)audio[idx] * sc * hann[i];
    s[__brev((unsigned)i) >> 22] = make_float2(v, 0.f);
  }
  __syncthreads();
  fft1024(s, tw, false, tid);
  for (int k = tid; k <= 512; k += 256) mag[k] = sqrtf(s[k].x * s[k].x + s[k].y * s[k].y);
  __syncthreads();
  if (tid < n_mel) {
    float acc = 0.f;
    for (int k = 0; k <= 512; ++k) acc += mag[k] * __ldg(fb + k * n_mel + tid);
    mel_out[(size_t)f * n_mel + tid] = logf(fmaxf(acc, clamp_min));
  }
}

void launch_mel(const int16_t* audio, int64_t n, float target_rms, float* scale_tmp, const float* hann,
                const float2* tw, const float* fb, int n_mel, float clamp_min, int frames, float* mel_out,
                cudaStream_t st) {
  rms_scale_kernel<<<1, 1024, 0, st>>>(audio, n, target_rms, scale_tmp);
  if (frames > 0) mel_kernel<<<frames, 256, 0, st>>>(audio, n, scale_tmp, hann, tw, fb, n_mel, clamp_min, frames, mel_out);
}

// ------------------------------------------------------------------------------------------------ text path
// tx[row, :] = embed[ids[row]] + pos_table[min(pos, L-1)]  (valid rows only; others 0)
__global__ void text_gather_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ row_pos,
                                   const uint8_t* __restrict__ row_mask, const float* __restrict__ embed,
                                   const float* __restrict__ pos_table, int pos_len, int rows, int td,
                                   float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * td) return;
  const int r = i / td, c = i - (size_t)r * td;
  float v = 0.f;
  if (row_mask[r]) {
    int p = row_pos[r];
    if (p > pos_len - 1) p = pos_len - 1;
    v = embed[(size_t)ids[r] * td + c] + pos_table[(size_t)p * td + c];
  }
  out[i] = v;
}
void launch_text_gather(const int32_t* ids, const int32_t* row_pos, const uint8_t* row_mask, const float* embed,
                        const float* pos_table, int pos_len, int rows, int td, float* out, cudaStream_t st) {
  const size_t n = (size_t)rows * td;
  if (n) text_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ids, row_pos, row_mask, embed, pos_table, pos_len, rows, td, out);
}

// depthwise conv over rows with per-sequence zero padding: out[r,c] = b[c] + sum_k w[c,k] * x[r+k-K/2, c]
__global__ void dwconv_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_pos,
                                   const int32_t* __restrict__ row_len, const float* __restrict__ w,
                                   const float* __restrict__ b, int rows, int C, int K, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * C) return;
  const int r = i / C, c = i - (size_t)r * C;
  const int pos = row_pos[r], len = row_len[r];
  float acc = b[c];
  const int h = K / 2;
  for (int k = 0; k < K; ++k) {
    const int p = pos + k - h;
    if (p >= 0 && p < len) acc += w[c * K + k] * x[(size_t)(r + k - h) * C + c];
  }
  out[i] = acc;
}
// the same, four channels per thread (C % 4 == 0, K <= 8): 16-byte loads / stores, the taps of the four channels in
// registers.  Bandwidth-bound: a row is read once from HBM and 7x from L1/L2 by its neighbours.
template <int K>
__global__ void __launch_bounds__(256)
dwconv_rows4_kernel(const float4* __restrict__ x, const int32_t* __restrict__ row_pos,
                    const int32_t* __restrict__ row_len, const float* __restrict__ w, const float* __restrict__ b,
                    int rows, int C4, float4* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * C4) return;
  const int r = i / C4, c4 = i - (size_t)r * C4;
  const int pos = row_pos[r], len = row_len[r];
  float wk[4][K];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) wk[j][k] = __ldg(w + (size_t)(c4 * 4 + j) * K + k);
  float4 acc = __ldg(reinterpret_cast<const float4*>(b) + c4);
  constexpr int h = K / 2;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int p = pos + k - h;
    if (p >= 0 && p < len) {
      const float4 v = x[(size_t)(r + k - h) * C4 + c4];
      acc.x += wk[0][k] * v.x; acc.y += wk[1][k] * v.y; acc.z += wk[2][k] * v.z; acc.w += wk[3][k] * v.w;
    }
  }
  out[i] = acc;
}
void launch_dwconv_rows(const float* x, const int32_t* row_pos, const int32_t* row_len, const float* w,
                        const float* b, int rows, int C, int K, float* out, cudaStream_t st) {
  const size_t n = (size_t)rows * C;
  if (!n) return;
  if (K == 7 && C % 4 == 0) {
    const size_t n4 = n / 4;
    dwconv_rows4_kernel<7><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(x), row_pos, row_len, w, b, rows, C / 4, reinterpret_cast<float4*>(out));
    return;
  }
  dwconv_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, row_pos, row_len, w, b, rows, C, K, out);
}

// GRN (ConvNeXt-V2): per sequence, per channel L2 norm over TIME.  Two-stage reduction with a fixed summation
// order (no float atomics): the same inputs give bit-identical outputs on every run.
__global__ void grn_sumsq_kernel(const float* __restrict__ h, const int32_t* __restrict__ seq_off,
                                 const int32_t* __restrict__ seq_len, int C, int rows_per_block,
                                 float* __restrict__ partial /* [n_seq][gridDim.y][C] */) {
  const int seq = blockIdx.z;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int len = seq_len[seq], off = seq_off[seq];
  const int r0 = blockIdx.y * rows_per_block;
  int r1 = r0 + rows_per_block;
  if (r1 > len) r1 = len;
  float s = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float v = h[(size_t)(off + r) * C + c];
    s += v * v;
  }
  partial[((size_t)seq * gridDim.y + blockIdx.y) * C + c] = s;
}
__global__ void grn_norm_kernel(const float* __restrict__ partial, int n_blk, int C, float* __restrict__ nx) {
  // one block per sequence
  __shared__ float red[32];
  const int seq = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float g2 = 0.f;
    for (int k = 0; k < n_blk; ++k) g2 += partial[((size_t)seq * n_blk + k) * C + c];
    const float g = sqrtf(g2);
    nx[(size_t)seq * C + c] = g;      // Gx for now; normalised below
    s += g;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) red[0] = s / (float)C;
  }
  __syncthreads();
  const float mean = red[0];
  for (int c = threadIdx.x; c < C; c += blockDim.x) nx[(size_t)seq * C + c] = nx[(size_t)seq * C + c] / (mean + 1e-6f);
}
__global__ void grn_apply_kernel(const float* __restrict__ h, const int32_t* __restrict__ row_seq,
                                 const float* __restrict__ nx, const float* __restrict__ g,
                                 const float* __restrict__ b, int rows, int C, bf16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * C) return;
  const int r = i / C, c = i - (size_t)r * C;
  const int seq = row_seq[r];
  float v = 0.f;
  if (seq >= 0) {
    const float x = h[i];
    v = g[c] * (x * nx[(size_t)seq * C + c]) + b[c] + x;
  }
  out[i] = __float2bfloat16(v);
}
// four channels per thread: one 16-byte load of h, one 8-byte store of bf16
__global__ void __launch_bounds__(256)
grn_apply4_kernel(const float4* __restrict__ h, const int32_t* __restrict__ row_seq, const float* __restrict__ nx,
                  const float* __restrict__ g, const float* __restrict__ b, int rows, int C4, uint2* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * C4) return;
  const int r = i / C4, c4 = i - (size_t)r * C4;
  const int seq = row_seq[r];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (seq >= 0) {
    const float4 x = h[i];
    const float4 n4 = __ldg(reinterpret_cast<const float4*>(nx + (size_t)seq * C4 * 4) + c4);
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + c4);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(b) + c4);
    v.x = g4.x * (x.x * n4.x) + b4.x + x.x;
    v.y = g4.y * (x.y * n4.y) + b4.y + x.y;
    v.z = g4.z * (x.z * n4.z) + b4.z + x.z;
    v.w = g4.w * (x.w * n4.w) + b4.w + x.w;
  }
  uint2 u;
  u.x = pack_bf16(v.x, v.y);
  u.y = pack_bf16(v.z, v.w);
  out[i] = u;
}
void launch_grn(const float* h, const int32_t* seq_off, const int32_t* seq_len, const int32_t* row_seq, int n_seq,
                int max_len, int rows, int C, const float* g, const float* b, float* gx2, float* nx, bf16* out,
                cudaStream_t st) {
  if (rows == 0 || n_seq == 0) return;
  const int rpb = 64;   // gx2 holds [n_seq][ceil(max_len/64)][C] partial sums
  dim3 grid((C + 127) / 128, (max_len + rpb - 1) / rpb, n_seq);
  grn_sumsq_kernel<<<grid, 128, 0, st>>>(h, seq_off, seq_len, C, rpb, gx2);
  grn_norm_kernel<<<n_seq, 256, 0, st>>>(gx2, grid.y, C, nx);
  const size_t n = (size_t)rows * C;
  if (C % 4 == 0)
    grn_apply4_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(h), row_seq, nx, g, b,
                                                                      rows, C / 4, reinterpret_cast<uint2*>(out));
  else
    grn_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h, row_seq, nx, g, b, rows, C, out);
}

// cat_b[row, :] = [ mel (cond rows) or 0 (uncond rows) | text | zero pad ]  -> bf16 [rows, ld]
__global__ void cat_cond_kernel(const float* __restrict__ mel, const float* __restrict__ tx,
                                const uint8_t* __restrict__ row_mask, int rows, int R, int n_mel, int td, int ld,
                                bf16* __restrict__ out, float* __restrict__ out_f32) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * ld) return;
  const int r = i / ld, c = i - (size_t)r * ld;
  float v = 0.f;
  if (row_mask[r]) {
    if (c < n_mel) v = r < R ? mel[(size_t)r * n_mel + c] : 0.f;
    else if (c < n_mel + td) v = tx[(size_t)r * td + (c - n_mel)];
  }
  out[i] = __float2bfloat16(v);
  if (out_f32) out_f32[i] = v;
}
void launch_cat_cond(const float* mel, const float* tx, const uint8_t* row_mask, int rows, int R, int n_mel, int td,
                     int ld, bf16* out, float* out_f32, cudaStream_t st) {
  const size_t n = (size_t)rows * ld;
  if (n) cat_cond_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mel, tx, row_mask, rows, R, n_mel, td, ld, out, out_f32);
}

// ------------------------------------------------------------------------------------------------ Philox N(0,1)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__global__ void philox_normal_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t key) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)key, (uint32_t)(key >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  float z[4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float u1 = ((float)c[2 * i] + 1.0f) * 2.3283064365386963e-10f;       // (0,1]
    const float u2 = (float)c[2 * i + 1] * 2.3283064365386963e-10f;            // [0,1)
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z[2 * i] = rad * cs;
    z[2 * i + 1] = rad * sn;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (q * 4 + i < n) out[q * 4 + i] = z[i];
}
void launch_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t key, cudaStream_t st) {
  const int64_t q = (n + 3) / 4;
  if (q) philox_normal_kernel<<<(unsigned)((q + 255) / 256), 256, 0, st>>>(out, n, seed, key);
}

// noise fp32 [rows_u, n_mel] (one utterance) -> bf16 copies in both CFG halves of noise_b [M, ld]
__global__ void noise_to_bf16_kernel(const float* __restrict__ noise, int rows_u, int n_mel, bf16* __restrict__ nb0,
                                     bf16* __restrict__ nb1, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_u * n_mel) return;
  const int r = i / n_mel, c = i - r * n_mel;
  const bf16 v = __float2bfloat16(noise[i]);
  nb0[(size_t)r * ld + c] = v;
  nb1[(size_t)r * ld + c] = v;
}
void launch_noise_to_bf16(const float* noise, int rows_u, int n_mel, bf16* nb0, bf16* nb1, int ld, cudaStream_t st) {
  const int n = rows_u * n_mel;
  if (n) noise_to_bf16_kernel<<<(n + 255) / 256, 256, 0, st>>>(noise, rows_u, n_mel, nb0, nb1, ld);
}

// ------------------------------------------------------------------------------------------------ Vocos / iSTFT
// im2col for Conv1d(n_mel -> voc_dim, k): A[r, tap*n_mel + c] = mel[src_row[r] + tap - k/2, c] inside the sequence
__global__ void voc_im2col_kernel(const float* __restrict__ mel, const int32_t* __restrict__ src_row,
                                  const int32_t* __restrict__ row_pos, const int32_t* __restrict__ row_len, int rows,
                                  int n_mel, int K, int ld, bf16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * ld) return;
  const int r = i / ld, col = i - (size_t)r * ld;
  float v = 0.f;
  if (col < K * n_mel) {
    const int tap = col / n_mel, c = col - tap * n_mel;
    const int p = row_pos[r] + tap - K / 2;
    if (p >= 0 && p < row_len[r]) v = mel[(size_t)(src_row[r] + tap - K / 2) * n_mel + c];
  }
  out[i] = __float2bfloat16(v);
}
void launch_voc_im2col(const float* mel, const int32_t* src_row, const int32_t* row_pos, const int32_t* row_len,
                       int rows, int n_mel, int K, int ld, bf16* out, cudaStream_t st) {
  const size_t n = (size_t)rows * ld;
  if (n) voc_im2col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mel, src_row, row_pos, row_len, rows, n_mel, K, ld, out);
}

// head [logmag(513) | phase(513)] per frame -> irFFT-1024 -> * hann -> overlap-add -> / window envelope -> trim
// n_fft/2 -> scale, clamp -> int16 (truncate toward zero), for every chunk of the batch in ONE launch.
//
// A block owns OLA_HOPS consecutive output hops (256 samples each) of one chunk.  Output hop h receives frames
// h-1 .. h+2, so the block inverse-transforms OLA_HOPS + 3 frames in shared memory and accumulates their windowed
// samples straight into a shared output strip: the 4 KB/frame `frames` array of the two-kernel version never exists
// (HBM traffic per frame: 1026 fp32 read (x 16/13 for the halo frames) + 256 int16 written).  Frames are visited in
// DEscending order, which is the order the per-sample loop of the former ola kernel added them in: same bits out.
constexpr int OLA_HOPS = 13;
__global__ void __launch_bounds__(256)
istft_ola_kernel(const float* __restrict__ head, int ld_head, const float* __restrict__ hann_g,
                 const float2* __restrict__ tw_g, float mag_clip, const int32_t* __restrict__ dec_off,
                 const int32_t* __restrict__ dec_len, const int64_t* __restrict__ pcm_off, float pcm_scale,
                 int16_t* __restrict__ pcm) {
  __shared__ float2 s[1024];
  __shared__ float2 tw[512];
  __shared__ float hann[1024];
  __shared__ float acc[OLA_HOPS * 256];
  const int chunk = blockIdx.y, tid = threadIdx.x;
  const int n_frames = dec_len[chunk];
  const int n_hops = n_frames - 1;                 // output samples = (n_frames - 1) * 256
  const int h0 = blockIdx.x * OLA_HOPS;
  if (h0 >= n_hops) return;
  const int hops = min(OLA_HOPS, n_hops - h0);
  for (int i = tid; i < 512; i += 256) tw[i] = tw_g[i];
  for (int i = tid; i < 1024; i += 256) hann[i] = hann_g[i];
  for (int i = tid; i < OLA_HOPS * 256; i += 256) acc[i] = 0.f;
  const int base = 256 * h0 + 512;                 // padded position of this block's first output sample
  const int f_lo = max(h0 - 1, 0), f_hi = min(h0 + hops + 1, n_frames - 1);
  const float* hbase = head + (size_t)dec_off[chunk] * ld_head;
  for (int f = f_hi; f >= f_lo; --f) {
    __syncthreads();                               // previous frame's samples consumed, tables / acc initialised
    const float* hr = hbase + (size_t)f * ld_head;
    for (int k = tid; k <= 512; k += 256) {
      const float mag = fminf(expf(hr[k]), mag_clip);
      float sn, cs;
      sincosf(hr[513 + k], &sn, &cs);
      float2 X = make_float2(mag * cs, mag * sn);
      if (k == 0 || k == 512) X.y = 0.f;           // irfft ignores Im of DC / Nyquist
      s[__brev((unsigned)k) >> 22] = X;
      if (k > 0 && k < 512) s[__brev((unsigned)(1024 - k)) >> 22] = make_float2(X.x, -X.y);
    }
    __syncthreads();
    fft1024(s, tw, true, tid);
    const int shift = 256 * f - base;              // acc index of sample 0 of this frame
    for (int n = tid; n < 1024; n += 256) {
      const int i = n + shift;
      // separate multiply / add roundings (no FMA contraction): bit-identical to summing stored fp32 frames
      if (i >= 0 && i < hops * 256)
        acc[i] = __fadd_rn(acc[i], __fmul_rn(__fmul_rn(s[n].x, 1.0f / 1024.0f), hann[n]));
    }
  }
  __syncthreads();
  int16_t* out = pcm + pcm_off[chunk] + (int64_t)h0 * 256;
  for (int i = tid; i < hops * 256; i += 256) {
    const int p = base + i;
    const int fh = p >> 8;
    float env = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int f = fh - d;
      if (f >= 0 && f < n_frames) {
        const float w = hann[p - f * 256];
        env += w * w;
      }
    }
    const float a = acc[i];
    float v = env > 1e-11f ? a / env : a;
    v = fminf(fmaxf(v * pcm_scale, -32768.f), 32767.f);
    out[i] = (int16_t)v;   // float -> int conversion truncates toward zero, as numpy astype does
  }
}

void launch_istft_ola(const float* head, int ld_head, const float* hann, const float2* tw, float mag_clip,
                      const int32_t* dec_off, const int32_t* dec_len, const int64_t* pcm_off, int n_chunks,
                      int max_frames, float pcm_scale, int16_t* pcm, cudaStream_t st) {
  if (n_chunks <= 0 || max_frames <= 1) return;
  dim3 grid((max_frames - 1 + OLA_HOPS - 1) / OLA_HOPS, n_chunks);
  istft_ola_kernel<<<grid, 256, 0, st>>>(head, ld_head, hann, tw, mag_clip, dec_off, dec_len, pcm_off, pcm_scale, pcm);
}

// ------------------------------------------------------------------------------------------------ cross-fade
// Clip-fix + RMS-matched cos^2 / sin^2 cross-fade of a list of int16 chunks, bit-exact with the reference's numpy code
// (/root/reference/vietvoicetts/core/audio_processor.py: fix_clipped_audio :47-58,
// concatenate_with_crossfade_improved :123-193).  What "bit-exact" needs:
//   * np.abs on int16 leaves -32768 negative, so only +-32767 trips the clip fix; the fix multiplies in float64;
//   * the level ratio is float32 arithmetic on np.mean(x.astype(float32) ** 2), and numpy sums float32 arrays
//     PAIRWISE (blocks of <= 128 elements with 8 interleaved accumulators, halves split at multiples of 8): the same
//     tree is walked here;
//   * every float -> int16 cast truncates toward zero and wraps modulo 2^16 (a ratio of 1.5 can overflow int16: the
//     reference wraps, so do we);
//   * the fade tables are float64 and are computed by numpy on the host (cos / sin differ in the last bit between
//     libraries), multiplied and added without FMA contraction.
// The fold is sequential only through the level ratio: ratio[i] depends on the tail of chunk i-1 AFTER its own
// adjustment.  One block walks the junctions (xf_ratio_kernel); the samples are then written fully in parallel.
struct XfChunk {
  const int16_t* src;
  int64_t len;
  int64_t out_off;     // where sample 0 of this chunk lands in the joined wave
};

__device__ __forceinline__ int16_t xf_fixed(int16_t v, int clipped) {
  if (!clipped) return v;
  const double s = __dmul_rn((double)v, 26214.0 / 32767.0);
  return (int16_t)(int32_t)s;                       // trunc toward zero
}
__device__ __forceinline__ int16_t xf_adjusted(int16_t fixed, int has_ratio, float ratio) {
  if (!has_ratio) return fixed;
  return (int16_t)__float2int_rz(__fmul_rn((float)fixed, ratio));     // wraps like the numpy cast
}

__global__ void __launch_bounds__(1024) xf_clip_kernel(const XfChunk* __restrict__ ch, int* __restrict__ clipped) {
  __shared__ int red[32];
  const XfChunk c = ch[blockIdx.x];
  int mx = -32768;
  for (int64_t i = threadIdx.x; i < c.len; i += blockDim.x) {
    const int v = c.src[i];
    const int a = v == -32768 ? -32768 : (v < 0 ? -v : v);           // np.abs(int16)
    mx = max(mx, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) clipped[blockIdx.x] = mx >= 32767 ? 1 : 0;
  }
}

// numpy's pairwise float32 sum (numpy/_core/src/umath/loops_utils.h.src), same association order
__device__ float xf_pairwise_sum(const float* a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(xf_pairwise_sum(a, n2), xf_pairwise_sum(a + n2, n - n2));
}

// one block: for every junction i = 1 .. n-1 the level ratio of chunk i (has[i] = 0: levels too low, no adjustment)
__global__ void __launch_bounds__(256) xf_ratio_kernel(const XfChunk* __restrict__ ch, const int* __restrict__ clipped,
                                                       int n, int nf, int* __restrict__ has, float* __restrict__ ratio) {
  extern __shared__ float sq[];                     // [2][nf]: squares of the previous tail, of the next head
  __shared__ int s_has;
  __shared__ float s_ratio;
  if (threadIdx.x == 0) { s_has = 0; s_ratio = 1.f; has[0] = 0; ratio[0] = 1.f; }
  __syncthreads();
  for (int i = 1; i < n; ++i) {
    const XfChunk p = ch[i - 1], c = ch[i];
    const int ph = s_has;
    const float pr = s_ratio;
    const int pc = clipped[i - 1], cc = clipped[i];
    for (int k = threadIdx.x; k < nf; k += blockDim.x) {
      const float a = (float)xf_adjusted(xf_fixed(p.src[p.len - nf + k], pc), ph, pr);
      const float b = (float)xf_fixed(c.src[k], cc);
      sq[k] = __fmul_rn(a, a);
      sq[nf + k] = __fmul_rn(b, b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float rp = __fsqrt_rn(__fdiv_rn(xf_pairwise_sum(sq, nf), (float)nf));
      const float rn = __fsqrt_rn(__fdiv_rn(xf_pairwise_sum(sq + nf, nf), (float)nf));
      int h = 0;
      float r = 1.f;
      if (rp > 100.f && rn > 100.f) {
        h = 1;
        r = fminf(fmaxf(__fdiv_rn(rp, rn), 0.7f), 1.5f);
      }
      s_has = h; s_ratio = r;
      has[i] = h; ratio[i] = r;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) xf_apply_kernel(const XfChunk* __restrict__ ch, const int* __restrict__ clipped,
                                                       const int* __restrict__ has, const float* __restrict__ ratio,
                                                       int n, int nf, const double* __restrict__ fade_out,
                                                       const double* __restrict__ fade_in, int16_t* __restrict__ out) {
  const int i = blockIdx.y;
  const XfChunk c = ch[i];
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.len) return;
  if (i + 1 < n && k >= c.len - nf) return;                 // the tail is written as the next chunk's seam
  const int16_t a = xf_adjusted(xf_fixed(c.src[k], clipped[i]), has[i], ratio[i]);
  if (i > 0 && k < nf) {
    const XfChunk p = ch[i - 1];
    const int16_t t = xf_adjusted(xf_fixed(p.src[p.len - nf + k], clipped[i - 1]), has[i - 1], ratio[i - 1]);
    const double v = __dadd_rn(__dmul_rn((double)(float)t, fade_out[k]), __dmul_rn((double)(float)a, fade_in[k]));
    out[c.out_off + k] = (int16_t)(int32_t)v;
  } else {
    out[c.out_off + k] = a;
  }
}

void launch_crossfade(const void* chunks, int n, int64_t max_len, int nf, const double* fade_out, const double* fade_in,
                      int* clipped, int* has, float* ratio, int16_t* out, cudaStream_t st) {
  const XfChunk* ch = static_cast<const XfChunk*>(chunks);
  xf_clip_kernel<<<n, 1024, 0, st>>>(ch, clipped);
  xf_ratio_kernel<<<1, 256, (size_t)2 * nf * sizeof(float), st>>>(ch, clipped, n, nf, has, ratio);
  dim3 grid((unsigned)((max_len + 255) / 256), n);
  xf_apply_kernel<<<grid, 256, 0, st>>>(ch, clipped, has, ratio, n, nf, fade_out, fade_in, out);
}

// ------------------------------------------------------------------------------------------------ weight layout
// conv_pos weight [dim, cg, taps] fp32 -> bf16 [groups][taps][co(cg)][ci(cg)]
__global__ void permute_conv_w_kernel(const float* __restrict__ w, int dim, int cg, int taps, bf16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)dim * cg * taps) return;
  // out index: ((g*taps + tap)*cg + co)*cg + ci
  const int ci = i % cg;
  size_t t = i / cg;
  const int co = t % cg; t /= cg;
  const int tap = t % taps;
  const int g = t / taps;
  out[i] = __float2bfloat16(w[((size_t)(g * cg + co) * cg + ci) * taps + tap]);
}
void launch_permute_conv_w(const float* w, int dim, int cg, int taps, bf16* out, cudaStream_t st) {
  const size_t n = (size_t)dim * cg * taps;
  permute_conv_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, dim, cg, taps, out);
}
// Vocos embed weight [vd, n_mel, K] fp32 -> bf16 [vd, ld] with column tap*n_mel + c
__global__ void permute_embed_w_kernel(const float* __restrict__ w, int vd, int n_mel, int K, int ld, bf16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)vd * ld) return;
  const int o = i / ld, col = i - (size_t)o * ld;
  float v = 0.f;
  if (col < K * n_mel) {
    const int tap = col / n_mel, c = col - tap * n_mel;
    v = w[((size_t)o * n_mel + c) * K + tap];
  }
  out[i] = __float2bfloat16(v);
}
void launch_permute_embed_w(const float* w, int vd, int n_mel, int K, int ld, bf16* out, cudaStream_t st) {
  const size_t n = (size_t)vd * ld;
  permute_embed_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, vd, n_mel, K, ld, out);
}

}  // namespace vv
