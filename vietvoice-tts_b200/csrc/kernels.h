// Internal launcher interface between the engine (engine.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>

#include <mutex>

namespace vv {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting of a kernel: a process that creates engines
// on several GPUs (the C ABI allows it) must set it once on each of them, and two threads may launch at the same time.
// `once(fn)` runs fn the first time it is called with a given device current, under a mutex.
class DeviceOnce {
 public:
  template <typename F>
  void once(F&& fn) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    std::lock_guard<std::mutex> g(mu_);
    if (!done_[dev]) {
      fn();
      done_[dev] = true;
    }
  }

 private:
  static constexpr int kMaxDevices = 64;
  std::mutex mu_;
  bool done_[kMaxDevices] = {};
};

typedef __nv_bfloat16 bf16;

enum Act { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2, ACT_MISH = 3 };

// Fused GEMM epilogue, applied per accumulator element (row m, column n):
//   v = acc (+ bias[n]); RoPE on the rotated column ranges; v = act(v); v *= gate[n];
//   v += resid[m, n]; rows with row_mask[m]==0 -> 0; store fp32 and/or bf16.
struct GemmEpi {
  const float* bias = nullptr;
  const float* gate = nullptr;
  const float* resid = nullptr;
  int ld_resid = 0;
  float* out_f32 = nullptr;
  int ld_f32 = 0;
  bf16* out_bf16 = nullptr;
  int ld_bf16 = 0;
  const uint8_t* row_mask = nullptr;
  const int32_t* row_pos = nullptr;   // RoPE position of each row
  const float2* rope_cs = nullptr;    // [max_pos][32] (cos, sin)
  int rope_dim = 0;                   // columns [0,rope_dim) and [rope_off2, rope_off2+rope_dim) are rotated
  int rope_off2 = 0;
  int act = ACT_NONE;
};

struct GemmShape {
  int M = 0, N = 0, K = 0;  // C[M,N] = A[M,K] * B[N,K]^T ; K % 64 == 0
  // grouped-conv mode (conv_taps > 0): A is [M, groups*64] activations, B is [groups*taps*64, 64] weights,
  // C[m, g*64+j] = sum_{tap,ci} A[m + tap - taps/2, g*64+ci] * B[(g*taps+tap)*64 + j, ci]
  int conv_taps = 0;
  int conv_groups = 0;
};

// Launch with programmatic stream serialization (see pdl_wait / pdl_trigger in ptx.cuh) when pdl_set(true) is in
// effect on this thread, as an ordinary stream-ordered launch otherwise.  Every kernel launched through this MUST call
// pdl_wait() before its first access to global memory another kernel may write or read.
// The engine turns it on for small batches only: with few tiles per kernel the launch gap, prologue and pipeline fill
// of the next kernel are a sizeable part of a 5-20 us kernel and overlap the tail of the current one; at the bench
// batch (M = 24 272) early-resident dependents cost 2 % instead.  VVB200_PDL=0 / 1 forces it off / on.
void pdl_set(bool on);
bool pdl_get();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_get() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 2-D bf16 tensor map, 128B swizzle, box = {64 columns, box_rows rows}
CUtensorMap make_tmap_bf16(const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows);

// bn in {64,128,256}: 1-CTA kernel, tmA box rows 128, tmB box rows bn.
// bn == 512: CTA-pair kernel (256 x 256 tile per 2-CTA cluster, tcgen05 cta_group::2), tmA and tmB box rows 128;
//            only when gemm_pair_supported(s, e).
bool gemm_pair_supported(const GemmShape& s, const GemmEpi& e);
// tmBt: the B matrix again with 32-row boxes, for the column slices the last partial wave of tiles is cut into
//       (nullptr: whole tiles only).
void launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmBt, const GemmShape& s,
                      const GemmEpi& e, int num_sms, cudaStream_t st);
void plan_pair_tail(int total_tiles, int clusters, int* full_tiles, int* split);
void launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& s, const GemmEpi& e, int bn,
                 int num_sms, cudaStream_t st, const CUtensorMap* tmBt = nullptr);

// conv_pos_embed with the activations resident in shared memory (conv.cu).  tmA128 / tmA32: the activation matrix
// [rows, groups*64] with 128-row and 32-row boxes; tmB: weights [groups*taps*64, 64] with 64-row boxes.
bool conv_pos_supported(const GemmShape& s, const GemmEpi& e);
void launch_conv_pos(const CUtensorMap& tmA128, const CUtensorMap& tmA32, const CUtensorMap& tmB, const GemmShape& s,
                     const GemmEpi& e, int num_sms, cudaStream_t st);

// Non-causal attention over packed rows. qkv [rows, 3*dim] bf16 (q | k | v, heads of 64), out [rows, dim] bf16.
// Sequence s covers rows [seq_off[s], seq_off[s]+seq_len[s]). q-tile list: tile_seq[i], tile_q0[i] (row within seq).
struct AttnParams {
  const int32_t* seq_off;
  const int32_t* seq_len;
  const int32_t* tile_seq;
  const int32_t* tile_q0;
  int n_tiles;        // number of attn_q_tile()-row q tiles
  int heads;
  int dim;
  bf16* out;
  float scale_log2;   // (1/sqrt(64)) * log2(e)
};
int attn_q_tile();    // query rows per tile_q0 entry expected by launch_attention (128)
void launch_attention(const CUtensorMap& tmQKV, const AttnParams& p, cudaStream_t st);


// LayerNorm (no affine) + AdaLN modulation: out = LN(x)*(1+scale)+shift -> bf16.  One warp per row.
// If x2 != null computes the CFG-combined row:  (1+cfg)*h(x) - cfg*h(x2)   (used by nobody yet)
void launch_ln_mod(const float* x, int rows, int dim, const float* shift, const float* scale, float eps, bf16* out,
                   cudaStream_t st);
// LayerNorm with affine (gamma,beta), fp32 in -> bf16 and/or fp32 out.
void launch_ln_affine(const float* x, int rows, int dim, const float* g, const float* b, float eps, bf16* out_bf16,
                      float* out_f32, cudaStream_t st);

// CFG combine + Euler update: noise[r,c] += dt * (v_c + cfg*(v_c - v_u)) on valid rows; refreshes the bf16 copy
// of the noise (both CFG halves).  v is [2R, ldv] fp32 (cond rows then uncond rows).
void launch_cfg_euler(float* noise, bf16* noise_bf16, int ld_nb, const float* v, int ldv, const uint8_t* row_mask,
                      int R, int n_mel, float dt, float cfg, cudaStream_t st);

// Small dense fp32 helpers (one-off / tiny work)
void launch_linear_f32(const float* x, const float* w, const float* b, float* y, int rows, int in_f, int out_f,
                       int ldy, int act_silu, cudaStream_t st);
void launch_f32_to_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st);
void launch_f32_to_bf16_2d(const float* src, int rows, int cols, int ld_src, bf16* dst, int ld_dst, int dst_cols,
                           cudaStream_t st);

}  // namespace vv
