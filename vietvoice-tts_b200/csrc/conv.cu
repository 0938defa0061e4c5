// conv_pos_embed (grouped Conv1d, 64 channels per group, `taps` taps, zero padding) as an implicit GEMM that keeps
// the activations RESIDENT in shared memory.
//
// The generic kernel (gemm.cu, CONV mode) runs one k-iteration per tap and re-fetches the 128-row activation tile at a
// row coordinate shifted by one for every tap: 24 KB of TMA traffic per 128 x 64 x 64 MMA block, three times what the
// L2 -> SM path delivers in the 128 cycles the block takes, so the tensor pipe idles two thirds of the time.
// Here a CTA owns 256 output rows x one group (64 channels):
//   * the (256 + taps - 1)-row halo of the group's 64 input channels is loaded ONCE per tile (36 KB, SW128 rows of
//     128 B, double-buffered across tiles);
//   * tap t reads it through a shared-memory matrix descriptor whose start address is simply advanced by t rows
//     (t x 128 B).  The tensor core applies the 128-byte swizzle from the ABSOLUTE shared-memory address bits — the
//     same bits TMA used when it wrote the rows — so a start address that is not 1024-byte aligned needs nothing else
//     (measured: with the descriptor's base-offset field set to (start >> 7) & 7 the results are wrong, with 0 they
//     are exact for every tap phase, tests/test_kernels_gpu.py::test_conv_rows_grouped).  The sliding window costs no
//     data movement at all;
//   * only the 64 x 64 tap weights (8 KB, L2-resident: 4 MB for all groups) stream through a 12-stage TMA ring, shared
//     by the two 128-row halves of the tile (8 MMAs of M128 N64 K16 per stage = 256 tensor cycles per 8 KB).
//   Accumulators: two 128-lane x 64-column blocks per tile, double-buffered (256 TMEM columns).  The sequence padding
//   is the 16-row zero gap between sequences plus TMA's out-of-bounds zero fill, as in the generic kernel.
//
// Replaces the two grouped Conv nodes (+ Mish) of the DiT's ConvPositionEmbedding inside `transformer.onnx`
// (/root/reference/vietvoicetts/core/tts_engine.py:161-172).
#include "kernels.h"
#include "ptx.cuh"
#include "gemm_epi.cuh"

#include <stdlib.h>

namespace vv {

namespace convk {
constexpr int ROWS = 256;                       // output rows per tile (two M = 128 halves)
constexpr int HALO_ROWS = 288;                  // 2 x 128-row boxes + one 32-row box >= 256 + taps - 1 (taps <= 33)
constexpr int A_BYTES = HALO_ROWS * 128;        // 36 KB
constexpr int B_BYTES = 64 * 128;               // one tap: 64 output channels x 64 input channels
constexpr int B_STAGES = 12;
constexpr int PAR_BYTES = 8 * 2 * 32 * 16;      // bias + gate staging per epilogue warp (as in gemm.cu)
constexpr int BAR_BYTES = 512;
constexpr int SMEM = 2 * A_BYTES + B_STAGES * B_BYTES + PAR_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 256;                  // 2 accumulator sets x 2 halves x 64 columns
}  // namespace convk

// descriptor for a K-major SW128 operand whose first row sits `row` rows into a 1024-byte aligned buffer
__device__ __forceinline__ uint64_t make_sdesc_sw128_row(uint32_t base_addr, int row) {
  return make_sdesc_sw128(base_addr + row * 128);
}

__global__ void __launch_bounds__(384, 1)
conv_pos_kernel(const __grid_constant__ CUtensorMap tmA128, const __grid_constant__ CUtensorMap tmA32,
                const __grid_constant__ CUtensorMap tmB, const GemmShape s, const GemmEpi e) {
  using namespace convk;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // 2 halo buffers
  uint8_t* sB = smem + 2 * A_BYTES;                     // tap ring
  uint8_t* par_base = sB + B_STAGES * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(par_base + PAR_BYTES);
  uint64_t* a_full = bars;                 // 2
  uint64_t* a_empty = bars + 2;            // 2
  uint64_t* b_full = bars + 4;             // B_STAGES
  uint64_t* b_empty = b_full + B_STAGES;   // B_STAGES
  uint64_t* tfull = b_empty + B_STAGES;    // 2
  uint64_t* tempty = tfull + 2;            // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    for (int i = 0; i < B_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA128);
    tma_prefetch_desc(&tmA32);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                    // prologue done; from here on the kernel touches what its predecessor produced

  const int taps = s.conv_taps;
  const int groups = s.conv_groups;
  const int m_tiles = (s.M + ROWS - 1) / ROWS;
  const int total = m_tiles * groups;       // group fastest: neighbouring CTAs share the halo rows in L2

  if (warp == 0) {
    if (lane == 0) {
      int st = 0, it = 0;
      uint32_t bph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int m_blk = tile / groups, g = tile % groups;
        const int ab = it & 1;
        mbar_wait(&a_empty[ab], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&a_full[ab], A_BYTES);
        uint8_t* a = sA + ab * A_BYTES;
        const int r0 = m_blk * ROWS - taps / 2;           // may be negative / past the end: TMA zero-fills
        tma_load_2d(a, &tmA128, g * 64, r0, &a_full[ab]);
        tma_load_2d(a + 128 * 128, &tmA128, g * 64, r0 + 128, &a_full[ab]);
        tma_load_2d(a + 256 * 128, &tmA32, g * 64, r0 + 256, &a_full[ab]);
        for (int t = 0; t < taps; ++t) {
          mbar_wait(&b_empty[st], bph ^ 1);
          mbar_expect_tx(&b_full[st], B_BYTES);
          tma_load_2d(sB + st * B_BYTES, &tmB, 0, (g * taps + t) * 64, &b_full[st]);
          if (++st == B_STAGES) { st = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64);
      int st = 0, it = 0;
      uint32_t bph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(&tempty[ab], par ^ 1);
        mbar_wait(&a_full[ab], par);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + ab * A_BYTES);
        const uint32_t d = tmem_base + ab * 128;
        for (int t = 0; t < taps; ++t) {
          mbar_wait(&b_full[st], bph);
          tc_fence_after();
          const uint64_t b0 = make_sdesc_sw128(smem_u32(sB + st * B_BYTES));
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint64_t a0 = make_sdesc_sw128_row(a_addr, u * 128 + t);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(d + u * 64, a0 + 2 * k, b0 + 2 * k, idesc, (t | k) != 0);
          }
          umma_commit(&b_empty[st]);
          if (++st == B_STAGES) { st = 0; bph ^= 1; }
        }
        umma_commit(&a_empty[ab]);
        umma_commit(&tfull[ab]);
      }
    }
  } else if (warp >= 4) {
    // 8 epilogue warps: warp % 4 = TMEM lane quadrant (32 rows), (warp - 4) / 4 = which 128-row half of the tile
    const int w = warp & 3;
    const int u = (warp - 4) >> 2;
    const bool wide = ((reinterpret_cast<uintptr_t>(e.resid) | reinterpret_cast<uintptr_t>(e.out_f32) |
                        reinterpret_cast<uintptr_t>(e.out_bf16)) & 31) == 0 &&
                      (e.ld_resid % 8) == 0 && (e.ld_f32 % 8) == 0 && (e.ld_bf16 % 16) == 0;
    float4* sbias = reinterpret_cast<float4*>(par_base) + (warp - 4) * 64;
    float4* sgate = sbias + 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int m_blk = tile / groups, g = tile % groups;
      const int ab = it & 1;
      const uint32_t par = (it >> 1) & 1;
      const int row = m_blk * ROWS + u * 128 + w * 32 + lane;
      const bool row_ok = row < s.M;
      const int nbase = g * 64;
      const bool has_res = e.resid != nullptr && row_ok;
      float4 rnext[8];
      auto load_res = [&](int c, float4 (&r)[8]) {
        if (has_res) {
          const float* rp = e.resid + (size_t)row * e.ld_resid + nbase + c * 32;
          if (wide) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ldg256_stream(rp + 8 * j, r[2 * j], r[2 * j + 1]);
          } else {
            const float4* r4 = reinterpret_cast<const float4*>(rp);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = r4[j];
          }
        }
      };
      load_res(0, rnext);
      {
        const int ncol = nbase + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (lane < 16) {
          if (e.bias) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + ncol));
          if (e.gate) g4 = __ldg(reinterpret_cast<const float4*>(e.gate + ncol));
        }
        __syncwarp();
        sbias[lane] = b4;
        sgate[lane] = g4;
        __syncwarp();
      }
      mbar_wait(&tfull[ab], par);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + (uint32_t(w * 32) << 16) + ab * 128 + u * 64 + c * 32, raw);
        float4 rcur[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
        if (c == 0) load_res(1, rnext);
        tmem_ld_wait();
        if (row_ok) epilogue_chunk(e, s.N, row, nbase + c * 32, raw, rcur, sbias + c * 8, sgate + c * 8, wide);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

bool conv_pos_supported(const GemmShape& s, const GemmEpi& e) {
  static const bool on = [] {   // VVB200_CONV_RESIDENT=0: the generic one-tap-per-k-iteration kernel (A/B runs)
    const char* v = getenv("VVB200_CONV_RESIDENT");
    return !(v && v[0] == '0');
  }();
  if (!on || s.conv_taps <= 0 || s.conv_taps > convk::HALO_ROWS - convk::ROWS + 1 || s.conv_groups <= 0) return false;
  if (e.bias && (reinterpret_cast<uintptr_t>(e.bias) & 15)) return false;
  if (e.gate && (reinterpret_cast<uintptr_t>(e.gate) & 15)) return false;
  return s.N == s.conv_groups * 64;
}

void launch_conv_pos(const CUtensorMap& tmA128, const CUtensorMap& tmA32, const CUtensorMap& tmB, const GemmShape& s,
                     const GemmEpi& e, int num_sms, cudaStream_t st) {
  using namespace convk;
  static DeviceOnce attr;
  attr.once([] { cudaFuncSetAttribute(conv_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); });
  int grid = ((s.M + ROWS - 1) / ROWS) * s.conv_groups;
  if (grid > num_sms) grid = num_sms;
  if (grid < 1) return;
  launch_k(conv_pos_kernel, grid, 384, SMEM, st, tmA128, tmA32, tmB, s, e);
}

}  // namespace vv
