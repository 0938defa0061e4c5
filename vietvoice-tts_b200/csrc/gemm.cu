// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (accumulators in TMEM,
// double-buffered) -> tcgen05.ld epilogue with the DiT's fused element-wise work (bias, RoPE, GELU/Mish,
// AdaLN gate, residual, row mask).  Also runs the grouped conv_pos_embed as an implicit GEMM (one k-iteration
// per tap, activation tile re-fetched at a shifted row coordinate; TMA zero-fills the sequence padding).
//
// Replaces the MatMul/Gemm/Conv nodes ONNX Runtime executes inside `transformer.onnx`
// (/root/reference/vietvoicetts/core/tts_engine.py:161-172).
#include "kernels.h"
#include "ptx.cuh"
#include "gemm_epi.cuh"

#include <mutex>
#include <stdio.h>

namespace vv {

static constexpr int BM = 128;
static constexpr int BK = 64;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulators
  static constexpr int BAR_BYTES = 256;
  static constexpr int PAR_BYTES = 8 * 2 * 32 * 16;  // per epilogue warp: 32 float4 of bias + 32 float4 of gate
  static constexpr int SMEM = STAGES * STAGE_BYTES + BAR_BYTES + PAR_BYTES + 1024;
};

template <int BN, bool CONV>
__global__ void __launch_bounds__(384, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape s,
            const GemmEpi e) {
  using C = GemmCfg<BN>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                    // prologue done; from here on the kernel touches what its predecessor produced

  const int m_tiles = (s.M + BM - 1) / BM;
  const int n_tiles = CONV ? s.conv_groups : (s.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int kiters = CONV ? s.conv_taps : s.K / BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < kiters; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          uint8_t* sA = smem + stage * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          if (CONV) {
            tma_load_2d(sA, &tmA, n_blk * 64, m_blk * BM + kb - s.conv_taps / 2, &full[stage]);
            tma_load_2d(sB, &tmB, 0, (n_blk * s.conv_taps + kb) * 64, &full[stage]);
          } else {
            tma_load_2d(sA, &tmA, kb * BK, m_blk * BM, &full[stage]);
            tma_load_2d(sB, &tmB, kb * BK, n_blk * BN, &full[stage]);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], aphase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int kb = 0; kb < kiters; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a0 = make_sdesc_sw128(a_addr);
          const uint64_t b0 = make_sdesc_sw128(a_addr + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_ss(d, a0 + 2 * k, b0 + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 4) {
    // 8 epilogue warps: warp%4 selects the TMEM lane quadrant (rows), (warp-4)/4 selects the column half
    const int w = warp & 3;
    const int half = (warp - 4) >> 2;
    const bool wide = ((reinterpret_cast<uintptr_t>(e.resid) | reinterpret_cast<uintptr_t>(e.out_f32) |
                        reinterpret_cast<uintptr_t>(e.out_bf16)) & 31) == 0 &&
                      (e.ld_resid % 8) == 0 && (e.ld_f32 % 8) == 0 && (e.ld_bf16 % 16) == 0;
    constexpr int CH = BN / 32;                      // 32-column chunks per tile
    constexpr int CH_PER = CH >= 2 ? CH / 2 : 1;     // chunks per epilogue warp
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row = m_blk * BM + w * 32 + lane;
      const bool row_ok = row < s.M;
      const int nbase = CONV ? n_blk * 64 : n_blk * BN;
      const int c0 = half * CH_PER;
      // residual prefetch for the first chunk: independent of the accumulator, overlaps the wait for the MMA
      float4 rnext[8];
      const bool has_res = e.resid != nullptr && row_ok;
      auto load_res = [&](int c, float4 (&r)[8]) {
        const int n0 = nbase + c * 32;
        if (has_res && n0 + 32 <= s.N) {
          const float* rp = e.resid + (size_t)row * e.ld_resid + n0;
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
      if (CH >= 2 || half == 0) load_res(c0, rnext);
      // stage this warp's bias / gate columns (CH_PER*32 floats each) in shared memory
      float4* sbias = reinterpret_cast<float4*>(smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES) + (warp - 4) * 64;
      float4* sgate = sbias + 32;
      {
        const int ncol = nbase + c0 * 32 + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (lane < CH_PER * 8 && ncol + 4 <= s.N) {
          if (e.bias) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + ncol));
          if (e.gate) g4 = __ldg(reinterpret_cast<const float4*>(e.gate + ncol));
        }
        __syncwarp();
        sbias[lane] = b4;
        sgate[lane] = g4;
        __syncwarp();
      }
      mbar_wait(&tfull[acc], aphase);
      tc_fence_after();
      if (CH >= 2 || half == 0) {
#pragma unroll 1
        for (int c = c0; c < c0 + CH_PER; ++c) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + (uint32_t(w * 32) << 16) + acc * BN + c * 32, raw);
          float4 rcur[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
          if (c + 1 < c0 + CH_PER) load_res(c + 1, rnext);
          tmem_ld_wait();
          if (row_ok) epilogue_chunk(e, s.N, row, nbase + c * 32, raw, rcur, sbias + (c - c0) * 8, sgate + (c - c0) * 8, wide);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

CUtensorMap make_tmap_bf16(const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    fprintf(stderr, "vvb200: cuTensorMapEncodeTiled entry point not available\n");
    abort();
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "vvb200: cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box_rows=%u base=%p\n",
            (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, base);
    abort();
  }
  return m;
}

CUtensorMap make_tmap_f32_box16x32(const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems) {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    fprintf(stderr, "vvb200: cuTensorMapEncodeTiled entry point not available\n");
    abort();
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 4};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "vvb200: cuTensorMapEncodeTiled(f32) failed (%d) rows=%llu cols=%llu ld=%llu base=%p\n", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, base);
    abort();
  }
  return m;
}

template <int BN, bool CONV>
static void launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& s, const GemmEpi& e,
                     int num_sms, cudaStream_t st) {
  using C = GemmCfg<BN>;
  static DeviceOnce attr;
  attr.once([] { cudaFuncSetAttribute(gemm_kernel<BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM); });
  const int m_tiles = (s.M + BM - 1) / BM;
  const int n_tiles = CONV ? s.conv_groups : (s.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  if (grid > num_sms) grid = num_sms;
  if (grid < 1) return;
  launch_k(gemm_kernel<BN, CONV>, grid, 384, C::SMEM, st, tmA, tmB, s, e);
}

void launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& s, const GemmEpi& e, int bn,
                 int num_sms, cudaStream_t st, const CUtensorMap* tmBt) {
  if (s.conv_taps > 0) {
    launch_t<64, true>(tmA, tmB, s, e, num_sms, st);
  } else if (bn == 512) {
    launch_gemm_pair(tmA, tmB, tmBt, s, e, num_sms, st);
  } else if (bn == 256) {
    launch_t<256, false>(tmA, tmB, s, e, num_sms, st);
  } else if (bn == 128) {
    launch_t<128, false>(tmA, tmB, s, e, num_sms, st);
  } else {
    launch_t<64, false>(tmA, tmB, s, e, num_sms, st);
  }
}

}  // namespace vv
