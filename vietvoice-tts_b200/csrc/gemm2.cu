// CTA-pair bf16 GEMM for sm_100a: a cluster of two CTAs (one TPC) computes a 256 x 256 tile with
// tcgen05.mma.cta_group::2.  Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 columns),
// so the smem fill traffic per MMA flop drops by a third against the 1-CTA 128x256 kernel in gemm.cu (48 -> 32 KB
// per 128x256x64 of work): tensor-pipe activity 58 % -> 81 % under ncu, and — the loop runs at the 1 kW power cap —
// less energy per flop (profiles/README.md).  Six 32 KB stages per CTA: the ring is latency-bound, depth matters.
//
//   warp 0   TMA producer (both CTAs; all transaction bytes are credited to the LEADER CTA's full barrier)
//   warp 1   MMA issuer (leader CTA only): M 256 x N 256 x K 16 per instruction, accumulators double-buffered in the
//            TMEM of both CTAs; tcgen05.commit multicasts the smem-slot release and the accumulator-ready signal
//   warp 2   TMEM allocation (cta_group::2, both CTAs)
//   warps 4-11  epilogue: tcgen05.ld -> fused bias / RoPE / activation / gate / residual / mask -> global.
//            Residual-update GEMMs (out-proj, ffn-down: x += gate * (acc + bias), in place) never load x at all:
//            each warp writes its 32 x 16 fp32 delta box to shared memory and a TMA reduce-add
//            (cp.reduce.async.bulk.tensor .add) lets the L2 do x += delta — one addition per element, so the result
//            is deterministic; the 8 B/element residual round trip stays between L2 and HBM.
//
// Replaces the MatMul/Gemm nodes ONNX Runtime executes for the 22 DiT blocks inside `transformer.onnx`
// (/root/reference/vietvoicetts/core/tts_engine.py:161-172).
#include "kernels.h"
#include "ptx.cuh"
#include "gemm_epi.cuh"

#include <stdio.h>
#include <stdlib.h>

#ifndef VV_GEMM_TIMING
#define VV_GEMM_TIMING 0
#endif
#if VV_GEMM_TIMING
#define GT_DECL long long gt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long gt_last = clock64()
#define GT(i) do { long long _t = clock64(); gt_acc[i] += _t - gt_last; gt_last = _t; } while (0)
#else
#define GT_DECL do { } while (0)
#define GT(i) do { } while (0)
#endif

namespace vv {

#if VV_GEMM_TIMING
__device__ unsigned long long g_gemm_timing[12];
#endif

namespace pair {
constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BN = 256;            // columns per pair tile; each CTA stages BN/2 rows of B
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = (BN / 2) * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;          // 32 KB per CTA
constexpr int EPI_WARPS = 8;
constexpr int STG_COLS = 16;
constexpr int STG_BYTES = 32 * STG_COLS * 4;            // one 32-row x 16-column fp32 box (64-byte rows, 64B swizzle)
constexpr int PAR_BYTES = EPI_WARPS * 2 * 32 * 16;      // bias + gate columns per epilogue warp
constexpr int BAR_BYTES = 512;
template <bool TMA_EPI>
struct Cfg {
  // ONE delta-box staging buffer per epilogue warp: the 16 KB a second buffer would take buy a sixth smem stage, and
  // the reduce-add GEMMs wait on TMA data, not on their epilogue (measured: out-proj +3.6 %, FFN-down +6 %)
  static constexpr int STG_BUFS = 1;
  static constexpr int STAGES = TMA_EPI ? (STG_BUFS == 1 ? 6 : 5) : 6;
  static constexpr int STG_TOTAL = TMA_EPI ? EPI_WARPS * STG_BUFS * STG_BYTES : 0;
  static constexpr int SMEM = STAGES * STAGE_BYTES + STG_TOTAL + PAR_BYTES + BAR_BYTES + 1024;
};
}  // namespace pair

// Work unit of a cluster.  Units [0, full) are whole 256 x 256 tiles; the tiles of the last, partial wave are cut into
// `split` column slices of 256 / split columns each, so that the wave's work spreads over all clusters instead of
// leaving most of them idle for a whole tile time (N = 1024: 380 tiles on 74 clusters = 5.13 waves, run as 5 waves +
// one quarter-tile wave instead of 6).  Slices of one tile are adjacent units: they run at the same time on
// neighbouring clusters and share the A rows in L2.
struct Unit {
  int m2;       // 256-row block
  int n0;       // first column
  int bn;       // columns (256, 128 or 64)
};
__device__ __forceinline__ Unit decode_unit(int u, int full, int n_tiles, int split) {
  int tile = u, sub = 0, bn = pair::BN;
  if (u >= full) {
    const int t = u - full;
    tile = full + t / split;
    sub = t % split;
    bn = pair::BN / split;
  }
  Unit r;
  r.m2 = tile / n_tiles;
  r.n0 = (tile % n_tiles) * pair::BN + sub * bn;
  r.bn = bn;
  return r;
}

template <bool TMA_EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmBt, const GemmShape s,
                 const GemmEpi e, const int tail_full, const int tail_split) {
  using namespace pair;
  using C = Cfg<TMA_EPI>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + C::STAGES * STAGE_BYTES;
  uint8_t* par_base = stg_base + C::STG_TOTAL;
  uint64_t* bars = reinterpret_cast<uint64_t*>(par_base + PAR_BYTES);
  uint64_t* full = bars;                              // leader's are used
  uint64_t* empty = bars + C::STAGES;                 // per CTA
  uint64_t* tfull = bars + 2 * C::STAGES;             // per CTA
  uint64_t* tempty = bars + 2 * C::STAGES + 2;        // leader's are used (both CTAs' epilogue warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * EPI_WARPS);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TMA_EPI) tma_prefetch_desc(&tmO);
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                    // prologue done; from here on the kernel touches what its predecessor produced

  const int m2_tiles = (s.M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = s.N / BN;
  const int total = m2_tiles * n_tiles;
  const int kiters = s.K / BK;
  const int cid = blockIdx.x >> 1;
  const int ncl = gridDim.x >> 1;
  const int units = tail_full + (total - tail_full) * tail_split;   // tail_full == total when nothing is split

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      GT_DECL;
      for (int u = cid; u < units; u += ncl) {
        const Unit un = decode_unit(u, tail_full, n_tiles, tail_split);
        const int a_row = (2 * un.m2 + (int)rank) * BM;
        const int b_row = un.n0 + (int)rank * (un.bn / 2);
        const int b_boxes = un.bn / 64;                       // 32-row boxes of B per CTA (tail slices only)
        const uint32_t tx = 2 * (A_BYTES + (un.bn / 2) * BK * 2);
        for (int kb = 0; kb < kiters; ++kb) {
          GT(6);
          mbar_wait(&empty[stage], phase ^ 1);
          GT(7);   // producer waiting for a free smem slot
          if (leader) mbar_expect_tx(&full[stage], tx);
          const uint32_t fb = mapa_u32(smem_u32(&full[stage]), 0);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          tma_load_2d_pair(sA, &tmA, kb * BK, a_row, fb);
          if (un.bn == BN) {
            tma_load_2d_pair(sA + A_BYTES, &tmB, kb * BK, b_row, fb);
          } else {
            for (int i = 0; i < b_boxes; ++i)
              tma_load_2d_pair(sA + A_BYTES + i * (32 * BK * 2), &tmBt, kb * BK, b_row + 32 * i, fb);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
#if VV_GEMM_TIMING
      GT(6);
      if (leader) for (int i = 6; i < 8; ++i) atomicAdd(&g_gemm_timing[i], (unsigned long long)gt_acc[i]);
#endif
      // tail: do not leave while the leader's tensor core may still multicast slot releases into this CTA
      for (int i = 0; i < C::STAGES; ++i) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      GT_DECL;
      for (int u = cid; u < units; u += ncl, ++it) {
        const int ubn = decode_unit(u, tail_full, n_tiles, tail_split).bn;
        const uint32_t idesc = ubn == BN ? make_idesc_bf16(2 * BM, BN)
                                         : (ubn == BN / 2 ? make_idesc_bf16(2 * BM, BN / 2) : make_idesc_bf16(2 * BM, BN / 4));
        const int acc = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        GT(0);
        mbar_wait(&tempty[acc], aphase ^ 1);
        GT(1);   // waiting for the epilogue to free the accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int kb = 0; kb < kiters; ++kb) {
          GT(0);
          mbar_wait(&full[stage], phase);
          GT(2);   // waiting for TMA data
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a0 = make_sdesc_sw128(a_addr);
          const uint64_t b0 = make_sdesc_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_ss_pair(d, a0 + 2 * k, b0 + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty[stage], 3);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull[acc], 3);
      }
#if VV_GEMM_TIMING
      GT(0);
      for (int i = 0; i < 3; ++i) atomicAdd(&g_gemm_timing[i], (unsigned long long)gt_acc[i]);
#endif
    }
  } else if (warp >= 4) {
    const int w = warp & 3;                  // TMEM lane quadrant
    const int half = (warp - 4) >> 2;        // column half of the tile
    const int ew = warp - 4;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty[0]), 0);
    const uint32_t tempty_leader1 = mapa_u32(smem_u32(&tempty[1]), 0);
    float4* sbias = reinterpret_cast<float4*>(par_base) + ew * 64;
    float4* sgate = sbias + 32;
    uint8_t* stg = stg_base + ew * C::STG_BUFS * STG_BYTES;
    const bool wide = ((reinterpret_cast<uintptr_t>(e.resid) | reinterpret_cast<uintptr_t>(e.out_f32) |
                        reinterpret_cast<uintptr_t>(e.out_bf16)) & 31) == 0 &&
                      (e.ld_resid % 8) == 0 && (e.ld_f32 % 8) == 0 && (e.ld_bf16 % 16) == 0;
    uint32_t g = 0;                          // running delta-box counter of this warp (staging buffer parity)
    int it = 0;
    GT_DECL;
    for (int u = cid; u < units; u += ncl, ++it) {
      const Unit un = decode_unit(u, tail_full, n_tiles, tail_split);
      const int m2 = un.m2;
      const int ch_per = un.bn / 64;           // 32-column chunks per epilogue warp (4; 2 or 1 in tail slices)
      const int c0 = half * ch_per;
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = (2 * m2 + (int)rank) * BM + w * 32;
      const int row = row0 + lane;
      const bool row_ok = row < s.M;
      const int nbase = un.n0;
      float4 rnext[8];
      const bool has_res = !TMA_EPI && e.resid != nullptr && row_ok;
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
      if (!TMA_EPI) load_res(c0, rnext);
      {  // stage this warp's bias / gate columns (ch_per*32 floats each)
        const int ncol = nbase + c0 * 32 + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (lane * 4 < ch_per * 32) {
          if (e.bias) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + ncol));
          if (e.gate) g4 = __ldg(reinterpret_cast<const float4*>(e.gate + ncol));
        }
        __syncwarp();
        sbias[lane] = b4;
        sgate[lane] = g4;
        __syncwarp();
      }
      GT(3);
      mbar_wait(&tfull[acc], aphase);
      GT(4);   // epilogue warp waiting for the accumulator
      tc_fence_after();
      if (TMA_EPI) {
        const int NBOX = ch_per * 32 / STG_COLS;         // delta boxes per warp per unit (8; 4 or 2 in tail slices)
#pragma unroll 1
        for (int q = 0; q < NBOX; ++q, ++g) {
          uint32_t raw[STG_COLS];
          tmem_ld16(tmem_base + (uint32_t(w * 32) << 16) + acc * BN + c0 * 32 + q * STG_COLS, raw);
          if (lane == 0) tma_store_wait_read<C::STG_BUFS - 1>();   // the reduce that last read this buffer drained it
          __syncwarp();
          tmem_ld_wait();
          if (q + 1 == NBOX) {                 // accumulator fully copied to registers: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
          }
          uint8_t* buf = stg + (g % C::STG_BUFS) * STG_BYTES;
          uint8_t* rowp = buf + lane * (STG_COLS * 4);
          const float4* sb = sbias + q * (STG_COLS / 4);
          const float4* sg = sgate + q * (STG_COLS / 4);
#pragma unroll
          for (int j = 0; j < STG_COLS / 4; ++j) {
            const float4 b = sb[j], gt = sg[j];
            float4 v = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                   __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
            fadd2(v.x, v.y, b.x, b.y);
            fadd2(v.z, v.w, b.z, b.w);
            fmul2(v.x, v.y, v.x, v.y, gt.x, gt.y);
            fmul2(v.z, v.w, v.z, v.w, gt.z, gt.w);
            *reinterpret_cast<float4*>(rowp + ((j ^ ((lane >> 1) & 3)) << 4)) = v;   // 64B swizzle
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmO, buf, nbase + c0 * 32 + q * STG_COLS, row0);
            tma_store_commit();
          }
        }
      } else {
#pragma unroll 1
        for (int c = c0; c < c0 + ch_per; ++c) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + (uint32_t(w * 32) << 16) + acc * BN + c * 32, raw);
          float4 rcur[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
          if (c + 1 < c0 + ch_per) load_res(c + 1, rnext);
          tmem_ld_wait();
          if (row_ok)
            epilogue_chunk(e, s.N, row, nbase + c * 32, raw, rcur, sbias + (c - c0) * 8, sgate + (c - c0) * 8, wide);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
      }
    }
    if (TMA_EPI && lane == 0) tma_store_wait_all<0>();
#if VV_GEMM_TIMING
    GT(3);
    if (lane == 0 && ew == 0 && leader)
      for (int i = 3; i < 6; ++i) atomicAdd(&g_gemm_timing[i], (unsigned long long)gt_acc[i]);
#endif
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs are done with each other's shared memory / TMEM
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

// fp32 row-major [rows, cols] (row stride ld elements), box = 16 columns x 32 rows, 64B swizzle
CUtensorMap make_tmap_f32_box16x32(const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems);

bool gemm_pair_supported(const GemmShape& s, const GemmEpi& e) {
  if (s.conv_taps > 0 || s.N % pair::BN != 0 || s.K % pair::BK != 0 || s.M < 1) return false;
  if (e.bias && (reinterpret_cast<uintptr_t>(e.bias) & 15)) return false;
  if (e.gate && (reinterpret_cast<uintptr_t>(e.gate) & 15)) return false;
  return true;
}

// in-place residual update with nothing else to do per element: x += gate * (acc + bias)
static bool reduce_epilogue_ok(const GemmEpi& e) {
  return e.resid && e.out_f32 == e.resid && e.ld_f32 == e.ld_resid && !e.out_bf16 && !e.row_mask &&
         e.rope_dim == 0 && e.act == ACT_NONE && (reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0 &&
         e.ld_f32 % 4 == 0;
}

// How to cut the tiles of the last, partial wave: the split (1, 2 or 4 column slices per tile) with the smallest
// estimated wave time; narrower MMAs re-read A from shared memory per column and are charged 10 % / 30 %.
// VVB200_GEMM_TAIL=0 disables the split (A/B runs).
void plan_pair_tail(int total, int clusters, int* full, int* split) {
  static const bool on = [] {
    const char* v = getenv("VVB200_GEMM_TAIL");
    return !(v && v[0] == '0');
  }();
  *full = total;
  *split = 1;
  const int rem = total % clusters;
  if (!on || rem == 0) return;
  const float penalty[3] = {1.0f, 1.1f, 1.3f};
  float best = 0.9f;                        // a split has to promise at least 10 %
  for (int i = 1; i < 3; ++i) {
    const int sp = 1 << i;
    const float cost = float((rem * sp + clusters - 1) / clusters) / float(sp) * penalty[i];
    if (cost < best - 1e-6f) {
      best = cost;
      *split = sp;
    }
  }
  if (*split > 1) *full = total - rem;
}

void launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmBt, const GemmShape& s,
                      const GemmEpi& e, int num_sms, cudaStream_t st) {
  using namespace pair;
  static DeviceOnce attr;
  attr.once([] {
    cudaFuncSetAttribute(gemm_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::SMEM);
    cudaFuncSetAttribute(gemm_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::SMEM);
  });
  const int m2_tiles = (s.M + 2 * BM - 1) / (2 * BM);
  const int total = m2_tiles * (s.N / BN);
  int clusters = num_sms / 2;
  if (clusters < 1 || total < 1) return;
  int full = total, split = 1;
  if (tmBt) plan_pair_tail(total, clusters, &full, &split);
  const int units = full + (total - full) * split;
  if (clusters > units) clusters = units;
  const CUtensorMap& tBt = tmBt ? *tmBt : tmB;
  if (reduce_epilogue_ok(e)) {
    const CUtensorMap tmO = make_tmap_f32_box16x32(e.out_f32, s.M, s.N, e.ld_f32);
    launch_k(gemm_pair_kernel<true>, 2 * clusters, 384, Cfg<true>::SMEM, st, tmA, tmB, tmO, tBt, s, e, full, split);
  } else {
    launch_k(gemm_pair_kernel<false>, 2 * clusters, 384, Cfg<false>::SMEM, st, tmA, tmB, tmA, tBt, s, e, full, split);
  }
}

#if VV_GEMM_TIMING
extern "C" void vv_gemm_timing_dump() {
  unsigned long long h[12];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_gemm_timing, sizeof(h));
  const char* names[8] = {"mma: issue/other", "mma: wait tempty (epilogue)", "mma: wait full (TMA)",
                          "epi: work", "epi: wait tfull (MMA)", "epi: (unused)",
                          "tma: issue/other", "tma: wait empty slot"};
  const double tm = double(h[0] + h[1] + h[2]), te = double(h[3] + h[4] + h[5]), tp = double(h[6] + h[7]);
  for (int i = 0; i < 8; ++i)
    printf("  %-30s %14llu  %5.1f%%\n", names[i], h[i],
           100.0 * double(h[i]) / ((i < 3 ? tm : (i < 6 ? te : tp)) + 1e-9));
  unsigned long long z[12] = {0};
  cudaMemcpyToSymbol(g_gemm_timing, z, sizeof(z));
}
#endif

}  // namespace vv
