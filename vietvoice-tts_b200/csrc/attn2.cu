// Non-causal flash attention for sm_100a (d_h = 64), generation 2: the exponentials of one score row are SPLIT BY
// PIPE between two threads.
//
// attn.cu (generation 1) is bound by MUFU.EX2 (16/clk/SM: 1024 cycles per 128x128 score tile against 512 tensor
// cycles) and reaches ~70 % of that pipe: with one softmax warp per scheduler and CTA, a scheduler's MUFU idles whenever
// its two warps (two CTAs per SM) are both outside their exp2 loops, and a warp alone cannot saturate it.  Moving
// exponentials to the FMA pipe (degree-3 polynomial, attn_softmax.cuh) helped only up to one pair in eight, because the
// twelve polynomial instructions sit in the SAME warp's instruction stream as its MUFU work.
//
// Here a 128-row query tile has TWO softmax warpgroups:
//   main   (warps 0-3, thread = row): columns [0, MAIN) of every 128-key score tile, exp2 on MUFU (optionally one pair
//          in eight on the FMA pipe, as before);
//   assist (warps 4-7, thread = row): columns [MAIN, 128), exp2 ONLY by the FMA/ALU-pipe polynomial — it never
//          touches MUFU, so it never competes with the main warps for it; it has its own issue slots and registers.
// With two CTAs per SM every scheduler now holds four softmax warps (two of each kind): MUFU work per tile drops to
// MAIN/128 of a row, and the gaps of one warp (S readout from TMEM, barrier round trips, P store) are covered by three
// others instead of one.  Both threads of a row use the same reference m_ref (fixed by the first kv tile, as in
// generation 1: scaling by a power of two is exact), write their halves of P to the same TMEM columns the PV MMA
// reads, and keep their own partial row sums, combined once per CTA.
//
// Overflow: a row whose scores outgrow the first-tile reference by more than 2^80 cannot be repaired per warp any more
// (the two threads of a row would have to agree on a new reference every tile).  Instead every thread keeps a sticky
// flag (one compare per tile on the row sum it computes anyway); the CTA votes after its last tile and, only if some row
// tripped it, redoes the whole item in two passes: an exact row maximum (QK^T only), then the same loop against that
// exact maximum, which cannot overflow.  The fallback costs that CTA 2.5x; it needs scores 55 nats above the first
// tile's maximum (tests/test_kernels_gpu.py::test_attention_reference_outgrown_by_overflow exercises it).
//
//   warps 0-3  main softmax      144 registers
//   warps 4-7  assist softmax     64 registers
//   warp  8    TMA producer       32 registers   (Q once, K / V double-buffered 128-key tiles)
//   warp  9    MMA issuer + TMEM allocation
//   S = Q K^T : tcgen05.mma M128 N128 K64 -> TMEM cols [0,128);  P (bf16) -> cols [128,192);  O += P V -> cols [192,256)
//
// Replaces the attention sub-graph of `transformer.onnx` (/root/reference/vietvoicetts/core/tts_engine.py:161-172).
#include "kernels.h"
#include "ptx.cuh"
#include "attn_softmax.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef VV_ATTN_TIMING
#define VV_ATTN_TIMING 0
#endif
#if VV_ATTN_TIMING
#define T2(i) do { long long _t = clock64(); tacc[i] += _t - tlast; tlast = _t; } while (0)
#else
#define T2(i) do { } while (0)
#endif

namespace vv {

#if VV_ATTN_TIMING
__device__ unsigned long long g_attn2_timing[2][12];
#endif

namespace attn2 {
constexpr int K_STAGES = 2;
constexpr int V_STAGES = 2;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int Q_OFF = 0;
constexpr int K_OFF = Q_OFF + TILE_BYTES;
constexpr int V_OFF = K_OFF + K_STAGES * TILE_BYTES;
constexpr int BAR_OFF = V_OFF + V_STAGES * TILE_BYTES;
constexpr int XCH_OFF = BAR_OFF + 256;                 // float xch[2][128]: row values handed between the two threads
constexpr int SMEM = XCH_OFF + 1024 + 1024;
constexpr int THREADS = 384;
constexpr uint32_t TM_S = 0;
constexpr uint32_t TM_P = 128;
constexpr uint32_t TM_O = 192;
constexpr uint32_t TM_COLS = 256;
constexpr float SUM_LIMIT = 1.2089258196146292e24f;    // 2^80: row sum of one kv tile against the reference
// registers per thread after setmaxnreg; every split sums to 240 = 3 x 80 (the launch allocation of 384 x 80)
constexpr int AUX_REGS = 32;
constexpr int MAIN_REGS_FMA_ASSIST = 144, ASSIST_REGS_FMA = 64;      // assist = FMA-pipe polynomial only
constexpr int MAIN_REGS_SYM = 104, ASSIST_REGS_SYM = 104;            // both groups on MUFU (two threads per row)
}  // namespace attn2

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// exp2 of `N` consecutive scores (N = 32 or 16) against reference m -> N/2 packed bf16 pairs; adds to the two packed
// row-sum chains.  POLY8: of every 8 score pairs this many go to the FMA pipe (8 = all of them: the assist warps).
// MASKED: columns >= valid count as -inf.
template <bool MASKED, int N, int POLY8>
__device__ __forceinline__ void exp_group(const uint32_t* s, int col0, int valid, float scale_log2, float m,
                                          uint32_t* pk, float& sa0, float& sa1, float& sb0, float& sb1) {
  auto val = [&](int i) { return (MASKED && col0 + i >= valid) ? __uint_as_float(0xff800000u) : __uint_as_float(s[i]); };
  if (MASKED && col0 >= valid) {                 // whole group past the end of the sequence (warp-uniform)
#pragma unroll
    for (int i = 0; i < N / 2; ++i) pk[i] = 0u;
    return;
  }
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    float x0, x1, x2, x3, e0, e1, e2, e3;
    ffma2(x0, x1, val(i), val(i + 1), scale_log2, -m);
    ffma2(x2, x3, val(i + 2), val(i + 3), scale_log2, -m);
    if (((i / 2) % 8) >= 8 - POLY8) {
      poly_exp2_pair(x0, x1, e0, e1);
    } else {
      e0 = fast_exp2(x0);
      e1 = fast_exp2(x1);
    }
    if (((i / 2 + 1) % 8) >= 8 - POLY8) {
      poly_exp2_pair(x2, x3, e2, e3);
    } else {
      e2 = fast_exp2(x2);
      e3 = fast_exp2(x3);
    }
    if (MASKED && POLY8 > 0) {      // the polynomial clamps -inf to 2^-125 instead of returning 0 (tail tile only)
      if (col0 + i >= valid) e0 = 0.f;
      if (col0 + i + 1 >= valid) e1 = 0.f;
      if (col0 + i + 2 >= valid) e2 = 0.f;
      if (col0 + i + 3 >= valid) e3 = 0.f;
    }
    fadd2(sa0, sa1, e0, e1);
    fadd2(sb0, sb1, e2, e3);
    pk[i / 2] = pack_bf16(e0, e1);
    pk[i / 2 + 1] = pack_bf16(e2, e3);
  }
}

// maximum of the first `COLS` registers of s (columns col0 .. col0+COLS of the tile), masked past `valid`
template <int COLS>
__device__ __forceinline__ float row_max_cols(const uint32_t (&s)[COLS], int col0, int valid, bool partial) {
  float mx = __uint_as_float(0xff800000u), my = mx, mz = mx, mw = mx;
#pragma unroll
  for (int i = 0; i < COLS; i += 4) {
    float a = __uint_as_float(s[i]), b = __uint_as_float(s[i + 1]), c = __uint_as_float(s[i + 2]),
          d = __uint_as_float(s[i + 3]);
    if (partial) {
      if (col0 + i >= valid) a = __uint_as_float(0xff800000u);
      if (col0 + i + 1 >= valid) b = __uint_as_float(0xff800000u);
      if (col0 + i + 2 >= valid) c = __uint_as_float(0xff800000u);
      if (col0 + i + 3 >= valid) d = __uint_as_float(0xff800000u);
    }
    mx = fmaxf(mx, a); my = fmaxf(my, b); mz = fmaxf(mz, c); mw = fmaxf(mw, d);
  }
  return fmaxf(fmaxf(mx, my), fmaxf(mz, mw));
}

template <int COLS>
__device__ __forceinline__ void load_cols(uint32_t taddr, uint32_t (&s)[COLS]) {
  static_assert(COLS % 16 == 0, "column split must be a multiple of 16");
#pragma unroll
  for (int c = 0; c + 32 <= COLS; c += 32) tmem_ld32(taddr + c, *reinterpret_cast<uint32_t(*)[32]>(&s[c]));
  if (COLS % 32) tmem_ld16(taddr + (COLS / 32) * 32, *reinterpret_cast<uint32_t(*)[16]>(&s[(COLS / 32) * 32]));
  tmem_ld_wait();
}

// pass kinds
enum { PASS_OPT = 0, PASS_MAX = 1, PASS_EXACT = 2 };

// One softmax warpgroup (IS_MAIN: columns [0, MAIN) on MUFU; else columns [MAIN, 128) on the FMA pipe).  A separate
// instantiation per group: ptxas sizes the register allocation of a region by the setmaxnreg that dominates it, so the
// two groups must not share code after their setmaxnreg.
template <int MAIN, int POLY8, int APOLY, bool IS_MAIN>
__device__ __forceinline__ void softmax_role(const AttnParams& p, uint32_t tmem_base, int warp, int q0, int seq_row0,
                                             int kv_len, int n_kv, int head, uint64_t* s_full, uint64_t* s_free,
                                             uint64_t* p_full, uint64_t* pv_done, float* xch0, float* xch1) {
  using namespace attn2;
  constexpr int ASSIST = 128 - MAIN;
  {
    // ------------------------------------------------------------------ the two softmax warpgroups: thread = row
    const int r = threadIdx.x & 127;           // row within tile == TMEM lane
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t ts = tmem_base + lane_base + TM_S;
    const uint32_t tp = tmem_base + lane_base + TM_P;
    const uint32_t to = tmem_base + lane_base + TM_O;
    float m_ref = 0.0f, l = 0.0f;
    bool bad = false;
    int it = 0, pt = 0;
#if VV_ATTN_TIMING
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif

    // one pass over the kv tiles.  kind: PASS_OPT (reference from the first tile, overflow detection), PASS_MAX (row
    // maximum only), PASS_EXACT (reference = exact maximum).  Written once for both groups: `IS_MAIN` is warp-uniform
    // and the column counts are compile-time in each branch.
    auto run_pass = [&](const int kind) {
      float mx_run = __uint_as_float(0xff800000u);
      bool s_ready = false;
      l = 0.0f;
      for (int j = 0; j < n_kv; ++j, ++it) {
        T2(7);     // loop overhead / previous arrive
        if (!s_ready) mbar_wait(s_full, it & 1);
        s_ready = false;
        T2(0);     // wait S
        tc_fence_after();
        const int kv_valid = kv_len - j * 128;
        const bool partial = kv_valid < 128;
        const bool pv_wait = kind != PASS_MAX && j > 0;
        uint64_t* nb = (kind != PASS_MAX && j + 1 < n_kv) ? s_full : nullptr;
        if (IS_MAIN) {
          uint32_t s[MAIN];
          load_cols<MAIN>(ts, s);
          tc_fence_before();
          mbar_arrive(s_free);
          T2(1);   // S readout
          if (kind == PASS_MAX) {
            mx_run = fmaxf(mx_run, row_max_cols<MAIN>(s, 0, kv_valid, partial));
            continue;
          }
          if (kind == PASS_OPT && j == 0) {      // reference = maximum of this thread's columns of the first tile
            m_ref = row_max_cols<MAIN>(s, 0, kv_valid, partial) * p.scale_log2;
            xch0[r] = m_ref;
            named_bar_arrive(1, 256);            // the assist thread of this row picks it up
          }
          T2(2);   // first-tile maximum
          float sa0 = 0.f, sa1 = 0.f, sb0 = 0.f, sb1 = 0.f;
          constexpr int G = MAIN / 32, R16 = (MAIN % 32) / 16;
          uint32_t pk[G][16];
          uint32_t pk8[8];
          const bool pv_ready = pv_wait ? mbar_test(pv_done, (pt - 1) & 1) : true;
#pragma unroll
          for (int c = 0; c < G; ++c) {
            if (c == (G >= 3 ? 2 : G - 1) && nb) s_ready = mbar_test(nb, (it + 1) & 1);
            if (partial) exp_group<true, 32, POLY8>(&s[c * 32], c * 32, kv_valid, p.scale_log2, m_ref, pk[c], sa0, sa1, sb0, sb1);
            else exp_group<false, 32, POLY8>(&s[c * 32], c * 32, kv_valid, p.scale_log2, m_ref, pk[c], sa0, sa1, sb0, sb1);
            if (c == 1) {
              T2(3);   // exp2 groups 0-1
              if (pv_wait) {
                if (!pv_ready) mbar_wait(pv_done, (pt - 1) & 1);    // P buffer free again, O quiescent
                tc_fence_after();
              }
              T2(4);   // wait PV(j-1)
              tmem_st16(tp, pk[0]);
              tmem_st16(tp + 16, pk[1]);
            } else if (c > 1) {
              tmem_st16(tp + c * 16, pk[c]);
            }
          }
          if (R16) {
            if (partial) exp_group<true, 16, POLY8>(&s[G * 32], G * 32, kv_valid, p.scale_log2, m_ref, pk8, sa0, sa1, sb0, sb1);
            else exp_group<false, 16, POLY8>(&s[G * 32], G * 32, kv_valid, p.scale_log2, m_ref, pk8, sa0, sa1, sb0, sb1);
            tmem_st8(tp + G * 16, pk8);
          }
          const float sum = (sa0 + sa1) + (sb0 + sb1);
          bad |= !(sum < SUM_LIMIT);
          l += sum;
          T2(5);   // remaining exp2 groups + P stores
        } else {
          uint32_t s[ASSIST];
          load_cols<ASSIST>(ts + MAIN, s);
          tc_fence_before();
          mbar_arrive(s_free);
          T2(1);
          if (kind == PASS_MAX) {
            mx_run = fmaxf(mx_run, row_max_cols<ASSIST>(s, MAIN, kv_valid, partial));
            continue;
          }
          if (kind == PASS_OPT && j == 0) {
            named_bar_sync(1, 256);
            m_ref = xch0[r];
          }
          T2(2);
          float sa0 = 0.f, sa1 = 0.f, sb0 = 0.f, sb1 = 0.f;
          uint32_t pk[ASSIST / 2];
          const bool pv_ready = pv_wait ? mbar_test(pv_done, (pt - 1) & 1) : true;
          if (nb) s_ready = mbar_test(nb, (it + 1) & 1);
          if (partial) exp_group<true, ASSIST, APOLY>(s, MAIN, kv_valid, p.scale_log2, m_ref, pk, sa0, sa1, sb0, sb1);
          else exp_group<false, ASSIST, APOLY>(s, MAIN, kv_valid, p.scale_log2, m_ref, pk, sa0, sa1, sb0, sb1);
          T2(3);   // polynomial exp2
          if (pv_wait) {
            if (!pv_ready) mbar_wait(pv_done, (pt - 1) & 1);
            tc_fence_after();
          }
          T2(4);   // wait PV(j-1)
#pragma unroll
          for (int g = 0; g < ASSIST / 32; ++g)
            tmem_st16(tp + MAIN / 2 + g * 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[g * 16]));
          if (ASSIST % 32) tmem_st8(tp + MAIN / 2 + (ASSIST / 32) * 16, *reinterpret_cast<uint32_t(*)[8]>(&pk[(ASSIST / 32) * 16]));
          const float sum = (sa0 + sa1) + (sb0 + sb1);
          bad |= !(sum < SUM_LIMIT);
          l += sum;
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
        ++pt;
        T2(6);     // P retire + arrive
      }
      if (kind == PASS_MAX) {
        // exact row maximum = max over both threads of the row
        (IS_MAIN ? xch0 : xch1)[r] = mx_run;
        named_bar_sync(1, 256);
        m_ref = fmaxf(xch0[r], xch1[r]) * p.scale_log2;
        named_bar_sync(2, 256);                // both have read before anybody writes xch again
      } else {
        mbar_wait(pv_done, (pt - 1) & 1);      // all MMAs of the pass retired: TMEM quiescent, O complete
        tc_fence_after();
      }
    };

    run_pass(PASS_OPT);
    if (__syncthreads_or(bad ? 1 : 0)) {
      run_pass(PASS_MAX);
      run_pass(PASS_EXACT);
    }
#if VV_ATTN_TIMING
    tacc[11] = n_kv;
    if ((threadIdx.x & 31) == 0 && blockIdx.x % 97 == 0)
      for (int i = 0; i < 12; ++i) atomicAdd(&g_attn2_timing[IS_MAIN ? 0 : 1][i], (unsigned long long)tacc[i]);
#endif
    // ---- finalize: the assist hands over its partial row sum, the main thread writes O / l as bf16
    if (!IS_MAIN) {
      xch1[r] = l;
      named_bar_arrive(3, 256);
    } else {
      named_bar_sync(3, 256);
      l += xch1[r];
      const int qrow = q0 + r;
      const float inv = 1.0f / l;
      bf16* orow = p.out + (size_t)(seq_row0 + qrow) * p.dim + head * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(to + c * 32, o);
        tmem_ld_wait();
        if (qrow < kv_len) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {   // 2 x 256-bit stores: whole 32-byte sectors per instruction
            uint4 u0, u1;
            u0.x = pack_bf16(__uint_as_float(o[16 * g]) * inv, __uint_as_float(o[16 * g + 1]) * inv);
            u0.y = pack_bf16(__uint_as_float(o[16 * g + 2]) * inv, __uint_as_float(o[16 * g + 3]) * inv);
            u0.z = pack_bf16(__uint_as_float(o[16 * g + 4]) * inv, __uint_as_float(o[16 * g + 5]) * inv);
            u0.w = pack_bf16(__uint_as_float(o[16 * g + 6]) * inv, __uint_as_float(o[16 * g + 7]) * inv);
            u1.x = pack_bf16(__uint_as_float(o[16 * g + 8]) * inv, __uint_as_float(o[16 * g + 9]) * inv);
            u1.y = pack_bf16(__uint_as_float(o[16 * g + 10]) * inv, __uint_as_float(o[16 * g + 11]) * inv);
            u1.z = pack_bf16(__uint_as_float(o[16 * g + 12]) * inv, __uint_as_float(o[16 * g + 13]) * inv);
            u1.w = pack_bf16(__uint_as_float(o[16 * g + 14]) * inv, __uint_as_float(o[16 * g + 15]) * inv);
            stg256_u(orow + c * 32 + g * 16, u0, u1);
          }
        }
      }
    }
    }
}

// SWAP: the MUFU group runs on warps 4-7 and the FMA-pipe group on warps 0-3 (the warp scheduler prefers the higher
// warp id among eligible warps: the group that is bound by a scarce pipe should win the issue slot)
template <int MAIN, int POLY8, bool SWAP, int APOLY>
__global__ void __launch_bounds__(attn2::THREADS, 2)
attn_split_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  using namespace attn2;
  static_assert(MAIN % 16 == 0 && MAIN >= 64 && MAIN <= 112 && (APOLY == 8 || MAIN == 64), "main column count");
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // K_STAGES
  uint64_t* k_empty = k_full + K_STAGES;
  uint64_t* v_full = k_empty + K_STAGES;   // V_STAGES
  uint64_t* v_empty = v_full + V_STAGES;
  uint64_t* s_full = v_empty + V_STAGES;   // S(t) accumulator complete                    (MMA -> both softmax groups)
  uint64_t* s_free = s_full + 1;           // S(t) copied to registers by all 256 threads  (softmax -> MMA)
  uint64_t* p_full = s_free + 1;           // P(t) in TMEM, all 256 threads                (softmax -> MMA)
  uint64_t* pv_done = p_full + 1;          // O += P(t) V(t) complete                      (MMA -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
  float* xch0 = reinterpret_cast<float*>(smem + XCH_OFF);      // written by main
  float* xch1 = xch0 + 128;                                     // written by assist

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tile = blockIdx.x % p.n_tiles;
  const int head = blockIdx.x / p.n_tiles;
  const int seq = p.tile_seq[tile];
  const int q0 = p.tile_q0[tile];
  const int seq_row0 = p.seq_off[seq];
  const int kv_len = p.seq_len[seq];
  const int n_kv = (kv_len + 127) >> 7;

  constexpr int PRE = K_STAGES < V_STAGES ? K_STAGES : V_STAGES;
  const int n_pre = n_kv < PRE ? n_kv : PRE;
  if (warp == 8 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < K_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < V_STAGES; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
    pdl_wait();                  // qkv is the predecessor's output
    tma_prefetch_desc(&tmQKV);
    mbar_expect_tx(q_full, TILE_BYTES);
    tma_load_2d(smem + Q_OFF, &tmQKV, head * 64, seq_row0 + q0, q_full);
    for (int j = 0; j < n_pre; ++j) {      // ring slots are empty: no wait
      mbar_expect_tx(&k_full[j], TILE_BYTES);
      tma_load_2d(smem + K_OFF + j * TILE_BYTES, &tmQKV, p.dim + head * 64, seq_row0 + j * 128, &k_full[j]);
      mbar_expect_tx(&v_full[j], TILE_BYTES);
      tma_load_2d(smem + V_OFF + j * TILE_BYTES, &tmQKV, 2 * p.dim + head * 64, seq_row0 + j * 128, &v_full[j]);
    }
  }
  if (warp == 9) tmem_alloc(tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // ------------------------------------------------------------------------------------------------------------
  // Every role runs the same sequence of passes: the optimistic one, and — only if the CTA-wide vote after it says
  // that some row overflowed — a maximum pass and an exact pass.  `it` counts kv tiles over all passes (parity of
  // s_full / s_free), `pt` the tiles of the passes that produce P (parity of p_full / pv_done).
  // ------------------------------------------------------------------------------------------------------------
  if (warp >= 8) {
    setmaxnreg_dec<AUX_REGS>();
    // The CTA-wide vote (bar.red) must be executed by whole, converged warps: every lane of warps 8 / 9 walks the
    // pass loop, lane 0 alone does the work inside it, and the warp reconverges before the vote.
    if (warp == 8) {
      // ---------------------------------------------------------------- TMA producer: one sweep over K / V per pass
      const int kcol = p.dim + head * 64, vcol = 2 * p.dim + head * 64;
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      int sweeps = 1;
      for (int sweep = 0; sweep < sweeps; ++sweep) {
        if (lane == 0) {
          for (int j = 0; j < n_kv; ++j) {
            const bool pre = sweep == 0 && j < n_pre;          // issued before the CTA barrier
            if (!pre) {
              mbar_wait(&k_empty[ks], kph ^ 1);
              mbar_expect_tx(&k_full[ks], TILE_BYTES);
              tma_load_2d(smem + K_OFF + ks * TILE_BYTES, &tmQKV, kcol, seq_row0 + j * 128, &k_full[ks]);
            }
            if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
            if (!pre) {
              mbar_wait(&v_empty[vs], vph ^ 1);
              mbar_expect_tx(&v_full[vs], TILE_BYTES);
              tma_load_2d(smem + V_OFF + vs * TILE_BYTES, &tmQKV, vcol, seq_row0 + j * 128, &v_full[vs]);
            }
            if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
          }
        }
        __syncwarp();
        if (sweep == 0 && __syncthreads_or(0)) sweeps = 3;     // the CTA voted for the two fallback passes
      }
    } else if (warp == 9) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 1);
      const uint32_t k_addr = smem_u32(smem + K_OFF);
      const uint32_t v_addr = smem_u32(smem + V_OFF);
      const uint64_t a0 = make_sdesc_sw128(smem_u32(smem + Q_OFF));
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      int it = 0, pt = 0;
      if (lane == 0) mbar_wait(q_full, 0);
      int passes = 1;
      for (int pass = 0; pass < passes; ++pass) {
        if (lane == 0) {
          const bool pv = pass != 1;          // pass 1 of a fallback = maximum pass: scores only
          // S(n+1) is issued as soon as S(n) has been copied out of TMEM: it runs under softmax(n)
          for (int n = 0; n <= n_kv; ++n) {
            if (n < n_kv) {
              if (it > 0) mbar_wait(s_free, (it - 1) & 1);
              mbar_wait(&k_full[ks], kph);
              tc_fence_after();
              const uint64_t b0 = make_sdesc_sw128(k_addr + ks * TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_ss(tmem_base + TM_S, a0 + 2 * k, b0 + 2 * k, idesc_s, k != 0);
              umma_commit(s_full);
              umma_commit(&k_empty[ks]);
              if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
              ++it;
            }
            if (n > 0) {
              if (pv) {
                mbar_wait(p_full, pt & 1);
                mbar_wait(&v_full[vs], vph);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const uint64_t b = make_sdesc_sw128(v_addr + vs * TILE_BYTES + k * 2048);
                  umma_ts(tmem_base + TM_O, tmem_base + TM_P + k * 8, b, idesc_o, !(n == 1 && k == 0));
                }
                umma_commit(pv_done);
                umma_commit(&v_empty[vs]);
                ++pt;
              } else {                        // the V stage is consumed unread
                mbar_wait(&v_full[vs], vph);
                mbar_arrive(&v_empty[vs]);
              }
              if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
            }
          }
        }
        __syncwarp();
        if (pass == 0 && __syncthreads_or(0)) passes = 3;
      }
    } else {
      (void)__syncthreads_or(0);              // idle warps only take part in the vote
    }
  } else {
    // ------------------------------------------------------------------ the two softmax warpgroups: thread = row
    if ((warp < 4) != SWAP) {
      setmaxnreg_inc<(APOLY == 8 ? MAIN_REGS_FMA_ASSIST : MAIN_REGS_SYM)>();
      softmax_role<MAIN, POLY8, APOLY, true>(p, tmem_base, warp, q0, seq_row0, kv_len, n_kv, head, s_full, s_free, p_full,
                                      pv_done, xch0, xch1);
    } else {
      if (APOLY == 8) setmaxnreg_dec<ASSIST_REGS_FMA>();
      else setmaxnreg_inc<ASSIST_REGS_SYM>();
      softmax_role<MAIN, POLY8, APOLY, false>(p, tmem_base, warp, q0, seq_row0, kv_len, n_kv, head, s_full, s_free, p_full,
                                       pv_done, xch0, xch1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, TM_COLS);
}

template <int MAIN, int POLY8, bool SWAP, int APOLY = 8>
static void launch_split(const CUtensorMap& tmQKV, const AttnParams& p, cudaStream_t st) {
  static DeviceOnce attr;
  attr.once([] {
    cudaFuncSetAttribute(attn_split_kernel<MAIN, POLY8, SWAP, APOLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn2::SMEM);
  });
  launch_k(attn_split_kernel<MAIN, POLY8, SWAP, APOLY>, p.n_tiles * p.heads, attn2::THREADS, attn2::SMEM, st, tmQKV, p);
}

// VVB200_ATTN selects the kernel: "1" = generation 1 (attn.cu), "96" / "96p" / "112" / "112p" = split kernel with that
// many main columns, "p" = one main pair in eight on the FMA pipe as well.
bool launch_attention_split(const CUtensorMap& tmQKV, const AttnParams& p, cudaStream_t st) {
  static const int mode = [] {
    const char* v = getenv("VVB200_ATTN");
    if (!v || !v[0]) return VV_ATTN_DEFAULT;
    if (!strcmp(v, "1")) return 0;
    if (!strcmp(v, "3")) return 3;
    if (!strcmp(v, "96")) return 960;
    if (!strcmp(v, "96p")) return 961;
    if (!strcmp(v, "112")) return 1120;
    if (!strcmp(v, "112p")) return 1121;
    if (!strcmp(v, "80")) return 800;
    if (!strcmp(v, "64")) return 640;
    if (!strcmp(v, "64p")) return 641;
    if (!strcmp(v, "96s")) return 962;
    if (!strcmp(v, "112s")) return 1122;
    return VV_ATTN_DEFAULT;
  }();
  if (mode == 0 || p.n_tiles <= 0) return mode != 0;
  switch (mode) {
    case 3: launch_attention_gen3(tmQKV, p, st); break;
    case 960: launch_split<96, 0, false>(tmQKV, p, st); break;
    case 961: launch_split<96, 1, false>(tmQKV, p, st); break;
    case 1120: launch_split<112, 0, false>(tmQKV, p, st); break;
    case 1121: launch_split<112, 1, false>(tmQKV, p, st); break;
    case 800: launch_split<80, 0, false>(tmQKV, p, st); break;
    case 640: launch_split<64, 0, false, 0>(tmQKV, p, st); break;
    case 641: launch_split<64, 1, false, 1>(tmQKV, p, st); break;
    case 962: launch_split<96, 0, true>(tmQKV, p, st); break;
    case 1122: launch_split<112, 0, true>(tmQKV, p, st); break;
    default: return false;
  }
  return true;
}

#if VV_ATTN_TIMING
extern "C" void vv_attn2_timing_dump() {
  unsigned long long h[2][12];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_attn2_timing, sizeof(h));
  const char* names[8] = {"wait_S", "S readout", "first max / m_ref", "exp2 (to PV wait)", "wait PV(j-1)",
                          "exp2 rest + P store", "P retire+arrive", "loop"};
  for (int g = 0; g < 2; ++g) {
    double tot = 0;
    for (int i = 0; i < 8; ++i) tot += double(h[g][i]);
    const double n = double(h[g][11] ? h[g][11] : 1);
    printf(" %s warps: %.0f cycles per kv tile per warp\n", g == 0 ? "MAIN" : "ASSIST", tot / n);
    for (int i = 0; i < 8; ++i)
      printf("  %-22s %5.1f%%  %8.0f cyc/kv-tile\n", names[i], 100.0 * double(h[g][i]) / (tot + 1e-9), double(h[g][i]) / n);
  }
  unsigned long long z[2][12] = {};
  cudaMemcpyToSymbol(g_attn2_timing, z, sizeof(z));
}
#endif

}  // namespace vv
