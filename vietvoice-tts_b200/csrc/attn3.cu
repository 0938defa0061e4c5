// GENERATION 3 of the attention kernel = generation 1 (attn.cu, read that header first) + two changes that follow from
// its phase timing (profiles/r02_attn_phase_timing.txt): the two CTAs of an SM run in a convoy (their exp2 phases
// coincide, MUFU 100 % busy, and so do their gaps, MUFU idle), so a kv tile costs (2 x MUFU work) + (gap), and the gap
// is what can still be removed:
//   * S(j+1) IS READ OUT OF TMEM DURING softmax(j): as soon as a 32-column group of S(j) has been exponentiated its
//     registers are dead, and the tcgen05.ld of the same group of S(j+1) — which the MMA warp finished long ago — is
//     issued into them.  At the end of the tile one tcgen05.wait::ld remains; the 4 x LDTM + wait_S of the next tile
//     (~250 of ~2450 cycles per tile) are gone.
//   * S(j) no longer exists when the row sum of tile j is known, so the per-warp "reference outgrown" repair of
//     generation 1 cannot redo a tile.  Overflow is a sticky per-thread flag; the CTA votes after its last tile and, if
//     any row tripped it, runs the item again with the generation-1 loop (no prefetch, per-warp repair).
//
// Non-causal flash attention for sm_100a (d_h = 64), tcgen05 + TMEM: one 128-row query tile per CTA, TWO CTAs resident
// per SM (256 TMEM columns, ~82 KB smem, 256 threads each: softmax warpgroup | TMA warp, MMA warp, 2 idle warps).
//
// At d_h = 64 this op is bound by the exponential, not the tensor pipe: a 128x128 score tile costs 512 tensor cycles
// (QK^T + PV) but 1024 MUFU cycles (16 ex2/clk/SM).  Design points, each measured against the alternative
// (profiles/README.md; the other generations are in the git history):
//   * two free-running CTAs per SM instead of two tiles per CTA in lock step: the exp2 phases of the two tiles drift
//     apart and one CTA's Q/K load latency and O store run under the other's main loop (MUFU pipe 49 % -> 65 %);
//   * no running maximum at all after the first kv tile: exp2 runs against the reference m_ref fixed by the first tile.
//     Scaling by a power of two is exact in floating point, so a stale reference costs no precision as long as nothing
//     overflows: P (bf16) and the fp32 accumulators l and O have 2^127 of head room.  The row sum the loop computes
//     anyway is the detector — only if it exceeds 2^40 (a score 40 bits above the reference, or inf) the warp takes
//     the tile's true maximum as new reference, redoes the tile's exp2 and rescales O and l.  The inner loop is
//     FFMA2 + 2 MUFU + FADD2 + F2FP per pair of scores, nothing else;
//   * packed FFMA2 / FADD2 for scale-subtract and the row sum, no compare/select on full tiles (the tail mask is a
//     separate loop instance), P streamed to TMEM in 16-column groups (attn_softmax.cuh);
//   * tried and dropped: explicit MUFU ping-pong between two tiles in a persistent CTA (a single softmax warp cannot
//     saturate its scheduler's MUFU, so taking turns only serialises two latency-bound streams), two threads per
//     query row (the per-tile maximum exchange and the exposed PV round trip cost more than the extra warps gain);
//   * one score pair in EIGHT takes its exp2 on the FMA/ALU pipes (packed degree-3 polynomial, attn_softmax.cuh)
//     instead of MUFU.EX2.  Sweep, back to back on one box: 0/8 272-277 us, 1/8 265 us, 2/8 270 us, 3/8 284-319 us,
//     4/8 297 us — 12 instructions replace 2, so the loop turns issue-bound quickly.
//   S = Q K^T : tcgen05.mma M128 N128 K64 -> TMEM cols [0,128)
//   P (bf16)  : tcgen05.st -> TMEM cols [128,192); O += P V : tcgen05.mma M128 N64 K128, A from TMEM, V MN-major
//               straight from the TMA tile; O in TMEM cols [192,256)
// RoPE has already been applied to q/k by the QKV GEMM epilogue.
//
// Replaces the attention sub-graph of `transformer.onnx` (/root/reference/vietvoicetts/core/tts_engine.py:161-172).
#include "kernels.h"
#include "ptx.cuh"
#include "attn_softmax.cuh"

#include <stdio.h>
#include <stdlib.h>

#ifndef VV_ATTN_TIMING
#define VV_ATTN_TIMING 0
#endif
#if VV_ATTN_TIMING
#define TICK3(i) do { long long _t = clock64(); tacc[i] += _t - tlast; tlast = _t; } while (0)
#else
#define TICK3(i) do { } while (0)
#endif

namespace vv {

#if VV_ATTN_TIMING
__device__ unsigned long long g_attn3_timing[10];
#endif

namespace attn3 {
constexpr int K_STAGES = 2;
constexpr int V_STAGES = 2;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int Q_OFF = 0;
constexpr int K_OFF = Q_OFF + TILE_BYTES;
constexpr int V_OFF = K_OFF + K_STAGES * TILE_BYTES;
constexpr int BAR_OFF = V_OFF + V_STAGES * TILE_BYTES;
constexpr int SMEM = BAR_OFF + 256 + 1024;
constexpr int THREADS = 256;   // softmax warpgroup | TMA warp, MMA warp, 2 idle warps
constexpr uint32_t TM_S = 0;
constexpr uint32_t TM_P = 128;
constexpr uint32_t TM_O = 192;
constexpr uint32_t TM_COLS = 256;
constexpr float SUM_LIMIT_PF = 1.2089258196146292e24f;   // 2^80: sticky overflow flag of the prefetching pass
constexpr float SUM_LIMIT = 1.099511627776e12f;           // 2^40: per-warp repair of the fallback pass (generation 1)
}  // namespace attn3

// softmax_row of attn_softmax.cuh with the readout of the NEXT tile folded in: after the 32-column group c of S(j) has
// been exponentiated, its registers receive group c of S(j+1) (tcgen05.ld is asynchronous: the data lands while the
// following groups are exponentiated; the caller issues the one tcgen05.wait::ld at the end of the tile).
template <bool MASKED>
__device__ __forceinline__ void softmax_row_pf(uint32_t (&s)[128], float scale_log2, float m, int kv_valid, uint32_t tp,
                                               uint64_t* pv_bar, uint32_t pv_parity, bool pv_wait, float& sum,
                                               bool prefetch, uint32_t ts, uint64_t* s_bar, uint32_t s_parity) {
  float sa0 = 0.0f, sa1 = 0.0f, sb0 = 0.0f, sb1 = 0.0f;
  auto val = [&](int i) { return (MASKED && i >= kv_valid) ? __uint_as_float(0xff800000u) : __uint_as_float(s[i]); };
  uint32_t pk_all[4][16];
  const bool pv_ready = pv_wait ? mbar_test(pv_bar, pv_parity) : true;
  bool s_ready = prefetch ? mbar_test(s_bar, s_parity) : false;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t (&pk)[16] = pk_all[c];
    if (MASKED && c * 32 >= kv_valid) {
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = 0u;
    } else
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const int k = c * 32 + i;
      float x0, x1, x2, x3;
      ffma2(x0, x1, val(k), val(k + 1), scale_log2, -m);
      ffma2(x2, x3, val(k + 2), val(k + 3), scale_log2, -m);
      float e0, e1, e2, e3;
      if (((i / 2) % 8) >= 8 - VV_ATTN_POLY8) {
        poly_exp2_pair(x0, x1, e0, e1);
      } else {
        e0 = fast_exp2(x0);
        e1 = fast_exp2(x1);
      }
      if (((i / 2 + 1) % 8) >= 8 - VV_ATTN_POLY8) {
        poly_exp2_pair(x2, x3, e2, e3);
      } else {
        e2 = fast_exp2(x2);
        e3 = fast_exp2(x3);
      }
      fadd2(sa0, sa1, e0, e1);
      fadd2(sb0, sb1, e2, e3);
      pk[i / 2] = pack_bf16(e0, e1);
      pk[i / 2 + 1] = pack_bf16(e2, e3);
    }
    if (c == 1) {
      if (pv_wait) {
        if (!pv_ready) mbar_wait(pv_bar, pv_parity);       // P buffer free again, O quiescent
        tc_fence_after();
      }
      tmem_st16(tp, pk_all[0]);
      tmem_st16(tp + 16, pk_all[1]);
      // the first two P groups were held in registers until here; now that they are stored, the registers of the two
      // score groups they came from take the next tile's scores (no more registers live than in generation 1)
      if (prefetch) {
        if (!s_ready) mbar_wait(s_bar, s_parity);          // S(j+1): issued when S(j) had been read out, long complete
        tc_fence_after();
        tmem_ld32(ts, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld32(ts + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      }
    } else if (c > 1) {
      tmem_st16(tp + c * 16, pk);
      if (prefetch) tmem_ld32(ts + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
    }
  }
  sum = (sa0 + sa1) + (sb0 + sb1);
}

__global__ void __launch_bounds__(attn3::THREADS, 2)
attn3_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  using namespace attn3;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // K_STAGES
  uint64_t* k_empty = k_full + K_STAGES;
  uint64_t* v_full = k_empty + K_STAGES;   // V_STAGES
  uint64_t* v_empty = v_full + V_STAGES;
  uint64_t* s_full = v_empty + V_STAGES;   // S(t) accumulator complete             (MMA -> softmax)
  uint64_t* s_free = s_full + 1;           // S(t) copied to registers               (softmax -> MMA)
  uint64_t* p_full = s_free + 1;           // P(t) in TMEM                           (softmax -> MMA)
  uint64_t* pv_done = p_full + 1;          // O += P(t) V(t) complete                (MMA -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

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
  if (warp == 4 && lane == 0) {
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
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
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
  if (warp == 5) tmem_alloc(tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // Every role runs the item once, and a second time if the CTA-wide vote after the first pass reports an overflowed
  // row.  The vote (bar.red) is executed by whole, converged warps: all lanes of warps 4 / 5 walk the pass loop, lane 0
  // alone works inside it.  `it` counts kv tiles over both passes (parities of s_full / s_free, p_full / pv_done).
  if (warp >= 4) {
    setmaxnreg_dec<48>();
    if (warp == 4) {
      // ------------------------------------------------------------------ TMA producer
      const int kcol = p.dim + head * 64, vcol = 2 * p.dim + head * 64;
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      int sweeps = 1;
      for (int sweep = 0; sweep < sweeps; ++sweep) {
        if (lane == 0) {
          for (int j = 0; j < n_kv; ++j) {
            const bool pre = sweep == 0 && j < n_pre;        // issued before the CTA barrier
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
        if (sweep == 0 && __syncthreads_or(0)) sweeps = 2;
      }
    } else if (warp == 5) {
      // ------------------------------------------------------------------ MMA issuer.  S(n+1) is issued as soon as the
      // softmax warps have copied S(n) to registers.
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
              if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
              ++pt;
            }
          }
        }
        __syncwarp();
        if (pass == 0 && __syncthreads_or(0)) passes = 2;
      }
    } else {
      (void)__syncthreads_or(0);            // idle warps only take part in the vote
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroup: thread = query row
    setmaxnreg_inc<208>();

    const int r = threadIdx.x;                 // row within tile == TMEM lane
    const uint32_t lane_base = uint32_t(warp * 32) << 16;
    const uint32_t ts = tmem_base + lane_base + TM_S;
    const uint32_t tp = tmem_base + lane_base + TM_P;
    const uint32_t to = tmem_base + lane_base + TM_O;
    float m_ref = 0.0f, l = 0.0f;
    int it = 0;                                // kv tiles done, both passes (= P tiles done: every tile makes a P)
    uint32_t s[128];
#if VV_ATTN_TIMING
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif
    auto load_s = [&]() {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(ts + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);                     // S may be overwritten by the next tile's scores from here on
    };
    // exact row maximum (x scale) of the tile held in s[], keys >= kv_valid masked out
    auto row_max = [&](int kv_valid) {
      if (kv_valid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
      }
      float mxa = __uint_as_float(s[0]), mxb = __uint_as_float(s[1]), mxc = __uint_as_float(s[2]),
            mxd = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i < 124; i += 8) {
        mxa = fmax3(mxa, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        mxb = fmax3(mxb, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mxc = fmax3(mxc, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mxd = fmax3(mxd, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
      }
      mxa = fmax3(mxa, __uint_as_float(s[124]), __uint_as_float(s[125]));
      mxb = fmax3(mxb, __uint_as_float(s[126]), __uint_as_float(s[127]));
      return fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd)) * p.scale_log2;
    };

    // ---------------- pass 0: S(j+1) read out under softmax(j); overflow only flagged
    bool bad = false;
    mbar_wait(s_full, 0);
    TICK3(0);   // wait S(0)
    tc_fence_after();
    load_s();
    TICK3(1);   // S readout (first tile only)
    m_ref = row_max(kv_len);                   // kv_len >= 128 -> no mask; shorter sequences: the first tile is the tail
    TICK3(2);
    for (int j = 0; j < n_kv; ++j, ++it) {
      const int kv_valid = kv_len - j * 128;
      const bool more = j + 1 < n_kv;
      float sum;
      if (kv_valid < 128)
        softmax_row_pf<true>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, (it - 1) & 1, j > 0, sum, more, ts, s_full, (it + 1) & 1);
      else
        softmax_row_pf<false>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, (it - 1) & 1, j > 0, sum, more, ts, s_full, (it + 1) & 1);
      TICK3(3);   // exp2 + pack + P store (+ next tile's readout issued)
      bad |= !(sum < SUM_LIMIT_PF);
      l += sum;
      if (more) {
        tmem_ld_wait();                        // s[] now holds S(j+1)
        tc_fence_before();
        mbar_arrive(s_free);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      TICK3(6);   // retire + arrives
    }
    mbar_wait(pv_done, (it - 1) & 1);          // every MMA of the pass retired: O complete, TMEM quiescent
    tc_fence_after();
    TICK3(7);
    if (__syncthreads_or(bad ? 1 : 0)) {
      // ---------------- pass 1 (rare): generation-1 loop — no prefetch, per-warp repair of an outgrown reference
      l = 0.0f;
      for (int j = 0; j < n_kv; ++j, ++it) {
        mbar_wait(s_full, it & 1);
        tc_fence_after();
        load_s();
        const int kv_valid = kv_len - j * 128;
        const bool partial = kv_valid < 128;
        if (j == 0) m_ref = row_max(kv_valid);
        float sum;
        if (partial) softmax_row<true>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, (it - 1) & 1, j > 0, sum);
        else softmax_row<false>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, (it - 1) & 1, j > 0, sum);
        if (j > 0 && __any_sync(0xffffffffu, !(sum < SUM_LIMIT))) {
          const float m_new = fmaxf(m_ref, row_max(kv_valid));
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          tmem_st_wait();                      // first-pass P stores retired before the same columns are rewritten
          if (partial) softmax_row<true>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, 0, false, sum);
          else softmax_row<false>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, 0, false, sum);
          l *= f;
          uint32_t o[32];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tmem_ld32(to + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(to + c * 32, o);
          }
        }
        tmem_st_wait();
        l += sum;
        tc_fence_before();
        mbar_arrive(p_full);
      }
      mbar_wait(pv_done, (it - 1) & 1);
      tc_fence_after();
    }
    // ---- finalize: O / l -> bf16
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
#if VV_ATTN_TIMING
    TICK3(8);   // O store
    tacc[9] = n_kv;
    if (lane == 0 && blockIdx.x % 97 == 0)
      for (int i = 0; i < 10; ++i) atomicAdd(&g_attn3_timing[i], (unsigned long long)tacc[i]);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, TM_COLS);
}

void launch_attention_gen3(const CUtensorMap& tmQKV, const AttnParams& p, cudaStream_t st) {
  static DeviceOnce attr;
  attr.once([] { cudaFuncSetAttribute(attn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn3::SMEM); });
  if (p.n_tiles <= 0) return;
  launch_k(attn3_kernel, p.n_tiles * p.heads, attn3::THREADS, attn3::SMEM, st, tmQKV, p);
}

#if VV_ATTN_TIMING
extern "C" void vv_attn3_timing_dump() {
  unsigned long long h[10];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_attn3_timing, sizeof(h));
  const char* names[9] = {"wait_S (first tile)", "S readout (first)", "first max", "exp2+pack+P+prefetch", "-", "-",
                          "retire+arrives", "final PV wait", "O store"};
  double tot = 0;
  for (int i = 0; i < 9; ++i) tot += double(h[i]);
  printf(" generation 3: %.0f cycles per kv tile per warp\n", tot / double(h[9] ? h[9] : 1));
  for (int i = 0; i < 9; ++i)
    if (h[i]) printf("  %-22s %5.1f%%  %8.0f cyc/kv-tile\n", names[i], 100.0 * double(h[i]) / (tot + 1e-9), double(h[i]) / double(h[9] ? h[9] : 1));
  unsigned long long z[10] = {0};
  cudaMemcpyToSymbol(g_attn3_timing, z, sizeof(z));
}
#endif

}  // namespace vv
