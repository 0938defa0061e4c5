// Non-causal flash attention for sm_100a (d_h = 64), tcgen05 + TMEM.
// One CTA = one head x 256 query rows (two 128-row tiles, each owned by one softmax warpgroup).
//   S_t = Q_t K^T      : tcgen05.mma M128 N128 K64, accumulator in TMEM (128 cols per tile)
//   softmax (online, lazy rescale) in registers: one thread per query row, no shuffles
//   P_t -> bf16 written back to TMEM (tcgen05.st, 64 cols per tile); O_t += P_t V : tcgen05.mma M128 N64 K128 with the
//   A operand read from TMEM and V consumed MN-major straight from the TMA tile (no transpose); O in TMEM (64 cols)
// RoPE has already been applied to q/k by the QKV GEMM epilogue.
//
// Replaces the attention sub-graph of `transformer.onnx` (/root/reference/vietvoicetts/core/tts_engine.py:161-172).
#include "kernels.h"
#include "ptx.cuh"

#include <stdio.h>

#ifndef VV_ATTN_PINGPONG
#define VV_ATTN_PINGPONG 0
#endif

#ifndef VV_ATTN_POLY_N
#define VV_ATTN_POLY_N 0     // of every VV_ATTN_POLY_MOD softmax elements, this many take the FMA-pipe exp2
#endif
#ifndef VV_ATTN_POLY_MOD
#define VV_ATTN_POLY_MOD 4
#endif
#ifndef VV_ATTN_P_TMEM
#define VV_ATTN_P_TMEM 1    // 1: P goes back to TMEM (tcgen05.st) and the PV MMA reads its A operand from TMEM
#endif
#ifndef VV_ATTN_SKEW_NS
#define VV_ATTN_SKEW_NS 0
#endif
#ifndef VV_ATTN_TIMING
#define VV_ATTN_TIMING 0
#endif
#if VV_ATTN_TIMING
#define TICK(i) do { long long _t = clock64(); tacc[i] += _t - tlast; tlast = _t; } while (0)
#else
#define TICK(i) do { } while (0)
#endif

namespace vv {

#if VV_ATTN_TIMING
__device__ long long g_attn_timing[8];
#endif

namespace attn {
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int Q_OFF = 0;                                   // 2 tiles
constexpr int K_OFF = Q_OFF + 2 * TILE_BYTES;              // KV_STAGES tiles
constexpr int V_OFF = K_OFF + KV_STAGES * TILE_BYTES;      // KV_STAGES tiles
constexpr int P_OFF = V_OFF + KV_STAGES * TILE_BYTES;      // 2 tiles x 2 atoms
constexpr int BAR_OFF = P_OFF + 4 * TILE_BYTES;
constexpr int SMEM = BAR_OFF + 256 + 1024;
constexpr int THREADS = 384;
constexpr uint32_t TM_S = 0;     // S0 @0, S1 @128
constexpr uint32_t TM_O = 256;   // O0 @256, O1 @320
constexpr uint32_t TM_P = 384;   // P0 @384, P1 @448 (bf16 pairs: 64 columns per tile) — only with VV_ATTN_P_TMEM
}  // namespace attn

__global__ void __launch_bounds__(attn::THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // KV_STAGES
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;  // 2: S_t(j) accumulator complete            (MMA -> softmax)
  uint64_t* s_free = s_full + 2;           // 2: S_t(j) copied to registers              (softmax -> MMA)
  uint64_t* p_full = s_free + 2;           // 2: P_t(j) in smem, O_t rescaled             (softmax -> MMA)
  uint64_t* pv_done = p_full + 2;          // 2: O_t += P_t(j) V(j) complete              (MMA -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tile = blockIdx.x % p.n_tiles;
  const int head = blockIdx.x / p.n_tiles;
  const int seq = p.tile_seq[tile];
  const int q0 = p.tile_q0[tile];
  const int seq_row0 = p.seq_off[seq];
  const int kv_len = p.seq_len[seq];
  const int n_kv = (kv_len + 127) >> 7;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);   // one commit per q-tile MMA warp
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
  // producer warpgroup (TMA, MMA issue, TMEM alloc, spare): hand registers to the two softmax warpgroups
  setmaxnreg_dec<88>();
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      tma_load_2d(smem + Q_OFF, &tmQKV, head * 64, seq_row0 + q0, q_full);
      tma_load_2d(smem + Q_OFF + TILE_BYTES, &tmQKV, head * 64, seq_row0 + q0 + 128, q_full);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_expect_tx(&k_full[stage], TILE_BYTES);
        tma_load_2d(smem + K_OFF + stage * TILE_BYTES, &tmQKV, p.dim + head * 64, seq_row0 + j * 128, &k_full[stage]);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_expect_tx(&v_full[stage], TILE_BYTES);
        tma_load_2d(smem + V_OFF + stage * TILE_BYTES, &tmQKV, 2 * p.dim + head * 64, seq_row0 + j * 128,
                    &v_full[stage]);
        if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9 || warp == 11) {
    // ------------------------------------------------------------------ MMA issuers: one warp per q tile, so the
    // two tiles form independent S -> softmax -> PV pipelines that only share the K/V ring and the tensor pipe.
    // S_t(n+1) is issued as soon as softmax has copied S_t(n) to registers, i.e. it runs under softmax_t(n).
    if (lane == 0) {
      const int t = warp == 9 ? 0 : 1;
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 1);
      const uint32_t q_addr = smem_u32(smem + Q_OFF) + t * TILE_BYTES;
      const uint32_t k_addr = smem_u32(smem + K_OFF);
      const uint32_t v_addr = smem_u32(smem + V_OFF);
      const uint32_t p_addr = smem_u32(smem + P_OFF) + 2 * t * TILE_BYTES;
      const uint32_t d_s = tmem_base + TM_S + t * 128;
      const uint32_t d_o = tmem_base + TM_O + t * 64;
      auto issue_s = [&](int stage) {
        const uint64_t a0 = make_sdesc_sw128(q_addr);
        const uint64_t b0 = make_sdesc_sw128(k_addr + stage * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d_s, a0 + 2 * k, b0 + 2 * k, idesc_s, k != 0);
      };
      auto issue_pv = [&](int stage, bool first) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t b = make_sdesc_sw128(v_addr + stage * TILE_BYTES + k * 2048);
#if VV_ATTN_P_TMEM
          umma_ts(d_o, tmem_base + TM_P + t * 64 + k * 8, b, idesc_o, !(first && k == 0));
#else
          const uint64_t a = make_sdesc_sw128(p_addr + (k >> 2) * TILE_BYTES) + 2 * (k & 3);
          umma_ss(d_o, a, b, idesc_o, !(first && k == 0));
#endif
        }
      };
      mbar_wait(q_full, 0);
      int stage = 0, pstage = 0;
      uint32_t phase = 0, pphase = 0;
      for (int n = 0; n <= n_kv; ++n) {
        if (n < n_kv) {
          if (n > 0) mbar_wait(&s_free[t], (n - 1) & 1);
          mbar_wait(&k_full[stage], phase);
          tc_fence_after();
          issue_s(stage);
          umma_commit(&s_full[t]);
          umma_commit(&k_empty[stage]);
          if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
        }
        if (n > 0) {
          mbar_wait(&p_full[t], (n - 1) & 1);
          mbar_wait(&v_full[pstage], pphase);
          tc_fence_after();
          issue_pv(pstage, n == 1);
          umma_commit(&pv_done[t]);
          umma_commit(&v_empty[pstage]);
          if (++pstage == KV_STAGES) { pstage = 0; pphase ^= 1; }
        }
      }
    }
  }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    setmaxnreg_inc<208>();
    const int t = warp >> 2;                 // q tile
    const int r = threadIdx.x & 127;         // row within tile == TMEM lane
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t ts = tmem_base + lane_base + TM_S + t * 128;
    const uint32_t to = tmem_base + lane_base + TM_O + t * 64;
#if !VV_ATTN_P_TMEM
    const uint32_t prow = smem_u32(smem + P_OFF) + (2 * t) * TILE_BYTES + r * 128;   // shared-space address
    const int sw = r & 7;
#endif
    float m_ref = 0.0f, l = 0.0f;
#if VV_ATTN_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif
#if VV_ATTN_PINGPONG
    if (t == 1) named_bar_arrive(2, 256);   // warpgroup 0 takes the first MUFU turn
#endif
#if VV_ATTN_SKEW_NS > 0
    if (t == 1) __nanosleep(VV_ATTN_SKEW_NS);   // start tile 1 half an iteration late: its exp2 burst then falls under tile 0's non-MUFU phases
#endif
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      TICK(0);   // wait S
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(ts + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[t]);               // S_t may be overwritten by S_t(j+1) from here on
      TICK(1);   // tmem ld
      const int kv_valid = kv_len - j * 128;
      if (kv_valid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
      }
      // 4 independent 3-input max chains instead of one 127-deep dependent chain
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
      float mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
      mx *= p.scale_log2;
      TICK(2);   // mask + max
      // lazy rescale decision (the O update itself is deferred until PV(j-1) has retired, below)
      bool rescale = false;
      float f = 1.0f;
      if (j == 0) {
        m_ref = mx;
      } else if (__any_sync(0xffffffffu, (mx - m_ref) > 8.0f)) {
        const float m_new = fmaxf(m_ref, mx);
        f = fast_exp2(m_ref - m_new);
        m_ref = m_new;
        rescale = true;
      }
      // The two softmax warpgroups take turns on the MUFU pipe (ping-pong): while one runs its 128x128 exp2
      // burst, the other does its TMEM load / max / P store / barrier traffic.
#if VV_ATTN_PINGPONG
      named_bar_sync(2 + t, 256);
#endif
      uint32_t pk[64];
      float sum0 = 0.0f, sum1 = 0.0f, sum2 = 0.0f, sum3 = 0.0f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {  // 16 chunks of 8 columns (16 bytes of bf16)
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xs = fmaf(__uint_as_float(s[c * 8 + i]), p.scale_log2, -m_ref);
          e[i] = (i % VV_ATTN_POLY_MOD) < VV_ATTN_POLY_N ? poly_exp2(xs) : fast_exp2(xs);
        }
        sum0 += e[0] + e[4];
        sum1 += e[1] + e[5];
        sum2 += e[2] + e[6];
        sum3 += e[3] + e[7];
        pk[4 * c] = pack_bf16(e[0], e[1]);
        pk[4 * c + 1] = pack_bf16(e[2], e[3]);
        pk[4 * c + 2] = pack_bf16(e[4], e[5]);
        pk[4 * c + 3] = pack_bf16(e[6], e[7]);
      }
#if VV_ATTN_PINGPONG
      named_bar_arrive(2 + (1 - t), 256);
#endif
      TICK(3);   // exp + pack (incl. ping-pong wait)
      if (j > 0) {
        mbar_wait(&pv_done[t], (j - 1) & 1);   // P_t buffer free again, O_t quiescent
        TICK(4);   // wait PV
        tc_fence_after();
        if (rescale) {
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
          tmem_st_wait();
        }
      }
#if VV_ATTN_P_TMEM
      tmem_st32(tmem_base + lane_base + TM_P + t * 64, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
      tmem_st32(tmem_base + lane_base + TM_P + t * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
#else
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int atom = c >> 3, chunk = c & 7;
        st_shared_v4(prow + atom * TILE_BYTES + ((chunk ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2],
                     pk[4 * c + 3]);
      }
#endif
      l += (sum0 + sum1) + (sum2 + sum3);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      TICK(5);   // rescale + P store + fence + arrive
    }
    // ---- finalize: O / l -> bf16
    mbar_wait(&pv_done[t], (n_kv - 1) & 1);
    TICK(6);   // final wait
    tc_fence_after();
    const int qrow = q0 + t * 128 + r;
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
    TICK(7);   // O store
    if (lane == 0 && blockIdx.x % 97 == 0)
      for (int i = 0; i < 8; ++i)
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_attn_timing[i]), (unsigned long long)tacc[i]);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc(tmem_base, 512);
}

#if VV_ATTN_TIMING
extern "C" void vv_attn2_timing_dump() {
  long long h[8];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_attn_timing, sizeof(h));
  const char* names[8] = {"wait_S", "tmem_ld", "mask_max", "exp_pack", "wait_PV", "store_arrive", "final_wait", "O_store"};
  long long tot = 0;
  for (int i = 0; i < 8; ++i) tot += h[i];
  for (int i = 0; i < 8; ++i) printf("%-14s %12lld  %5.1f%%\n", names[i], h[i], 100.0 * h[i] / (tot ? tot : 1));
  long long z[8] = {0};
  cudaMemcpyToSymbol(g_attn_timing, z, sizeof(z));
}
#endif

void launch_attention2(const CUtensorMap& tmQKV, const AttnParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM);
    attr_set = true;
  }
  if (p.n_tiles <= 0) return;
  attn_kernel<<<p.n_tiles * p.heads, attn::THREADS, attn::SMEM, st>>>(tmQKV, p);
}

}  // namespace vv
