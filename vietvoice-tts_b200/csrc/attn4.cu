// Non-causal flash attention for sm_100a (d_h = 64), tcgen05 + TMEM — version 4: persistent, two independent
// "lanes" per CTA with an explicit MUFU ping-pong.
//
// At d_h = 64 the op is bound by the exponential, not by the tensor pipe: a 128 x 128 score tile costs 512 tensor
// cycles (QK^T + PV) but 1024 MUFU cycles (16 ex2/clk/SM).  The SM therefore has to keep its MUFU pipe busy all the
// time.  v2 (two tiles per CTA in lock step) ran it at ~49 %, v3 (two free-running CTAs per SM) at ~69 %: the exp2
// phases of the two tiles overlapped at random and both then left the pipe idle together.
//
// One CTA per SM, 384 threads, persistent over (head, 128-row query tile) work items:
//   lane t in {0,1}:  softmax warpgroup t (thread = query row) + its own TMA thread + its own MMA thread + its own
//                     Q / K / V shared-memory ring + 256 TMEM columns (S | P | O).  A lane walks its own list of work
//                     items (idx = 2*cta + t, += 2*gridDim.x); its TMA and MMA threads run ahead into the next item, so
//                     the Q/K load latency and the first QK^T of an item are hidden under the previous item's tail.
//   ping-pong:        the two softmax warpgroups take strict turns on the exp2 burst (mbarrier tokens): while one is
//                     on the MUFU pipe, the other does its TMEM load, P store, O rescale, barrier traffic.
//   lazy maximum:     exp2 runs against the reference maximum of the previous kv tiles (tau = 8) and tracks the new
//                     maximum on the ALU pipe in the same loop; only if a row outgrew the reference the warp redoes
//                     the tile and rescales O.
//   S = Q K^T : tcgen05.mma M128 N128 K64, accumulator in TMEM; P (bf16) written back to TMEM with tcgen05.st and
//   consumed as the A operand of O += P V (M128 N64 K128, V MN-major straight from the TMA tile).
// RoPE has already been applied to q/k by the QKV GEMM epilogue.
//
// Replaces the attention sub-graph of `transformer.onnx` (/root/reference/vietvoicetts/core/tts_engine.py:161-172).
#include "kernels.h"
#include "ptx.cuh"
#include "attn_softmax.cuh"

#include <stdio.h>
#include <stdlib.h>

#ifndef VV_ATTN_PINGPONG
#define VV_ATTN_PINGPONG 1
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
__device__ unsigned long long g_attn4_timing[12];
#endif

namespace attn4 {
constexpr int K_STAGES = 3;
constexpr int V_STAGES = 2;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int LANE_BYTES = (1 + K_STAGES + V_STAGES) * TILE_BYTES;   // Q | K ring | V ring
constexpr int MAX_ITEMS = 32;             // work items per lane (grid is sized so that this holds)
constexpr int ITEM_OFF = 2 * LANE_BYTES;                     // int4 items[2][MAX_ITEMS]
constexpr int BAR_OFF = ITEM_OFF + 2 * MAX_ITEMS * 16;
constexpr int SMEM = BAR_OFF + 512 + 1024;
constexpr int THREADS = 384;
constexpr uint32_t TM_S = 0;
constexpr uint32_t TM_P = 128;
constexpr uint32_t TM_O = 192;
constexpr uint32_t TM_LANE = 256;
constexpr float TAU = 8.0f;
constexpr int BARS_PER_LANE = 2 + 2 * K_STAGES + 2 * V_STAGES + 4;
}  // namespace attn4

__global__ void __launch_bounds__(attn4::THREADS, 1)
attn4_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  using namespace attn4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  int4* items = reinterpret_cast<int4*>(smem + ITEM_OFF);   // [2][MAX_ITEMS]: (first row of the sequence, q0, kv_len, head)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* pp = bars + 2 * BARS_PER_LANE;                  // [2] exp2 burst of lane t finished
  int* n_items_s = reinterpret_cast<int*>(pp + 2);          // [2] items per lane, [2] kv iterations per lane
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(n_items_s + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = p.n_tiles * p.heads;
  const int stride = 2 * gridDim.x;

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      uint64_t* b = bars + t * BARS_PER_LANE;
      mbar_init(b + 0, 1);                                   // q_full
      mbar_init(b + 1, 1);                                   // q_empty
      for (int i = 0; i < 2 * K_STAGES + 2 * V_STAGES; ++i) mbar_init(b + 2 + i, 1);
      uint64_t* c = b + 2 + 2 * K_STAGES + 2 * V_STAGES;
      mbar_init(c + 0, 1);                                   // s_full
      mbar_init(c + 1, 128);                                 // s_free
      mbar_init(c + 2, 128);                                 // p_full
      mbar_init(c + 3, 1);                                   // pv_done
      mbar_init(&pp[t], 128);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp < 2) {
    // work list of lane t = warp: one item per thread, kv-iteration total by warp reduction
    const int t = warp;
    const int idx = 2 * blockIdx.x + t + lane * stride;
    int nkv = 0;
    if (idx < total_items) {
      const int tile = idx % p.n_tiles;
      const int head = idx / p.n_tiles;
      const int seq = p.tile_seq[tile];
      const int kv_len = p.seq_len[seq];
      items[t * MAX_ITEMS + lane] = make_int4(p.seq_off[seq], p.tile_q0[tile], kv_len, head);
      nkv = (kv_len + 127) >> 7;
    }
    const unsigned have = __ballot_sync(0xffffffffu, idx < total_items);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nkv += __shfl_xor_sync(0xffffffffu, nkv, o);
    if (lane == 0) {
      n_items_s[t] = __popc(have);
      n_items_s[2 + t] = nkv;
    }
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    setmaxnreg_dec<80>();
    const int t = (warp - 8) >> 1;                           // lane served by this warp
    uint64_t* b = bars + t * BARS_PER_LANE;
    uint64_t* q_full = b;
    uint64_t* q_empty = b + 1;
    uint64_t* k_full = b + 2;
    uint64_t* k_empty = k_full + K_STAGES;
    uint64_t* v_full = k_empty + K_STAGES;
    uint64_t* v_empty = v_full + V_STAGES;
    uint64_t* s_full = v_empty + V_STAGES;
    uint64_t* s_free = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    uint8_t* lsm = smem + t * LANE_BYTES;
    const int n_items = n_items_s[t];
    if (((warp - 8) & 1) == 0) {
      // ------------------------------------------------------------------ TMA producer of lane t
      if (lane == 0) {
        tma_prefetch_desc(&tmQKV);
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int it = 0; it < n_items; ++it) {
          const int4 w = items[t * MAX_ITEMS + it];
          const int n_kv = (w.z + 127) >> 7;
          mbar_wait(q_empty, (it & 1) ^ 1);
          mbar_expect_tx(q_full, TILE_BYTES);
          tma_load_2d(lsm, &tmQKV, w.w * 64, w.x + w.y, q_full);
          for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&k_empty[ks], kph ^ 1);
            mbar_expect_tx(&k_full[ks], TILE_BYTES);
            tma_load_2d(lsm + (1 + ks) * TILE_BYTES, &tmQKV, p.dim + w.w * 64, w.x + j * 128, &k_full[ks]);
            if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
            mbar_wait(&v_empty[vs], vph ^ 1);
            mbar_expect_tx(&v_full[vs], TILE_BYTES);
            tma_load_2d(lsm + (1 + K_STAGES + vs) * TILE_BYTES, &tmQKV, 2 * p.dim + w.w * 64, w.x + j * 128,
                        &v_full[vs]);
            if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
          }
        }
      }
    } else {
      // ------------------------------------------------------------------ MMA issuer of lane t.  S(n+1) is issued as
      // soon as the softmax warps have copied S(n) to registers, i.e. it runs under softmax(n); S(0) of the next
      // item is issued right behind the last PV of the current one.
      if (lane == 0) {
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0);
        constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 1);
        const uint32_t q_addr = smem_u32(lsm);
        const uint32_t k_addr = q_addr + TILE_BYTES;
        const uint32_t v_addr = k_addr + K_STAGES * TILE_BYTES;
        const uint32_t d_s = tmem_base + t * TM_LANE + TM_S;
        const uint32_t d_p = tmem_base + t * TM_LANE + TM_P;
        const uint32_t d_o = tmem_base + t * TM_LANE + TM_O;
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        uint32_t gs = 0, gp = 0;                             // S tiles issued / PV tiles issued so far
        for (int it = 0; it < n_items; ++it) {
          const int4 w = items[t * MAX_ITEMS + it];
          const int n_kv = (w.z + 127) >> 7;
          mbar_wait(q_full, it & 1);
          for (int n = 0; n <= n_kv; ++n) {
            if (n < n_kv) {
              if (gs > 0) mbar_wait(s_free, (gs - 1) & 1);
              mbar_wait(&k_full[ks], kph);
              tc_fence_after();
              const uint64_t a0 = make_sdesc_sw128(q_addr);
              const uint64_t b0 = make_sdesc_sw128(k_addr + ks * TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_ss(d_s, a0 + 2 * k, b0 + 2 * k, idesc_s, k != 0);
              umma_commit(s_full);
              umma_commit(&k_empty[ks]);
              if (n == n_kv - 1) umma_commit(q_empty);       // Q tile no longer needed: the next item's Q may land
              if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
              ++gs;
            }
            if (n > 0) {
              mbar_wait(p_full, gp & 1);
              mbar_wait(&v_full[vs], vph);
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint64_t bd = make_sdesc_sw128(v_addr + vs * TILE_BYTES + k * 2048);
                umma_ts(d_o, d_p + k * 8, bd, idesc_o, !(n == 1 && k == 0));
              }
              umma_commit(pv_done);
              umma_commit(&v_empty[vs]);
              if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
              ++gp;
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroup of lane t: thread = query row
    setmaxnreg_inc<208>();
    const int t = warp >> 2;
    uint64_t* b = bars + t * BARS_PER_LANE;
    uint64_t* s_full = b + 2 + 2 * K_STAGES + 2 * V_STAGES;
    uint64_t* s_free = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    const int r = threadIdx.x & 127;                         // row within tile == TMEM lane
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t ts = tmem_base + lane_base + t * TM_LANE + TM_S;
    const uint32_t tp = tmem_base + lane_base + t * TM_LANE + TM_P;
    const uint32_t to = tmem_base + lane_base + t * TM_LANE + TM_O;
    const int n_items = n_items_s[t];
    const uint32_t other_total = (uint32_t)n_items_s[2 + (1 - t)];   // kv iterations the other lane will run
    uint32_t g = 0;                                          // kv iterations of this lane so far
#if VV_ATTN_TIMING
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif
    for (int it = 0; it < n_items; ++it) {
      const int4 w = items[t * MAX_ITEMS + it];
      const int kv_len = w.z;
      const int n_kv = (kv_len + 127) >> 7;
      float m_ref = 0.0f, l = 0.0f;
      for (int j = 0; j < n_kv; ++j, ++g) {
        mbar_wait(s_full, g & 1);
        TICK(0);   // wait S
        tc_fence_after();
        uint32_t s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(ts + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free);                   // S may be overwritten by S(j+1) from here on
        TICK(1);   // tmem ld
        const int kv_valid = kv_len - j * 128;
        const bool partial = kv_valid < 128;   // warp-uniform: last kv tile of a sequence whose length is not k*128
        if (j == 0) {
          // first tile of an item: no reference yet -> plain row maximum (4 independent 3-input chains)
          if (partial) {
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
          m_ref = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd)) * p.scale_log2;
        }
        TICK(2);   // first-tile max
#if VV_ATTN_PINGPONG
        // strict turns on the MUFU pipe: lane 0 runs burst g after lane 1 finished burst g-1, lane 1 runs burst g
        // after lane 0 finished burst g (as long as the other lane still has bursts to run)
        if (t == 0) {
          if (g >= 1 && g - 1 < other_total) mbar_wait(&pp[1], (g - 1) & 1);
        } else {
          if (g < other_total) mbar_wait(&pp[0], g & 1);
        }
        TICK(8);   // wait for the MUFU turn
#endif
        float sum, mx;
        if (partial) softmax_row<true, true>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, (g - 1) & 1, j > 0, sum, mx);
        else softmax_row<true, false>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, (g - 1) & 1, j > 0, sum, mx);
        mx *= p.scale_log2;
        if (j > 0 && __any_sync(0xffffffffu, (mx - m_ref) > TAU)) {
          // some row of this warp outgrew the reference: redo this tile against the new maximum, rescale O and l
          const float m_new = fmaxf(m_ref, mx);
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          float dummy;
          tmem_st_wait();                      // first-pass P stores retired before the same columns are rewritten
          if (partial) softmax_row<false, true>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, 0, false, sum, dummy);
          else softmax_row<false, false>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, 0, false, sum, dummy);
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
#if VV_ATTN_PINGPONG
        mbar_arrive(&pp[t]);
#endif
        TICK(3);   // exp2 + max tracking + pack + P store (+ redo)
        tmem_st_wait();
        l += sum;
        tc_fence_before();
        mbar_arrive(p_full);
        TICK(6);   // O rescale + P store + arrive
      }
      // ---- finalize the item: O / l -> bf16
      mbar_wait(pv_done, (g - 1) & 1);
      tc_fence_after();
      const int qrow = w.y + r;
      const float inv = 1.0f / l;
      bf16* orow = p.out + (size_t)(w.x + qrow) * p.dim + w.w * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(to + c * 32, o);
        tmem_ld_wait();
        if (qrow < kv_len) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {   // 2 x 256-bit stores: whole 32-byte sectors per instruction
            uint4 u0, u1;
            u0.x = pack_bf16(__uint_as_float(o[16 * q]) * inv, __uint_as_float(o[16 * q + 1]) * inv);
            u0.y = pack_bf16(__uint_as_float(o[16 * q + 2]) * inv, __uint_as_float(o[16 * q + 3]) * inv);
            u0.z = pack_bf16(__uint_as_float(o[16 * q + 4]) * inv, __uint_as_float(o[16 * q + 5]) * inv);
            u0.w = pack_bf16(__uint_as_float(o[16 * q + 6]) * inv, __uint_as_float(o[16 * q + 7]) * inv);
            u1.x = pack_bf16(__uint_as_float(o[16 * q + 8]) * inv, __uint_as_float(o[16 * q + 9]) * inv);
            u1.y = pack_bf16(__uint_as_float(o[16 * q + 10]) * inv, __uint_as_float(o[16 * q + 11]) * inv);
            u1.z = pack_bf16(__uint_as_float(o[16 * q + 12]) * inv, __uint_as_float(o[16 * q + 13]) * inv);
            u1.w = pack_bf16(__uint_as_float(o[16 * q + 14]) * inv, __uint_as_float(o[16 * q + 15]) * inv);
            stg256_u(orow + c * 32 + q * 16, u0, u1);
          }
        }
      }
      tc_fence_before();   // the O reads above are ordered before the next item's p_full arrive (-> PV overwrite)
      TICK(7);   // final wait + O store
    }
#if VV_ATTN_TIMING
    tacc[9] = g;
    if (lane == 0 && blockIdx.x % 37 == 0)
      for (int i = 0; i < 10; ++i) atomicAdd(&g_attn4_timing[i], (unsigned long long)tacc[i]);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

#if VV_ATTN_TIMING
extern "C" void vv_attn_timing_dump() {
  unsigned long long h[12];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_attn4_timing, sizeof(h));
  const char* names[9] = {"wait_S", "tmem_ld", "mask/first max", "exp2+max+pack", "-", "wait_PV", "P store+arrive",
                          "final+O store", "wait MUFU turn"};
  double tot = 0;
  for (int i = 0; i < 9; ++i) tot += double(h[i]);
  for (int i = 0; i < 9; ++i)
    printf("  %-16s %14llu  %5.1f%%  %8.0f cyc/kv-iter\n", names[i], h[i], 100.0 * double(h[i]) / (tot + 1e-9),
           double(h[i]) / double(h[9] ? h[9] : 1));
  printf("  %.0f cycles per kv iteration per warp (%llu iterations sampled)\n", tot / double(h[9] ? h[9] : 1), h[9]);
  unsigned long long z[12] = {0};
  cudaMemcpyToSymbol(g_attn4_timing, z, sizeof(z));
}
#endif

void launch_attention4(const CUtensorMap& tmQKV, const AttnParams& p, cudaStream_t st) {
  using namespace attn4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    attr_set = true;
  }
  const int total = p.n_tiles * p.heads;
  if (total <= 0) return;
  int grid = p.num_sms > 0 ? p.num_sms : 148;
  if (grid > (total + 1) / 2) grid = (total + 1) / 2;
  while ((total + 2 * grid - 1) / (2 * grid) > MAX_ITEMS) grid *= 2;   // very long work lists: more CTAs than SMs
  attn4_kernel<<<grid, THREADS, SMEM, st>>>(tmQKV, p);
}

}  // namespace vv
