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
#define TICK(i) do { long long _t = clock64(); tacc[i] += _t - tlast; tlast = _t; } while (0)
#else
#define TICK(i) do { } while (0)
#endif

namespace vv {

#if VV_ATTN_TIMING
__device__ unsigned long long g_attn_timing[10];
#endif
#ifndef VV_ATTN_TRACE
#define VV_ATTN_TRACE 0
#endif
#if VV_ATTN_TRACE
// exp2-phase trace of every CTA that runs on ONE SM (smid = VV_ATTN_TRACE - 1): (cta, warp, kv tile, start, end)
__device__ long long g_attn_trace[5 * 16384];
__device__ int g_attn_trace_n;
#endif

namespace attn {
constexpr int K_STAGES = 2;
constexpr int V_STAGES = 2;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int Q_OFF = 0;
constexpr int K_OFF = Q_OFF + TILE_BYTES;
constexpr int V_OFF = K_OFF + K_STAGES * TILE_BYTES;
constexpr int BAR_OFF = V_OFF + V_STAGES * TILE_BYTES;
constexpr int SMEM = BAR_OFF + 256 + 1024;
constexpr int THREADS = 256;   // 2 full warpgroups (setmaxnreg is warpgroup-aligned): softmax | TMA, MMA, 2 idle warps
constexpr uint32_t TM_S = 0;
constexpr uint32_t TM_P = 128;
constexpr uint32_t TM_O = 192;
constexpr uint32_t TM_COLS = 256;
constexpr float SUM_LIMIT = 1.099511627776e12f;   // 2^40: row sum of one kv tile against the current reference
}  // namespace attn

__global__ void __launch_bounds__(attn::THREADS, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  using namespace attn;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // K_STAGES
  uint64_t* k_empty = k_full + K_STAGES;
  uint64_t* v_full = k_empty + K_STAGES;   // V_STAGES
  uint64_t* v_empty = v_full + V_STAGES;
  uint64_t* s_full = v_empty + V_STAGES;   // S(j) accumulator complete             (MMA -> softmax)
  uint64_t* s_free = s_full + 1;           // S(j) copied to registers               (softmax -> MMA)
  uint64_t* p_full = s_free + 1;           // P(j) in TMEM, O rescaled               (softmax -> MMA)
  uint64_t* pv_done = p_full + 1;          // O += P(j) V(j) complete                (MMA -> softmax)
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

  // The TMA thread initialises the barriers itself and issues Q and the first K/V stages at once: their L2/HBM
  // latency then runs under the TMEM allocation and the CTA-wide barrier instead of after them.
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

  if (warp >= 4) {
    setmaxnreg_dec<48>();       // 128 x 208 + 128 x 48 = 256 x 128 registers: two such CTAs fill the SM's register file
    if (warp == 4) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int j = 0; j < n_kv; ++j) {
          if (j >= n_pre) {                    // the first n_pre stages were issued before the CTA barrier
            mbar_wait(&k_empty[ks], kph ^ 1);
            mbar_expect_tx(&k_full[ks], TILE_BYTES);
            tma_load_2d(smem + K_OFF + ks * TILE_BYTES, &tmQKV, p.dim + head * 64, seq_row0 + j * 128, &k_full[ks]);
          }
          if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
          if (j >= n_pre) {
            mbar_wait(&v_empty[vs], vph ^ 1);
            mbar_expect_tx(&v_full[vs], TILE_BYTES);
            tma_load_2d(smem + V_OFF + vs * TILE_BYTES, &tmQKV, 2 * p.dim + head * 64, seq_row0 + j * 128, &v_full[vs]);
          }
          if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
        }
      }
    } else if (warp == 5) {
      // ------------------------------------------------------------------ MMA issuer.  S(n+1) is issued as soon as the
      // softmax warps have copied S(n) to registers, i.e. it runs under softmax(n).
      if (lane == 0) {
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0);
        constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 1);
        const uint32_t q_addr = smem_u32(smem + Q_OFF);
        const uint32_t k_addr = smem_u32(smem + K_OFF);
        const uint32_t v_addr = smem_u32(smem + V_OFF);
        const uint32_t d_s = tmem_base + TM_S;
        const uint32_t d_o = tmem_base + TM_O;
        mbar_wait(q_full, 0);
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int n = 0; n <= n_kv; ++n) {
          if (n < n_kv) {
            if (n > 0) mbar_wait(s_free, (n - 1) & 1);
            mbar_wait(&k_full[ks], kph);
            tc_fence_after();
            const uint64_t a0 = make_sdesc_sw128(q_addr);
            const uint64_t b0 = make_sdesc_sw128(k_addr + ks * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(d_s, a0 + 2 * k, b0 + 2 * k, idesc_s, k != 0);
            umma_commit(s_full);
            umma_commit(&k_empty[ks]);
            if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
          }
          if (n > 0) {
            mbar_wait(p_full, (n - 1) & 1);
            mbar_wait(&v_full[vs], vph);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t b = make_sdesc_sw128(v_addr + vs * TILE_BYTES + k * 2048);
              umma_ts(d_o, tmem_base + TM_P + k * 8, b, idesc_o, !(n == 1 && k == 0));
            }
            umma_commit(pv_done);
            umma_commit(&v_empty[vs]);
            if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
          }
        }
      }
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
#if VV_ATTN_TIMING
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif
    bool s_ready = false;                      // S(j) already seen complete by the probe inside softmax(j-1)
    for (int j = 0; j < n_kv; ++j) {
      if (!s_ready) mbar_wait(s_full, j & 1);
#if VV_ATTN_TIMING
      { long long _t = clock64(); tacc[j == 0 ? 0 : 5] += _t - tlast; tlast = _t; }   // wait S: first tile | later tiles
#endif
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(ts + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);                     // S may be overwritten by S(j+1) from here on
      TICK(1);   // tmem ld
      const int kv_valid = kv_len - j * 128;
      const bool partial = kv_valid < 128;     // warp-uniform: last kv tile of a sequence whose length is not k*128
      // exact row maximum of the tile in registers (4 independent 3-input chains); only the first tile of an item and
      // the rare re-reference path need it
      auto row_max = [&]() {
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
        return fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd)) * p.scale_log2;
      };
      if (j == 0) m_ref = row_max();           // first tile: no reference yet
      TICK(2);   // first-tile max
#if VV_ATTN_TRACE
      const long long tr0 = clock64();
#endif
      float sum;
      s_ready = false;
      uint64_t* nb = j + 1 < n_kv ? s_full : nullptr;
      if (partial)
        softmax_row<true>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, (j - 1) & 1, j > 0, sum, nb, (j + 1) & 1,
                          nb ? &s_ready : nullptr);
      else
        softmax_row<false>(s, p.scale_log2, m_ref, kv_valid, tp, pv_done, (j - 1) & 1, j > 0, sum, nb, (j + 1) & 1,
                           nb ? &s_ready : nullptr);
      TICK(3);   // exp2 + pack + P store
#if VV_ATTN_TRACE
      {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (lane == 0 && smid == VV_ATTN_TRACE - 1) {
          const int k = atomicAdd(&g_attn_trace_n, 1);
          if (k < 16384) {
            g_attn_trace[5 * k] = blockIdx.x; g_attn_trace[5 * k + 1] = warp; g_attn_trace[5 * k + 2] = j;
            g_attn_trace[5 * k + 3] = tr0; g_attn_trace[5 * k + 4] = clock64();
          }
        }
      }
#endif
      if (j > 0 && __any_sync(0xffffffffu, !(sum < SUM_LIMIT))) {
        // some row of this warp outgrew the reference by more than 2^40 (or overflowed): take the true maximum as the
        // new reference, redo this tile, rescale O and l
        const float m_new = fmaxf(m_ref, row_max());
        const float f = fast_exp2(m_ref - m_new);
        m_ref = m_new;
        tmem_st_wait();                        // first-pass P stores retired before the same columns are rewritten
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
#if VV_ATTN_TIMING
        tacc[8] += 1;
#endif
      }
      TICK(4);   // redo (reference outgrown)
      tmem_st_wait();
      l += sum;
      tc_fence_before();
      mbar_arrive(p_full);
      TICK(6);   // O rescale + P store + arrive
    }
    // ---- finalize: O / l -> bf16
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
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
    TICK(7);   // final wait + O store
    tacc[9] = n_kv;
    if (lane == 0 && blockIdx.x % 97 == 0)
      for (int i = 0; i < 10; ++i) atomicAdd(&g_attn_timing[i], (unsigned long long)tacc[i]);
#endif
  }
#if VV_ATTN_TIMING
  if (warp < 4) {
    // (softmax threads only; tacc lives in their scope)
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, TM_COLS);
}

void launch_attention_impl(const CUtensorMap& tmQKV, const AttnParams& p, cudaStream_t st) {
  static DeviceOnce attr;
  attr.once([] { cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM); });
  if (p.n_tiles <= 0) return;
  launch_k(attn_kernel, p.n_tiles * p.heads, attn::THREADS, attn::SMEM, st, tmQKV, p);
}

#if VV_ATTN_TIMING
extern "C" void vv_attn_timing_dump() {
  unsigned long long h[10];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_attn_timing, sizeof(h));
  const char* names[8] = {"wait_S (j=0)", "tmem_ld", "mask/first max", "exp2+max+pack", "redo", "wait_S (j>0)", "P store+arrive",
                          "final+O store"};
  double tot = 0;
  for (int i = 0; i < 8; ++i) tot += double(h[i]);
  for (int i = 0; i < 8; ++i)
    printf("  %-16s %14llu  %5.1f%%  %8.0f cyc/kv-iter\n", names[i], h[i], 100.0 * double(h[i]) / (tot + 1e-9),
           double(h[i]) / double(h[9] ? h[9] : 1));
  printf("  redo count %llu of %llu kv iterations (sampled warps); %.0f cycles per kv iteration per warp\n", h[8], h[9],
         tot / double(h[9] ? h[9] : 1));
  unsigned long long z[10] = {0};
  cudaMemcpyToSymbol(g_attn_timing, z, sizeof(z));
}
#endif

#if VV_ATTN_TRACE
extern "C" void vv_attn_trace_dump(const char* path) {
  static long long h[5 * 16384];
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_attn_trace_n, sizeof(n));
  cudaMemcpyFromSymbol(h, g_attn_trace, sizeof(h));
  if (n > 16384) n = 16384;
  FILE* f = fopen(path, "w");
  if (!f) return;
  fprintf(f, "cta,warp,kv,start,end\n");
  for (int i = 0; i < n; ++i)
    fprintf(f, "%lld,%lld,%lld,%lld,%lld\n", h[5 * i], h[5 * i + 1], h[5 * i + 2], h[5 * i + 3], h[5 * i + 4]);
  fclose(f);
  n = 0;
  cudaMemcpyToSymbol(g_attn_trace_n, &n, sizeof(n));
}
#endif

int attn_q_tile() { return 128; }
void launch_attention(const CUtensorMap& tmQKV, const AttnParams& p, cudaStream_t st) { launch_attention_impl(tmQKV, p, st); }

}  // namespace vv
