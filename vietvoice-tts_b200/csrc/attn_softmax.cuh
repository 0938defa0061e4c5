// Softmax inner loop of the attention kernel (attn.cu).
#pragma once
#include "ptx.cuh"


#ifndef VV_ATTN_POLY8
#define VV_ATTN_POLY8 1     // of every 8 score pairs, this many take exp2 on the FMA pipe (0: use VV_ATTN_POLY of every 4)
#endif
#ifndef VV_ATTN_POLY
#define VV_ATTN_POLY 1      // of every 4 score pairs, this many take the FMA-pipe exp2 instead of MUFU.EX2
#endif

namespace vv {

// exp2 of a pair on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 magic add,
// degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max rel. error 1.0e-4, 40x below bf16 resolution), exponent
// re-attached with an integer shift-add.  Packed: 6 FFMA2 + 4 FMNMX + 2 shift-adds for two elements.
// The clamp keeps the exponent arithmetic from wrapping: below -125 the result is ~2^-125 (P rounds it to 0 in bf16
// terms of the row sum), above +126 it is ~2^126, which trips the row-sum overflow detector like MUFU's inf does.
__device__ __forceinline__ void poly_exp2_pair(float x0, float x1, float& e0, float& e1) {
  constexpr float MAGIC = 12582912.0f;     // 1.5 * 2^23
  x0 = fminf(fmaxf(x0, -125.0f), 126.0f);
  x1 = fminf(fmaxf(x1, -125.0f), 126.0f);
  float t0, t1, r0, r1, f0, f1, p0, p1;
  ffma2(t0, t1, x0, x1, 1.0f, MAGIC);      // low mantissa bits of t hold round(x)
  ffma2(r0, r1, t0, t1, 1.0f, -MAGIC);     // r = round(x)
  ffma2v(f0, f1, r0, r1, -1.0f, -1.0f, x0, x1);
  ffma2(p0, p1, f0, f1, 0.05500871316f, 0.24221068621f);
  ffma2v(p0, p1, p0, p1, f0, f1, 0.69328290224f, 0.69328290224f);
  ffma2v(p0, p1, p0, p1, f0, f1, 1.0f, 1.0f);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// exp2 of one 128-wide score row against reference `m`; P goes straight to TMEM as bf16 pairs, 32 elements (16
// columns) per tcgen05.st, so only one 16-register group of P is live at a time.  Returns the row sum in `sum`.
// Scale-and-subtract and the row sum use packed FFMA2 / FADD2.
// MASKED: columns >= kv_valid are keys past the end of the sequence (last kv tile only) and count as -inf.  The mask
// lives inside this variant so that full tiles run a loop without a single compare/select.
// The P columns are still being read by PV(j-1) when the loop starts: the first two groups are kept in registers
// and stored after the wait on `pv_bar` (by then PV(j-1) has long retired), the last two are stored as produced.
template <bool MASKED>
__device__ __forceinline__ void softmax_row(const uint32_t (&s)[128], float scale_log2, float m, int kv_valid,
                                            uint32_t tp, uint64_t* pv_bar, uint32_t pv_parity, bool pv_wait,
                                            float& sum, uint64_t* next_bar = nullptr, uint32_t next_parity = 0,
                                            bool* next_ready = nullptr) {
  float sa0 = 0.0f, sa1 = 0.0f, sb0 = 0.0f, sb1 = 0.0f;   // two packed (FADD2) row-sum chains
  auto val = [&](int i) { return (MASKED && i >= kv_valid) ? __uint_as_float(0xff800000u) : __uint_as_float(s[i]); };
  uint32_t pk_all[4][16];
  // both barriers this row will need next are probed early and without blocking (PV(j-1) here, S(j+1) two groups
  // further down): they completed long ago in the steady state, and the probe's latency runs under the exp2 stream
  const bool pv_ready = pv_wait ? mbar_test(pv_bar, pv_parity) : true;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t (&pk)[16] = pk_all[c];
    if (c == 2 && next_ready) *next_ready = mbar_test(next_bar, next_parity);
    if (MASKED && c * 32 >= kv_valid) {        // whole group past the end of the sequence (warp-uniform): P = 0,
#pragma unroll                                 // no exp2 issued at all
      for (int i = 0; i < 16; ++i) pk[i] = 0u;
    } else
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const int k = c * 32 + i;
      float x0, x1, x2, x3;
      ffma2(x0, x1, val(k), val(k + 1), scale_log2, -m);
      ffma2(x2, x3, val(k + 2), val(k + 3), scale_log2, -m);
      // pair slots 0..3 repeat every 8 elements; the LAST VV_ATTN_POLY slots of each group go to the FMA pipe
      float e0, e1, e2, e3;
      if (VV_ATTN_POLY8 ? (((i / 2) % 8) >= 8 - VV_ATTN_POLY8) : (((i / 4) % 2) * 2 + 0 >= 4 - VV_ATTN_POLY)) {
        poly_exp2_pair(x0, x1, e0, e1);
      } else {
        e0 = fast_exp2(x0);
        e1 = fast_exp2(x1);
      }
      if (VV_ATTN_POLY8 ? (((i / 2 + 1) % 8) >= 8 - VV_ATTN_POLY8) : (((i / 4) % 2) * 2 + 1 >= 4 - VV_ATTN_POLY)) {
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
        if (!pv_ready) mbar_wait(pv_bar, pv_parity);   // P buffer free again, O quiescent
        tc_fence_after();
      }
      tmem_st16(tp, pk_all[0]);
      tmem_st16(tp + 16, pk_all[1]);
    } else if (c > 1) {
      tmem_st16(tp + c * 16, pk);
    }
  }
  sum = (sa0 + sa1) + (sb0 + sb1);
}

}  // namespace vv
