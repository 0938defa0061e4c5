// Softmax inner loop of the attention kernel (attn.cu).
#pragma once
#include "ptx.cuh"


namespace vv {

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
                                            float& sum) {
  float sa0 = 0.0f, sa1 = 0.0f, sb0 = 0.0f, sb1 = 0.0f;   // two packed (FADD2) row-sum chains
  auto val = [&](int i) { return (MASKED && i >= kv_valid) ? __uint_as_float(0xff800000u) : __uint_as_float(s[i]); };
  uint32_t pk_all[4][16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t (&pk)[16] = pk_all[c];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const int k = c * 32 + i;
      float x0, x1, x2, x3;
      ffma2(x0, x1, val(k), val(k + 1), scale_log2, -m);
      ffma2(x2, x3, val(k + 2), val(k + 3), scale_log2, -m);
      const float e0 = fast_exp2(x0);
      const float e1 = fast_exp2(x1);
      const float e2 = fast_exp2(x2);
      const float e3 = fast_exp2(x3);
      fadd2(sa0, sa1, e0, e1);
      fadd2(sb0, sb1, e2, e3);
      pk[i / 2] = pack_bf16(e0, e1);
      pk[i / 2 + 1] = pack_bf16(e2, e3);
    }
    if (c == 1) {
      if (pv_wait) {
        mbar_wait(pv_bar, pv_parity);          // P buffer free again, O quiescent
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
