// Fused GEMM epilogue shared by the 1-CTA (gemm.cu) and 2-CTA pair (gemm2.cu) kernels: bias, RoPE, GELU/Mish,
// AdaLN gate, fp32 residual, row mask, fp32 / bf16 stores — the element-wise nodes ONNX Runtime runs around each
// MatMul inside `transformer.onnx` (/root/reference/vietvoicetts/core/tts_engine.py:161-172).
#pragma once
#include "kernels.h"
#include "ptx.cuh"

namespace vv {

__device__ __forceinline__ float gelu_tanh_f(float x) {
  float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}
// the same on a pair of values with packed FMUL2 / FFMA2: 5 FMA-pipe issues + 2 MUFU per pair instead of 12 + 2
__device__ __forceinline__ void gelu_tanh_f2(float& x0, float& x1) {
  float q0, q1, w0, w1, u0, u1, h0, h1, t0, t1;
  fmul2(q0, q1, x0, x1, x0, x1);
  ffma2(w0, w1, q0, q1, 0.7978845608028654f * 0.044715f, 0.7978845608028654f);
  fmul2(u0, u1, w0, w1, x0, x1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  fmul2(h0, h1, x0, x1, 0.5f, 0.5f);
  ffma2v(x0, x1, h0, h1, t0, t1, h0, h1);
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }
__device__ __forceinline__ float mish_f(float x) {
  float ex = __expf(fminf(x, 20.0f));
  float n = ex * (ex + 2.0f);
  return x * __fdividef(n, n + 2.0f);
}

// sb / sg: this chunk's 32 bias / gate values staged in shared memory by the warp at tile start (a global __ldg
// right before use cost one exposed L2 latency per chunk for each of them: the L1 is thrashed by the residual stream)
__device__ __forceinline__ void epilogue_chunk(const GemmEpi& e, int N, int row, int n0, const uint32_t (&raw)[32],
                                               const float4 (&rpre)[8], const float4* sb, const float4* sg,
                                               bool wide) {
  int nvalid = N - n0;
  if (nvalid <= 0) return;
  const bool full = nvalid >= 32;
  if (!full && nvalid > 32) nvalid = 32;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);

  if (e.bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = sb[j];
        fadd2(v[4 * j], v[4 * j + 1], b.x, b.y);
        fadd2(v[4 * j + 2], v[4 * j + 3], b.z, b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += __ldg(e.bias + n0 + j);
    }
  }
  if (e.rope_dim > 0) {
    const bool in_q = n0 < e.rope_dim;
    const bool in_k = n0 >= e.rope_off2 && n0 < e.rope_off2 + e.rope_dim;
    if (in_q || in_k) {
      const int pos = e.row_pos[row];
      const float2* cs = e.rope_cs + (size_t)pos * 32 + ((n0 & 63) >> 1);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 t = __ldg(cs + i);
        float x0 = v[2 * i], x1 = v[2 * i + 1];
        v[2 * i] = x0 * t.x - x1 * t.y;
        v[2 * i + 1] = x1 * t.x + x0 * t.y;
      }
    }
  }
  if (e.act == ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_tanh_f2(v[j], v[j + 1]);
  } else if (e.act == ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
  } else if (e.act == ACT_MISH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = mish_f(v[j]);
  }
  if (e.gate) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g = sg[j];
        v[4 * j] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] *= __ldg(e.gate + n0 + j);
    }
  }
  if (e.resid) {
    const float* r = e.resid + (size_t)row * e.ld_resid + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = rpre[j];   // prefetched by the caller before the accumulator was ready
        v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += r[j];
    }
  }
  if (e.row_mask && e.row_mask[row] == 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.0f;
  }
  if (e.out_f32) {
    float* o = e.out_f32 + (size_t)row * e.ld_f32 + n0;
    if (full && wide) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        stg256(o + 8 * j, make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
               make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));
    } else if (full) {
      float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) o[j] = v[j];
    }
  }
  if (e.out_bf16) {
    bf16* o = e.out_bf16 + (size_t)row * e.ld_bf16 + n0;
    if (full && wide) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint4 u0, u1;
        u0.x = pack_bf16(v[16 * j], v[16 * j + 1]);       u0.y = pack_bf16(v[16 * j + 2], v[16 * j + 3]);
        u0.z = pack_bf16(v[16 * j + 4], v[16 * j + 5]);   u0.w = pack_bf16(v[16 * j + 6], v[16 * j + 7]);
        u1.x = pack_bf16(v[16 * j + 8], v[16 * j + 9]);   u1.y = pack_bf16(v[16 * j + 10], v[16 * j + 11]);
        u1.z = pack_bf16(v[16 * j + 12], v[16 * j + 13]); u1.w = pack_bf16(v[16 * j + 14], v[16 * j + 15]);
        stg256_u(o + 16 * j, u0, u1);
      }
    } else if (full) {
      uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16(v[8 * j], v[8 * j + 1]);
        u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
        o4[j] = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) o[j] = __float2bfloat16(v[j]);
    }
  }
}

}  // namespace vv
