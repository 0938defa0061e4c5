// Issue rates of the instructions of the attention softmax loop, per scheduler, for 1 / 2 / 4 warps per scheduler:
//   mode 0  F2FP.BF16.F32.PACK_AB alone (cvt.rn.bf16x2.f32)
//   mode 1  the loop body of attn_softmax.cuh per score pair: FFMA2 + 2 MUFU.EX2 + FADD2 + F2FP, consumers right behind
//           the MUFU pair the way ptxas schedules them (dependent)
//   mode 2  the same with the consumers of a pair issued 4 pairs later (register-rotated by hand in the source)
//   mode 3  PRMT-based truncating pack instead of F2FP (for comparison)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_pipe_bench.bin softmax_pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, float a) {
  float x[32];
  for (int i = 0; i < 32; ++i) x[i] = a * (threadIdx.x + i);
  uint32_t acc = 0;
  float s0 = 0.f, s1 = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < 128; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) acc ^= pack(x[i], x[i + 1]);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) acc ^= __byte_perm(__float_as_uint(x[i]), __float_as_uint(x[i + 1]), 0x7632);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float y0 = fmaf(x[i], a, -1.0f), y1 = fmaf(x[i + 1], a, -1.0f);
        const float e0 = ex2(y0), e1 = ex2(y1);
        s0 += e0; s1 += e1;
        acc ^= pack(e0, e1);
      }
    } else {
      float e[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float y0 = fmaf(x[i], a, -1.0f), y1 = fmaf(x[i + 1], a, -1.0f);
        e[i] = ex2(y0); e[i + 1] = ex2(y1);
        if (i >= 8) { s0 += e[i - 8]; s1 += e[i - 7]; acc ^= pack(e[i - 8], e[i - 7]); }
      }
#pragma unroll
      for (int i = 24; i < 32; i += 2) { s0 += e[i]; s1 += e[i + 1]; acc ^= pack(e[i], e[i + 1]); }
    }
    for (int i = 0; i < 32; ++i) x[i] += 1e-3f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(s0 + s1);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const char* names[4] = {"F2FP pack alone (16 per iteration)", "softmax pair, dependent (16 pairs)", "softmax pair, consumers 4 pairs late", "PRMT pack alone (16 per iteration)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, 1e-3f);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, 1e-3f);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, 1e-3f);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, cyc, 1e-3f);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-40s warps/scheduler %d : %8lld cycles = %6.2f cycles per pair per warp, %6.2f per pair per scheduler\n", names[mode], warps / 4, h,
             double(h) / (128.0 * 16.0), double(h) / (128.0 * 16.0 * (warps / 4)));
    }
  return 0;
}
