// Does the 16-bit exponential raise the MUFU rate?  VERDICT r1 asked for `ex2.approx.ftz.bf16x2` as a way past the
// MUFU bound of the attention kernel (P is rounded to bf16 anyway).  ptxas splits the packed form into TWO MUFU.EX2.BF16
// (one per half, `R.H1` operand) plus a PRMT, so the question is whether MUFU.EX2.BF16 / .F16 issue faster than the
// fp32 MUFU.EX2.  Cycles per warp-level MUFU per scheduler, 1..4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu16_bench.bin mufu16_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>   // 0: ex2.approx.ftz.f32, 1: ex2.approx.ftz.bf16x2, 2: ex2.approx.f16x2, 3: ex2.approx.ftz.bf16 (scalar)
__global__ void k(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t x[16];
  for (int i = 0; i < 16; ++i) x[i] = MODE == 0 ? __float_as_uint(-1e-3f * (threadIdx.x + i)) : (0xBC00BC00u ^ (seed + i));
  long long t0 = clock64();
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(x[i]) : "r"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(x[i]) : "r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(x[i]) : "r"(x[i]));
      if (MODE == 3) {
        uint16_t h = (uint16_t)x[i], r;
        asm volatile("ex2.approx.ftz.bf16 %0, %1;" : "=h"(r) : "h"(h));
        x[i] = r;
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 16; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const char* names[4] = {"ex2.approx.ftz.f32      (1 result  / instr)", "ex2.approx.ftz.bf16x2   (2 results / instr)",
                          "ex2.approx.f16x2        (2 results / instr)", "ex2.approx.ftz.bf16     (1 result  / instr)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps = 4; warps <= 16; warps += 4) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, 1);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, 1);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, 1);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, cyc, 1);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double per = double(h) / (256.0 * 16.0 * (warps / 4));
      const int res = (mode == 1 || mode == 2) ? 2 : 1;
      printf("%s warps/scheduler %d : %.2f cycles per PTX instr per scheduler = %.2f cycles per 32 results\n",
             names[mode], warps / 4, per, per / res);
    }
  return 0;
}
