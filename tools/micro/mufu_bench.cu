// MUFU.EX2 issue-rate microbenchmark: cycles per warp-level ex2 per SM sub-partition, for 1..4 warps per scheduler,
// alone and mixed with the FMA-pipe work of the softmax loop.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, float a) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = a * (threadIdx.x + i);
  float acc0 = 0, acc1 = 0;
  long long t0 = clock64();
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float e;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[i]));
      if (MODE == 1) { x[i] = fmaf(e, 0.5f, -1.0f); acc0 += e; }
      else x[i] = e;
      if (MODE == 2) { acc0 += e; acc1 = fmaf(e, a, acc1); x[i] = fmaf(x[i], 0.25f, -2.0f); }
    }
  }
  long long t1 = clock64();
  float s = acc0 + acc1;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 16; warps += 4) {   // warps per CTA, 1 CTA per SM -> warps/4 per scheduler
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, 1e-3f);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, 1e-3f);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, 1e-3f);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double per = double(h) / (256.0 * 16.0 * (warps / 4));
      printf("mode %d  warps/scheduler %d : %lld cycles, %.2f cycles per warp-level MUFU.EX2 per scheduler\n", mode, warps / 4, h, per);
    }
  return 0;
}
