// Latency of mbarrier operations as seen by one warp (clock64 around each, 1 warp per CTA, 1 CTA).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __noinline__ long long clk_after(uint32_t dep) {
  // the branch on `dep` has to resolve before either clock read can issue
  if (dep == 0x7fffffffu) return clock64() + 1;
  return clock64();
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(long long* out) {
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[i])), "r"(i == 1 ? 32 : 1));
  }
  __syncthreads();
  long long acc[6] = {0, 0, 0, 0, 0, 0};
  uint32_t ph = 0;
  for (int it = 0; it < 64; ++it) {
    long long t0 = clock64();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[0])) : "memory");
    long long t1 = clock64();
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(s32(&bar[0])), "r"(ph) : "memory");
    long long t2 = clk_after(ok);
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(s32(&bar[0])), "r"(ph) : "memory");
    long long t3 = clk_after(ok);
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(s32(&bar[0])), "r"(ph) : "memory");
    long long t4 = clk_after(ok);
    // all 32 lanes arrive on a count-32 barrier
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[1])) : "memory");
    long long t5 = clock64();
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(s32(&bar[2])), "r"(1u) : "memory");  // fresh barrier, parity 1: complete
    long long t6 = clk_after(ok);
    acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4; acc[5] += t6 - t5;
    ph ^= 1;
    if (ok == 77) out[100] = 1;
  }
  if (threadIdx.x == 0) for (int i = 0; i < 6; ++i) out[i] = acc[i] / 64;
}
int main() {
  long long* d; cudaMalloc(&d, 1024); cudaMemset(d, 0, 1024);
  k<<<1, 32>>>(d); cudaDeviceSynchronize();
  k<<<1, 32>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  const char* n[6] = {"arrive (1 lane) issue", "try_wait right after arrive (completes)", "try_wait again (already complete)",
                      "test_wait (already complete)", "arrive (32 lanes, same word) issue", "try_wait on another complete barrier after it"};
  for (int i = 0; i < 6; ++i) printf("%-50s %lld cycles\n", n[i], h[i]);
  printf("\n");
  return 0;
}
