// How fast can the SMs be FILLED with GEMM operand tiles, and does TMA multicast across a 4-CTA cluster help?
// Emulates the operand traffic of the CTA-pair GEMM (gemm2.cu) at the bench shape without any MMA: per k-step every
// CTA receives a 128x64 bf16 A box and a 128x64 B box (32 KB) into a 6-stage ring; a consumer thread holds each
// stage for `delay` clocks (the 512 cycles four M256 N256 K16 MMAs take) and releases it.
//   mode 0  clusters of 2, unicast (what gemm_pair_kernel does)
//   mode 1  clusters of 4 = two pairs on adjacent row blocks sharing the B tile: every CTA loads 64 B rows and
//           multicasts them to the CTA of the other pair that needs the same half
//   mode 2  clusters of 4 = two pairs on adjacent column tiles sharing the A rows (A multicast)
// Prints time, delivered bytes/clk/SM, the share of time the consumer waited for data, SMs used.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_fill_bench.bin tma_fill_bench.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <set>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int STAGES = 6;
constexpr int BOX = 128 * 64 * 2;          // 16 KB
constexpr int STAGE_BYTES = 2 * BOX;
constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void tma_ld(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(dst)), "l"((uint64_t)m), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_ld_mc(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(s32(dst)), "l"((uint64_t)m), "r"(s32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct Stats { unsigned long long total, wait; unsigned int smid, pad; };

// tmA: box 128 rows, tmAh: box 64 rows (same matrix); likewise tmB / tmBh
template <int CL>
__global__ void __launch_bounds__(64, 1)
fill_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAh,
            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBh, int M, int N, int K,
            int mode, int delay, Stats* stats) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const uint32_t rank = CL > 1 ? ctarank() : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool mc = mode != 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], mc ? 2 : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (CL > 1) cluster_sync();
  // mode 0 works in pairs whatever the cluster size is (cluster 4 + mode 0 = the stranding effect alone)
  const int cid = mode == 0 ? blockIdx.x / 2 : blockIdx.x / CL, ncl = mode == 0 ? gridDim.x / 2 : gridDim.x / CL;
  const int r2 = blockIdx.x & 1;
  const int kiters = K / 64;
  // units: mode 0: (m2, n) 256x256; mode 1: (m4, n) 512x256; mode 2: (m2, n2) 256x512
  const int m_units = mode == 1 ? (M + 511) / 512 : (M + 255) / 256;
  const int n_units = mode == 2 ? N / 512 : N / 256;
  const int units = m_units * n_units;
  const uint32_t p = rank >> 1, h = rank & 1;
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int u = cid; u < units; u += ncl) {
      const int mu = u / n_units, nu = u % n_units;
      for (int kb = 0; kb < kiters; ++kb) {
        mbar_wait(&empty[stage], ph ^ 1);
        mbar_expect(&full[stage], STAGE_BYTES);
        uint8_t* sA = smem + stage * STAGE_BYTES;
        uint8_t* sB = sA + BOX;
        if (mode == 0) {
          tma_ld(sA, &tmA, kb * 64, (2 * mu + r2) * 128, &full[stage]);
          tma_ld(sB, &tmB, kb * 64, nu * 256 + r2 * 128, &full[stage]);
        } else if (mode == 1) {
          tma_ld(sA, &tmA, kb * 64, (4 * mu + (int)rank) * 128, &full[stage]);
          tma_ld_mc(sB + p * (BOX / 2), &tmBh, kb * 64, nu * 256 + (int)h * 128 + (int)p * 64, &full[stage],
                    (uint16_t)((1u << h) | (1u << (h + 2))));
        } else {
          tma_ld_mc(sA + p * (BOX / 2), &tmAh, kb * 64, (2 * mu + (int)h) * 128 + (int)p * 64, &full[stage],
                    (uint16_t)((1u << h) | (1u << (h + 2))));
          tma_ld(sB, &tmB, kb * 64, (2 * nu + (int)p) * 256 + (int)h * 128, &full[stage]);
        }
        if (++stage == STAGES) { stage = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t ph = 0;
    long long t_wait = 0;
    const long long t0 = clock64();
    for (int u = cid; u < units; u += ncl) {
      for (int kb = 0; kb < kiters; ++kb) {
        const long long a = clock64();
        mbar_wait(&full[stage], ph);
        const long long b = clock64();
        t_wait += b - a;
        while (clock64() - b < delay) { }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[stage])) : "memory");
        if (mc) mbar_arrive_remote(s32(&empty[stage]), rank ^ 2);
        if (++stage == STAGES) { stage = 0; ph ^= 1; }
      }
    }
    Stats s;
    s.total = clock64() - t0; s.wait = t_wait;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s.smid));
    s.pad = 0;
    stats[blockIdx.x] = s;
  }
  __syncthreads();
  if (CL > 1) cluster_sync();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap mk(EncodeFn enc, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int CL>
static void run(const char* name, int mode, int delay, const CUtensorMap& a, const CUtensorMap& ah, const CUtensorMap& b,
                const CUtensorMap& bh, int M, int N, int K, Stats* d_stats, int nsm) {
  auto kern = fill_kernel<CL>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  if (CL > 1) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(nsm / CL * CL);
  int maxcl = 0;
  CK(cudaOccupancyMaxActiveClusters(&maxcl, kern, &cfg));
  int ncl = maxcl < nsm / CL ? maxcl : nsm / CL;
  cfg.gridDim = dim3(ncl * CL);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 5;
  float best = 1e9f;
  for (int r = 0; r < reps + 1; ++r) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, a, ah, b, bh, M, N, K, mode, delay, d_stats));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
  }
  std::vector<Stats> hs(ncl * CL);
  CK(cudaMemcpy(hs.data(), d_stats, hs.size() * sizeof(Stats), cudaMemcpyDeviceToHost));
  double tot = 0, wait = 0; std::set<unsigned> sms;
  for (auto& s : hs) { tot += s.total; wait += s.wait; sms.insert(s.smid); }
  const int m_units = mode == 1 ? (M + 511) / 512 : (M + 255) / 256;
  const int n_units = mode == 2 ? N / 512 : N / 256;
  const double steps = double(m_units) * n_units * (K / 64) * (mode == 0 ? 2 : CL);     // CTA k-steps
  const double bytes = steps * STAGE_BYTES;
  const double avg_clk = tot / hs.size();
  printf("%-34s delay %4d  max_active_clusters %3d  grid %3d  SMs %3zu  %8.1f us  %7.2f TB/s into smem  %6.1f B/clk/SM  "
         "consumer wait %5.1f%%  equivalent MMA rate %6.1f TFLOP/s\n",
         name, delay, maxcl, ncl * CL, sms.size(), best * 1e3, bytes / (best * 1e-3) / 1e12,
         bytes / hs.size() / avg_clk, 100.0 * wait / tot, 2.0 * M * N * K / (best * 1e-3) / 1e12);
}

int main(int argc, char** argv) {
  int M = 24320, N = 3072, K = 1024;
  if (argc > 3) { M = atoi(argv[1]); N = atoi(argv[2]); K = atoi(argv[3]); }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  __nv_bfloat16 *A, *B; Stats* st;
  CK(cudaMalloc(&A, (size_t)M * K * 2)); CK(cudaMalloc(&B, (size_t)N * K * 2)); CK(cudaMalloc(&st, 1024 * sizeof(Stats)));
  CK(cudaMemset(A, 0, (size_t)M * K * 2)); CK(cudaMemset(B, 0, (size_t)N * K * 2));
  CUtensorMap a = mk(enc, A, M, K, 128), ah = mk(enc, A, M, K, 64), b = mk(enc, B, N, K, 128), bh = mk(enc, B, N, K, 64);
  printf("M %d N %d K %d, %d SMs\n", M, N, K, nsm);
  const int delays[4] = {0, 384, 512, 640};
  for (int d = 0; d < 4; ++d) {
    run<2>("cluster 2 unicast", 0, delays[d], a, ah, b, bh, M, N, K, st, nsm);
    run<4>("cluster 4 unicast", 0, delays[d], a, ah, b, bh, M, N, K, st, nsm);
    run<4>("cluster 4 B multicast (512x256)", 1, delays[d], a, ah, b, bh, M, N, K, st, nsm);
    run<4>("cluster 4 A multicast (256x512)", 2, delays[d], a, ah, b, bh, M, N, K, st, nsm);
  }
  return 0;
}
