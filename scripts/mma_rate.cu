// Micro-benchmark: issue rate of tcgen05.mma kind::f16 shapes on B200 (cycles per instruction, operands static in
// shared memory, one accumulator, back-to-back accumulate).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
//   -I pixel-nerf_b200/csrc scripts/mma_rate.cu -o gpurun_out/mma_rate ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include "pnr_tc_ptx.cuh"
using namespace pnr::tcptx;

template <int CTAS, int M, int N, int COMMIT_EVERY = 0, int STREAM = 0>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* out, const uint8_t* src) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(8) unsigned long long dummy[8];
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&dummy[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (warp == 1 && (CTAS == 1 || rank == 0)) {
    const uint64_t a = make_desc(smem_u32(smem)), b = make_desc(smem_u32(smem) + 32768);
    const bool issuer = elect_one();
    const long long t0 = clock64();
    if (issuer) {
      for (int i = 0; i < iters; ++i) {
        if (CTAS == 2) {
          umma_f16_2sm(tm, a, b, idesc, 1u); umma_f16_2sm(tm, a + 2, b + 2, idesc, 1u);
          umma_f16_2sm(tm, a + 4, b + 4, idesc, 1u); umma_f16_2sm(tm, a + 6, b + 6, idesc, 1u);
          if (COMMIT_EVERY == 4 || (COMMIT_EVERY == 8 && (i & 1))) umma_commit_pair(smem_u32(&dummy[i & 7]));
        } else {
          umma_f16_1sm(tm, a, b, idesc, 1u); umma_f16_1sm(tm, a + 2, b + 2, idesc, 1u);
          umma_f16_1sm(tm, a + 4, b + 4, idesc, 1u); umma_f16_1sm(tm, a + 6, b + 6, idesc, 1u);
        }
      }
      if (CTAS == 2) umma_commit_pair(smem_u32(&bar)); else umma_commit_local(smem_u32(&bar));
    }
    __syncwarp();
    int st = 0;
    mbar_wait(smem_u32(&bar), 0, &st, 1);
    const long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (STREAM && warp == 2 && threadIdx.x == 64) {
    // endless 16 KB bulk copies into 4 slots behind the operands until the MMA warp is done (bounded by iters)
    __shared__ __align__(8) unsigned long long sb[4];
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&sb[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int st = 0;
    const int n = iters * 4 * 64 / 400;   // about as long as the MMA loop at ~400 cycles per copy
    for (int i = 0; i < n + 4; ++i) {
      const int sl = i & 3;
      if (i >= 4) mbar_wait(smem_u32(&sb[sl]), ((i >> 2) - 1) & 1, &st, 2);
      if (i < n) {
        mbar_expect_tx(smem_u32(&sb[sl]), 16384);
        bulk_g2s(smem_u32(smem) + 65536 + sl * 8192 * 0 + (sl & 1) * 16384, src + (size_t)((i * 37 + blockIdx.x) % 600) * 16384, 16384, smem_u32(&sb[sl]));
      }
    }
  } else if (CTAS == 2 && warp == 1) {
    int st = 0;
    mbar_wait(smem_u32(&bar), 0, &st, 1);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

static uint8_t* g_src = nullptr;
template <int CTAS, int M, int N, int CE = 0, int ST = 0>
void run(const char* name, int grid) {
  if (!g_src) { cudaMalloc(&g_src, 10u << 20); cudaMemset(g_src, 0, 10u << 20); }
  long long* d; cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
  auto kern = k_rate<CTAS, M, N, CE, ST>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  const int iters = 2000;
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 98304;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {CTAS, 1, 1};
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, kern, iters, d, (const uint8_t*)g_src);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  double per = (double)c / (4.0 * iters);
  double macs = (double)M * N * 16 / per / (CTAS);
  printf("%-28s grid %3d: %8.1f cycles/MMA  -> %7.0f MAC/cycle/SM  (%s)\n", name, grid, per, macs, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<1, 128, 256>("1-CTA M128 N256", 148);
  run<1, 128, 128>("1-CTA M128 N128", 148);
  run<1, 64, 256>("1-CTA M64  N256", 148);
  run<2, 128, 256>("2-CTA M128 N256 (64/CTA)", 148);
  run<2, 256, 256>("2-CTA M256 N256 (128/CTA)", 148);
  run<1, 128, 16>("1-CTA M128 N16", 148);
  run<2, 128, 256, 8>("2-CTA M128 N256 commit/8", 148);
  run<2, 128, 256, 4>("2-CTA M128 N256 commit/4", 148);
  run<2, 128, 256, 8, 1>("2-CTA M128 N256 + bulk stream", 148);
  run<1, 128, 256, 0, 1>("1-CTA M128 N256 + bulk stream", 148);
  return 0;
}
