// Micro-benchmark: execution rate of tcgen05.mma.cta_group::2 (kind::f16, SS mode, M = 256 over a CTA pair) on operands
// already resident in shared memory.  One pair per TPC (74 clusters of 2), the leader issues REPS x (K_STEPS MMAs + commit
// + wait).  Prints cycles per MMA for N = 64, 128, 256 next to the per-SM floor (128 x N x 16 MACs at 8 192 MAC/clk/SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_pair mma_rate_pair.cu && ./mma_rate_pair
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc(uint32_t a, uint32_t sbo = 1024) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int N, int ACCS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int reps, int ksteps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (65536 + (N / 2) * 128) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  if (warp == 0 && rank == 0) {
    const uint64_t ad = desc(smem_u32(smem) + 128, 3072), bd = desc(smem_u32(smem + 65536));
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
        for (int s = 0; s < ksteps; ++s) {
          const uint64_t a = ad + 2 * (s & 3) + (uint64_t)(((s >> 2) % 9) * 8), b = bd + 2 * (s & 3);
          const uint32_t d = tmem + (uint32_t)((s % ACCS) * N);      // ACCS accumulators in rotation (no back-to-back dependency)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)(s >= ACCS)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
      }
      __syncwarp();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(r & 1)) : "memory");
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N, int ACCS>
void run(int grid) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const int smem = 65536 + (N / 2) * 128 + 2048;
  cudaFuncSetAttribute(k<N, ACCS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 200, ksteps = 288;
  k<N, ACCS><<<grid, 128, smem>>>(reps, ksteps, d);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<N, ACCS><<<grid, 128, smem>>>(reps, ksteps, d);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)cyc / ((double)reps * ksteps);
  const double tf = 2.0 * 256 * N * 16 * (double)reps * ksteps * (grid / 2) / (ms * 1e-3) / 1e12;
  printf("pair M=256 N=%3d accumulators=%d grid=%3d: %s  %.1f cycles/MMA (floor %d), %.0f TFLOP/s\n", N, ACCS, grid,
         cudaGetErrorString(e), per, 128 * N / 256, tf);
  cudaFree(d);
}

int main() {
  run<64, 1>(148); run<64, 2>(148); run<128, 1>(148); run<128, 2>(148); run<128, 3>(148); run<256, 1>(148); run<256, 2>(148);
  return 0;
}
