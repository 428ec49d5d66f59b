// Micro-benchmark: issue rate / execution rate of tcgen05.mma (kind::f16, SS mode, cta_group::1) on operands that
// are already resident in shared memory.  One CTA per SM, REPS x (K_STEPS MMAs + commit + wait).  Prints cycles per MMA
// for N = 32, 64, 128, 256.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
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

template <int N>
__global__ void __launch_bounds__(128, 1) k(int reps, int ksteps, long long* out, int sbo, int shift_rows, const uint8_t* gsrc, int copy_kb) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t cbar[2];
  __shared__ volatile int done_flag;
  __shared__ long long copied;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (65536 + N * 128) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    done_flag = 0;
    copied = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(N < 32 ? 32 : N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint64_t ad = desc(smem_u32(smem) + shift_rows * 128, sbo), bd = desc(smem_u32(smem + 65536));
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
        for (int s = 0; s < ksteps; ++s) {
          const uint64_t a = ad + 2 * (s & 3) + (uint64_t)(((s >> 2) % 9) * 8), b = bd + 2 * (s & 3);   // tap-like shifted views
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)(s != 0)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(r & 1)) : "memory");
    }
    t1 = clock64();
    if (threadIdx.x == 0) done_flag = 1;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp == 1 && copy_kb > 0 && threadIdx.x == 32) {
    // stream bulk copies global -> shared (two 16 KB slots in flight) next to the MMAs
    uint8_t* cdst = smem + 65536 + N * 128;
    const uint32_t bytes = 16384;
    long long n = 0;
    uint32_t it = 0;
    const uint8_t* src = gsrc + (size_t)blockIdx.x * 65536;
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[s])), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(cdst + s * bytes)), "l"(src + s * bytes), "r"(bytes), "r"(smem_u32(&cbar[s])) : "memory");
    }
    while (!done_flag) {
      const int s = it & 1;
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&cbar[s])), "r"((it >> 1) & 1u) : "memory");
      ++n;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[s])), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(cdst + s * bytes)), "l"(src + ((it + 2) & 3) * bytes), "r"(bytes), "r"(smem_u32(&cbar[s])) : "memory");
      ++it;
    }
    // drain
    for (int s = 0; s < 2; ++s) {
      uint32_t ok = 0;
      const uint32_t ph = ((it + s) >> 1) & 1u;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&cbar[(it + s) & 1])), "r"(ph) : "memory");
    }
    if (blockIdx.x == 0) out[1] = n;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N < 32 ? 32 : N) : "memory");
}

template <int N>
void run(int grid, int sbo = 1024, int shift = 0, int copy_kb = 0) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  static uint8_t* gsrc = nullptr;
  if (!gsrc) { cudaMalloc(&gsrc, 148 * 65536); cudaMemset(gsrc, 0, 148 * 65536); }
  const int smem = 65536 + N * 128 + 32768 + 2048;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 200, ksteps = 288;
  k<N><<<grid, 128, smem>>>(reps, ksteps, d, sbo, shift, gsrc, copy_kb);
  k<N><<<grid, 128, smem>>>(reps, ksteps, d, sbo, shift, gsrc, copy_kb);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<N><<<grid, 128, smem>>>(reps, ksteps, d, sbo, shift, gsrc, copy_kb);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long both[2]; cudaMemcpy(both, d, 16, cudaMemcpyDeviceToHost);
  long long cyc = both[0];
  const double per = (double)cyc / ((double)reps * ksteps);
  const double tf = 2.0 * 128 * N * 16 * (double)reps * ksteps * grid / (ms * 1e-3) / 1e12;
  printf("N=%3d grid=%3d sbo=%d shift=%d copies=%lld (%.1f B/clk): %s  %.1f cycles/MMA (floor %d), %.0f TFLOP/s, clock %.0f MHz\n", N, grid, sbo, shift, both[1], (double)both[1] * 16384.0 / (double)cyc, cudaGetErrorString(e), per,
         128 * N / 256, tf, (double)cyc / (ms * 1e-3) / 1e6);
  cudaFree(d);
}

int main() {
  run<128>(148, 3072, 1, 0); run<128>(148, 3072, 1, 1); run<64>(148, 3072, 1, 0); run<64>(148, 3072, 1, 1); run<256>(148, 1024, 0, 1); run<32>(148, 3072, 1, 1);
  return 0;
}
