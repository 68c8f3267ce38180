// Micro-benchmark: latency / throughput of tcgen05.mma.kind::tf32 chains on sm_100a (one CTA per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_latency scripts/umma_latency.cu && /tmp/umma_latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
template <int BN>
__global__ void __launch_bounds__(128, 1) bench(int n_mma, int n_rep, long long *out, int same_acc) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < (16384 + BN * 128) / 4; i += 128) ((float *)smem)[i] = 1.0f;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | (8u << 24);
  if (threadIdx.x == 0) {
    const uint64_t ad = make_desc(smem_u32(smem)), bd = make_desc(smem_u32(smem + 16384));
    long long total = 0;
    uint32_t parity = 0;
    for (int r = 0; r < n_rep; ++r) {
      long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t d = tb + (same_acc ? 0 : (uint32_t)((i & 1) * BN));
        const uint32_t acc = 1;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(d), "l"(ad + (uint64_t)((i & 3) * 2)), "l"(bd + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      mbar_wait(&bar, parity); parity ^= 1;
      total += clock64() - t0;
    }
    out[blockIdx.x] = total / n_rep;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
template <int BN> void run(int n_mma, int same) {
  long long *d; cudaMalloc(&d, 148 * 8);
  size_t smem = 16384 + BN * 128 + 1024;
  cudaFuncSetAttribute(bench<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<BN><<<148, 128, smem>>>(n_mma, 20, d, same);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("BN=%3d n_mma=%3d same_acc=%d : %lld cycles per batch (%.1f per MMA)  [%s]\n", BN, n_mma, same, h[0], (double)h[0] / n_mma, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int same = 1; same >= 0; --same) {
    run<32>(1, same); run<32>(4, same); run<32>(24, same); run<32>(96, same);
    run<128>(1, same); run<128>(4, same); run<128>(24, same); run<128>(96, same);
    run<256>(1, same); run<256>(4, same); run<256>(24, same); run<256>(96, same);
  }
  return 0;
}
