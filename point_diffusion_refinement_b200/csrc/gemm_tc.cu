// tcgen05 (5th-gen tensor core) TF32 GEMM with fused prologue/epilogue for the denoiser -- sm_100a only.
//
//   C[M, N] = pro(A)[M, K] . W[N, K]^T + bias (+ rowadd)          fp32 in HBM, TF32 MMA, fp32 accumulate
//
// Same contract as the SIMT kernel in net.cu (PdrGemmArgs); this is the production path for the large
// grouped 1x1 convolutions (M = B*npoint*nsample rows).  Per CTA: one 128-row tile x BN columns.
//   * A cannot come through TMA: GroupNorm scale/shift + ReLU + per-sample embedding + residual are applied
//     to it on the way in.  All 256 threads load A (coalesced 128 B rows) and W chunks into registers,
//     transform, round to TF32 (cvt.rna) and store into shared memory in the canonical K-major
//     SWIZZLE_128B layout (16-byte chunk c of row r lands at chunk c ^ (r & 7)); fence.proxy.async makes the
//     generic-proxy stores visible to the tensor core.
//   * One thread issues tcgen05.mma.cta_group::1.kind::tf32 (128 x BN x 8 per instruction, 4 per 32-float
//     K chunk) with the accumulator in TMEM; tcgen05.commit arrives on an mbarrier that recycles the
//     shared-memory stage, so the loads of chunk k+1 overlap the MMAs of chunk k.
//   * Epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> bias / broadcast row-add -> global store,
//     plus the per-column (sum, sum^2, relu-sum, relu-sum^2) tile statistics consumed by the next GroupNorm.
// smem <= 97 KiB and TMEM <= 256 columns per CTA, so two CTAs share an SM and one's epilogue overlaps the
// other's loads.
#include "common.cuh"

namespace pdr {
namespace {

constexpr int kTcThreads = 256;
constexpr int kTcTileM = 128;
constexpr int kTcBK = 32;                       // floats per K chunk = one 128-byte swizzle row
constexpr int kATileBytes = kTcTileM * 128;     // 16 KiB

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor: start>>4 | LBO=1<<16 |
// SBO=(1024>>4)<<32 | version=1<<46 | layout SWIZZLE_128B=2<<61)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float pro1(int mode, float x, float sc, float sh) {
  if (mode == PDR_PRO_GN_RELU) return fmaxf(fmaf(x, sc, sh), 0.f);
  if (mode == PDR_PRO_RELU_GN) return fmaf(fmaxf(x, 0.f), sc, sh);
  return x;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 2)
gemm_tf32_kernel(const PdrGemmArgs a) {
  constexpr int kBTileBytes = BN * 128;
  constexpr int kStageBytes = kATileBytes + kBTileBytes;
  constexpr int kWLoads = BN * 8 / kTcThreads;          // float4 per thread per chunk (>= 1 for BN >= 32)
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_empty[STAGES];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t s_tmem_base;

  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int tiles_per_sample = (a.rows_per_sample + kTcTileM - 1) / kTcTileM;
  const int tile = blockIdx.y;
  const int b = tile / tiles_per_sample;
  const int r0 = (tile % tiles_per_sample) * kTcTileM;
  const size_t row_base = (size_t)b * a.rows_per_sample + r0;
  const int rows_valid = min(kTcTileM, a.rows_per_sample - r0);
  const int n0 = blockIdx.x * BN;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bar_empty[s], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcTileM >> 4) << 24);

  // ---- loader mapping: A chunk = 128 rows x 8 float4; thread handles rows (tid>>3) + 32*i, float4 #(tid&7) ----
  const int chunk = tid & 7;
  const int arow = tid >> 3;
  float4 ra[4];
  float4 rw[kWLoads];

  auto load_chunk = [&](int k0) {
    const int k = k0 + chunk * 4;
    const bool kin = k < a.K;
    float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f), e4 = h4;
    if (kin) {
      if (a.pro_mode != PDR_PRO_NONE) {
        s4 = __ldg(reinterpret_cast<const float4 *>(a.sc + (size_t)b * a.ld_scsh + k));
        h4 = __ldg(reinterpret_cast<const float4 *>(a.sh + (size_t)b * a.ld_scsh + k));
      }
      if (a.add) e4 = __ldg(reinterpret_cast<const float4 *>(a.add + (size_t)b * a.ld_add + k));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = arow + 32 * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kin && row < rows_valid) {
        const size_t grow = row_base + row;
        v = *reinterpret_cast<const float4 *>(a.A + grow * a.lda + k);
        v.x = pro1(a.pro_mode, v.x, s4.x, h4.x) + e4.x; v.y = pro1(a.pro_mode, v.y, s4.y, h4.y) + e4.y;
        v.z = pro1(a.pro_mode, v.z, s4.z, h4.z) + e4.z; v.w = pro1(a.pro_mode, v.w, s4.w, h4.w) + e4.w;
        if (a.R) {
          const float4 r = *reinterpret_cast<const float4 *>(a.R + grow * a.ldr + k);
          v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < kWLoads; ++i) {
      const int n = arow + 32 * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kin && n0 + n < a.N) {
        v = __ldg(reinterpret_cast<const float4 *>(a.W + (size_t)(n0 + n) * a.ldw + k));
        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
      }
      rw[i] = v;
    }
  };
  auto store_chunk = [&](int s) {
    uint8_t *sa = smem + (size_t)s * kStageBytes;
    uint8_t *sb = sa + kATileBytes;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = arow + 32 * i;
      *reinterpret_cast<float4 *>(sa + row * 128 + ((chunk ^ (row & 7)) << 4)) = ra[i];
    }
#pragma unroll
    for (int i = 0; i < kWLoads; ++i) {
      const int n = arow + 32 * i;
      *reinterpret_cast<float4 *>(sb + n * 128 + ((chunk ^ (n & 7)) << 4)) = rw[i];
    }
  };

  const int nk = (a.K + kTcBK - 1) / kTcBK;
  load_chunk(0);
  for (int kc = 0; kc < nk; ++kc) {
    const int s = kc % STAGES;
    if (kc >= STAGES) mbar_wait(&bar_empty[s], (uint32_t)((kc / STAGES - 1) & 1));
    store_chunk(s);
    if (kc + 1 < nk) load_chunk((kc + 1) * kTcBK);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes);
      const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + kATileBytes);
#pragma unroll
      for (int k8 = 0; k8 < kTcBK / 8; ++k8)
        umma_tf32(tmem_base, adesc + (uint64_t)(k8 * 2), bdesc + (uint64_t)(k8 * 2), kIdesc, (kc | k8) ? 1u : 0u);
      umma_commit(&bar_empty[s]);
      if (kc == nk - 1) umma_commit(&bar_done);
    }
  }
  mbar_wait(&bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue ---------------------------------------------------------------------------------------
  // stage memory is free now: per-warp 32x33 transpose tiles, then per-quarter column partials
  float *s_t = reinterpret_cast<float *>(smem) + warp * (32 * 33);
  float *s_part = reinterpret_cast<float *>(smem) + 8 * 32 * 33;     // [4 quarters][BN][4]
  const int quarter = warp & 3;
  constexpr int kColsPerWarp = BN >= 64 ? BN / 2 : BN;
  const bool warp_active = BN >= 64 || warp < 4;
  const int cbeg = BN >= 64 ? (warp >> 2) * kColsPerWarp : 0;
  const int row = quarter * 32 + lane;
  const bool rvalid = row < rows_valid;
  const size_t grow = row_base + row;
  const bool vec_ok = (a.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0);
  if (warp_active) {
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + kColsPerWarp; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const float *radd = (a.rowadd && rvalid) ? a.rowadd + (grow / a.rowadd_div) * a.ld_rowadd : nullptr;
      float y[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + c0 + j;
        float t = 0.f;
        if (rvalid && n < a.N) {
          t = __uint_as_float(v[j]);
          if (a.bias) t += __ldg(a.bias + n);
          if (radd) t += __ldg(radd + n);
        }
        y[j] = t;
      }
      if (rvalid) {
        float *crow = a.C + grow * a.ldc + n0 + c0;
        if (vec_ok && n0 + c0 + 32 <= a.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(crow + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + c0 + j;
            if (n < a.N) crow[j] = y[j];
            else if (n < a.ldc_zero_to) crow[j] = 0.f;
          }
        }
      }
      if (a.stats) {
#pragma unroll
        for (int j = 0; j < 32; ++j) s_t[lane * 33 + j] = y[j];
        __syncwarp();
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const float t = s_t[r * 33 + lane];
          const float p = fmaxf(t, 0.f);
          q0 += t; q1 = fmaf(t, t, q1); q2 += p; q3 = fmaf(p, p, q3);
        }
        __syncwarp();
        float *dst = s_part + ((size_t)quarter * BN + c0 + lane) * 4;
        dst[0] = q0; dst[1] = q1; dst[2] = q2; dst[3] = q3;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (a.stats) {
    for (int f = tid; f < BN * 4; f += kTcThreads) {
      const int col = f >> 2, q = f & 3;
      if (n0 + col < a.N) {
        const float s = s_part[((size_t)0 * BN + col) * 4 + q] + s_part[((size_t)1 * BN + col) * 4 + q] +
                        s_part[((size_t)2 * BN + col) * 4 + q] + s_part[((size_t)3 * BN + col) * 4 + q];
        a.stats[((size_t)tile * a.N + n0 + col) * 4 + q] = s;
      }
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

template <int BN, int STAGES>
int launch_tc(const PdrGemmArgs &a, cudaStream_t stream) {
  constexpr size_t kSmem = (size_t)STAGES * (kATileBytes + BN * 128) + 1024;
  static_assert(kSmem >= (size_t)(8 * 32 * 33 + 4 * BN * 4) * 4 + 1024, "epilogue scratch must fit in the stages");
  auto kern = gemm_tf32_kernel<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) { set_error("gemm_tf32: smem attr: %s", cudaGetErrorString(e)); return PDR_ERR_CUDA; }
    configured = true;
  }
  const int tiles_per_sample = ceil_div(a.rows_per_sample, kTcTileM);
  const long long tiles = (long long)a.batch * tiles_per_sample;
  if (tiles > 65535) { set_error("gemm_tf32: %lld tiles > 65535", tiles); return PDR_ERR_INVALID_ARGUMENT; }
  dim3 grid(ceil_div(a.N, BN), (unsigned)tiles);
  kern<<<grid, kTcThreads, kSmem, stream>>>(a);
  return check_launch("gemm_tf32_kernel");
}

}  // namespace

int launch_gemm_tf32(const PdrGemmArgs &a, cudaStream_t stream) {
  if (a.N > 128) return launch_tc<256, 2>(a, stream);
  if (a.N > 64) return launch_tc<128, 3>(a, stream);
  if (a.N > 32) return launch_tc<64, 4>(a, stream);
  return launch_tc<32, 4>(a, stream);
}

}  // namespace pdr
