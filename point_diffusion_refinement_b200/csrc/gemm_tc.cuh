// tcgen05 (5th-gen tensor core) TF32 GEMM with fused prologue/epilogue for the denoiser -- sm_100a only.
//
//   C[M, N] = pro(A)[M, K] . W[N, K]^T + bias (+ rowadd)          fp32 in HBM, TF32 MMA, fp32 accumulate
//
// Same contract as the SIMT kernel in net.cu (PdrGemmArgs).  These GEMMs are tall and skinny (M = B*npoint*
// nsample up to 2 M rows, K and N tens to hundreds): they are HBM-bound, the tensor core is there so that
// the math never is.  The first version (one CTA per 128-row tile, load -> MMA -> epilogue in sequence) ran
// at 9-12 % of HBM peak: every CTA paid a DRAM round trip, a TMEM allocation and an epilogue with nothing
// else in flight (profiles/r01_ncu_gemm_tcgen05_v1_summary.csv).  This version is PERSISTENT and
// WARP-SPECIALISED, one CTA of 608 threads per SM looping over (row tile, column tile) work items:
//
//   warps 8..15  producers  direct mode (no prologue): cp.async of A (dense or gathered rows) and streamed W straight
//                           into the MMA stage in the canonical K-major SWIZZLE_128B layout (16-byte chunk c of row r
//                           at c ^ (r & 7)), completion by cp.async.mbarrier.arrive on full[s];
//                           transform mode: GroupNorm scale/shift, ReLU, per-sample embedding, residual -> TF32, in
//                           place in the stage the loaders filled -> fence.proxy.async -> full[s].
//   warps 16,17  loaders    transform mode only: raw A (+ residual, + the raw gathered K tail) and streamed W by cp.async.
//   warp 18      MMA        one thread: tcgen05.mma.cta_group::1.kind::tf32 (128 x BN x 8), accumulators in
//                           TMEM (double buffered: 2 x BN columns); tcgen05.commit -> empty[s] / tmem_full[a].
//   warps 0..7   epilogue   two groups of four (one warp per TMEM lane quarter), one accumulator / every second tile
//                           each: tcgen05.ld 32 lanes x 32 columns -> + bias / broadcast row-add -> 32 x 32 staging tile
//                           -> TMA tensor store, and the per-column (sum, sum^2, relu-sum, relu-sum^2) statistics the
//                           next GroupNorm needs, read back from the staging tile -> tmem_empty[a].  (Older flavours:
//                           transpose + coalesced STG, float4, pooling.)
//   When the whole weight matrix fits in shared memory next to the ring it is staged once per CTA (W-resident mode)
//   and the ring carries A only.
//
// Round 2 (ncu role analysis in profiles/r02_gather_producer_notes.txt: these kernels are bound by instruction ISSUE, not by HBM):
//   * the warp index comes through a shuffle from lane 0, so ptxas knows it is warp-uniform and keeps role / tile / barrier state in
//     uniform registers (no ELECT + R2UR loop around every UTMASTG / UTCHMMA / SYNCS): -7 % step;
//   * launched as a programmatic dependent (griddepcontrol.wait after the barrier / TMEM set-up; the MMA warp releases the next
//     kernel after its last MMA; resident weights staged before the wait when PdrGemmArgs.w_static);
//   * column tiles of the widest template are balanced (TcPlan.bn), short GEMMs take narrower tiles (gemm_tc.cu);
//   * the gathered operand has its own lean producer loop (and an optional TMA tile::gather4 path), the GroupNorm -> ReLU
//     prologue and the epilogue statistics / bias run on packed fp32 pairs.
//
// This header holds the kernel template and its launcher; gemm_tc_bn{32,64,128,256}.cu instantiate one column-tile
// width each (so that nvcc compiles them in parallel), gemm_tc.cu picks the width.
#pragma once
#include <cuda.h>      // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

namespace pdr {
namespace {

constexpr int kEpiWarps = 8, kProdWarps = 8, kLoadWarps = 2;
// warps [0,8): epilogue, two per TMEM lane quarter, alternating 32-column blocks
// warps [8,16): producers -- direct mode: cp.async into the MMA ring; transform mode: raw ring -> prologue -> TF32
// warps [16,18): loaders (transform mode only): raw A/R into the raw ring, streamed W into the MMA ring
// warp 18: MMA issue
constexpr int kEpiThreads = kEpiWarps * 32, kProdThreads = kProdWarps * 32, kLoadThreads = kLoadWarps * 32;
constexpr int kTcThreads = kEpiThreads + kProdThreads + kLoadThreads + 32;   // 608
constexpr int kMmaWarp = kEpiWarps + kProdWarps + kLoadWarps;   // 18
constexpr int kTcTileM = 128;
constexpr int kTcBK = 32;                       // floats per K chunk = one 128-byte swizzle row
constexpr int kATileBytes = kTcTileM * 128;     // 16 KiB
constexpr int kWResidentBytes = 160 * 1024;   // upper bound; the planner checks what actually fits
constexpr int kMaxStages = 10;

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
// same, for waits that are expected to be long (a producer that is ahead of the ring): sleep between polls so that the
// retry loop does not take issue slots from the warps that are the bottleneck
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity, uint32_t ns) {
  if (ns == 0) { mbar_wait(bar, parity); return; }
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    asm volatile("nanosleep.u32 %0;" ::"r"(ns));
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
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
// Round-to-nearest (ties away) to TF32 for a tensor core that TRUNCATES fp32 operands to 10 mantissa bits: adding
// half a TF32 ulp to the bit pattern is all that is needed, the low 13 bits are ignored by the MMA.  (ptxas expands
// cvt.rna.tf32.f32 into a 4-instruction Inf/NaN-preserving sequence on sm_100a, which made the transform warps the
// bottleneck; Inf/NaN inputs poison the output either way.)
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ float pro1(int mode, float x, float sc, float sh) {
  if (mode == PDR_PRO_GN_RELU) return fmaxf(fmaf(x, sc, sh), 0.f);
  if (mode == PDR_PRO_RELU_GN) return fmaf(fmaxf(x, 0.f), sc, sh);
  return x;
}
__device__ __forceinline__ float4 tf32x4(float4 v) {
  return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool pred) {
  // src-size 0 zero-fills the 16 destination bytes (rows beyond the tile, K tail)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(pred ? 16 : 0) : "memory");
}
// the mbarrier receives one arrival from this thread once ALL its earlier cp.async copies have landed; the
// thread itself does not wait (and, unlike wait_group + fence, is not stalled behind younger copies)
__device__ __forceinline__ void cp_async16_sz(uint32_t dst, const void *src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
// ignore-src form: one LDGSTS.ZFILL with a predicate, no src-size arithmetic; with `ignore` the source is not read
__device__ __forceinline__ void cp_async16_ignore(uint32_t dst, const void *src, bool ignore) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "cp.async.cg.shared.global [%0], [%1], 16, p;\n\t"
      "}\n" ::"r"(dst), "l"(src), "r"((int)ignore)
      : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA tile::gather4: four rows (one 128-byte K chunk each) of a 2-D row-major table, picked by four row indices, land as
// four consecutive 128-byte rows of the SWIZZLE_128B stage (the swizzle is a function of the shared-memory address, so rows
// 4l .. 4l+3 of the canonical K-major tile are exactly what the copy writes at tile + 512 l).  A row index outside the
// table ([0, rows)) reads as zeros and still counts its bytes on the mbarrier.
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *tm, int col, int4 rows, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(tm), "r"(col), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kDirectDepth = 3;   // chunks of cp.async in flight per producer thread, direct mode (needs >= 3 stages)

struct TcPlan {
  int n_tiles_n;        // column tiles (N / BN rounded up)
  int bn;               // columns per column tile: BN, or for the widest template the balanced width ceil(N / n_tiles_n)
                        // rounded up to 32 (N = 300: 160 + 140 instead of 256 + 44, whose light tiles always fell to the
                        // same CTAs of an even-sized persistent grid)
  int tiles_per_sample; // row tiles per sample
  int total_items;      // n_tiles_n * batch * tiles_per_sample, column tile fastest
  int nk;               // K chunks
  int stages;           // MMA-layout ring depth
  int direct;           // 1: no prologue -> cp.async lands straight in the MMA ring; 0: raw ring + transform
  int stage_bytes;      // bytes of one ring stage: A tile [+ W tile when streamed] [+ residual tile]
  int r_off;            // offset of the residual tile inside a stage (transform mode with R)
  int epi_alt;          // 1: the two groups of 4 epilogue warps take alternate TILES (one accumulator each);
                        // 0: both groups share every tile and alternate its column blocks
  int prod_sleep_ns;    // > 0: producers / loaders sleep this long between polls of an EMPTY-stage barrier
  int w_early;          // 1: the resident weight matrix is staged BEFORE the wait on the previous kernel (PdrGemmArgs.w_static)
  int tma_gather;       // 1: whole 32-column chunks of the gathered table come by TMA tile::gather4 (one warp, one
                        //    instruction per 4 rows) instead of 8 cp.async pieces per row with per-thread address arithmetic
  int lean;             // 1: packed GroupNorm -> ReLU prologue (PDR_GEMM_LEAN=0: the generic clamp-fma-clamp, A/B)
  int pdl_late;         // 1: the MMA warp releases the dependent launch after its last MMA, 0: every thread at kernel start
};

// bytes of the TMA-store staging tiles (EPI 3): two 32 x 32 fp32 boxes per epilogue warp
// two staging tiles per epilogue warp; ONE for the 256-column tiles, whose 48 KB ring stages (A + streamed W) otherwise get
// a 2-deep ring next to 96 KB of epilogue buffers (46 launches, 2.6 ms / step at ~1.5 TB/s: latency-bound, r02k)
constexpr int tma_bufs(int bn) { return bn == 256 ? 1 : 2; }
constexpr int tma_stage_bytes(int bn) { return kEpiWarps * tma_bufs(bn) * 4096; }

// EPI: 0 = scalar epilogue, 1 = float4 epilogue, 2 = attention pooling over the K neighbour rows (no C store),
//      3 = TMA-store epilogue (32 x 32 boxes through a 128B-swizzled staging tile, statistics read back column-wise)
//
// The raw gathered K tail (PdrGemmArgs.tail_rows) is copied by the 8 transform warps, which have nothing to transform there
// (4 rows per thread), not by the 2 loader warps (16 rows per thread, saturated on the folded-residual GEMMs): measured
// -0.07 ms / step, profiles/r02_experiments_ab.txt.  (The two other experiments of round 1 -- a shared-memory ring for the
// gathered-A indices, GroupNorm finalisation inside the epilogue -- measured slower and were removed, same file.)
template <int BN, bool WRES, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32_persistent(const PdrGemmArgs a, const TcPlan plan, const __grid_constant__ CUtensorMap tmap_c,
                     const __grid_constant__ CUtensorMap tmap_a) {
  constexpr int kBTileBytes = BN * 128;
  constexpr int kWLoads = BN * 8 / kProdThreads;        // float4 of W per producer thread per chunk
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_tfull[2], bar_tempty[2], bar_wready;
  __shared__ uint64_t bar_rfull[kMaxStages];   // transform mode: raw A (+R) of the stage has landed
  __shared__ uint32_t s_tmem_base;
  // epilogue scratch is STATIC shared memory so that the compiler emits LDS/STS (the first persistent version
  // indexed it through a generic pointer into the dynamic region: LD.E/ST.E at ~3x the latency, which made the
  // 4 epilogue warps the bottleneck of the whole pipeline -- profiles/r01_ncu_gemm_tcgen05_v2_hotspots.txt)
  __shared__ __align__(16) float s_epi[kEpiWarps][EPI == 3 ? 192 : 32 * 36];   // EPI 3: per-column addends, <= 5 row groups
  // the column partials (4 lane quarters x BN x 4 sums) live in the dynamic region and are touched only through
  // explicit ld/st.shared (a handful of accesses per block), which keeps static shared memory under 48 KB

  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // layout: [W resident (WRES)] [MMA stages] [raw ring (transform mode)] [column partials 4 x BN x float4]
  uint8_t *s_wres = smem;
  uint8_t *s_stages = smem + (WRES ? (size_t)plan.nk * kBTileBytes : 0);
  const uint32_t kStageBytes = (uint32_t)plan.stage_bytes;
  const uint32_t s_part = smem_u32(s_stages + (size_t)plan.stages * kStageBytes);   // [4][BN] float4
  // EPI 3: staging tiles of the TMA stores, after the column partials (every region before is a multiple of 1 KiB)
  constexpr uint32_t kPartRegion = (uint32_t)(BN <= 128 ? 4 : 2) * 4u * BN * 16u;
  const uint32_t s_tma = s_part + kPartRegion;

  // (warp index through a shuffle from lane 0: nvcc then knows it is warp-uniform, keeps what derives from it in uniform registers
  //  and issues UTMASTG / UTCHMMA / SYNCS operands without a per-lane R2UR loop)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int S = plan.stages, nk = plan.nk;
  if (!plan.pdl_late) pdl_launch_dependents();      // PDL (common.cuh): the next kernel's grid may be launched; it waits for this one

  if (tid == 0) {
    const int full_count = plan.direct ? kProdThreads : kProdThreads + kLoadThreads;
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], full_count); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_tfull[0], 1); mbar_init(&bar_tfull[1], 1);
    const uint32_t tempty_count = plan.epi_alt ? kEpiThreads / 2 : kEpiThreads;
    mbar_init(&bar_tempty[0], tempty_count); mbar_init(&bar_tempty[1], tempty_count);
    mbar_init(&bar_wready, kProdThreads);
    for (int r = 0; r < S; ++r) mbar_init(&bar_rfull[r], kLoadThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"((uint32_t)(2 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, s_tmem_base, 0);   // (uniform for ptxas, like the warp index)
  // barrier init and the TMEM allocation above overlap the tail of the previous kernel; from here on global memory is touched
  const bool stages_w_first = WRES && plan.w_early && warp >= kEpiWarps && warp < kEpiWarps + kProdWarps;
  if (!stages_w_first) pdl_wait();

  // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
  const int bn = BN == 256 ? plan.bn : BN;       // runtime tile width for the widest template only
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kTcTileM >> 4) << 24);

  if (warp >= kEpiWarps && warp < kMmaWarp) {
    // =============================== PRODUCERS ===============================================
    const int ptid = tid - kEpiThreads;
    const int chunk = ptid & 7;       // which 16-byte piece of the 128-byte K chunk
    const int arow = ptid >> 3;       // rows arow + 32*i
    const bool is_loader = warp >= kEpiWarps + kProdWarps;
    if (WRES && !is_loader) {          // stage the whole weight matrix once
      for (int kc = 0; kc < nk; ++kc) {
        const int k = kc * kTcBK + chunk * 4;
#pragma unroll
        for (int i = 0; i < kWLoads; ++i) {
          const int n = arow + 32 * i;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < a.K && n < a.N) v = tf32x4(__ldg(reinterpret_cast<const float4 *>(a.W + (size_t)n * a.ldw + k)));
          *reinterpret_cast<float4 *>(s_wres + (size_t)kc * kBTileBytes + n * 128 + ((chunk ^ (n & 7)) << 4)) = v;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&bar_wready);
      if (plan.w_early) pdl_wait();
    }
    // chunk sequence of this CTA: items blockIdx.x, +gridDim.x, ... ; nk chunks each.
    // The producer loop runs once per 16 KiB chunk in every producer thread, so it is kept lean: item geometry
    // is cached per cursor and advanced incrementally (no integer division in the common case), global and
    // shared addresses are running pointers / per-thread constants.
    const int my_items = plan.total_items > (int)blockIdx.x
                             ? (plan.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int my_chunks = my_items * nk;
    const int G = (int)gridDim.x;
    const bool fast_adv = plan.n_tiles_n == 1 && plan.tiles_per_sample >= G;
    struct Cur {                 // one (item, K-chunk) position in this CTA's sequence + cached geometry
      int item, kc, b, tis, n0, rows_valid;
      const float *pa;           // &A[row_base + arow][chunk*4]
      const float *pr;           // same for the residual
      const float *pg[4];        // gathered A: &A[a_rows[row_base + arow + 32 i]][chunk*4], nullptr = zero row
      const float *p2;           // gathered A: &A2[row_base + arow][chunk*4]
    };
    auto locate = [&](Cur &c) {  // full (division) geometry of c.item
      const int tile = c.item / plan.n_tiles_n;
      c.n0 = (c.item - tile * plan.n_tiles_n) * bn;
      c.b = tile / plan.tiles_per_sample;
      c.tis = tile - c.b * plan.tiles_per_sample;
    };
    // gathered A (PdrGemmArgs.a_rows): the neighbour rows of the CURRENT item sit in cidx, those of the NEXT item of
    // this CTA are already in flight in nidx, so the index loads never stall the copy loop
    const bool gath = a.a_rows != nullptr;
    // (kIdxAhead items ahead; 3 measured neutral-to-slower than 1 on B200, gpurun call r01s3b: the registers of all
    // look-ahead loads share one scoreboard, so the consumer of the oldest waits for the youngest)
    constexpr int kIdxAhead = 1;
    int cidx[4] = {-1, -1, -1, -1}, nidx[kIdxAhead][4];
#pragma unroll
    for (int d = 0; d < kIdxAhead; ++d)
#pragma unroll
      for (int i = 0; i < 4; ++i) nidx[d][i] = -1;
    // the look-ahead position (the item whose neighbour rows are requested next) is stepped, not divided: the index
    // prefetch runs once per item in every producer warp, and two integer divisions were a fifth of the instructions a
    // producer warp spends on an item (ncu, profiles/r02_gather_producer_notes.txt)
    struct Pos { int item, b, tis, ct; };
    const int step_q = G / plan.n_tiles_n, step_r = G - step_q * plan.n_tiles_n;
    const bool step_divides = step_q + 1 >= plan.tiles_per_sample;      // a step may cross more than one sample
    auto locate_pos = [&](Pos &q) {
      const int tile = q.item / plan.n_tiles_n;
      q.ct = q.item - tile * plan.n_tiles_n;
      q.b = tile / plan.tiles_per_sample;
      q.tis = tile - q.b * plan.tiles_per_sample;
    };
    auto step_pos = [&](Pos &q) {
      q.item += G;
      if (step_divides) { locate_pos(q); return; }
      q.ct += step_r;
      int dt = step_q;
      if (q.ct >= plan.n_tiles_n) { q.ct -= plan.n_tiles_n; ++dt; }
      q.tis += dt;
      if (q.tis >= plan.tiles_per_sample) { q.tis -= plan.tiles_per_sample; ++q.b; }
    };
    auto fetch_idx = [&](const Pos &q, int (&out)[4]) {
      if (q.item >= plan.total_items) return;
      const int b = q.b, r0 = q.tis * kTcTileM;
      const int *p = a.a_rows + (size_t)b * a.rows_per_sample + r0 + arow;
#pragma unroll
      for (int i = 0; i < 4; ++i) out[i] = (r0 + arow + 32 * i < a.rows_per_sample) ? __ldg(p + 32 * i) : -1;
    };
    // TMA gather (TcPlan.tma_gather): lanes 0..3 of the 8 producer warps each own 4 rows (rows 4 g4 .. 4 g4 + 3, g4 < 32) and
    // hold their neighbour rows for the current and for the next item -- UTMALDG takes its operands from uniform registers,
    // so ptxas serialises the lanes of a warp; four per warp keeps that loop short and the eight warps issue side by side.
    // The per-thread indices above are needed only when the table part ends inside a chunk.
    const bool tg = gath && plan.tma_gather;
    const bool tg_warp = tg && !is_loader && lane < 4;
    const int g4 = (warp - kEpiWarps) * 4 + lane;
    const bool need_cidx = gath && !is_loader && !(tg && a.k_split % kTcBK == 0);
    const bool g4_vec = a.rows_per_sample % 4 == 0 && (reinterpret_cast<uintptr_t>(a.a_rows) & 15) == 0;
    int4 g4c = make_int4(-1, -1, -1, -1), g4n = make_int4(-1, -1, -1, -1);
    Pos ahead = {0, 0, 0, 0};
    auto fetch_g4 = [&](const Pos &q, int4 &out) {
      out = make_int4(-1, -1, -1, -1);
      if (q.item >= plan.total_items) return;
      const int b = q.b, r = q.tis * kTcTileM + 4 * g4;
      const int *p = a.a_rows + (size_t)b * a.rows_per_sample + r;
      if (g4_vec) {
        if (r < a.rows_per_sample) out = __ldg(reinterpret_cast<const int4 *>(p));
      } else {
        if (r < a.rows_per_sample) out.x = __ldg(p);
        if (r + 1 < a.rows_per_sample) out.y = __ldg(p + 1);
        if (r + 2 < a.rows_per_sample) out.z = __ldg(p + 2);
        if (r + 3 < a.rows_per_sample) out.w = __ldg(p + 3);
      }
    };
    auto derive = [&](Cur &c) {  // pointers and row count from (b, tis)
      const int r0 = c.tis * kTcTileM;
      c.rows_valid = min(kTcTileM, a.rows_per_sample - r0);
      const size_t row = (size_t)c.b * a.rows_per_sample + r0 + arow;
      if (gath) {                // (the dense pointers are not used by the gathered producer)
        c.pa = a.A; c.pr = nullptr;
        if (need_cidx) {
#pragma unroll
          for (int i = 0; i < 4; ++i) c.pg[i] = cidx[i] >= 0 ? a.A + (size_t)cidx[i] * a.lda + chunk * 4 : nullptr;
        }
        c.p2 = a.A2 + row * a.lda2 + chunk * 4;
      } else {
        c.pa = a.A + row * a.lda + chunk * 4;
        c.pr = a.R ? a.R + row * a.ldr + chunk * 4 : nullptr;
      }
    };
    auto advance = [&](Cur &c) {
      if (++c.kc < nk) return;
      c.kc = 0;
      c.item += G;
      if (fast_adv) { c.tis += G; if (c.tis >= plan.tiles_per_sample) { c.tis -= plan.tiles_per_sample; ++c.b; } }
      else locate(c);
      if (gath && !is_loader) {
        step_pos(ahead);
        if (tg_warp) { g4c = g4n; fetch_g4(ahead, g4n); }
        if (need_cidx) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cidx[i] = nidx[0][i];
#pragma unroll
          for (int d = 0; d + 1 < kIdxAhead; ++d)
#pragma unroll
            for (int i = 0; i < 4; ++i) nidx[d][i] = nidx[d + 1][i];
          fetch_idx(ahead, nidx[kIdxAhead - 1]);
        }
      }
      derive(c);
    };
    const uint32_t sw_off = (uint32_t)(arow * 128 + ((chunk ^ (arow & 7)) << 4));   // row arow+32i: + i*4096
    const uint32_t raw_off = (uint32_t)(arow * 128 + (chunk << 4));
    const size_t a_step = (size_t)32 * a.lda, r_step = (size_t)32 * a.ldr;
    const size_t w_step = (size_t)32 * a.ldw;
    Cur ci;
    ci.item = (int)blockIdx.x; ci.kc = 0;
    static_assert(kIdxAhead == 1, "one look-ahead position");
    // the gathered operand without TMA has its own, lean, loop below
    const bool lean_gather = plan.direct && gath && !tg;
    if (gath && !is_loader && !lean_gather) {
      ahead.item = ci.item; locate_pos(ahead);
      if (tg_warp) fetch_g4(ahead, g4c);
      if (need_cidx) fetch_idx(ahead, cidx);
      step_pos(ahead);
      if (tg_warp) fetch_g4(ahead, g4n);
      if (need_cidx) fetch_idx(ahead, nidx[0]);
    }
    locate(ci); derive(ci);

    if (lean_gather) {
      // ---- gathered A (feature-table rows picked by a_rows | dense geometric channels), no prologue.  ncu showed these
      //      producers busy 95 % of the kernel at 170 - 200 warp-instructions per 16 KiB chunk in the general loop below
      //      (kernel parameters re-read from the constant bank, cursor divisions, pointers derived per chunk, the W tile
      //      behind two levels of predicates): profiles/r02_gather_producer_notes.txt.  Here everything that depends on the
      //      item only is computed once per item -- 32-bit element offsets of my four table rows, validity masks, the W
      //      row count -- and a chunk costs one 64-bit add per operand plus (IMAD.WIDE, LDGSTS) per copy. ----
      if (!is_loader && my_items > 0) {
        const int lda16 = a.lda >> 2, ksplit = a.k_split, Kk = a.K, rps = a.rows_per_sample;     // lda % 4 == 0
        const size_t st2 = (size_t)32 * a.lda2 * sizeof(float);
        const size_t wst = (size_t)32 * a.ldw * sizeof(float);
        const char *const Abase = reinterpret_cast<const char *>(a.A + chunk * 4);
        const int t_last = Kk - 4;                                // last 16-byte piece inside K (K % 4 == 0)
        Pos cur;
        cur.item = (int)blockIdx.x; locate_pos(cur);
        ahead = cur; step_pos(ahead);
        int ci4[4] = {-1, -1, -1, -1}, ni4[4] = {-1, -1, -1, -1};
        fetch_idx(cur, ci4);
        fetch_idx(ahead, ni4);
        int stage = 0, phase = 0;
        uint32_t sa = smem_u32(s_stages) + sw_off;
        for (int it = 0; it < my_items; ++it) {
          const int r0 = cur.tis * kTcTileM;
          const int rows_valid = min(kTcTileM, rps - r0);
          // geometric channels of my row arow (+ 32 i): column t - ksplit of piece t = kofs + chunk * 4
          const char *const g0 = reinterpret_cast<const char *>(a.A2 + ((size_t)cur.b * rps + r0 + arow) * a.lda2);
          int toff[4], tsz[4], gsz[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            toff[i] = ci4[i] < 0 ? 0 : ci4[i] * lda16;            // in 16-byte units: tables up to 32 GiB
            tsz[i] = ci4[i] < 0 ? 0 : 16;
            gsz[i] = arow + 32 * i < rows_valid ? 16 : 0;
          }
          const int n0 = cur.ct * bn;
          const char *const w0 = reinterpret_cast<const char *>(a.W + (size_t)(n0 + arow) * a.ldw + chunk * 4);
          const int w_rows = min(bn, a.N - n0) - arow;            // W rows n0 + arow + 32 i exist for 32 i < w_rows
          // the next item's rows are requested now, a whole item ahead of their use
#pragma unroll
          for (int i = 0; i < 4; ++i) ci4[i] = ni4[i];
          step_pos(cur);
          step_pos(ahead);
#pragma unroll
          for (int i = 0; i < 4; ++i) ni4[i] = -1;
          fetch_idx(ahead, ni4);
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait_sleep(&bar_empty[stage], (uint32_t)(phase ^ 1), (uint32_t)plan.prod_sleep_ns);
            const int kofs = kc * kTcBK, t = kofs + chunk * 4;
            if (t < ksplit) {
              const char *const ak = Abase + (size_t)kofs * sizeof(float);
#pragma unroll
              for (int i = 0; i < 4; ++i) cp_async16_sz(sa + i * 4096, ak + (long long)toff[i] * 16, tsz[i]);
            } else {
              const bool in = t < Kk;                              // beyond K: zeros (no bytes are read)
              const char *const gk = g0 + (size_t)(min(t, t_last) - ksplit) * sizeof(float);
#pragma unroll
              for (int i = 0; i < 4; ++i) cp_async16_sz(sa + i * 4096, (in && gsz[i]) ? gk + i * st2 : Abase, in ? gsz[i] : 0);
            }
            if (!WRES) {
              const bool in = t < Kk;
              const char *const wk = w0 + (size_t)min(t, t_last) * sizeof(float) - (size_t)chunk * 4 * sizeof(float);
#pragma unroll
              for (int i = 0; i < kWLoads; ++i)
                if (BN < 256 || i * 32 < bn) {
                  const bool ok = in && 32 * i < w_rows;
                  cp_async16_sz(sa + kATileBytes + i * 4096, ok ? wk + i * wst : Abase, ok ? 16 : 0);
                }
            }
            cp_async_arrive_noinc(&bar_full[stage]);
            sa += (uint32_t)kStageBytes;
            if (++stage == S) { stage = 0; phase ^= 1; sa = smem_u32(s_stages) + sw_off; }
          }
        }
      }
    } else if (plan.direct) {
      // ---- no prologue: all 8 producer warps cp.async straight into the swizzled MMA stage.  Completion is
      //      signalled by cp.async.mbarrier.arrive (no wait_group, no fence: a fence.proxy.async compiles to
      //      MEMBAR.ALL.CTA, which also waits for the YOUNGER copies in flight and serialised the ring --
      //      2.5 us per 16 KiB chunk in the previous version).  Depth = number of stages. ----
      auto issue = [&](const Cur &c, int stage) {
        const int kofs = c.kc * kTcBK;
        const bool kin = kofs + chunk * 4 < a.K;
        const uint32_t sa = smem_u32(s_stages + (size_t)stage * kStageBytes) + sw_off;
        const float *src = c.pa + kofs;
        // interior chunks (full rows, full K chunk, full column tile) take a predicate-free path: the copy loops
        // are what the producer warps spend their issue slots on
        // (K tail: this thread's 16-byte piece is either wholly inside K or wholly zero-filled)
        const int ksz = kin ? 16 : 0;
        if (gath) {
          if (tg && kofs + kTcBK <= a.k_split) {       // a whole chunk of table columns: 32 gather4 copies, 4 per warp
            if (tg_warp) {
              mbar_expect_tx(&bar_full[stage], 512u);
              tma_gather4(smem_u32(s_stages + (size_t)stage * kStageBytes) + (uint32_t)g4 * 512u, &tmap_a, kofs, g4c,
                          &bar_full[stage]);
            }
          } else if (kofs + chunk * 4 < a.k_split) {   // feature part: one table row per grouped row
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float *p = c.pg[i];
              cp_async16_sz(sa + i * 4096, p ? p + kofs : a.A, p ? 16 : 0);
            }
          } else {                                       // geometric channels, dense (M, lda2)
            const float *p2 = c.p2 + (kofs - a.k_split);
            const size_t st2 = (size_t)32 * a.lda2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const bool ok = kin && arow + 32 * i < c.rows_valid;
              cp_async16_sz(sa + i * 4096, ok ? p2 + i * st2 : a.A, ok ? 16 : 0);
            }
          }
        } else if (c.rows_valid == kTcTileM) {
          const float *p = kin ? src : a.A;
          const size_t st = kin ? a_step : 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) { cp_async16_sz(sa + i * 4096, p, ksz); p += st; }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = kin && arow + 32 * i < c.rows_valid;
            cp_async16(sa + i * 4096, ok ? src + i * a_step : a.A, ok);
          }
        }
        if (!WRES) {
          const float *wsrc = a.W + (size_t)(c.n0 + arow) * a.ldw + kofs + chunk * 4;
          if (c.n0 + bn <= a.N) {
            const float *p = kin ? wsrc : a.W;
            const size_t st = kin ? w_step : 0;
#pragma unroll
            for (int i = 0; i < kWLoads; ++i)
              if (BN < 256 || i * 32 < bn) { cp_async16_sz(sa + kATileBytes + i * 4096, p, ksz); p += st; }
          } else {
#pragma unroll
            for (int i = 0; i < kWLoads; ++i) {
              const bool ok = kin && c.n0 + arow + 32 * i < a.N;
              if (BN < 256 || i * 32 < bn) cp_async16(sa + kATileBytes + i * 4096, ok ? wsrc + i * w_step : a.W, ok);
            }
          }
        }
      };
      int stage = 0, phase = 0;
      for (int j = 0; j < (is_loader ? 0 : my_chunks); ++j) {
        mbar_wait_sleep(&bar_empty[stage], (uint32_t)(phase ^ 1), (uint32_t)plan.prod_sleep_ns);
        issue(ci, stage);
        cp_async_arrive_noinc(&bar_full[stage]);
        advance(ci);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    } else if (is_loader) {
      // ---- transform mode, LOADERS (warps 16,17): raw A (+ residual) into the raw ring, weights (when streamed)
      //      straight into the MMA stage; completion is signalled through cp.async.mbarrier.arrive, so these two
      //      warps never block on their own copies ----
      const int ltid = tid - kEpiThreads - kProdThreads;   // 0..63
      const int lchunk = ltid & 7, lrow = ltid >> 3;        // rows lrow + 8*i, i < 16; (row & 7) == lrow
      const uint32_t l_sw = (uint32_t)(lrow * 128 + ((lchunk ^ lrow) << 4));
      Cur cl;
      cl.item = (int)blockIdx.x; cl.kc = 0;
      auto derive_l = [&](Cur &c) {
        const int r0 = c.tis * kTcTileM;
        c.rows_valid = min(kTcTileM, a.rows_per_sample - r0);
        const size_t row = (size_t)c.b * a.rows_per_sample + r0 + lrow;
        c.pa = a.A + row * a.lda + lchunk * 4;
        c.pr = a.R ? a.R + row * a.ldr + lchunk * 4 : nullptr;
      };
      locate(cl); derive_l(cl);
      const size_t a8 = (size_t)8 * a.lda, r8 = (size_t)8 * a.ldr, w8 = (size_t)8 * a.ldw;
      // raw gathered K tail (PdrGemmArgs.tail_rows): copied by the transform warps (below), not here
      const bool has_tail = a.tail_rows != nullptr;
      int stage = 0, phase = 0;
      for (int j = 0; j < my_chunks; ++j) {
        const int kofs = cl.kc * kTcBK;
        const bool kin = kofs + lchunk * 4 < a.K;
        mbar_wait_sleep(&bar_empty[stage], (uint32_t)(phase ^ 1), (uint32_t)plan.prod_sleep_ns);   // the MMAs that read this stage have retired
        const uint32_t sa = smem_u32(s_stages + (size_t)stage * kStageBytes) + l_sw;
        // interior chunks take a predicate-free path (see the direct producer)
        const int ksz = kin ? 16 : 0;
        const bool afull = cl.rows_valid == kTcTileM;
        if (has_tail && kofs >= a.k_pro) {
          // the transform warps copy this chunk (4 rows per thread instead of 16 here: the two loader warps were saturated
          // on the folded-residual GEMMs); the loaders only keep the raw-data barrier's phases in step (below)
        } else if (afull) {
          const float *src = kin ? cl.pa + kofs : a.A;
          const size_t st = kin ? a8 : 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) { cp_async16_sz(sa + i * 1024, src, ksz); src += st; }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool ok = kin && lrow + 8 * i < cl.rows_valid;
            cp_async16(sa + i * 1024, ok ? cl.pa + kofs + i * a8 : a.A, ok);
          }
        }
        if (a.R) {
          if (afull) {
            const float *src = kin ? cl.pr + kofs : a.R;
            const size_t st = kin ? r8 : 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) { cp_async16_sz(sa + plan.r_off + i * 1024, src, ksz); src += st; }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const bool ok = kin && lrow + 8 * i < cl.rows_valid;
              cp_async16(sa + plan.r_off + i * 1024, ok ? cl.pr + kofs + i * r8 : a.R, ok);
            }
          }
        }
        cp_async_arrive_noinc(&bar_rfull[stage]);
        if (!WRES) {
          const float *wsrc = a.W + (size_t)(cl.n0 + lrow) * a.ldw + kofs + lchunk * 4;
          if (cl.n0 + bn <= a.N) {
            const float *src = kin ? wsrc : a.W;
            const size_t st = kin ? w8 : 0;
#pragma unroll
            for (int i = 0; i < BN / 8; ++i)
              if (BN < 256 || i * 8 < bn) { cp_async16_sz(sa + kATileBytes + i * 1024, src, ksz); src += st; }
          } else {
#pragma unroll
            for (int i = 0; i < BN / 8; ++i) {
              const bool ok = kin && cl.n0 + lrow + 8 * i < a.N;
              if (BN < 256 || i * 8 < bn) cp_async16(sa + kATileBytes + i * 1024, ok ? wsrc + i * w8 : a.W, ok);
            }
          }
        }
        cp_async_arrive_noinc(&bar_full[stage]);
        if (++cl.kc == nk) {
          cl.kc = 0; cl.item += G;
          if (fast_adv) { cl.tis += G; if (cl.tis >= plan.tiles_per_sample) { cl.tis -= plan.tiles_per_sample; ++cl.b; } }
          else locate(cl);
          derive_l(cl);
        }
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    } else {
      // ---- transform mode, TRANSFORMERS (warps 8..15): raw ring -> GroupNorm/ReLU/embedding/residual -> TF32 ->
      //      swizzled MMA stage.  This loop bounds every GEMM with a prologue, so it is kept to ~6 instructions
      //      per element: both prologue flavours are one clamp-fma-clamp
      //          y = max(fma(max(x, lo1), sc, sh), lo2) + e (+ r)
      //      with (lo1, lo2) = (-inf, 0) for GN->ReLU, (0, -inf) for ReLU->GN, (-inf, -inf) for none; rows beyond
      //      the tile and the K tail need no test (the loaders zero-filled them, sc/sh/e default to 1/0/0 there,
      //      rows are independent in the MMA and the epilogue never reads the padding rows).  sc/sh/e of the NEXT
      //      chunk are fetched before the current one is processed. ----
      const uint32_t t_sw = sw_off;                                          // rows arow + 32*i, i < 4
      const float ninf = __int_as_float(0xff800000);
      const float lo1 = a.pro_mode == PDR_PRO_RELU_GN ? 0.f : ninf;
      const float lo2 = a.pro_mode == PDR_PRO_GN_RELU ? 0.f : ninf;
      Cur ct;
      ct.item = (int)blockIdx.x; ct.kc = 0;
      locate(ct);
      const int k_pro = a.tail_rows ? a.k_pro : a.K;          // columns that get the prologue
      // per-sample rows of the prologue constants (scale / shift / embedding), re-derived only when the sample changes:
      // the transform warps are bound by instruction issue (ncu, profiles/r02_gather_producer_notes.txt), and three 64-bit
      // multiply-adds per chunk for pointers that change every few hundred chunks were part of it
      const bool has_pro = a.pro_mode != PDR_PRO_NONE, has_add = a.add != nullptr;
      int cb_b = -1;
      const float *psc = nullptr, *psh = nullptr, *pad = nullptr;
      auto fetch = [&](const Cur &c, float4 &s4, float4 &h4, float4 &e4) {
        const int k = c.kc * kTcBK + chunk * 4;
        s4 = make_float4(1.f, 1.f, 1.f, 1.f); h4 = make_float4(0.f, 0.f, 0.f, 0.f); e4 = h4;
        if (c.b != cb_b) {
          cb_b = c.b;
          if (has_pro) { psc = a.sc + (size_t)c.b * a.ld_scsh + chunk * 4; psh = a.sh + (size_t)c.b * a.ld_scsh + chunk * 4; }
          if (has_add) pad = a.add + (size_t)c.b * a.ld_add + chunk * 4;
        }
        if (k < k_pro) {
          const int kk = c.kc * kTcBK;
          if (has_pro) {
            s4 = __ldg(reinterpret_cast<const float4 *>(psc + kk));
            h4 = __ldg(reinterpret_cast<const float4 *>(psh + kk));
          }
          if (has_add) e4 = __ldg(reinterpret_cast<const float4 *>(pad + kk));
        }
      };
      float4 s4, h4, e4;
      if (my_chunks > 0) fetch(ct, s4, h4, e4);
      int stage = 0, phase = 0;
      int tidx4[4] = {-1, -1, -1, -1};                          // raw K tail: table rows of my 4 tile rows (arow + 32 i)
      for (int j = 0; j < my_chunks; ++j) {
        const bool raw_chunk = ct.kc * kTcBK >= k_pro;          // gathered tail: lands ready for the tensor core
        const int kc_cur = ct.kc, b_cur = ct.b, tis_cur = ct.tis;
        if (a.tail_rows && kc_cur == 0) {                       // requested k_pro / 32 chunks before their first use
          const int r0 = tis_cur * kTcTileM;
          const int *p = a.tail_rows + (size_t)b_cur * a.rows_per_sample + r0 + arow;
#pragma unroll
          for (int i = 0; i < 4; ++i) tidx4[i] = (r0 + arow + 32 * i < a.rows_per_sample) ? __ldg(p + 32 * i) : -1;
        }
        if (++ct.kc == nk) {                                    // cursor of chunk j + 1
          ct.kc = 0; ct.item += G;
          if (fast_adv) { ct.tis += G; if (ct.tis >= plan.tiles_per_sample) { ct.tis -= plan.tiles_per_sample; ++ct.b; } }
          else locate(ct);
        }
        float4 ns4 = s4, nh4 = h4, ne4 = e4;
        if (j + 1 < my_chunks) fetch(ct, ns4, nh4, ne4);
        if (raw_chunk) {
          // nothing to transform.  The wait keeps this warp from running ahead of the ring (an arrival for the NEXT use
          // of a stage must not land in the phase of the current one); the data needs no fence, cp.async wrote it.
          mbar_wait(&bar_rfull[stage], (uint32_t)phase);
          {
            // the loaders passed this stage's empty barrier before arriving on rfull, so the stage is free: copy my four
            // 16-byte pieces of the tail (one address form: base + sel * mul, sel < 0 -> zeros) and let the copies arrive
            const int kofs = kc_cur * kTcBK;
            const int t0 = kofs - a.k_pro + chunk * 4;
            const bool is_g = t0 < a.t_split;
            const int r0 = tis_cur * kTcTileM;
            const int rows_valid = min(kTcTileM, a.rows_per_sample - r0);
            const int nval = (kofs + chunk * 4 < a.K) ? (rows_valid - arow + 31) >> 5 : 0;
            const float *base = a.T + t0;
            if (!is_g) {
              const size_t row = (size_t)b_cur * a.rows_per_sample + r0 + arow;
              base = nval > 0 ? a.T2 + row * a.ldt2 + (t0 - a.t_split) : a.T2;
            }
            const int mul = is_g ? a.ldt : 32 * a.ldt2;
            const uint32_t dst = smem_u32(s_stages + (size_t)stage * kStageBytes) + t_sw;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int sel = is_g ? tidx4[i] : (i < nval ? i : -1);
              cp_async16_ignore(dst + i * 4096, base + (long long)max(sel, 0) * mul, sel < 0);
            }
            cp_async_arrive_noinc(&bar_full[stage]);
          }
          s4 = ns4; h4 = nh4; e4 = ne4;
          if (++stage == S) { stage = 0; phase ^= 1; }
          continue;
        }
        mbar_wait(&bar_rfull[stage], (uint32_t)phase);
        // in place: the loaders land raw A (and R) at the swizzled position the tensor core expects, so every
        // thread rewrites exactly the 16-byte pieces it read
        uint8_t *sa = s_stages + (size_t)stage * kStageBytes + t_sw;
        const uint8_t *sr = sa;
        auto xf = [&](float x, float sc, float sh, float e) {
          return fmaxf(fmaf(fmaxf(x, lo1), sc, sh), lo2) + e;
        };
        if (a.pro_mode == PDR_PRO_GN_RELU && !a.R && plan.lean) {
          // the common flavour (GroupNorm -> ReLU [+ embedding]), on packed fp32 pairs: y = max(fma(x, sc, sh), 0) [+ e]; every
          // operation is the .rn form of the scalar one in xf() (whose first clamp is the identity here), bit for bit
          unsigned long long s01, s23, h01, h23, e01, e23;
          asm("mov.b64 %0, {%1, %2};" : "=l"(s01) : "f"(s4.x), "f"(s4.y));
          asm("mov.b64 %0, {%1, %2};" : "=l"(s23) : "f"(s4.z), "f"(s4.w));
          asm("mov.b64 %0, {%1, %2};" : "=l"(h01) : "f"(h4.x), "f"(h4.y));
          asm("mov.b64 %0, {%1, %2};" : "=l"(h23) : "f"(h4.z), "f"(h4.w));
          asm("mov.b64 %0, {%1, %2};" : "=l"(e01) : "f"(e4.x), "f"(e4.y));
          asm("mov.b64 %0, {%1, %2};" : "=l"(e23) : "f"(e4.z), "f"(e4.w));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 v = *reinterpret_cast<const float4 *>(sr + i * 4096);
            unsigned long long v01, v23;
            asm("mov.b64 %0, {%1, %2};" : "=l"(v01) : "f"(v.x), "f"(v.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(v23) : "f"(v.z), "f"(v.w));
            asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v01) : "l"(s01), "l"(h01));
            asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v23) : "l"(s23), "l"(h23));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(v01));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v.z), "=f"(v.w) : "l"(v23));
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            if (has_add) {
              asm("mov.b64 %0, {%1, %2};" : "=l"(v01) : "f"(v.x), "f"(v.y));
              asm("mov.b64 %0, {%1, %2};" : "=l"(v23) : "f"(v.z), "f"(v.w));
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v01) : "l"(e01));
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v23) : "l"(e23));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(v01));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(v.z), "=f"(v.w) : "l"(v23));
            }
            *reinterpret_cast<float4 *>(sa + i * 4096) = tf32x4(v);
          }
        } else if (a.R) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 v = *reinterpret_cast<const float4 *>(sr + i * 4096);
            const float4 r = *reinterpret_cast<const float4 *>(sr + plan.r_off + i * 4096);
            v.x = xf(v.x, s4.x, h4.x, e4.x) + r.x; v.y = xf(v.y, s4.y, h4.y, e4.y) + r.y;
            v.z = xf(v.z, s4.z, h4.z, e4.z) + r.z; v.w = xf(v.w, s4.w, h4.w, e4.w) + r.w;
            *reinterpret_cast<float4 *>(sa + i * 4096) = tf32x4(v);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 v = *reinterpret_cast<const float4 *>(sr + i * 4096);
            v.x = xf(v.x, s4.x, h4.x, e4.x); v.y = xf(v.y, s4.y, h4.y, e4.y);
            v.z = xf(v.z, s4.z, h4.z, e4.z); v.w = xf(v.w, s4.w, h4.w, e4.w);
            *reinterpret_cast<float4 *>(sa + i * 4096) = tf32x4(v);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&bar_full[stage]);
        s4 = ns4; h4 = nh4; e4 = ne4;
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA ISSUER ==============================================
    // the whole warp walks the pipeline (so it stays convergent for the block-wide barrier at the end);
    // lane 0 alone issues tcgen05.mma / tcgen05.commit
    if (WRES) mbar_wait(&bar_wready, 0);
    int stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < plan.total_items; item += gridDim.x) {
      mbar_wait(&bar_tempty[acc], (uint32_t)(acc_phase ^ 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&bar_full[stage], (uint32_t)phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t sa = smem_u32(s_stages + (size_t)stage * kStageBytes);
          const uint32_t sb = WRES ? smem_u32(s_wres + (size_t)kc * kBTileBytes) : sa + kATileBytes;
          const uint64_t adesc = make_desc(sa), bdesc = make_desc(sb);
#pragma unroll
          for (int k8 = 0; k8 < kTcBK / 8; ++k8)
            umma_tf32(d_tmem, adesc + (uint64_t)(k8 * 2), bdesc + (uint64_t)(k8 * 2), idesc, (kc | k8) ? 1u : 0u);
          umma_commit(&bar_empty[stage]);
          if (kc == nk - 1) umma_commit(&bar_tfull[acc]);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (plan.pdl_late) pdl_launch_dependents();
  } else {
    // =============================== EPILOGUE ================================================
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32)
    // statistics pairs nobody reads in the 32-column block that starts at column n (PdrGemmArgs.stats_skip / stats_skip_blocks)
    auto skip_of = [&](int n) {
      const int blk = n >> 5;
      return a.stats_skip | (blk < 32 ? (int)((a.stats_skip_blocks >> (2 * blk)) & 3ull) : 0);
    };
    const int half = warp >> 2;                   // which of the two warps sharing this lane quarter
    constexpr bool VEC = EPI == 1, POOL = EPI == 2, TMA = EPI == 3;
    // epilogue block width: 32 columns; 16 for the narrowest tile so that all 8 warps have work there (the pooling
    // epilogue keeps whole 32-row groups in one warp instead: lane = column, rows = the K neighbours of 32 / K points)
    constexpr int CW = (BN == 32 && !POOL && !TMA) ? 16 : 32;
    uint32_t tma_buf = 0;                         // EPI 3: which of this warp's two staging tiles is next
    // transpose tile row stride (floats).  CW + 4 keeps rows 16-byte aligned, so a lane parks its row with CW / 4
    // STS.128 (conflict-free per quarter warp: lane * 36 floats = lane * 4 banks) and the column-wise read-back
    // (bank = 4 r + lane) is conflict-free as well.  The 16-column blocks keep the odd stride: with 20 the two lane
    // halves (rows r and r + 16) would collide.
    constexpr int kTs = (VEC || CW == 32) ? CW + 4 : CW + 1;
    float *s_t = s_epi[warp];
    // Two groups of 4 warps (one warp per TMEM lane quarter each).  epi_alt: group g owns accumulator g and every
    // second tile of this CTA, so the epilogues of two tiles overlap -- the small-tile GEMMs are bound by the LATENCY
    // of one tile's epilogue (TMEM load -> transpose -> stores -> statistics barrier), not by its instruction count.
    // Otherwise both groups work on the same tile and alternate its column blocks.
    const bool alt = plan.epi_alt != 0;
    int acc = alt ? half : 0, acc_phase = 0;
    const int G = (int)gridDim.x;
    const int item_step = alt ? 2 * G : G;
    const int cb0 = alt ? 0 : half * CW, cb_step = alt ? CW : 2 * CW;
    // column partials: one buffer per group, and (narrow tiles) per tile parity, so that a tile's partials can be
    // summed without a second barrier before the next tile's are written
    constexpr bool kPartDB = BN <= 128;
    constexpr uint32_t kPartBytes = 4u * BN * 16u;
    const uint32_t s_part_base = s_part + (alt ? (uint32_t)half * (kPartDB ? 2u : 1u) * kPartBytes : 0u);
    uint32_t part_parity = 0;
    const int stat_bar = alt ? 1 + half : 1, stat_threads = alt ? kEpiThreads / 2 : kEpiThreads;
    const int stat_tid = alt ? (tid & (kEpiThreads / 2 - 1)) : tid;
    const bool radd_split = a.rowadd && a.rows_per_sample % a.rowadd_div == 0;   // groups never straddle samples
    const int groups_per_sample = a.rowadd ? a.rows_per_sample / a.rowadd_div : 0;
    for (int item = (int)blockIdx.x + (alt ? half * G : 0); item < plan.total_items; item += item_step) {
      const int tile = item / plan.n_tiles_n, n0 = (item - tile * plan.n_tiles_n) * bn;
      const int b = tile / plan.tiles_per_sample, tis = tile - b * plan.tiles_per_sample;
      const uint32_t s_part_g = s_part_base + (kPartDB ? part_parity * kPartBytes : 0u);
      const int r0 = tis * kTcTileM;
      const size_t row_base = (size_t)b * a.rows_per_sample + r0;
      const int rows_valid = min(kTcTileM, a.rows_per_sample - r0);
      const int wrows = max(0, min(32, rows_valid - quarter * 32));   // valid rows among this warp's 32
      const size_t wrow0 = row_base + quarter * 32;                   // first global row of this warp
      size_t radd_g0 = 0;
      int radd_rem0 = 0;
      if (a.rowadd) {
        if (radd_split) {
          const int g = (r0 + quarter * 32) / a.rowadd_div;
          radd_rem0 = r0 + quarter * 32 - g * a.rowadd_div;
          radd_g0 = (size_t)b * groups_per_sample + g;
        } else {
          radd_g0 = wrow0 / (size_t)a.rowadd_div; radd_rem0 = (int)(wrow0 % (size_t)a.rowadd_div);
        }
      }
      // EPI 3 with the broadcast row-add: the row groups this warp's valid rows span (at most 5: launch_tc keeps
      // rowadd_div >= 8 here) and the group of my row (lane = row)
      int tma_ng = 0, tma_gi = 0;
      if constexpr (TMA) {
        if (a.rowadd && wrows > 0) {
          tma_ng = (radd_rem0 + wrows - 1) / a.rowadd_div + 1;
          tma_gi = min((radd_rem0 + lane) / a.rowadd_div, tma_ng - 1);
        }
      }
      // EPI 3: the per-column addends (bias + the broadcast row of each group; lane = column) of the tile's first block
      // are requested before the accumulator is waited for, so that single-block tiles (N <= 32) have no global latency
      // between the TMEM load and the bulk store
      float addn[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      auto load_addends = [&](int cb_) {
        const int n = n0 + cb_ + lane;
        const bool nin = n < a.N;
        const float bias_n = (nin && a.bias) ? __ldg(a.bias + n) : 0.f;
        addn[0] = bias_n;
        if (a.rowadd) {
          // same association as the scalar flavour: y = acc + (bias + rowadd)
          const float *rp = a.rowadd + radd_g0 * a.ld_rowadd + (nin ? n : 0);
#pragma unroll
          for (int g = 0; g < 5; ++g)
            if (g < tma_ng) addn[g] = bias_n + (nin ? __ldg(rp + (size_t)g * a.ld_rowadd) : 0.f);
        }
      };
      if constexpr (TMA) load_addends(cb0);
      mbar_wait(&bar_tfull[acc], (uint32_t)acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cb = cb0; cb < bn; cb += cb_step) {
        // CW-column blocks: CW = 32 normally, 16 for the narrowest tile (BN = 32)
        if (n0 + cb >= a.N && (POOL || n0 + cb >= a.ldc_zero_to)) break;   // nothing to write in this or later blocks
        uint32_t v[CW];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + cb);
        if constexpr (CW == 32) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr));
        } else {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(taddr));
        }
        if constexpr (POOL) {
          // ---- soft-attention pooling fused into the score GEMM (AttentionModule, attention.py:85-96): this GEMM's
          //      output IS the score tensor; instead of storing it, every lane (= output channel) runs the masked
          //      softmax over the K neighbour rows of each point held by this warp and accumulates the GroupNorm-ed,
          //      ReLU-ed values read from V.  Same operation order as attention_pool_kernel -> same bits. ----
          const int n = n0 + cb + lane;
          const bool nin = n < a.N;
          const float bias_n = (nin && a.bias) ? __ldg(a.bias + n) : 0.f;
          const float gs = nin ? __ldg(a.pool_sc + (size_t)b * a.pool_ld_scsh + n) : 0.f;
          const float gh = nin ? __ldg(a.pool_sh + (size_t)b * a.pool_ld_scsh + n) : 0.f;
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if constexpr (kTs % 4 == 0) {
  #pragma unroll
            for (int j = 0; j < CW / 4; ++j)
              *reinterpret_cast<float4 *>(s_t + lane * kTs + 4 * j) =
                  make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                              __uint_as_float(v[4 * j + 3]));
          } else {
  #pragma unroll
            for (int j = 0; j < CW; ++j) s_t[lane * kTs + j] = __uint_as_float(v[j]);
          }
          __syncwarp();
          const int PK = a.pool_K;
          const float *st = s_t + lane;
          const float *vp = a.pool_V + wrow0 * (size_t)a.pool_ldv + (nin ? n : 0);
          for (int r0g = 0; r0g < wrows; r0g += PK) {             // wrows is a multiple of PK (rows_per_sample % PK == 0)
            const size_t point = (wrow0 + r0g) / (size_t)PK;      // b*P + p
            int cnt = PK;
            if (a.pool_counts) { cnt = __ldg(a.pool_counts + point); cnt = cnt < 1 ? 1 : cnt; }
            float mx = -3.0e38f;
  #pragma unroll 8
            for (int k = 0; k < PK; ++k) mx = fmaxf(mx, k < cnt ? st[(r0g + k) * kTs] + bias_n : -1e9f);
            float den = 0.f, num = 0.f;
            const float *vk = vp + (size_t)r0g * a.pool_ldv;
  #pragma unroll 8
            for (int k = 0; k < PK; ++k) {
              const float sk = k < cnt ? st[(r0g + k) * kTs] + bias_n : -1e9f;
              const float e = expf(sk - mx);
              den += e;
              num = fmaf(e, fmaxf(fmaf(__ldg(vk), gs, gh), 0.f), num);
              vk += a.pool_ldv;
            }
            if (nin) a.pool_out[point * (size_t)a.pool_ldo + n] = num / den;
          }
          __syncwarp();                                         // before the next block overwrites the tile
        } else if constexpr (VEC) {
          // lane mapping of the store phase: LPR lanes cover one row of the block as float4s, a store instruction
          // writes RPI rows, a lane walks ITERS consecutive rows.  The per-column loads are issued before the TMEM
          // load is waited for.
          constexpr int LPR = CW / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
          const int rsub = lane / LPR, cg = lane % LPR;
          const int n = n0 + cb + 4 * cg;                   // first of this lane's 4 columns
          const int lim = max(a.N, a.ldc_zero_to);          // columns < lim are written (values or zero padding)
          const bool in0 = n < a.N, in1 = n + 1 < a.N, in2 = n + 2 < a.N, in3 = n + 3 < a.N;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias) {
            if (in0) bias4.x = __ldg(a.bias + n);
            if (in1) bias4.y = __ldg(a.bias + n + 1);
            if (in2) bias4.z = __ldg(a.bias + n + 2);
            if (in3) bias4.w = __ldg(a.bias + n + 3);
          }
          const int my_rows = max(0, min(ITERS, wrows - rsub * ITERS));
          // broadcast row-add: rows (wrow0 + r) / div share one row of `rowadd` (the query term of AttentionModule,
          // expanded over the K neighbours)
          const float *rp = nullptr;
          int rem = 0;
          float4 cur = bias4;
          if (a.rowadd) {
            const int first = radd_rem0 + rsub * ITERS;
            const int gskip = first / a.rowadd_div;
            rem = first - gskip * a.rowadd_div;
            rp = a.rowadd + (radd_g0 + gskip) * a.ld_rowadd + n;
            if (in0 && my_rows > 0) {
              const float4 q = __ldg(reinterpret_cast<const float4 *>(rp));
              cur.x += q.x; cur.y += q.y; cur.z += q.z; cur.w += q.w;
            }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // park my row (lane = row) in the transpose tile as float4s; the row stride CW + 4 floats keeps 16-byte
          // alignment and both phases bank-conflict free
  #pragma unroll
          for (int j = 0; j < CW / 4; ++j)
            *reinterpret_cast<float4 *>(s_t + lane * kTs + 4 * j) =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
          __syncwarp();
          const size_t ldc = (size_t)a.ldc;
          float *cp = a.C + (wrow0 + rsub * ITERS) * ldc + n;
          const float *st = s_t + (rsub * ITERS) * kTs + 4 * cg;
          const bool st_all = n + 3 < lim, st_any = n < lim;
          // statistics of the 4 columns as packed pairs (columns 0|1 and 2|3): sum, sum of squares, relu-sum,
          // relu-sum of squares -- one FADD2 / FFMA2 serves two columns
          unsigned long long s01 = 0ull, s23 = 0ull, q01 = 0ull, q23 = 0ull, rs01 = 0ull, rs23 = 0ull, rq01 = 0ull, rq23 = 0ull;
          auto accum2 = [&](float x, float y, unsigned long long &sm, unsigned long long &sq, unsigned long long &rs,
                            unsigned long long &rq) {
            unsigned long long t2, p2;
            asm("mov.b64 %0, {%1, %2};" : "=l"(t2) : "f"(x), "f"(y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(fmaxf(x, 0.f)), "f"(fmaxf(y, 0.f)));
            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sm) : "l"(t2));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(sq) : "l"(t2));
            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(rs) : "l"(p2));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(rq) : "l"(p2));
          };
          // fast path (warp-uniform): full rows, one broadcast row per lane, no float4 straddling the column limit
          const bool one_group = !a.rowadd || rem + ITERS <= a.rowadd_div;
          const bool fast = __all_sync(0xffffffffu, my_rows == ITERS && one_group && (st_all || !st_any));
          if (fast) {
            // the hot loop of the epilogue warps (they are issue-bound: profiles/r01_ncu_gemm_epilogue_hotspots_v5.txt),
            // specialised at compile time on what is needed: P = (sum, sum^2), R = the relu pair, AI = all four columns
            // of every lane are inside N (no select, packed bias add).  ~3-5 instructions per element.
            const bool all_in = __all_sync(0xffffffffu, in3);
            const int sk = skip_of(n0 + cb);
            const bool needP = a.stats && !(sk & 1), needR = a.stats && !(sk & 2);
            unsigned long long cur01, cur23;
            asm("mov.b64 %0, {%1, %2};" : "=l"(cur01) : "f"(cur.x), "f"(cur.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(cur23) : "f"(cur.z), "f"(cur.w));
            auto loop = [&](auto p_c, auto r_c, auto ai_c) {
              constexpr bool P = decltype(p_c)::value, R = decltype(r_c)::value, AI = decltype(ai_c)::value;
  #pragma unroll
              for (int r = 0; r < ITERS; ++r) {
                float4 t = *reinterpret_cast<const float4 *>(st + r * kTs);
                unsigned long long t01, t23;
                if constexpr (AI) {
                  asm("mov.b64 %0, {%1, %2};" : "=l"(t01) : "f"(t.x), "f"(t.y));
                  asm("mov.b64 %0, {%1, %2};" : "=l"(t23) : "f"(t.z), "f"(t.w));
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(t01) : "l"(cur01));
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(t23) : "l"(cur23));
                  asm("mov.b64 {%0, %1}, %2;" : "=f"(t.x), "=f"(t.y) : "l"(t01));
                  asm("mov.b64 {%0, %1}, %2;" : "=f"(t.z), "=f"(t.w) : "l"(t23));
                  *reinterpret_cast<float4 *>(cp) = t;
                } else {
                  t.x = in0 ? t.x + cur.x : 0.f; t.y = in1 ? t.y + cur.y : 0.f;
                  t.z = in2 ? t.z + cur.z : 0.f; t.w = in3 ? t.w + cur.w : 0.f;
                  if (st_all) *reinterpret_cast<float4 *>(cp) = t;
                  asm("mov.b64 %0, {%1, %2};" : "=l"(t01) : "f"(t.x), "f"(t.y));
                  asm("mov.b64 %0, {%1, %2};" : "=l"(t23) : "f"(t.z), "f"(t.w));
                }
                cp += ldc;
                if constexpr (P) {
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s01) : "l"(t01));
                  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q01) : "l"(t01));
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s23) : "l"(t23));
                  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q23) : "l"(t23));
                }
                if constexpr (R) {
                  unsigned long long p01, p23;
                  asm("mov.b64 %0, {%1, %2};" : "=l"(p01) : "f"(fmaxf(t.x, 0.f)), "f"(fmaxf(t.y, 0.f)));
                  asm("mov.b64 %0, {%1, %2};" : "=l"(p23) : "f"(fmaxf(t.z, 0.f)), "f"(fmaxf(t.w, 0.f)));
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(rs01) : "l"(p01));
                  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(rq01) : "l"(p01));
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(rs23) : "l"(p23));
                  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(rq23) : "l"(p23));
                }
              }
            };
            using T = std::true_type; using F = std::false_type;
            const int variant = (needP ? 4 : 0) | (needR ? 2 : 0) | (all_in ? 1 : 0);
            switch (variant) {
              case 7: loop(T{}, T{}, T{}); break;
              case 6: loop(T{}, T{}, F{}); break;
              case 5: loop(T{}, F{}, T{}); break;
              case 4: loop(T{}, F{}, F{}); break;
              case 3: loop(F{}, T{}, T{}); break;
              case 2: loop(F{}, T{}, F{}); break;
              case 1: loop(F{}, F{}, T{}); break;
              default: loop(F{}, F{}, F{}); break;
            }
          } else {
  #pragma unroll 1
            for (int r = 0; r < my_rows; ++r) {
              if (a.rowadd && rem == a.rowadd_div) {          // next broadcast row
                rem = 0;
                rp += a.ld_rowadd;
                cur = bias4;
                if (in0) {
                  const float4 q = __ldg(reinterpret_cast<const float4 *>(rp));
                  cur.x += q.x; cur.y += q.y; cur.z += q.z; cur.w += q.w;
                }
              }
              ++rem;
              float4 t = *reinterpret_cast<const float4 *>(st + r * kTs);
              t.x = in0 ? t.x + cur.x : 0.f; t.y = in1 ? t.y + cur.y : 0.f;
              t.z = in2 ? t.z + cur.z : 0.f; t.w = in3 ? t.w + cur.w : 0.f;
              if (st_all) {
                *reinterpret_cast<float4 *>(cp) = t;
              } else if (st_any) {
                cp[0] = t.x;
                if (n + 1 < lim) cp[1] = t.y;
                if (n + 2 < lim) cp[2] = t.z;
              }
              cp += ldc;
              accum2(t.x, t.y, s01, q01, rs01, rq01);
              accum2(t.z, t.w, s23, q23, rs23, rq23);
            }
          }
          __syncwarp();                                         // every lane is done with the transpose tile
          if (a.stats) {
            // fold the RPI row sub-groups through the (now free) transpose tile in a fixed order: quad index
            // rsub * CW + c * LPR + cg holds (sum, sumsq, relu-sum, relu-sumsq) of column 4 * cg + c
            {
              float sa, sb, qa, qb, ra, rb, ua, ub;
              asm("mov.b64 {%0, %1}, %2;" : "=f"(sa), "=f"(sb) : "l"(s01));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(qa), "=f"(qb) : "l"(q01));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(ra), "=f"(rb) : "l"(rs01));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(ua), "=f"(ub) : "l"(rq01));
              *reinterpret_cast<float4 *>(s_t + (rsub * CW + 0 * LPR + cg) * 4) = make_float4(sa, qa, ra, ua);
              *reinterpret_cast<float4 *>(s_t + (rsub * CW + 1 * LPR + cg) * 4) = make_float4(sb, qb, rb, ub);
              asm("mov.b64 {%0, %1}, %2;" : "=f"(sa), "=f"(sb) : "l"(s23));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(qa), "=f"(qb) : "l"(q23));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(ra), "=f"(rb) : "l"(rs23));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(ua), "=f"(ub) : "l"(rq23));
              *reinterpret_cast<float4 *>(s_t + (rsub * CW + 2 * LPR + cg) * 4) = make_float4(sa, qa, ra, ua);
              *reinterpret_cast<float4 *>(s_t + (rsub * CW + 3 * LPR + cg) * 4) = make_float4(sb, qb, rb, ub);
            }
            __syncwarp();
            if (lane < CW) {
              float4 tot = *reinterpret_cast<const float4 *>(s_t + lane * 4);
  #pragma unroll
              for (int rs = 1; rs < RPI; ++rs) {
                const float4 q = *reinterpret_cast<const float4 *>(s_t + (rs * CW + lane) * 4);
                tot.x += q.x; tot.y += q.y; tot.z += q.z; tot.w += q.w;
              }
              const int col = 4 * (lane % LPR) + lane / LPR;
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(s_part_g + (uint32_t)((quarter * BN + cb + col) * 16)),
                           "f"(tot.x), "f"(tot.y), "f"(tot.z), "f"(tot.w) : "memory");
            }
            __syncwarp();                                       // before the next block overwrites the tile
          }
        } else if constexpr (TMA) {
          // ---- TMA-store epilogue: lane = row parks its 32 columns (+ bias) in a 32 x 32 staging tile laid out as the
          //      tensor map expects (128-byte rows, 16-byte chunk j of row r at j ^ (r & 7)), one lane issues the bulk
          //      store (rows beyond the sample and columns beyond max(N, ldc_zero_to) are clipped by the TMA unit), and the
          //      column statistics are read back from the same tile, lane = column, while the store drains.  No
          //      per-element STG, no 64-bit address arithmetic: ~3 (store only) to ~9 instructions per element-row. ----
          // (later blocks fetch theirs here: requesting them during the previous block was measured ~10 % slower on every
          //  multi-block tile, gpurun call r01s3d -- the loads in flight share a scoreboard with the TMEM load)
          if (cb != cb0) load_addends(cb);
#pragma unroll
          for (int g = 0; g < 5; ++g)
            if (g == 0 || g < tma_ng) s_t[g * 36 + lane] = addn[g];
          __syncwarp();
          // the 32 addends of my row (broadcast reads): the first half is fetched under the TMEM load, the second under
          // the first four stores -- all eight at once cost 16 more live registers than the 96 this kernel can have
          const float *sadd = s_t + tma_gi * 36;
          float4 b4[4];
  #pragma unroll
          for (int j = 0; j < 4; ++j) b4[j] = *reinterpret_cast<const float4 *>(sadd + 4 * j);
          // the store that last read the tile about to be overwritten (two blocks ago) must have finished reading
          if (lane == 0) {
            if constexpr (tma_bufs(BN) == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          __syncwarp();
          // (explicit shared-space accesses: through a generic pointer into the dynamic region these would be LD.E/ST.E)
          const uint32_t tile = s_tma + ((uint32_t)warp * (uint32_t)tma_bufs(BN) + tma_buf) * 4096u;
          if constexpr (tma_bufs(BN) == 2) tma_buf ^= 1u;
          {
            // chunk j of my row sits at row + ((j ^ (lane & 7)) << 4); the row start is 128-byte aligned, so that is
            // (row | (lane & 7) << 4) ^ (j << 4): one LOP3 with an immediate per store.  The addends go on in packed pairs.
            const uint32_t trow = (tile + (uint32_t)lane * 128u) | (((uint32_t)lane & 7u) << 4);
            float4 c4[4];
  #pragma unroll
            for (int j = 0; j < 4; ++j) c4[j] = *reinterpret_cast<const float4 *>(sadd + 16 + 4 * j);
  #pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = j < 4 ? b4[j] : c4[j - 4];
              unsigned long long y01, y23, b01, b23;
              asm("mov.b64 %0, {%1, %2};" : "=l"(y01) : "r"(v[4 * j]), "r"(v[4 * j + 1]));
              asm("mov.b64 %0, {%1, %2};" : "=l"(y23) : "r"(v[4 * j + 2]), "r"(v[4 * j + 3]));
              asm("mov.b64 %0, {%1, %2};" : "=l"(b01) : "f"(b.x), "f"(b.y));
              asm("mov.b64 %0, {%1, %2};" : "=l"(b23) : "f"(b.z), "f"(b.w));
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(y01) : "l"(b01));
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(y23) : "l"(b23));
              asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(trow ^ ((uint32_t)j << 4)), "l"(y01), "l"(y23) : "memory");
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (wrows > 0)
              asm volatile(
                  "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                      reinterpret_cast<uint64_t>(&tmap_c)),
                  "r"(n0 + cb), "r"(r0 + quarter * 32), "r"(b), "r"(tile)
                  : "memory");
            // always a group (possibly empty): wait_group.read 1 above counts groups, one per block
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (a.stats) {
            const int sk = skip_of(n0 + cb);
            const bool needP = !(sk & 1), needR = !(sk & 2);
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
            // element (r, c) sits at r * 128 + (((c >> 2) ^ (r & 7)) << 4) + (c & 3) * 4 bytes: 8 per-lane offsets, one
            // per value of r & 7, then immediates
            uint32_t toff[8];
  #pragma unroll
            for (int k = 0; k < 8; ++k)      // tile is 4 KiB aligned: (tile | chunk << 4 | word << 2) ^ (k << 4)
              toff[k] = (tile | (((uint32_t)lane >> 2) << 4) | (((uint32_t)lane & 3u) << 2)) ^ ((uint32_t)k << 4);
            // two rows per step on packed fp32 pairs (add / fma.rn.f32x2): the epilogue warps are bound by instruction
            // issue, and the even and the odd rows get one partial sum each, added at the end
            unsigned long long s01 = 0ull, q01 = 0ull, r01 = 0ull, u01 = 0ull;
            auto rows_loop = [&](auto p_c, auto r_c, auto full_c) {
              constexpr bool P = decltype(p_c)::value, R = decltype(r_c)::value, FULL = decltype(full_c)::value;
  #pragma unroll
              for (int r = 0; r < 32; r += 2) {
                if (!FULL && r >= wrows) break;
                float t0, t1;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t0) : "r"(toff[r & 7] + (uint32_t)r * 128u) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t1) : "r"(toff[(r + 1) & 7] + (uint32_t)(r + 1) * 128u) : "memory");
                if (!FULL && r + 1 >= wrows) t1 = 0.f;          // a row beyond the sample adds nothing to any of the sums
                unsigned long long t2;
                asm("mov.b64 %0, {%1, %2};" : "=l"(t2) : "f"(t0), "f"(t1));
                if constexpr (P) {
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s01) : "l"(t2));
                  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q01) : "l"(t2));
                }
                if constexpr (R) {
                  unsigned long long p2;
                  asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(fmaxf(t0, 0.f)), "f"(fmaxf(t1, 0.f)));
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(r01) : "l"(p2));
                  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(u01) : "l"(p2));
                }
              }
            };
            using T = std::true_type; using F = std::false_type;
            if (wrows == 32) {
              if (needP && needR) rows_loop(T{}, T{}, T{});
              else if (needP) rows_loop(T{}, F{}, T{});
              else if (needR) rows_loop(F{}, T{}, T{});
            } else {
              rows_loop(T{}, T{}, F{});
            }
            {
              float lo, hi;
              asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s01)); q0 = lo + hi;
              asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q01)); q1 = lo + hi;
              asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r01)); q2 = lo + hi;
              asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(u01)); q3 = lo + hi;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(s_part_g + (uint32_t)((quarter * BN + cb + lane) * 16)),
                         "f"(q0), "f"(q1), "f"(q2), "f"(q3) : "memory");
          }
          __syncwarp();                                         // s_t (the addends) is rewritten by the next block
        } else {
          // per-column constants are fetched while the TMEM load is in flight
          const int hi = lane / CW, cl = lane % CW;
          const int n = n0 + cb + cl;
          const bool nin = n < a.N;
          const bool nstore = nin || n < a.ldc_zero_to;
          const float bias_n = (nin && a.bias) ? __ldg(a.bias + n) : 0.f;
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // park my row (lane = row) in the transpose tile; the odd stride keeps both phases bank-conflict free
          if constexpr (kTs % 4 == 0) {
  #pragma unroll
            for (int j = 0; j < CW / 4; ++j)
              *reinterpret_cast<float4 *>(s_t + lane * kTs + 4 * j) =
                  make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                              __uint_as_float(v[4 * j + 3]));
          } else {
  #pragma unroll
            for (int j = 0; j < CW; ++j) s_t[lane * kTs + j] = __uint_as_float(v[j]);
          }
          __syncwarp();
          // from here on lane = (row group hi, column cl): a store instruction writes 32 / CW row segments of CW
          // consecutive floats; bias and the broadcast row-add are per column
          constexpr int kRowsPer = CW;                  // rows walked by one lane group (32 rows * CW / 32 lanes)
          float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
          const size_t ldc = (size_t)a.ldc;
          float *cp = a.C + (wrow0 + kRowsPer * hi) * ldc + n;
          const float *st = s_t + (kRowsPer * hi) * kTs + cl;
          const int my_rows = max(0, min(kRowsPer, wrows - kRowsPer * hi));
          // the loops below are the hot path of the epilogue warps: keep them branch-free and free of 64-bit
          // index arithmetic (running pointers only)
          if (!a.rowadd) {
            // specialised on the statistics pairs some GroupNorm will read (PdrGemmArgs.stats_skip)
            auto rows_loop = [&](auto p_c, auto r_c) {
              constexpr bool P = decltype(p_c)::value, R = decltype(r_c)::value;
  #pragma unroll 8
              for (int r = 0; r < my_rows; ++r) {
                float t = st[r * kTs] + bias_n;
                t = nin ? t : 0.f;
                if (nstore) *cp = t;
                cp += ldc;
                if constexpr (P) { q0 += t; q1 = fmaf(t, t, q1); }
                if constexpr (R) { const float p = fmaxf(t, 0.f); q2 += p; q3 = fmaf(p, p, q3); }
              }
            };
            using T = std::true_type; using F = std::false_type;
            const int sk = skip_of(n0 + cb);
            const bool needP = a.stats && !(sk & 1), needR = a.stats && !(sk & 2);
            if (needP && needR) rows_loop(T{}, T{});
            else if (needP) rows_loop(T{}, F{});
            else if (needR) rows_loop(F{}, T{});
            else rows_loop(F{}, F{});
          } else {
            // rows (wrow0 + r) / div share one broadcast row (the query term of AttentionModule, expanded over
            // the K neighbours): one load per group of rows, issued one group ahead of its use
            const int first = radd_rem0 + kRowsPer * hi;
            const int gskip = first / a.rowadd_div;
            int rem = first - gskip * a.rowadd_div;
            const float *rp = a.rowadd + (radd_g0 + gskip) * a.ld_rowadd + (nin ? n : 0);
            int r = 0;
            float nxt = (nin && my_rows > 0) ? __ldg(rp) : 0.f;
            while (r < my_rows) {
              const float cur = bias_n + nxt;
              const int rend = min(my_rows, r + (a.rowadd_div - rem));
              rp += a.ld_rowadd;
              if (nin && rend < my_rows) nxt = __ldg(rp);
  #pragma unroll 8
              for (; r < rend; ++r) {
                float t = st[r * kTs] + cur;
                t = nin ? t : 0.f;
                if (nstore) *cp = t;
                cp += ldc;
                const float p = fmaxf(t, 0.f);
                q0 += t; q1 = fmaf(t, t, q1); q2 += p; q3 = fmaf(p, p, q3);
              }
              rem = 0;
            }
          }
          __syncwarp();
          if (a.stats) {
            if constexpr (CW == 16) {
              q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
              q2 += __shfl_xor_sync(0xffffffffu, q2, 16); q3 += __shfl_xor_sync(0xffffffffu, q3, 16);
            }
            if (hi == 0)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(s_part_g + (uint32_t)((quarter * BN + cb + cl) * 16)),
                           "f"(q0), "f"(q1), "f"(q2), "f"(q3) : "memory");
          }
        }
      }
      // all TMEM reads of this accumulator are complete: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&bar_tempty[acc]);
      if (a.stats) {
        asm volatile("bar.sync %0, %1;" ::"r"(stat_bar), "r"(stat_threads) : "memory");
        for (int f = stat_tid; f < bn * 4; f += stat_threads) {
          const int col = f >> 2, q = f & 3;
          if (n0 + col < a.N) {
            float p0, p1, p2, p3;
            const uint32_t pa = s_part_g + (uint32_t)(f * 4);
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p0) : "r"(pa) : "memory");
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p1) : "r"(pa + BN * 16) : "memory");
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p2) : "r"(pa + 2 * BN * 16) : "memory");
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p3) : "r"(pa + 3 * BN * 16) : "memory");
            a.stats[((size_t)tile * a.N + n0 + col) * 4 + q] = p0 + p1 + p2 + p3;
          }
        }
        if constexpr (kPartDB) part_parity ^= 1u;
        else asm volatile("bar.sync %0, %1;" ::"r"(stat_bar), "r"(stat_threads) : "memory");
      }
      if (alt) acc_phase ^= 1;
      else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    // the staging tiles must outlive the bulk stores that read them
    if (TMA && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
                 : "memory");
  }
}

constexpr int kPlanDoesNotFit = 12345;

// Epilogue flavour.  Measured on B200:
//  * profiles/r01_epilogue_ab_v7.txt: float4 / packed-f32x2 beats scalar only where the broadcast row-add is used on
//    32-column blocks (0.240 -> 0.177 ms on the 524288 x 172 x 128 score GEMM); scalar wins elsewhere.
//  * profiles/r01_tma_epilogue_ab_v10.txt: the TMA-store flavour takes the dense 2 M x 32 x 32 GEMMs from 3.9 to
//    4.9-5.9 TB/s and is neutral to +3 % elsewhere.
// Default = tma (row groups of at least 8 rows; shorter ones fall back to the old choice).
// PDR_GEMM_BALANCED=0: column tiles of the full template width (A/B of TcPlan.bn)
bool balanced_tiles() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("PDR_GEMM_BALANCED"); on = !(e && e[0] == '0'); }
  return on == 1;
}
// PDR_GEMM_EPILOGUE=auto|scalar|vec4|tma forces one (tests, A/B); auto = the pre-TMA hybrid.
int epilogue_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("PDR_GEMM_EPILOGUE");
    mode = !e ? 3 : (e[0] == 's' ? 0 : (e[0] == 'v' ? 1 : (e[0] == 't' ? 3 : 2)));
  }
  return mode;
}

// cuTensorMapEncodeTiled without linking libcuda: the entry point comes from the runtime
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

// C seen by the TMA unit: (columns written, rows of one sample, samples), 32 x 32 x 1 boxes, 128-byte swizzle.
// Clipping at dims 0 and 1 is what keeps a partial tile from touching the pad columns / the next sample.
int make_c_tensor_map(const PdrGemmArgs &a, CUtensorMap *tm) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("gemm_tf32: cuTensorMapEncodeTiled is not available"); return PDR_ERR_UNSUPPORTED; }
  const int cols = a.N > a.ldc_zero_to ? a.N : a.ldc_zero_to;
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)a.rows_per_sample, (cuuint64_t)a.batch};
  const cuuint64_t gstr[2] = {(cuuint64_t)a.ldc * 4u, (cuuint64_t)a.rows_per_sample * (cuuint64_t)a.ldc * 4u};
  const cuuint32_t box[3] = {32u, 32u, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)a.C, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tf32: cuTensorMapEncodeTiled failed (%d)", (int)r); return PDR_ERR_CUDA; }
  return 0;
}

// the gathered table (PdrGemmArgs.A with a_rows) as a 2-D map for TMA tile::gather4: box = one 128-byte chunk of one row
int make_table_tensor_map(const PdrGemmArgs &a, CUtensorMap *tm) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("gemm_tf32: cuTensorMapEncodeTiled is not available"); return PDR_ERR_UNSUPPORTED; }
  const cuuint64_t gdim[2] = {(cuuint64_t)(a.k_split / kTcBK * kTcBK), (cuuint64_t)a.table_rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)a.lda * 4u};
  const cuuint32_t box[2] = {(cuuint32_t)kTcBK, 1u};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)a.A, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tf32: cuTensorMapEncodeTiled (gathered table) failed (%d)", (int)r); return PDR_ERR_CUDA; }
  return 0;
}
// PDR_GEMM_PROD_SLEEP=<ns>: back-off of the producers' empty-stage polls (0 = spin on try_wait)
int producer_sleep_ns() {
  static int ns = -1;
  if (ns < 0) {
    const char *e = getenv("PDR_GEMM_PROD_SLEEP");
    ns = e ? atoi(e) : 0;
    if (ns < 0) ns = 0;
  }
  return ns;
}

bool epilogue_alternates_tiles() {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("PDR_GEMM_EPI_ALT");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode == 1;
}

template <int BN, bool WRES>
int launch_tc(const PdrGemmArgs &a, cudaStream_t stream) {
  TcPlan plan;
  plan.n_tiles_n = ceil_div(a.N, BN);
  plan.bn = BN == 256 && balanced_tiles() ? ceil_div(ceil_div(a.N, plan.n_tiles_n), 32) * 32 : BN;
  plan.tiles_per_sample = ceil_div(a.rows_per_sample, kTcTileM);
  const long long items = (long long)plan.n_tiles_n * a.batch * plan.tiles_per_sample;
  if (items > 0x7fffffffll) { set_error("gemm_tf32: too many tiles"); return PDR_ERR_INVALID_ARGUMENT; }
  plan.total_items = (int)items;
  plan.nk = ceil_div(a.K, kTcBK);
  plan.epi_alt = epilogue_alternates_tiles() ? 1 : 0;
  plan.prod_sleep_ns = producer_sleep_ns();
  plan.pdl_late = pdl_mode() == 2;
  {
    static int lean = -1;
    if (lean < 0) { const char *e = getenv("PDR_GEMM_LEAN"); lean = !(e && e[0] == '0'); }
    plan.lean = lean;
  }
  plan.w_early = pdl_mode() != 0 && a.w_static;
  // epilogue flavour: pooling when asked for; float4 for the broadcast row-add on 32-column blocks; otherwise scalar, or
  // (PDR_GEMM_EPILOGUE=tma) the TMA-store flavour (row groups of at least 8 rows)
  const int mode = epilogue_mode();
  int vec;
  if (a.pool_K > 0) vec = 2;
  else if (mode == 3) vec = (a.rowadd && a.rowadd_div < 8) ? (BN > 32 ? 1 : 0) : 3;
  else vec = (mode == 2 ? (a.rowadd != nullptr && BN > 32) : mode == 1) ? 1 : 0;
  const size_t epi = (size_t)(BN <= 128 ? 4 : 2) * 4 * BN * 16 +   // column partials: 2 epilogue groups (x 2 tile parities)
                     (vec == 3 ? (size_t)tma_stage_bytes(BN) : 0);  // + the staging tiles of the TMA stores
  const size_t static_smem = (size_t)(kEpiWarps * (vec == 3 ? 192 : 32 * 36)) * sizeof(float) + 512;
  const size_t budget = 226 * 1024 - static_smem;
  size_t smem = 0;
  bool planned = false;
  // prefer the direct (no-transform) producer when the GEMM has no prologue
  for (int direct = (a.pro_mode == PDR_PRO_NONE && !a.add && !a.R) ? 1 : 0; direct >= 0 && !planned; --direct) {
    if (a.a_rows && !direct) break;               // gathered A exists in the cp.async (direct) producer only
    plan.direct = direct;
    const size_t w_tile = WRES ? 0 : (size_t)BN * 128;
    const size_t stage = kATileBytes + w_tile + ((!direct && a.R) ? kATileBytes : 0);
    plan.stage_bytes = (int)stage;
    plan.r_off = (int)(kATileBytes + w_tile);
    const size_t fixed = 1024 + epi + (WRES ? (size_t)plan.nk * BN * 128 : 0);
    if (fixed + 2 * stage > budget) continue;
    int stages = (int)((budget - fixed) / stage);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) continue;
    plan.stages = stages;
    smem = fixed + (size_t)stages * stage;
    planned = true;
  }
  if (!planned) {
    if (WRES || BN > 128) return kPlanDoesNotFit;
    if (a.a_rows) { set_error("gemm_tf32: gathered A does not fit (K=%d N=%d)", a.K, a.N); return PDR_ERR_UNSUPPORTED; }   // caller retries with streamed weights / narrower column tiles
    set_error("gemm_tf32: shared memory budget exceeded (K=%d N=%d)", a.K, a.N);
    return PDR_ERR_UNSUPPORTED;
  }
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (vec == 3) {
    const int rc = make_c_tensor_map(a, &tmap);
    if (rc != 0) return rc;
  }
  CUtensorMap tmap_a;
  memset(&tmap_a, 0, sizeof(tmap_a));
  plan.tma_gather = 0;
  if (a.a_rows && plan.direct && a.table_rows > 0 && a.k_split >= kTcBK &&
      (reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && a.lda % 4 == 0) {
    const int rc = make_table_tensor_map(a, &tmap_a);
    if (rc != 0) return rc;
    plan.tma_gather = 1;
  }
  auto kern = vec == 3 ? gemm_tf32_persistent<BN, WRES, 3>
              : vec == 2 ? gemm_tf32_persistent<BN, WRES, 2>
                         : (vec == 1 ? gemm_tf32_persistent<BN, WRES, 1> : gemm_tf32_persistent<BN, WRES, 0>);
  static bool configured[4] = {false, false, false, false};
  const int cfg_slot = vec;
  if (!configured[cfg_slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    if (e != cudaSuccess) { set_error("gemm_tf32: smem attr: %s", cudaGetErrorString(e)); return PDR_ERR_CUDA; }
    configured[cfg_slot] = true;
  }
  const int sm_cap = (a.max_ctas > 0 && a.max_ctas < kNumSMs) ? a.max_ctas : kNumSMs;
  const int grid = plan.total_items < sm_cap ? plan.total_items : sm_cap;
  {
    const cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kTcThreads), smem, stream, a, plan, tmap, tmap_a);
    if (e != cudaSuccess) { set_error("gemm_tf32_persistent: %s", cudaGetErrorString(e)); return PDR_ERR_CUDA; }
  }
  return check_launch("gemm_tf32_persistent");
}

template <int BN>
int dispatch_wres(const PdrGemmArgs &a, cudaStream_t stream) {
  const int nk = ceil_div(a.K, kTcBK);
  // resident weights pay off once a CTA reuses them over several row tiles; with about one tile per CTA the
  // up-front staging is pure latency and streaming (overlapped with the MMAs) wins
  const long long row_tiles = (long long)a.batch * ceil_div(a.rows_per_sample, kTcTileM);
  const bool wres = ceil_div(a.N, BN) == 1 && (size_t)nk * BN * 128 <= (size_t)kWResidentBytes && row_tiles > kNumSMs;
  if (wres) {
    const int rc = launch_tc<BN, true>(a, stream);
    if (rc != kPlanDoesNotFit) return rc;
  }
  return launch_tc<BN, false>(a, stream);
}

}  // namespace
}  // namespace pdr
