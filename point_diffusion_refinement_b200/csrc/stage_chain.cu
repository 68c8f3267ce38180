// Fused stage kernel for the denoiser -- sm_100a only.
//
// One grouped stage of the network (pointnet2_modules.py:57-65,129-174 + attention.py:70-96) is a chain of 1x1
// convolutions over the SAME grouped rows, with a per-sample GroupNorm between any two of them:
//
//     X0 --W1--> [y1 | key]      y1 --GN,ReLU--> a1 --W2--> y2 ...  --Wv (+ folded residual of X0)--> V
//                                key --ReLU,GN--> k1 --W1k (+ query row)--> s1 --ReLU,GN--> s1' --Ws--> S
//     out[point] = sum_k softmax_k(S) * relu(GN(V))
//
// The per-layer engine (gemm_tc.cuh) writes every one of these tensors to HBM and reads it back: ~400 floats per grouped
// row for a 32-channel stage.  A GroupNorm needs the statistics of the WHOLE sample before its output can be used, so the
// chain cannot be evaluated in one sweep -- but it can be evaluated in L + 2 sweeps that each RECOMPUTE the chain from X0
// up to the layer whose statistics are still missing, keeping every intermediate on chip:
//
//     sweep d:  for every 128-row tile:  X0 (gathered rows, 13-60 floats per row, shared memory)
//               -> tcgen05.mma (A from shared memory) -> accumulator in TMEM
//               -> epilogue warps: tcgen05.ld, + bias, GroupNorm scale/shift + ReLU + embedding, TF32 rounding,
//                  tcgen05.st back into TMEM as the A OPERAND of the next layer's tcgen05.mma (A from TMEM)
//               -> ... -> layer d: per-tile column statistics only (sweeps 1 .. L+1), or the soft-attention pooling
//                  over the K neighbour rows straight from the two accumulators (last sweep).
//
// Nothing but X0 is read and nothing but statistics / the pooled rows is written: the tensor pipe (idle at 3-12 % in the
// per-layer engine) pays for the recomputation, HBM traffic drops by ~8x for the stages this kernel takes.
//
// The layer chain is data: a small step program (PdrChainArgs) built by the host (fused.py) -- per step a list of MMAs
// (A = X0 tile in shared memory | TMEM columns written by an earlier epilogue; B = a weight matrix resident in shared
// memory in the canonical K-major SWIZZLE_128B layout; D = TMEM columns) followed by a list of epilogue operations
// (XFORM / STATS / POOL over blocks of 32 accumulator columns).
//
// Persistent, warp-specialised, one CTA per SM, every CTA a contiguous range of row tiles.  With W = 8 (or 4 when shared
// memory is short) epilogue warps per tile group:
//   warps [0, 2W)      epilogue: two GROUPS; group g owns the tiles g, g+2, ... of the CTA's range and TMEM columns
//                      [256 g, 256 g + 256), so two tiles are in flight and one group's epilogue overlaps the other group's
//                      MMAs.  W / 4 warps share a TMEM lane quarter and take alternate 32-column blocks of a step.
//                      The per-column constants of every operation (GroupNorm scale / shift with the bias folded in,
//                      embeddings) are folded ONCE per sample into a shared-memory table by the group itself.
//   warps [2W, 2W+2)   producers: cp.async of the gathered X0 rows (table row + geometric channels) into a ring of
//                      SWIZZLE_128B tiles, completion by cp.async.mbarrier.arrive.
//   warp 2W+2          MMA issue (one lane): walks both groups' step programs in lock step.
// (First version, r02d/r02e profiles: 4 warps per group re-reading the operation descriptors from the constant bank and
//  their constants from global memory for every float4 -- 870, then 318 SASS instructions per 32-column block and an
//  issue rate of 0.2 per epilogue warp; this version: ~120 instructions per block and twice the warps.)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace pdr {
namespace {

// (registers are allocated for warps in fours: W = 8 runs 16 + 2 + 1 = 19 warps of 96 registers, 21 would not fit the file;
//  W = 4 runs 8 + 4 + 1 = 13 warps and gets 152 registers)
constexpr int kMaxWpg = 8;                                     // epilogue warps per tile group (4 or 8)
constexpr int kTileM = 128;
constexpr int kChunkBytes = kTileM * 128;                      // one 32-float K chunk of a 128-row tile
constexpr int kGroupCols = 256;                                // TMEM columns per tile group
constexpr int kMaxSlots = 8;
constexpr int kMaxStatCols = 128;                              // statistics columns per sweep
constexpr int kScratchBytes = 32 * 36 * 4;                     // per-warp transposition tile (32 rows x 36 floats)
constexpr int kConstCols = 512;                                // 32-padded columns of all operations of a sweep
constexpr int kMaxOps = PDR_CHAIN_MAX_STEPS * PDR_CHAIN_MAX_EPI;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "CH_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra CH_WAIT_DONE;\n\t"
      "bra CH_WAIT_LOOP;\n\t"
      "CH_WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)      // suspend-time hint: the warp sleeps in hardware until
      : "memory");                                                   // the phase completes instead of re-issuing the poll
}
// for waits that are expected to be long (a producer ahead of the ring): sleep between polls, the retry loop must not take
// issue slots from the epilogue warps
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
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
    asm volatile("nanosleep.u32 200;");
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem: lane = row, one column per K element] . B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (same encoding as gemm_tc.cuh)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// explicit shared-space accesses for the transposition tiles (they live in the dynamic region: through a generic pointer
// these would be LD.E / ST.E at ~3x the latency)
__device__ __forceinline__ float lds1(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one tcgen05.mma (one K step of 8) of the step program, everything precomputed but the tile group / X0 slot bases
// A value every lane of the warp holds alike, made visibly so for ptxas (shuffle from lane 0): what derives from it lives in
// uniform registers, and UTCHMMA / LDTM / STTM / SYNCS take it without an ELECT + R2UR.BROADCAST loop around each instruction
// (the step tables below are read from shared memory, which ptxas must otherwise treat as per-lane data).
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ int uni(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ uint64_t uni(uint64_t v) {
  return ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(v >> 32), 0) << 32) | __shfl_sync(0xffffffffu, (uint32_t)v, 0);
}

struct MmaEntry {
  uint64_t bdesc;        // complete shared-memory descriptor of the weight slice
  uint32_t d_col;        // accumulator column inside the group
  uint32_t a;            // A in TMEM: column inside the group; A = X0: (byte offset inside the slot) >> 4
  uint32_t idesc;
  uint32_t flags;        // bit 0: A in TMEM, bit 1: accumulate
};
constexpr int kMaxMmaEntries = 160;
struct EpiEntry {
  int kind, d_col, ncols, nblk, pro_mode, a_col, stat_col0, stat_skip, v_col, cblk, ld_rowadd;
  const float *rowadd;
};

struct ChainPlan {
  int tiles_per_sample, total_tiles;
  int nk0;          // 32-float K chunks of the gathered X0 tile
  int slots;        // ring depth
  int w_region;     // bytes reserved for the weight image (multiple of 1024)
  int wpg;          // epilogue warps per tile group (4 or 8)
};

// Folded per-column constants of one operation for one sample: three arrays of 32 floats per 32-column block,
//   XFORM GN->ReLU :  t = max(fma(acc, c0, c1), 0) + c2        c0 = sc, c1 = sh + bias sc, c2 = emb
//   XFORM ReLU->GN :  t = fma(max(acc + c0 [+ rowadd], 0), c1, c2)   c0 = bias, c1 = sc, c2 = sh + emb
//   XFORM none     :  t = acc + c0 [+ rowadd] + c2                c0 = bias, c2 = emb
//   STATS          :  y = acc + c0 [+ rowadd]
//   POOL           :  score = acc + c0;  value = max(fma(acc_v, c1, c2), 0)     c1 = v_sc, c2 = v_sh + v_bias v_sc
// (columns >= ncols: zeros, so pad columns of an XFORM come out as zeros)
__device__ __forceinline__ void fold_constants(const PdrChainEpi &op, int b, int c, float &c0, float &c1, float &c2) {
  c0 = c1 = c2 = 0.f;
  if (c >= op.ncols) return;
  const float bias = op.bias ? __ldg(op.bias + c) : 0.f;
  if (op.kind == PDR_CHAIN_XFORM) {
    const bool pro = op.pro_mode != PDR_PRO_NONE;
    const float sc = pro ? __ldg(op.sc + (size_t)b * op.ld_scsh + c) : 1.f;
    const float sh = pro ? __ldg(op.sh + (size_t)b * op.ld_scsh + c) : 0.f;
    const float e = op.emb ? __ldg(op.emb + (size_t)b * op.ld_emb + c) : 0.f;
    if (op.pro_mode == PDR_PRO_GN_RELU) { c0 = sc; c1 = fmaf(bias, sc, sh); c2 = e; }
    else if (op.pro_mode == PDR_PRO_RELU_GN) { c0 = bias; c1 = sc; c2 = sh + e; }
    else { c0 = bias; c2 = e; }
  } else if (op.kind == PDR_CHAIN_STATS) {
    c0 = bias;
  } else {
    const float vb = op.v_bias ? __ldg(op.v_bias + c) : 0.f;
    const float sc = __ldg(op.v_sc + (size_t)b * op.v_ld_scsh + c);
    const float sh = __ldg(op.v_sh + (size_t)b * op.v_ld_scsh + c);
    c0 = bias; c1 = sc; c2 = fmaf(vb, sc, sh);
  }
}

template <int WPG, int kProdWarps>
__global__ void __launch_bounds__((2 * WPG + kProdWarps + 1) * 32, 1)
stage_chain_kernel(const __grid_constant__ PdrChainArgs a, const ChainPlan plan) {
  constexpr int kProdThreads = kProdWarps * 32;
  constexpr int kRowsPerProd = kTileM * 8 / kProdThreads;        // rows a producer thread copies per chunk (16 or 8)
  constexpr int kRowStep = kTileM / kRowsPerProd;                // arow + kRowStep * r
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[kMaxSlots], bar_empty[kMaxSlots], bar_mma_done[2], bar_epi_done[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_part[2][4][kMaxStatCols][2];              // per group / lane quarter column partials
  __shared__ __align__(16) float s_const[2][kConstCols * 3];                 // per group: folded constants of one sample
  // The step program, decoded ONCE into shared memory.  Read through the kernel parameter block with a dynamic index every
  // field is an LDC with its own scoreboard wait (r02g profile: the single issuing lane of the MMA warp spent ~1 us per step
  // on them while both epilogue groups waited; the epilogue warps stalled on them between blocks).
  __shared__ MmaEntry s_mma[kMaxMmaEntries];
  __shared__ int s_mma_first[PDR_CHAIN_MAX_STEPS + 1];
  __shared__ EpiEntry s_epi[kMaxOps];

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp-uniform for ptxas (gemm_tc.cuh)
  constexpr int wpg = WPG;
  const int n_epi_warps = 2 * wpg, mma_warp = n_epi_warps + kProdWarps;
  const int nthreads = (mma_warp + 1) * 32;

  // dynamic region: [transposition tiles 2 wpg x 4608 B][weight image][ring of X0 tiles], 1 KiB aligned
  const uint32_t dyn0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_scr_u = dyn0;
  const uint32_t s_w_u = (dyn0 + (uint32_t)(n_epi_warps * kScratchBytes) + 1023u) & ~1023u;
  const uint32_t s_x0_u = s_w_u + (uint32_t)plan.w_region;
  uint8_t *s_w = smem_raw + (s_w_u - smem_u32(smem_raw));
  const uint32_t slot_bytes = (uint32_t)plan.nk0 * kChunkBytes;

  if (tid == 0) {
    for (int s = 0; s < plan.slots; ++s) { mbar_init(&bar_full[s], kProdThreads); mbar_init(&bar_empty[s], 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(&bar_mma_done[g], 1); mbar_init(&bar_epi_done[g], (uint32_t)(wpg * 32)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int off = 0;
    for (int s = 0; s < a.n_steps; ++s)
      for (int e = 0; e < PDR_CHAIN_MAX_EPI; ++e) {
        EpiEntry &t = s_epi[s * PDR_CHAIN_MAX_EPI + e];
        t.cblk = off;
        if (e < a.steps[s].n_epi) {
          const PdrChainEpi &op = a.steps[s].epi[e];
          t.kind = op.kind; t.d_col = op.d_col; t.ncols = op.ncols; t.nblk = (op.ncols + 31) >> 5; t.pro_mode = op.pro_mode;
          t.a_col = op.a_col; t.stat_col0 = op.stat_col0; t.stat_skip = op.stat_skip; t.v_col = op.v_col;
          t.ld_rowadd = op.ld_rowadd; t.rowadd = op.rowadd;
          off += t.nblk;
        }
      }
  }
  if (tid == 32) {
    // every MMA of the sweep, one entry per K step of 8
    constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileM >> 4) << 24);
    int n = 0;
    for (int s = 0; s < a.n_steps; ++s) {
      s_mma_first[s] = n;
      for (int m = 0; m < a.steps[s].n_mma; ++m) {
        const PdrChainMma &op = a.steps[s].mma[m];
        for (int kk = 0; kk < op.k; kk += 8, ++n) {
          const int kw = op.w_k0 + kk;
          MmaEntry &t = s_mma[n];
          t.bdesc = make_desc(s_w_u + (uint32_t)op.w_off + (uint32_t)(kw >> 5) * (uint32_t)op.w_rows * 128u +
                              (uint32_t)op.w_row0 * 128u) + (uint64_t)((kw & 31) >> 2);
          t.d_col = (uint32_t)op.d_col;
          t.idesc = kIdescBase | ((uint32_t)(op.n >> 3) << 17);
          t.flags = (op.a_tmem ? 1u : 0u) | ((kk > 0 || op.accumulate) ? 2u : 0u);
          const int ka = op.a_col + kk;
          t.a = op.a_tmem ? (uint32_t)ka : (((uint32_t)(ka >> 5) * kChunkBytes) >> 4) + (uint32_t)((ka & 31) >> 2);
        }
      }
    }
    s_mma_first[a.n_steps] = n;
  }
  if (warp == mma_warp) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weight image: a byte-exact copy (the host laid it out in the swizzled K-major form the tensor core reads)
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(a.w_image);
    uint4 *dst = reinterpret_cast<uint4 *>(s_w);
    for (int i = tid; i < a.w_bytes / 16; i += nthreads) dst[i] = __ldg(src + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  // this CTA's contiguous range of row tiles
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int base_n = plan.total_tiles / G, rem = plan.total_tiles % G;
  const int t_lo = cta * base_n + min(cta, rem);
  const int n_my = base_n + (cta < rem ? 1 : 0);

  if (warp >= n_epi_warps && warp < mma_warp) {
    // =============================== PRODUCERS: gathered X0 tiles ================================
    const int ptid = tid - n_epi_warps * 32;
    const int piece = ptid & 7;       // 16-byte piece of the 128-byte chunk row
    const int arow = ptid >> 3;       // rows arow + kRowStep i ((arow + kRowStep i) & 7 == arow & 7)
    const uint32_t sw_off = (uint32_t)(arow * 128 + ((piece ^ (arow & 7)) << 4));
    int slot = 0, phase = 0;
    int idx[kRowsPerProd], nidx[kRowsPerProd];
    if (n_my > 0) {
#pragma unroll
      for (int r = 0; r < kRowsPerProd; ++r) nidx[r] = __ldg(a.src_rows + (size_t)t_lo * kTileM + arow + kRowStep * r);
    }
    for (int i = 0; i < n_my; ++i) {
      const int tile = t_lo + i;
      const size_t row0 = (size_t)tile * kTileM + arow;             // rows_per_sample % 128 == 0: tiles are dense
#pragma unroll
      for (int r = 0; r < kRowsPerProd; ++r) idx[r] = nidx[r];
      if (i + 1 < n_my) {                                          // the next tile's rows: one full tile of latency hidden
#pragma unroll
        for (int r = 0; r < kRowsPerProd; ++r) nidx[r] = __ldg(a.src_rows + row0 + kTileM + kRowStep * r);
      }
      mbar_wait_sleep(&bar_empty[slot], (uint32_t)(phase ^ 1));
      const uint32_t sbase = s_x0_u + (uint32_t)slot * slot_bytes + sw_off;
      for (int kc = 0; kc < plan.nk0; ++kc) {
        const int k = kc * 32 + piece * 4;
        const uint32_t dst = sbase + (uint32_t)kc * kChunkBytes;
        if (k < a.k_split) {
#pragma unroll
          for (int r = 0; r < kRowsPerProd; ++r)
            cp_async16_ignore(dst + r * (kRowStep * 128), a.table + (size_t)max(idx[r], 0) * a.ld_table + k, idx[r] < 0);
        } else {
          const bool in = k < a.k0;
          const float *p = a.geo + row0 * (size_t)a.ld_geo + (in ? k - a.k_split : 0);
#pragma unroll
          for (int r = 0; r < kRowsPerProd; ++r)
            cp_async16_ignore(dst + r * (kRowStep * 128), in ? p + (size_t)kRowStep * r * a.ld_geo : a.geo, !in);
        }
      }
      cp_async_arrive_noinc(&bar_full[slot]);
      if (++slot == plan.slots) { slot = 0; phase ^= 1; }
    }
  } else if (warp == mma_warp) {
    // =============================== MMA ISSUER ==================================================
    // both tile groups in lock step: step s of tile 2j (group 0), step s of tile 2j + 1 (group 1), step s + 1 ...
    uint32_t ph_epi[2] = {0u, 0u};
    bool first[2] = {true, true};
    uint32_t release_mask = 0u;
    for (int s = 0; s < a.n_steps; ++s) release_mask |= a.steps[s].release_x0 ? (1u << s) : 0u;
    const int n_steps = a.n_steps;
    for (int pair = 0; pair * 2 < n_my; ++pair) {
      for (int s = 0; s < n_steps; ++s) {
        for (int g = 0; g < 2; ++g) {
          const int i = pair * 2 + g;
          if (i >= n_my) continue;
          const int slot = i % plan.slots;
          const uint32_t slot_phase = (uint32_t)((i / plan.slots) & 1);
          if (!first[g]) {
            // the epilogue of this group's previous step (or of its previous tile's last step) has drained the
            // accumulators and published the A operands
            mbar_wait(&bar_epi_done[g], ph_epi[g]);
            ph_epi[g] ^= 1u;
          }
          first[g] = false;
          if (s == 0) mbar_wait(&bar_full[slot], slot_phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          {
            // every lane reads the table entry (one broadcast LDS) and the fields are made uniform; lane 0 issues
            const uint32_t tg = tmem_base + (uint32_t)(g * kGroupCols);
            const uint64_t x0desc = make_desc(s_x0_u + (uint32_t)slot * slot_bytes);
            const int m_end = uni(s_mma_first[s + 1]);
            for (int m = uni(s_mma_first[s]); m < m_end; ++m) {
              const MmaEntry t = s_mma[m];
              const uint64_t bdesc = uni(t.bdesc);
              const uint32_t d_col = uni(t.d_col), ta = uni(t.a), idesc = uni(t.idesc), flags = uni(t.flags);
              if (lane == 0) {
                if (flags & 1u) umma_ts(tg + d_col, tg + ta, bdesc, idesc, flags >> 1);
                else umma_ss(tg + d_col, x0desc + (uint64_t)ta, bdesc, idesc, flags >> 1);
              }
            }
            if (lane == 0) {
              umma_commit(&bar_mma_done[g]);
              if (release_mask & (1u << s)) umma_commit(&bar_empty[slot]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== EPILOGUE ====================================================
    const int g = warp / wpg, wi = warp - g * wpg;
    const int quarter = wi & 3, half = wi >> 2, halves = wpg >> 2;
    const int gthreads = wpg * 32;
    const int gtid = tid - g * gthreads;                     // thread index inside the group
    const int gbar = 1 + g;                                  // named barrier of the group
    const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * kGroupCols);
    const uint32_t scr = s_scr_u + (uint32_t)warp * kScratchBytes;
    const uint32_t scr_row = scr + (uint32_t)lane * 144u;    // my row of the transposition tile (lane = row)
    const uint32_t scr_col = scr + (uint32_t)lane * 4u;      // my column (lane = column)
    const float *cg = s_const[g];
    uint32_t ph = 0u;
    int cached_b = -1;
    const int n_steps = a.n_steps;
    int n_epi_packed = 0;
    for (int s = 0; s < n_steps; ++s) n_epi_packed |= a.steps[s].n_epi << (4 * s);
    for (int i = g; i < n_my; i += 2) {
      const int tile = t_lo + i;
      const int b = tile / plan.tiles_per_sample;
      const size_t grow = (size_t)tile * kTileM + quarter * 32 + lane;     // my global grouped row
      const size_t point = grow / (size_t)a.group_k;
      if (b != cached_b) {
        // fold this sample's per-column constants of every operation of the sweep (once per sample and group)
        asm volatile("bar.sync %0, %1;" ::"r"(gbar), "r"(gthreads) : "memory");
        for (int s = 0; s < a.n_steps; ++s)
          for (int e = 0; e < a.steps[s].n_epi; ++e) {
            const PdrChainEpi &op = a.steps[s].epi[e];
            const int blk0 = s_epi[s * PDR_CHAIN_MAX_EPI + e].cblk;
            const int ncp = ((op.ncols + 31) >> 5) << 5;
            for (int c = gtid; c < ncp; c += gthreads) {
              float c0, c1, c2;
              fold_constants(op, b, c, c0, c1, c2);
              float *dst = s_const[g] + (blk0 + (c >> 5)) * 96 + (c & 31);
              dst[0] = c0; dst[32] = c1; dst[64] = c2;
            }
          }
        asm volatile("bar.sync %0, %1;" ::"r"(gbar), "r"(gthreads) : "memory");
        cached_b = b;
      }
      for (int s = 0; s < n_steps; ++s) {
        const int n_epi = (n_epi_packed >> (4 * s)) & 15;
        mbar_wait(&bar_mma_done[g], ph);
        ph ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bool wrote_tmem = false, did_stats = false;
        int unit = 0;
        for (int e = 0; e < n_epi; ++e) {
          EpiEntry opd = s_epi[s * PDR_CHAIN_MAX_EPI + e];
          opd.a_col = uni(opd.a_col); opd.stat_col0 = uni(opd.stat_col0); opd.stat_skip = uni(opd.stat_skip);
          opd.v_col = uni(opd.v_col); opd.cblk = uni(opd.cblk);
          const int kind = uni(opd.kind), d_col = uni(opd.d_col), ncols = uni(opd.ncols), pro_mode = uni(opd.pro_mode);
          const int nblk = uni(opd.nblk);
          const float *rowadd = opd.rowadd ? opd.rowadd + point * (size_t)opd.ld_rowadd : nullptr;
          const float *cop = cg + opd.cblk * 96;
          wrote_tmem |= kind == PDR_CHAIN_XFORM;
          did_stats |= kind == PDR_CHAIN_STATS;
          for (int blk = 0; blk < nblk; ++blk, ++unit) {
            if (halves > 1 && (unit & 1) != half) continue;
            const int c0 = blk * 32;
            const float *cb = cop + blk * 96;
            uint32_t v[32];
            tmem_ld32(tq + (uint32_t)(d_col + c0), v);
            // (the broadcast query row of my point is read in whole 32-column blocks: zero padded by the caller)
            const float4 *ra = reinterpret_cast<const float4 *>(rowadd + c0);
            tmem_wait_ld();
            if (kind == PDR_CHAIN_XFORM) {
              // -> TF32 A operand of a later MMA (TMEM, lane = row, one column per channel)
              if (pro_mode == PDR_PRO_GN_RELU) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 k0 = *reinterpret_cast<const float4 *>(cb + 4 * j4);
                  const float4 k1 = *reinterpret_cast<const float4 *>(cb + 32 + 4 * j4);
                  const float4 k2 = *reinterpret_cast<const float4 *>(cb + 64 + 4 * j4);
                  v[4 * j4 + 0] = __float_as_uint(to_tf32(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), k0.x, k1.x), 0.f) + k2.x));
                  v[4 * j4 + 1] = __float_as_uint(to_tf32(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), k0.y, k1.y), 0.f) + k2.y));
                  v[4 * j4 + 2] = __float_as_uint(to_tf32(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), k0.z, k1.z), 0.f) + k2.z));
                  v[4 * j4 + 3] = __float_as_uint(to_tf32(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), k0.w, k1.w), 0.f) + k2.w));
                }
              } else {
                // ReLU -> GN (or no prologue: k1 would be 1 -- not used by any stage; handled as scale 1 by the fold)
                const bool relu = pro_mode == PDR_PRO_RELU_GN;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  float4 k0 = *reinterpret_cast<const float4 *>(cb + 4 * j4);
                  const float4 k1 = *reinterpret_cast<const float4 *>(cb + 32 + 4 * j4);
                  const float4 k2 = *reinterpret_cast<const float4 *>(cb + 64 + 4 * j4);
                  if (rowadd) { const float4 r = __ldg(ra + j4); k0.x += r.x; k0.y += r.y; k0.z += r.z; k0.w += r.w; }
                  float y0 = __uint_as_float(v[4 * j4 + 0]) + k0.x, y1 = __uint_as_float(v[4 * j4 + 1]) + k0.y;
                  float y2 = __uint_as_float(v[4 * j4 + 2]) + k0.z, y3 = __uint_as_float(v[4 * j4 + 3]) + k0.w;
                  if (relu) {
                    y0 = fmaf(fmaxf(y0, 0.f), k1.x, k2.x); y1 = fmaf(fmaxf(y1, 0.f), k1.y, k2.y);
                    y2 = fmaf(fmaxf(y2, 0.f), k1.z, k2.z); y3 = fmaf(fmaxf(y3, 0.f), k1.w, k2.w);
                  } else {
                    y0 += k2.x; y1 += k2.y; y2 += k2.z; y3 += k2.w;
                  }
                  v[4 * j4 + 0] = __float_as_uint(to_tf32(y0)); v[4 * j4 + 1] = __float_as_uint(to_tf32(y1));
                  v[4 * j4 + 2] = __float_as_uint(to_tf32(y2)); v[4 * j4 + 3] = __float_as_uint(to_tf32(y3));
                }
              }
              tmem_st32(tq + (uint32_t)(opd.a_col + c0), v);
            } else {
              // STATS and POOL start alike: y = acc + (bias + rowadd), my row parked in the transposition tile
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                float4 k0 = *reinterpret_cast<const float4 *>(cb + 4 * j4);
                if (rowadd) { const float4 r = __ldg(ra + j4); k0.x += r.x; k0.y += r.y; k0.z += r.z; k0.w += r.w; }
                sts4(scr_row + 16u * j4, make_float4(__uint_as_float(v[4 * j4 + 0]) + k0.x, __uint_as_float(v[4 * j4 + 1]) + k0.y,
                                                     __uint_as_float(v[4 * j4 + 2]) + k0.z, __uint_as_float(v[4 * j4 + 3]) + k0.w));
              }
              if (kind == PDR_CHAIN_STATS) {
                // per-tile column statistics: read the tile back column-wise (lane = column).  Every consumer reads ONE
                // of the two pairs (plain or relu), stat_skip says which one is not needed.
                __syncwarp();
                float q0 = 0.f, q1 = 0.f;
                if (opd.stat_skip & 1) {
#pragma unroll
                  for (int r = 0; r < 32; ++r) { const float p = fmaxf(lds1(scr_col + 144u * r), 0.f); q0 += p; q1 = fmaf(p, p, q1); }
                } else {
#pragma unroll
                  for (int r = 0; r < 32; ++r) { const float t = lds1(scr_col + 144u * r); q0 += t; q1 = fmaf(t, t, q1); }
                }
                *reinterpret_cast<float2 *>(&s_part[g][quarter][opd.stat_col0 + c0 + lane][0]) = make_float2(q0, q1);
                __syncwarp();
              } else {
                // soft-attention pooling over the group_k neighbour rows of each point (attention.py:85-96): scores = y,
                // values = relu(GN(V accumulator + bias)).  One transposition tile, three phases:
                //   (1) lane = column takes the masked maximum and the denominator of each point and leaves
                //       exp(score - max) in place;
                //   (2) lane = row multiplies its row of weights with its values (TMEM) in place;
                //   (3) lane = column adds the group_k products of each point and divides.
                tmem_ld32(tq + (uint32_t)(opd.v_col + c0), v);        // the values, in flight during phase 1
                __syncwarp();
                const int PK = a.group_k;
                const size_t point0 = ((size_t)tile * kTileM + quarter * 32) / (size_t)PK;
                float den[4] = {1.f, 1.f, 1.f, 1.f};                  // 32 / group_k <= 4 points per warp
#pragma unroll
                for (int pi = 0; pi < 4; ++pi) {
                  const int r0 = pi * PK;
                  if (r0 < 32) {
                    int cnt = PK;
                    if (a.counts) { cnt = __ldg(a.counts + point0 + pi); cnt = cnt < 1 ? 1 : cnt; }
                    float mx = -3.0e38f;
#pragma unroll 8
                    for (int k = 0; k < PK; ++k) mx = fmaxf(mx, k < cnt ? lds1(scr_col + 144u * (r0 + k)) : -1e9f);
                    float dsum = 0.f;
#pragma unroll 8
                    for (int k = 0; k < PK; ++k) {
                      const float sk = k < cnt ? lds1(scr_col + 144u * (r0 + k)) : -1e9f;
                      const float ex = expf(sk - mx);
                      dsum += ex;
                      sts1(scr_col + 144u * (r0 + k), ex);
                    }
                    den[pi] = dsum;
                  }
                }
                tmem_wait_ld();
                __syncwarp();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 k1 = *reinterpret_cast<const float4 *>(cb + 32 + 4 * j4);
                  const float4 k2 = *reinterpret_cast<const float4 *>(cb + 64 + 4 * j4);
                  float4 w = lds4(scr_row + 16u * j4);
                  w.x *= fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), k1.x, k2.x), 0.f);
                  w.y *= fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), k1.y, k2.y), 0.f);
                  w.z *= fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), k1.z, k2.z), 0.f);
                  w.w *= fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), k1.w, k2.w), 0.f);
                  sts4(scr_row + 16u * j4, w);
                }
                __syncwarp();
                const int n = c0 + lane;                             // my output channel
#pragma unroll
                for (int pi = 0; pi < 4; ++pi) {
                  const int r0 = pi * PK;
                  if (r0 < 32) {
                    float num = 0.f;
#pragma unroll 8
                    for (int k = 0; k < PK; ++k) num += lds1(scr_col + 144u * (r0 + k));
                    if (n < ncols) {
                      const float o = num / den[pi];
                      a.out[(point0 + pi) * (size_t)a.ld_out + n] =
                          a.round_out ? __uint_as_float((__float_as_uint(o) + 0x1000u) & 0xffffe000u) : o;
                    }
                  }
                }
                __syncwarp();
              }
            }
          }
        }
        // the accumulators of this step are drained and (XFORM) the next A operands are in TMEM
        if (wrote_tmem) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&bar_epi_done[g]);
        if (did_stats) {
          // fold the four lane quarters in a fixed order and publish this tile's column partials
          asm volatile("bar.sync %0, %1;" ::"r"(gbar), "r"(gthreads) : "memory");
          for (int col = gtid; col < a.stats_n; col += gthreads) {
            const float s0 = s_part[g][0][col][0] + s_part[g][1][col][0] + s_part[g][2][col][0] + s_part[g][3][col][0];
            const float s1 = s_part[g][0][col][1] + s_part[g][1][col][1] + s_part[g][2][col][1] + s_part[g][3][col][1];
            const bool relu = (a.stats_relu_mask[col >> 5] >> (col & 31)) & 1u;
            *reinterpret_cast<float4 *>(a.stats + ((size_t)tile * a.stats_n + col) * 4) =
                relu ? make_float4(0.f, 0.f, s0, s1) : make_float4(s0, s1, 0.f, 0.f);
          }
          asm volatile("bar.sync %0, %1;" ::"r"(gbar), "r"(gthreads) : "memory");
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == mma_warp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace
}  // namespace pdr

using namespace pdr;

extern "C" int pdr_stage_chain_tile_rows(void) { return kTileM; }

extern "C" int pdr_stage_chain(const PdrChainArgs *args, void *stream_) {
  PDR_REQUIRE(args, "stage_chain: null args");
  const PdrChainArgs &a = *args;
  PDR_REQUIRE(a.table && a.src_rows && a.geo && a.w_image, "stage_chain: null pointer");
  PDR_REQUIRE(a.batch > 0 && a.rows_per_sample > 0 && a.rows_per_sample % kTileM == 0,
              "stage_chain: rows_per_sample=%d must be a positive multiple of %d", a.rows_per_sample, kTileM);
  PDR_REQUIRE(a.group_k == 8 || a.group_k == 16 || a.group_k == 32, "stage_chain: group_k=%d not in {8,16,32}", a.group_k);
  PDR_REQUIRE(a.k_split > 0 && a.k_split % 4 == 0 && a.ld_table % 4 == 0 && a.ld_table >= a.k_split && a.ld_geo % 4 == 0 &&
                  a.k0 > a.k_split && a.k0 - a.k_split <= a.ld_geo && ((uintptr_t)a.table % 16) == 0 &&
                  ((uintptr_t)a.geo % 16) == 0,
              "stage_chain: bad gathered-operand layout (k_split=%d ld_table=%d ld_geo=%d k0=%d)", a.k_split, a.ld_table,
              a.ld_geo, a.k0);
  PDR_REQUIRE(a.w_bytes > 0 && a.w_bytes % 1024 == 0 && ((uintptr_t)a.w_image % 16) == 0, "stage_chain: weight image");
  PDR_REQUIRE(a.n_steps >= 1 && a.n_steps <= PDR_CHAIN_MAX_STEPS, "stage_chain: n_steps=%d", a.n_steps);
  ChainPlan plan;
  memset(&plan, 0, sizeof(plan));
  bool any_stats = false;
  int const_blocks = 0;
  plan.nk0 = ceil_div(a.k0, 32);
  plan.tiles_per_sample = a.rows_per_sample / kTileM;
  const long long tiles = (long long)a.batch * plan.tiles_per_sample;
  PDR_REQUIRE(tiles < (1ll << 24), "stage_chain: too many tiles");
  plan.total_tiles = (int)tiles;
  plan.w_region = a.w_bytes;
  int releases = 0;
  for (int s = 0; s < a.n_steps; ++s) {
    const PdrChainStep &st = a.steps[s];
    PDR_REQUIRE(st.n_mma >= 1 && st.n_mma <= PDR_CHAIN_MAX_MMA && st.n_epi >= 1 && st.n_epi <= PDR_CHAIN_MAX_EPI,
                "stage_chain: step %d has %d MMAs / %d epilogue operations", s, st.n_mma, st.n_epi);
    releases += st.release_x0 ? 1 : 0;
    for (int m = 0; m < st.n_mma; ++m) {
      const PdrChainMma &op = st.mma[m];
      PDR_REQUIRE(op.n >= 16 && op.n <= 256 && op.n % 16 == 0 && op.k > 0 && op.k % 8 == 0 && op.d_col >= 0 &&
                      op.d_col + op.n <= kGroupCols && op.w_off >= 0 && op.w_off % 1024 == 0 && op.w_rows % 8 == 0 &&
                      op.w_row0 % 8 == 0 && op.w_row0 + op.n <= op.w_rows && op.w_k0 % 8 == 0 && op.a_col >= 0 &&
                      op.a_col % 8 == 0,
                  "stage_chain: step %d MMA %d is malformed", s, m);
      const long long w_end = (long long)op.w_off + (long long)ceil_div(op.w_k0 + op.k, 32) * op.w_rows * 128;
      PDR_REQUIRE(w_end <= a.w_bytes, "stage_chain: step %d MMA %d reads past the weight image", s, m);
      if (op.a_tmem) PDR_REQUIRE(op.a_col + op.k <= kGroupCols, "stage_chain: step %d MMA %d: A operand past the group's TMEM", s, m);
      else PDR_REQUIRE(op.a_col + op.k <= plan.nk0 * 32, "stage_chain: step %d MMA %d: A operand past the X0 tile", s, m);
    }
    for (int e = 0; e < st.n_epi; ++e) {
      const PdrChainEpi &op = st.epi[e];
      const int padded = ceil_div(op.ncols, 32) * 32;
      const_blocks += padded / 32;
      // (ncols need not be a multiple of 32: bias / sc / sh / emb / rowadd rows are read, unguarded, up to the end of their last
      //  32-column block; the caller keeps them readable and zero there, so the pad columns come out as zeros)
      PDR_REQUIRE(op.kind >= PDR_CHAIN_XFORM && op.kind <= PDR_CHAIN_POOL && op.ncols > 0 &&
                      op.d_col >= 0 && op.d_col + padded <= kGroupCols,
                  "stage_chain: step %d epilogue operation %d is malformed", s, e);
      PDR_REQUIRE((!op.bias || ((uintptr_t)op.bias % 16) == 0) &&
                      (!op.rowadd || (op.ld_rowadd % 4 == 0 && ((uintptr_t)op.rowadd % 16) == 0)),
                  "stage_chain: bias / rowadd alignment");
      if (op.kind == PDR_CHAIN_XFORM) {
        PDR_REQUIRE(op.a_col >= 0 && op.a_col + padded <= kGroupCols, "stage_chain: XFORM destination");
        PDR_REQUIRE(op.pro_mode == PDR_PRO_NONE || (op.sc && op.sh && op.ld_scsh % 4 == 0 && ((uintptr_t)op.sc % 16) == 0 &&
                                                     ((uintptr_t)op.sh % 16) == 0),
                    "stage_chain: XFORM needs aligned sc / sh");
        PDR_REQUIRE(!op.emb || (op.ld_emb % 4 == 0 && ((uintptr_t)op.emb % 16) == 0), "stage_chain: emb alignment");
        PDR_REQUIRE(!(op.rowadd && op.pro_mode == PDR_PRO_GN_RELU), "stage_chain: rowadd with a GN->ReLU prologue is not implemented");
      } else if (op.kind == PDR_CHAIN_STATS) {
        PDR_REQUIRE(a.stats && op.stat_col0 >= 0 && op.stat_col0 + padded <= kMaxStatCols && op.stat_col0 + op.ncols <= a.stats_n,
                    "stage_chain: STATS columns [%d, +%d) do not fit (stats_n=%d, cap %d)", op.stat_col0, op.ncols, a.stats_n,
                    kMaxStatCols);
        any_stats = true;
      } else {
        PDR_REQUIRE(a.out && a.ld_out >= op.ncols && op.v_sc && op.v_sh && op.v_ld_scsh % 4 == 0 && op.v_col >= 0 &&
                        op.v_col + padded <= kGroupCols && ((uintptr_t)op.v_sc % 16) == 0 && ((uintptr_t)op.v_sh % 16) == 0 &&
                        (!op.v_bias || ((uintptr_t)op.v_bias % 16) == 0),
                    "stage_chain: POOL operands");
      }
    }
  }
  int mma_entries = 0;
  for (int s = 0; s < a.n_steps; ++s)
    for (int m = 0; m < a.steps[s].n_mma; ++m) mma_entries += a.steps[s].mma[m].k / 8;
  PDR_REQUIRE(mma_entries <= kMaxMmaEntries, "stage_chain: %d MMA instructions per tile (cap %d)", mma_entries, kMaxMmaEntries);
  PDR_REQUIRE(releases == 1, "stage_chain: exactly one step must release the X0 tile (%d do)", releases);
  PDR_REQUIRE(!any_stats || (a.stats_n > 0 && a.stats_n <= kMaxStatCols), "stage_chain: stats_n=%d", a.stats_n);
  PDR_REQUIRE(const_blocks * 32 <= kConstCols, "stage_chain: %d columns of per-sample constants (cap %d)", const_blocks * 32, kConstCols);
  // shared memory: transposition tiles (one per epilogue warp) + weight image + ring of X0 tiles, next to the static part
  // (barriers, column partials, folded constants).  8 epilogue warps per tile group when that leaves room for >= 2 X0
  // tiles (two tiles are in flight), else 4.
  const size_t static_smem = 25600;          // barriers, column partials, folded constants, decoded step program (25.2 KB)
  const size_t budget = 227 * 1024 - 512 - static_smem;      // 227 KB per CTA (static + dynamic) on sm_100
  const size_t slot_bytes = (size_t)plan.nk0 * kChunkBytes;
  size_t fixed = 0;
  plan.wpg = 0;
  // 4 epilogue warps per tile group by default: 8 measured SLOWER (r02i/r02j: 2.23 ms against 1.61 ms on the 2 M-row stage)
  static const int wpg_max = getenv("PDR_CHAIN_WPG") ? atoi(getenv("PDR_CHAIN_WPG")) : 4;      // A/B: 4 or 8
  for (int wpg = (wpg_max == 8 ? kMaxWpg : 4); wpg >= 4; wpg -= 4) {
    fixed = 1024 + (size_t)(2 * wpg) * kScratchBytes + (size_t)plan.w_region;     // (4608 B tiles: 8 of them are 36 KiB)
    if (fixed + 2 * slot_bytes <= budget) { plan.wpg = wpg; break; }
  }
  if (plan.wpg == 0) {
    set_error("stage_chain: weights (%d B) + two X0 tiles (%zu B each) exceed shared memory", a.w_bytes, slot_bytes);
    return PDR_ERR_UNSUPPORTED;
  }
  int slots = (int)((budget - fixed) / slot_bytes);
  if (slots > kMaxSlots) slots = kMaxSlots;
  plan.slots = slots;
  const size_t smem = fixed + (size_t)slots * slot_bytes;
  const int sm_cap = (a.max_ctas > 0 && a.max_ctas < kNumSMs) ? a.max_ctas : kNumSMs;
  const int grid = plan.total_tiles < sm_cap ? plan.total_tiles : sm_cap;
  const int threads = plan.wpg == 8 ? (16 + 2 + 1) * 32 : (8 + 4 + 1) * 32;
  auto kern = plan.wpg == 8 ? stage_chain_kernel<8, 2> : stage_chain_kernel<4, 4>;
  static bool configured[2] = {false, false};
  if (!configured[plan.wpg == 8]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    if (e != cudaSuccess) { set_error("stage_chain: smem attr: %s", cudaGetErrorString(e)); return PDR_ERR_CUDA; }
    configured[plan.wpg == 8] = true;
  }
  kern<<<grid, threads, smem, (cudaStream_t)stream_>>>(a, plan);
  return check_launch("stage_chain_kernel");
}
