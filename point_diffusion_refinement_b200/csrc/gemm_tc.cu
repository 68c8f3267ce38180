// Column-tile width of the tcgen05 GEMM (kernel and launcher: gemm_tc.cuh; one translation unit per width).
#include <stdlib.h>

#include "common.cuh"

namespace pdr {

int gemm_tf32_bn32(const PdrGemmArgs &a, cudaStream_t stream);
int gemm_tf32_bn64(const PdrGemmArgs &a, cudaStream_t stream);
int gemm_tf32_bn128(const PdrGemmArgs &a, cudaStream_t stream);
int gemm_tf32_bn256(const PdrGemmArgs &a, cudaStream_t stream);
constexpr int kPlanDoesNotFit = 12345;   // same value as in gemm_tc.cuh: the planner could not fit this width

// PDR_GEMM_SPREAD=0: column-tile width from N alone (A/B)
static bool spread_small() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("PDR_GEMM_SPREAD"); on = !(e && e[0] == '0'); }
  return on == 1;
}

int launch_gemm_tf32(const PdrGemmArgs &a, cudaStream_t stream) {
  // widest tile that covers N ...
  int bn = a.N > 128 ? 256 : (a.N > 64 ? 128 : (a.N > 32 ? 64 : 32));
  // ... unless the GEMM is so short (deep levels: 4 .. 64 row tiles; the t-embedding GEMM: one) that it would occupy less than
  // half of the SMs: narrower tiles put more CTAs on it, and each streams a thinner weight slice through a deeper ring.
  // A is re-read once per column tile, from L2 (these A's are a few MB).
  if (spread_small()) {
    const long long row_tiles = (long long)a.batch * ((a.rows_per_sample + 127) / 128);
    while (bn > 32 && row_tiles * ((a.N + bn - 1) / bn) * 2 <= 148) bn >>= 1;
  }
  if (bn == 256) {
    const int rc = gemm_tf32_bn256(a, stream);
    if (rc != kPlanDoesNotFit) return rc;
    bn = 128;
  }
  if (bn == 128) return gemm_tf32_bn128(a, stream);
  if (bn == 64) return gemm_tf32_bn64(a, stream);
  return gemm_tf32_bn32(a, stream);
}

}  // namespace pdr
