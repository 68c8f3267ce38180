// Error plumbing and library identity for libpdr_b200.so.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace pdr {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int pdl_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("PDR_PDL");
    mode = e ? (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2)) : 2;
  }
  return mode;
}
bool pdl_enabled() { return pdl_mode() != 0; }

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PDR_ERR_CUDA;
  }
  return PDR_OK;
}

}  // namespace pdr

extern "C" {
int pdr_version(void) { return 100; }
const char *pdr_last_error_string(void) { return pdr::g_err; }
int pdr_built_for_sm(void) { return 100; }
}
