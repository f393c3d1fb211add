// Error reporting + misc entry points of the C ABI (include/acsr.h).
#include "acsr_common.cuh"
#include "../../include/acsr.h"
#include <cstdarg>
#include <cstdio>

namespace acsr {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ACSR_ERR_CUDA;
  }
  return ACSR_OK;
}

__global__ void rng_advance_kernel(RngState* s) { s->step += 1ull; }
}  // namespace acsr

extern "C" {
int acsr_version(void) { return ACSR_ABI_VERSION; }
const char* acsr_last_error(void) { return acsr::g_err; }
int acsr_num_sms(void) { return acsr::kNumSMs; }

int acsr_rng_advance(void* rng, void* stream) {
  ACSR_REQUIRE(rng != nullptr, "acsr_rng_advance: rng is NULL");
  acsr::rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((acsr::RngState*)rng);
  return acsr::check_launch("acsr_rng_advance");
}
}
