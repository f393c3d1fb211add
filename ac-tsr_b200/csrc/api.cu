// Error reporting + misc entry points of the C ABI (include/acsr.h).
#include "acsr_common.cuh"
#include "../../include/acsr.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

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

static int g_pdl = -1;       // -1: not decided yet (environment), 0 / 1: set
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("ACSR_PDL");
    g_pdl = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_pdl == 1;
}
void pdl_set(int on) { g_pdl = on ? 1 : 0; }

__global__ void rng_advance_kernel(RngState* s) {
  pdl_launch_dependents();
  pdl_wait(); s->step += 1ull; }
}  // namespace acsr

extern "C" {
int acsr_version(void) { return ACSR_ABI_VERSION; }
const char* acsr_last_error(void) { return acsr::g_err; }
int acsr_num_sms(void) { return acsr::kNumSMs; }
int acsr_set_pdl(int on) { int was = acsr::pdl_enabled() ? 1 : 0; acsr::pdl_set(on); return was; }

int acsr_rng_advance(void* rng, void* stream) {
  ACSR_REQUIRE(rng != nullptr, "acsr_rng_advance: rng is NULL");
  launch_pdl(acsr::rng_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, (acsr::RngState*)rng);
  return acsr::check_launch("acsr_rng_advance");
}
}
