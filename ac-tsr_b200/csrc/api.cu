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

struct StaticRange { const char* lo; const char* hi; };
static StaticRange g_static[64];
static int g_n_static = 0;
bool is_static_memory(const void* p) {
  const char* c = (const char*)p;
  for (int i = 0; i < g_n_static; ++i)
    if (c >= g_static[i].lo && c < g_static[i].hi) return true;
  return false;
}

// caller-owned scratch memory of the current device (acsr_set_workspace): the library allocates nothing itself
static void* g_ws_ptr[64] = {};
static size_t g_ws_bytes[64] = {};
void* workspace_ptr(size_t* bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { *bytes = 0; return nullptr; }
  *bytes = g_ws_bytes[dev];
  return g_ws_ptr[dev];
}

__global__ void rng_advance_kernel(RngState* s) {
  pdl_launch_dependents();
  pdl_wait(); s->step += 1ull; }
}  // namespace acsr

extern "C" {
int acsr_version(void) { return ACSR_ABI_VERSION; }
const char* acsr_last_error(void) { return acsr::g_err; }
int acsr_num_sms(void) { return acsr::kNumSMs; }
int acsr_set_pdl(int on) { int was = acsr::pdl_enabled() ? 1 : 0; acsr::pdl_set(on); return was; }

int acsr_register_static(const void* ptr, int64_t bytes) {
  if (ptr == nullptr) { acsr::g_n_static = 0; return ACSR_OK; }           // NULL clears the registry
  ACSR_REQUIRE(bytes > 0, "acsr_register_static: bad size");
  for (int i = 0; i < acsr::g_n_static; ++i)
    if (acsr::g_static[i].lo == (const char*)ptr) { acsr::g_static[i].hi = (const char*)ptr + bytes; return ACSR_OK; }
  ACSR_REQUIRE(acsr::g_n_static < 64, "acsr_register_static: registry full");
  acsr::g_static[acsr::g_n_static].lo = (const char*)ptr;
  acsr::g_static[acsr::g_n_static].hi = (const char*)ptr + bytes;
  ++acsr::g_n_static;
  return ACSR_OK;
}

int acsr_set_workspace(void* ptr, int64_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { acsr::set_error("acsr_set_workspace: %s", cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  ACSR_REQUIRE(dev >= 0 && dev < 64 && bytes >= 0 && (ptr != nullptr || bytes == 0), "acsr_set_workspace: bad arguments");
  acsr::g_ws_ptr[dev] = ptr;
  acsr::g_ws_bytes[dev] = (size_t)bytes;
  return ACSR_OK;
}
int64_t acsr_attn_workspace_bytes(int L, int H, int n_streams) {
  if (L <= 64) return 0;
  return (int64_t)((size_t)(2 * n_streams + 2) * L * L + (size_t)2 * n_streams * L) * 4 * H;
}

int acsr_rng_advance(void* rng, void* stream) {
  ACSR_REQUIRE(rng != nullptr, "acsr_rng_advance: rng is NULL");
  launch_pdl(acsr::rng_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, (acsr::RngState*)rng);
  return acsr::check_launch("acsr_rng_advance");
}
}
