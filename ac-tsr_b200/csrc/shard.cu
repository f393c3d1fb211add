// Vocab-sharded item table (north_star: "the item table and logits vocab-sharded"; the reference only has the dense
// nn.Embedding of acsasrec.py:37).  Rank r stores rows [lo, hi) of the table, their gradient and their Adam moments.  The
// embedding rows a rank's own sequences need live on other ranks, so every step
//   forward : all ranks' item ids are all-gathered, each owner answers with the rows it holds (zeros elsewhere) and a
//             reduce-scatter(sum) delivers every rank the rows of its own tokens           -> acsr_shard_gather_rows
//   backward: the per-token gradient rows are all-gathered and each owner adds the ones it owns into its gradient shard
//             (row 0 = padding never receives a gradient, nn.Embedding(padding_idx=0))      -> acsr_shard_scatter_add_rows
// Both are memory-bound row movers: one warp per row, 16-byte vectors.
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

__global__ void __launch_bounds__(256) shard_gather_rows_kernel(const long long* __restrict__ ids, long long n, const float* __restrict__ shard,
                                                                long long lo, long long hi, int d4, float4* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long i = (long long)blockIdx.x * 8 + warp; i < n; i += (long long)gridDim.x * 8) {
    const long long id = ids[i];
    const bool own = id >= lo && id < hi;
    const float4* src = reinterpret_cast<const float4*>(shard) + (own ? (id - lo) * d4 : 0);
    for (int j = lane; j < d4; j += 32) out[i * d4 + j] = own ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(256) shard_scatter_add_rows_kernel(const long long* __restrict__ ids, long long n, const float* __restrict__ rows,
                                                                     long long lo, long long hi, int d4, float* __restrict__ shard_grad) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long i = (long long)blockIdx.x * 8 + warp; i < n; i += (long long)gridDim.x * 8) {
    const long long id = ids[i];
    if (id == 0 || id < lo || id >= hi) continue;          // padding row, or a row another rank owns
    const float4* src = reinterpret_cast<const float4*>(rows) + i * d4;
    float* dst = shard_grad + (id - lo) * d4 * 4;
    for (int j = lane; j < d4; j += 32) {
      const float4 v = src[j];
      red_add_v4(dst + 4 * j, v.x, v.y, v.z, v.w);
    }
  }
}

}  // namespace acsr

using namespace acsr;

extern "C" {

int acsr_shard_gather_rows(const int64_t* ids, int64_t n, const float* shard, int64_t row_lo, int64_t row_hi, int d, float* out,
                           void* stream) {
  ACSR_REQUIRE(ids && shard && out, "shard_gather_rows: NULL pointer");
  ACSR_REQUIRE(n >= 0 && d > 0 && (d & 3) == 0 && row_lo <= row_hi, "shard_gather_rows: bad sizes");
  if (n == 0) return ACSR_OK;
  long long blocks = (n + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  launch_pdl(shard_gather_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const long long*)ids, (long long)n, shard,
             (long long)row_lo, (long long)row_hi, d / 4, reinterpret_cast<float4*>(out));
  return check_launch("shard_gather_rows");
}

int acsr_shard_scatter_add_rows(const int64_t* ids, int64_t n, const float* rows, int64_t row_lo, int64_t row_hi, int d,
                                float* shard_grad, void* stream) {
  ACSR_REQUIRE(ids && rows && shard_grad, "shard_scatter_add_rows: NULL pointer");
  ACSR_REQUIRE(n >= 0 && d > 0 && (d & 3) == 0 && row_lo <= row_hi, "shard_scatter_add_rows: bad sizes");
  if (n == 0) return ACSR_OK;
  long long blocks = (n + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  launch_pdl(shard_scatter_add_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const long long*)ids, (long long)n,
             rows, (long long)row_lo, (long long)row_hi, d / 4, shard_grad);
  return check_launch("shard_scatter_add_rows");
}

}  // extern "C"
