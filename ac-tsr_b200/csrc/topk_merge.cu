// Merge of the per-chunk (and per-vocab-shard) partial top-k lists: one CTA per user row loads the
// n_parts*k candidates into shared memory, bitonic-sorts them (descending score, ascending slot on
// ties) and writes the final top-k item ids, scores and the hit flags the collector needs
// (evaluator/collector.py:147-153: pos_matrix gather -> [B,k] flags + pos_len).
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

constexpr int kMergeThreads = 256;
constexpr int kMergeMaxCand = 16384;

__global__ void __launch_bounds__(kMergeThreads)
topk_merge_kernel(const float* __restrict__ pval, const long long* __restrict__ pidx, int n_cand, int P, int k,
                  const long long* __restrict__ positive, float* __restrict__ oval, long long* __restrict__ oidx,
                  int* __restrict__ rec) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float smem_m[];
  float* sv = smem_m;                                   // [P]
  int* ss = reinterpret_cast<int*>(smem_m + P);         // [P] candidate slot
  const int row = blockIdx.x;
  const float* rv = pval + (long long)row * n_cand;
  const long long* ri = pidx + (long long)row * n_cand;
  for (int i = threadIdx.x; i < P; i += kMergeThreads) {
    float v = -INFINITY;
    if (i < n_cand && ri[i] >= 0) v = rv[i];
    sv[i] = v;
    ss[i] = i < n_cand ? i : 0x7fffffff;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < P / 2; t += kMergeThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);           // first half of each block sorted descending
        const float a = sv[lo], b = sv[hi];
        const int sa = ss[lo], sb = ss[hi];
        const bool a_before_b = (a > b) || (a == b && sa < sb);
        if (a_before_b != desc) { sv[lo] = b; sv[hi] = a; ss[lo] = sb; ss[hi] = sa; }
      }
      __syncthreads();
    }
  }
  const long long pos_item = positive ? positive[row] : -1;
  for (int j = threadIdx.x; j < k; j += kMergeThreads) {
    const int slot = ss[j];
    const bool ok = slot < n_cand && ri[slot] >= 0;
    const long long id = ok ? ri[slot] : -1;
    oval[(long long)row * k + j] = ok ? sv[j] : -INFINITY;
    oidx[(long long)row * k + j] = id;
    if (rec) rec[(long long)row * (k + 1) + j] = (id == pos_item && id >= 0) ? 1 : 0;
  }
  if (rec && threadIdx.x == 0) rec[(long long)row * (k + 1) + k] = 1;   // pos_len: one held-out item per user
}

}  // namespace acsr

using namespace acsr;

extern "C" int acsr_topk_merge(const float* partial_val, const int64_t* partial_idx, int M, int n_parts, int k,
                               const int64_t* positive, float* topk_val, int64_t* topk_idx, int32_t* rec_topk, void* stream) {
  ACSR_REQUIRE(partial_val && partial_idx && topk_val && topk_idx, "topk_merge: NULL pointer");
  ACSR_REQUIRE(M > 0 && n_parts > 0 && k > 0, "topk_merge: bad sizes");
  const int n_cand = n_parts * k;
  if (n_cand > kMergeMaxCand) { set_error("topk_merge: %d candidates per row exceed %d", n_cand, kMergeMaxCand); return ACSR_ERR_UNSUPPORTED; }
  ACSR_REQUIRE(n_cand >= k, "topk_merge: fewer candidates than k");
  int P = 2;
  while (P < n_cand) P <<= 1;
  const size_t smem = (size_t)P * 8;
  cudaError_t e = cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("topk_merge: smem attr: %s", cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  launch_pdl(topk_merge_kernel, dim3(M), dim3(kMergeThreads), smem, (cudaStream_t)stream, partial_val, (const long long*)partial_idx, n_cand, P, k,
                                                                      (const long long*)positive, topk_val, (long long*)topk_idx,
                                                                      rec_topk);
  return check_launch("topk_merge");
}
