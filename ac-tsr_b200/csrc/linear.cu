// Weight / bias gradient of a token-parallel linear layer:  dW[N,K] += dY[T,N]^T . X[T,K],  db[N] += sum_t dY.
// T (tokens) is the long axis (12,800 per step) and the output is tiny (64x64 ... 256x64), the
// shape library GEMMs handle worst (one split-K CTA per output tile).  Here the token axis is split
// over the whole GPU: every CTA stages 16-token slabs of dY and X in shared memory, accumulates a
// register tile per thread in exact fp32 FMAs and finishes with one atomic add per output element,
// so the gradient lands directly in the caller's (flat) gradient buffer.
// Also used for d_out = G.E of the CE backward (reduction over the catalogue axis).
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

constexpr int kWgThreads = 256;
constexpr int kWgTT = 16;          // tokens per shared-memory slab

template <int TN, int TK>
__global__ void __launch_bounds__(kWgThreads)
linear_wgrad_kernel(const float* __restrict__ dY, const float* __restrict__ X, int T, int N, int K, int tok_per_cta,
                    int tiles_k, float* __restrict__ dW, float* __restrict__ db) {
  constexpr int TILE_N = 16 * TN, TILE_K = 16 * TK;
  __shared__ __align__(16) float sY[kWgTT][TILE_N];
  __shared__ __align__(16) float sX[kWgTT][TILE_K];
  const int tile_n = blockIdx.y / tiles_k, tile_k = blockIdx.y % tiles_k;
  const int n0 = tile_n * TILE_N, k0 = tile_k * TILE_K;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // tx -> k sub-tile, ty -> n sub-tile
  const int t_begin = blockIdx.x * tok_per_cta;
  const int t_end = min(T, t_begin + tok_per_cta);
  float acc[TN][TK];
  float accb[TN];
#pragma unroll
  for (int a = 0; a < TN; ++a) {
    accb[a] = 0.f;
#pragma unroll
    for (int c = 0; c < TK; ++c) acc[a][c] = 0.f;
  }
  for (int t0 = t_begin; t0 < t_end; t0 += kWgTT) {
    // stage (zero-fill outside the matrix so the inner loop is branch-free)
    for (int e = threadIdx.x; e < kWgTT * TILE_N; e += kWgThreads) {
      const int tt = e / TILE_N, n = e % TILE_N;
      const int t = t0 + tt;
      sY[tt][n] = (t < t_end && n0 + n < N) ? dY[(long long)t * N + n0 + n] : 0.f;
    }
    for (int e = threadIdx.x; e < kWgTT * TILE_K; e += kWgThreads) {
      const int tt = e / TILE_K, k = e % TILE_K;
      const int t = t0 + tt;
      sX[tt][k] = (t < t_end && k0 + k < K) ? X[(long long)t * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < kWgTT; ++tt) {
      float y[TN], x[TK];
#pragma unroll
      for (int a = 0; a < TN; a += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&sY[tt][ty * TN + a]);
        y[a] = v.x; y[a + 1] = v.y; y[a + 2] = v.z; y[a + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < TK; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&sX[tt][tx * TK + c]);
        x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
      }
#pragma unroll
      for (int a = 0; a < TN; ++a) {
        accb[a] += y[a];
#pragma unroll
        for (int c = 0; c < TK; ++c) acc[a][c] = fmaf(y[a], x[c], acc[a][c]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < TN; ++a) {
    const int n = n0 + ty * TN + a;
    if (n >= N) continue;
#pragma unroll
    for (int c = 0; c < TK; ++c) {
      const int k = k0 + tx * TK + c;
      if (k < K) atomicAdd(dW + (long long)n * K + k, acc[a][c]);
    }
    if (db != nullptr && tx == 0 && tile_k == 0) atomicAdd(db + n, accb[a]);
  }
}

}  // namespace acsr

using namespace acsr;

extern "C" int acsr_linear_wgrad(const float* dY, const float* X, int T, int N, int K, float* dW, float* db, void* stream) {
  ACSR_REQUIRE(dY && X && dW, "linear_wgrad: NULL pointer");
  ACSR_REQUIRE(T >= 0 && N > 0 && K > 0, "linear_wgrad: bad sizes");
  if (T == 0) return ACSR_OK;
  const bool big = (long long)N * K > 4096;
  const int tile = big ? 128 : 64;
  const int tiles_n = (N + tile - 1) / tile, tiles_k = (K + tile - 1) / tile;
  // split the token axis so that ~2 CTAs per SM are busy, at least 64 tokens each
  int splits = (2 * kNumSMs) / (tiles_n * tiles_k);
  if (splits < 1) splits = 1;
  int tok = (T + splits - 1) / splits;
  if (tok < 64) tok = 64;
  tok = (tok + kWgTT - 1) / kWgTT * kWgTT;
  dim3 grid((T + tok - 1) / tok, tiles_n * tiles_k);
  if (big)
    linear_wgrad_kernel<8, 8><<<grid, kWgThreads, 0, (cudaStream_t)stream>>>(dY, X, T, N, K, tok, tiles_k, dW, db);
  else
    linear_wgrad_kernel<4, 4><<<grid, kWgThreads, 0, (cudaStream_t)stream>>>(dY, X, T, N, K, tok, tiles_k, dW, db);
  return check_launch("linear_wgrad");
}
