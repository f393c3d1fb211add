// Time-interval aware key / value terms of ACTiSASRec.
//
// The reference (actisasrec.py:104-124, transformer_layers.py:1116-1134, 1085-1091) gathers two [B,L,L,d] tensors per forward --
// time_matrix_emb_K/V[t_ij], t_ij = min(|ts_i - ts_j|, time_span) -- applies element-wise dropout to them and to two position
// tensors [B,L,d], and then adds  q_i.posK_j + q_i.timeK_ij  to the raw attention scores and  sum_j prob_ij (posV_j + timeV_ij)
// to the context of every head.  Here the [B,L,L,d] tensors never exist.  With the pair embedding
//
//     E[b,i,j,c] = P[j,c] * Dp[b,j,c] + T[t[b,i,j], c] * Dt[b,i,j,c]          (P: position table, T: interval table,
//                                                                              Dp / Dt: dropout multipliers, Philox or explicit)
// three kernels cover forward and backward of both uses:
//     pair_score    s[b,h,i,j] = sum_{c in head h} x[b,i,c] * E[b,i,j,c]        (score bias; also d prob of the context term)
//     pair_context  y[b,i,c]   = sum_j p[b,h(c),i,j] * E[b,i,j,c]               (context term; also d q of the score term)
//     pair_wgrad    dP[j,c] += sum_{b,i} a[b,h,i,j] * v[b,i,c] * Dp[b,j,c] ;  dT[t,c] += sum_{t_ij = t} a_ij * v[b,i,c] * Dt[b,i,j,c]
// (row 0 of both tables is nn.Embedding's padding_idx and gets no gradient, actisasrec.py:55-58).  One CTA works on one
// sequence; x / P.Dp / p / t tiles sit in shared memory, T rows (<= 257 x d floats) are read through L1/L2, and pair_wgrad
// accumulates the interval-table gradient in shared memory when it fits and flushes it once per CTA.
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

struct PairSpec {
  const float* P;          // [L, d]
  const float* T;          // [span1, d]
  const int32_t* t;        // [B, L, L]
  const float* Dp;         // [B, L, d] explicit multipliers or NULL
  const float* Dt;         // [B, L, L, d] explicit multipliers or NULL
  const RngState* rng;     // Philox (p > 0 and no explicit multipliers)
  uint32_t stream_p, stream_t;
  float p;
  int B, L, H, dh, d, span1;
};

constexpr int kPairThreads = 256;

// dropout multipliers of 4 consecutive channels c..c+3 of element row `row` (row-major [rows, d])
__device__ __forceinline__ float4 mult4(const PairSpec& s, const float* explicit_m, uint32_t stream, long long row, int c, float inv_keep,
                                        unsigned long long seed, unsigned long long step) {
  if (explicit_m != nullptr) return *reinterpret_cast<const float4*>(explicit_m + row * s.d + c);
  if (s.p <= 0.f || s.rng == nullptr) return make_float4(1.f, 1.f, 1.f, 1.f);
  const unsigned long long e = (unsigned long long)row * s.d + c;
  const uint4 r = philox4x32(seed, step, stream, e >> 2);
  return make_float4(drop_mult(r.x, s.p, inv_keep), drop_mult(r.y, s.p, inv_keep), drop_mult(r.z, s.p, inv_keep), drop_mult(r.w, s.p, inv_keep));
}

// P[j, :] * Dp[b, j, :] of sequence b into a padded [L][d+4] tile
__device__ __forceinline__ void stage_pos(const PairSpec& s, int b, float* Ps, float inv_keep, unsigned long long seed, unsigned long long step) {
  const int dp = s.d + 4;
  for (int e = threadIdx.x; e < s.L * (s.d / 4); e += blockDim.x) {
    const int j = e / (s.d / 4), c = (e % (s.d / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(s.P + (long long)j * s.d + c);
    const float4 m = mult4(s, s.Dp, s.stream_p, (long long)b * s.L + j, c, inv_keep, seed, step);
    *reinterpret_cast<float4*>(Ps + j * dp + c) = make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
  }
}

// thread per (i, j) pair.  (A variant with d/4 lanes per pair -- coalesced reads of the interval-table row, shuffle reductions per
// head -- measured 2x SLOWER on B200, 156 vs 79 us at B=256, L=50, d=64: with one pair in flight per lane the dependent chain
// t_ij -> table row -> FMA is exposed, while here a thread has the 16 row loads of its pair in flight at once.)
__global__ void __launch_bounds__(kPairThreads) pair_score_kernel(const PairSpec s, const float* __restrict__ x, float* __restrict__ out, int causal) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, L = s.L, d = s.d, dp = d + 4;
  float* Xs = sm;                // [L][d+4]
  float* Ps = Xs + L * dp;       // [L][d+4]
  const float inv_keep = s.p > 0.f ? 1.0f / (1.0f - s.p) : 1.0f;
  unsigned long long seed = 0, step = 0;
  if (s.rng != nullptr) { seed = s.rng->seed; step = s.rng->step; }
  for (int e = threadIdx.x; e < L * (d / 4); e += blockDim.x) {
    const int i = e / (d / 4), c = (e % (d / 4)) * 4;
    *reinterpret_cast<float4*>(Xs + i * dp + c) = *reinterpret_cast<const float4*>(x + ((long long)b * L + i) * d + c);
  }
  stage_pos(s, b, Ps, inv_keep, seed, step);
  __syncthreads();
  for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
    const int i = e / L, j = e % L;
    const long long pair = ((long long)b * L + i) * L + j;
    const bool skip = causal && j > i;
    const float* Trow = s.T + (long long)(skip ? 0 : s.t[pair]) * d;
    for (int h = 0; h < s.H; ++h) {
      float acc = 0.f;
      if (!skip) {
        for (int c = h * s.dh; c < (h + 1) * s.dh; c += 4) {
          const float4 xv = *reinterpret_cast<const float4*>(Xs + i * dp + c);
          const float4 pv = *reinterpret_cast<const float4*>(Ps + j * dp + c);
          const float4 tv = __ldg(reinterpret_cast<const float4*>(Trow + c));
          const float4 m = mult4(s, s.Dt, s.stream_t, pair, c, inv_keep, seed, step);
          acc = fmaf(xv.x, fmaf(tv.x, m.x, pv.x), acc);
          acc = fmaf(xv.y, fmaf(tv.y, m.y, pv.y), acc);
          acc = fmaf(xv.z, fmaf(tv.z, m.z, pv.z), acc);
          acc = fmaf(xv.w, fmaf(tv.w, m.w, pv.w), acc);
        }
      }
      out[(((long long)b * s.H + h) * L + i) * L + j] = acc;
    }
  }
}

__global__ void __launch_bounds__(kPairThreads) pair_context_kernel(const PairSpec s, const float* __restrict__ prob, float* __restrict__ y, int accumulate) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, L = s.L, d = s.d, dp = d + 4, H = s.H;
  float* Ps = sm;                                        // [L][d+4]
  float* Pr = Ps + L * dp;                               // [H][L][L]
  int* Ts = reinterpret_cast<int*>(Pr + H * L * L);      // [L][L]
  const float inv_keep = s.p > 0.f ? 1.0f / (1.0f - s.p) : 1.0f;
  unsigned long long seed = 0, step = 0;
  if (s.rng != nullptr) { seed = s.rng->seed; step = s.rng->step; }
  stage_pos(s, b, Ps, inv_keep, seed, step);
  for (int e = threadIdx.x; e < H * L * L; e += blockDim.x) Pr[e] = prob[(long long)b * H * L * L + e];
  for (int e = threadIdx.x; e < L * L; e += blockDim.x) Ts[e] = s.t[(long long)b * L * L + e];
  __syncthreads();
  for (int e = threadIdx.x; e < L * (d / 4); e += blockDim.x) {
    const int i = e / (d / 4), c = (e % (d / 4)) * 4, h = c / s.dh;
    const float* pr = Pr + (h * L + i) * L;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < L; ++j) {
      const float pj = pr[j];
      if (pj == 0.f) continue;
      const float4 pv = *reinterpret_cast<const float4*>(Ps + j * dp + c);
      const float4 tv = __ldg(reinterpret_cast<const float4*>(s.T + (long long)Ts[i * L + j] * d + c));
      const float4 m = mult4(s, s.Dt, s.stream_t, ((long long)b * L + i) * L + j, c, inv_keep, seed, step);
      acc.x = fmaf(pj, fmaf(tv.x, m.x, pv.x), acc.x);
      acc.y = fmaf(pj, fmaf(tv.y, m.y, pv.y), acc.y);
      acc.z = fmaf(pj, fmaf(tv.z, m.z, pv.z), acc.z);
      acc.w = fmaf(pj, fmaf(tv.w, m.w, pv.w), acc.w);
    }
    float4* dst = reinterpret_cast<float4*>(y + ((long long)b * L + i) * d + c);
    if (accumulate) { const float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    *dst = acc;
  }
}

// a [B,H,L,L], v [B,L,d] -> dP [L,d] +=, dT [span1,d] +=.  acc_in_smem: the interval-table gradient is accumulated in shared
// memory and flushed once per CTA
__global__ void __launch_bounds__(kPairThreads) pair_wgrad_kernel(const PairSpec s, const float* __restrict__ a, const float* __restrict__ v,
                                                                  float* __restrict__ dP, float* __restrict__ dT, int acc_in_smem) {
  extern __shared__ __align__(16) float sm[];
  const int L = s.L, d = s.d, dp = d + 4, H = s.H;
  float* Vs = sm;                                        // [L][d+4]
  float* As = Vs + L * dp;                               // [H][L][L]
  int* Ts = reinterpret_cast<int*>(As + H * L * L);      // [L][L]
  float* accT = acc_in_smem ? reinterpret_cast<float*>(Ts + L * L) : dT;      // [span1][d]
  const float inv_keep = s.p > 0.f ? 1.0f / (1.0f - s.p) : 1.0f;
  unsigned long long seed = 0, step = 0;
  if (s.rng != nullptr) { seed = s.rng->seed; step = s.rng->step; }
  if (acc_in_smem) for (int e = threadIdx.x; e < s.span1 * d; e += blockDim.x) accT[e] = 0.f;
  for (int b = blockIdx.x; b < s.B; b += gridDim.x) {
    __syncthreads();
    for (int e = threadIdx.x; e < L * (d / 4); e += blockDim.x) {
      const int i = e / (d / 4), c = (e % (d / 4)) * 4;
      *reinterpret_cast<float4*>(Vs + i * dp + c) = *reinterpret_cast<const float4*>(v + ((long long)b * L + i) * d + c);
    }
    for (int e = threadIdx.x; e < H * L * L; e += blockDim.x) As[e] = a[(long long)b * H * L * L + e];
    for (int e = threadIdx.x; e < L * L; e += blockDim.x) Ts[e] = s.t[(long long)b * L * L + e];
    __syncthreads();
    for (int e = threadIdx.x; e < L * (d / 4); e += blockDim.x) {
      const int j = e / (d / 4), c = (e % (d / 4)) * 4, h = c / s.dh;
      float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = 0; i < L; ++i) {
        const float av = As[(h * L + i) * L + j];
        if (av == 0.f) continue;
        const float4 vv = *reinterpret_cast<const float4*>(Vs + i * dp + c);
        const float4 pr = make_float4(av * vv.x, av * vv.y, av * vv.z, av * vv.w);
        sp.x += pr.x; sp.y += pr.y; sp.z += pr.z; sp.w += pr.w;
        const int t = Ts[i * L + j];
        if (t == 0) continue;                            // padding_idx row of the interval table
        const float4 m = mult4(s, s.Dt, s.stream_t, ((long long)b * L + i) * L + j, c, inv_keep, seed, step);
        float* dst = accT + (long long)t * d + c;
        if (m.x != 0.f) atomicAdd(dst + 0, pr.x * m.x);
        if (m.y != 0.f) atomicAdd(dst + 1, pr.y * m.y);
        if (m.z != 0.f) atomicAdd(dst + 2, pr.z * m.z);
        if (m.w != 0.f) atomicAdd(dst + 3, pr.w * m.w);
      }
      if (j != 0) {                                      // position 0 is the padding_idx row of the position tables
        const float4 m = mult4(s, s.Dp, s.stream_p, (long long)b * L + j, c, inv_keep, seed, step);
        float* dst = dP + j * d + c;                     // L*d reductions per sequence: straight to global
        if (sp.x * m.x != 0.f) atomicAdd(dst + 0, sp.x * m.x);
        if (sp.y * m.y != 0.f) atomicAdd(dst + 1, sp.y * m.y);
        if (sp.z * m.z != 0.f) atomicAdd(dst + 2, sp.z * m.z);
        if (sp.w * m.w != 0.f) atomicAdd(dst + 3, sp.w * m.w);
      }
    }
  }
  __syncthreads();
  if (acc_in_smem) for (int e = threadIdx.x; e < s.span1 * d; e += blockDim.x) if (accT[e] != 0.f) atomicAdd(dT + e, accT[e]);
}

// actisasrec.py:146-155: |ts_i - ts_j| (fp32, like the reference's float field) clipped to time_span, truncated to int
__global__ void __launch_bounds__(256) time_matrix_kernel(const float* __restrict__ ts, int B, int L, int span, int32_t* __restrict__ out) {
  const long long n = (long long)B * L * L;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % L), i = (int)((e / L) % L);
    const long long b = e / ((long long)L * L);
    float df = fabsf(ts[b * L + i] - ts[b * L + j]);
    if (df > (float)span) df = (float)span;
    out[e] = (int32_t)df;
  }
}

static int pair_validate(const PairSpec& s, const char* who) {
  ACSR_REQUIRE(s.P && s.T && s.t, "%s: NULL table / interval matrix", who);
  ACSR_REQUIRE(s.B > 0 && s.L > 0 && s.L <= 64 && s.H > 0 && s.dh > 0 && (s.dh & 3) == 0 && s.span1 > 0, "%s: bad sizes (L <= 64, head size a multiple of 4)", who);
  ACSR_REQUIRE(s.p >= 0.f && s.p < 1.f, "%s: dropout p=%f", who, s.p);
  ACSR_REQUIRE(!(s.p > 0.f && s.rng == nullptr && (s.Dp == nullptr || s.Dt == nullptr)), "%s: p>0 needs explicit multipliers or rng", who);
  return ACSR_OK;
}

template <typename K>
static int pair_prep(K kernel, size_t smem, const char* who) {
  if (smem > 227 * 1024) { set_error("%s: needs %zu bytes of shared memory (> 227 KB)", who, smem); return ACSR_ERR_UNSUPPORTED; }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) { set_error("%s: smem %zu: %s", who, smem, cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  return ACSR_OK;
}

static PairSpec make_spec(const float* P, const float* T, const int32_t* tmat, int B, int L, int H, int dh, int span1, float p,
                          const float* Dp, const float* Dt, const void* rng, uint32_t stream_p, uint32_t stream_t) {
  PairSpec s;
  s.P = P; s.T = T; s.t = tmat; s.Dp = Dp; s.Dt = Dt; s.rng = (const RngState*)rng; s.stream_p = stream_p; s.stream_t = stream_t;
  s.p = p; s.B = B; s.L = L; s.H = H; s.dh = dh; s.d = H * dh; s.span1 = span1;
  return s;
}

}  // namespace acsr

using namespace acsr;

extern "C" {

int acsr_time_matrix(const float* time_seq, int B, int L, int time_span, int32_t* tmat, void* stream) {
  ACSR_REQUIRE(time_seq && tmat && B > 0 && L > 0 && time_span >= 0, "time_matrix: bad arguments");
  const long long n = (long long)B * L * L;
  time_matrix_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(time_seq, B, L, time_span, tmat);
  return check_launch("time_matrix");
}

int acsr_pair_score(const float* x, const float* P, const float* T, const int32_t* tmat, int B, int L, int H, int dh, int span1,
                    float p, const float* Dp, const float* Dt, const void* rng, uint32_t stream_p, uint32_t stream_t, int causal,
                    float* s_out, void* stream) {
  const PairSpec s = make_spec(P, T, tmat, B, L, H, dh, span1, p, Dp, Dt, rng, stream_p, stream_t);
  int rc = pair_validate(s, "pair_score");
  if (rc) return rc;
  ACSR_REQUIRE(x && s_out, "pair_score: NULL pointer");
  const size_t smem = (size_t)2 * L * (s.d + 4) * sizeof(float);
  rc = pair_prep(pair_score_kernel, smem, "pair_score");
  if (rc) return rc;
  pair_score_kernel<<<B, kPairThreads, smem, (cudaStream_t)stream>>>(s, x, s_out, causal);
  return check_launch("pair_score");
}

int acsr_pair_context(const float* prob, const float* P, const float* T, const int32_t* tmat, int B, int L, int H, int dh, int span1,
                      float p, const float* Dp, const float* Dt, const void* rng, uint32_t stream_p, uint32_t stream_t,
                      int accumulate, float* y, void* stream) {
  const PairSpec s = make_spec(P, T, tmat, B, L, H, dh, span1, p, Dp, Dt, rng, stream_p, stream_t);
  int rc = pair_validate(s, "pair_context");
  if (rc) return rc;
  ACSR_REQUIRE(prob && y, "pair_context: NULL pointer");
  const size_t smem = ((size_t)L * (s.d + 4) + (size_t)H * L * L + (size_t)L * L) * sizeof(float);
  rc = pair_prep(pair_context_kernel, smem, "pair_context");
  if (rc) return rc;
  pair_context_kernel<<<B, kPairThreads, smem, (cudaStream_t)stream>>>(s, prob, y, accumulate);
  return check_launch("pair_context");
}

int acsr_pair_wgrad(const float* a, const float* v, const int32_t* tmat, int B, int L, int H, int dh, int span1,
                    float p, const float* Dp, const float* Dt, const void* rng, uint32_t stream_p, uint32_t stream_t,
                    float* dP, float* dT, void* stream) {
  PairSpec s = make_spec(dP, dT, tmat, B, L, H, dh, span1, p, Dp, Dt, rng, stream_p, stream_t);     // tables are not read here
  int rc = pair_validate(s, "pair_wgrad");
  if (rc) return rc;
  ACSR_REQUIRE(a && v && dP && dT, "pair_wgrad: NULL pointer");
  const size_t base = ((size_t)L * (s.d + 4) + (size_t)H * L * L + (size_t)L * L) * sizeof(float);
  const size_t with_t = base + (size_t)span1 * s.d * sizeof(float);
  const int in_smem = with_t <= 200 * 1024;        // (<= 112 KB, the C2 shape: two CTAs share an SM)
  const size_t smem = in_smem ? with_t : base;
  rc = pair_prep(pair_wgrad_kernel, smem, "pair_wgrad");
  if (rc) return rc;
  pair_wgrad_kernel<<<std::min(B, 4 * 148), kPairThreads, smem, (cudaStream_t)stream>>>(s, a, v, dP, dT, in_smem);
  return check_launch("pair_wgrad");
}

}  // extern "C"
