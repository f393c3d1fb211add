// The last encoder layer after its attention, on the rows that feed the losses only (layers.py:676-684 dense +
// dropout + residual + LayerNorm, layers.py:790-798 FeedForward, abstract_recommender.py:130-134 gather_indexes).  Only position
// len-1 of every sequence reaches the loss, so this part of the layer runs on 2B rows ([calibrated ; attacked]) instead of 2T.
// At B = 256 that is 512 rows: six launches of tensor-core GEMMs / row-wise kernels, each a few hundred CTAs-worth of fixed cost
// for ~20 MFLOP.  Here the whole chain is ONE launch in each direction: a CTA owns four rows, keeps them in shared memory between
// the stages, and reads the weights (144 KB at d = 64, inner = 256, L2-resident) with one coalesced pass per stage; the math is
// plain fp32 FMA (exact fp32 accumulation -- no TF32 split needed at this size).  Hidden size 64, inner size <= 256.
//   forward : gather(len-1) -> hz = ctx.Wo^T -> h = LN(drop(hz + bo) + x) -> z1 = h.W1^T -> a1 = act(z1 + b1) -> z2 = a1.W2^T
//             -> out = LN(drop(z2 + b2) + h); everything the backward needs is written once (same tensors as the unfused path)
//   backward: LN / dropout / activation backward and the three input-gradient products in the reverse order, parameter-gradient
//             column sums reduced per CTA then one atomic per column, d_ctx / d_x scattered straight to position len-1 of the
//             token-major buffers the attention backward and the layer below read (gather_indexes backward)
// Dropout masks use the same Philox counters as bdrl_{fwd,bwd}_kernel / linear_tok's fused epilogue (element e -> call e>>2, word e&3).
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

constexpr int kTailRows = 4;
constexpr int kTailThreads = 256;
constexpr int kTD = 64;
constexpr int kTailMaxI = 256;

struct TailParams {
  // geometry
  int B, L, I, act, n_groups;          // C = n_groups * B rows; group g reads ctx[g]
  const long long* item_len;
  // forward inputs
  const float* ctx[2]; const float* x;
  // parameters
  const float *Wo, *bo, *lnA_w, *lnA_b, *W1, *b1, *W2, *b2, *lnF_w, *lnF_b;
  float epsA, epsF, p_drop;
  const float *mask_a, *mask_f; const RngState* rng; uint32_t stream_a, stream_f;
  // saved activations (written by the forward, read by the backward)
  float *c_ctx, *c_x, *hz, *st_a, *h, *z1, *a1, *z2, *st_f, *out;
  // backward
  const float* d_out;
  float *d_z2, *d_z1, *d_hz;
  float* d_x[2]; float* d_ctx[2];
  float *g_bo, *g_lnA_w, *g_lnA_b, *g_b1, *g_b2, *g_lnF_w, *g_lnF_b;
};

__device__ __forceinline__ float tail_drop(const TailParams& p, const float* mask, uint32_t stream, long long row, int col) {
  if (mask != nullptr) return mask[row * kTD + col];
  if (p.p_drop <= 0.f || p.rng == nullptr) return 1.0f;
  const unsigned long long e = (unsigned long long)row * kTD + col;
  const uint4 r = philox4x32(p.rng->seed, p.rng->step, stream, e >> 2);
  const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
  return drop_mult(bits[e & 3], p.p_drop, 1.0f / (1.0f - p.p_drop));
}

// dot of a shared-memory row (broadcast reads) with 64 register-resident weights
__device__ __forceinline__ float dot64(const float* __restrict__ srow, const float* w) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 64; k += 4) {
    const float4 s = *reinterpret_cast<const float4*>(srow + k);
    acc = fmaf(s.x, w[k], acc); acc = fmaf(s.y, w[k + 1], acc); acc = fmaf(s.z, w[k + 2], acc); acc = fmaf(s.w, w[k + 3], acc);
  }
  return acc;
}

__device__ __forceinline__ void load_w64(const float* __restrict__ g, float* w) {
#pragma unroll
  for (int k = 0; k < 64; k += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(g + k));
    w[k] = t.x; w[k + 1] = t.y; w[k + 2] = t.z; w[k + 3] = t.w;
  }
}

// LayerNorm of s_x[r][0..64) for the thread's own row (every thread of a row walks the same 64 values in the same order)
__device__ __forceinline__ void row_stats(const float* __restrict__ srow, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) s += srow[c];
  mean = s * (1.0f / 64);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) { const float t = srow[c] - mean; var = fmaf(t, t, var); }
  rstd = 1.0f / sqrtf(var * (1.0f / 64) + eps);
}

__global__ void __launch_bounds__(kTailThreads) tail_fwd_kernel(const TailParams p) {
  __shared__ __align__(16) float s_ctx[kTailRows][kTD], s_res[kTailRows][kTD], s_x[kTailRows][kTD], s_h[kTailRows][kTD];
  __shared__ __align__(16) float s_a1[kTailRows][kTailMaxI], s_part[4][kTailRows][kTD];
  const int tid = threadIdx.x;
  const int C = p.n_groups * p.B;
  const int row0 = blockIdx.x * kTailRows;
  const int n = tid & 63, r = tid >> 6;                 // (column, row) role of the 64-wide stages
  const int rg = row0 + r;
  const bool ok = rg < C;
  // weights of the first stage are parameters: in flight before the previous kernel has finished
  float w[64];
  load_w64(p.Wo + n * kTD, w);
  pdl_launch_dependents();
  pdl_wait();
  // ---- gather position len-1 (gather_indexes) ----
  if (tid < kTailRows * 16) {
    const int gr = tid >> 4, q = tid & 15;
    const int grow = row0 + gr;
    float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f), x4 = c4;
    if (grow < C) {
      const int g = grow / p.B, b = grow - g * p.B;
      const long long t = (long long)b * p.L + (p.item_len[b] - 1);
      c4 = *reinterpret_cast<const float4*>(p.ctx[g] + t * kTD + q * 4);
      x4 = *reinterpret_cast<const float4*>(p.x + t * kTD + q * 4);
      *reinterpret_cast<float4*>(p.c_ctx + (long long)grow * kTD + q * 4) = c4;
      if (g == 0) *reinterpret_cast<float4*>(p.c_x + (long long)b * kTD + q * 4) = x4;
    }
    *reinterpret_cast<float4*>(&s_ctx[gr][q * 4]) = c4;
    *reinterpret_cast<float4*>(&s_res[gr][q * 4]) = x4;
  }
  __syncthreads();
  // ---- attention output projection + dropout + residual + LayerNorm ----
  {
    const float acc = dot64(s_ctx[r], w);
    if (ok) p.hz[(long long)rg * kTD + n] = acc;
    const float m = ok ? tail_drop(p, p.mask_a, p.stream_a, rg, n) : 1.0f;
    s_x[r][n] = (acc + p.bo[n]) * m + s_res[r][n];
  }
  __syncthreads();
  {
    float mean, rstd;
    row_stats(s_x[r], p.epsA, mean, rstd);
    const float hv = (s_x[r][n] - mean) * rstd * p.lnA_w[n] + p.lnA_b[n];
    s_h[r][n] = hv;
    if (ok) {
      p.h[(long long)rg * kTD + n] = hv;
      if (n == 0) { p.st_a[2 * rg] = mean; p.st_a[2 * rg + 1] = rstd; }
    }
  }
  __syncthreads();
  // ---- feed-forward 1: thread = inner feature ----
  if (tid < p.I) {
    load_w64(p.W1 + (long long)tid * kTD, w);
    const float b1 = p.b1[tid];
#pragma unroll
    for (int rr = 0; rr < kTailRows; ++rr) {
      const float acc = dot64(s_h[rr], w);
      const float a = act_fwd(p.act, acc + b1);
      s_a1[rr][tid] = a;
      if (row0 + rr < C) {
        p.z1[(long long)(row0 + rr) * p.I + tid] = acc;
        p.a1[(long long)(row0 + rr) * p.I + tid] = a;
      }
    }
  }
  __syncthreads();
  // ---- feed-forward 2: thread = (output column, quarter of the inner axis); the quarters meet in shared memory ----
  {
    const int I4 = p.I >> 2, kq = r;                     // I is a multiple of 16: quarters of I4 <= 64 (multiple of 4) features
    float acc[kTailRows] = {0.f, 0.f, 0.f, 0.f};
    const float* wrow = p.W2 + (long long)n * p.I + kq * I4;
    for (int k = 0; k < I4; k += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
      for (int rr = 0; rr < kTailRows; ++rr) {
        const float4 a4 = *reinterpret_cast<const float4*>(&s_a1[rr][kq * I4 + k]);
        acc[rr] = fmaf(a4.x, w4.x, acc[rr]); acc[rr] = fmaf(a4.y, w4.y, acc[rr]);
        acc[rr] = fmaf(a4.z, w4.z, acc[rr]); acc[rr] = fmaf(a4.w, w4.w, acc[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < kTailRows; ++rr) s_part[kq][rr][n] = acc[rr];
  }
  __syncthreads();
  {
    const float z = (s_part[0][r][n] + s_part[1][r][n]) + (s_part[2][r][n] + s_part[3][r][n]);
    if (ok) p.z2[(long long)rg * kTD + n] = z;
    const float m = ok ? tail_drop(p, p.mask_f, p.stream_f, rg, n) : 1.0f;
    s_x[r][n] = (z + p.b2[n]) * m + s_h[r][n];
  }
  __syncthreads();
  {
    float mean, rstd;
    row_stats(s_x[r], p.epsF, mean, rstd);
    if (ok) {
      p.out[(long long)rg * kTD + n] = (s_x[r][n] - mean) * rstd * p.lnF_w[n] + p.lnF_b[n];
      if (n == 0) { p.st_f[2 * rg] = mean; p.st_f[2 * rg + 1] = rstd; }
    }
  }
}

// sum of s[r][n] over the CTA's rows that feed the parameter gradients, then one atomic per column (threads r == 0)
__device__ __forceinline__ void col_atomic(float (*s)[kTD], int n, int r, int n_param_rows, float* dst) {
  if (r != 0 || dst == nullptr || n_param_rows <= 0) return;
  float t = 0.f;
  for (int rr = 0; rr < n_param_rows; ++rr) t += s[rr][n];
  atomicAdd(dst + n, t);
}

__global__ void __launch_bounds__(kTailThreads) tail_bwd_kernel(const TailParams p) {
  __shared__ __align__(16) float s_a[kTailRows][kTD], s_b[kTailRows][kTD], s_c[kTailRows][kTD];
  __shared__ __align__(16) float s_dz2[kTailRows][kTD], s_dh[kTailRows][kTD], s_dhz[kTailRows][kTD];
  __shared__ __align__(16) float s_dz1[kTailRows][kTailMaxI], s_part[4][kTailRows][kTD];
  const int tid = threadIdx.x;
  const int C = p.n_groups * p.B;
  const int row0 = blockIdx.x * kTailRows;
  const int n = tid & 63, r = tid >> 6;
  const int rg = row0 + r;
  const bool ok = rg < C;
  // rows [0, B) (the calibrated stream) feed the parameter gradients
  const int n_param_rows = min(kTailRows, max(0, p.B - row0));
  const bool prow = rg < p.B;
  pdl_launch_dependents();
  pdl_wait();
  // ---- LayerNorm + dropout backward of the feed-forward output ----
  float xhat = 0.f, dy = 0.f, m = 1.0f, rstd = 0.f;
  if (ok) {
    m = tail_drop(p, p.mask_f, p.stream_f, rg, n);
    const float mean = p.st_f[2 * rg];
    rstd = p.st_f[2 * rg + 1];
    xhat = ((p.z2[(long long)rg * kTD + n] + p.b2[n]) * m + p.h[(long long)rg * kTD + n] - mean) * rstd;
    dy = p.d_out[(long long)rg * kTD + n];
  }
  const float gf = dy * p.lnF_w[n];
  s_a[r][n] = gf; s_b[r][n] = gf * xhat;
  s_c[r][n] = prow ? dy * xhat : 0.f;
  s_dhz[r][n] = prow ? dy : 0.f;                         // (scratch use: column sums of dy)
  __syncthreads();
  {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < 64; ++c) { s1 += s_a[r][c]; s2 += s_b[r][c]; }
    s1 *= (1.0f / 64); s2 *= (1.0f / 64);
    const float dx = rstd * (gf - s1 - xhat * s2);       // grad wrt (dropout(z2 + b2) + h)
    const float dz2 = dx * m;
    col_atomic(s_c, n, r, n_param_rows, p.g_lnF_w);
    col_atomic(s_dhz, n, r, n_param_rows, p.g_lnF_b);
    s_dz2[r][n] = dz2; s_dh[r][n] = dx;
    if (ok) p.d_z2[(long long)rg * kTD + n] = dz2;
  }
  __syncthreads();
  col_atomic(s_dz2, n, r, n_param_rows, p.g_b2);
  // ---- d_a1 = d_z2 . W2, activation backward: thread = inner feature (W2 column, coalesced across the threads) ----
  if (tid < p.I) {
    float acc[kTailRows] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < kTD; ++k) {
      const float wv = __ldg(p.W2 + (long long)k * p.I + tid);
#pragma unroll
      for (int rr = 0; rr < kTailRows; ++rr) acc[rr] = fmaf(s_dz2[rr][k], wv, acc[rr]);
    }
    const float b1 = p.b1[tid];
    float bsum = 0.f;
#pragma unroll
    for (int rr = 0; rr < kTailRows; ++rr) {
      float g = 0.f;
      if (row0 + rr < C) {
        g = acc[rr] * act_bwd(p.act, p.z1[(long long)(row0 + rr) * p.I + tid] + b1);
        p.d_z1[(long long)(row0 + rr) * p.I + tid] = g;
        if (row0 + rr < p.B) bsum += g;
      }
      s_dz1[rr][tid] = g;
    }
    if (p.g_b1 != nullptr && n_param_rows > 0) atomicAdd(p.g_b1 + tid, bsum);
  }
  __syncthreads();
  // ---- d_h += d_z1 . W1: thread = (column, quarter of the inner axis) ----
  {
    const int I4 = p.I >> 2, kq = r;
    float acc[kTailRows] = {0.f, 0.f, 0.f, 0.f};
    for (int k = kq * I4; k < (kq + 1) * I4; ++k) {
      const float wv = __ldg(p.W1 + (long long)k * kTD + n);
#pragma unroll
      for (int rr = 0; rr < kTailRows; ++rr) acc[rr] = fmaf(s_dz1[rr][k], wv, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < kTailRows; ++rr) s_part[kq][rr][n] = acc[rr];
  }
  __syncthreads();
  // ---- LayerNorm + dropout backward of the attention output ----
  const float dh = s_dh[r][n] + ((s_part[0][r][n] + s_part[1][r][n]) + (s_part[2][r][n] + s_part[3][r][n]));
  int b = 0, g = 0;
  long long tok = 0;
  if (ok) {
    g = rg / p.B; b = rg - g * p.B;
    tok = (long long)b * p.L + (p.item_len[b] - 1);
    m = tail_drop(p, p.mask_a, p.stream_a, rg, n);
    const float mean = p.st_a[2 * rg];
    rstd = p.st_a[2 * rg + 1];
    xhat = ((p.hz[(long long)rg * kTD + n] + p.bo[n]) * m + p.c_x[(long long)b * kTD + n] - mean) * rstd;
  } else { xhat = 0.f; rstd = 0.f; m = 1.0f; }
  const float ga = (ok ? dh : 0.f) * p.lnA_w[n];
  s_a[r][n] = ga; s_b[r][n] = ga * xhat;
  s_c[r][n] = prow ? dh * xhat : 0.f;
  s_dz2[r][n] = prow ? dh : 0.f;                         // (scratch use: column sums of the incoming gradient)
  __syncthreads();
  {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < 64; ++c) { s1 += s_a[r][c]; s2 += s_b[r][c]; }
    s1 *= (1.0f / 64); s2 *= (1.0f / 64);
    const float dx = rstd * (ga - s1 - xhat * s2);       // grad wrt (dropout(hz + bo) + x): the residual path of the layer input
    const float dhz = dx * m;
    col_atomic(s_c, n, r, n_param_rows, p.g_lnA_w);
    col_atomic(s_dz2, n, r, n_param_rows, p.g_lnA_b);
    s_dhz[r][n] = dhz;
    if (ok) {
      p.d_hz[(long long)rg * kTD + n] = dhz;
      p.d_x[g][tok * kTD + n] = dx;                      // gather_indexes backward: position len-1 of the token-major buffer
    }
  }
  __syncthreads();
  col_atomic(s_dhz, n, r, n_param_rows, p.g_bo);
  // ---- d_ctx = d_hz . Wo ----
  {
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < kTD; ++k) acc = fmaf(s_dhz[r][k], __ldg(p.Wo + k * kTD + n), acc);
    if (ok) p.d_ctx[g][tok * kTD + n] = acc;
  }
}

static int tail_validate(const TailParams& p, const char* who) {
  ACSR_REQUIRE(p.B > 0 && p.L > 0 && (p.n_groups == 1 || p.n_groups == 2), "%s: bad sizes B=%d L=%d groups=%d", who, p.B, p.L, p.n_groups);
  if (p.I <= 0 || p.I > kTailMaxI || (p.I & 15)) {
    set_error("%s: inner size %d unsupported (multiples of 16 up to %d)", who, p.I, kTailMaxI);
    return ACSR_ERR_UNSUPPORTED;
  }
  ACSR_REQUIRE(p.act >= 0 && p.act <= 4, "%s: unknown activation %d", who, p.act);
  ACSR_REQUIRE(p.p_drop >= 0.f && p.p_drop < 1.f, "%s: dropout p=%f", who, p.p_drop);
  ACSR_REQUIRE(!(p.p_drop > 0.f && p.rng == nullptr && (p.mask_a == nullptr || p.mask_f == nullptr)), "%s: p>0 needs masks or rng", who);
  return ACSR_OK;
}

}  // namespace acsr

using namespace acsr;

extern "C" {

int acsr_tail_fwd(const float* ctx_first, const float* ctx_second, const float* x, const int64_t* item_len, int B, int L, int d, int I,
                  int act, const float* Wo, const float* bo, const float* lnA_w, const float* lnA_b, float epsA, const float* W1,
                  const float* b1, const float* W2, const float* b2, const float* lnF_w, const float* lnF_b, float epsF, float p_drop,
                  const float* mask_a, const float* mask_f, const void* rng, uint32_t stream_a, uint32_t stream_f, float* c_ctx,
                  float* c_x, float* hz, float* st_a, float* h, float* z1, float* a1, float* z2, float* st_f, float* out, void* stream) {
  if (d != kTD) { set_error("tail_fwd: hidden size %d unsupported (64)", d); return ACSR_ERR_UNSUPPORTED; }
  ACSR_REQUIRE(ctx_first && x && item_len && Wo && bo && lnA_w && lnA_b && W1 && b1 && W2 && b2 && lnF_w && lnF_b, "tail_fwd: NULL input");
  ACSR_REQUIRE(c_ctx && c_x && hz && st_a && h && z1 && a1 && z2 && st_f && out, "tail_fwd: NULL output");
  TailParams p = {};
  p.B = B; p.L = L; p.I = I; p.act = act; p.n_groups = ctx_second ? 2 : 1; p.item_len = (const long long*)item_len;
  p.ctx[0] = ctx_first; p.ctx[1] = ctx_second; p.x = x;
  p.Wo = Wo; p.bo = bo; p.lnA_w = lnA_w; p.lnA_b = lnA_b; p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2; p.lnF_w = lnF_w; p.lnF_b = lnF_b;
  p.epsA = epsA; p.epsF = epsF; p.p_drop = p_drop; p.mask_a = mask_a; p.mask_f = mask_f; p.rng = (const RngState*)rng;
  p.stream_a = stream_a; p.stream_f = stream_f;
  p.c_ctx = c_ctx; p.c_x = c_x; p.hz = hz; p.st_a = st_a; p.h = h; p.z1 = z1; p.a1 = a1; p.z2 = z2; p.st_f = st_f; p.out = out;
  int rc = tail_validate(p, "tail_fwd");
  if (rc) return rc;
  const int C = p.n_groups * B;
  launch_pdl(tail_fwd_kernel, dim3((C + kTailRows - 1) / kTailRows), dim3(kTailThreads), 0, (cudaStream_t)stream, p);
  return check_launch("tail_fwd");
}

int acsr_tail_bwd(const float* d_out, const int64_t* item_len, int B, int L, int d, int I, int act, int n_groups, const float* c_x,
                  const float* hz, const float* st_a, const float* h, const float* z1, const float* z2, const float* st_f,
                  const float* Wo, const float* bo, const float* lnA_w, const float* W1, const float* b1, const float* W2,
                  const float* b2, const float* lnF_w, float p_drop, const float* mask_a, const float* mask_f, const void* rng,
                  uint32_t stream_a, uint32_t stream_f, float* d_z2, float* d_z1, float* d_hz, float* d_x_first, float* d_x_second,
                  float* d_ctx_first, float* d_ctx_second, float* g_bo, float* g_lnA_w, float* g_lnA_b, float* g_b1, float* g_b2,
                  float* g_lnF_w, float* g_lnF_b, void* stream) {
  if (d != kTD) { set_error("tail_bwd: hidden size %d unsupported (64)", d); return ACSR_ERR_UNSUPPORTED; }
  ACSR_REQUIRE(d_out && item_len && c_x && hz && st_a && h && z1 && z2 && st_f && Wo && bo && lnA_w && W1 && b1 && W2 && b2 && lnF_w,
               "tail_bwd: NULL input");
  ACSR_REQUIRE(d_z2 && d_z1 && d_hz && d_x_first && d_ctx_first, "tail_bwd: NULL output");
  ACSR_REQUIRE(n_groups == 1 || (d_x_second && d_ctx_second), "tail_bwd: the second group needs its own d_x / d_ctx");
  TailParams p = {};
  p.B = B; p.L = L; p.I = I; p.act = act; p.n_groups = n_groups; p.item_len = (const long long*)item_len;
  p.d_out = d_out; p.c_x = (float*)c_x; p.hz = (float*)hz; p.st_a = (float*)st_a; p.h = (float*)h; p.z1 = (float*)z1; p.z2 = (float*)z2;
  p.st_f = (float*)st_f;
  p.Wo = Wo; p.bo = bo; p.lnA_w = lnA_w; p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2; p.lnF_w = lnF_w;
  p.p_drop = p_drop; p.mask_a = mask_a; p.mask_f = mask_f; p.rng = (const RngState*)rng; p.stream_a = stream_a; p.stream_f = stream_f;
  p.d_z2 = d_z2; p.d_z1 = d_z1; p.d_hz = d_hz; p.d_x[0] = d_x_first; p.d_x[1] = d_x_second; p.d_ctx[0] = d_ctx_first; p.d_ctx[1] = d_ctx_second;
  p.g_bo = g_bo; p.g_lnA_w = g_lnA_w; p.g_lnA_b = g_lnA_b; p.g_b1 = g_b1; p.g_b2 = g_b2; p.g_lnF_w = g_lnF_w; p.g_lnF_b = g_lnF_b;
  int rc = tail_validate(p, "tail_bwd");
  if (rc) return rc;
  const int C = n_groups * B;
  launch_pdl(tail_bwd_kernel, dim3((C + kTailRows - 1) / kTailRows), dim3(kTailThreads), 0, (cudaStream_t)stream, p);
  return check_launch("tail_bwd");
}

}  // extern "C"
