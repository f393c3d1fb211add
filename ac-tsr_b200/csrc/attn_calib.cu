// Fused calibrated causal attention of one AC layer (forward and backward), L <= 64.
//
// One CTA owns one (sequence b, head h): the five projected [L,dh] tiles are staged in shared
// memory once, each warp then owns query rows i = warp, warp+8, ... and its lanes own key
// columns j = lane, lane+32, so the chained softmaxes (spatially-calibrated P, attack mask M,
// attacked A, calibrated C, combined R) are warp-shuffle reductions and none of the [B,H,L,L]
// intermediates of the reference (layers.py:686-742, 657-674, 917-925) ever reaches HBM.
// The additive mask (abstract_recommender.py:136-143) is derived from item_seq, the spatial-
// calibrator affine over cat(q_i,k_j) is evaluated in its rank-1 form, dropout masks and the attack
// noise come from Philox (or from explicit tensors in parity mode), and the penalty sum (1-M)^2
// (acsasrec.py:135) is reduced in the same pass.
// Causal structure is exploited exactly: a 32-column group that lies entirely above the diagonal
// is skipped (its probabilities are exactly 0 in the reference as well: exp(-10000-max) underflows),
// and all probs.V / gradient contractions run over j <= i only.
// The backward recomputes the row's probabilities from the same tiles (and the same Philox
// counters), keeps dS, dS', R, A as lower-triangular shared-memory matrices and finishes the
// column-side gradients (dK, dK', dV) in a second, column-parallel phase.
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

constexpr int kAttnWarps = 8;
constexpr int kAttnThreads = kAttnWarps * 32;

struct AttnParams {
  const float *mq, *mk, *mv, *aq, *ak, *gate;
  const int64_t* item_seq;
  const float *ow, *ob, *dw, *db, *scalar;
  int B, L, H, dh, d;
  int two_level, combine, rich;
  float comb_scalar;
  const float* rich_ratio;
  float p;
  const float *D1, *D2, *D3, *noise;
  const RngState* rng;
  uint32_t stream;
  // forward outputs
  float *ctx_att, *ctx_cal;
  double* pen_sq;
  float* probs;
  // backward
  const float *d_ctx_att, *d_ctx_cal, *d_pen;
  float *d_mq, *d_mk, *d_mv, *d_aq, *d_ak, *d_gate, *d_ow, *d_ob, *d_dw, *d_db, *d_scalar, *d_ratio;
};

// lane <-> head-dim mapping for "lane = channel" loops.  dh >= 32: lane owns channels lane, lane+32;
// dh < 32: the warp splits into 32/dh groups that share the reduction index.
template <int DH>
struct CMap {
  static constexpr int G = DH >= 32 ? 1 : 32 / DH;
  static constexpr int CPL = DH >= 32 ? DH / 32 : 1;
  __device__ static __forceinline__ int c(int lane, int k) { return DH >= 32 ? lane + 32 * k : lane % DH; }
  __device__ static __forceinline__ int grp(int lane) { return DH >= 32 ? 0 : lane / DH; }
  __device__ static __forceinline__ float reduce(float v) {
#pragma unroll
    for (int o = DH; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
};

template <int JPL>
struct RowP {
  float S[JPL], S2[JPL], sig[JPL], delta[JPL], msk[JPL];
  float Psoft[JPL], P[JPL], P0soft[JPL], P0[JPL], Msoft[JPL], M[JPL];
  float D1[JPL], D2[JPL], D3[JPL], nz[JPL];
  float O[JPL], A[JPL], expm[JPL], C[JPL], g[JPL], F[JPL], R[JPL], Rf[JPL];
  bool inb[JPL];      // column exists (j < L)
  bool act[JPL];      // column group intersects the causal triangle of this row (warp-uniform)
};

struct AttnSmem {
  float *Q, *K, *V, *Q2, *K2;                        // [L][dh+1]
  float *rowO, *rowD, *colO, *colD, *logd, *keyok;   // [L]
};

// fast-math forms (ex2/lg2/rcp approx, ~2 ulp): far inside the 1e-3 parity budget, ~10x fewer instructions
__device__ __forceinline__ float fexp(float x) { return __expf(x); }
__device__ __forceinline__ float flog(float x) { return __logf(x); }
__device__ __forceinline__ float fsigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

template <int JPL>
__device__ __forceinline__ void softmax_row(const float* z, const bool* inb, const bool* act, float* y) {
  float m = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) if (act[jj] && inb[jj]) m = fmaxf(m, z[jj]);
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) { y[jj] = (act[jj] && inb[jj]) ? fexp(z[jj] - m) : 0.f; s += y[jj]; }
  s = warp_sum(s);
  const float inv = 1.0f / s;
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) y[jj] *= inv;
}

// Y .* (dY - sum(Y .* dY))
template <int JPL>
__device__ __forceinline__ void softmax_bwd_row(const float* y, const float* dy, float* dz) {
  float s = 0.f;
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) s += y[jj] * dy[jj];
  s = warp_sum(s);
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) dz[jj] = y[jj] * (dy[jj] - s);
}

template <int DH>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, int b, int h, int L, int d) {
  constexpr int dhp = DH + 1;
  const float* base = src + (long long)b * L * d + h * DH;
  for (int e = threadIdx.x; e < L * DH; e += blockDim.x) {
    int r = e / DH, c = e % DH;
    dst[r * dhp + c] = base[(long long)r * d + c];
  }
}

// stage tiles + per-row / per-column scalars of the spatial calibrator
template <int DH>
__device__ __forceinline__ void stage_common(const AttnParams& p, const AttnSmem& sm, int b, int h) {
  constexpr int dhp = DH + 1;
  const int L = p.L;
  load_tile<DH>(sm.Q, p.mq, b, h, L, p.d);
  load_tile<DH>(sm.K, p.mk, b, h, L, p.d);
  load_tile<DH>(sm.V, p.mv, b, h, L, p.d);
  load_tile<DH>(sm.Q2, p.aq, b, h, L, p.d);
  load_tile<DH>(sm.K2, p.ak, b, h, L, p.d);
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    sm.logd[j] = logf((float)j + 1.0f);
    sm.keyok[j] = p.item_seq[(long long)b * L + j] != 0 ? 1.0f : 0.0f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float ro = 0.f, rd = 0.f, co = 0.f, cd = 0.f;
    for (int c = 0; c < DH; ++c) {
      float q = sm.Q[j * dhp + c], k = sm.K[j * dhp + c];
      if (p.ow) { ro += q * p.ow[c]; co += k * p.ow[DH + c]; }
      if (p.dw) { rd += q * p.dw[c]; cd += k * p.dw[DH + c]; }
    }
    sm.rowO[j] = ro; sm.rowD[j] = rd; sm.colO[j] = co; sm.colD[j] = cd;
  }
  __syncthreads();
}

template <int DH, int JPL>
__device__ __forceinline__ void row_forward(const AttnParams& p, const AttnSmem& sm, int b, int h, int i, int lane,
                                            bool need_att, RowP<JPL>& r) {
  constexpr int dhp = DH + 1;
  const int L = p.L;
  const float inv_sq = 1.0f / sqrtf((float)DH);
  int jr[JPL];
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) {
    int j = lane + 32 * jj;
    r.inb[jj] = j < L;
    r.act[jj] = (32 * jj) <= i;
    jr[jj] = j < L ? j : L - 1;
    r.S[jj] = 0.f; r.S2[jj] = 0.f;
  }
  const float* qi = sm.Q + i * dhp;
  const float* q2i = sm.Q2 + i * dhp;
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) {
    if (!r.act[jj]) continue;
    const float* kj = sm.K + jr[jj] * dhp;
    const float* k2j = sm.K2 + jr[jj] * dhp;
    float s = 0.f, s2 = 0.f;
#pragma unroll 8
    for (int c = 0; c < DH; ++c) {
      s = fmaf(qi[c], kj[c], s);
      s2 = fmaf(q2i[c], k2j[c], s2);
    }
    r.S[jj] = s; r.S2[jj] = s2;
  }
  const float ob = p.ob ? p.ob[0] : 0.f, db = p.db ? p.db[0] : 0.f;
  const float sc = p.scalar ? p.scalar[0] : 0.f;
  const float sc2h = sc * sc * 0.5f;
  const float rowO = sm.rowO[i], rowD = sm.rowD[i];
  float zP[JPL], z0[JPL], zM[JPL];
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) {
    const int j = jr[jj];
    const bool valid = r.inb[jj] && (j <= i) && (sm.keyok[j] != 0.f);
    r.msk[jj] = valid ? 0.f : kMaskNeg;
    r.sig[jj] = 0.f; r.delta[jj] = 0.f;
    zP[jj] = z0[jj] = zM[jj] = kMaskNeg;
    if (!r.act[jj]) continue;
    float eo = 0.f, ed = 0.f;
    if (p.ow) {
      const float sg = fsigmoid(rowO + sm.colO[j] + ob);
      r.sig[jj] = sg;
      eo = (j > i) ? flog(sg + kOrderEps) : flog((1.0f - sg) + kOrderEps);
    }
    if (p.dw) {
      const int dist = i > j ? i - j : j - i;
      const float dl = sm.logd[dist] - (rowD + sm.colD[j] + db);
      r.delta[jj] = dl;
      ed = -(dl * dl) * sc2h;
    }
    zP[jj] = (r.S[jj] + eo + ed) * inv_sq + r.msk[jj];
    z0[jj] = r.S[jj] * inv_sq + r.msk[jj];
    zM[jj] = r.S2[jj] * inv_sq + r.msk[jj];
  }
  softmax_row<JPL>(zP, r.inb, r.act, r.Psoft);
  softmax_row<JPL>(zM, r.inb, r.act, r.Msoft);
  const bool need_p0 = !p.two_level || p.probs != nullptr;
  if (need_p0) softmax_row<JPL>(z0, r.inb, r.act, r.P0soft);
  // randomness
  const float inv_keep = p.p > 0.f ? 1.0f / (1.0f - p.p) : 1.0f;
  const bool philox_drop = p.p > 0.f && p.D1 == nullptr;
  const bool philox_noise = need_att && p.noise == nullptr && p.rng != nullptr;
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) {
    r.D1[jj] = r.D2[jj] = r.D3[jj] = 1.0f;
    r.nz[jj] = 0.f;
    if (r.act[jj]) {
      const long long e = (((long long)b * p.H + h) * L + i) * L + jr[jj];
      if (philox_drop || philox_noise) {
        const uint4 w = philox4x32(p.rng->seed, p.rng->step, p.stream, (unsigned long long)e);
        if (philox_drop) { r.D1[jj] = drop_mult(w.x, p.p, inv_keep); r.D3[jj] = drop_mult(w.y, p.p, inv_keep); }
        if (philox_noise) {
          const float u1 = u32_to_unit(w.z), u2 = u32_to_unit(w.w);
          r.nz[jj] = sqrtf(-2.0f * flog(u1)) * __cosf(6.283185307179586f * u2);
        }
        if (philox_drop && need_p0) {
          const uint4 w2 = philox4x32(p.rng->seed, p.rng->step, p.stream + 1u, (unsigned long long)e);
          r.D2[jj] = drop_mult(w2.x, p.p, inv_keep);
        }
      }
      if (p.D1) r.D1[jj] = p.D1[e];
      if (p.D2) r.D2[jj] = p.D2[e];
      if (p.D3) r.D3[jj] = p.D3[e];
      if (p.noise) r.nz[jj] = p.noise[e];
    }
    r.P[jj] = r.Psoft[jj] * r.D1[jj];
    r.P0[jj] = need_p0 ? r.P0soft[jj] * r.D2[jj] : 0.f;
    r.M[jj] = r.Msoft[jj] * r.D3[jj];
    r.O[jj] = p.two_level ? r.P[jj] : r.P0[jj];
    r.expm[jj] = r.act[jj] ? fexp(1.0f - r.M[jj]) : 0.f;
  }
  float z[JPL];
  if (need_att) {
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) z[jj] = r.O[jj] * r.M[jj] + r.nz[jj] * (1.0f - r.M[jj]) + r.msk[jj];
    softmax_row<JPL>(z, r.inb, r.act, r.A);
  } else {
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) r.A[jj] = 0.f;
  }
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) z[jj] = r.O[jj] * r.expm[jj] + r.msk[jj];
  softmax_row<JPL>(z, r.inb, r.act, r.C);
  if (p.combine == ACSR_ATTN_COMBINE_FIXED) {
    // layers.py:885: softmax(origin + 0.5*calibrated) has NO mask: columns above the diagonal hold exp(0)
    bool all[JPL];
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) { z[jj] = r.O[jj] + 0.5f * r.C[jj]; r.g[jj] = 0.f; all[jj] = true; }
    softmax_row<JPL>(z, r.inb, all, r.F);
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) z[jj] = r.F[jj] + r.msk[jj];
  } else {
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      float g = p.comb_scalar;
      if (p.combine == ACSR_ATTN_COMBINE_GATE && r.act[jj]) g = fsigmoid(p.gate[((long long)b * L + i) * L + jr[jj]]);
      r.g[jj] = g; r.F[jj] = 0.f;
      z[jj] = g * r.O[jj] + (1.0f - g) * r.C[jj] + r.msk[jj];
    }
  }
  softmax_row<JPL>(z, r.inb, r.act, r.R);
  float rr = 1.0f;
  if (!p.two_level) rr = (p.rich == ACSR_ATTN_RICH_TRAINABLE) ? p.rich_ratio[0] : 0.5f;
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) r.Rf[jj] = p.two_level ? r.R[jj] : (rr * r.R[jj] + (1.0f - rr) * r.P[jj]);
}

__device__ __forceinline__ AttnSmem carve_common(float*& ptr, int L, int dh) {
  AttnSmem sm;
  const int tile = L * (dh + 1);
  sm.Q = ptr; ptr += tile;
  sm.K = ptr; ptr += tile;
  sm.V = ptr; ptr += tile;
  sm.Q2 = ptr; ptr += tile;
  sm.K2 = ptr; ptr += tile;
  sm.rowO = ptr; ptr += L;
  sm.rowD = ptr; ptr += L;
  sm.colO = ptr; ptr += L;
  sm.colD = ptr; ptr += L;
  sm.logd = ptr; ptr += L;
  sm.keyok = ptr; ptr += L;
  return sm;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int DH, int JPL>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ float smem_f[];
  constexpr int dhp = DH + 1;
  using CM = CMap<DH>;
  const int L = p.L;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ptr = smem_f;
  AttnSmem sm = carve_common(ptr, L, DH);
  float* rowbuf = ptr;            // [warps][2][L]
  ptr += kAttnWarps * 2 * L;
  double* pen_red = reinterpret_cast<double*>(smem_f + ((ptr - smem_f + 1) & ~1));   // [warps], 8-byte aligned
  stage_common<DH>(p, sm, b, h);
  const bool need_att = p.ctx_att != nullptr;
  float pen = 0.f;
  RowP<JPL> r;
  // rows are dealt so that every warp gets a similar amount of causal work (long and short rows alternate)
  for (int rnd = 0; rnd * kAttnWarps < L; ++rnd) {
    const int i = rnd * kAttnWarps + ((rnd & 1) ? (kAttnWarps - 1 - warp) : warp);   // snake order: balanced causal work
    if (i >= L) continue;
    row_forward<DH, JPL>(p, sm, b, h, i, lane, need_att, r);
    float* bufR = rowbuf + (warp * 2 + 0) * L;
    float* bufA = rowbuf + (warp * 2 + 1) * L;
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      const int j = lane + 32 * jj;
      if (r.inb[jj]) {
        const float om = 1.0f - r.M[jj];         // columns above the diagonal: M == 0 -> contributes 1
        pen += om * om;
        bufR[j] = r.Rf[jj];
        bufA[j] = r.A[jj];
        if (p.probs) {
          const long long e = (((long long)b * p.H + h) * L + i) * L + j;
          const long long plane = (long long)p.B * p.H * L * L;
          p.probs[0 * plane + e] = r.P0[jj]; p.probs[1 * plane + e] = r.P[jj]; p.probs[2 * plane + e] = r.M[jj];
          p.probs[3 * plane + e] = r.A[jj]; p.probs[4 * plane + e] = r.C[jj]; p.probs[5 * plane + e] = r.Rf[jj];
        }
      }
    }
    __syncwarp();
    // ctx[i][c] = sum_{j<=i} prob[j] * V[j][c]
    float accR[CM::CPL], accA[CM::CPL];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) accR[k] = accA[k] = 0.f;
    for (int j = CM::grp(lane); j <= i; j += CM::G) {
      const float pr = bufR[j], pa = bufA[j];
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) {
        const float v = sm.V[j * dhp + CM::c(lane, k)];
        accR[k] = fmaf(pr, v, accR[k]);
        accA[k] = fmaf(pa, v, accA[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      accR[k] = CM::reduce(accR[k]);
      accA[k] = CM::reduce(accA[k]);
      if (CM::grp(lane) == 0) {
        const long long o = ((long long)b * L + i) * p.d + h * DH + CM::c(lane, k);
        p.ctx_cal[o] = accR[k];
        if (need_att) p.ctx_att[o] = accA[k];
      }
    }
    __syncwarp();
  }
  double pd = warp_sum_d((double)pen);
  if (lane == 0) pen_red[warp] = pd;
  __syncthreads();
  if (threadIdx.x == 0 && p.pen_sq != nullptr) {
    double s = 0.0;
    for (int w = 0; w < kAttnWarps; ++w) s += pen_red[w];
    atomicAdd(p.pen_sq, s);
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

template <int DH, int JPL>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_bwd_kernel(const AttnParams p) {
  extern __shared__ float smem_f[];
  constexpr int dhp = DH + 1;
  using CM = CMap<DH>;
  const int L = p.L;
  const int ntri = L * (L + 1) / 2;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ptr = smem_f;
  AttnSmem sm = carve_common(ptr, L, DH);
  float* sDC = ptr; ptr += L * dhp;       // d_ctx_cal head slice
  float* sDA = ptr; ptr += L * dhp;       // d_ctx_att head slice
  float* matS = ptr; ptr += ntri;         // dS   (lower triangle, packed)
  float* matS2 = ptr; ptr += ntri;        // dS'
  float* matR = ptr; ptr += ntri;         // R_final
  float* matA = ptr; ptr += ntri;         // A
  float* colDU = ptr; ptr += L;
  float* colDT = ptr; ptr += L;
  float* red = ptr; ptr += kAttnWarps * 8;   // scalar partials per warp
  const bool has_att = p.d_ctx_att != nullptr;
  const bool has_cal = p.d_ctx_cal != nullptr;

  for (int e = threadIdx.x; e < L * DH; e += blockDim.x) {
    const int rr = e / DH, c = e % DH;
    const long long o = ((long long)b * L + rr) * p.d + h * DH + c;
    sDC[rr * dhp + c] = has_cal ? p.d_ctx_cal[o] : 0.f;
    sDA[rr * dhp + c] = has_att ? p.d_ctx_att[o] : 0.f;
  }
  for (int j = threadIdx.x; j < L; j += blockDim.x) { colDU[j] = 0.f; colDT[j] = 0.f; }
  stage_common<DH>(p, sm, b, h);     // ends with __syncthreads()

  const float inv_sq = 1.0f / sqrtf((float)DH);
  const float dpen = p.d_pen ? p.d_pen[0] : 0.f;
  const float sc = p.scalar ? p.scalar[0] : 0.f;
  const float sc2 = sc * sc;
  float rr = 1.0f;
  if (!p.two_level) rr = (p.rich == ACSR_ATTN_RICH_TRAINABLE) ? p.rich_ratio[0] : 0.5f;

  float accOq[CM::CPL], accDq[CM::CPL];
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) accOq[k] = accDq[k] = 0.f;
  float s_ob = 0.f, s_db = 0.f, s_scalar = 0.f, s_ratio = 0.f;

  RowP<JPL> r;
  for (int rnd = 0; rnd * kAttnWarps < L; ++rnd) {
    const int i = rnd * kAttnWarps + ((rnd & 1) ? (kAttnWarps - 1 - warp) : warp);   // snake order: balanced causal work
    if (i >= L) continue;
    row_forward<DH, JPL>(p, sm, b, h, i, lane, has_att, r);
    int jr[JPL];
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) jr[jj] = (lane + 32 * jj) < L ? lane + 32 * jj : L - 1;
    // dRf_j = dctx_cal_i . v_j ; dA_j = dctx_att_i . v_j   (only column groups that touch the triangle)
    float dRf[JPL], dA[JPL];
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      dRf[jj] = dA[jj] = 0.f;
      if (!r.act[jj]) continue;
      const float* vj = sm.V + jr[jj] * dhp;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
      for (int c = 0; c < DH; ++c) {
        const float v = vj[c];
        a0 = fmaf(sDC[i * dhp + c], v, a0);
        a1 = fmaf(sDA[i * dhp + c], v, a1);
      }
      if (r.inb[jj]) { dRf[jj] = a0; dA[jj] = a1; }
    }
    float dO[JPL], dP[JPL], dM[JPL], dR[JPL], dC[JPL], tmp[JPL], dcm[JPL];
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      dO[jj] = 0.f; dP[jj] = 0.f; dM[jj] = 0.f;
      if (p.two_level) dR[jj] = dRf[jj];
      else {
        dR[jj] = dRf[jj] * rr;
        dP[jj] = dRf[jj] * (1.0f - rr);
        s_ratio += dRf[jj] * (r.R[jj] - r.P[jj]);
      }
    }
    softmax_bwd_row<JPL>(r.R, dR, dcm);            // grad wrt (comb + mask)
    if (p.combine == ACSR_ATTN_COMBINE_FIXED) {
      softmax_bwd_row<JPL>(r.F, dcm, tmp);          // grad wrt (O + 0.5 C)
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj) { dO[jj] += tmp[jj]; dC[jj] = 0.5f * tmp[jj]; }
    } else {
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj) {
        const float g = r.g[jj];
        dO[jj] += dcm[jj] * g;
        dC[jj] = dcm[jj] * (1.0f - g);
        if (p.combine == ACSR_ATTN_COMBINE_GATE && r.inb[jj] && r.act[jj]) {
          const float dgl = dcm[jj] * (r.O[jj] - r.C[jj]) * g * (1.0f - g);
          if (dgl != 0.f) atomicAdd(p.d_gate + ((long long)b * L + i) * L + jr[jj], dgl);
        }
      }
    }
    softmax_bwd_row<JPL>(r.C, dC, tmp);             // grad wrt (O*expm + mask)
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      dO[jj] += tmp[jj] * r.expm[jj];
      dM[jj] -= tmp[jj] * r.O[jj] * r.expm[jj];
    }
    if (has_att) {
      softmax_bwd_row<JPL>(r.A, dA, tmp);           // grad wrt (O*M + n(1-M) + mask)
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj) {
        dO[jj] += tmp[jj] * r.M[jj];
        dM[jj] += tmp[jj] * (r.O[jj] - r.nz[jj]);
      }
    }
    float dP0[JPL];
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      if (r.inb[jj] && r.act[jj]) dM[jj] += dpen * (-2.0f) * (1.0f - r.M[jj]);
      dM[jj] *= r.D3[jj];
      if (p.two_level) { dP[jj] += dO[jj]; dP0[jj] = 0.f; }
      else dP0[jj] = dO[jj] * r.D2[jj];
      dP[jj] *= r.D1[jj];
    }
    float dS2[JPL], dz[JPL], dS[JPL];
    softmax_bwd_row<JPL>(r.Msoft, dM, dS2);
    softmax_bwd_row<JPL>(r.Psoft, dP, dz);
    float row_du = 0.f, row_dt = 0.f;
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) { dS2[jj] *= inv_sq; dz[jj] *= inv_sq; dS[jj] = dz[jj]; }
    if (!p.two_level) {
      softmax_bwd_row<JPL>(r.P0soft, dP0, tmp);
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj) dS[jj] += tmp[jj] * inv_sq;
    }
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      const int j = lane + 32 * jj;
      if (!r.inb[jj] || j > i) continue;             // everything above the diagonal is exactly zero
      if (p.ow) {
        const float sg = r.sig[jj];
        const float de = -sg * (1.0f - sg) / ((1.0f - sg) + kOrderEps);     // j <= i branch of layers.py:719
        const float du = dz[jj] * de;
        row_du += du;
        if (du != 0.f) atomicAdd(colDU + j, du);
      }
      if (p.dw) {
        const float dl = r.delta[jj];
        const float dt = dz[jj] * dl * sc2;
        row_dt += dt;
        if (dt != 0.f) atomicAdd(colDT + j, dt);
        s_scalar += dz[jj] * (-(dl * dl) * sc);
      }
      const int t = tri(i, j);
      matS[t] = dS[jj];
      matS2[t] = dS2[jj];
      matR[t] = r.Rf[jj];
      matA[t] = r.A[jj];
    }
    row_du = warp_sum(row_du);
    row_dt = warp_sum(row_dt);
    s_ob += row_du;            // identical on every lane; lane 0 publishes it
    s_db += row_dt;
    __syncwarp();
    // row-side gradients: dq_i, dq'_i  (lane = channel, j <= i)
    float aq_[CM::CPL], aq2_[CM::CPL];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) aq_[k] = aq2_[k] = 0.f;
    const int tb = tri(i, 0);
    for (int j = CM::grp(lane); j <= i; j += CM::G) {
      const float s1 = matS[tb + j], s2 = matS2[tb + j];
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) {
        const int c = CM::c(lane, k);
        aq_[k] = fmaf(s1, sm.K[j * dhp + c], aq_[k]);
        aq2_[k] = fmaf(s2, sm.K2[j * dhp + c], aq2_[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      const int c = CM::c(lane, k);
      float v1 = CM::reduce(aq_[k]), v2 = CM::reduce(aq2_[k]);
      const float qic = sm.Q[i * dhp + c];
      if (p.ow) { v1 += row_du * p.ow[c]; accOq[k] += row_du * qic; }
      if (p.dw) { v1 += row_dt * p.dw[c]; accDq[k] += row_dt * qic; }
      if (CM::grp(lane) == 0) {
        const long long o = ((long long)b * L + i) * p.d + h * DH + c;
        p.d_mq[o] = v1;
        p.d_aq[o] = v2;
      }
    }
  }
  __syncthreads();
  // column-side gradients: dk_j, dk'_j, dv_j  (warp per column, lane = channel, rows i >= j)
  float accOk[CM::CPL], accDk[CM::CPL];
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) accOk[k] = accDk[k] = 0.f;
  for (int j = warp; j < L; j += kAttnWarps) {
    float ak_[CM::CPL], ak2_[CM::CPL], av_[CM::CPL];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) ak_[k] = ak2_[k] = av_[k] = 0.f;
    for (int i = j + CM::grp(lane); i < L; i += CM::G) {
      const int t = tri(i, j);
      const float s1 = matS[t], s2 = matS2[t], pr = matR[t], pa = matA[t];
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) {
        const int c = CM::c(lane, k);
        ak_[k] = fmaf(s1, sm.Q[i * dhp + c], ak_[k]);
        ak2_[k] = fmaf(s2, sm.Q2[i * dhp + c], ak2_[k]);
        av_[k] = fmaf(pr, sDC[i * dhp + c], av_[k]);
        av_[k] = fmaf(pa, sDA[i * dhp + c], av_[k]);
      }
    }
    const float cdu = colDU[j], cdt = colDT[j];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      const int c = CM::c(lane, k);
      float v1 = CM::reduce(ak_[k]), v2 = CM::reduce(ak2_[k]), v3 = CM::reduce(av_[k]);
      const float kjc = sm.K[j * dhp + c];
      if (p.ow) { v1 += cdu * p.ow[DH + c]; accOk[k] += cdu * kjc; }
      if (p.dw) { v1 += cdt * p.dw[DH + c]; accDk[k] += cdt * kjc; }
      if (CM::grp(lane) == 0) {
        const long long o = ((long long)b * L + j) * p.d + h * DH + c;
        p.d_mk[o] = v1;
        p.d_ak[o] = v2;
        p.d_mv[o] = v3;
      }
    }
  }
  // parameter gradients: lane-held channel partials -> one atomic per lane per warp
  if (CM::grp(lane) == 0) {
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      const int c = CM::c(lane, k);
      if (p.d_ow) { atomicAdd(p.d_ow + c, accOq[k]); atomicAdd(p.d_ow + DH + c, accOk[k]); }
      if (p.d_dw) { atomicAdd(p.d_dw + c, accDq[k]); atomicAdd(p.d_dw + DH + c, accDk[k]); }
    }
  }
  s_scalar = warp_sum(s_scalar);
  s_ratio = warp_sum(s_ratio);
  if (lane == 0) {
    red[warp * 4 + 0] = s_ob; red[warp * 4 + 1] = s_db; red[warp * 4 + 2] = s_scalar; red[warp * 4 + 3] = s_ratio;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < kAttnWarps; ++w) s += red[w * 4 + threadIdx.x];
    float* dst = threadIdx.x == 0 ? p.d_ob : threadIdx.x == 1 ? p.d_db : threadIdx.x == 2 ? p.d_scalar : p.d_ratio;
    if (dst != nullptr && s != 0.f) atomicAdd(dst, s);
  }
}

static size_t fwd_smem_bytes(int L, int dh) {
  size_t f = (size_t)5 * L * (dh + 1) + 6 * L + (size_t)kAttnWarps * 2 * L + 2;
  return f * sizeof(float) + kAttnWarps * sizeof(double);
}
static size_t bwd_smem_bytes(int L, int dh) {
  size_t f = (size_t)7 * L * (dh + 1) + 6 * L + (size_t)4 * (L * (L + 1) / 2) + 2 * L + kAttnWarps * 8;
  return f * sizeof(float);
}

template <typename K>
static int prep_kernel(K kernel, size_t smem, const char* who) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) { set_error("%s: smem %zu: %s", who, smem, cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  return ACSR_OK;
}

template <int DH, int JPL>
static int launch_fwd(const AttnParams& p, cudaStream_t st) {
  size_t smem = fwd_smem_bytes(p.L, DH);
  int rc = prep_kernel(attn_fwd_kernel<DH, JPL>, smem, "attn_calib_fwd");
  if (rc) return rc;
  attn_fwd_kernel<DH, JPL><<<p.B * p.H, kAttnThreads, smem, st>>>(p);
  return check_launch("attn_calib_fwd");
}
template <int DH, int JPL>
static int launch_bwd(const AttnParams& p, cudaStream_t st) {
  size_t smem = bwd_smem_bytes(p.L, DH);
  int rc = prep_kernel(attn_bwd_kernel<DH, JPL>, smem, "attn_calib_bwd");
  if (rc) return rc;
  attn_bwd_kernel<DH, JPL><<<p.B * p.H, kAttnThreads, smem, st>>>(p);
  return check_launch("attn_calib_bwd");
}

#define ATTN_DISPATCH(FN, p, st)                                                              \
  do {                                                                                        \
    const int jpl = (p.L + 31) / 32;                                                          \
    if (jpl == 1) {                                                                           \
      switch (p.dh) {                                                                         \
        case 8: return FN<8, 1>(p, st);                                                       \
        case 16: return FN<16, 1>(p, st);                                                     \
        case 32: return FN<32, 1>(p, st);                                                     \
        case 64: return FN<64, 1>(p, st);                                                     \
      }                                                                                       \
    } else {                                                                                  \
      switch (p.dh) {                                                                         \
        case 8: return FN<8, 2>(p, st);                                                       \
        case 16: return FN<16, 2>(p, st);                                                     \
        case 32: return FN<32, 2>(p, st);                                                     \
        case 64: return FN<64, 2>(p, st);                                                     \
      }                                                                                       \
    }                                                                                         \
  } while (0)

static int validate(const AttnParams& p, const char* who) {
  ACSR_REQUIRE(p.mq && p.mk && p.mv && p.aq && p.ak && p.item_seq, "%s: NULL input", who);
  ACSR_REQUIRE(p.B > 0 && p.H > 0, "%s: bad B/H", who);
  if (p.L < 1 || p.L > 64) { set_error("%s: L=%d unsupported in ABI v1 (1..64)", who, p.L); return ACSR_ERR_UNSUPPORTED; }
  if (!(p.dh == 8 || p.dh == 16 || p.dh == 32 || p.dh == 64)) {
    set_error("%s: head size %d unsupported (8/16/32/64)", who, p.dh);
    return ACSR_ERR_UNSUPPORTED;
  }
  ACSR_REQUIRE((p.ow == nullptr) == (p.ob == nullptr), "%s: order_w/order_b mismatch", who);
  ACSR_REQUIRE((p.dw == nullptr) == (p.db == nullptr) && (p.dw == nullptr) == (p.scalar == nullptr), "%s: distance params mismatch", who);
  ACSR_REQUIRE(p.combine >= 0 && p.combine <= 2, "%s: unknown combine_option %d", who, p.combine);
  ACSR_REQUIRE(p.combine != ACSR_ATTN_COMBINE_GATE || p.gate != nullptr, "%s: combine_option gate needs gate_logit", who);
  ACSR_REQUIRE(p.two_level || p.rich == ACSR_ATTN_RICH_FIXED || (p.rich == ACSR_ATTN_RICH_TRAINABLE && p.rich_ratio),
               "%s: two_level=False needs rich_calibrated_combine fixed/trainable", who);
  ACSR_REQUIRE(p.p >= 0.f && p.p < 1.f, "%s: dropout p=%f", who, p.p);
  ACSR_REQUIRE(!(p.p > 0.f && p.D1 == nullptr && p.rng == nullptr), "%s: p>0 needs explicit masks or rng", who);
  ACSR_REQUIRE((p.D1 == nullptr) == (p.D3 == nullptr), "%s: D1/D3 must be given together", who);
  return ACSR_OK;
}

}  // namespace acsr

using namespace acsr;

extern "C" {

int acsr_attn_calib_fwd(const float* mq, const float* mk, const float* mv, const float* aq, const float* ak,
                        const float* gate_logit, const int64_t* item_seq, const float* order_w, const float* order_b,
                        const float* dist_w, const float* dist_b, const float* scalar, int B, int L, int H, int dh,
                        int two_level, int combine_option, float comb_scalar, int rich_mode, const float* rich_ratio,
                        float p_attn, const float* D1, const float* D2, const float* D3, const float* noise, const void* rng,
                        uint32_t rng_stream, float* ctx_att, float* ctx_cal, double* pen_sq, float* probs_out, void* stream) {
  AttnParams p = {};
  p.mq = mq; p.mk = mk; p.mv = mv; p.aq = aq; p.ak = ak; p.gate = gate_logit; p.item_seq = item_seq;
  p.ow = order_w; p.ob = order_b; p.dw = dist_w; p.db = dist_b; p.scalar = scalar;
  p.B = B; p.L = L; p.H = H; p.dh = dh; p.d = H * dh;
  p.two_level = two_level; p.combine = combine_option; p.rich = rich_mode; p.comb_scalar = comb_scalar; p.rich_ratio = rich_ratio;
  p.p = p_attn; p.D1 = D1; p.D2 = D2; p.D3 = D3; p.noise = noise; p.rng = (const RngState*)rng; p.stream = rng_stream;
  p.ctx_att = ctx_att; p.ctx_cal = ctx_cal; p.pen_sq = pen_sq; p.probs = probs_out;
  int rc = validate(p, "attn_calib_fwd");
  if (rc) return rc;
  ACSR_REQUIRE(ctx_cal != nullptr, "attn_calib_fwd: ctx_cal is NULL");
  ATTN_DISPATCH(launch_fwd, p, (cudaStream_t)stream);
  return ACSR_ERR_UNSUPPORTED;
}

int acsr_attn_calib_bwd(const float* d_ctx_att, const float* d_ctx_cal, const float* d_pen_sq, const float* mq, const float* mk,
                        const float* mv, const float* aq, const float* ak, const float* gate_logit, const int64_t* item_seq,
                        const float* order_w, const float* order_b, const float* dist_w, const float* dist_b, const float* scalar,
                        int B, int L, int H, int dh, int two_level, int combine_option, float comb_scalar, int rich_mode,
                        const float* rich_ratio, float p_attn, const float* D1, const float* D2, const float* D3,
                        const float* noise, const void* rng, uint32_t rng_stream, float* d_mq, float* d_mk, float* d_mv,
                        float* d_aq, float* d_ak, float* d_gate_logit, float* d_order_w, float* d_order_b, float* d_dist_w,
                        float* d_dist_b, float* d_scalar, float* d_rich_ratio, void* stream) {
  AttnParams p = {};
  p.mq = mq; p.mk = mk; p.mv = mv; p.aq = aq; p.ak = ak; p.gate = gate_logit; p.item_seq = item_seq;
  p.ow = order_w; p.ob = order_b; p.dw = dist_w; p.db = dist_b; p.scalar = scalar;
  p.B = B; p.L = L; p.H = H; p.dh = dh; p.d = H * dh;
  p.two_level = two_level; p.combine = combine_option; p.rich = rich_mode; p.comb_scalar = comb_scalar; p.rich_ratio = rich_ratio;
  p.p = p_attn; p.D1 = D1; p.D2 = D2; p.D3 = D3; p.noise = noise; p.rng = (const RngState*)rng; p.stream = rng_stream;
  p.d_ctx_att = d_ctx_att; p.d_ctx_cal = d_ctx_cal; p.d_pen = d_pen_sq;
  p.d_mq = d_mq; p.d_mk = d_mk; p.d_mv = d_mv; p.d_aq = d_aq; p.d_ak = d_ak; p.d_gate = d_gate_logit;
  p.d_ow = d_order_w; p.d_ob = d_order_b; p.d_dw = d_dist_w; p.d_db = d_dist_b; p.d_scalar = d_scalar; p.d_ratio = d_rich_ratio;
  int rc = validate(p, "attn_calib_bwd");
  if (rc) return rc;
  ACSR_REQUIRE(d_mq && d_mk && d_mv && d_aq && d_ak, "attn_calib_bwd: NULL output");
  ACSR_REQUIRE(combine_option != ACSR_ATTN_COMBINE_GATE || d_gate_logit != nullptr, "attn_calib_bwd: d_gate_logit is NULL");
  // d_order_*, d_dist_*, d_scalar, d_rich_ratio may be NULL: that cotangent stream does not own those parameters
  ATTN_DISPATCH(launch_bwd, p, (cudaStream_t)stream);
  return ACSR_ERR_UNSUPPORTED;
}

}  // extern "C"
