// Fused calibrated causal attention of one AC layer (forward and backward), L <= 64.
//
// One CTA owns one (sequence b, head h).  The projected [L,dh] tiles (and the gate-logit tile) are
// staged in shared memory with 16-byte cp.async copies (rows padded to dh+4 floats: 16-byte aligned
// and bank-conflict free for both "lane = key" float4 reads and "lane = channel" scalar reads).
// Each warp then owns query rows and its lanes own key columns j = lane, lane+32, so the chained
// softmaxes (spatially-calibrated P, attack mask M, attacked A, calibrated C, combined R) are
// warp-shuffle reductions and none of the [B,H,L,L] intermediates of the reference
// (layers.py:686-742, 657-674, 917-925) ever reaches HBM.  The additive mask
// (abstract_recommender.py:136-143) is derived from item_seq, the spatial-calibrator affine over
// cat(q_i,k_j) is evaluated in its rank-1 form, dropout masks and the attack noise come from Philox
// (or from explicit tensors in parity mode), and the penalty sum (1-M)^2 (acsasrec.py:135) is reduced
// in the same pass.
// Causal structure is exploited exactly: a 32-column group that lies entirely above the diagonal is
// skipped (its probabilities are exactly 0 in the reference as well: exp(-10000-max) underflows) and
// all probs.V / gradient contractions run over j <= i only.
// The backward recomputes the row's probabilities from the same tiles (and the same Philox counters),
// finishes the row-side gradients (dQ, dQ') from per-warp row buffers, stores dS, dS', R, A TRANSPOSED
// in shared memory and finishes the column-side gradients (dK, dK', dV) in a second, column-parallel
// phase with float4 broadcast reads.
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
// dh < 32: the warp splits into 32/dh groups, each reducing a contiguous quarter-aligned slice.
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
  // slice [lo,hi) of the reduction range [begin,end) owned by this lane's group; begin % 4 == 0, lo % 4 == 0
  __device__ static __forceinline__ void slice(int lane, int begin, int end, int& lo, int& hi) {
    if (G == 1) { lo = begin; hi = end; return; }
    const int chunk = (((end - begin + G - 1) / G) + 3) & ~3;
    lo = begin + grp(lane) * chunk;
    hi = min(end, lo + chunk);
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
  float *Q, *K, *V, *Q2, *K2;                        // [LP][dh+4]
  float *G;                                          // gate logits [L*L]
  float *rowO, *rowD, *colO, *colD, *logd, *keyok;   // [LP]
  float *wo, *wd;                                    // [2*dh] spatial-calibrator weights (0 when absent)
};

struct RowConst {      // per-launch scalars hoisted out of the row loop
  float ob, db, sc, sc2h, rr, inv_sq, inv_keep;
  unsigned long long seed, step;
  bool philox_drop, philox_noise, need_p0;
};

// fast-math forms (ex2/lg2/rcp approx, ~2 ulp): far inside the 1e-3 parity budget, ~10x fewer instructions
__device__ __forceinline__ float fexp(float x) { return __expf(x); }
__device__ __forceinline__ float flog(float x) { return __logf(x); }
__device__ __forceinline__ float fsigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

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

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// async copy of one [L, DH] head slice into a padded smem tile; src == nullptr -> zero tile
template <int DH>
__device__ __forceinline__ void stage_tile(float* dst, const float* __restrict__ src, int b, int h, int L, int d) {
  constexpr int dhp = DH + 4;
  if (src == nullptr) {
    for (int e = threadIdx.x; e < L * (DH / 4); e += blockDim.x) {
      const int r = e / (DH / 4), c4 = e % (DH / 4);
      *reinterpret_cast<float4*>(dst + r * dhp + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const float* base = src + (long long)b * L * d + h * DH;
  for (int e = threadIdx.x; e < L * (DH / 4); e += blockDim.x) {
    const int r = e / (DH / 4), c4 = e % (DH / 4);
    cp_async16(dst + r * dhp + c4 * 4, base + (long long)r * d + c4 * 4);
  }
}

// stage tiles + gate logits + per-row / per-column scalars of the spatial calibrator
template <int DH>
__device__ __forceinline__ void stage_common(const AttnParams& p, const AttnSmem& sm, int b, int h, int LP) {
  constexpr int dhp = DH + 4;
  const int L = p.L;
  stage_tile<DH>(sm.Q, p.mq, b, h, L, p.d);
  stage_tile<DH>(sm.K, p.mk, b, h, L, p.d);
  stage_tile<DH>(sm.V, p.mv, b, h, L, p.d);
  stage_tile<DH>(sm.Q2, p.aq, b, h, L, p.d);
  stage_tile<DH>(sm.K2, p.ak, b, h, L, p.d);
  if (p.gate != nullptr) {
    const float* g = p.gate + (long long)b * L * L;
    if (((L * L) & 3) == 0) {
      for (int e = threadIdx.x; e < L * L / 4; e += blockDim.x) cp_async16(sm.G + 4 * e, g + 4 * e);
    } else {
      for (int e = threadIdx.x; e < L * L; e += blockDim.x) cp_async4(sm.G + e, g + e);
    }
  }
  // zero the padding rows [L, LP) of every tile (float4 reads may touch them; they must be finite)
  for (int e = threadIdx.x; e < (LP - L) * dhp; e += blockDim.x) {
    const int o = L * dhp + e;
    sm.Q[o] = 0.f; sm.K[o] = 0.f; sm.V[o] = 0.f; sm.Q2[o] = 0.f; sm.K2[o] = 0.f;
  }
  for (int j = threadIdx.x; j < LP; j += blockDim.x) {
    sm.logd[j] = logf((float)j + 1.0f);
    sm.keyok[j] = (j < L && p.item_seq[(long long)b * L + j] != 0) ? 1.0f : 0.0f;
  }
  for (int c = threadIdx.x; c < 2 * DH; c += blockDim.x) {
    sm.wo[c] = p.ow ? p.ow[c] : 0.f;
    sm.wd[c] = p.dw ? p.dw[c] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  for (int j = threadIdx.x; j < LP; j += blockDim.x) {
    float ro = 0.f, rd = 0.f, co = 0.f, cd = 0.f;
    if (j < L) {
      for (int c = 0; c < DH; ++c) {
        const float q = sm.Q[j * dhp + c], k = sm.K[j * dhp + c];
        ro = fmaf(q, sm.wo[c], ro); co = fmaf(k, sm.wo[DH + c], co);
        rd = fmaf(q, sm.wd[c], rd); cd = fmaf(k, sm.wd[DH + c], cd);
      }
    }
    sm.rowO[j] = ro; sm.rowD[j] = rd; sm.colO[j] = co; sm.colD[j] = cd;
  }
  __syncthreads();
}

template <int DH>
__device__ __forceinline__ RowConst make_consts(const AttnParams& p, bool need_att) {
  RowConst k;
  k.ob = p.ob ? p.ob[0] : 0.f;
  k.db = p.db ? p.db[0] : 0.f;
  k.sc = p.scalar ? p.scalar[0] : 0.f;
  k.sc2h = k.sc * k.sc * 0.5f;
  k.rr = 1.0f;
  if (!p.two_level) k.rr = (p.rich == ACSR_ATTN_RICH_TRAINABLE) ? p.rich_ratio[0] : 0.5f;
  k.inv_sq = 1.0f / sqrtf((float)DH);
  k.inv_keep = p.p > 0.f ? 1.0f / (1.0f - p.p) : 1.0f;
  k.philox_drop = p.p > 0.f && p.D1 == nullptr;
  k.philox_noise = need_att && p.noise == nullptr && p.rng != nullptr;
  k.need_p0 = !p.two_level || p.probs != nullptr;
  k.seed = 0; k.step = 0;
  if (p.rng != nullptr) { k.seed = p.rng->seed; k.step = p.rng->step; }
  return k;
}

template <int DH, int JPL>
__device__ __forceinline__ void row_forward(const AttnParams& p, const AttnSmem& sm, const RowConst& kc, int b, int h, int i,
                                            int lane, bool need_att, RowP<JPL>& r) {
  constexpr int dhp = DH + 4;
  const int L = p.L;
  int jr[JPL];
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) {
    const int j = lane + 32 * jj;
    r.inb[jj] = j < L;
    r.act[jj] = (32 * jj) <= i;
    jr[jj] = j < L ? j : L - 1;
    r.S[jj] = 0.f; r.S2[jj] = 0.f;
  }
  const float4* qi = reinterpret_cast<const float4*>(sm.Q + i * dhp);
  const float4* q2i = reinterpret_cast<const float4*>(sm.Q2 + i * dhp);
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) {
    if (!r.act[jj]) continue;
    const float4* kj = reinterpret_cast<const float4*>(sm.K + jr[jj] * dhp);
    const float4* k2j = reinterpret_cast<const float4*>(sm.K2 + jr[jj] * dhp);
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; ++c4) {
      s = dot4(qi[c4], kj[c4], s);
      s2 = dot4(q2i[c4], k2j[c4], s2);
    }
    r.S[jj] = s; r.S2[jj] = s2;
  }
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
      const float sg = fsigmoid(rowO + sm.colO[j] + kc.ob);
      r.sig[jj] = sg;
      eo = (j > i) ? flog(sg + kOrderEps) : flog((1.0f - sg) + kOrderEps);
    }
    if (p.dw) {
      const int dist = i > j ? i - j : j - i;
      const float dl = sm.logd[dist] - (rowD + sm.colD[j] + kc.db);
      r.delta[jj] = dl;
      ed = -(dl * dl) * kc.sc2h;
    }
    zP[jj] = (r.S[jj] + eo + ed) * kc.inv_sq + r.msk[jj];
    z0[jj] = r.S[jj] * kc.inv_sq + r.msk[jj];
    zM[jj] = r.S2[jj] * kc.inv_sq + r.msk[jj];
  }
  softmax_row<JPL>(zP, r.inb, r.act, r.Psoft);
  softmax_row<JPL>(zM, r.inb, r.act, r.Msoft);
  if (kc.need_p0) softmax_row<JPL>(z0, r.inb, r.act, r.P0soft);
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) {
    r.D1[jj] = r.D2[jj] = r.D3[jj] = 1.0f;
    r.nz[jj] = 0.f;
    if (r.act[jj]) {
      const long long e = (((long long)b * p.H + h) * L + i) * L + jr[jj];
      if (kc.philox_drop || kc.philox_noise) {
        const uint4 w = philox4x32(kc.seed, kc.step, p.stream, (unsigned long long)e);
        if (kc.philox_drop) { r.D1[jj] = drop_mult(w.x, p.p, kc.inv_keep); r.D3[jj] = drop_mult(w.y, p.p, kc.inv_keep); }
        if (kc.philox_noise) {
          const float u1 = u32_to_unit(w.z), u2 = u32_to_unit(w.w);
          r.nz[jj] = sqrtf(-2.0f * flog(u1)) * __cosf(6.283185307179586f * u2);
        }
        if (kc.philox_drop && kc.need_p0) {
          const uint4 w2 = philox4x32(kc.seed, kc.step, p.stream + 1u, (unsigned long long)e);
          r.D2[jj] = drop_mult(w2.x, p.p, kc.inv_keep);
        }
      }
      if (p.D1) r.D1[jj] = p.D1[e];
      if (p.D2) r.D2[jj] = p.D2[e];
      if (p.D3) r.D3[jj] = p.D3[e];
      if (p.noise) r.nz[jj] = p.noise[e];
    }
    r.P[jj] = r.Psoft[jj] * r.D1[jj];
    r.P0[jj] = kc.need_p0 ? r.P0soft[jj] * r.D2[jj] : 0.f;
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
      if (p.combine == ACSR_ATTN_COMBINE_GATE && r.act[jj]) g = fsigmoid(sm.G[i * L + jr[jj]]);
      r.g[jj] = g; r.F[jj] = 0.f;
      z[jj] = g * r.O[jj] + (1.0f - g) * r.C[jj] + r.msk[jj];
    }
  }
  softmax_row<JPL>(z, r.inb, r.act, r.R);
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj) r.Rf[jj] = p.two_level ? r.R[jj] : (kc.rr * r.R[jj] + (1.0f - kc.rr) * r.P[jj]);
}

__device__ __forceinline__ AttnSmem carve_common(float*& ptr, int L, int LP, int dh, bool has_gate) {
  AttnSmem sm;
  const int tile = LP * (dh + 4);
  sm.Q = ptr; ptr += tile;
  sm.K = ptr; ptr += tile;
  sm.V = ptr; ptr += tile;
  sm.Q2 = ptr; ptr += tile;
  sm.K2 = ptr; ptr += tile;
  sm.G = ptr; ptr += has_gate ? ((L * L + 3) & ~3) : 0;
  sm.rowO = ptr; ptr += LP;
  sm.rowD = ptr; ptr += LP;
  sm.colO = ptr; ptr += LP;
  sm.colD = ptr; ptr += LP;
  sm.logd = ptr; ptr += LP;
  sm.keyok = ptr; ptr += LP;
  sm.wo = ptr; ptr += 2 * dh;
  sm.wd = ptr; ptr += 2 * dh;
  return sm;
}
static size_t common_floats(int L, int LP, int dh, bool has_gate) {
  return (size_t)5 * LP * (dh + 4) + (has_gate ? ((L * L + 3) & ~3) : 0) + 6 * LP + 4 * dh;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int DH, int JPL>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int dhp = DH + 4;
  using CM = CMap<DH>;
  const int L = p.L, LP = (L + 3) & ~3;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ptr = smem_f;
  AttnSmem sm = carve_common(ptr, L, LP, DH, p.gate != nullptr);
  float* rowbuf = ptr;            // [warps][2][LP]
  ptr += kAttnWarps * 2 * LP;
  double* pen_red = reinterpret_cast<double*>(ptr);   // [warps]; ptr offset is a multiple of 4 floats
  stage_common<DH>(p, sm, b, h, LP);
  const bool need_att = p.ctx_att != nullptr;
  const RowConst kc = make_consts<DH>(p, need_att);
  float pen = 0.f;
  RowP<JPL> r;
  for (int rnd = 0; rnd * kAttnWarps < L; ++rnd) {
    const int i = rnd * kAttnWarps + ((rnd & 1) ? (kAttnWarps - 1 - warp) : warp);   // snake order: balanced causal work
    if (i >= L) continue;
    row_forward<DH, JPL>(p, sm, kc, b, h, i, lane, need_att, r);
    float* bufR = rowbuf + (warp * 2 + 0) * LP;
    float* bufA = rowbuf + (warp * 2 + 1) * LP;
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      const int j = lane + 32 * jj;
      if (j < LP) {                               // padding entries [L,LP) are written as 0
        bufR[j] = r.inb[jj] ? r.Rf[jj] : 0.f;
        bufA[j] = r.inb[jj] ? r.A[jj] : 0.f;
      }
      if (r.inb[jj]) {
        const float om = 1.0f - r.M[jj];         // columns above the diagonal: M == 0 -> contributes 1
        pen += om * om;
        if (p.probs) {
          const long long e = (((long long)b * p.H + h) * L + i) * L + j;
          const long long plane = (long long)p.B * p.H * L * L;
          p.probs[0 * plane + e] = r.P0[jj]; p.probs[1 * plane + e] = r.P[jj]; p.probs[2 * plane + e] = r.M[jj];
          p.probs[3 * plane + e] = r.A[jj]; p.probs[4 * plane + e] = r.C[jj]; p.probs[5 * plane + e] = r.Rf[jj];
        }
      }
    }
    __syncwarp();
    // ctx[i][c] = sum_{j<=i} prob[j] * V[j][c]   (entries j > i of the row buffers are exactly 0)
    float accR[CM::CPL], accA[CM::CPL];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) accR[k] = accA[k] = 0.f;
    int lo, hi;
    CM::slice(lane, 0, (i + 4) & ~3, lo, hi);
    for (int j = lo; j < hi; j += 4) {
      const float4 pr = *reinterpret_cast<const float4*>(bufR + j);
      const float4 pa = *reinterpret_cast<const float4*>(bufA + j);
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) {
        const int c = CM::c(lane, k);
        const float v0 = sm.V[(j + 0) * dhp + c], v1 = sm.V[(j + 1) * dhp + c];
        const float v2 = sm.V[(j + 2) * dhp + c], v3 = sm.V[(j + 3) * dhp + c];
        accR[k] = fmaf(pr.x, v0, accR[k]); accR[k] = fmaf(pr.y, v1, accR[k]);
        accR[k] = fmaf(pr.z, v2, accR[k]); accR[k] = fmaf(pr.w, v3, accR[k]);
        if (need_att) {
          accA[k] = fmaf(pa.x, v0, accA[k]); accA[k] = fmaf(pa.y, v1, accA[k]);
          accA[k] = fmaf(pa.z, v2, accA[k]); accA[k] = fmaf(pa.w, v3, accA[k]);
        }
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
template <int DH, int JPL>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_bwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int dhp = DH + 4;
  using CM = CMap<DH>;
  const int L = p.L, LP = (L + 3) & ~3;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ptr = smem_f;
  AttnSmem sm = carve_common(ptr, L, LP, DH, p.gate != nullptr);
  float* sDC = ptr; ptr += LP * dhp;      // d_ctx_cal head slice
  float* sDA = ptr; ptr += LP * dhp;      // d_ctx_att head slice
  float* matST = ptr; ptr += L * LP;      // dS^T   [j][i]  (zero above the diagonal)
  float* matS2T = ptr; ptr += L * LP;     // dS'^T
  float* matRT = ptr; ptr += L * LP;      // R_final^T
  float* matAT = ptr; ptr += L * LP;      // A^T
  float* rowbuf = ptr; ptr += kAttnWarps * 2 * LP;   // per-warp dS / dS' rows
  float* colDU = ptr; ptr += LP;
  float* colDT = ptr; ptr += LP;
  float* red = ptr; ptr += kAttnWarps * 4;
  const bool has_att = p.d_ctx_att != nullptr;

  stage_tile<DH>(sDC, p.d_ctx_cal, b, h, L, p.d);
  stage_tile<DH>(sDA, p.d_ctx_att, b, h, L, p.d);
  for (int e = threadIdx.x; e < L * LP; e += blockDim.x) { matST[e] = 0.f; matS2T[e] = 0.f; matRT[e] = 0.f; matAT[e] = 0.f; }
  for (int e = threadIdx.x; e < (LP - L) * dhp; e += blockDim.x) { sDC[L * dhp + e] = 0.f; sDA[L * dhp + e] = 0.f; }
  for (int j = threadIdx.x; j < LP; j += blockDim.x) { colDU[j] = 0.f; colDT[j] = 0.f; }
  stage_common<DH>(p, sm, b, h, LP);     // waits for all cp.async groups, ends with __syncthreads()

  const RowConst kc = make_consts<DH>(p, has_att);
  const float dpen = p.d_pen ? p.d_pen[0] : 0.f;
  const float sc2 = kc.sc * kc.sc;

  float accOq[CM::CPL], accDq[CM::CPL];
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) accOq[k] = accDq[k] = 0.f;
  float s_ob = 0.f, s_db = 0.f, s_scalar = 0.f, s_ratio = 0.f;
  float* bufS = rowbuf + (warp * 2 + 0) * LP;
  float* bufS2 = rowbuf + (warp * 2 + 1) * LP;

  RowP<JPL> r;
  for (int rnd = 0; rnd * kAttnWarps < L; ++rnd) {
    const int i = rnd * kAttnWarps + ((rnd & 1) ? (kAttnWarps - 1 - warp) : warp);
    if (i >= L) continue;
    row_forward<DH, JPL>(p, sm, kc, b, h, i, lane, has_att, r);
    int jr[JPL];
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) jr[jj] = (lane + 32 * jj) < L ? lane + 32 * jj : L - 1;
    // dRf_j = dctx_cal_i . v_j ; dA_j = dctx_att_i . v_j   (only column groups that touch the triangle)
    float dRf[JPL], dA[JPL];
    const float4* dci = reinterpret_cast<const float4*>(sDC + i * dhp);
    const float4* dai = reinterpret_cast<const float4*>(sDA + i * dhp);
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      dRf[jj] = dA[jj] = 0.f;
      if (!r.act[jj]) continue;
      const float4* vj = reinterpret_cast<const float4*>(sm.V + jr[jj] * dhp);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 v = vj[c4];
        a0 = dot4(dci[c4], v, a0);
        if (has_att) a1 = dot4(dai[c4], v, a1);
      }
      if (r.inb[jj]) { dRf[jj] = a0; dA[jj] = a1; }
    }
    float dO[JPL], dP[JPL], dM[JPL], dR[JPL], dC[JPL], tmp[JPL], dcm[JPL];
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      dO[jj] = 0.f; dP[jj] = 0.f; dM[jj] = 0.f;
      if (p.two_level) dR[jj] = dRf[jj];
      else {
        dR[jj] = dRf[jj] * kc.rr;
        dP[jj] = dRf[jj] * (1.0f - kc.rr);
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
    for (int jj = 0; jj < JPL; ++jj) { dS2[jj] *= kc.inv_sq; dz[jj] *= kc.inv_sq; dS[jj] = dz[jj]; }
    if (!p.two_level) {
      softmax_bwd_row<JPL>(r.P0soft, dP0, tmp);
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj) dS[jj] += tmp[jj] * kc.inv_sq;
    }
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj) {
      const int j = lane + 32 * jj;
      const bool tri_ok = r.inb[jj] && j <= i;       // everything above the diagonal is exactly zero
      if (j < LP) { bufS[j] = tri_ok ? dS[jj] : 0.f; bufS2[j] = tri_ok ? dS2[jj] : 0.f; }
      if (!tri_ok) continue;
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
        s_scalar += dz[jj] * (-(dl * dl) * kc.sc);
      }
      const int t = j * LP + i;
      matST[t] = dS[jj];
      matS2T[t] = dS2[jj];
      matRT[t] = r.Rf[jj];
      matAT[t] = r.A[jj];
    }
    row_du = warp_sum(row_du);
    row_dt = warp_sum(row_dt);
    s_ob += row_du;            // identical on every lane; lane 0 publishes it
    s_db += row_dt;
    __syncwarp();
    // row-side gradients: dq_i, dq'_i  (lane = channel, j <= i; float4 broadcast of the row buffers)
    float aq_[CM::CPL], aq2_[CM::CPL];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) aq_[k] = aq2_[k] = 0.f;
    int lo, hi;
    CM::slice(lane, 0, (i + 4) & ~3, lo, hi);
    for (int j = lo; j < hi; j += 4) {
      const float4 s1 = *reinterpret_cast<const float4*>(bufS + j);
      const float4 s2 = *reinterpret_cast<const float4*>(bufS2 + j);
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) {
        const int c = CM::c(lane, k);
        aq_[k] = fmaf(s1.x, sm.K[(j + 0) * dhp + c], aq_[k]); aq_[k] = fmaf(s1.y, sm.K[(j + 1) * dhp + c], aq_[k]);
        aq_[k] = fmaf(s1.z, sm.K[(j + 2) * dhp + c], aq_[k]); aq_[k] = fmaf(s1.w, sm.K[(j + 3) * dhp + c], aq_[k]);
        aq2_[k] = fmaf(s2.x, sm.K2[(j + 0) * dhp + c], aq2_[k]); aq2_[k] = fmaf(s2.y, sm.K2[(j + 1) * dhp + c], aq2_[k]);
        aq2_[k] = fmaf(s2.z, sm.K2[(j + 2) * dhp + c], aq2_[k]); aq2_[k] = fmaf(s2.w, sm.K2[(j + 3) * dhp + c], aq2_[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      const int c = CM::c(lane, k);
      float v1 = CM::reduce(aq_[k]), v2 = CM::reduce(aq2_[k]);
      const float qic = sm.Q[i * dhp + c];
      v1 += row_du * sm.wo[c] + row_dt * sm.wd[c];
      accOq[k] += row_du * qic;
      accDq[k] += row_dt * qic;
      if (CM::grp(lane) == 0) {
        const long long o = ((long long)b * L + i) * p.d + h * DH + c;
        p.d_mq[o] = v1;
        p.d_aq[o] = v2;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // column-side gradients: dk_j, dk'_j, dv_j  (warp per column, lane = channel, rows i >= j)
  float accOk[CM::CPL], accDk[CM::CPL];
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) accOk[k] = accDk[k] = 0.f;
  for (int rnd = 0; rnd * kAttnWarps < L; ++rnd) {
    const int j = rnd * kAttnWarps + ((rnd & 1) ? (kAttnWarps - 1 - warp) : warp);
    if (j >= L) continue;
    float ak_[CM::CPL], ak2_[CM::CPL], av_[CM::CPL];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) ak_[k] = ak2_[k] = av_[k] = 0.f;
    int lo, hi;
    CM::slice(lane, j & ~3, LP, lo, hi);
    for (int i = lo; i < hi; i += 4) {
      const float4 s1 = *reinterpret_cast<const float4*>(matST + j * LP + i);
      const float4 s2 = *reinterpret_cast<const float4*>(matS2T + j * LP + i);
      const float4 pr = *reinterpret_cast<const float4*>(matRT + j * LP + i);
      const float4 pa = *reinterpret_cast<const float4*>(matAT + j * LP + i);
      const float s1v[4] = {s1.x, s1.y, s1.z, s1.w}, s2v[4] = {s2.x, s2.y, s2.z, s2.w};
      const float prv[4] = {pr.x, pr.y, pr.z, pr.w}, pav[4] = {pa.x, pa.y, pa.z, pa.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < CM::CPL; ++k) {
          const int c = CM::c(lane, k);
          ak_[k] = fmaf(s1v[u], sm.Q[(i + u) * dhp + c], ak_[k]);
          ak2_[k] = fmaf(s2v[u], sm.Q2[(i + u) * dhp + c], ak2_[k]);
          av_[k] = fmaf(prv[u], sDC[(i + u) * dhp + c], av_[k]);
          if (has_att) av_[k] = fmaf(pav[u], sDA[(i + u) * dhp + c], av_[k]);
        }
      }
    }
    const float cdu = colDU[j], cdt = colDT[j];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      const int c = CM::c(lane, k);
      float v1 = CM::reduce(ak_[k]), v2 = CM::reduce(ak2_[k]), v3 = CM::reduce(av_[k]);
      const float kjc = sm.K[j * dhp + c];
      v1 += cdu * sm.wo[DH + c] + cdt * sm.wd[DH + c];
      accOk[k] += cdu * kjc;
      accDk[k] += cdt * kjc;
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

static size_t fwd_smem_bytes(int L, int dh, bool has_gate) {
  const int LP = (L + 3) & ~3;
  size_t f = common_floats(L, LP, dh, has_gate) + (size_t)kAttnWarps * 2 * LP;
  return f * sizeof(float) + kAttnWarps * sizeof(double);
}
static size_t bwd_smem_bytes(int L, int dh, bool has_gate) {
  const int LP = (L + 3) & ~3;
  size_t f = common_floats(L, LP, dh, has_gate) + (size_t)2 * LP * (dh + 4) + (size_t)4 * L * LP + (size_t)kAttnWarps * 2 * LP +
             2 * LP + kAttnWarps * 4;
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
  size_t smem = fwd_smem_bytes(p.L, DH, p.gate != nullptr);
  int rc = prep_kernel(attn_fwd_kernel<DH, JPL>, smem, "attn_calib_fwd");
  if (rc) return rc;
  attn_fwd_kernel<DH, JPL><<<p.B * p.H, kAttnThreads, smem, st>>>(p);
  return check_launch("attn_calib_fwd");
}
template <int DH, int JPL>
static int launch_bwd(const AttnParams& p, cudaStream_t st) {
  size_t smem = bwd_smem_bytes(p.L, DH, p.gate != nullptr);
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
  p.mq = mq; p.mk = mk; p.mv = mv; p.aq = aq; p.ak = ak; p.gate = combine_option == ACSR_ATTN_COMBINE_GATE ? gate_logit : nullptr;
  p.item_seq = item_seq;
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
  p.mq = mq; p.mk = mk; p.mv = mv; p.aq = aq; p.ak = ak; p.gate = combine_option == ACSR_ATTN_COMBINE_GATE ? gate_logit : nullptr;
  p.item_seq = item_seq;
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
