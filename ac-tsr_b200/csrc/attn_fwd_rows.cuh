// Row iterations of the fused calibrated attention forward, shared by attn_fwd.cu (L <= 64, one CTA holds all five
// tiles) and attn_long.cu (L <= 256, keys/values resident, query rows streamed).
#pragma once
#include "attn_common.cuh"

namespace acsr {

struct FwdCtx {
  float* rowbuf;     // this warp's row buffers
  float pen;
};

// one row group iteration: probabilities of row i, penalty, probs.V
template <int DH, int G, int NJ>
__device__ __forceinline__ void fwd_row_iter(const AttnParams& p, const AttnSmem& sm, const RowConst& kc, int b, int h, int i,
                                             bool rowok, int bound, int grp, int sub, bool need_att, int rstride, FwdCtx& cx) {
  constexpr int dhp = DH + 4;
  using CM = CMap<DH, G>;
  const int L = p.L;
  RowF<NJ> r;
  row_forward<DH, G, NJ>(p, sm, kc, b, h, i, bound, sub, need_att, r);
  float* bufR = cx.rowbuf + (grp * 2 + 0) * rstride;
  float* bufA = cx.rowbuf + (grp * 2 + 1) * rstride;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const int j = sub + G * jj;
    const bool a = (r.act >> jj) & 1u;
    const float Pj = r.Psoft[jj] * r.D1[jj];
    const float Rf = p.two_level ? r.R[jj] : (kc.rr * r.R[jj] + (1.0f - kc.rr) * Pj);
    const float Mj = r.Msoft[jj] * r.D3[jj];
    if (j < rstride) {                            // entries outside the row's range are written as 0
      bufR[j] = a ? Rf : 0.f;
      bufA[j] = a ? r.A[jj] : 0.f;
    }
    if (a && rowok) { const float om = 1.0f - Mj; cx.pen = fmaf(om, om, cx.pen); }
    if (p.prob_cal_out != nullptr && a && rowok) {
      const long long e = (((long long)b * p.H + h) * L + i) * L + j;
      p.prob_cal_out[e] = Rf;
      if (p.prob_att_out != nullptr && need_att) p.prob_att_out[e] = r.A[jj];
    }
    if (p.probs && a && rowok) {
      const long long e = (((long long)b * p.H + h) * L + i) * L + j;
      const long long plane = (long long)p.B * p.H * L * L;
      p.probs[0 * plane + e] = r.P0soft[jj] * r.D2[jj]; p.probs[1 * plane + e] = Pj;
      p.probs[2 * plane + e] = Mj; p.probs[3 * plane + e] = r.A[jj];
      p.probs[4 * plane + e] = r.C[jj]; p.probs[5 * plane + e] = Rf;
    }
  }
  if (rowok && sub == 0) cx.pen += (float)(L - bound);      // columns outside the range: M == 0 -> (1-M)^2 == 1
  __syncwarp();
  // ctx[i][c] = sum_{j<bound} prob[j] * V[j][c]
  float accR[CM::CPL], accA[CM::CPL];
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) accR[k] = accA[k] = 0.f;
  int lo, hi;
  CM::slice(sub, 0, (bound + 3) & ~3, lo, hi);
  const int c0 = CM::c0(sub);
  for (int j = lo; j < hi; j += 4) {
    const float4 pr = *reinterpret_cast<const float4*>(bufR + j);
    const float4 pa = *reinterpret_cast<const float4*>(bufA + j);
    const float prv[4] = {pr.x, pr.y, pr.z, pr.w}, pav[4] = {pa.x, pa.y, pa.z, pa.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[CM::CPL];
      VecLd<CM::CPL>::ld(sm.V + (j + u) * dhp + c0, v);
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) {
        accR[k] = fmaf(prv[u], v[k], accR[k]);
        if (need_att) accA[k] = fmaf(pav[u], v[k], accA[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) { accR[k] = CM::reduce(accR[k]); accA[k] = CM::reduce(accA[k]); }
  if (CM::split(sub) == 0 && rowok) {
    const long long o = ((long long)b * L + i) * p.d + h * DH + c0;
    VecLd<CM::CPL>::st(p.ctx_cal + o, accR);
    if (need_att) VecLd<CM::CPL>::st(p.ctx_att + o, accA);
  }
  __syncwarp();
}

// row group iteration for rows whose context nobody reads: attack mask -> penalty only
template <int DH, int G, int NJ>
__device__ __forceinline__ void fwd_row_iter_m(const AttnParams& p, const AttnSmem& sm, const RowConst& kc, int b, int h, int i,
                                               bool rowok, int bound, int sub, FwdCtx& cx) {
  float Msoft[NJ], D3[NJ];
  unsigned act;
  row_forward_m<DH, G, NJ>(p, sm, kc, b, h, i, bound, sub, Msoft, D3, act);
  if (!rowok) return;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj)
    if ((act >> jj) & 1u) { const float om = 1.0f - Msoft[jj] * D3[jj]; cx.pen = fmaf(om, om, cx.pen); }
  if (sub == 0) cx.pen += (float)(p.L - bound);
}

}  // namespace acsr
