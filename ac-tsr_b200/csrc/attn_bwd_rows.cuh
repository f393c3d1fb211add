// Row iterations of the fused calibrated attention backward, shared by attn_bwd.cu (L <= 64: dS, dS', R, A kept transposed
// in shared memory) and attn_long.cu (L <= 256: the same matrices go row-major to a global scratch, template LONG).
#pragma once
#include "attn_common.cuh"

namespace acsr {

template <int NS>
struct BwdSmem {
  float *sT0, *sT1;                  // cotangent tiles [LP][dh+4]
  float *matST[NS], *matS2T[NS];     // dS^T, dS'^T  [j][i] packed lower triangular
  float *matRT, *matAT;              // R_final^T, A^T
  float *rowbuf;                     // per warp: dS / dS' rows of every stream, per row group
  float *colDU, *colDT;              // [NS][LP]
  float *pacc;                       // [4*dh] CTA partials of d_ow / d_dw
  float *red;                        // [warps][4]
  // LONG variant (attn_long.cu): the four matrices live row-major [i][j] in a global scratch (pre-zeroed), this (b,h)'s slice
  float *gS[NS], *gS2[NS], *gR, *gA;
};

struct BwdAcc {                      // lane partials of the scalar parameter gradients (stream 0)
  float s_ob, s_db, s_scalar, s_ratio;
};

struct BwdFlags {
  bool has_t0, has_t1, t1_att, has_att, gate;
  float sc2;
};

// one row group iteration: recompute row i, run the chain of every stream, finish dq_i, dq'_i
template <int DH, int G, int NJ, int NS, bool LONG = false>
__device__ __forceinline__ void bwd_row_iter(const AttnParams& p, const AttnSmem& sm, const BwdSmem<NS>& bs, const RowConst& kc,
                                             const BwdFlags& f, const float* dpen, int b, int h, int i, bool rowok, int bound,
                                             int grp, int sub, int rstride, float* wbuf, BwdAcc& acc, float* accOq,
                                             float* accDq) {
  constexpr int dhp = DH + 4;
  using CM = CMap<DH, G>;
  const int L = p.L, LP = (L + 3) & ~3;
  RowF<NJ> r;
  row_forward<DH, G, NJ>(p, sm, kc, b, h, i, bound, sub, f.has_att, r);
  const unsigned act = r.act;
  // cotangent dots with the value rows: d0_j = t0_i . v_j ; d1_j = t1_i . v_j
  float d0[NJ], d1[NJ];
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) d0[jj] = d1[jj] = 0.f;
  if (f.has_t0 || f.has_t1) {
    const float4* a0 = reinterpret_cast<const float4*>(bs.sT0 + i * dhp);
    const float4* a1 = reinterpret_cast<const float4*>(bs.sT1 + i * dhp);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int jcl = ((act >> jj) & 1u) ? sub + G * jj : 0;
      const float4* vj = reinterpret_cast<const float4*>(sm.V + jcl * dhp);
      float x0 = 0.f, x1 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 v = vj[c4];
        if (f.has_t0) x0 = dot4(a0[c4], v, x0);
        if (f.has_t1) x1 = dot4(a1[c4], v, x1);
      }
      if ((act >> jj) & 1u) { d0[jj] = x0; d1[jj] = x1; }
    }
  }
  if (p.dprob_cal != nullptr || p.dprob_att != nullptr) {       // cotangents of the attention matrices themselves (ACTiSASRec)
    const long long ebase = (((long long)b * p.H + h) * L + i) * L;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj)
      if ((act >> jj) & 1u) {
        const int j = sub + G * jj;
        if (p.dprob_cal != nullptr) d0[jj] += __ldg(p.dprob_cal + ebase + j);
        if (p.dprob_att != nullptr) d1[jj] += __ldg(p.dprob_att + ebase + j);
      }
  }
  // forward quantities shared by the streams
  float Pj[NJ], Mj[NJ], Oj[NJ], expm[NJ];
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const bool a = (act >> jj) & 1u;
    Pj[jj] = r.Psoft[jj] * r.D1[jj];
    Mj[jj] = r.Msoft[jj] * r.D3[jj];
    Oj[jj] = p.two_level ? Pj[jj] : r.P0soft[jj] * r.D2[jj];
    expm[jj] = a ? fexp(1.0f - Mj[jj]) : 0.f;
    if (a && rowok) {
      const int j = sub + G * jj;
      const float rfin = p.two_level ? r.R[jj] : (kc.rr * r.R[jj] + (1.0f - kc.rr) * Pj[jj]);
      if (LONG) {
        bs.gR[(long long)i * L + j] = rfin;
        bs.gA[(long long)i * L + j] = r.A[jj];
      } else {
        const int t = mat_row(p, j, LP) + i;
        bs.matRT[t] = rfin;
        bs.matAT[t] = r.A[jj];
      }
    }
  }
  float* gbuf = wbuf + grp * 2 * NS * rstride;     // [s][2][rstride]
  bool live[NS];
  float row_du[NS], row_dt[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    float* bufS = gbuf + (2 * s + 0) * rstride;
    float* bufS2 = gbuf + (2 * s + 1) * rstride;
    row_du[s] = row_dt[s] = 0.f;
    // which cotangents feed this stream
    const float* dRf = (NS == 1 || s == 0) ? d0 : d1;
    const float* dA = d1;
    bool useR = (NS == 1 || s == 0) ? f.has_t0 : (f.has_t1 && !f.t1_att);
    bool useA = (NS == 1) ? f.has_t1 : (s == 1 && f.has_t1 && f.t1_att);
    bool nzR = false, nzA = false;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) { nzR |= dRf[jj] != 0.f; nzA |= dA[jj] != 0.f; }
    useR = useR && __any_sync(kFull, nzR && rowok);       // exact-zero cotangent rows skip the chain (warp-uniform)
    useA = useA && __any_sync(kFull, nzA && rowok);
    const float dp = dpen[s];
    live[s] = useR || useA || dp != 0.f;
    if (!live[s]) {
      if (LONG && rowok) {                  // the workspace is not pre-zeroed: every entry in play is written
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj)
          if ((act >> jj) & 1u) {
            const long long e = (long long)i * L + sub + G * jj;
            bs.gS[s][e] = 0.f;
            bs.gS2[s][e] = 0.f;
          }
      }
      continue;
    }
    const bool owner = s == 0;
    float dO[NJ], dP[NJ], dM[NJ], tmp[NJ];
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) { dO[jj] = 0.f; dP[jj] = 0.f; dM[jj] = 0.f; }
    if (useR) {
      float dR[NJ], dcm[NJ], dC[NJ];
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        if (p.two_level) dR[jj] = dRf[jj];
        else {
          dR[jj] = dRf[jj] * kc.rr;
          dP[jj] = dRf[jj] * (1.0f - kc.rr);
          if (owner && rowok) acc.s_ratio += dRf[jj] * (r.R[jj] - Pj[jj]);
        }
      }
      if (p.plain) {                                   // transformer_layers.py: no re-normalising softmax to go through
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) dcm[jj] = dR[jj];
      } else softmax_bwd_row<G, NJ>(r.R, dR, dcm);     // grad wrt (comb + mask)
      if (p.combine == ACSR_ATTN_COMBINE_FIXED) {
        softmax_bwd_row<G, NJ>(r.F, dcm, tmp);          // grad wrt (O + 0.5 C); columns outside the range carry no cotangent
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) { dO[jj] += tmp[jj]; dC[jj] = 0.5f * tmp[jj]; }
      } else {
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const float g = r.g[jj];
          dO[jj] += dcm[jj] * g;
          dC[jj] = dcm[jj] * (1.0f - g);
          if (f.gate && ((act >> jj) & 1u) && rowok) {
            const float dgl = dcm[jj] * (Oj[jj] - r.C[jj]) * g * (1.0f - g);
            if (dgl != 0.f) atomicAdd(p.d_gate + s * p.s1_ll + ((long long)b * L + i) * L + sub + G * jj, dgl);
          }
        }
      }
      if (p.plain) {
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) tmp[jj] = dC[jj];
      } else softmax_bwd_row<G, NJ>(r.C, dC, tmp);      // grad wrt (O*expm + mask)
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        dO[jj] += tmp[jj] * expm[jj];
        dM[jj] -= tmp[jj] * Oj[jj] * expm[jj];
      }
    }
    if (useA) {
      if (p.plain) {
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) tmp[jj] = dA[jj];
      } else softmax_bwd_row<G, NJ>(r.A, dA, tmp);      // grad wrt (O*M + n(1-M) + mask)
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        dO[jj] += tmp[jj] * Mj[jj];
        dM[jj] += tmp[jj] * (Oj[jj] - r.nz[jj]);
      }
    }
    float dS2[NJ], dS[NJ], dz[NJ];
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      if ((act >> jj) & 1u) dM[jj] += dp * (-2.0f) * (1.0f - Mj[jj]);
      dM[jj] *= r.D3[jj];
    }
    softmax_bwd_row<G, NJ>(r.Msoft, dM, dS2);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) { dS2[jj] *= kc.inv_sq; dz[jj] = 0.f; dS[jj] = 0.f; }
    if (useR || useA) {
      float dP0[NJ];
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        if (p.two_level) { dP[jj] += dO[jj]; dP0[jj] = 0.f; }
        else dP0[jj] = dO[jj] * r.D2[jj];
        dP[jj] *= r.D1[jj];
      }
      softmax_bwd_row<G, NJ>(r.Psoft, dP, dz);
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) { dz[jj] *= kc.inv_sq; dS[jj] = dz[jj]; }
      if (!p.two_level) {
        softmax_bwd_row<G, NJ>(r.P0soft, dP0, tmp);
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) dS[jj] += tmp[jj] * kc.inv_sq;
      }
    }
    float rdu = 0.f, rdt = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = sub + G * jj;
      const bool a = ((act >> jj) & 1u) && rowok;       // columns outside the range are exactly zero
      if (j < rstride) { bufS[j] = a ? dS[jj] : 0.f; bufS2[j] = a ? dS2[jj] : 0.f; }
      if (!a) continue;
      if (p.ow) {
        const float sg = r.sig[jj];
        // layers.py:719: d/du log(1 - sg + eps) for j <= i; d/du log(sg + eps) for j > i (only in play under the bidirectional mask)
        const float de = (j > i) ? sg * (1.0f - sg) * frcp(sg + kOrderEps) : -sg * (1.0f - sg) * frcp((1.0f - sg) + kOrderEps);
        const float du = dz[jj] * de;
        rdu += du;
        if (du != 0.f) atomicAdd(bs.colDU + s * LP + j, du);
      }
      if (p.dw) {
        const float dl = r.delta[jj];
        const float dt = dz[jj] * dl * f.sc2;
        rdt += dt;
        if (dt != 0.f) atomicAdd(bs.colDT + s * LP + j, dt);
        if (owner) acc.s_scalar += dz[jj] * (-(dl * dl) * kc.sc);
      }
      if (LONG) {
        bs.gS[s][(long long)i * L + j] = dS[jj];
        bs.gS2[s][(long long)i * L + j] = dS2[jj];
      } else {
        const int t = mat_row(p, j, LP) + i;
        bs.matST[s][t] = dS[jj];
        bs.matS2T[s][t] = dS2[jj];
        if (NS == 1 && p.d_s_bias != nullptr) p.d_s_bias[(((long long)b * p.H + h) * L + i) * L + j] = dS[jj];
      }
    }
    row_du[s] = grp_sum<G>(rdu);
    row_dt[s] = grp_sum<G>(rdt);
  }
  if (sub == 0) { acc.s_ob += row_du[0]; acc.s_db += row_dt[0]; }
  __syncwarp();
  // row-side gradients: dq_i, dq'_i of every stream (lane = channel, j < bound; float4 broadcast of the row buffers)
  float aq_[NS][CM::CPL], aq2_[NS][CM::CPL];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) aq_[s][k] = aq2_[s][k] = 0.f;
  int lo, hi;
  CM::slice(sub, 0, (bound + 3) & ~3, lo, hi);
  const int c0 = CM::c0(sub);
  bool any_live = false;
#pragma unroll
  for (int s = 0; s < NS; ++s) any_live |= live[s];
  if (any_live) {
    for (int j = lo; j < hi; j += 4) {
      float s1v[NS][4], s2v[NS][4];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float4 s1 = *reinterpret_cast<const float4*>(gbuf + (2 * s + 0) * rstride + j);
        const float4 s2 = *reinterpret_cast<const float4*>(gbuf + (2 * s + 1) * rstride + j);
        s1v[s][0] = s1.x; s1v[s][1] = s1.y; s1v[s][2] = s1.z; s1v[s][3] = s1.w;
        s2v[s][0] = s2.x; s2v[s][1] = s2.y; s2v[s][2] = s2.z; s2v[s][3] = s2.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float kv[CM::CPL], k2v[CM::CPL];
        VecLd<CM::CPL>::ld(sm.K + (j + u) * dhp + c0, kv);
        VecLd<CM::CPL>::ld(sm.K2 + (j + u) * dhp + c0, k2v);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          if (!live[s]) continue;
#pragma unroll
          for (int k = 0; k < CM::CPL; ++k) {
            aq_[s][k] = fmaf(s1v[s][u], kv[k], aq_[s][k]);
            aq2_[s][k] = fmaf(s2v[s][u], k2v[k], aq2_[s][k]);
          }
        }
      }
    }
  }
  {
    float qv[CM::CPL];
    VecLd<CM::CPL>::ld(sm.Q + i * dhp + c0, qv);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      float o1[CM::CPL], o2[CM::CPL];
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) {
        o1[k] = CM::reduce(aq_[s][k]) + row_du[s] * sm.wo[c0 + k] + row_dt[s] * sm.wd[c0 + k];
        o2[k] = CM::reduce(aq2_[s][k]);
      }
      if (CM::split(sub) == 0 && rowok) {
        const long long o = s * p.s1_td + ((long long)b * L + i) * p.d + h * DH + c0;
        VecLd<CM::CPL>::st(p.d_mq + o, o1);
        VecLd<CM::CPL>::st(p.d_aq + o, o2);
      }
    }
    if (CM::split(sub) == 0 && rowok) {
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) { accOq[k] = fmaf(row_du[0], qv[k], accOq[k]); accDq[k] = fmaf(row_dt[0], qv[k], accDq[k]); }
    }
  }
  __syncwarp();
}

// row group iteration for rows whose context carries no cotangent by contract (ctx_rows): only the penalty reaches them,
// through the attack mask: dS' of the streams with a penalty cotangent, dq'_i; everything else of the row is zero
template <int DH, int G, int NJ, int NS, bool LONG = false>
__device__ __forceinline__ void bwd_row_iter_m(const AttnParams& p, const AttnSmem& sm, const BwdSmem<NS>& bs, const RowConst& kc,
                                               const float* dpen, int b, int h, int i, bool rowok, int bound, int grp, int sub,
                                               int rstride, float* wbuf) {
  constexpr int dhp = DH + 4;
  using CM = CMap<DH, G>;
  const int L = p.L, LP = (L + 3) & ~3;
  float Msoft[NJ], D3[NJ];
  unsigned act;
  row_forward_m<DH, G, NJ>(p, sm, kc, b, h, i, bound, sub, Msoft, D3, act);
  float* gbuf = wbuf + grp * 2 * NS * rstride;
  const int c0 = CM::c0(sub);
  if (LONG && rowok) {                      // the workspace is not pre-zeroed: this row's dS, R, A (and dS' without a penalty) are zero
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj)
      if ((act >> jj) & 1u) {
        const long long e = (long long)i * L + sub + G * jj;
        bs.gR[e] = 0.f;
        bs.gA[e] = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          bs.gS[s][e] = 0.f;
          if (dpen[s] == 0.f) bs.gS2[s][e] = 0.f;
        }
      }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const float dp = dpen[s];
    float o2[CM::CPL];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) o2[k] = 0.f;
    if (dp != 0.f) {
      float dM[NJ], dS2[NJ];
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) dM[jj] = ((act >> jj) & 1u) ? dp * (-2.0f) * (1.0f - Msoft[jj] * D3[jj]) * D3[jj] : 0.f;
      softmax_bwd_row<G, NJ>(Msoft, dM, dS2);
      float* bufS2 = gbuf + (2 * s + 1) * rstride;
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        const int j = sub + G * jj;
        const bool a = ((act >> jj) & 1u) && rowok;
        const float v = dS2[jj] * kc.inv_sq;
        if (j < rstride) bufS2[j] = a ? v : 0.f;
        if (a) {
          if (LONG) bs.gS2[s][(long long)i * L + j] = v;
          else bs.matS2T[s][mat_row(p, j, LP) + i] = v;
        }
      }
      __syncwarp();
      int lo, hi;
      CM::slice(sub, 0, (bound + 3) & ~3, lo, hi);
      for (int j = lo; j < hi; j += 4) {
        const float4 s2 = *reinterpret_cast<const float4*>(bufS2 + j);
        const float s2v[4] = {s2.x, s2.y, s2.z, s2.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float k2v[CM::CPL];
          VecLd<CM::CPL>::ld(sm.K2 + (j + u) * dhp + c0, k2v);
#pragma unroll
          for (int k = 0; k < CM::CPL; ++k) o2[k] = fmaf(s2v[u], k2v[k], o2[k]);
        }
      }
      __syncwarp();
    }
    float o1[CM::CPL];
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) { o2[k] = CM::reduce(o2[k]); o1[k] = 0.f; }
    if (CM::split(sub) == 0 && rowok) {
      const long long o = s * p.s1_td + ((long long)b * L + i) * p.d + h * DH + c0;
      VecLd<CM::CPL>::st(p.d_mq + o, o1);
      VecLd<CM::CPL>::st(p.d_aq + o, o2);
    }
  }
}

}  // namespace acsr
