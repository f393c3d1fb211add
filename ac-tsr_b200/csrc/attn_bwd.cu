// Fused calibrated causal attention, backward (see attn_common.cuh for the design).
//
// The backward recomputes a row's probabilities once (same tiles, same Philox counters) and then
// runs the (linear) backward chain for up to TWO cotangent streams -- the reference's two backward()
// traversals, trainer.py:672-684 -- finishing the row-side gradients (dQ, dQ') from per-group row
// buffers.  dS, dS' (per stream) and R, A are kept TRANSPOSED in shared memory, packed lower
// triangular, and the column-side gradients (dK, dK', dV) are finished in a second, column-parallel
// phase with float4 broadcast reads.  A row whose cotangents are exactly zero (most rows of the
// calibrated-loss stream: only position len-1 feeds the loss) skips the chain.
#include "attn_bwd_rows.cuh"

namespace acsr {

// row phase of the (b,h) tile with G-lane row groups
template <int DH, int G, int MAXNJ, int NS>
__device__ __forceinline__ void bwd_body(const AttnParams& p, const AttnSmem& sm, const BwdSmem<NS>& bs, const RowConst& kc,
                                         const BwdFlags& f, const float* dpen, int b, int h, int nkey, int rstride, BwdAcc& acc) {
  constexpr int dhp = DH + 4;
  constexpr int RPW = 32 / G;
  using CM = CMap<DH, G>;
  const int L = p.L, LP = (L + 3) & ~3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / G, sub = lane % G;
  const int c0 = CM::c0(sub);
  float* wbuf = bs.rowbuf + warp * rowbuf_floats_per_warp(LP, 2 * NS);
  const int rt = p.ctx_rows ? (int)p.ctx_rows[b] - 1 : -1;     // the only row with a context cotangent, or -1: all rows
  float accOq[CM::CPL], accDq[CM::CPL];
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) accOq[k] = accDq[k] = 0.f;
  for (int t0 = next_task(sm.misc + 2, RPW); t0 < L; t0 = next_task(sm.misc + 2, RPW)) {      // heaviest rows first
    const int iw = L - 1 - t0;
    const int iraw = iw - grp;
    const bool rowok = iraw >= 0;
    const int i = rowok ? iraw : 0;
    const int bound = p.full ? nkey : min(i + 1, nkey);
    const int nj = ((p.full ? nkey : min(iw + 1, nkey)) + G - 1) / G;
    if (rt >= 0 && (rt > iw || rt <= iw - RPW)) {       // rows without a context cotangent: penalty path only
#define ACSR_BWD_ROW_M(NJV) bwd_row_iter_m<DH, G, NJV, NS>(p, sm, bs, kc, dpen, b, h, i, rowok, bound, grp, sub, rstride, wbuf)
      if (MAXNJ == 1 || nj == 1) ACSR_BWD_ROW_M(1);
      else if (nj == 2) ACSR_BWD_ROW_M((MAXNJ >= 2 ? 2 : 1));
      else ACSR_BWD_ROW_M((MAXNJ >= 4 ? 4 : 1));
#undef ACSR_BWD_ROW_M
      continue;
    }
#define ACSR_BWD_ROW(NJV) \
  bwd_row_iter<DH, G, NJV, NS>(p, sm, bs, kc, f, dpen, b, h, i, rowok, bound, grp, sub, rstride, wbuf, acc, accOq, accDq)
    if (MAXNJ == 1 || nj == 1) ACSR_BWD_ROW(1);
    else if (nj == 2) ACSR_BWD_ROW((MAXNJ >= 2 ? 2 : 1));
    else ACSR_BWD_ROW((MAXNJ >= 4 ? 4 : 1));
#undef ACSR_BWD_ROW
  }
  // lane partials of the query-side halves of d_ow / d_dw -> CTA partials in smem
  if ((p.d_ow || p.d_dw) && CM::split(sub) == 0) {
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      const int c = c0 + k;
      if (accOq[k] != 0.f) atomicAdd(bs.pacc + c, accOq[k]);
      if (accDq[k] != 0.f) atomicAdd(bs.pacc + 2 * DH + c, accDq[k]);
    }
  }
}

// column phase of the (b,h) tile: dk_j, dk'_j, dv_j.  CG lanes per column (lane = channel, rows i >= j); short sequences
// use whole warps per column so that every warp has a column to work on
template <int DH, int CG, int NS>
__device__ __forceinline__ void bwd_cols(const AttnParams& p, const AttnSmem& sm, const BwdSmem<NS>& bs, const BwdFlags& f, int b, int h,
                                         int nkey) {
  constexpr int dhp = DH + 4;
  constexpr int RPW = 32 / CG;
  using CM = CMap<DH, CG>;
  const int L = p.L, LP = (L + 3) & ~3;
  const int lane = threadIdx.x & 31;
  const int grp = lane / CG, sub = lane % CG;
  const int c0 = CM::c0(sub);
  float accOk[CM::CPL], accDk[CM::CPL];
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) accOk[k] = accDk[k] = 0.f;
  // column-side gradients: dk_j, dk'_j, dv_j  (group per column, lane = channel, rows i >= j)
  for (int t0 = next_task(sm.misc + 3, RPW); t0 < nkey; t0 = next_task(sm.misc + 3, RPW)) {    // heaviest columns (small j) first
    const int jraw = t0 + grp;
    const bool colok = jraw < nkey;
    const int j = colok ? jraw : 0;
    float ak_[NS][CM::CPL], ak2_[NS][CM::CPL], av_[NS][CM::CPL];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) ak_[s][k] = ak2_[s][k] = av_[s][k] = 0.f;
    int lo, hi;
    CM::slice(sub, p.full ? 0 : (j & ~3), LP, lo, hi);      // rows i that see column j
    const int off = mat_row(p, j, LP);
    for (int i = lo; i < hi; i += 4) {
      float s1v[NS][4], s2v[NS][4];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float4 s1 = *reinterpret_cast<const float4*>(bs.matST[s] + off + i);
        const float4 s2 = *reinterpret_cast<const float4*>(bs.matS2T[s] + off + i);
        s1v[s][0] = s1.x; s1v[s][1] = s1.y; s1v[s][2] = s1.z; s1v[s][3] = s1.w;
        s2v[s][0] = s2.x; s2v[s][1] = s2.y; s2v[s][2] = s2.z; s2v[s][3] = s2.w;
      }
      const float4 pr = *reinterpret_cast<const float4*>(bs.matRT + off + i);
      const float4 pa = *reinterpret_cast<const float4*>(bs.matAT + off + i);
      const float prv[4] = {pr.x, pr.y, pr.z, pr.w}, pav[4] = {pa.x, pa.y, pa.z, pa.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float qv[CM::CPL], q2v[CM::CPL], t0v[CM::CPL], t1v[CM::CPL];
        VecLd<CM::CPL>::ld(sm.Q + (i + u) * dhp + c0, qv);
        VecLd<CM::CPL>::ld(sm.Q2 + (i + u) * dhp + c0, q2v);
        VecLd<CM::CPL>::ld(bs.sT0 + (i + u) * dhp + c0, t0v);
        VecLd<CM::CPL>::ld(bs.sT1 + (i + u) * dhp + c0, t1v);
#pragma unroll
        for (int k = 0; k < CM::CPL; ++k) {
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            ak_[s][k] = fmaf(s1v[s][u], qv[k], ak_[s][k]);
            ak2_[s][k] = fmaf(s2v[s][u], q2v[k], ak2_[s][k]);
          }
          av_[0][k] = fmaf(prv[u], t0v[k], av_[0][k]);
          av_[NS - 1][k] = fmaf(f.t1_att ? pav[u] : prv[u], t1v[k], av_[NS - 1][k]);
        }
      }
    }
    float kv[CM::CPL];
    VecLd<CM::CPL>::ld(sm.K + j * dhp + c0, kv);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float cdu = bs.colDU[s * LP + j], cdt = bs.colDT[s * LP + j];
      float o1[CM::CPL], o2[CM::CPL], o3[CM::CPL];
#pragma unroll
      for (int k = 0; k < CM::CPL; ++k) {
        o1[k] = CM::reduce(ak_[s][k]) + cdu * sm.wo[DH + c0 + k] + cdt * sm.wd[DH + c0 + k];
        o2[k] = CM::reduce(ak2_[s][k]);
        o3[k] = CM::reduce(av_[s][k]);
      }
      if (CM::split(sub) == 0 && colok) {
        const long long o = s * p.s1_td + ((long long)b * L + j) * p.d + h * DH + c0;
        VecLd<CM::CPL>::st(p.d_mk + o, o1);
        VecLd<CM::CPL>::st(p.d_ak + o, o2);
        VecLd<CM::CPL>::st(p.d_mv + o, o3);
        if (s == 0) {
#pragma unroll
          for (int k = 0; k < CM::CPL; ++k) { accOk[k] = fmaf(cdu, kv[k], accOk[k]); accDk[k] = fmaf(cdt, kv[k], accDk[k]); }
        }
      }
    }
  }
  if ((p.d_ow || p.d_dw) && CM::split(sub) == 0) {
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      const int c = c0 + k;
      if (accOk[k] != 0.f) atomicAdd(bs.pacc + DH + c, accOk[k]);
      if (accDk[k] != 0.f) atomicAdd(bs.pacc + 3 * DH + c, accDk[k]);
    }
  }
}

template <int DH, int NS>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_bwd_kernel(const AttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem_f[];
  constexpr int dhp = DH + 4;
  const int L = p.L, LP = (L + 3) & ~3;
  const int TRI = mat_floats(p.full, L, LP);
  const int b = p.order ? p.order[blockIdx.x / p.H] : (int)(blockIdx.x / p.H), h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ptr = smem_f;
  AttnSmem sm = carve_common(ptr, LP, DH);
  BwdSmem<NS> bs;
  bs.sT0 = ptr; ptr += LP * dhp;
  bs.sT1 = ptr; ptr += LP * dhp;
#pragma unroll
  for (int s = 0; s < NS; ++s) { bs.matST[s] = ptr; ptr += TRI; bs.matS2T[s] = ptr; ptr += TRI; }
  bs.matRT = ptr; ptr += TRI;
  bs.matAT = ptr; ptr += TRI;
  bs.rowbuf = ptr; ptr += kAttnWarps * rowbuf_floats_per_warp(LP, 2 * NS);
  bs.colDU = ptr; ptr += NS * LP;
  bs.colDT = ptr; ptr += NS * LP;
  bs.pacc = ptr; ptr += 4 * DH;
  bs.red = ptr; ptr += kAttnWarps * 4;
  BwdFlags f;
  f.has_t0 = p.t0 != nullptr || p.dprob_cal != nullptr; f.has_t1 = p.t1 != nullptr || p.dprob_att != nullptr;
  f.t1_att = NS == 1 ? true : (p.t1_is_att != 0);
  f.has_att = f.has_t1 && f.t1_att;          // the attacked probabilities A are needed
  f.gate = p.combine == ACSR_ATTN_COMBINE_GATE;

  stage_tile<DH>(bs.sT0, p.t0, b, h, L, p.d, L, LP);
  stage_tile<DH>(bs.sT1, p.t1, b, h, L, p.d, L, LP);
  {
    const int ntri = (2 * NS + 2) * TRI;       // the packed matrices are contiguous
    float4* z4 = reinterpret_cast<float4*>(bs.matST[0]);
    for (int e = threadIdx.x; e < ntri / 4; e += blockDim.x) z4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int j = threadIdx.x; j < NS * LP; j += blockDim.x) { bs.colDU[j] = 0.f; bs.colDT[j] = 0.f; }
  for (int j = threadIdx.x; j < 4 * DH; j += blockDim.x) bs.pacc[j] = 0.f;
  const int nkey = stage_common<DH>(p, sm, b, h, LP);     // waits for all cp.async groups, ends with __syncthreads()

  const RowConst kc = make_consts<DH>(p, f.has_att);
  f.sc2 = kc.sc * kc.sc;
  float dpen[NS];
  dpen[0] = p.d_pen0 ? p.d_pen0[0] : 0.f;
  if (NS == 2) dpen[NS - 1] = p.d_pen1 ? p.d_pen1[0] : 0.f;
  BwdAcc acc;
  acc.s_ob = acc.s_db = acc.s_scalar = acc.s_ratio = 0.f;

  if (nkey <= 8) bwd_body<DH, 8, 1, NS>(p, sm, bs, kc, f, dpen, b, h, nkey, 8, acc);
  else bwd_body<DH, 16, 4, NS>(p, sm, bs, kc, f, dpen, b, h, nkey, LP, acc);
  __syncthreads();
  if (nkey <= 8) bwd_cols<DH, 32, NS>(p, sm, bs, f, b, h, nkey);
  else bwd_cols<DH, 16, NS>(p, sm, bs, f, b, h, nkey);

  // key positions behind the last real item receive no gradient
  for (int e = threadIdx.x; e < (L - nkey) * (DH / 4); e += blockDim.x) {
    const int j = nkey + e / (DH / 4), c4 = e % (DH / 4);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const long long o = s * p.s1_td + ((long long)b * L + j) * p.d + h * DH + c4 * 4;
      *reinterpret_cast<float4*>(p.d_mk + o) = z;
      *reinterpret_cast<float4*>(p.d_ak + o) = z;
      *reinterpret_cast<float4*>(p.d_mv + o) = z;
    }
  }
  // parameter gradients (stream 0 owns them): one atomic per address per CTA
  const float s_ob = warp_sum(acc.s_ob), s_db = warp_sum(acc.s_db);
  const float s_scalar = warp_sum(acc.s_scalar), s_ratio = warp_sum(acc.s_ratio);
  if (lane == 0) {
    bs.red[warp * 4 + 0] = s_ob; bs.red[warp * 4 + 1] = s_db; bs.red[warp * 4 + 2] = s_scalar; bs.red[warp * 4 + 3] = s_ratio;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < kAttnWarps; ++w) s += bs.red[w * 4 + threadIdx.x];
    float* dst = threadIdx.x == 0 ? p.d_ob : threadIdx.x == 1 ? p.d_db : threadIdx.x == 2 ? p.d_scalar : p.d_ratio;
    if (dst != nullptr && s != 0.f) atomicAdd(dst, s);
  }
  for (int c = threadIdx.x; c < 4 * DH; c += blockDim.x) {
    const float v = bs.pacc[c];
    float* dst = c < 2 * DH ? p.d_ow : p.d_dw;
    if (dst != nullptr && v != 0.f) atomicAdd(dst + (c < 2 * DH ? c : c - 2 * DH), v);
  }
}

static size_t bwd_smem_bytes(int L, int dh, int ns, int bidir) {
  const int LP = (L + 3) & ~3;
  size_t f = common_floats(LP, dh) + (size_t)2 * LP * (dh + 4) + (size_t)(2 * ns + 2) * mat_floats(bidir, L, LP) +
             (size_t)kAttnWarps * rowbuf_floats_per_warp(LP, 2 * ns) + 2 * ns * LP + 4 * dh + kAttnWarps * 4;
  return f * sizeof(float);
}

template <int DH, int NS>
static int launch_bwd(const AttnParams& p, cudaStream_t st) {
  size_t smem = bwd_smem_bytes(p.L, DH, NS, p.full);
  int rc = prep_kernel(attn_bwd_kernel<DH, NS>, smem, "attn_calib_bwd");
  if (rc) return rc;
  launch_pdl(attn_bwd_kernel<DH, NS>, dim3(p.B * p.H), dim3(kAttnThreads), smem, st, p);
  return check_launch("attn_calib_bwd");
}

template <int NS>
static int dispatch_bwd(const AttnParams& p, cudaStream_t st) {
  switch (p.dh) {
    case 8: return launch_bwd<8, NS>(p, st);
    case 16: return launch_bwd<16, NS>(p, st);
    case 32: return launch_bwd<32, NS>(p, st);
    case 64: return launch_bwd<64, NS>(p, st);
  }
  return ACSR_ERR_UNSUPPORTED;
}

}  // namespace acsr

using namespace acsr;

extern "C" {

int acsr_attn_calib_bwd(const float* d_ctx_att, const float* d_ctx_cal, const float* d_pen_sq, const float* mq, const float* mk,
                        const float* mv, const float* aq, const float* ak, const float* gate_logit, const int64_t* item_seq,
                        const float* order_w, const float* order_b, const float* dist_w, const float* dist_b, const float* scalar,
                        int B, int L, int H, int dh, int two_level, int combine_option, float comb_scalar, int rich_mode,
                        const float* rich_ratio, float p_attn, const float* D1, const float* D2, const float* D3,
                        const float* noise, const void* rng, uint32_t rng_stream, float* d_mq, float* d_mk, float* d_mv,
                        float* d_aq, float* d_ak, float* d_gate_logit, float* d_order_w, float* d_order_b, float* d_dist_w,
                        float* d_dist_b, float* d_scalar, float* d_rich_ratio, const int32_t* order, const int64_t* ctx_rows,
                        void* stream) {
  AttnParams p = {};
  attn_fill_common(p, mq, mk, mv, aq, ak, gate_logit, item_seq, order_w, order_b, dist_w, dist_b, scalar, B, L, H, dh, two_level,
                   combine_option, comb_scalar, rich_mode, rich_ratio, p_attn, D1, D2, D3, noise, rng, rng_stream, order, ctx_rows);
  p.t0 = d_ctx_cal; p.t1 = d_ctx_att; p.t1_is_att = 1; p.d_pen0 = d_pen_sq;
  p.d_mq = d_mq; p.d_mk = d_mk; p.d_mv = d_mv; p.d_aq = d_aq; p.d_ak = d_ak; p.d_gate = d_gate_logit;
  p.d_ow = d_order_w; p.d_ob = d_order_b; p.d_dw = d_dist_w; p.d_db = d_dist_b; p.d_scalar = d_scalar; p.d_ratio = d_rich_ratio;
  plain_range(p, d_ctx_att != nullptr);
  int rc = attn_validate(p, "attn_calib_bwd");
  if (rc) return rc;
  ACSR_REQUIRE(d_mq && d_mk && d_mv && d_aq && d_ak, "attn_calib_bwd: NULL output");
  ACSR_REQUIRE(combine_option != ACSR_ATTN_COMBINE_GATE || d_gate_logit != nullptr, "attn_calib_bwd: d_gate_logit is NULL");
  // d_order_*, d_dist_*, d_scalar, d_rich_ratio may be NULL: that cotangent stream does not own those parameters
  if (L > 64) return attn_long_bwd(p, 1, (cudaStream_t)stream);
  return dispatch_bwd<1>(p, (cudaStream_t)stream);
}

/* ACTiSASRec: acsr_attn_calib_bwd with the raw-score bias, the cotangents of the attention matrices handed out by
 * acsr_attn_calib_ti_fwd, and the gradient of the bias (written where a row's chain runs; the caller zero-fills). */
int acsr_attn_calib_ti_bwd(const float* d_ctx_att, const float* d_ctx_cal, const float* d_pen_sq, const float* d_prob_att,
                           const float* d_prob_cal, const float* s_bias, const float* mq, const float* mk,
                           const float* mv, const float* aq, const float* ak, const float* gate_logit, const int64_t* item_seq,
                           const float* order_w, const float* order_b, const float* dist_w, const float* dist_b, const float* scalar,
                           int B, int L, int H, int dh, int two_level, int combine_option, float comb_scalar, int rich_mode,
                           const float* rich_ratio, float p_attn, const float* D1, const float* D2, const float* D3,
                           const float* noise, const void* rng, uint32_t rng_stream, float* d_mq, float* d_mk, float* d_mv,
                           float* d_aq, float* d_ak, float* d_gate_logit, float* d_order_w, float* d_order_b, float* d_dist_w,
                           float* d_dist_b, float* d_scalar, float* d_rich_ratio, float* d_s_bias, void* stream) {
  AttnParams p = {};
  attn_fill_common(p, mq, mk, mv, aq, ak, gate_logit, item_seq, order_w, order_b, dist_w, dist_b, scalar, B, L, H, dh, two_level,
                   combine_option, comb_scalar, rich_mode, rich_ratio, p_attn, D1, D2, D3, noise, rng, rng_stream, nullptr, nullptr);
  p.t0 = d_ctx_cal; p.t1 = d_ctx_att; p.t1_is_att = 1; p.d_pen0 = d_pen_sq;
  p.s_bias = s_bias; p.dprob_att = d_prob_att; p.dprob_cal = d_prob_cal; p.d_s_bias = d_s_bias;
  p.d_mq = d_mq; p.d_mk = d_mk; p.d_mv = d_mv; p.d_aq = d_aq; p.d_ak = d_ak; p.d_gate = d_gate_logit;
  p.d_ow = d_order_w; p.d_ob = d_order_b; p.d_dw = d_dist_w; p.d_db = d_dist_b; p.d_scalar = d_scalar; p.d_ratio = d_rich_ratio;
  plain_range(p, d_ctx_att != nullptr || d_prob_att != nullptr);
  int rc = attn_validate(p, "attn_calib_ti_bwd");
  if (rc) return rc;
  ACSR_REQUIRE(d_mq && d_mk && d_mv && d_aq && d_ak, "attn_calib_ti_bwd: NULL output");
  ACSR_REQUIRE(combine_option != ACSR_ATTN_COMBINE_GATE || d_gate_logit != nullptr, "attn_calib_ti_bwd: d_gate_logit is NULL");
  ACSR_REQUIRE(p.plain && L <= 64, "attn_calib_ti_bwd: the time-aware terms need ACSR_ATTN_PLAIN and L <= 64");
  return dispatch_bwd<1>(p, (cudaStream_t)stream);
}

int acsr_attn_calib_bwd2(const float* d_ctx_cal0, const float* d_pen_sq0, const float* d_ctx_att1, const float* d_ctx_cal1,
                         const float* d_pen_sq1, const float* mq, const float* mk, const float* mv, const float* aq,
                         const float* ak, const float* gate_logit, const int64_t* item_seq, const float* order_w,
                         const float* order_b, const float* dist_w, const float* dist_b, const float* scalar, int B, int L, int H,
                         int dh, int two_level, int combine_option, float comb_scalar, int rich_mode, const float* rich_ratio,
                         float p_attn, const float* D1, const float* D2, const float* D3, const float* noise, const void* rng,
                         uint32_t rng_stream, float* d_mq, float* d_mk, float* d_mv, float* d_aq, float* d_ak,
                         float* d_gate_logit, float* d_order_w, float* d_order_b, float* d_dist_w, float* d_dist_b,
                         float* d_scalar, float* d_rich_ratio, const int32_t* order, const int64_t* ctx_rows, void* stream) {
  AttnParams p = {};
  attn_fill_common(p, mq, mk, mv, aq, ak, gate_logit, item_seq, order_w, order_b, dist_w, dist_b, scalar, B, L, H, dh, two_level,
                   combine_option, comb_scalar, rich_mode, rich_ratio, p_attn, D1, D2, D3, noise, rng, rng_stream, order, ctx_rows);
  ACSR_REQUIRE(!(d_ctx_att1 && d_ctx_cal1), "attn_calib_bwd2: stream 1 takes d_ctx_att or d_ctx_cal, not both");
  p.t0 = d_ctx_cal0; p.t1 = d_ctx_att1 ? d_ctx_att1 : d_ctx_cal1; p.t1_is_att = d_ctx_att1 != nullptr;
  p.d_pen0 = d_pen_sq0; p.d_pen1 = d_pen_sq1;
  p.s1_td = (long long)B * L * H * dh; p.s1_ll = (long long)B * L * L;
  p.d_mq = d_mq; p.d_mk = d_mk; p.d_mv = d_mv; p.d_aq = d_aq; p.d_ak = d_ak; p.d_gate = d_gate_logit;
  p.d_ow = d_order_w; p.d_ob = d_order_b; p.d_dw = d_dist_w; p.d_db = d_dist_b; p.d_scalar = d_scalar; p.d_ratio = d_rich_ratio;
  plain_range(p, d_ctx_att1 != nullptr);
  int rc = attn_validate(p, "attn_calib_bwd2");
  if (rc) return rc;
  ACSR_REQUIRE(d_mq && d_mk && d_mv && d_aq && d_ak, "attn_calib_bwd2: NULL output");
  ACSR_REQUIRE(combine_option != ACSR_ATTN_COMBINE_GATE || d_gate_logit != nullptr, "attn_calib_bwd2: d_gate_logit is NULL");
  if (L > 64) return attn_long_bwd(p, 2, (cudaStream_t)stream);
  return dispatch_bwd<2>(p, (cudaStream_t)stream);
}

}  // extern "C"
