// Fused calibrated causal attention, forward (see attn_common.cuh for the design).
// Replaces layers.py:657-674, 686-742, 883-896, 917-936 and the probs.V of 676-678.
#include "attn_fwd_rows.cuh"

namespace acsr {

// all rows of the (b,h) tile with G-lane row groups; the heaviest rows go first
template <int DH, int G, int MAXNJ>
__device__ __forceinline__ void fwd_rows(const AttnParams& p, const AttnSmem& sm, const RowConst& kc, int b, int h, int nkey,
                                         bool need_att, int rstride, FwdCtx& cx) {
  constexpr int RPW = 32 / G;
  const int L = p.L;
  const int lane = threadIdx.x & 31;
  const int grp = lane / G, sub = lane % G;
  const int rt = p.ctx_rows ? (int)p.ctx_rows[b] - 1 : -1;     // the only consumed context row, or -1: all rows
  for (int t0 = next_task(sm.misc + 2, RPW); t0 < L; t0 = next_task(sm.misc + 2, RPW)) {   // heaviest rows first
    const int iw = L - 1 - t0;                    // largest row of this warp (>= 0)
    const int iraw = iw - grp;
    const bool rowok = iraw >= 0;
    const int i = rowok ? iraw : 0;
    const int bound = p.full ? nkey : min(i + 1, nkey);
    const int nj = ((p.full ? nkey : min(iw + 1, nkey)) + G - 1) / G;      // warp-uniform number of column groups
    if (rt >= 0 && (rt > iw || rt <= iw - RPW)) {       // none of this warp's rows is the consumed one: mask + penalty only
      if (p.pen_sq == nullptr) continue;
      if (MAXNJ == 1 || nj == 1) fwd_row_iter_m<DH, G, 1>(p, sm, kc, b, h, i, rowok, bound, sub, cx);
      else if (nj == 2) fwd_row_iter_m<DH, G, (MAXNJ >= 2 ? 2 : 1)>(p, sm, kc, b, h, i, rowok, bound, sub, cx);
      else if (nj == 3) fwd_row_iter_m<DH, G, (MAXNJ >= 3 ? 3 : 1)>(p, sm, kc, b, h, i, rowok, bound, sub, cx);
      else fwd_row_iter_m<DH, G, (MAXNJ >= 4 ? 4 : 1)>(p, sm, kc, b, h, i, rowok, bound, sub, cx);
      continue;
    }
    if (MAXNJ == 1 || nj == 1) fwd_row_iter<DH, G, 1>(p, sm, kc, b, h, i, rowok, bound, grp, sub, need_att, rstride, cx);
    else if (nj == 2) fwd_row_iter<DH, G, (MAXNJ >= 2 ? 2 : 1)>(p, sm, kc, b, h, i, rowok, bound, grp, sub, need_att, rstride, cx);
    else if (nj == 3) fwd_row_iter<DH, G, (MAXNJ >= 3 ? 3 : 1)>(p, sm, kc, b, h, i, rowok, bound, grp, sub, need_att, rstride, cx);
    else fwd_row_iter<DH, G, (MAXNJ >= 4 ? 4 : 1)>(p, sm, kc, b, h, i, rowok, bound, grp, sub, need_att, rstride, cx);
  }
}

template <int DH>
__global__ void __launch_bounds__(kAttnThreads, 4) attn_fwd_kernel(const AttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smem_f[];
  const int L = p.L, LP = (L + 3) & ~3;
  const int b = p.order ? p.order[blockIdx.x / p.H] : (int)(blockIdx.x / p.H), h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ptr = smem_f;
  AttnSmem sm = carve_common(ptr, LP, DH);
  const int rbw = rowbuf_floats_per_warp(LP, 2);
  FwdCtx cx;
  cx.rowbuf = ptr + warp * rbw;
  cx.pen = 0.f;
  ptr += kAttnWarps * rbw;
  double* pen_red = reinterpret_cast<double*>(ptr);   // CTA penalty sum; ptr offset is a multiple of 4 floats
  if (threadIdx.x == 0) *pen_red = 0.0;
  const int nkey = stage_common<DH>(p, sm, b, h, LP);
  const bool need_att = p.ctx_att != nullptr;
  const RowConst kc = make_consts<DH>(p, need_att);
  if (p.probs) {           // introspection output: columns outside a row's range are exactly 0
    const long long plane = (long long)p.B * p.H * L * L;
    float* base = p.probs + ((long long)b * p.H + h) * L * L;
    for (int e = threadIdx.x; e < L * L; e += blockDim.x)
#pragma unroll
      for (int k = 0; k < 6; ++k) base[k * plane + e] = 0.f;
    __syncthreads();
  }
  if (nkey <= 8) fwd_rows<DH, 8, 1>(p, sm, kc, b, h, nkey, need_att, 8, cx);
  else fwd_rows<DH, 16, 4>(p, sm, kc, b, h, nkey, need_att, LP, cx);
  // CTA penalty sum without a closing barrier: the last warp to arrive publishes it
  const double pd = warp_sum_d((double)cx.pen);
  if (lane == 0 && p.pen_sq != nullptr) {
    atomicAdd(pen_red, pd);
    __threadfence_block();
    if (atomicAdd(sm.misc + 4, 1) == kAttnWarps - 1) atomicAdd(p.pen_sq, *reinterpret_cast<volatile double*>(pen_red));
  }
}

static size_t fwd_smem_bytes(int L, int dh) {
  const int LP = (L + 3) & ~3;
  size_t f = common_floats(LP, dh) + (size_t)kAttnWarps * rowbuf_floats_per_warp(LP, 2);
  return f * sizeof(float) + kAttnWarps * sizeof(double);
}

template <int DH>
static int launch_fwd(const AttnParams& p, cudaStream_t st) {
  size_t smem = fwd_smem_bytes(p.L, DH);
  int rc = prep_kernel(attn_fwd_kernel<DH>, smem, "attn_calib_fwd");
  if (rc) return rc;
  launch_pdl(attn_fwd_kernel<DH>, dim3(p.B * p.H), dim3(kAttnThreads), smem, st, p);
  return check_launch("attn_calib_fwd");
}

// sequences sorted by the number of keys in play (1 + index of the last real item), longest first: the CTAs of the
// attention kernels are scheduled in that order, so the long sequences start first and the short ones fill the tail
__global__ void __launch_bounds__(256) seq_order_kernel(const int64_t* __restrict__ item_seq, int B, int L, int32_t* __restrict__ order) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int cnt[66], off[66];
  __shared__ unsigned char key[2048];
  for (int i = threadIdx.x; i < 66; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  // thread per sequence: the L loads of a row are independent, so they are all in flight at once
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long* row = reinterpret_cast<const long long*>(item_seq) + (long long)b * L;
    unsigned long long m = 0ull;
#pragma unroll 16
    for (int j = 0; j < L; ++j) m |= (unsigned long long)(__ldg(row + j) != 0) << j;
    const int nkey = m ? 64 - __clzll((long long)m) : 0;
    key[b] = (unsigned char)nkey;
    atomicAdd(cnt + nkey, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 65; k >= 0; --k) { off[k] = run; run += cnt[k]; }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) order[atomicAdd(off + key[b], 1)] = b;
}

}  // namespace acsr

using namespace acsr;

extern "C" int acsr_seq_order(const int64_t* item_seq, int B, int L, int32_t* order, void* stream) {
  ACSR_REQUIRE(item_seq && order && B > 0 && B <= 2048 && L > 0 && L <= 1024, "seq_order: bad arguments (B <= 2048, L <= 1024)");
  if (L > 64) return seq_order_long(item_seq, B, L, order, (cudaStream_t)stream);
  launch_pdl(seq_order_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, item_seq, B, L, order);
  return check_launch("seq_order");
}


extern "C" int acsr_attn_calib_fwd(const float* mq, const float* mk, const float* mv, const float* aq, const float* ak,
                                   const float* gate_logit, const int64_t* item_seq, const float* order_w, const float* order_b,
                                   const float* dist_w, const float* dist_b, const float* scalar, int B, int L, int H, int dh,
                                   int two_level, int combine_option, float comb_scalar, int rich_mode, const float* rich_ratio,
                                   float p_attn, const float* D1, const float* D2, const float* D3, const float* noise,
                                   const void* rng, uint32_t rng_stream, float* ctx_att, float* ctx_cal, double* pen_sq,
                                   float* probs_out, const int32_t* order, const int64_t* ctx_rows, void* stream) {
  AttnParams p = {};
  attn_fill_common(p, mq, mk, mv, aq, ak, gate_logit, item_seq, order_w, order_b, dist_w, dist_b, scalar, B, L, H, dh, two_level,
                   combine_option, comb_scalar, rich_mode, rich_ratio, p_attn, D1, D2, D3, noise, rng, rng_stream, order, ctx_rows);
  p.ctx_att = ctx_att; p.ctx_cal = ctx_cal; p.pen_sq = pen_sq; p.probs = probs_out;
  plain_range(p, ctx_att != nullptr || probs_out != nullptr);
  int rc = attn_validate(p, "attn_calib_fwd");
  if (rc) return rc;
  ACSR_REQUIRE(ctx_cal != nullptr, "attn_calib_fwd: ctx_cal is NULL");
  if (L > 64) return attn_long_fwd(p, (cudaStream_t)stream);
  switch (dh) {
    case 8: return launch_fwd<8>(p, (cudaStream_t)stream);
    case 16: return launch_fwd<16>(p, (cudaStream_t)stream);
    case 32: return launch_fwd<32>(p, (cudaStream_t)stream);
    case 64: return launch_fwd<64>(p, (cudaStream_t)stream);
  }
  return ACSR_ERR_UNSUPPORTED;
}

/* ACTiSASRec: acsr_attn_calib_fwd with an additive raw-score bias (q.posK + q.timeK[t_ij], acsr_pair_score) and the attacked /
 * final calibrated attention matrices written out [B,H,L,L] (the time-aware context terms are formed from them by
 * acsr_pair_context).  Plain (transformer_layers.py) variant, L <= 64: every entry of both matrices is written. */
extern "C" int acsr_attn_calib_ti_fwd(const float* s_bias, const float* mq, const float* mk, const float* mv, const float* aq,
                                      const float* ak, const float* gate_logit, const int64_t* item_seq, const float* order_w,
                                      const float* order_b, const float* dist_w, const float* dist_b, const float* scalar, int B,
                                      int L, int H, int dh, int two_level, int combine_option, float comb_scalar, int rich_mode,
                                      const float* rich_ratio, float p_attn, const float* D1, const float* D2, const float* D3,
                                      const float* noise, const void* rng, uint32_t rng_stream, float* ctx_att, float* ctx_cal,
                                      double* pen_sq, float* prob_att, float* prob_cal, void* stream) {
  AttnParams p = {};
  attn_fill_common(p, mq, mk, mv, aq, ak, gate_logit, item_seq, order_w, order_b, dist_w, dist_b, scalar, B, L, H, dh, two_level,
                   combine_option, comb_scalar, rich_mode, rich_ratio, p_attn, D1, D2, D3, noise, rng, rng_stream, nullptr, nullptr);
  p.ctx_att = ctx_att; p.ctx_cal = ctx_cal; p.pen_sq = pen_sq; p.probs = nullptr;
  p.s_bias = s_bias; p.prob_att_out = prob_att; p.prob_cal_out = prob_cal;
  plain_range(p, ctx_att != nullptr);       // without the attacked stream only the keys in play are written to prob_cal (caller zero-fills)
  int rc = attn_validate(p, "attn_calib_ti_fwd");
  if (rc) return rc;
  ACSR_REQUIRE(ctx_cal != nullptr && prob_cal != nullptr, "attn_calib_ti_fwd: ctx_cal / prob_cal is NULL");
  ACSR_REQUIRE(p.plain && L <= 64, "attn_calib_ti_fwd: the time-aware terms need ACSR_ATTN_PLAIN and L <= 64");
  switch (dh) {
    case 8: return launch_fwd<8>(p, (cudaStream_t)stream);
    case 16: return launch_fwd<16>(p, (cudaStream_t)stream);
    case 32: return launch_fwd<32>(p, (cudaStream_t)stream);
    case 64: return launch_fwd<64>(p, (cudaStream_t)stream);
  }
  return ACSR_ERR_UNSUPPORTED;
}
