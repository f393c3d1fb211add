// Fused calibrated causal attention for LONG sequences, 64 < L <= 256 (BASELINE configuration #5: L = 200, d = 256,
// 4 heads) -- same math, same Philox counters, same entry points as attn_fwd.cu / attn_bwd.cu (which keep all five
// [L,dh] tiles of a (sequence, head) in shared memory and therefore stop at L = 64).
//
// One CTA owns one (sequence, head).  The KEY-side tiles K, K', V (only the nkey rows in play) are resident in shared
// memory; the QUERY-side rows (q_i, q'_i and, in the backward, the cotangent rows) are streamed: a whole warp owns
// one query row (G = 32 lanes over the key columns, up to 8 columns per lane), copies the row's vectors into its
// private buffer and runs the same row iteration as the short kernels (attn_fwd_rows.cuh / attn_bwd_rows.cuh).
// Rows are handed out heaviest first.
//
// Backward: the row phase writes dS, dS' (per cotangent stream) and R, A row-major into a caller-provided global
// workspace (acsr_set_workspace; 4*(2*NS+2)*L*L bytes per (sequence, head), the batch is cut into chunks that fit),
// and a second kernel finishes the column-side gradients dK, dK', dV as [L,L]^T x [L,dh] products: thread = key
// column j, the [L,dh] right-hand tile broadcast from shared memory, rows i >= j only.
#pragma once
#include "attn_fwd_rows.cuh"
#include "attn_bwd_rows.cuh"

namespace acsr {

constexpr int kLongMaxL = 256;

void* workspace_ptr(size_t* bytes);     // api.cu: acsr_set_workspace of the current device

struct LongCarve {
  AttnSmem sm;
  float* wrows;     // [warps][NR][dh+4] streamed query-side rows
  float* rowbuf;    // [warps][NB][LP]
};

__device__ __forceinline__ LongCarve carve_long(float*& ptr, int LP, int dh, int nr, int nb) {
  LongCarve c;
  const int tile = LP * (dh + 4);
  c.sm.K = ptr; ptr += tile;
  c.sm.V = ptr; ptr += tile;
  c.sm.K2 = ptr; ptr += tile;
  c.sm.Q = nullptr; c.sm.Q2 = nullptr;
  c.sm.rowO = ptr; ptr += LP;
  c.sm.rowD = ptr; ptr += LP;
  c.sm.colO = ptr; ptr += LP;
  c.sm.colD = ptr; ptr += LP;
  c.sm.logd = ptr; ptr += LP;
  c.sm.keyok = ptr; ptr += LP;
  c.sm.wo = ptr; ptr += 2 * dh;
  c.sm.wd = ptr; ptr += 2 * dh;
  c.sm.misc = reinterpret_cast<int*>(ptr); ptr += 8;
  c.sm.G = nullptr;
  c.wrows = ptr; ptr += kAttnWarps * nr * (dh + 4);
  c.rowbuf = ptr; ptr += kAttnWarps * nb * LP;
  return c;
}
static inline size_t long_floats(int LP, int dh, int nr, int nb) {
  return (size_t)3 * LP * (dh + 4) + 6 * LP + 4 * dh + 8 + (size_t)kAttnWarps * nr * (dh + 4) + (size_t)kAttnWarps * nb * LP;
}

// key-side tiles + the per-row / per-column scalars of the spatial calibrator; returns nkey (>= 1)
template <int DH>
__device__ __forceinline__ int stage_long(const AttnParams& p, const AttnSmem& sm, int b, int h, int LP) {
  constexpr int dhp = DH + 4;
  const int L = p.L;
  if (threadIdx.x < 8) sm.misc[threadIdx.x] = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < LP; j += blockDim.x) {
    const bool ok = j < L && p.item_seq[(long long)b * L + j] != 0;
    sm.keyok[j] = ok ? 1.0f : 0.0f;
    sm.logd[j] = logf((float)j + 1.0f);
    if (ok) atomicMax(sm.misc, j + 1);
  }
  for (int c = threadIdx.x; c < 2 * DH; c += blockDim.x) {
    sm.wo[c] = p.ow ? p.ow[c] : 0.f;
    sm.wd[c] = p.dw ? p.dw[c] : 0.f;
  }
  __syncthreads();
  const int nkey = max(sm.misc[0], 1);
  const int nkp = (nkey + 3) & ~3;
  stage_tile<DH>(sm.K, p.mk, b, h, L, p.d, nkey, nkp);
  stage_tile<DH>(sm.V, p.mv, b, h, L, p.d, nkey, nkp);
  stage_tile<DH>(sm.K2, p.ak, b, h, L, p.d, nkey, nkp);
  cp_async_wait_all();
  __syncthreads();
  if (p.ow || p.dw) {      // rank-1 pieces of the affines: 4 threads per row, DH/4 channels each; query rows from global
    const float* qb = p.mq + (long long)b * L * p.d + h * DH;
    for (int j0 = 0; j0 < LP; j0 += kAttnThreads / 4) {
      const int j = j0 + (threadIdx.x >> 2), part = threadIdx.x & 3;
      float ro = 0.f, rd = 0.f, co = 0.f, cd = 0.f;
      if (j < L) {
#pragma unroll
        for (int cc = 0; cc < DH / 4; ++cc) {
          const int c = part * (DH / 4) + cc;
          const float q = qb[(long long)j * p.d + c];
          ro = fmaf(q, sm.wo[c], ro); rd = fmaf(q, sm.wd[c], rd);
          if (j < nkey) {
            const float k = sm.K[j * dhp + c];
            co = fmaf(k, sm.wo[DH + c], co); cd = fmaf(k, sm.wd[DH + c], cd);
          }
        }
      }
      ro += __shfl_xor_sync(kFull, ro, 1); ro += __shfl_xor_sync(kFull, ro, 2);
      rd += __shfl_xor_sync(kFull, rd, 1); rd += __shfl_xor_sync(kFull, rd, 2);
      co += __shfl_xor_sync(kFull, co, 1); co += __shfl_xor_sync(kFull, co, 2);
      cd += __shfl_xor_sync(kFull, cd, 1); cd += __shfl_xor_sync(kFull, cd, 2);
      if (part == 0 && j < LP) { sm.rowO[j] = ro; sm.rowD[j] = rd; sm.colO[j] = co; sm.colD[j] = cd; }
    }
    __syncthreads();
  }
  return nkey;
}

// copy one [DH] head slice of token row i into a warp-private buffer
template <int DH>
__device__ __forceinline__ void load_row(float* dst, const float* src, int b, int h, int i, int L, int d) {
  const int lane = threadIdx.x & 31;
  if (src == nullptr) return;
  const float* g = src + ((long long)b * L + i) * d + h * DH;
  for (int c = lane; c < DH; c += 32) dst[c] = g[c];
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_long_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int dhp = DH + 4;
  const int L = p.L, LP = (L + 3) & ~3;
  const int b = p.order ? p.order[blockIdx.x / p.H] : (int)(blockIdx.x / p.H), h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ptr = smem_f;
  LongCarve lc = carve_long(ptr, LP, DH, 2, 2);
  double* pen_red = reinterpret_cast<double*>(ptr);
  if (threadIdx.x == 0) *pen_red = 0.0;
  AttnSmem sm = lc.sm;
  const int nkey = stage_long<DH>(p, sm, b, h, LP);
  const bool need_att = p.ctx_att != nullptr;
  const RowConst kc = make_consts<DH>(p, need_att);
  if (p.probs) {
    const long long plane = (long long)p.B * p.H * L * L;
    float* base = p.probs + ((long long)b * p.H + h) * L * L;
    for (int e = threadIdx.x; e < L * L; e += blockDim.x)
#pragma unroll
      for (int k = 0; k < 6; ++k) base[k * plane + e] = 0.f;
    __syncthreads();
  }
  float* wq = lc.wrows + (warp * 2 + 0) * dhp;
  float* wq2 = lc.wrows + (warp * 2 + 1) * dhp;
  FwdCtx cx;
  cx.rowbuf = lc.rowbuf + warp * 2 * LP;
  cx.pen = 0.f;
  const int rt = p.ctx_rows ? (int)p.ctx_rows[b] - 1 : -1;
  for (int t0 = next_task(sm.misc + 2, 1); t0 < L; t0 = next_task(sm.misc + 2, 1)) {   // heaviest rows first
    const int i = L - 1 - t0;
    const int bound = min(i + 1, nkey);
    const int nj = (bound + 31) >> 5;
    AttnSmem s2 = sm;
    s2.Q = wq - i * dhp;
    s2.Q2 = wq2 - i * dhp;
    __syncwarp();
    if (rt >= 0 && rt != i) {            // context not consumed: attack mask -> penalty only
      if (p.pen_sq == nullptr) continue;
      load_row<DH>(wq2, p.aq, b, h, i, L, p.d);
      __syncwarp();
      if (nj <= 2) fwd_row_iter_m<DH, 32, 2>(p, s2, kc, b, h, i, true, bound, lane, cx);
      else if (nj <= 4) fwd_row_iter_m<DH, 32, 4>(p, s2, kc, b, h, i, true, bound, lane, cx);
      else if (nj <= 6) fwd_row_iter_m<DH, 32, 6>(p, s2, kc, b, h, i, true, bound, lane, cx);
      else fwd_row_iter_m<DH, 32, 8>(p, s2, kc, b, h, i, true, bound, lane, cx);
      continue;
    }
    load_row<DH>(wq, p.mq, b, h, i, L, p.d);
    load_row<DH>(wq2, p.aq, b, h, i, L, p.d);
    __syncwarp();
    if (nj <= 2) fwd_row_iter<DH, 32, 2>(p, s2, kc, b, h, i, true, bound, 0, lane, need_att, LP, cx);
    else if (nj <= 4) fwd_row_iter<DH, 32, 4>(p, s2, kc, b, h, i, true, bound, 0, lane, need_att, LP, cx);
    else if (nj <= 6) fwd_row_iter<DH, 32, 6>(p, s2, kc, b, h, i, true, bound, 0, lane, need_att, LP, cx);
    else fwd_row_iter<DH, 32, 8>(p, s2, kc, b, h, i, true, bound, 0, lane, need_att, LP, cx);
  }
  const double pd = warp_sum_d((double)cx.pen);
  if (lane == 0 && p.pen_sq != nullptr) {
    atomicAdd(pen_red, pd);
    __threadfence_block();
    if (atomicAdd(sm.misc + 4, 1) == kAttnWarps - 1) atomicAdd(p.pen_sq, *reinterpret_cast<volatile double*>(pen_red));
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward, row phase
// ------------------------------------------------------------------------------------------------------------
// per-(sequence, head) slice of the workspace, in floats: (2 NS + 2) [L,L] matrices, then colDU / colDT [NS][L] each
__host__ __device__ static inline size_t long_ws_floats(int L, int ns) { return (size_t)(2 * ns + 2) * L * L + (size_t)2 * ns * L; }

template <int DH, int NS>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_long_bwd_rows_kernel(const AttnParams p, float* __restrict__ ws, const int b0) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int dhp = DH + 4;
  const int L = p.L, LP = (L + 3) & ~3;
  const int g = b0 + blockIdx.x / p.H;
  const int b = p.order ? p.order[g] : g, h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ptr = smem_f;
  LongCarve lc = carve_long(ptr, LP, DH, 4, 2 * NS);
  AttnSmem sm = lc.sm;
  BwdSmem<NS> bs;
  bs.colDU = ptr; ptr += NS * LP;
  bs.colDT = ptr; ptr += NS * LP;
  bs.pacc = ptr; ptr += 4 * DH;
  bs.red = ptr; ptr += kAttnWarps * 4;
  bs.rowbuf = lc.rowbuf;
  bs.sT0 = bs.sT1 = nullptr;
  bs.matRT = bs.matAT = nullptr;
  const size_t LL = (size_t)L * L;
  float* slice = ws + (size_t)blockIdx.x * long_ws_floats(L, NS);
#pragma unroll
  for (int s = 0; s < NS; ++s) { bs.matST[s] = bs.matS2T[s] = nullptr; bs.gS[s] = slice + (2 * s) * LL; bs.gS2[s] = slice + (2 * s + 1) * LL; }
  bs.gR = slice + (2 * NS) * LL;
  bs.gA = slice + (2 * NS + 1) * LL;
  float* g_col = slice + (2 * NS + 2) * LL;        // [NS][2][L]

  BwdFlags f;
  f.has_t0 = p.t0 != nullptr; f.has_t1 = p.t1 != nullptr;
  f.t1_att = NS == 1 ? true : (p.t1_is_att != 0);
  f.has_att = f.has_t1 && f.t1_att;
  f.gate = p.combine == ACSR_ATTN_COMBINE_GATE;
  for (int j = threadIdx.x; j < NS * LP; j += blockDim.x) { bs.colDU[j] = 0.f; bs.colDT[j] = 0.f; }
  for (int j = threadIdx.x; j < 4 * DH; j += blockDim.x) bs.pacc[j] = 0.f;
  const int nkey = stage_long<DH>(p, sm, b, h, LP);
  const RowConst kc = make_consts<DH>(p, f.has_att);
  f.sc2 = kc.sc * kc.sc;
  float dpen[NS];
  dpen[0] = p.d_pen0 ? p.d_pen0[0] : 0.f;
  if (NS == 2) dpen[NS - 1] = p.d_pen1 ? p.d_pen1[0] : 0.f;
  BwdAcc acc;
  acc.s_ob = acc.s_db = acc.s_scalar = acc.s_ratio = 0.f;

  using CM = CMap<DH, 32>;
  const int c0 = CM::c0(lane);
  float accOq[CM::CPL], accDq[CM::CPL];
#pragma unroll
  for (int k = 0; k < CM::CPL; ++k) accOq[k] = accDq[k] = 0.f;
  float* wr = lc.wrows + warp * 4 * dhp;
  float* wbuf = bs.rowbuf + warp * 2 * NS * LP;
  for (int c = lane; c < 4 * dhp; c += 32) wr[c] = 0.f;
  const int rt = p.ctx_rows ? (int)p.ctx_rows[b] - 1 : -1;
  for (int t0 = next_task(sm.misc + 2, 1); t0 < L; t0 = next_task(sm.misc + 2, 1)) {
    const int i = L - 1 - t0;
    const int bound = min(i + 1, nkey);
    const int nj = (bound + 31) >> 5;
    AttnSmem s2 = sm;
    s2.Q = wr - i * dhp;
    s2.Q2 = wr + dhp - i * dhp;
    BwdSmem<NS> b2 = bs;
    b2.sT0 = wr + 2 * dhp - i * dhp;
    b2.sT1 = wr + 3 * dhp - i * dhp;
    __syncwarp();
    if (rt >= 0 && rt != i) {
      load_row<DH>(wr + dhp, p.aq, b, h, i, L, p.d);
      __syncwarp();
#define ACSR_LROW_M(NJV) bwd_row_iter_m<DH, 32, NJV, NS, true>(p, s2, b2, kc, dpen, b, h, i, true, bound, 0, lane, LP, wbuf)
      if (nj <= 2) ACSR_LROW_M(2);
      else if (nj <= 4) ACSR_LROW_M(4);
      else if (nj <= 6) ACSR_LROW_M(6);
      else ACSR_LROW_M(8);
#undef ACSR_LROW_M
      continue;
    }
    load_row<DH>(wr, p.mq, b, h, i, L, p.d);
    load_row<DH>(wr + dhp, p.aq, b, h, i, L, p.d);
    load_row<DH>(wr + 2 * dhp, p.t0, b, h, i, L, p.d);
    load_row<DH>(wr + 3 * dhp, p.t1, b, h, i, L, p.d);
    __syncwarp();
#define ACSR_LROW(NJV) \
  bwd_row_iter<DH, 32, NJV, NS, true>(p, s2, b2, kc, f, dpen, b, h, i, true, bound, 0, lane, LP, wbuf, acc, accOq, accDq)
    if (nj <= 2) ACSR_LROW(2);
    else if (nj <= 4) ACSR_LROW(4);
    else if (nj <= 6) ACSR_LROW(6);
    else ACSR_LROW(8);
#undef ACSR_LROW
  }
  if ((p.d_ow || p.d_dw) && CM::split(lane) == 0) {
#pragma unroll
    for (int k = 0; k < CM::CPL; ++k) {
      const int c = c0 + k;
      if (accOq[k] != 0.f) atomicAdd(bs.pacc + c, accOq[k]);
      if (accDq[k] != 0.f) atomicAdd(bs.pacc + 2 * DH + c, accDq[k]);
    }
  }
  const float s_ob = warp_sum(acc.s_ob), s_db = warp_sum(acc.s_db);
  const float s_scalar = warp_sum(acc.s_scalar), s_ratio = warp_sum(acc.s_ratio);
  if (lane == 0) {
    bs.red[warp * 4 + 0] = s_ob; bs.red[warp * 4 + 1] = s_db; bs.red[warp * 4 + 2] = s_scalar; bs.red[warp * 4 + 3] = s_ratio;
  }
  __syncthreads();
  // column sums of du / dt: to the workspace (rank-1 part of dK, column kernel) and, stream 0, the key halves of d_ow / d_dw
  for (int e = threadIdx.x; e < NS * L; e += blockDim.x) {
    const int s = e / L, j = e - s * L;
    g_col[(2 * s + 0) * L + j] = bs.colDU[s * LP + j];
    g_col[(2 * s + 1) * L + j] = bs.colDT[s * LP + j];
  }
  if ((p.d_ow || p.d_dw) && threadIdx.x < DH) {
    const int c = threadIdx.x;
    float so = 0.f, sd = 0.f;
    for (int j = 0; j < nkey; ++j) {
      const float k = sm.K[j * dhp + c];
      so = fmaf(bs.colDU[j], k, so);
      sd = fmaf(bs.colDT[j], k, sd);
    }
    bs.pacc[DH + c] += so;
    bs.pacc[3 * DH + c] += sd;
  }
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < kAttnWarps; ++w) s += bs.red[w * 4 + threadIdx.x];
    float* dst = threadIdx.x == 0 ? p.d_ob : threadIdx.x == 1 ? p.d_db : threadIdx.x == 2 ? p.d_scalar : p.d_ratio;
    if (dst != nullptr && s != 0.f) atomicAdd(dst, s);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 4 * DH; c += blockDim.x) {
    const float v = bs.pacc[c];
    float* dst = c < 2 * DH ? p.d_ow : p.d_dw;
    if (dst != nullptr && v != 0.f) atomicAdd(dst + (c < 2 * DH ? c : c - 2 * DH), v);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward, column phase: out[j][:] = sum_{i >= j} X[i][j] * Y[i][:]  for the (2 NS + 2) workspace matrices
// ------------------------------------------------------------------------------------------------------------
template <int DH, int NS>
__global__ void __launch_bounds__(kLongMaxL, 1) attn_long_bwd_cols_kernel(const AttnParams p, const float* __restrict__ ws, const int b0) {
  extern __shared__ __align__(16) float smem_f[];
  float* Y = smem_f;                                // [L][DH]
  const int L = p.L;
  const int g = b0 + blockIdx.x / p.H;
  const int b = p.order ? p.order[g] : g, h = blockIdx.x % p.H;
  const int j = threadIdx.x;
  const bool jok = j < L;
  const int i0 = threadIdx.x & ~31;                 // first row that can be non-zero for this warp's columns
  const size_t LL = (size_t)L * L;
  const float* slice = ws + (size_t)blockIdx.x * long_ws_floats(L, NS);
  const float* g_col = slice + (2 * NS + 2) * LL;
  const bool t1_att = NS == 1 ? true : (p.t1_is_att != 0);
  float acc[DH];

  auto zero = [&]() {
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
  };
  auto accumulate = [&](const float* X, const float* ysrc) {     // acc += X^T . Y  (Y = head slice of ysrc, NULL: nothing)
    if (ysrc == nullptr) return;                                  // uniform over the CTA
    __syncthreads();
    const float* yb = ysrc + (long long)b * L * p.d + h * DH;
    for (int e = threadIdx.x; e < L * (DH / 4); e += blockDim.x) {
      const int r = e / (DH / 4), c4 = e % (DH / 4);
      *reinterpret_cast<float4*>(Y + r * DH + c4 * 4) = *reinterpret_cast<const float4*>(yb + (long long)r * p.d + c4 * 4);
    }
    __syncthreads();
    if (!jok) return;
#pragma unroll 2
    for (int i = i0; i < L; ++i) {
      const float x = X[(size_t)i * L + j];
      const float4* y4 = reinterpret_cast<const float4*>(Y + i * DH);
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 y = y4[c4];
        acc[4 * c4 + 0] = fmaf(x, y.x, acc[4 * c4 + 0]);
        acc[4 * c4 + 1] = fmaf(x, y.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(x, y.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(x, y.w, acc[4 * c4 + 3]);
      }
    }
  };
  auto flush = [&](float* out, int s) {
    if (!jok) return;
    float* o = out + s * p.s1_td + ((long long)b * L + j) * p.d + h * DH;
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; ++c4)
      *reinterpret_cast<float4*>(o + 4 * c4) = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
  };

#pragma unroll
  for (int s = 0; s < NS; ++s) {
    // dK_s = dS_s^T Q + colDU_s (x) wo[dh:] + colDT_s (x) wd[dh:]
    zero();
    accumulate(slice + (2 * s) * LL, p.mq);
    if (jok) {
      const float cdu = g_col[(2 * s + 0) * L + j], cdt = g_col[(2 * s + 1) * L + j];
#pragma unroll
      for (int c = 0; c < DH; ++c) {
        if (p.ow) acc[c] = fmaf(cdu, __ldg(p.ow + DH + c), acc[c]);
        if (p.dw) acc[c] = fmaf(cdt, __ldg(p.dw + DH + c), acc[c]);
      }
    }
    flush(p.d_mk, s);
    // dK'_s = dS'_s^T Q'
    zero();
    accumulate(slice + (2 * s + 1) * LL, p.aq);
    flush(p.d_ak, s);
  }
  // dV: stream 0 <- R^T t0 ; last stream <- (A or R)^T t1
  const float* XR = slice + (2 * NS) * LL;
  const float* XA = slice + (2 * NS + 1) * LL;
  zero();
  accumulate(XR, p.t0);
  if (NS == 1) {
    accumulate(t1_att ? XA : XR, p.t1);
    flush(p.d_mv, 0);
  } else {
    flush(p.d_mv, 0);
    zero();
    accumulate(t1_att ? XA : XR, p.t1);
    flush(p.d_mv, NS - 1);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
template <typename K>
static int prep_long(K kernel, size_t smem, const char* who, int L, int dh) {
  if (smem > 227 * 1024) {
    set_error("%s: L=%d with head size %d needs %zu bytes of shared memory (> 227 KB)", who, L, dh, smem);
    return ACSR_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("%s: smem %zu: %s", who, smem, cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  return ACSR_OK;
}

template <int DH>
static int launch_long_fwd(const AttnParams& p, cudaStream_t st) {
  const int LP = (p.L + 3) & ~3;
  const size_t smem = long_floats(LP, DH, 2, 2) * sizeof(float) + sizeof(double) * 2;
  int rc = prep_long(attn_long_fwd_kernel<DH>, smem, "attn_calib_fwd", p.L, DH);
  if (rc) return rc;
  attn_long_fwd_kernel<DH><<<dim3(p.B * p.H), dim3(kAttnThreads), smem, st>>>(p);
  return check_launch("attn_calib_fwd");
}

template <int DH, int NS>
static int launch_long_bwd(const AttnParams& p, cudaStream_t st) {
  const char* who = NS == 1 ? "attn_calib_bwd" : "attn_calib_bwd2";
  const int LP = (p.L + 3) & ~3;
  const size_t smem_r = (long_floats(LP, DH, 4, 2 * NS) + 2 * NS * LP + 4 * DH + kAttnWarps * 4) * sizeof(float);
  const size_t smem_c = (size_t)p.L * DH * sizeof(float);
  int rc = prep_long(attn_long_bwd_rows_kernel<DH, NS>, smem_r, who, p.L, DH);
  if (rc) return rc;
  rc = prep_long(attn_long_bwd_cols_kernel<DH, NS>, smem_c, who, p.L, DH);
  if (rc) return rc;
  size_t ws_bytes = 0;
  float* ws = reinterpret_cast<float*>(workspace_ptr(&ws_bytes));
  const size_t per_seq = long_ws_floats(p.L, NS) * sizeof(float) * p.H;
  if (ws == nullptr || ws_bytes < per_seq) {
    set_error("%s: L=%d (> 64) needs a workspace of at least %zu bytes (acsr_set_workspace; %zu per sequence)", who, p.L, per_seq, per_seq);
    return ACSR_ERR_ARG;
  }
  const int chunk = (int)((ws_bytes / per_seq) < (size_t)p.B ? (ws_bytes / per_seq) : (size_t)p.B);
  for (int b0 = 0; b0 < p.B; b0 += chunk) {
    const int nb = p.B - b0 < chunk ? p.B - b0 : chunk;
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)nb * per_seq, st);
    if (e != cudaSuccess) { set_error("%s: workspace memset: %s", who, cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
    attn_long_bwd_rows_kernel<DH, NS><<<dim3(nb * p.H), dim3(kAttnThreads), smem_r, st>>>(p, ws, b0);
    attn_long_bwd_cols_kernel<DH, NS><<<dim3(nb * p.H), dim3(kLongMaxL), smem_c, st>>>(p, ws, b0);
  }
  return check_launch(who);
}

}  // namespace acsr
