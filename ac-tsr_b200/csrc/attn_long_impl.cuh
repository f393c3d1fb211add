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
// Backward: the row phase writes dS, dS' (per cotangent stream) and R, A row-major (every entry in play, no memset) into a caller-provided global
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

// the same copy split in two so that the global loads of the NEXT row are in flight while the current row is processed
template <int DH>
struct RowRegs { float v[(DH + 31) / 32]; };
template <int DH>
__device__ __forceinline__ void fetch_row(RowRegs<DH>& r, const float* src, int b, int h, int i, int L, int d) {
  const int lane = threadIdx.x & 31;
  if (src == nullptr) return;
  const float* g = src + ((long long)b * L + i) * d + h * DH;
#pragma unroll
  for (int q = 0; q < (DH + 31) / 32; ++q) if (lane + 32 * q < DH) r.v[q] = __ldg(g + lane + 32 * q);
}
template <int DH>
__device__ __forceinline__ void put_row(float* dst, const RowRegs<DH>& r, const float* src) {
  const int lane = threadIdx.x & 31;
  if (src == nullptr) return;
#pragma unroll
  for (int q = 0; q < (DH + 31) / 32; ++q) if (lane + 32 * q < DH) dst[lane + 32 * q] = r.v[q];
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
  RowRegs<DH> rq, rq2;
  int t0 = next_task(sm.misc + 2, 1);
  if (t0 < L) {
    if (!(rt >= 0 && rt != L - 1 - t0)) fetch_row<DH>(rq, p.mq, b, h, L - 1 - t0, L, p.d);
    fetch_row<DH>(rq2, p.aq, b, h, L - 1 - t0, L, p.d);
  }
  while (t0 < L) {                                   // heaviest rows first
    const int i = L - 1 - t0;
    const int bound = min(i + 1, nkey);
    const int nj = (bound + 31) >> 5;
    const bool mask_only = rt >= 0 && rt != i;       // context not consumed: attack mask -> penalty only
    AttnSmem s2 = sm;
    s2.Q = wq - i * dhp;
    s2.Q2 = wq2 - i * dhp;
    __syncwarp();
    if (!mask_only) put_row<DH>(wq, rq, p.mq);
    put_row<DH>(wq2, rq2, p.aq);
    __syncwarp();
    t0 = next_task(sm.misc + 2, 1);
    if (t0 < L) {                                    // next row's vectors: in flight while this row is processed
      if (!(rt >= 0 && rt != L - 1 - t0)) fetch_row<DH>(rq, p.mq, b, h, L - 1 - t0, L, p.d);
      fetch_row<DH>(rq2, p.aq, b, h, L - 1 - t0, L, p.d);
    }
    if (mask_only) {
      if (p.pen_sq == nullptr) continue;
      if (nj <= 2) fwd_row_iter_m<DH, 32, 2>(p, s2, kc, b, h, i, true, bound, lane, cx);
      else if (nj <= 4) fwd_row_iter_m<DH, 32, 4>(p, s2, kc, b, h, i, true, bound, lane, cx);
      else if (nj <= 6) fwd_row_iter_m<DH, 32, 6>(p, s2, kc, b, h, i, true, bound, lane, cx);
      else fwd_row_iter_m<DH, 32, 8>(p, s2, kc, b, h, i, true, bound, lane, cx);
      continue;
    }
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
  RowRegs<DH> rq, rq2, rt0, rt1;
  auto fetch = [&](int i) {
    fetch_row<DH>(rq2, p.aq, b, h, i, L, p.d);
    if (!(rt >= 0 && rt != i)) {
      fetch_row<DH>(rq, p.mq, b, h, i, L, p.d);
      fetch_row<DH>(rt0, p.t0, b, h, i, L, p.d);
      fetch_row<DH>(rt1, p.t1, b, h, i, L, p.d);
    }
  };
  int t0 = next_task(sm.misc + 2, 1);
  if (t0 < L) fetch(L - 1 - t0);
  while (t0 < L) {
    const int i = L - 1 - t0;
    const int bound = min(i + 1, nkey);
    const int nj = (bound + 31) >> 5;
    const bool mask_only = rt >= 0 && rt != i;
    AttnSmem s2 = sm;
    s2.Q = wr - i * dhp;
    s2.Q2 = wr + dhp - i * dhp;
    BwdSmem<NS> b2 = bs;
    b2.sT0 = wr + 2 * dhp - i * dhp;
    b2.sT1 = wr + 3 * dhp - i * dhp;
    __syncwarp();
    put_row<DH>(wr + dhp, rq2, p.aq);
    if (!mask_only) {
      put_row<DH>(wr, rq, p.mq);
      put_row<DH>(wr + 2 * dhp, rt0, p.t0);
      put_row<DH>(wr + 3 * dhp, rt1, p.t1);
    }
    __syncwarp();
    t0 = next_task(sm.misc + 2, 1);
    if (t0 < L) fetch(L - 1 - t0);                   // next row's vectors: in flight while this row is processed
    if (mask_only) {
#define ACSR_LROW_M(NJV) bwd_row_iter_m<DH, 32, NJV, NS, true>(p, s2, b2, kc, dpen, b, h, i, true, bound, 0, lane, LP, wbuf)
      if (nj <= 2) ACSR_LROW_M(2);
      else if (nj <= 4) ACSR_LROW_M(4);
      else if (nj <= 6) ACSR_LROW_M(6);
      else ACSR_LROW_M(8);
#undef ACSR_LROW_M
      continue;
    }
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
// backward, column phase: out[j][:] = sum_{i >= j} X[i][j] * Y[i][:]  for the (2 NS + 2) workspace matrices.
// Lane = key column (groups of 32 columns, only the nkey columns in play), the rows of a column group are dealt
// round-robin to the 8 warps -- so short sequences (few columns, all L rows: padded query rows still carry the attack
// mask's penalty gradient) keep every warp busy -- and the warps' partial sums meet in a transposed [dh][L] shared
// accumulator (lane = column: conflict-free shared atomics).  Y = the [L,dh] head slice of Q, Q', t0 or t1 is staged
// in shared memory once per product and read as float4 broadcasts.
// ------------------------------------------------------------------------------------------------------------
constexpr int kColThreads = 256;
__host__ __device__ static inline int cols_lpo(int L) { return ((L + 31) & ~31) + 1; }

template <int DH, int NS>
__global__ void __launch_bounds__(kColThreads, 2) attn_long_bwd_cols_kernel(const AttnParams p, const float* __restrict__ ws, const int b0) {
  extern __shared__ __align__(16) float smem_f[];
  const int L = p.L;
  const int LPO = cols_lpo(L);
  float* Y = smem_f;                                // [L][DH]
  float* OUT = smem_f + (size_t)((L + 3) & ~3) * DH;  // [DH][LPO]
  const int g = b0 + blockIdx.x / p.H;
  const int b = p.order ? p.order[g] : g, h = blockIdx.x % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int s_nkey;
  if (threadIdx.x == 0) s_nkey = 1;
  __syncthreads();
  for (int j = threadIdx.x; j < L; j += kColThreads)
    if (p.item_seq[(long long)b * L + j] != 0) atomicMax(&s_nkey, j + 1);
  __syncthreads();
  const int nkey = s_nkey;
  const int ng = (nkey + 31) >> 5;
  const size_t LL = (size_t)L * L;
  const float* slice = ws + (size_t)blockIdx.x * long_ws_floats(L, NS);
  const float* g_col = slice + (2 * NS + 2) * LL;
  const bool t1_att = NS == 1 ? true : (p.t1_is_att != 0);

  auto clear = [&]() {
    for (int e = threadIdx.x; e < DH * LPO; e += kColThreads) OUT[e] = 0.f;
  };
  // OUT += X^T . Y   (Y = head slice of ysrc; NULL: nothing to add).  Ends with a barrier.
  auto product = [&](const float* X, const float* ysrc) {
    if (ysrc == nullptr) { __syncthreads(); return; }            // uniform over the CTA
    const float* yb = ysrc + (long long)b * L * p.d + h * DH;
#pragma unroll 4
    for (int e = threadIdx.x; e < L * (DH / 4); e += kColThreads) {
      const int r = e / (DH / 4), c4 = e % (DH / 4);
      *reinterpret_cast<float4*>(Y + r * DH + c4 * 4) = __ldg(reinterpret_cast<const float4*>(yb + (long long)r * p.d + c4 * 4));
    }
    __syncthreads();
    for (int cg = 0; cg < ng; ++cg) {
      const int j = cg * 32 + lane;
      const bool jv = j < nkey;
      float acc[DH];
#pragma unroll
      for (int c = 0; c < DH; ++c) acc[c] = 0.f;
      const float* xc = X + (jv ? j : 0);
      for (int i = cg * 32 + warp; i < L; i += 4 * kAttnWarps) {
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int iu = i + u * kAttnWarps;
          x[u] = (jv && iu < L && iu >= j) ? xc[(size_t)iu * L] : 0.f;     // entries above the diagonal are never written
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int iu = i + u * kAttnWarps;
          if (iu >= L) break;
          const float4* y4 = reinterpret_cast<const float4*>(Y + iu * DH);
#pragma unroll
          for (int c4 = 0; c4 < DH / 4; ++c4) {
            const float4 y = y4[c4];
            acc[4 * c4 + 0] = fmaf(x[u], y.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(x[u], y.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(x[u], y.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(x[u], y.w, acc[4 * c4 + 3]);
          }
        }
      }
      if (jv) {
#pragma unroll
        for (int c = 0; c < DH; ++c) atomicAdd(OUT + c * LPO + j, acc[c]);
      }
    }
    __syncthreads();
  };
  // out rows of stream s <- OUT (+ the rank-1 part of the spatial calibrator for dK); columns behind the last key get zeros
  auto flush = [&](float* out, int s, bool rank1) {
    for (int e = threadIdx.x; e < L * (DH / 4); e += kColThreads) {
      const int j = e / (DH / 4), c4 = e % (DH / 4);
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (j < nkey) {
        float cdu = 0.f, cdt = 0.f;
        if (rank1) { cdu = g_col[(2 * s + 0) * L + j]; cdt = g_col[(2 * s + 1) * L + j]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = 4 * c4 + u;
          v[u] = OUT[c * LPO + j];
          if (rank1 && p.ow) v[u] = fmaf(cdu, __ldg(p.ow + DH + c), v[u]);
          if (rank1 && p.dw) v[u] = fmaf(cdt, __ldg(p.dw + DH + c), v[u]);
        }
      }
      *reinterpret_cast<float4*>(out + s * p.s1_td + ((long long)b * L + j) * p.d + h * DH + 4 * c4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
  };

#pragma unroll
  for (int s = 0; s < NS; ++s) {
    clear();                                         // (ordered before the atomics by the barrier inside product)
    product(slice + (2 * s) * LL, p.mq);             // dK_s = dS_s^T Q + colDU_s (x) wo[dh:] + colDT_s (x) wd[dh:]
    flush(p.d_mk, s, true);
    clear();
    product(slice + (2 * s + 1) * LL, p.aq);         // dK'_s = dS'_s^T Q'
    flush(p.d_ak, s, false);
  }
  // dV: stream 0 <- R^T t0 ; last stream <- (A or R)^T t1
  const float* XR = slice + (2 * NS) * LL;
  const float* XA = slice + (2 * NS + 1) * LL;
  clear();
  product(XR, p.t0);
  if (NS == 1) {
    product(t1_att ? XA : XR, p.t1);
    flush(p.d_mv, 0, false);
  } else {
    flush(p.d_mv, 0, false);
    clear();
    product(t1_att ? XA : XR, p.t1);
    flush(p.d_mv, NS - 1, false);
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
  const size_t smem_c = ((size_t)((p.L + 3) & ~3) * DH + (size_t)DH * cols_lpo(p.L)) * sizeof(float);
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
    attn_long_bwd_rows_kernel<DH, NS><<<dim3(nb * p.H), dim3(kAttnThreads), smem_r, st>>>(p, ws, b0);
    attn_long_bwd_cols_kernel<DH, NS><<<dim3(nb * p.H), dim3(kColThreads), smem_c, st>>>(p, ws, b0);
  }
  return check_launch(who);
}

}  // namespace acsr
