// Fused calibrated causal attention of one AC layer -- shared device code of attn_fwd.cu / attn_bwd.cu.
//
// One CTA owns one (sequence b, head h), L <= 64.  The projected [L,dh] tiles are staged in shared
// memory with 16-byte cp.async copies (rows padded to dh+4 floats: 16-byte aligned and bank-conflict
// free for both "lane = key" float4 reads and "lane = channel" vector reads).  A group of G lanes owns
// one query row and its lanes own the key columns j = sub, sub+G, ..., so the chained softmaxes
// (spatially-calibrated P, attack mask M, attacked A, calibrated C, combined R) are log2(G)-step
// shuffle reductions that serve 32/G rows per instruction, and none of the [B,H,L,L] intermediates of
// the reference (layers.py:686-742, 657-674, 917-925) ever reaches HBM.  The additive mask
// (abstract_recommender.py:136-143) is derived from item_seq, the spatial-calibrator affine over
// cat(q_i,k_j) is evaluated in its rank-1 form, dropout masks and the attack noise come from Philox
// (one call per column pair; or from explicit tensors in parity mode), and the penalty sum (1-M)^2
// (acsasrec.py:135) is reduced in the same pass.
//
// Exact work skipping.  With nkey = 1 + (index of the last non-padding key), row i only has the
// columns j < min(i+1, nkey) in play: every other column is masked, and a masked probability is
// exactly 0 in the reference too (exp(-10000 - max) underflows in fp32).  Those columns are never
// touched -- not staged, not computed -- except for their constant contribution 1 to the penalty.
// The number of G-wide column groups of a row (NJ) is a template parameter of the row code, chosen
// per row at run time; sequences with nkey <= 8 run with G = 8 (four rows per warp instruction),
// all others with G = 16.
// Right-padded input is assumed (sequential_dataset.py:128-132: position 0 holds a real item);
// a row whose allowed keys are all padding gets a finite softmax over its allowed columns instead
// of the reference's softmax over all L masked columns.
#pragma once
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

constexpr int kAttnWarps = 8;
constexpr int kAttnThreads = kAttnWarps * 32;
constexpr unsigned kFull = 0xffffffffu;

struct AttnParams {
  const float *mq, *mk, *mv, *aq, *ak, *gate;
  const int64_t* item_seq;
  const int32_t* order;      // sequence handled by CTA group g is order[g] (longest first), NULL = identity
  const int64_t* ctx_rows;   // NULL, or [B]: only context row ctx_rows[b]-1 is consumed (last layer); the other rows only
                             // contribute their attack mask to the penalty
  const float *ow, *ob, *dw, *db, *scalar;
  int B, L, H, dh, d;
  int two_level, combine, rich;
  int bidir;                 // bidirectional attention mask (AcBERT4Rec, get_attention_mask(bidirectional=True)): keys j > i stay in play
  int plain;                 // transformer_layers.py:873-953 (ACSSEPT): A, C and the combined attention are NOT re-normalised by a
                             // masked softmax (layers.py:917-925 does); the attacked weights of masked keys are then pure noise
  int full;                  // bidir | plain: all keys of a row are in play (no causal work skipping, dense [L][LP] matrices)
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
  // ACTiSASRec (timeaware.cu): additive raw-score bias [B,H,L,L] (q.posK + q.timeK), the attacked / final calibrated attention
  // written out [B,H,L,L] so that the time-aware context terms can be formed, and in the backward their cotangents coming
  // back plus the gradient of the bias going out.  All optional; plain variant only.
  const float* s_bias;
  float *prob_att_out, *prob_cal_out;
  const float *dprob_att, *dprob_cal;
  float* d_s_bias;
  // backward: two cotangent tiles.  One stream: t0 = d_ctx_cal, t1 = d_ctx_att.  Two streams: t0 = stream 0's
  // d_ctx_cal, t1 = stream 1's d_ctx_cal or (t1_is_att) d_ctx_att.
  const float *t0, *t1;
  int t1_is_att;
  const float *d_pen0, *d_pen1;
  long long s1_td, s1_ll;   // element offset of stream 1 inside the [.,d] outputs / inside d_gate
  float *d_mq, *d_mk, *d_mv, *d_aq, *d_ak, *d_gate, *d_ow, *d_ob, *d_dw, *d_db, *d_scalar, *d_ratio;
};

// ---- group (G lanes) reductions ----
template <int G>
__device__ __forceinline__ float grp_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ float grp_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// lane <-> head-dim mapping for the "lane = channel" loops of a G-lane group.  DH >= G: each lane owns
// CPL contiguous channels; DH < G: the group splits into G/DH parts, each reducing a slice of the range.
template <int DH, int G>
struct CMap {
  static constexpr int CPL = DH >= G ? DH / G : 1;
  static constexpr int SPL = DH >= G ? 1 : G / DH;
  __device__ static __forceinline__ int c0(int sub) { return DH >= G ? sub * CPL : sub % DH; }
  __device__ static __forceinline__ int split(int sub) { return DH >= G ? 0 : sub / DH; }
  __device__ static __forceinline__ float reduce(float v) {
#pragma unroll
    for (int o = DH; o < G; o <<= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
  }
  // slice [lo,hi) of [begin,end) owned by this lane's part; begin, lo multiples of 4
  __device__ static __forceinline__ void slice(int sub, int begin, int end, int& lo, int& hi) {
    if (SPL == 1) { lo = begin; hi = end; return; }
    const int chunk = (((end - begin + SPL - 1) / SPL) + 3) & ~3;
    lo = begin + split(sub) * chunk;
    hi = min(end, lo + chunk);
  }
};

template <int N> struct VecLd;
template <> struct VecLd<1> {
  __device__ static __forceinline__ void ld(const float* p, float* v) { v[0] = p[0]; }
  __device__ static __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; }
};
template <> struct VecLd<2> {
  __device__ static __forceinline__ void ld(const float* p, float* v) { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
  __device__ static __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <> struct VecLd<4> {
  __device__ static __forceinline__ void ld(const float* p, float* v) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ static __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct VecLd<8> {
  __device__ static __forceinline__ void ld(const float* p, float* v) { VecLd<4>::ld(p, v); VecLd<4>::ld(p + 4, v + 4); }
  __device__ static __forceinline__ void st(float* p, const float* v) { VecLd<4>::st(p, v); VecLd<4>::st(p + 4, v + 4); }
};

// per-lane state of one query row after the forward recomputation (NJ columns per lane)
template <int NJ>
struct RowF {
  float Psoft[NJ], Msoft[NJ], P0soft[NJ];
  float D1[NJ], D2[NJ], D3[NJ], nz[NJ];
  float A[NJ], C[NJ], g[NJ], F[NJ], R[NJ];
  float sig[NJ], delta[NJ];
  unsigned act;      // bit jj: column j = sub + G*jj takes part in this row's softmaxes (j < bound)
  unsigned valid;    // bit jj: ... and its key is a real item (additive mask 0)
};

struct AttnSmem {
  float *Q, *K, *V, *Q2, *K2;                        // [LP][dh+4]
  float *rowO, *rowD, *colO, *colD, *logd, *keyok;   // [LP]
  float *wo, *wd;                                    // [2*dh] spatial-calibrator weights (0 when absent)
  const float* G;                                    // gate-logit tile [L*L] staged in shared memory, or nullptr (read from global)
  int* misc;                                         // [8] 0,1: ballot words of the key-validity scan; 2,3: row / column task
                                                     // counters (dynamic heaviest-first scheduling); 4: warp arrival counter
};

struct RowConst {      // per-launch scalars hoisted out of the row loop
  float ob, db, sc, sc2h, rr, inv_sq, inv_keep;
  unsigned thr16;
  unsigned long long seed, step;
  bool philox_drop, philox_noise, need_p0;
  bool bounded;      // dropout scaling <= 16: the logits of the C and R softmaxes are bounded (no max subtraction)
};

// fast-math forms (ex2/lg2/rcp approx, ~2 ulp): far inside the 1e-3 parity budget
__device__ __forceinline__ float fex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float flg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fexp(float x) { return fex2(x * 1.4426950408889634f); }
__device__ __forceinline__ float flog(float x) { return flg2(x) * 0.6931471805599453f; }
__device__ __forceinline__ float fsigmoid(float x) { return frcp(1.0f + fex2(x * -1.4426950408889634f)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// softmax over the active columns of a row (at least one column is always active)
template <int G, int NJ>
__device__ __forceinline__ void softmax_row(const float* z, unsigned act, float* y) {
  float m = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) if ((act >> jj) & 1u) m = fmaxf(m, z[jj]);
  m = grp_max<G>(m) * 1.4426950408889634f;
  float s = 0.f;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) { y[jj] = ((act >> jj) & 1u) ? fex2(fmaf(z[jj], 1.4426950408889634f, -m)) : 0.f; s += y[jj]; }
  s = grp_sum<G>(s);
  const float inv = frcp(s);
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) y[jj] *= inv;
}

// softmax of logits known to lie in [0, ~45] (or at the mask value): no running maximum is needed.  A row whose
// columns are all masked (only possible for all-padding input) gets zeros.
template <int G, int NJ>
__device__ __forceinline__ void softmax_row_bounded(const float* z, unsigned act, float* y) {
  float s = 0.f;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) { y[jj] = ((act >> jj) & 1u) ? fex2(z[jj] * 1.4426950408889634f) : 0.f; s += y[jj]; }
  s = grp_sum<G>(s);
  const float inv = s > 0.f ? frcp(s) : 0.f;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) y[jj] *= inv;
}

// Y .* (dY - sum(Y .* dY))   (inactive columns have Y == 0)
template <int G, int NJ>
__device__ __forceinline__ void softmax_bwd_row(const float* y, const float* dy, float* dz) {
  float s = 0.f;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) s = fmaf(y[jj], dy[jj], s);
  s = grp_sum<G>(s);
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) dz[jj] = y[jj] * (dy[jj] - s);
}

// next task of a dynamically scheduled loop: tasks are handed out in order (heaviest first), `step` at a time
__device__ __forceinline__ int next_task(int* counter, int step) {
  int t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd(counter, step);
  return __shfl_sync(kFull, t, 0);
}

// packed lower-triangular storage of a transposed [j][i] matrix: row j keeps i in [j & ~3, LP)
__host__ __device__ __forceinline__ int tri_off(int j, int LP) {
  const int a = j >> 2, b = j & 3;
  return j * LP - 8 * a * (a - 1) - 4 * a * b;
}

// offset of row j of a transposed [j][i] shared-memory matrix such that element i sits at mat_row(...) + i: packed lower
// triangular for causal attention (i >= j & ~3), dense [L][LP] for the bidirectional mask
__device__ __forceinline__ int mat_row(const AttnParams& p, int j, int LP) { return p.full ? j * LP : tri_off(j, LP) - (j & ~3); }
__host__ __device__ __forceinline__ int mat_floats(int bidir, int L, int LP) { return bidir ? L * LP : tri_off(L, LP); }

// async copy of rows [0,rows) of one [L, DH] head slice into a padded smem tile; rows [rows, rows_pad) are zeroed
template <int DH>
__device__ __forceinline__ void stage_tile(float* dst, const float* __restrict__ src, int b, int h, int L, int d, int rows,
                                           int rows_pad) {
  constexpr int dhp = DH + 4;
  if (src == nullptr) rows = 0;
  const float* base = src + (long long)b * L * d + h * DH;
  for (int e = threadIdx.x; e < rows_pad * (DH / 4); e += blockDim.x) {
    const int r = e / (DH / 4), c4 = e % (DH / 4);
    if (r < rows) cp_async16(dst + r * dhp + c4 * 4, base + (long long)r * d + c4 * 4);
    else *reinterpret_cast<float4*>(dst + r * dhp + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// stage tiles + per-row / per-column scalars of the spatial calibrator; returns nkey (>= 1)
template <int DH>
__device__ __forceinline__ int stage_common(const AttnParams& p, const AttnSmem& sm, int b, int h, int LP) {
  constexpr int dhp = DH + 4;
  const int L = p.L;
  if (threadIdx.x < 64) {
    const int j = threadIdx.x;
    const bool ok = j < L && p.item_seq[(long long)b * L + j] != 0;
    const unsigned m = __ballot_sync(kFull, ok);
    if ((threadIdx.x & 31) == 0) sm.misc[threadIdx.x >> 5] = (int)m;
    if (threadIdx.x >= 2 && threadIdx.x < 5) sm.misc[threadIdx.x] = 0;
    if (j < LP) { sm.keyok[j] = ok ? 1.0f : 0.0f; sm.logd[j] = logf((float)j + 1.0f); }
  }
  stage_tile<DH>(sm.Q, p.mq, b, h, L, p.d, L, LP);
  stage_tile<DH>(sm.Q2, p.aq, b, h, L, p.d, L, LP);
  for (int c = threadIdx.x; c < 2 * DH; c += blockDim.x) {
    sm.wo[c] = p.ow ? p.ow[c] : 0.f;
    sm.wd[c] = p.dw ? p.dw[c] : 0.f;
  }
  __syncthreads();
  const unsigned m0 = (unsigned)sm.misc[0], m1 = (unsigned)sm.misc[1];
  int nkey = m1 ? 64 - __clz(m1) : (m0 ? 32 - __clz(m0) : 0);
  nkey = max(nkey, 1);
  if (p.plain && p.full) nkey = L;       // masked keys keep a weight (the attack noise): every key row is staged
  const int nkp = (nkey + 3) & ~3;
  stage_tile<DH>(sm.K, p.mk, b, h, L, p.d, nkey, nkp);
  stage_tile<DH>(sm.V, p.mv, b, h, L, p.d, nkey, nkp);
  stage_tile<DH>(sm.K2, p.ak, b, h, L, p.d, nkey, nkp);
  cp_async_wait_all();
  __syncthreads();
  // rank-1 pieces of the affines: 4 threads per row, DH/4 channels each
  if (p.ow || p.dw) {
    const int j = threadIdx.x >> 2, part = threadIdx.x & 3;
    float ro = 0.f, rd = 0.f, co = 0.f, cd = 0.f;
    if (j < L) {
#pragma unroll
      for (int cc = 0; cc < DH / 4; ++cc) {
        const int c = part * (DH / 4) + cc;
        const float q = sm.Q[j * dhp + c];
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
    __syncthreads();
  }
  return nkey;
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
  k.thr16 = (unsigned)(p.p * 65536.0f + 0.5f);
  k.bounded = k.inv_keep <= 16.0f;
  k.philox_drop = p.p > 0.f && p.D1 == nullptr;
  k.philox_noise = need_att && p.noise == nullptr && p.rng != nullptr;
  k.need_p0 = !p.two_level || p.probs != nullptr;
  k.seed = 0; k.step = 0;
  if (p.rng != nullptr) { k.seed = p.rng->seed; k.step = p.rng->step; }
  return k;
}

// Forward of query row i for one lane (columns j = sub + G*jj, jj < NJ).  bound = number of columns in play for this
// row; every one of the NJ column groups holds an active column of some row of the warp.
template <int DH, int G, int NJ>
__device__ __forceinline__ void row_forward(const AttnParams& p, const AttnSmem& sm, const RowConst& kc, int b, int h, int i,
                                            int bound, int sub, bool need_att, RowF<NJ>& r) {
  constexpr int dhp = DH + 4;
  const int L = p.L;
  int jc[NJ];
  float S[NJ], S2[NJ], gl[NJ];
  unsigned act = 0;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const int j = sub + G * jj;
    const bool a = j < bound;
    act |= (a ? 1u : 0u) << jj;
    jc[jj] = a ? j : 0;
    gl[jj] = 0.f;
  }
  r.act = act;
  const long long ebase = (((long long)b * p.H + h) * L + i) * L;
  if (p.combine == ACSR_ATTN_COMBINE_GATE) {     // gate logits straight from global (read once per element), early
    if (sm.G != nullptr) {
      const float* gp = sm.G + i * L;
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) if ((act >> jj) & 1u) gl[jj] = gp[jc[jj]];
    } else {
      const float* gp = p.gate + ((long long)b * L + i) * L;
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) if ((act >> jj) & 1u) gl[jj] = __ldg(gp + jc[jj]);
    }
  }
  {
    const float4* qi = reinterpret_cast<const float4*>(sm.Q + i * dhp);
    const float4* q2i = reinterpret_cast<const float4*>(sm.Q2 + i * dhp);
    if (NJ == 1) {
      const float4* kj = reinterpret_cast<const float4*>(sm.K + jc[0] * dhp);
      const float4* k2j = reinterpret_cast<const float4*>(sm.K2 + jc[0] * dhp);
      float s = 0.f, s2 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) { s = dot4(qi[c4], kj[c4], s); s2 = dot4(q2i[c4], k2j[c4], s2); }
      S[0] = s; S2[0] = s2;
    } else {
      {
        float4 q[DH / 4];
#pragma unroll
        for (int c4 = 0; c4 < DH / 4; ++c4) q[c4] = qi[c4];
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const float4* kj = reinterpret_cast<const float4*>(sm.K + jc[jj] * dhp);
          float s = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < DH / 4; ++c4) s = dot4(q[c4], kj[c4], s);
          S[jj] = s;
        }
      }
      {
        float4 q[DH / 4];
#pragma unroll
        for (int c4 = 0; c4 < DH / 4; ++c4) q[c4] = q2i[c4];
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const float4* kj = reinterpret_cast<const float4*>(sm.K2 + jc[jj] * dhp);
          float s = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < DH / 4; ++c4) s = dot4(q[c4], kj[c4], s);
          S2[jj] = s;
        }
      }
    }
  }
  if (p.s_bias != nullptr) {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) if ((act >> jj) & 1u) S[jj] += __ldg(p.s_bias + ebase + jc[jj]);
  }
  const float rowO = sm.rowO[i], rowD = sm.rowD[i];
  float zP[NJ], z0[NJ], zM[NJ];
  unsigned valid = 0;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const int j = jc[jj];
    const bool a = (act >> jj) & 1u;
    const bool v = a && (p.bidir || j <= i) && (sm.keyok[j] != 0.f);
    valid |= (v ? 1u : 0u) << jj;
    const float msk = v ? 0.f : kMaskNeg;
    r.sig[jj] = 0.f; r.delta[jj] = 0.f;
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
    zP[jj] = (S[jj] + eo + ed) * kc.inv_sq + msk;
    z0[jj] = S[jj] * kc.inv_sq + msk;
    zM[jj] = S2[jj] * kc.inv_sq + msk;
  }
  r.valid = valid;
  softmax_row<G, NJ>(zP, act, r.Psoft);
  softmax_row<G, NJ>(zM, act, r.Msoft);
  if (kc.need_p0) softmax_row<G, NJ>(z0, act, r.P0soft);
  else {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) r.P0soft[jj] = 0.f;
  }
  // ---- randomness: one Philox call per column pair (jj, jj+1) ----
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) { r.D1[jj] = r.D2[jj] = r.D3[jj] = 1.0f; r.nz[jj] = 0.f; }
  if (kc.philox_drop || kc.philox_noise) {
#pragma unroll
    for (int jj = 0; jj < NJ; jj += 2) {
      const uint4 w = philox4x32(kc.seed, kc.step, p.stream, (unsigned long long)(ebase + jc[jj]));
      if (kc.philox_drop) {
        r.D1[jj] = ((w.x & 0xffffu) >= kc.thr16) ? kc.inv_keep : 0.f;
        r.D3[jj] = ((w.x >> 16) >= kc.thr16) ? kc.inv_keep : 0.f;
        if (jj + 1 < NJ) {
          r.D1[jj + 1] = ((w.y & 0xffffu) >= kc.thr16) ? kc.inv_keep : 0.f;
          r.D3[jj + 1] = ((w.y >> 16) >= kc.thr16) ? kc.inv_keep : 0.f;
        }
      }
      if (kc.philox_noise) {
        const float rad = sqrtf(-2.0f * flog(u32_to_unit(w.z)));
        float sn, cs;
        __sincosf(6.283185307179586f * u32_to_unit(w.w), &sn, &cs);
        r.nz[jj] = rad * cs;
        if (jj + 1 < NJ) r.nz[jj + 1] = rad * sn;
      }
      if (kc.philox_drop && kc.need_p0) {
        const uint4 w2 = philox4x32(kc.seed, kc.step, p.stream + 1u, (unsigned long long)(ebase + jc[jj]));
        r.D2[jj] = ((w2.x & 0xffffu) >= kc.thr16) ? kc.inv_keep : 0.f;
        if (jj + 1 < NJ) r.D2[jj + 1] = ((w2.x >> 16) >= kc.thr16) ? kc.inv_keep : 0.f;
      }
    }
  }
  if (p.D1 || p.noise) {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      if (!((act >> jj) & 1u)) continue;
      const long long e = ebase + jc[jj];
      if (p.D1) r.D1[jj] = p.D1[e];
      if (p.D2) r.D2[jj] = p.D2[e];
      if (p.D3) r.D3[jj] = p.D3[e];
      if (p.noise) r.nz[jj] = p.noise[e];
    }
  }
  float O[NJ], M[NJ], expm[NJ], z[NJ];
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const bool a = (act >> jj) & 1u;
    M[jj] = r.Msoft[jj] * r.D3[jj];
    O[jj] = p.two_level ? r.Psoft[jj] * r.D1[jj] : r.P0soft[jj] * r.D2[jj];
    expm[jj] = a ? fexp(1.0f - M[jj]) : 0.f;
  }
  if (need_att) {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj)
      z[jj] = O[jj] * M[jj] + r.nz[jj] * (1.0f - M[jj]) + ((p.plain || ((valid >> jj) & 1u)) ? 0.f : kMaskNeg);
    if (p.plain) {              // transformer_layers.py:919: used as is
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) r.A[jj] = ((act >> jj) & 1u) ? z[jj] : 0.f;
    } else softmax_row<G, NJ>(z, act, r.A);
  } else {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) r.A[jj] = 0.f;
  }
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) z[jj] = O[jj] * expm[jj] + ((p.plain || ((valid >> jj) & 1u)) ? 0.f : kMaskNeg);
  if (p.plain) {                // transformer_layers.py:921
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) r.C[jj] = z[jj];
  } else if (kc.bounded) softmax_row_bounded<G, NJ>(z, act, r.C);
  else softmax_row<G, NJ>(z, act, r.C);
  if (p.combine == ACSR_ATTN_COMBINE_FIXED) {
    // layers.py:885: softmax(origin + 0.5*calibrated) has NO mask: each of the L - bound columns outside the row's
    // range holds exp(0 - max) of the mass; they only enter through the normaliser.
    const float ninact = (float)(L - bound);
    float m = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) { z[jj] = O[jj] + 0.5f * r.C[jj]; r.g[jj] = 0.f; if ((act >> jj) & 1u) m = fmaxf(m, z[jj]); }
    m = grp_max<G>(m);
    if (ninact > 0.f) m = fmaxf(m, 0.f);
    float s = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) { r.F[jj] = ((act >> jj) & 1u) ? fexp(z[jj] - m) : 0.f; s += r.F[jj]; }
    s = grp_sum<G>(s) + ninact * fexp(-m);
    const float inv = frcp(s);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) { r.F[jj] *= inv; z[jj] = r.F[jj] + (((valid >> jj) & 1u) ? 0.f : kMaskNeg); }
  } else {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      float g = p.comb_scalar;
      if (p.combine == ACSR_ATTN_COMBINE_GATE) g = fsigmoid(gl[jj]);
      r.g[jj] = g; r.F[jj] = 0.f;
      z[jj] = g * O[jj] + (1.0f - g) * r.C[jj] + ((p.plain || ((valid >> jj) & 1u)) ? 0.f : kMaskNeg);
    }
  }
  if (p.plain) {                // transformer_layers.py:897-908: the combination itself is the attention
    const bool fixed = p.combine == ACSR_ATTN_COMBINE_FIXED;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) r.R[jj] = fixed ? r.F[jj] : (((act >> jj) & 1u) ? z[jj] : 0.f);
    return;
  }
  if (kc.bounded) softmax_row_bounded<G, NJ>(z, act, r.R);
  else softmax_row<G, NJ>(z, act, r.R);
}

// Attack mask of query row i only (softmax M and its dropout): all a row contributes when its context is not consumed.
// Same column ownership and the same Philox words as row_forward.
template <int DH, int G, int NJ>
__device__ __forceinline__ void row_forward_m(const AttnParams& p, const AttnSmem& sm, const RowConst& kc, int b, int h, int i,
                                              int bound, int sub, float* Msoft, float* D3, unsigned& act_out) {
  constexpr int dhp = DH + 4;
  const int L = p.L;
  int jc[NJ];
  float zM[NJ];
  unsigned act = 0;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const int j = sub + G * jj;
    const bool a = j < bound;
    act |= (a ? 1u : 0u) << jj;
    jc[jj] = a ? j : 0;
  }
  act_out = act;
  const long long ebase = (((long long)b * p.H + h) * L + i) * L;
  const float4* q2i = reinterpret_cast<const float4*>(sm.Q2 + i * dhp);
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const float4* kj = reinterpret_cast<const float4*>(sm.K2 + jc[jj] * dhp);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; c4 += 2) { s0 = dot4(q2i[c4], kj[c4], s0); s1 = dot4(q2i[c4 + 1], kj[c4 + 1], s1); }
    const bool v = ((act >> jj) & 1u) && (p.bidir || jc[jj] <= i) && (sm.keyok[jc[jj]] != 0.f);
    zM[jj] = (s0 + s1) * kc.inv_sq + (v ? 0.f : kMaskNeg);
  }
  softmax_row<G, NJ>(zM, act, Msoft);
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) D3[jj] = 1.0f;
  if (kc.philox_drop) {
#pragma unroll
    for (int jj = 0; jj < NJ; jj += 2) {
      const uint4 w = philox4x32(kc.seed, kc.step, p.stream, (unsigned long long)(ebase + jc[jj]));
      D3[jj] = ((w.x >> 16) >= kc.thr16) ? kc.inv_keep : 0.f;
      if (jj + 1 < NJ) D3[jj + 1] = ((w.y >> 16) >= kc.thr16) ? kc.inv_keep : 0.f;
    }
  }
  if (p.D3) {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) if ((act >> jj) & 1u) D3[jj] = p.D3[ebase + jc[jj]];
  }
}

__device__ __forceinline__ AttnSmem carve_common(float*& ptr, int LP, int dh) {
  AttnSmem sm;
  const int tile = LP * (dh + 4);
  sm.Q = ptr; ptr += tile;
  sm.K = ptr; ptr += tile;
  sm.V = ptr; ptr += tile;
  sm.Q2 = ptr; ptr += tile;
  sm.K2 = ptr; ptr += tile;
  sm.rowO = ptr; ptr += LP;
  sm.rowD = ptr; ptr += LP;
  sm.colO = ptr; ptr += LP;
  sm.colD = ptr; ptr += LP;
  sm.logd = ptr; ptr += LP;
  sm.keyok = ptr; ptr += LP;
  sm.wo = ptr; ptr += 2 * dh;
  sm.wd = ptr; ptr += 2 * dh;
  sm.misc = reinterpret_cast<int*>(ptr); ptr += 8;
  sm.G = nullptr;
  return sm;
}
static inline size_t common_floats(int LP, int dh) { return (size_t)5 * LP * (dh + 4) + 6 * LP + 4 * dh + 8; }

// per-group row buffers: sequences with nkey <= 8 run four 8-lane groups per warp with 8-float rows, all others two
// 16-lane groups with LP-float rows; the allocation covers both
__host__ __device__ __forceinline__ int rowbuf_floats_per_warp(int LP, int bufs_per_row) {
  const int a = 2 * bufs_per_row * LP, b = 4 * bufs_per_row * 8;
  return a > b ? a : b;
}

template <typename K>
static int prep_kernel(K kernel, size_t smem, const char* who) {
  if (smem > 227 * 1024) { set_error("%s: needs %zu bytes of shared memory (> 227 KB)", who, smem); return ACSR_ERR_UNSUPPORTED; }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) { set_error("%s: smem %zu: %s", who, smem, cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  return ACSR_OK;
}

static inline int attn_validate(const AttnParams& p, const char* who) {
  ACSR_REQUIRE(p.mq && p.mk && p.mv && p.aq && p.ak && p.item_seq, "%s: NULL input", who);
  ACSR_REQUIRE(p.B > 0 && p.H > 0, "%s: bad B/H", who);
  if (p.L < 1 || p.L > 256) { set_error("%s: L=%d unsupported (1..256)", who, p.L); return ACSR_ERR_UNSUPPORTED; }
  if ((p.full || p.plain) && p.L > 64) { set_error("%s: the bidirectional mask / the transformer_layers variant are implemented for L <= 64 (L=%d)", who, p.L); return ACSR_ERR_UNSUPPORTED; }
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

// The plain variant needs every key of a row only because of the ATTACKED weights (bare noise on masked keys) and of the
// unmasked softmax of combine_option fixed.  A call without the attacked stream (evaluation, non-final layers, the calibrated
// pass of the routed backward) and with gate / annealing has origin = calibrated = combined = 0 on every masked key, so the
// exact causal work skipping of the layers.py kernels applies to it as well.
static inline void plain_range(AttnParams& p, bool attacked_stream) {
  if (p.plain && !p.bidir && !attacked_stream && p.combine != ACSR_ATTN_COMBINE_FIXED) p.full = 0;
}

// attn_long.cu: 64 < L <= 256 (key-side tiles resident, query rows streamed, backward matrices in a global workspace)
int attn_long_fwd(const AttnParams& p, cudaStream_t st);
int attn_long_bwd(const AttnParams& p, int ns, cudaStream_t st);
int seq_order_long(const int64_t* item_seq, int B, int L, int32_t* order, cudaStream_t st);

static inline void attn_fill_common(AttnParams& p, const float* mq, const float* mk, const float* mv, const float* aq,
                                    const float* ak, const float* gate_logit, const int64_t* item_seq, const float* order_w,
                                    const float* order_b, const float* dist_w, const float* dist_b, const float* scalar, int B,
                                    int L, int H, int dh, int two_level, int combine_option, float comb_scalar, int rich_mode,
                                    const float* rich_ratio, float p_attn, const float* D1, const float* D2, const float* D3,
                                    const float* noise, const void* rng, uint32_t rng_stream, const int32_t* order, const int64_t* ctx_rows) {
  p.order = order; p.ctx_rows = ctx_rows;
  p.mq = mq; p.mk = mk; p.mv = mv; p.aq = aq; p.ak = ak; p.gate = combine_option == ACSR_ATTN_COMBINE_GATE ? gate_logit : nullptr;
  p.item_seq = item_seq;
  p.ow = order_w; p.ob = order_b; p.dw = dist_w; p.db = dist_b; p.scalar = scalar;
  p.B = B; p.L = L; p.H = H; p.dh = dh; p.d = H * dh;
  p.two_level = two_level & 1; p.bidir = (two_level >> 1) & 1; p.plain = (two_level >> 2) & 1; p.full = p.bidir | p.plain; p.combine = combine_option; p.rich = rich_mode; p.comb_scalar = comb_scalar; p.rich_ratio = rich_ratio;
  p.p = p_attn; p.D1 = D1; p.D2 = D2; p.D3 = D3; p.noise = noise; p.rng = (const RngState*)rng; p.stream = rng_stream;
}

}  // namespace acsr
