// The dense part of an encoder layer after its attention, forward, as ONE tcgen05 kernel per 128-token tile:
//   hz = ctx.Wo^T ; h = LN(drop(hz + bo) + x) ; z1 = h.W1^T ; a1 = act(z1 + b1) ; z2 = a1.W2^T ; out = LN(drop(z2 + b2) + h)
// (layers.py:676-684 dense + dropout + residual + LayerNorm, layers.py:790-798 FeedForward), hidden size 64, inner size a
// multiple of 64 up to 256.  It replaces four launches (two GEMMs with fused LayerNorm epilogue, one GEMM, one activation
// kernel) whose intermediates each made a round trip through L2; here the activations stay in TENSOR MEMORY between the GEMMs:
// every epilogue leaves its result (split into hi / lo TF32 halves) in TMEM with tcgen05.st, and the next tcgen05.mma takes it
// as its A operand (TS form) -- the mechanism proven in logits_bwd.cu.  The feed-forward runs in 64-feature chunks of the inner
// axis (z1 chunk -> activation -> partial z2), so the [128, inner] activation never exists as a whole on the SM.
//
// TMEM columns: acc [0,64): stage-1 accumulator (hz), later the z2 accumulator | A [64,192): ctx hi / lo, overwritten by h hi / lo
//               | C[2] [192,320): z1 chunk, overwritten IN PLACE by a1 hi | L[2] [320,448): a1 lo
// Weights arrive pre-split: acsr_dense_prep writes Wo, the W1 chunks and the W2 chunks as (hi, lo) TF32 operands in the canonical
// K-major UMMA layout once per step (parameters change once per step), so staging a weight block is one 32 KB bulk copy and
// nothing in this kernel touches weights with ordinary loads.
// Roles (352 threads, one CTA per SM, persistent over tiles): warp 0 bulk-copy producer, warp 1 issues the stage-1 and stage-2
// MMAs, warp 2 the stage-3 MMAs (one thread issues an MMA every ~90 cycles; the tensor pipe takes them from several threads),
// warps 3-10 are the epilogue (two per TMEM lane quarter; one thread owns one token row, so LayerNorm / dropout / activation are
// thread-private); the upper four of them first load the ctx tile (one token row per thread) and store it to TMEM.
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

constexpr int kDfThreads = 352;
constexpr int kDfD = 64;
constexpr int kDfBM = 128;
constexpr int kDfOpFloats = 64 * 64;              // one 64 x 64 operand (hi or lo)
constexpr int kDfBlkBytes = 2 * kDfOpFloats * 4;  // hi + lo: 32 KB
constexpr int kDfMaxI = 256;

struct DfParams {
  const float* ctx; const float* res; long long rows, res_rows; int I, act, passes;
  const float* ops;
  const float *bo, *lnA_w, *lnA_b, *b1, *b2, *lnF_w, *lnF_b;
  float epsA, epsF, p_drop;
  const float *mask_a, *mask_f; const RngState* rng; uint32_t stream_a, stream_f;
  float *hz, *st_a, *h, *z1, *a1, *z2, *st_f, *out;
  int m_tiles;
};

// ---- weights -> (hi, lo) TF32 operands, canonical K-major layout: block 0 = Wo, 1..nc = W1 chunks, nc+1..2nc = W2 chunks ----
__global__ void __launch_bounds__(256) dense_prep_kernel(const float* __restrict__ Wo, const float* __restrict__ W1,
                                                         const float* __restrict__ W2, int I, float* __restrict__ ops) {
  pdl_wait();
  const int nc = I / 64, blk = blockIdx.x;
  float* hi = ops + (long long)blk * 2 * kDfOpFloats;
  float* lo = hi + kDfOpFloats;
  for (int e = threadIdx.x; e < 64 * 16; e += blockDim.x) {
    const int n = e >> 4, kc = e & 15;
    const float* src;
    if (blk == 0) src = Wo + n * 64 + kc * 4;
    else if (blk <= nc) src = W1 + (long long)((blk - 1) * 64 + n) * 64 + kc * 4;          // (n, k) = W1[c*64 + n][k]
    else src = W2 + (long long)n * I + (blk - 1 - nc) * 64 + kc * 4;                        // (n, k) = W2[n][c*64 + k]
    const float4 x = *reinterpret_cast<const float4*>(src);
    const float4 h4 = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
    const int off = kc * 256 + n * 4;
    *reinterpret_cast<float4*>(hi + off) = h4;
    *reinterpret_cast<float4*>(lo + off) = make_float4(x.x - h4.x, x.y - h4.y, x.z - h4.z, x.w - h4.w);
  }
}

struct DfCfg {
  static constexpr int kOffWo = 0;
  static constexpr int kOffW1 = kOffWo + kDfBlkBytes;          // [2]
  static constexpr int kOffW2 = kOffW1 + 2 * kDfBlkBytes;      // [2]
  static constexpr int kOffPar = kOffW2 + 2 * kDfBlkBytes;     // bo, b2, lnA_w, lnA_b, lnF_w, lnF_b [64 each], b1 [256]
  static constexpr int kParFloats = 6 * 64 + kDfMaxI;
  static constexpr int kOffBar = kOffPar + kParFloats * 4;
  static constexpr int kNumBars = 24;
  static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16;
  static constexpr int kColAcc = 0, kColA = 64, kColC = 192, kColL = 320;
};

__device__ __forceinline__ float df_drop(const DfParams& p, const float* mask, uint32_t stream, long long row, int q, float* m4) {
  // multipliers of columns 4q..4q+3 of `row`: same Philox counters as bdrl_{fwd,bwd}_kernel (element e -> call e>>2, word e&3)
  if (mask != nullptr) {
    const float4 mm = __ldg(reinterpret_cast<const float4*>(mask + row * kDfD + q * 4));
    m4[0] = mm.x; m4[1] = mm.y; m4[2] = mm.z; m4[3] = mm.w;
  } else if (p.p_drop > 0.f && p.rng != nullptr) {
    const float inv_keep = 1.0f / (1.0f - p.p_drop);
    const uint4 w = philox4x32(p.rng->seed, p.rng->step, stream, (unsigned long long)row * 16 + q);
    m4[0] = drop_mult(w.x, p.p_drop, inv_keep); m4[1] = drop_mult(w.y, p.p_drop, inv_keep);
    m4[2] = drop_mult(w.z, p.p_drop, inv_keep); m4[3] = drop_mult(w.w, p.p_drop, inv_keep);
  } else {
    m4[0] = m4[1] = m4[2] = m4[3] = 1.0f;
  }
  return 0.f;
}

__global__ void __launch_bounds__(kDfThreads, 1) dense_fwd_kernel(const DfParams p) {
  using Cfg = DfCfg;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nc = p.I / 64;
  float* sPar = reinterpret_cast<float*>(smem + Cfg::kOffPar);
  float *s_bo = sPar, *s_b2 = sPar + 64, *s_lnAw = sPar + 128, *s_lnAb = sPar + 192, *s_lnFw = sPar + 256, *s_lnFb = sPar + 320,
        *s_b1 = sPar + 384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* wo_full = bars;
  uint64_t* w1_full = bars + 1;      // [2]
  uint64_t* w1_empty = bars + 3;     // [2]
  uint64_t* w2_full = bars + 5;      // [2]
  uint64_t* w2_empty = bars + 7;     // [2]
  uint64_t* a_full = bars + 9;       // ctx tile sits in TMEM
  uint64_t* s1_full = bars + 10;     // hz accumulator ready
  uint64_t* h_full = bars + 11;      // h (hi, lo) sits in TMEM, acc free
  uint64_t* c_full = bars + 12;      // [2] z1 chunk ready
  uint64_t* a_ready = bars + 14;     // [2] a1 chunk (hi, lo) sits in TMEM
  uint64_t* c_empty = bars + 16;     // [2] stage 3 of the chunk retired
  uint64_t* s3_full = bars + 18;     // z2 accumulator ready
  uint64_t* tile_done = bars + 19;   // final epilogue has read acc and h: the tile's TMEM regions are free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(wo_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(w1_full + s, 1); mbar_init(w1_empty + s, 1); mbar_init(w2_full + s, 1); mbar_init(w2_empty + s, 1);
      mbar_init(c_full + s, 1); mbar_init(a_ready + s, 256); mbar_init(c_empty + s, 1);
    }
    mbar_init(a_full, 128); mbar_init(s1_full, 1); mbar_init(h_full, 128); mbar_init(s3_full, 1); mbar_init(tile_done, 128);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // parameters (never written inside a step): staged before the dependency wait
  for (int i = threadIdx.x; i < 64; i += kDfThreads) {
    s_bo[i] = p.bo[i]; s_b2[i] = p.b2[i]; s_lnAw[i] = p.lnA_w[i]; s_lnAb[i] = p.lnA_b[i]; s_lnFw[i] = p.lnF_w[i]; s_lnFb[i] = p.lnF_b[i];
  }
  for (int i = threadIdx.x; i < p.I; i += kDfThreads) s_b1[i] = p.b1[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (int)blockIdx.x < p.m_tiles ? (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ------------------------------ producer: prepared weight blocks, one bulk copy each ------------------------------
    if (lane == 0 && my_tiles > 0) {
      mbar_arrive_expect_tx(wo_full, kDfBlkBytes);
      bulk_g2s(smem + Cfg::kOffWo, p.ops, kDfBlkBytes, wo_full);
      for (int it = 0; it < my_tiles; ++it)
        for (int c = 0; c < nc; ++c) {
          const int g = it * nc + c, b = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          mbar_wait(w1_empty + b, ph ^ 1);
          mbar_arrive_expect_tx(w1_full + b, kDfBlkBytes);
          bulk_g2s(smem + Cfg::kOffW1 + b * kDfBlkBytes, p.ops + (long long)(1 + c) * 2 * kDfOpFloats, kDfBlkBytes, w1_full + b);
          mbar_wait(w2_empty + b, ph ^ 1);
          mbar_arrive_expect_tx(w2_full + b, kDfBlkBytes);
          bulk_g2s(smem + Cfg::kOffW2 + b * kDfBlkBytes, p.ops + (long long)(1 + nc + c) * 2 * kDfOpFloats, kDfBlkBytes, w2_full + b);
        }
    }
  } else if (warp == 1 || warp == 2) {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = umma_idesc_tf32(kDfBM, 64);
      constexpr uint32_t kLbo = 64 * 16, kSbo = 128;
      const int npass = p.passes == 3 ? 3 : 1;
      const uint32_t a_hi = tmem_base + Cfg::kColA, a_lo = a_hi + 64;
      auto mma24 = [&](uint32_t d_tmem, uint32_t ahi, uint32_t alo, uint32_t bsm, bool first_zero) {
        uint32_t acc = first_zero ? 0u : 1u;
        for (int ps = 0; ps < npass; ++ps) {
          const uint32_t a_base = (npass == 3 && ps == 0) ? alo : ahi;
          const uint32_t b_base = bsm + ((npass == 3 && ps == 1) ? kDfOpFloats * 4 : 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            umma_tf32_ts(d_tmem, a_base + ks * 8, umma_desc_kmajor(b_base + ks * 2 * kLbo, kLbo, kSbo), idesc, acc);
            acc = 1;
          }
        }
      };
      if (warp == 1) {
        // ---- issuer A: stage 1 (hz = ctx.Wo^T) and stage 2 (z1 chunk = h.W1_c^T) ----
        mbar_wait(wo_full, 0);
        for (int it = 0; it < my_tiles; ++it) {
          mbar_wait(a_full, it & 1);            // (the loaders stored the tile after tile_done of the previous one: acc is free too)
          tc_fence_after();
          mma24(tmem_base + Cfg::kColAcc, a_hi, a_lo, smem_u32(smem + Cfg::kOffWo), true);
          umma_commit(s1_full);
          mbar_wait(h_full, it & 1);
          for (int c = 0; c < nc; ++c) {
            const int g = it * nc + c, b = g & 1;
            const uint32_t ph = (g >> 1) & 1;
            mbar_wait(w1_full + b, ph);
            mbar_wait(c_empty + b, ph ^ 1);
            tc_fence_after();
            mma24(tmem_base + Cfg::kColC + b * 64, a_hi, a_lo, smem_u32(smem + Cfg::kOffW1 + b * kDfBlkBytes), true);
            umma_commit(w1_empty + b);
            umma_commit(c_full + b);
          }
        }
      } else {
        // ---- issuer B: stage 3 (z2 += a1 chunk . W2_c^T) ----
        for (int it = 0; it < my_tiles; ++it) {
          for (int c = 0; c < nc; ++c) {
            const int g = it * nc + c, b = g & 1;
            const uint32_t ph = (g >> 1) & 1;
            mbar_wait(a_ready + b, ph);
            mbar_wait(w2_full + b, ph);
            tc_fence_after();
            mma24(tmem_base + Cfg::kColAcc, tmem_base + Cfg::kColC + b * 64, tmem_base + Cfg::kColL + b * 64,
                  smem_u32(smem + Cfg::kOffW2 + b * kDfBlkBytes), c == 0);
            umma_commit(w2_empty + b);
            umma_commit(c_empty + b);
          }
          umma_commit(s3_full);
        }
      }
    }
  } else if (warp < 11) {
    // ------------------------------ epilogue: one thread = one token row ------------------------------
    pdl_wait();
    const int quarter = warp & 3;
    const int half = (warp - 3) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int it = 0; it < my_tiles; ++it) {
      const long long grow = (long long)(blockIdx.x + it * gridDim.x) * kDfBM + row;
      const bool row_ok = grow < p.rows;
      if (half == 1) {
        // ---- ctx tile -> TMEM (hi, lo), the A operand of stage 1: one token row per thread ----
        float4 cx[16];
#pragma unroll
        for (int q = 0; q < 16; ++q)
          cx[q] = row_ok ? __ldg(reinterpret_cast<const float4*>(p.ctx + grow * kDfD + q * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(tile_done, (it & 1) ^ 1);       // the previous tile no longer needs its h (same TMEM columns)
        tc_fence_after();
#pragma unroll
        for (int hcol = 0; hcol < 2; ++hcol) {
          float hi[32], lo[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v = cx[hcol * 8 + q];
            hi[4 * q] = to_tf32(v.x); hi[4 * q + 1] = to_tf32(v.y); hi[4 * q + 2] = to_tf32(v.z); hi[4 * q + 3] = to_tf32(v.w);
            lo[4 * q] = v.x - hi[4 * q]; lo[4 * q + 1] = v.y - hi[4 * q + 1]; lo[4 * q + 2] = v.z - hi[4 * q + 2]; lo[4 * q + 3] = v.w - hi[4 * q + 3];
          }
          tmem_st32(t_lane + Cfg::kColA + hcol * 32, hi);
          tmem_st32(t_lane + Cfg::kColA + 64 + hcol * 32, lo);
        }
        tc_fence_before();
        mbar_arrive(a_full);
      }
      if (half == 0) {
        // ---- hz -> h = LN(drop(hz + bo) + x): h to global and, split, back to TMEM as the A operand of stage 2 ----
        mbar_wait(s1_full, it & 1);
        tc_fence_after();
        float x[64];
        tmem_ld32(t_lane + Cfg::kColAcc, x);
        tmem_ld32(t_lane + Cfg::kColAcc + 32, x + 32);
        float s = 0.f;
        if (row_ok) {
          float* hz = p.hz + grow * kDfD;
#pragma unroll
          for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(hz + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
          const float* rr = p.res + (grow % p.res_rows) * kDfD;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float m4[4];
            df_drop(p, p.mask_a, p.stream_a, grow, q, m4);
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(rr + q * 4));
            const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int c = q * 4 + v;
              x[c] = (x[c] + s_bo[c]) * m4[v] + rv[v];
              s += x[c];
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) x[c] = 0.f;
        }
        const float mean = s * (1.0f / 64);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) { const float t = x[c] - mean; var = fmaf(t, t, var); }
        const float rstd = 1.0f / sqrtf(var * (1.0f / 64) + p.epsA);
#pragma unroll
        for (int c = 0; c < 64; ++c) x[c] = row_ok ? (x[c] - mean) * rstd * s_lnAw[c] + s_lnAb[c] : 0.f;
        if (row_ok) {
          float* ho = p.h + grow * kDfD;
#pragma unroll
          for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(ho + c) = make_float4(x[c], x[c + 1], x[c + 2], x[c + 3]);
          p.st_a[2 * grow] = mean; p.st_a[2 * grow + 1] = rstd;
        }
#pragma unroll
        for (int hcol = 0; hcol < 2; ++hcol) {
          float hi[32], lo[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { hi[i] = to_tf32(x[hcol * 32 + i]); lo[i] = x[hcol * 32 + i] - hi[i]; }
          tmem_st32(t_lane + Cfg::kColA + hcol * 32, hi);
          tmem_st32(t_lane + Cfg::kColA + 64 + hcol * 32, lo);
        }
        tc_fence_before();
        mbar_arrive(h_full);
      }
      // ---- z1 chunk -> a1 = act(z1 + b1): both to global, a1 split back to TMEM (hi in place, lo next to it) ----
      for (int c = 0; c < nc; ++c) {
        const int g = it * nc + c, b = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        mbar_wait(c_full + b, ph);
        tc_fence_after();
        float v[32], lo[32];
        tmem_ld32(t_lane + Cfg::kColC + b * 64 + half * 32, v);
        const int col0 = c * 64 + half * 32;
        if (row_ok) {
          float* z = p.z1 + grow * p.I + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(z + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = row_ok ? act_fwd(p.act, v[i] + s_b1[col0 + i]) : 0.f;
        if (row_ok) {
          float* a = p.a1 + grow * p.I + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(a + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float hi = to_tf32(v[i]); lo[i] = v[i] - hi; v[i] = hi; }
        tmem_st32(t_lane + Cfg::kColC + b * 64 + half * 32, v);
        tmem_st32(t_lane + Cfg::kColL + b * 64 + half * 32, lo);
        tc_fence_before();
        mbar_arrive(a_ready + b);
      }
      if (half == 0) {
        // ---- z2 -> out = LN(drop(z2 + b2) + h), h re-read exactly from its (hi, lo) halves in TMEM ----
        mbar_wait(s3_full, it & 1);
        tc_fence_after();
        float x[64];
        tmem_ld32(t_lane + Cfg::kColAcc, x);
        tmem_ld32(t_lane + Cfg::kColAcc + 32, x + 32);
        if (row_ok) {
          float* z = p.z2 + grow * kDfD;
#pragma unroll
          for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(z + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
        }
        float s = 0.f;
#pragma unroll
        for (int hcol = 0; hcol < 2; ++hcol) {
          float hi[32], lo[32];
          tmem_ld32(t_lane + Cfg::kColA + hcol * 32, hi);
          tmem_ld32(t_lane + Cfg::kColA + 64 + hcol * 32, lo);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float m4[4] = {1.f, 1.f, 1.f, 1.f};
            if (row_ok) df_drop(p, p.mask_f, p.stream_f, grow, hcol * 8 + q, m4);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int i = q * 4 + v, c = hcol * 32 + i;
              x[c] = (x[c] + s_b2[c]) * m4[v] + (hi[i] + lo[i]);
              s += x[c];
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tile_done);                 // acc and h are in registers: the next tile may use their TMEM columns
        const float mean = s * (1.0f / 64);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) { const float t = x[c] - mean; var = fmaf(t, t, var); }
        const float rstd = 1.0f / sqrtf(var * (1.0f / 64) + p.epsF);
        if (row_ok) {
          float* o = p.out + grow * kDfD;
#pragma unroll
          for (int c = 0; c < 64; c += 4)
            *reinterpret_cast<float4*>(o + c) = make_float4((x[c] - mean) * rstd * s_lnFw[c] + s_lnFb[c], (x[c + 1] - mean) * rstd * s_lnFw[c + 1] + s_lnFb[c + 1],
                                                            (x[c + 2] - mean) * rstd * s_lnFw[c + 2] + s_lnFb[c + 2], (x[c + 3] - mean) * rstd * s_lnFw[c + 3] + s_lnFb[c + 3]);
          p.st_f[2 * grow] = mean; p.st_f[2 * grow + 1] = rstd;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace acsr

using namespace acsr;

extern "C" {

int64_t acsr_dense_prep_floats(int I) { return (I > 0 && I % 64 == 0) ? (int64_t)(1 + 2 * (I / 64)) * 2 * kDfOpFloats : 0; }

int acsr_dense_prep(const float* Wo, const float* W1, const float* W2, int d, int I, float* ops, void* stream) {
  if (d != kDfD || I <= 0 || I > kDfMaxI || (I & 63)) {
    set_error("dense_prep: hidden size %d / inner size %d unsupported (64; multiples of 64 up to %d)", d, I, kDfMaxI);
    return ACSR_ERR_UNSUPPORTED;
  }
  ACSR_REQUIRE(Wo && W1 && W2 && ops, "dense_prep: NULL pointer");
  launch_pdl(dense_prep_kernel, dim3(1 + 2 * (I / 64)), dim3(256), 0, (cudaStream_t)stream, Wo, W1, W2, I, ops);
  return check_launch("dense_prep");
}

int acsr_dense_fwd(const float* ctx, const float* res, int64_t rows, int64_t res_rows, int d, int I, int act, const float* ops,
                   const float* bo, const float* lnA_w, const float* lnA_b, float epsA, const float* b1, const float* b2,
                   const float* lnF_w, const float* lnF_b, float epsF, float p_drop, const float* mask_a, const float* mask_f,
                   const void* rng, uint32_t stream_a, uint32_t stream_f, float* hz, float* st_a, float* h, float* z1, float* a1,
                   float* z2, float* st_f, float* out, int passes, void* stream) {
  if (d != kDfD || I <= 0 || I > kDfMaxI || (I & 63)) {
    set_error("dense_fwd: hidden size %d / inner size %d unsupported (64; multiples of 64 up to %d)", d, I, kDfMaxI);
    return ACSR_ERR_UNSUPPORTED;
  }
  ACSR_REQUIRE(ctx && res && ops && bo && lnA_w && lnA_b && b1 && b2 && lnF_w && lnF_b, "dense_fwd: NULL input");
  ACSR_REQUIRE(hz && st_a && h && z1 && a1 && z2 && st_f && out, "dense_fwd: NULL output");
  ACSR_REQUIRE(rows >= 0 && res_rows > 0 && act >= 0 && act <= 4, "dense_fwd: bad sizes");
  ACSR_REQUIRE(passes == 1 || passes == 3, "dense_fwd: passes must be 1 or 3");
  ACSR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "dense_fwd: dropout p=%f", p_drop);
  ACSR_REQUIRE(!(p_drop > 0.f && rng == nullptr && (mask_a == nullptr || mask_f == nullptr)), "dense_fwd: p>0 needs masks or rng");
  if (rows == 0) return ACSR_OK;
  DfParams p = {};
  p.ctx = ctx; p.res = res; p.rows = rows; p.res_rows = res_rows; p.I = I; p.act = act; p.passes = passes; p.ops = ops;
  p.bo = bo; p.lnA_w = lnA_w; p.lnA_b = lnA_b; p.b1 = b1; p.b2 = b2; p.lnF_w = lnF_w; p.lnF_b = lnF_b;
  p.epsA = epsA; p.epsF = epsF; p.p_drop = p_drop; p.mask_a = mask_a; p.mask_f = mask_f; p.rng = (const RngState*)rng;
  p.stream_a = stream_a; p.stream_f = stream_f;
  p.hz = hz; p.st_a = st_a; p.h = h; p.z1 = z1; p.a1 = a1; p.z2 = z2; p.st_f = st_f; p.out = out;
  p.m_tiles = (int)((rows + kDfBM - 1) / kDfBM);
  static_assert(DfCfg::kSmemBytes <= 227 * 1024, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(dense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DfCfg::kSmemBytes);
  if (e != cudaSuccess) { set_error("dense_fwd: smem attr: %s", cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
  launch_pdl(dense_fwd_kernel, dim3(grid), dim3(kDfThreads), DfCfg::kSmemBytes, (cudaStream_t)stream, p);
  return check_launch("dense_fwd");
}

}  // extern "C"
