// Cross-entropy backward over the full catalogue WITHOUT the [M,V] gradient matrix (acsasrec.py:117-121, the backward of
// nn.CrossEntropyLoss(seq_output . E^T) as autograd runs it).  G = (softmax - onehot) * row_scale is recomputed tile by tile
// from the logits and never leaves the SM: the TMEM epilogue turns a logits tile into G (split into hi / lo TF32 halves) and
// writes it BACK to tensor memory with tcgen05.st, where the next tcgen05.mma reads it as its A operand (the "TS" form: A from
// TMEM, B from shared memory), so both contractions of the backward run on the tensor cores:
//   acsr_ce_bwd_dout   : d_out[m,:] += sum_v G[m,v] . E[v,:]      rows of `out` on the UMMA M axis (TMEM lanes), K = table rows
//   acsr_ce_bwd_dtable : d_E[v,:]   += sum_m G[m,v] . out[m,:]    table rows on the UMMA M axis, K = rows of `out`
// Pipeline of both (448 threads, one CTA per SM): warp 0 bulk-copy producer (table tiles -> shared-memory ring), warps 2-5
// splitters (fp32 -> hi/lo TF32 operands in the canonical K-major UMMA layout, plus the TRANSPOSED operand the second MMA
// contracts over), warp 1 issues MMA1 (logits, 3xTF32) and MMA2 (G . operand, 3xTF32), warps 6-13 are the epilogue (two per TMEM
// lane quarter).  Results leave the SM as 16-byte vector reductions (red.global.add.v4.f32).  Per step this removes the write +
// two reads of Gt [V, 2B] (24.8 MB at 12k items, 2 GB at 1M items) and one launch from the critical path.
// (A first version applied the rank-1 updates with fp32 FMAs from the thread owning the logits row: correct, but bound by the
// shared-memory broadcast loads -- 2.7 ms for d_out at 1M items against 0.9 ms for the Gt path; profiles/r02_ce_backward.txt.)
#include "logits_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

constexpr int kBwThreads = 480;        // warp 0 producer, warp 1 MMA1 issuer, warps 2-5 splitter, warps 6-13 epilogue, warp 14 MMA2 issuer
constexpr int kBwIssuer2 = 14;         // (one thread issues a 128 x N x 8 TF32 MMA every ~90 cycles whatever N <= 128 is -- scripts/umma_rate.py;
                                       // two issuing threads double the rate, so the logits MMAs and the G-consuming MMAs get a warp each)
constexpr int kBwEpiThreads = 256;
constexpr int kBwTmemCols = 512;

struct CeBwdParams {
  const float* out;        // [M,64]
  const float* table;      // [V,64]
  const float* lse;        // [M]
  const long long* target; // [M] column of the one-hot (outside [0,V): none)
  const float* row_scale;  // [M]
  int M;
  long long V;
  int passes;
  int m_tiles, n_tiles, n_chunks;
  float* dst;              // d_out [M,64] or d_table [V,64], accumulated
};

__device__ __forceinline__ float4 rot4(float4 x, int rot) {   // x'[j] = x[(j + rot) & 3]
  if (rot & 1) x = make_float4(x.y, x.z, x.w, x.x);
  if (rot & 2) x = make_float4(x.z, x.w, x.x, x.y);
  return x;
}

// Transposed K-major operand: source rows r (4 consecutive ones, x0..x3 = their 16-byte chunk kc) become the K axis.  Element
// (n = source column, k = source row r) lives at float offset (r/4) * (64*4) + n*4 + r%4 (8x16B core matrices, 64 rows per chunk
// plane).  The four chunks a thread writes are rotated by kc so that a quarter-warp's STS.128 hit 8 different bank groups.
__device__ __forceinline__ void split_store_transposed(float* Thi, float* Tlo, int rb, int kc, float4 x0, float4 x1, float4 x2, float4 x3) {
  const int rot = (kc >> 1) & 3;
  x0 = rot4(x0, rot); x1 = rot4(x1, rot); x2 = rot4(x2, rot); x3 = rot4(x3, rot);
  const float a0[4] = {x0.x, x0.y, x0.z, x0.w}, a1[4] = {x1.x, x1.y, x1.z, x1.w};
  const float a2[4] = {x2.x, x2.y, x2.z, x2.w}, a3[4] = {x3.x, x3.y, x3.z, x3.w};
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int j = (jj + rot) & 3;
    const float4 c = make_float4(a0[jj], a1[jj], a2[jj], a3[jj]);
    const float4 hi = make_float4(to_tf32(c.x), to_tf32(c.y), to_tf32(c.z), to_tf32(c.w));
    const float4 lo = make_float4(c.x - hi.x, c.y - hi.y, c.z - hi.z, c.w - hi.w);
    const int off = rb * (kD * 4) + (4 * kc + j) * 4;
    *reinterpret_cast<float4*>(Thi + off) = hi;
    *reinterpret_cast<float4*>(Tlo + off) = lo;
  }
}

// The epilogues are instruction-issue bound (one thread per logits row), so G costs four instructions per element:
//   |G| = softmax * |scale| = ex2(logit * log2e + c),  c = log2|scale| - lse * log2e   (c = -inf when scale == 0)
// then cvt.rna.tf32 and one subtraction for the hi / lo halves.  The SIGN of row_scale is applied to the other MMA operand or to
// the finished accumulator row, and the one-hot term -scale * (row of the other operand) is added separately in exact fp32.
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float g_exponent_offset(float lse, float scale) {
  return scale == 0.f ? -INFINITY : (log2f(fabsf(scale)) - lse * kLog2e);
}
__device__ __forceinline__ void g_split(float x, float c, float& hi, float& lo) {
  const float g = ex2_approx(fmaf(x, kLog2e, c));
  hi = to_tf32(g);
  lo = g - hi;
}

// ------------------------------------------------------------------------------------------------------------------
// d_out: CTA = (128-row tile of `out`, chunk of 64-row table tiles).  The stationary tile of `out` lives in TENSOR MEMORY (hi / lo
// halves written once with tcgen05.st), so MMA1 is a TS-form MMA too and reads only the 2 KB table operand from shared memory:
// an SS-form 128x64x8 TF32 MMA would read 6 KB per 32-cycle dispatch, more than the 128 B/clk of shared-memory bandwidth.
//   TMEM columns: X[2] 0..127 (logits, overwritten IN PLACE by G hi) | Y[2] 128..255 (G lo) | d_out accumulator 256..319 |
//                 out hi 320..383 | out lo 384..447
// tcgen05 MMAs retire in issue order, so MMA1 of tile t+2 (which overwrites X[b]) cannot pass MMA2 of tile t (which reads it).
// ------------------------------------------------------------------------------------------------------------------
struct DoutCfg {
  static constexpr int kStages = 4;
  static constexpr int kBbytes = kBN * kD * 4;          // 16 KB
  static constexpr int kOffOps = 0;                      // [2] x {E hi, E lo, E^T hi, E^T lo}
  static constexpr int kOpsBytes = 4 * kBbytes;
  static constexpr int kOffStg = kOffOps + 2 * kOpsBytes;
  static constexpr int kOffBar = kOffStg + kStages * kBbytes;
  static constexpr int kNumBars = 2 * kStages + 4 + 4 + 2 + 2 + 2 + 1;
  static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16;
  static constexpr int kColX = 0, kColY = 128, kColD = 256, kColAhi = 320, kColAlo = 384;
};

__global__ void __launch_bounds__(kBwThreads, 1) ce_dout_kernel(const CeBwdParams p) {
  using Cfg = DoutCfg;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x % p.m_tiles;
  const int chunk = blockIdx.x / p.m_tiles;
  const int my_tiles = chunk < p.n_tiles ? (p.n_tiles - chunk + p.n_chunks - 1) / p.n_chunks : 0;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* stg_full = bars;
  uint64_t* stg_empty = stg_full + Cfg::kStages;
  uint64_t* e_full = stg_empty + Cfg::kStages;     // E   (K = hidden)     : splitter -> MMA1
  uint64_t* e_empty = e_full + 2;
  uint64_t* t_full = e_empty + 2;                  // E^T (K = table rows) : splitter -> MMA2
  uint64_t* t_empty = t_full + 2;
  uint64_t* l_full = t_empty + 2;                  // logits in X[b]       : MMA1 -> epilogue
  uint64_t* g_full = l_full + 2;                   // G in X[b] / Y[b]     : epilogue -> MMA2
  uint64_t* g_empty = g_full + 2;                  // MMA2 retired: X[b] / Y[b] may take the next logits (the two issuers are not ordered)
  uint64_t* d_full = g_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(stg_full + s, 1); mbar_init(stg_empty + s, 128); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(e_full + s, 128); mbar_init(e_empty + s, 1);
      mbar_init(t_full + s, 128); mbar_init(t_empty + s, 1);
      mbar_init(l_full + s, 1); mbar_init(g_full + s, kBwEpiThreads); mbar_init(g_empty + s, 1);
    }
    mbar_init(d_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kBwTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();                 // everything above overlaps the tail of the previous kernel (which wrote out / lse)

  const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
  const int half = (warp - 6) >> 2;             // epilogue warps: 32-column half of every 64-column tile
  const int row = quarter * 32 + lane;
  const int grow = m_tile * kBM + row;
  const bool row_ok = grow < p.M;
  const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 32;
  if (warp >= 6 && warp < kBwIssuer2) {   // the thread's row of `out` (its 32 columns) -> TMEM, hi and lo
    float hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) x = *reinterpret_cast<const float4*>(p.out + (long long)grow * kD + half * 32 + 4 * j);
      hi[4 * j + 0] = to_tf32(x.x); hi[4 * j + 1] = to_tf32(x.y); hi[4 * j + 2] = to_tf32(x.z); hi[4 * j + 3] = to_tf32(x.w);
      lo[4 * j + 0] = x.x - hi[4 * j + 0]; lo[4 * j + 1] = x.y - hi[4 * j + 1]; lo[4 * j + 2] = x.z - hi[4 * j + 2]; lo[4 * j + 3] = x.w - hi[4 * j + 3];
    }
    tmem_st32(t_lane + Cfg::kColAhi, hi);
    tmem_st32(t_lane + Cfg::kColAlo, lo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % Cfg::kStages;
        const uint32_t ph = (it / Cfg::kStages) & 1;
        const long long n0 = (long long)(chunk + it * p.n_chunks) * kBN;
        const long long rows = (p.V - n0) < kBN ? (p.V - n0) : kBN;
        mbar_wait(stg_empty + s, ph ^ 1);
        mbar_arrive_expect_tx(stg_full + s, (uint32_t)(rows * kD * 4));
        bulk_g2s(smem + Cfg::kOffStg + s * Cfg::kBbytes, p.table + n0 * kD, (uint32_t)(rows * kD * 4), stg_full + s);
      }
    }
  } else if (warp == 1 || warp == kBwIssuer2) {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = umma_idesc_tf32(kBM, kBN);          // both MMAs are 128 x 64 x 8
      const uint32_t a_hi = tmem_base + Cfg::kColAhi, a_lo = tmem_base + Cfg::kColAlo;
      constexpr uint32_t kBLbo = kBN * 16, kSbo = 128;
      const int npass = p.passes == 3 ? 3 : 1;
      if (warp == 1) {
        // ---- MMA1: logits of every tile into X[b] ----
        for (int it = 0; it < my_tiles; ++it) {
          const int ob = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(e_full + ob, ph);
          mbar_wait(g_empty + ob, ph ^ 1);        // MMA2 of tile it-2 no longer reads X[b] / Y[b]
          tc_fence_after();
          const uint32_t b_hi = smem_u32(smem + Cfg::kOffOps + ob * Cfg::kOpsBytes);
          const uint32_t b_lo = b_hi + Cfg::kBbytes;
          const uint32_t d_tmem = tmem_base + Cfg::kColX + ob * kBN;
          uint32_t acc = 0;
          for (int ps = 0; ps < npass; ++ps) {
            const uint32_t a_base = (npass == 3 && ps == 0) ? a_lo : a_hi;
            const uint32_t b_base = (npass == 3 && ps == 1) ? b_lo : b_hi;
#pragma unroll
            for (int ks = 0; ks < kD / 8; ++ks) {
              umma_tf32_ts(d_tmem, a_base + ks * 8, umma_desc_kmajor(b_base + ks * 2 * kBLbo, kBLbo, kSbo), idesc, acc);
              acc = 1;
            }
          }
          umma_commit(e_empty + ob);      // the K = hidden layout of this table tile is free once the logits MMAs retired
          umma_commit(l_full + ob);
        }
      } else {
        // ---- MMA2: d_out += G . E of every tile, from its own issuing thread ----
        for (int it = 0; it < my_tiles; ++it) {
          const int ob = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(t_full + ob, ph);
          mbar_wait(g_full + ob, ph);
          tc_fence_after();
          const uint32_t t_hi = smem_u32(smem + Cfg::kOffOps + ob * Cfg::kOpsBytes + 2 * Cfg::kBbytes);   // E^T: N = hidden, K = table rows
          const uint32_t t_lo = t_hi + Cfg::kBbytes;
          const uint32_t g_hi = tmem_base + Cfg::kColX + ob * kBN, g_lo = tmem_base + Cfg::kColY + ob * kBN;
          const uint32_t d_tmem = tmem_base + Cfg::kColD;
          for (int ps = 0; ps < npass; ++ps) {
            const uint32_t a_base = (npass == 3 && ps == 0) ? g_lo : g_hi;
            const uint32_t b_base = (npass == 3 && ps == 1) ? t_lo : t_hi;
#pragma unroll
            for (int ks = 0; ks < kBN / 8; ++ks)
              umma_tf32_ts(d_tmem, a_base + ks * 8, umma_desc_kmajor(b_base + ks * 2 * kBLbo, kBLbo, kSbo), idesc,
                           (it > 0 || ps > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(t_empty + ob);
          umma_commit(g_empty + ob);
        }
        umma_commit(d_full);
      }
    }
  } else if (warp < 6) {
    const int tid = threadIdx.x - 64;   // 0..127
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % Cfg::kStages;
      const uint32_t sph = (it / Cfg::kStages) & 1;
      const int ob = it & 1;
      const uint32_t oph = (it >> 1) & 1;
      const long long n0 = (long long)(chunk + it * p.n_chunks) * kBN;
      const int rows = (int)((p.V - n0) < kBN ? (p.V - n0) : kBN);
      mbar_wait(stg_full + s, sph);
      mbar_wait(e_empty + ob, oph ^ 1);
      const float* stg = reinterpret_cast<const float*>(smem + Cfg::kOffStg + s * Cfg::kBbytes);
      float* Bhi = reinterpret_cast<float*>(smem + Cfg::kOffOps + ob * Cfg::kOpsBytes);
      float* Blo = Bhi + kBN * kD;
      float* Thi = Blo + kBN * kD;
      float* Tlo = Thi + kBN * kD;
#pragma unroll
      for (int q = 0; q < (kBN * kKC) / 128; ++q) {
        const int item = q * 128 + tid;
        const int r = item % kBN;
        const int kc = (item / kBN + r) % kKC;      // diagonal rotation: conflict-free LDS and STS
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) x = *reinterpret_cast<const float4*>(stg + r * kD + kc * 4);
        float4 hi = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
        float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
        const int off = kc * (kBN * 4) + r * 4;
        *reinterpret_cast<float4*>(Bhi + off) = hi;
        *reinterpret_cast<float4*>(Blo + off) = lo;
      }
      fence_proxy_async();               // generic-proxy stores -> visible to the tensor-core (async) proxy
      mbar_arrive(e_full + ob);          // the logits MMAs of this tile may start while the transposed layout is built
      mbar_wait(t_empty + ob, oph ^ 1);
#pragma unroll
      for (int q = 0; q < (kBN / 4 * kKC) / 128; ++q) {     // the same tile with the table rows on the K axis
        const int item = q * 128 + tid;
        const int kc = item & (kKC - 1), rb = item / kKC;
        float4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          x[i] = (4 * rb + i < rows) ? *reinterpret_cast<const float4*>(stg + (4 * rb + i) * kD + kc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        split_store_transposed(Thi, Tlo, rb, kc, x[0], x[1], x[2], x[3]);
      }
      fence_proxy_async();
      mbar_arrive(t_full + ob);
      mbar_arrive(stg_empty + s);
    }
  } else if (warp < kBwIssuer2) {
    // ------------------------------ epilogue: logits -> G (hi in place, lo next to it); at the end the accumulator -> d_out ------------------------------
    float g_c = -INFINITY, g_scale = 0.f;
    long long g_tgt = -1;
    if (row_ok) { g_scale = p.row_scale[grow]; g_c = g_exponent_offset(p.lse[grow], g_scale); g_tgt = p.target[grow]; }
    for (int it = 0; it < my_tiles; ++it) {
      const int ob = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const long long c0 = (long long)(chunk + it * p.n_chunks) * kBN + half * 32;
      mbar_wait(l_full + ob, ph);
      tc_fence_after();
      float v[32], lo[32];
      tmem_ld32(t_lane + Cfg::kColX + ob * kBN, v);
      if (p.V - c0 >= 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { float hi; g_split(v[i], g_c, hi, lo[i]); v[i] = hi; }
      } else {                           // ragged last tile: columns past the catalogue carry no gradient
        const int nvalid = (int)((p.V - c0) > 0 ? (p.V - c0) : 0);
#pragma unroll
        for (int i = 0; i < 32; ++i) { float hi; g_split(v[i], (i < nvalid) ? g_c : -INFINITY, hi, lo[i]); v[i] = hi; }
      }
      tmem_st32(t_lane + Cfg::kColX + ob * kBN, v);
      tmem_st32(t_lane + Cfg::kColY + ob * kBN, lo);
      tc_fence_before();
      mbar_arrive(g_full + ob);
    }
    if (my_tiles > 0) {
      mbar_wait(d_full, 0);
      tc_fence_after();
      float v[32];
      tmem_ld32(t_lane + Cfg::kColD, v);
      if (row_ok) {
        const float sgn = g_scale < 0.f ? -1.f : 1.f;
        float* dst = p.dst + (long long)grow * kD + half * 32;
        // the one-hot term of this row, -scale * E[target], once per row: the CTA that owns the row's first chunk adds it
        const bool onehot = chunk == 0 && g_tgt >= 0 && g_tgt < p.V && g_scale != 0.f;
        const float* et = p.table + (onehot ? g_tgt : 0) * kD + half * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
          if (onehot) e = *reinterpret_cast<const float4*>(et + 4 * j);
          red_add_v4(dst + 4 * j, fmaf(-g_scale, e.x, sgn * v[4 * j]), fmaf(-g_scale, e.y, sgn * v[4 * j + 1]),
                     fmaf(-g_scale, e.z, sgn * v[4 * j + 2]), fmaf(-g_scale, e.w, sgn * v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kBwTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// d_table: CTA = (chunk of 128-row table tiles, 128-row block of `out`).  logits^T tile [128 table rows, 128 rows of out] in
// TMEM, thread = table row; G^T (hi, lo) goes back into TMEM and is the A operand of MMA2, whose B operand is the block of `out`
// with its rows on the K axis (stationary, built once per CTA).
//   TMEM columns: logits^T 0..127 | G^T hi 128..255 | G^T lo 256..383 | d_E accumulator [2] 384..511
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTN = 128;               // table rows per tile (UMMA M)
constexpr int kTM = 128;               // rows of `out` per CTA (UMMA N of MMA1, K of MMA2)
struct DtabCfg {
  static constexpr int kObytes = kTM * kD * 4;          // 32 KB per (hi|lo)
  static constexpr int kEbytes = kTN * kD * 4;          // 32 KB per (hi|lo|stage)
  static constexpr int kOffOhi = 0;                      // out block, K = hidden  (B operand of MMA1)
  static constexpr int kOffOlo = kOffOhi + kObytes;
  static constexpr int kOffThi = kOffOlo + kObytes;      // out block, K = its rows (B operand of MMA2)
  static constexpr int kOffTlo = kOffThi + kObytes;
  static constexpr int kOffEhi = kOffTlo + kObytes;      // table tile (A operand of MMA1)
  static constexpr int kOffElo = kOffEhi + kEbytes;
  static constexpr int kOffStg = kOffElo + kEbytes;      // one raw stage
  static constexpr int kOffMeta = kOffStg + kEbytes;     // exponent offset c[m] = log2|scale| - lse * log2e per row of the block
  static constexpr int kOffBar = kOffMeta + kTM * 4;
  static constexpr int kNumBars = 12;
  static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16;
  static constexpr int kColL = 0, kColGhi = 128, kColGlo = 256, kColD = 384;
};

__global__ void __launch_bounds__(kBwThreads, 1) ce_dtable_kernel(const CeBwdParams p) {
  using Cfg = DtabCfg;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_block = blockIdx.x % p.m_tiles;
  const int chunk = blockIdx.x / p.m_tiles;
  const int my_tiles = chunk < p.n_tiles ? (p.n_tiles - chunk + p.n_chunks - 1) / p.n_chunks : 0;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* stg_full = bars;
  uint64_t* stg_empty = bars + 1;
  uint64_t* op_full = bars + 2;
  uint64_t* op_empty = bars + 3;
  uint64_t* l_full = bars + 4;
  uint64_t* l_empty = bars + 5;
  uint64_t* g_full = bars + 6;
  uint64_t* g_empty = bars + 7;
  uint64_t* d_full = bars + 8;     // [2]
  uint64_t* d_empty = bars + 10;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);
  float* meta = reinterpret_cast<float*>(smem + Cfg::kOffMeta);

  if (threadIdx.x == 0) {
    mbar_init(stg_full, 1); mbar_init(stg_empty, 128);
    mbar_init(op_full, 128); mbar_init(op_empty, 1);
    mbar_init(l_full, 1); mbar_init(l_empty, kBwEpiThreads);
    mbar_init(g_full, kBwEpiThreads); mbar_init(g_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(d_full + s, 1); mbar_init(d_empty + s, kBwEpiThreads); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kBwTmemCols>(tmem_slot);
  pdl_launch_dependents();
  pdl_wait();

  {  // stationary block of `out` in both operand layouts + the per-row CE terms
    float* Ohi = reinterpret_cast<float*>(smem + Cfg::kOffOhi);
    float* Olo = reinterpret_cast<float*>(smem + Cfg::kOffOlo);
    float* Thi = reinterpret_cast<float*>(smem + Cfg::kOffThi);
    float* Tlo = reinterpret_cast<float*>(smem + Cfg::kOffTlo);
    for (int item = threadIdx.x; item < kTM * kKC; item += kBwThreads) {
      const int r = item / kKC, kc = item % kKC;
      const int grow = m_block * kTM + r;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < p.M) x = *reinterpret_cast<const float4*>(p.out + (long long)grow * kD + kc * 4);
      float4 hi = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
      float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
      const int off = kc * (kTM * 4) + r * 4;
      *reinterpret_cast<float4*>(Ohi + off) = hi;
      *reinterpret_cast<float4*>(Olo + off) = lo;
    }
    for (int item = threadIdx.x; item < (kTM / 4) * kKC; item += kBwThreads) {
      const int kc = item & (kKC - 1), rb = item / kKC;
      float4 x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int grow = m_block * kTM + 4 * rb + i;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < p.M) {
          x[i] = *reinterpret_cast<const float4*>(p.out + (long long)grow * kD + kc * 4);
          if (p.row_scale[grow] < 0.f) x[i] = make_float4(-x[i].x, -x[i].y, -x[i].z, -x[i].w);   // |G| . (sign * out)
        }
      }
      split_store_transposed(Thi, Tlo, rb, kc, x[0], x[1], x[2], x[3]);
    }
    for (int r = threadIdx.x; r < kTM; r += kBwThreads) {
      const int grow = m_block * kTM + r;
      meta[r] = grow < p.M ? g_exponent_offset(p.lse[grow], p.row_scale[grow]) : -INFINITY;
    }
    if (chunk == 0) {
      // one-hot terms of this block's rows, d_E[target[m], :] -= scale[m] * out[m, :], in exact fp32 (one CTA per block adds them)
      for (int item = threadIdx.x; item < kTM * kKC; item += kBwThreads) {
        const int r = item / kKC, kc = item % kKC;
        const int grow = m_block * kTM + r;
        if (grow >= p.M) continue;
        const long long t = p.target[grow];
        const float sc = p.row_scale[grow];
        if (t < 0 || t >= p.V || sc == 0.f) continue;
        const float4 x = *reinterpret_cast<const float4*>(p.out + (long long)grow * kD + kc * 4);
        red_add_v4(p.dst + t * kD + kc * 4, -sc * x.x, -sc * x.y, -sc * x.z, -sc * x.w);
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < my_tiles; ++it) {
        const long long n0 = (long long)(chunk + it * p.n_chunks) * kTN;
        const long long rows = (p.V - n0) < kTN ? (p.V - n0) : kTN;
        mbar_wait(stg_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(stg_full, (uint32_t)(rows * kD * 4));
        bulk_g2s(smem + Cfg::kOffStg, p.table + n0 * kD, (uint32_t)(rows * kD * 4), stg_full);
      }
    }
  } else if (warp == 1 || warp == kBwIssuer2) {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc1 = umma_idesc_tf32(kTN, kTM);      // logits^T: 128 table rows x 128 rows of out, K = hidden
      const uint32_t idesc2 = umma_idesc_tf32(kTN, kD);       // d_E tile : 128 table rows x 64, K = rows of out
      const uint32_t e_hi = smem_u32(smem + Cfg::kOffEhi), e_lo = smem_u32(smem + Cfg::kOffElo);
      const uint32_t o_hi = smem_u32(smem + Cfg::kOffOhi), o_lo = smem_u32(smem + Cfg::kOffOlo);
      const uint32_t t_hi = smem_u32(smem + Cfg::kOffThi), t_lo = smem_u32(smem + Cfg::kOffTlo);
      constexpr uint32_t kELbo = kTN * 16, kOLbo = kTM * 16, kTLbo = kD * 16, kSbo = 128;
      const int npass = p.passes == 3 ? 3 : 1;
      auto issue_logits = [&](int it) {
        mbar_wait(op_full, it & 1);
        mbar_wait(l_empty, (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + Cfg::kColL;
        uint32_t acc = 0;
        for (int ps = 0; ps < npass; ++ps) {
          const uint32_t a_base = (npass == 3 && ps == 0) ? e_lo : e_hi;
          const uint32_t b_base = (npass == 3 && ps == 1) ? o_lo : o_hi;
#pragma unroll
          for (int ks = 0; ks < kD / 8; ++ks) {
            umma_tf32(d_tmem, umma_desc_kmajor(a_base + ks * 2 * kELbo, kELbo, kSbo), umma_desc_kmajor(b_base + ks * 2 * kOLbo, kOLbo, kSbo),
                      idesc1, acc);
            acc = 1;
          }
        }
        umma_commit(op_empty);          // the table tile's operand buffer is free once the logits MMAs retired
        umma_commit(l_full);
      };
      auto issue_dtable = [&](int it) {
        const int db = it & 1;
        mbar_wait(g_full, it & 1);
        mbar_wait(d_empty + db, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + Cfg::kColD + db * kD;
        uint32_t acc = 0;
        for (int ps = 0; ps < npass; ++ps) {
          const uint32_t a_base = tmem_base + ((npass == 3 && ps == 0) ? Cfg::kColGlo : Cfg::kColGhi);
          const uint32_t b_base = (npass == 3 && ps == 1) ? t_lo : t_hi;
#pragma unroll
          for (int ks = 0; ks < kTM / 8; ++ks) {
            umma_tf32_ts(d_tmem, a_base + ks * 8, umma_desc_kmajor(b_base + ks * 2 * kTLbo, kTLbo, kSbo), idesc2, acc);
            acc = 1;
          }
        }
        umma_commit(g_empty);
        umma_commit(d_full + db);
      };
      if (warp == 1) {
        for (int it = 0; it < my_tiles; ++it) issue_logits(it);
      } else {
        for (int it = 0; it < my_tiles; ++it) issue_dtable(it);
      }
    }
  } else if (warp < 6) {
    const int tid = threadIdx.x - 64;   // 0..127
    for (int it = 0; it < my_tiles; ++it) {
      const long long n0 = (long long)(chunk + it * p.n_chunks) * kTN;
      const int rows = (int)((p.V - n0) < kTN ? (p.V - n0) : kTN);
      mbar_wait(stg_full, it & 1);
      mbar_wait(op_empty, (it & 1) ^ 1);
      const float* stg = reinterpret_cast<const float*>(smem + Cfg::kOffStg);
      float* Ehi = reinterpret_cast<float*>(smem + Cfg::kOffEhi);
      float* Elo = reinterpret_cast<float*>(smem + Cfg::kOffElo);
#pragma unroll
      for (int q = 0; q < (kTN * kKC) / 128; ++q) {
        const int item = q * 128 + tid;
        const int r = item % kTN;
        const int kc = (item / kTN + r) % kKC;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) x = *reinterpret_cast<const float4*>(stg + r * kD + kc * 4);
        float4 hi = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
        float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
        const int off = kc * (kTN * 4) + r * 4;
        *reinterpret_cast<float4*>(Ehi + off) = hi;
        *reinterpret_cast<float4*>(Elo + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(op_full);
      mbar_arrive(stg_empty);
    }
  } else if (warp < kBwIssuer2) {
    // ------------------------------ epilogue: thread = table row; 64 of the block's 128 rows of `out` per warp ------------------------------
    const int quarter = warp & 3;
    const int half = (warp - 6) >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    auto flush = [&](int it) {           // d_E accumulator of tile `it` -> global (32 of its 64 columns per warp)
      const int db = it & 1;
      const long long vrow = (long long)(chunk + it * p.n_chunks) * kTN + quarter * 32 + lane;
      mbar_wait(d_full + db, (it >> 1) & 1);
      tc_fence_after();
      float v[32];
      tmem_ld32(t_lane + Cfg::kColD + db * kD + half * 32, v);
      tc_fence_before();
      mbar_arrive(d_empty + db);
      if (vrow < p.V) {
        float* dst = p.dst + vrow * kD + half * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) red_add_v4(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      mbar_wait(l_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int m0 = half * 64 + c * 32;
        float v[32], lo[32];
        tmem_ld32(t_lane + Cfg::kColL + m0, v);
        if (c == 1) { tc_fence_before(); mbar_arrive(l_empty); }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 cm = *reinterpret_cast<const float4*>(meta + m0 + i);      // one broadcast load per four columns
          float hi;
          g_split(v[i + 0], cm.x, hi, lo[i + 0]); v[i + 0] = hi;
          g_split(v[i + 1], cm.y, hi, lo[i + 1]); v[i + 1] = hi;
          g_split(v[i + 2], cm.z, hi, lo[i + 2]); v[i + 2] = hi;
          g_split(v[i + 3], cm.w, hi, lo[i + 3]); v[i + 3] = hi;
        }
        if (c == 0) { mbar_wait(g_empty, (it & 1) ^ 1); tc_fence_after(); }    // MMA2 of the previous tile has retired
        tmem_st32(t_lane + Cfg::kColGhi + m0, v);
        tmem_st32(t_lane + Cfg::kColGlo + m0, lo);
      }
      tc_fence_before();
      mbar_arrive(g_full);
      if (it > 0) flush(it - 1);         // the previous tile's accumulator drains while MMA2 of this one runs
    }
    if (my_tiles > 0) flush(my_tiles - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kBwTmemCols>(tmem_base);
  }
}

static int ce_bwd_validate(const float* out, const float* table, const float* lse, const int64_t* target, const float* row_scale, int M,
                           long long V, int d, int passes, float* dst, const char* who) {
  ACSR_REQUIRE(out && table && lse && target && row_scale && dst, "%s: NULL pointer", who);
  ACSR_REQUIRE(M > 0 && V > 0 && V < (1ll << 31), "%s: bad sizes M=%d V=%lld", who, M, V);
  if (d != kD) { set_error("%s: hidden size %d unsupported (64; other widths keep acsr_logits_ce_grad + acsr_gemm_batch)", who, d); return ACSR_ERR_UNSUPPORTED; }
  ACSR_REQUIRE(passes == 1 || passes == 3, "%s: passes must be 1 (TF32) or 3 (3xTF32)", who);
  ACSR_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(table) & 15) == 0, "%s: pointers must be 16-byte aligned", who);
  return ACSR_OK;
}

template <typename Cfg, typename K>
static int ce_bwd_launch(K kernel, CeBwdParams& p, int rows_per_tile, int rows_per_block, cudaStream_t st, const char* who) {
  p.m_tiles = (p.M + rows_per_block - 1) / rows_per_block;
  p.n_tiles = (int)((p.V + rows_per_tile - 1) / rows_per_tile);
  int nc = kNumSMs / p.m_tiles;
  if (nc < 1) nc = 1;
  if (nc > p.n_tiles) nc = p.n_tiles;
  p.n_chunks = nc;
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (e != cudaSuccess) { set_error("%s: smem attr: %s", who, cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  launch_pdl(kernel, dim3(p.m_tiles * p.n_chunks), dim3(kBwThreads), Cfg::kSmemBytes, st, p);
  return check_launch(who);
}

}  // namespace acsr

using namespace acsr;

extern "C" {

int acsr_ce_bwd_dout(const float* out, const float* table, const float* lse, const int64_t* target, const float* row_scale, int M,
                     int64_t V, int d, int passes, float* d_out, void* stream) {
  int rc = ce_bwd_validate(out, table, lse, target, row_scale, M, V, d, passes, d_out, "ce_bwd_dout");
  if (rc) return rc;
  CeBwdParams p = {};
  p.out = out; p.table = table; p.lse = lse; p.target = (const long long*)target; p.row_scale = row_scale;
  p.M = M; p.V = V; p.passes = passes; p.dst = d_out;
  return ce_bwd_launch<DoutCfg>(ce_dout_kernel, p, kBN, kBM, (cudaStream_t)stream, "ce_bwd_dout");
}

int acsr_ce_bwd_dtable(const float* out, const float* table, const float* lse, const int64_t* target, const float* row_scale, int M,
                       int64_t V, int d, int passes, float* d_table, void* stream) {
  int rc = ce_bwd_validate(out, table, lse, target, row_scale, M, V, d, passes, d_table, "ce_bwd_dtable");
  if (rc) return rc;
  CeBwdParams p = {};
  p.out = out; p.table = table; p.lse = lse; p.target = (const long long*)target; p.row_scale = row_scale;
  p.M = M; p.V = V; p.passes = passes; p.dst = d_table;
  return ce_bwd_launch<DtabCfg>(ce_dtable_kernel, p, kTN, kTM, (cudaStream_t)stream, "ce_bwd_dtable");
}

}  // extern "C"
