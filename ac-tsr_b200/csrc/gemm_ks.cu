// General encoder GEMM  C[M,N] (+)= A[M,K] . B[N,K]^T  on the 5th-gen tensor cores (tcgen05), 3xTF32,
// for ANY hidden size: both operands are streamed along K in 32-wide slabs through a shared-memory ring.
//
// Replaces the nn.Linear forward / input-gradient / weight-gradient GEMMs of layers.py:658-659, 680,
// 687-689, 791-794, 887 wherever the width-64 kernel (linear_tok.cu: weights stationary, K <= 256) does
// not apply: d = 128 (config/yelp.yaml:39-42 and the two other shipped d=128 YAMLs), d = 256 / I = 1024
// (BASELINE config #5), and -- at every width -- the weight gradients dW = dY^T.X, whose contraction runs
// over the token axis (12,800 ... 409,600 tokens) and is split over the CTAs (split-K + vector atomics
// straight into the flat gradient buffer).
//
// One launch carries a LIST of problems (up to 16: all weight gradients of a layer, or the projections
// that read the same input); the persistent CTAs walk the work items (problem, m-tile, n-block, k-split)
// round robin.
//
// Roles (416 threads, one CTA per SM):
//   warp 0        tcgen05.mma issuer (one lane) + TMEM allocation (2 accumulator stages x 256 columns)
//   warps 1-4     A loaders, warps 5-8 B loaders: fp32 global -> (hi, lo) TF32 operands (hi = rna(x),
//                 lo = x - hi) written in the canonical no-swizzle K-major UMMA layout.  Operands are
//                 addressed with (row, k, k-block) strides; two lane mappings keep both the global loads
//                 and the 16-byte shared stores efficient: k-contiguous operands (8 rows x 4 chunks per
//                 warp instruction: 64-byte global segments, conflict-free stores) and row-contiguous
//                 operands (transposed use: lane = row, four 128-byte coalesced loads per 16-byte chunk).
//   warps 9-12    epilogue: tcgen05.ld 32x32b -> ONE THREAD OWNS ONE OUTPUT ROW: bias / activation /
//                 dropout / residual / LayerNorm are thread-private; for the LayerNorm epilogue the
//                 pre-norm row is parked back in TMEM (tcgen05.st) between the passes.
#include "acsr_common.cuh"
#include "../../include/acsr.h"

namespace acsr {

constexpr int kGkThreads = 416;
constexpr int kGkBM = 128;                       // output rows per tile (UMMA M)
constexpr int kGkKS = 32;                        // K slab (floats)
constexpr int kGkChunks = kGkKS / 4;             // 16-byte chunks per row per slab
constexpr int kGkAOp = kGkBM * kGkKS * 4;        // bytes of one (hi | lo) A operand slab
constexpr int kGkMaxStages = 4;
constexpr int kGkAccCols = 256;                  // TMEM columns per accumulator stage
constexpr int kGkLoaderThreads = 128;            // per operand

struct GkProblem {
  acsr_gemm_problem p;
  int m_tiles, n_blocks, BN, ksplit, slabs, slabs_per_split, item_begin, a_vec, b_vec;
};

struct GkParams {
  int n_problems, total_items, stages, stage_bytes, passes;
  int b_static;          // every B operand is registered parameter memory: the B loaders need not wait for the previous kernel
  GkProblem pr[ACSR_GEMM_MAX_PROBLEMS];
};

struct GkItem {
  int q, m_tile, n_block, s0, s1;
};

__device__ __forceinline__ GkItem gk_decode(const GkParams& P, int item) {
  int q = 0;
#pragma unroll 1
  for (int i = 1; i < P.n_problems; ++i)
    if (item >= P.pr[i].item_begin) q = i;
  const GkProblem& g = P.pr[q];
  int local = item - g.item_begin;
  GkItem it;
  it.q = q;
  const int ks = local % g.ksplit;
  local /= g.ksplit;
  it.m_tile = local % g.m_tiles;
  it.n_block = local / g.m_tiles;
  it.s0 = ks * g.slabs_per_split;
  it.s1 = min(g.slabs, it.s0 + g.slabs_per_split);
  return it;
}

struct GkOp {
  const float* base;
  long long s_r, s_k, kbs;
  int kblk;
};

// 4 consecutive k of one row (zero beyond K / invalid rows)
__device__ __forceinline__ float4 gk_load4(const GkOp& v, long long row, bool row_ok, int k, int K, bool vec_ok) {
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok || k >= K) return x;
  const int kb = k >= v.kblk ? k / v.kblk : 0;
  const float* p = v.base + (long long)kb * v.kbs + row * v.s_r + (long long)(k - kb * v.kblk) * v.s_k;
  if (vec_ok && k + 4 <= K) return __ldg(reinterpret_cast<const float4*>(p));
  x.x = __ldg(p);
  if (k + 1 < K) x.y = __ldg(p + v.s_k);
  if (k + 2 < K) x.z = __ldg(p + 2 * v.s_k);
  if (k + 3 < K) x.w = __ldg(p + 3 * v.s_k);
  return x;
}

__device__ __forceinline__ void gk_split_store(uint8_t* hi, uint8_t* lo, int off, const float4 x) {
  const float4 h = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
  const float4 l = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

// One operand slab: R rows (multiple of 16) x 32 k -> (hi, lo) in the canonical layout (chunk plane of R rows x 16 B).
// 128 threads.  `wait` (the ring slot is free) is executed after the first batch of loads is in flight.
//
// Generic path: any strides / alignment / K tail; unit u = (row, chunk), index arithmetic per unit.
template <typename WaitFn>
__device__ __forceinline__ float gk_load_slab_generic(const GkOp& v, long long row0, long long rows_total, int R, int k0, int K, bool kc_map,
                                                      bool vec_ok, int tid, uint8_t* hi, uint8_t* lo, WaitFn wait) {
  const int upt = R / 16;                          // units per thread (R * 8 / 128)
  float rowsum = 0.f;
#pragma unroll 1
  for (int g0 = 0; g0 < upt; g0 += 8) {
    float4 x[8];
    int off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      off[j] = -1;
      if (g0 + j < upt) {
        const int u = (g0 + j) * kGkLoaderThreads + tid;
        int row, chunk;
        if (kc_map) {
          const int w = u >> 5, l = u & 31;
          row = (w >> 1) * 8 + (l & 7);
          chunk = (w & 1) * 4 + (l >> 3);
        } else {
          row = u % R;
          chunk = u / R;
        }
        const long long grow = row0 + row;
        x[j] = gk_load4(v, grow, grow < rows_total, k0 + chunk * 4, K, vec_ok);
        off[j] = chunk * (R * 16) + row * 16;
      }
    }
    if (g0 == 0) wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (off[j] >= 0) {
        rowsum += (x[j].x + x[j].y) + (x[j].z + x[j].w);
        gk_split_store(hi, lo, off[j], x[j]);
      }
    }
  }
  return rowsum;
}

// Fast path, k-contiguous operand (forward GEMMs, input gradients' activations): 16-byte aligned rows, the slab lies inside one
// k-block and inside K.  A warp instruction covers 8 rows x 4 chunks (64-byte global segments; the 16-byte shared stores of a
// quarter warp fall into 8 different bank groups); unit j of a thread is 16 rows further down: one pointer, one stride.
template <typename WaitFn>
__device__ __forceinline__ void gk_load_slab_kc(const float* p0, long long stride16, long long rows_left, int upt, int off0, uint8_t* hi,
                                                uint8_t* lo, WaitFn wait) {
  // p0: this thread's (row_in, chunk) element of the slab; rows_left = rows_total - (row0 + row_in)
#pragma unroll 1
  for (int g0 = 0; g0 < upt; g0 += 8) {
    float4 x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g0 + j < upt && (long long)(g0 + j) * 16 < rows_left) x[j] = __ldg(reinterpret_cast<const float4*>(p0 + (g0 + j) * stride16));
    }
    if (g0 == 0) wait();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (g0 + j < upt) gk_split_store(hi, lo, off0 + (g0 + j) * 256, x[j]);
  }
}

// Fast path, row-contiguous mapping (transposed operands: the weight of an input gradient, both operands of a weight gradient):
// lane = row, so each of the four scalar loads of a 16-byte chunk is one coalesced 128-byte request; a thread owns row `tid`
// (and tid + 128) for all 8 chunks of the slab.
template <typename WaitFn>
__device__ __forceinline__ float gk_load_slab_rc(const float* pk, long long s_r, long long s_k, long long row0, long long rows_total, int R,
                                                 int kleft, int tid, uint8_t* hi, uint8_t* lo, WaitFn wait) {
  // pk: element (row 0, k0) of the operand; kleft = K - k0 (> 0)
  float rowsum = 0.f;
  bool waited = false;
#pragma unroll 1
  for (int rr = tid; rr < R; rr += kGkLoaderThreads) {
    const long long grow = row0 + rr;
    float x[32];
    if (grow < rows_total) {
      const float* p = pk + grow * s_r;
      if (kleft >= 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __ldg(p + i * s_k);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = i < kleft ? __ldg(p + i * s_k) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = 0.f;
    }
    if (!waited) { wait(); waited = true; }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v4 = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
      rowsum += (v4.x + v4.y) + (v4.z + v4.w);
      gk_split_store(hi, lo, c * (R * 16) + rr * 16, v4);
    }
  }
  if (!waited) wait();
  return rowsum;
}

__global__ void __launch_bounds__(kGkThreads, 1) gemm_ks_kernel(const __grid_constant__ GkParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.stages * P.stage_bytes);
  uint64_t* full = bars;                           // [stages] loaders -> MMA
  uint64_t* empty = bars + kGkMaxStages;           // [stages] MMA -> loaders
  uint64_t* tm_full = bars + 2 * kGkMaxStages;     // [2] MMA -> epilogue
  uint64_t* tm_empty = tm_full + 2;                // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGkMaxStages; ++s) { mbar_init(full + s, 2 * kGkLoaderThreads); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tm_full + s, 1); mbar_init(tm_empty + s, 128); }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<2 * kGkAccCols>(tmem_slot);
  // Everything above overlaps the tail of the previous kernel; when all B operands are registered parameters (weights) the B
  // loader warps start streaming them at once -- they touch nothing the previous kernel wrote.
  if (!(P.b_static && warp >= 5 && warp <= 8)) pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int S = P.stages;

  if (warp == 0) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const int npass = P.passes == 3 ? 3 : 1;
      int cnt = 0, it = 0;
      for (int item = blockIdx.x; item < P.total_items; item += gridDim.x, ++it) {
        const GkItem w = gk_decode(P, item);
        const int BN = P.pr[w.q].BN;
        const uint32_t idesc = umma_idesc_tf32(kGkBM, BN);
        const int acc_stage = it & 1;
        mbar_wait(tm_empty + acc_stage, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_stage * kGkAccCols;
        constexpr uint32_t kALbo = kGkBM * 16, kSbo = 128;
        const uint32_t kBLbo = BN * 16;
        uint32_t acc = 0;
        for (int s = w.s0; s < w.s1; ++s, ++cnt) {
          const int stage = cnt % S;
          mbar_wait(full + stage, (cnt / S) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * P.stage_bytes), a_lo = a_hi + kGkAOp;
          const uint32_t b_hi = a_hi + 2 * kGkAOp, b_lo = b_hi + BN * (kGkKS * 4);
          for (int ps = 0; ps < npass; ++ps) {
            // small cross terms first, the dominant hi.hi product last
            const uint32_t a_base = (npass == 3 && ps == 0) ? a_lo : a_hi;
            const uint32_t b_base = (npass == 3 && ps == 1) ? b_lo : b_hi;
#pragma unroll
            for (int ks = 0; ks < kGkKS / 8; ++ks) {
              const uint64_t ad = umma_desc_kmajor(a_base + ks * 2 * kALbo, kALbo, kSbo);
              const uint64_t bd = umma_desc_kmajor(b_base + ks * 2 * kBLbo, kBLbo, kSbo);
              umma_tf32(d_tmem, ad, bd, idesc, acc);
              acc = 1;
            }
          }
          umma_commit(empty + stage);       // the slab may be overwritten once these MMAs retire
        }
        umma_commit(tm_full + acc_stage);   // accumulator stage ready for the epilogue
      }
    }
  } else if (warp <= 8) {
    // ------------------------------ loaders / hi-lo splitters ------------------------------
    const bool is_a = warp <= 4;
    const int tid = threadIdx.x - (is_a ? 32 : 32 + kGkLoaderThreads);
    int cnt = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      const GkItem w = gk_decode(P, item);
      const GkProblem& g = P.pr[w.q];
      const acsr_gemm_problem& p = g.p;
      GkOp v;
      long long row0, rows_total;
      int R, vec;
      if (is_a) {
        v.base = p.A; v.s_r = p.a_row_stride; v.s_k = p.a_k_stride; v.kbs = p.a_kb_stride; v.kblk = p.a_kblk;
        row0 = (long long)w.m_tile * kGkBM; rows_total = p.M; R = kGkBM; vec = g.a_vec;
      } else {
        v.base = p.B; v.s_r = p.b_row_stride; v.s_k = p.b_k_stride; v.kbs = p.b_kb_stride; v.kblk = p.b_kblk;
        row0 = (long long)w.n_block * g.BN; rows_total = p.N; R = g.BN; vec = g.b_vec;
      }
      const bool kc_map = v.s_k == 1;
      const int lo_off = is_a ? kGkAOp : g.BN * (kGkKS * 4);
      // fast paths need slabs that never straddle a k-block
      const bool blk_ok = v.kblk >= p.K || (v.kblk % kGkKS) == 0;
      const int wid = tid >> 5, ln = tid & 31;
      const int kc_chunk = (wid & 1) * 4 + (ln >> 3), kc_row = (wid >> 1) * 8 + (ln & 7);
      float rowsum = 0.f;
      for (int s = w.s0; s < w.s1; ++s, ++cnt) {
        const int stage = cnt % S;
        uint8_t* hi = smem + stage * P.stage_bytes + (is_a ? 0 : 2 * kGkAOp);
        uint64_t* eb = empty + stage;
        const uint32_t par = ((cnt / S) & 1) ^ 1;
        auto wait = [&]() { mbar_wait(eb, par); };
        const int k0 = s * kGkKS;
        const int kb = k0 >= v.kblk ? k0 / v.kblk : 0;
        const float* pk = v.base + (long long)kb * v.kbs + (long long)(k0 - kb * v.kblk) * v.s_k;     // element (row 0, k0)
        if (kc_map && vec != 0 && blk_ok && k0 + kGkKS <= p.K) {
          gk_load_slab_kc(pk + (row0 + kc_row) * v.s_r + kc_chunk * 4, 16 * v.s_r, rows_total - (row0 + kc_row), R / 16,
                          kc_chunk * (R * 16) + kc_row * 16, hi, hi + lo_off, wait);
        } else if (!kc_map && blk_ok) {
          rowsum += gk_load_slab_rc(pk, v.s_r, v.s_k, row0, rows_total, R, p.K - k0, tid, hi, hi + lo_off, wait);
        } else {
          rowsum += gk_load_slab_generic(v, row0, rows_total, R, k0, p.K, kc_map, vec != 0, tid, hi, hi + lo_off, wait);
        }
        fence_proxy_async();               // generic-proxy stores -> visible to the tensor-core (async) proxy
        mbar_arrive(full + stage);
      }
      // bias gradient of a weight-gradient problem: row m of A = dY^T is owned by this thread (row-contiguous mapping)
      if (is_a && p.colsum != nullptr && w.n_block == 0 && !kc_map) {
        const long long m = row0 + tid;
        if (m < p.M) atomicAdd(p.colsum + m, rowsum);
      }
    }
  } else {
    // ------------------------------ epilogue: one thread = one output row ------------------------------
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x, ++it) {
      const GkItem w = gk_decode(P, item);
      const GkProblem& g = P.pr[w.q];
      const acsr_gemm_problem& p = g.p;
      const int acc_stage = it & 1;
      const uint32_t t_acc = t_lane + acc_stage * kGkAccCols;
      const long long grow = (long long)w.m_tile * kGkBM + row;
      const bool row_ok = grow < p.M;
      const int n0 = w.n_block * g.BN;
      const int ncols = min(g.BN, p.N - n0);        // valid columns of this block
      const int nch = (ncols + 31) / 32;
      mbar_wait(tm_full + acc_stage, (it >> 1) & 1);
      tc_fence_after();
      if (p.epilogue == ACSR_EPI_BDRL) {
        // C = acc ; y = dropout(acc + bias) + res ; out = LN(y).  The pre-norm row y is parked in TMEM between the passes.
        const int N = p.N;
        const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
        const bool philox = p.p_drop > 0.f && p.mask == nullptr && p.rng != nullptr;
        unsigned long long seed = 0, step = 0;
        if (philox) { const RngState* r = reinterpret_cast<const RngState*>(p.rng); seed = r->seed; step = r->step; }
        const float* rr = row_ok ? p.res + (grow % p.res_rows) * N : p.res;
        float sum = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < nch; ++cc) {
          float v[32];
          tmem_ld32(t_acc + cc * 32, v);
          const int c0 = cc * 32;
          const int nvalid = min(32, N - c0);       // multiple of 4
          if (row_ok) {
            if (p.C != nullptr) {
              float* hz = p.C + grow * p.ldc + c0;
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (i < nvalid) *reinterpret_cast<float4*>(hz + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (q * 4 < nvalid) {
                float m[4] = {1.f, 1.f, 1.f, 1.f};
                if (p.mask != nullptr) {
                  const float4 mm = __ldg(reinterpret_cast<const float4*>(p.mask + grow * N + c0 + q * 4));
                  m[0] = mm.x; m[1] = mm.y; m[2] = mm.z; m[3] = mm.w;
                } else if (philox) {       // same counters as bdrl_{fwd,bwd}_kernel (rowwise.cu): element e -> call e>>2, word e&3
                  const uint4 wd = philox4x32(seed, step, p.rng_stream, ((unsigned long long)grow * N + c0) / 4 + q);
                  m[0] = drop_mult(wd.x, p.p_drop, inv_keep); m[1] = drop_mult(wd.y, p.p_drop, inv_keep);
                  m[2] = drop_mult(wd.z, p.p_drop, inv_keep); m[3] = drop_mult(wd.w, p.p_drop, inv_keep);
                }
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rr + c0 + q * 4));
                const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + q * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const int c = q * 4 + t;
                  v[c] = (v[c] + bv[t]) * m[t] + rv[t];
                  sum += v[c];
                }
              }
            }
          }
          tmem_st32(t_acc + cc * 32, v);
        }
        const float mean = sum / (float)N;
        float var = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < nch; ++cc) {
          float v[32];
          tmem_ld32(t_acc + cc * 32, v);
          const int nvalid = min(32, N - cc * 32);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) { const float t = v[i] - mean; var = fmaf(t, t, var); }
        }
        const float rstd = 1.0f / sqrtf(var / (float)N + p.eps);
#pragma unroll 1
        for (int cc = 0; cc < nch; ++cc) {
          float v[32];
          tmem_ld32(t_acc + cc * 32, v);
          const int c0 = cc * 32;
          const int nvalid = min(32, N - c0);
          if (row_ok) {
            float* o = p.out + grow * N + c0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (i < nvalid) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.ln_w + c0 + i));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln_b + c0 + i));
                *reinterpret_cast<float4*>(o + i) = make_float4((v[i] - mean) * rstd * w4.x + b4.x, (v[i + 1] - mean) * rstd * w4.y + b4.y,
                                                                (v[i + 2] - mean) * rstd * w4.z + b4.z, (v[i + 3] - mean) * rstd * w4.w + b4.w);
              }
            }
          }
        }
        if (row_ok) { p.stats[2 * grow] = mean; p.stats[2 * grow + 1] = rstd; }
      } else if (p.epilogue == ACSR_EPI_CE) {
        // full-catalogue logits of any hidden size (acsasrec.py:118-120): (max, sum exp) of this block's columns per row;
        // C2 = partial [M, n_blocks, 2], combined by acsr_ce_finalize.  The logits never leave TMEM.
        float run_m = -INFINITY, run_s = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < nch; ++cc) {
          float v[32];
          tmem_ld32(t_acc + cc * 32, v);
          const int nvalid = min(32, p.N - (n0 + cc * 32));
          float cm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < nvalid) cm = fmaxf(cm, v[i]);
          const float nm = fmaxf(run_m, cm);
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < nvalid) sum += __expf(v[i] - nm);
          run_s = run_s * __expf(run_m - nm) + sum;
          run_m = nm;
        }
        if (row_ok) {
          float* o = p.C2 + (grow * g.n_blocks + w.n_block) * 2;
          o[0] = run_m; o[1] = run_s;
        }
      } else if (p.epilogue == ACSR_EPI_CE_GRAD) {
        // C = Gt [N, ldc]: transpose of (exp(logit - lse[m]) - onehot(target[m])) * row_scale[m]; lse = res, row_scale = ln_w,
        // target = rng (int64).  A warp's 32 rows are 32 consecutive floats of a Gt row: coalesced 128-byte lines.
        float g_lse = 0.f, g_scale = 0.f;
        long long g_tgt = -1;
        if (row_ok) { g_lse = p.res[grow]; g_scale = p.ln_w[grow]; g_tgt = reinterpret_cast<const long long*>(p.rng)[grow]; }
#pragma unroll 1
        for (int cc = 0; cc < nch; ++cc) {
          float v[32];
          tmem_ld32(t_acc + cc * 32, v);
          const long long c0 = n0 + cc * 32;
          const int nvalid = min(32, (int)(p.N - c0));
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < nvalid) {
                float gv = __expf(v[i] - g_lse);
                if (c0 + i == g_tgt) gv -= 1.0f;
                p.C[(c0 + i) * p.ldc + grow] = gv * g_scale;
              }
            }
          }
        }
      } else {
#pragma unroll 1
        for (int cc = 0; cc < nch; ++cc) {
          float v[32];
          tmem_ld32(t_acc + cc * 32, v);
          const int c0 = n0 + cc * 32;
          const int nvalid = min(32, p.N - c0);
          if (row_ok) {
          float* y = p.C + grow * p.ldc + c0;
          const bool vec = nvalid == 32 && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
          if (p.epilogue == ACSR_EPI_ATOMIC) {
            if (vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) red_add_v4(y + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nvalid) atomicAdd(y + i, v[i]);
            }
          } else {
          float b[32];
          if (p.bias != nullptr) {
            const float* bp = p.bias + c0;
            if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + i));
                b[i] = b4.x; b[i + 1] = b4.y; b[i + 2] = b4.z; b[i + 3] = b4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) b[i] = i < nvalid ? __ldg(bp + i) : 0.f;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) b[i] = 0.f;
          }
          if (p.epilogue == ACSR_EPI_ACT) {
            // C = raw GEMM output (pre-bias, saved for the backward), C2 = act(C + bias)
            float* y2 = p.C2 + grow * p.ldc + c0;
            if (vec && ((reinterpret_cast<uintptr_t>(y2) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                *reinterpret_cast<float4*>(y + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                *reinterpret_cast<float4*>(y2 + i) = make_float4(act_fwd(p.act, v[i] + b[i]), act_fwd(p.act, v[i + 1] + b[i + 1]),
                                                                 act_fwd(p.act, v[i + 2] + b[i + 2]), act_fwd(p.act, v[i + 3] + b[i + 3]));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nvalid) { y[i] = v[i]; y2[i] = act_fwd(p.act, v[i] + b[i]); }
            }
          } else {
            if (vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 o = make_float4(v[i] + b[i], v[i + 1] + b[i + 1], v[i + 2] + b[i + 2], v[i + 3] + b[i + 3]);
                if (p.accumulate) {
                  const float4 old = *reinterpret_cast<const float4*>(y + i);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(y + i) = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nvalid) y[i] = v[i] + b[i] + (p.accumulate ? y[i] : 0.f);
            }
          }
          }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tm_empty + acc_stage);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<2 * kGkAccCols>(tmem_base);
  }
}

static bool gk_vec_ok(const float* base, long long s_r, long long s_k, long long kbs, int kblk) {
  return s_k == 1 && (s_r & 3) == 0 && (kbs & 3) == 0 && (kblk & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
}

}  // namespace acsr

using namespace acsr;

extern "C" {

int acsr_gemm_batch(const acsr_gemm_problem* problems, int n_problems, int passes, void* stream) {
  ACSR_REQUIRE(problems != nullptr && n_problems > 0 && n_problems <= ACSR_GEMM_MAX_PROBLEMS, "gemm_batch: 1..%d problems per launch",
               ACSR_GEMM_MAX_PROBLEMS);
  ACSR_REQUIRE(passes == 1 || passes == 3, "gemm_batch: passes must be 1 (TF32) or 3 (3xTF32)");
  GkParams P = {};
  P.passes = passes;
  int bn_max = 16;
  long long base_items_atomic = 0;
  int n = 0;
  for (int i = 0; i < n_problems; ++i) {
    const acsr_gemm_problem& p = problems[i];
    ACSR_REQUIRE(p.A && p.B, "gemm_batch[%d]: NULL operand", i);
    ACSR_REQUIRE(p.M >= 0 && p.N > 0 && p.K > 0 && p.M < (1ll << 31), "gemm_batch[%d]: bad sizes M=%lld N=%d K=%d", i, (long long)p.M, p.N, p.K);
    ACSR_REQUIRE(p.a_kblk > 0 && p.b_kblk > 0 && (p.a_kblk >= p.K || (p.a_kblk & 3) == 0) && (p.b_kblk >= p.K || (p.b_kblk & 3) == 0),
                 "gemm_batch[%d]: k-block lengths must be multiples of 4 (or cover K)", i);
    ACSR_REQUIRE(p.epilogue >= ACSR_EPI_STORE && p.epilogue <= ACSR_EPI_CE_GRAD, "gemm_batch[%d]: unknown epilogue %d", i, p.epilogue);
    if (p.M == 0) continue;
    GkProblem& g = P.pr[n];
    g.p = p;
    if (p.epilogue == ACSR_EPI_BDRL) {
      if (p.N > 256 || (p.N & 3)) { set_error("gemm_batch[%d]: LayerNorm epilogue needs N <= 256 and a multiple of 4 (N=%d)", i, p.N); return ACSR_ERR_UNSUPPORTED; }
      ACSR_REQUIRE(p.res && p.ln_w && p.ln_b && p.out && p.stats && p.res_rows > 0, "gemm_batch[%d]: LayerNorm epilogue: NULL pointer", i);
      ACSR_REQUIRE(p.p_drop >= 0.f && p.p_drop < 1.f, "gemm_batch[%d]: dropout p=%f", i, p.p_drop);
      ACSR_REQUIRE(!(p.p_drop > 0.f && p.mask == nullptr && p.rng == nullptr), "gemm_batch[%d]: p>0 needs mask or rng", i);
      ACSR_REQUIRE(p.C == nullptr || (p.ldc >= p.N && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0), "gemm_batch[%d]: C layout", i);
    } else if (p.epilogue == ACSR_EPI_CE) {
      ACSR_REQUIRE(p.C2 != nullptr, "gemm_batch[%d]: CE epilogue needs the partial buffer in C2", i);
    } else if (p.epilogue == ACSR_EPI_CE_GRAD) {
      ACSR_REQUIRE(p.C && p.res && p.ln_w && p.rng && p.ldc >= p.M, "gemm_batch[%d]: CE-gradient epilogue arguments", i);
    } else {
      ACSR_REQUIRE(p.C != nullptr && p.ldc >= p.N, "gemm_batch[%d]: bad output", i);
      if (p.epilogue == ACSR_EPI_ACT) {
        ACSR_REQUIRE(p.C2 != nullptr && p.act >= 0 && p.act <= 4, "gemm_batch[%d]: activation epilogue arguments", i);
      }
      if (p.epilogue == ACSR_EPI_ATOMIC && p.colsum != nullptr)
        ACSR_REQUIRE(p.a_k_stride != 1, "gemm_batch[%d]: colsum needs the transposed-operand form (a_k_stride != 1)", i);
    }
    g.m_tiles = (int)((p.M + kGkBM - 1) / kGkBM);
    g.n_blocks = (p.N + 255) / 256;
    const int per = (p.N + g.n_blocks - 1) / g.n_blocks;
    g.BN = (per + 15) & ~15;
    g.slabs = (p.K + kGkKS - 1) / kGkKS;
    g.a_vec = gk_vec_ok(p.A, p.a_row_stride, p.a_k_stride, p.a_kb_stride, p.a_kblk) ? 1 : 0;
    g.b_vec = gk_vec_ok(p.B, p.b_row_stride, p.b_k_stride, p.b_kb_stride, p.b_kblk) ? 1 : 0;
    if (g.BN > bn_max) bn_max = g.BN;
    if (p.epilogue == ACSR_EPI_ATOMIC) base_items_atomic += (long long)g.m_tiles * g.n_blocks;
    ++n;
  }
  if (n == 0) return ACSR_OK;
  P.n_problems = n;
  P.b_static = 1;
  for (int i = 0; i < n; ++i)
    if (!is_static_memory(P.pr[i].p.B)) P.b_static = 0;
  int items = 0;
  for (int i = 0; i < n; ++i) {
    GkProblem& g = P.pr[i];
    int ks = 1;
    if (g.p.epilogue == ACSR_EPI_ATOMIC) {
      ks = g.p.k_splits;
      if (ks <= 0) {                     // spread the K range so that the launch has ~2 work items per SM in total
        ks = (int)((2 * kNumSMs + base_items_atomic - 1) / base_items_atomic);
        const int cap = (g.slabs + 3) / 4;             // at least 4 slabs (128 contraction steps) per item
        if (ks > cap) ks = cap;
      }
      // the tensor core truncates when it accumulates: keep the MMA chain of one accumulator short (<= 1024 contraction steps)
      const int min_ks = (g.slabs + 31) / 32;
      if (ks < min_ks) ks = min_ks;
      if (ks < 1) ks = 1;
      if (ks > g.slabs) ks = g.slabs;
    }
    g.slabs_per_split = (g.slabs + ks - 1) / ks;
    g.ksplit = (g.slabs + g.slabs_per_split - 1) / g.slabs_per_split;
    g.item_begin = items;
    items += g.m_tiles * g.n_blocks * g.ksplit;
  }
  P.total_items = items;
  P.stage_bytes = 2 * kGkAOp + 2 * bn_max * (kGkKS * 4);
  int stages = (227 * 1024 - 256) / P.stage_bytes;
  if (stages > kGkMaxStages) stages = kGkMaxStages;
  P.stages = stages;
  // at least half of the SM's shared memory, so that two of these CTAs (2 x 512 TMEM columns) never share an SM
  size_t smem = (size_t)stages * P.stage_bytes + (2 * kGkMaxStages + 4) * 8 + 16;
  if (smem < 120 * 1024) smem = 120 * 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_ks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("gemm_batch: smem attr %zu: %s", smem, cudaGetErrorString(e)); return ACSR_ERR_CUDA; }
  const int grid = items < kNumSMs ? items : kNumSMs;
  launch_pdl(gemm_ks_kernel, dim3(grid), dim3(kGkThreads), smem, (cudaStream_t)stream, P);
  return check_launch("gemm_batch");
}

int acsr_gemm_ce_parts(int64_t V) { return V > 0 ? (int)((V + 255) / 256) : 0; }

}  // extern "C"

namespace acsr {
// full-catalogue logits for hidden sizes other than 64 (logits_tc.cu keeps the A-stationary width-64 kernel): the K-streamed GEMM
// with rows of `out` on M, table rows on N and the consumer in the TMEM epilogue
int gemm_logits(int mode, const float* out, const float* table, int M, long long V, int d, int passes, float* C, long long ldc,
                float* partial, const float* lse, const long long* target, const float* row_scale, cudaStream_t st) {
  acsr_gemm_problem p = {};
  p.A = out; p.a_row_stride = d; p.a_k_stride = 1; p.a_kblk = d;
  p.B = table; p.b_row_stride = d; p.b_k_stride = 1; p.b_kblk = d;
  p.M = M; p.N = (int)V; p.K = d;
  p.C = C; p.ldc = ldc; p.C2 = partial; p.epilogue = mode;
  p.res = lse; p.ln_w = row_scale; p.rng = target;
  return acsr_gemm_batch(&p, 1, passes, (void*)st);
}
}  // namespace acsr
