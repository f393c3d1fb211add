// Shared device helpers for the AC-SASRec sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define ACSR_OK 0
#define ACSR_ERR_ARG (-1)
#define ACSR_ERR_UNSUPPORTED (-2)
#define ACSR_ERR_CUDA (-3)

namespace acsr {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define ACSR_REQUIRE(cond, ...)                      \
  do {                                               \
    if (!(cond)) {                                   \
      acsr::set_error(__VA_ARGS__);                  \
      return ACSR_ERR_ARG;                           \
    }                                                \
  } while (0)

constexpr int kNumSMs = 148;
constexpr float kMaskNeg = -10000.0f;   // abstract_recommender.py:142
constexpr float kOrderEps = 1e-24f;     // layers.py:719

// ------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the step is a few microseconds long, so the
// launch gap between dependent kernels is a large share of the chain.  Kernels are launched with
// programmaticStreamSerializationAllowed and start with `pdl_launch_dependents(); ... pdl_wait();`:
// the next kernel's CTAs are scheduled (and run their data-independent prologue: barrier init, TMEM
// allocation, WEIGHT staging) while this one still runs; pdl_wait() returns once every kernel it
// depends on has completed and its writes are visible.  Only parameters, which no kernel but Adam
// (the last node of the step) writes, may be read before pdl_wait().  Off unless acsr_set_pdl(1) / ACSR_PDL=1:
// the fused training step and the eval forward turn it on for their own launches (measured +1.5 % / +1 %).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// memory registered with acsr_register_static: parameters that no kernel of a training step but the optimizer (its last node)
// writes, so a kernel launched with programmatic dependent launch may read them before griddepcontrol.wait
bool is_static_memory(const void* p);
bool pdl_enabled();
void pdl_set(int on);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG.  key = seed, counter = (idx_lo, idx_hi, stream, step).
// The same (seed, step, stream, idx) reproduces the same bits in forward and backward.
// ------------------------------------------------------------------------------------
struct RngState {          // lives in device memory so CUDA-graph replays see a fresh step
  unsigned long long seed;
  unsigned long long step;
};

__device__ __forceinline__ uint4 philox4x32(unsigned long long seed, unsigned long long step,
                                            unsigned int stream, unsigned long long idx) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = (uint32_t)step;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) {   // (0,1]
  return ((x >> 8) + 1) * (1.0f / 16777216.0f);
}

// inverted-dropout multiplier from 32 random bits
__device__ __forceinline__ float drop_mult(uint32_t bits, float p, float inv_keep) {
  return (u32_to_unit(bits) > p) ? inv_keep : 0.0f;
}

__device__ __forceinline__ float box_muller(uint32_t a, uint32_t b) {
  float u1 = u32_to_unit(a), u2 = u32_to_unit(b);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ------------------------------------------------------------------------------------
// warp / block reductions
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// activation ids shared with the host (layers.py:764-773)
enum Act { ACT_GELU = 0, ACT_RELU = 1, ACT_SWISH = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };

// Normal cdf / pdf for the erf GELU (layers.py:776-785) from ONE exponential: u = exp(-x^2/2) is the pdf (up to
// 1/sqrt(2 pi)) and also the tail factor of erf(x/sqrt 2) = 1 - poly(t) * u, t = 1/(1 + p |x|/sqrt 2)
// (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7: fp32 level).
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x) * 0.70710678118654752440f;
  float u;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(x * x * -0.72134752044448170368f));   // exp(-x^2/2)
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfa = 1.0f - poly * t * u;            // erf(|x|/sqrt 2)
  cdf = 0.5f * (1.0f + copysignf(erfa, x));
  pdf = 0.39894228040143267794f * u;
}

__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case ACT_GELU: { float cdf, pdf; gelu_parts(x, cdf, pdf); return x * cdf; }
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_SWISH: return x * sigmoidf_(x);
    case ACT_TANH: return tanhf(x);
    default: return sigmoidf_(x);
  }
}
__device__ __forceinline__ float act_bwd(int act, float x) {   // d act / d x
  switch (act) {
    case ACT_GELU: { float cdf, pdf; gelu_parts(x, cdf, pdf); return cdf + x * pdf; }
    case ACT_RELU: return x > 0.0f ? 1.0f : 0.0f;
    case ACT_SWISH: { float s = sigmoidf_(x); return s + x * s * (1.0f - s); }
    case ACT_TANH: { float t = tanhf(x); return 1.0f - t * t; }
    default: { float s = sigmoidf_(x); return s * (1.0f - s); }
  }
}

// ------------------------------------------------------------------------------------
// sm_100a async primitives: mbarrier, bulk copy (TMA engine, UBLKCP), tcgen05 / TMEM
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (TMA engine).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {         // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS));
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4, [16,30) LBO>>4 (stride between core matrices adjacent in K),
//   [32,46) SBO>>4 (stride between 8-row groups), [46,48) version=1, [61,64) layout=0.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (M rows = TMEM lanes, K = consecutive 32-bit columns) is read from tensor
// memory, where an epilogue left it with tcgen05.st (cute SM100_MMA_TF32_TS); issued by ONE thread.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread t <-> lane t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// registers -> TMEM: mirror of tmem_ld32 (the thread rewrites its own lane's 32 columns)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// vector reduction into global memory (no return value): 4 consecutive floats, 16-byte aligned
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MN-major, no-swizzle ("interleave") operand: core matrix = 8 k-rows x 16 bytes (4 consecutive M/N elements);
// SBO = stride between core matrices adjacent in M/N, LBO = stride between core matrices adjacent in K.
// Bit 15 / 16 of the instruction descriptor select MN-major for A / B.
__host__ __device__ constexpr uint32_t umma_idesc_tf32_major(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float to_tf32(float x) {   // round-to-nearest tf32, low 13 mantissa bits zero
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace acsr
