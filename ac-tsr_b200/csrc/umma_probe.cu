// Debug tool (not part of the product path): runs ONE tcgen05.mma.kind::tf32 128 x N x 8 on a caller-supplied shared-memory image
// with caller-supplied raw descriptors and returns the accumulator, so the operand layouts a descriptor combination
// really addresses can be measured on the device (scripts/umma_probe.py).  TMEM is pre-filled with a sentinel to tell a
// skipped instruction from one that wrote zeros.
#include "acsr_common.cuh"

namespace acsr {

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const float* __restrict__ image, int image_bytes, unsigned long long adesc,
                                                            unsigned long long bdesc, unsigned int idesc, int ncols, float sentinel,
                                                            float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < image_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = image[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = sentinel;
  for (int cc = 0; cc < 8; ++cc) tmem_st32(t_lane + cc * 32, v);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const unsigned long long base16 = (unsigned long long)(smem_u32(smem) >> 4);
    // the start-address fields (low 14 bits) are relative to the image: add the shared-memory base
    const unsigned long long ad = (adesc & ~0x3FFFull) | (((adesc & 0x3FFFull) + base16) & 0x3FFFull);
    const unsigned long long bd = (bdesc & ~0x3FFFull) | (((bdesc & 0x3FFFull) + base16) & 0x3FFFull);
    umma_tf32(tmem_base, ad, bd, idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int cc = 0; cc < ncols / 32; ++cc) {
    tmem_ld32(t_lane + cc * 32, v);
    for (int i = 0; i < 32; ++i) D[row * ncols + cc * 32 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tmem_base); }
}

}  // namespace acsr

extern "C" int acsr_debug_umma_probe(const float* image, int image_bytes, unsigned long long adesc, unsigned long long bdesc,
                                     unsigned int idesc, int ncols, float sentinel, float* D, void* stream) {
  using namespace acsr;
  if (image_bytes > 200 * 1024 || ncols > 256 || (ncols & 31)) return ACSR_ERR_ARG;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  umma_probe_kernel<<<1, 128, 200 * 1024, (cudaStream_t)stream>>>(image, image_bytes, adesc, bdesc, idesc, ncols, sentinel, D);
  return check_launch("umma_probe");
}

// Debug tool: dispatch rate of tcgen05.mma.kind::tf32 by shape and operand source.  One CTA per SM issues `n_mma` back-to-back
// 128 x N x 8 MMAs (K-major no-swizzle operands over zeroed shared memory; mode 1: A from tensor memory) from one thread, commits,
// and reports clock64 cycles from first issue to completion.  With `loaders` > 0 that many extra warps keep streaming LDS.128 /
// STS.128 over another shared-memory region, the traffic the splitter warps of the real kernels generate.
namespace acsr {
// `issuers` = 1 or 2: with 2, lane 0 of warp 0 and lane 0 of warp 7 each issue n_mma MMAs into their own accumulator (separate
// mbarriers), i.e. 2 * n_mma MMAs in total: does the tensor pipe take more than one thread can issue?
__global__ void __launch_bounds__(256, 1) umma_rate_kernel(int N, int mode, int n_mma, int loaders, int issuers, int unrolled,
                                                           long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_fence_init(); done = 0; }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool is_issuer = (threadIdx.x == 0) || (issuers == 2 && threadIdx.x == 7 * 32);
  if (is_issuer) {
    const int who = threadIdx.x == 0 ? 0 : 1;
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32 * 1024);
    const uint32_t a_lbo = 128 * 16, b_lbo = N * 16;
    const uint32_t d_tmem = tmem_base + who * 256 * (N <= 128 ? 1 : 0);      // N = 256: both issuers share the accumulator columns
    const uint32_t a_tmem = tmem_base + (N <= 128 ? 128 : 256);
    const long long t0 = clock64();
    if (unrolled) {
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bd = umma_desc_kmajor(b + ks * 2 * b_lbo, b_lbo, 128);
          if (mode == 1) umma_tf32_ts(d_tmem, a_tmem + ks * 8, bd, idesc, 1);
          else umma_tf32(d_tmem, umma_desc_kmajor(a + ks * 2 * a_lbo, a_lbo, 128), bd, idesc, 1);
        }
      }
    } else {
      for (int i = 0; i < n_mma; ++i) {
        const int ks = i & 7;
        const uint64_t bd = umma_desc_kmajor(b + ks * 2 * b_lbo, b_lbo, 128);
        if (mode == 1) umma_tf32_ts(d_tmem, a_tmem + ks * 8, bd, idesc, 1);
        else umma_tf32(d_tmem, umma_desc_kmajor(a + ks * 2 * a_lbo, a_lbo, 128), bd, idesc, 1);
      }
    }
    umma_commit(bar + who);
    mbar_wait(bar + who, 0);
    const long long t1 = clock64();
    if (who == 0) cycles[blockIdx.x] = t1 - t0;
    if (who == 0) done = 1;
  } else if (warp >= 1 && warp <= loaders) {
    float4* region = reinterpret_cast<float4*>(smem + 96 * 1024);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = threadIdx.x & 127;
    while (!done) {
#pragma unroll 8
      for (int r = 0; r < 8; ++r) {
        const float4 x = region[(i + r * 128) & 2047];
        acc.x += x.x; acc.y += x.y;
        region[2048 + ((i + r * 128) & 2047)] = acc;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}
}  // namespace acsr

extern "C" int acsr_debug_umma_rate(int N, int mode, int n_mma, int loaders, int issuers, int unrolled, long long* cycles, int n_ctas,
                                    void* stream) {
  using namespace acsr;
  if (N < 16 || N > 256 || (N & 15) || loaders < 0 || loaders > 6 || n_ctas < 1 || issuers < 1 || issuers > 2) return ACSR_ERR_ARG;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  umma_rate_kernel<<<n_ctas, 256, 160 * 1024, (cudaStream_t)stream>>>(N, mode, n_mma, loaders, issuers, unrolled, cycles);
  return check_launch("umma_rate");
}
