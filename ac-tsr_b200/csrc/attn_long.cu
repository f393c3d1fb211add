// Long-sequence (64 < L <= 256) calibrated attention: dispatch over the head size.  The kernels live in
// attn_long_impl.cuh and are instantiated one per translation unit (attn_long_{fwd,bwd1,bwd2}_dh{16,32,64}.cu).
#include "attn_common.cuh"

namespace acsr {

constexpr int kLongMaxL = 256;
int attn_long_fwd_dh16(const AttnParams& p, cudaStream_t st);
int attn_long_bwd1_dh16(const AttnParams& p, cudaStream_t st);
int attn_long_bwd2_dh16(const AttnParams& p, cudaStream_t st);
int attn_long_fwd_dh32(const AttnParams& p, cudaStream_t st);
int attn_long_bwd1_dh32(const AttnParams& p, cudaStream_t st);
int attn_long_bwd2_dh32(const AttnParams& p, cudaStream_t st);
int attn_long_fwd_dh64(const AttnParams& p, cudaStream_t st);
int attn_long_bwd1_dh64(const AttnParams& p, cudaStream_t st);
int attn_long_bwd2_dh64(const AttnParams& p, cudaStream_t st);

int attn_long_validate(const AttnParams& p, const char* who) {
  if (p.L > kLongMaxL) { set_error("%s: L=%d unsupported (1..%d)", who, p.L, kLongMaxL); return ACSR_ERR_UNSUPPORTED; }
  if (!(p.dh == 16 || p.dh == 32 || p.dh == 64)) {
    set_error("%s: head size %d unsupported for L > 64 (16/32/64)", who, p.dh);
    return ACSR_ERR_UNSUPPORTED;
  }
  return ACSR_OK;
}

int attn_long_fwd(const AttnParams& p, cudaStream_t st) {
  int rc = attn_long_validate(p, "attn_calib_fwd");
  if (rc) return rc;
  switch (p.dh) {
    case 16: return attn_long_fwd_dh16(p, st);
    case 32: return attn_long_fwd_dh32(p, st);
    case 64: return attn_long_fwd_dh64(p, st);
  }
  return ACSR_ERR_UNSUPPORTED;
}

int attn_long_bwd(const AttnParams& p, int ns, cudaStream_t st) {
  int rc = attn_long_validate(p, ns == 1 ? "attn_calib_bwd" : "attn_calib_bwd2");
  if (rc) return rc;
#define ACSR_LB(DHV) return ns == 1 ? attn_long_bwd1_dh##DHV(p, st) : attn_long_bwd2_dh##DHV(p, st)
  switch (p.dh) {
    case 16: ACSR_LB(16);
    case 32: ACSR_LB(32);
    case 64: ACSR_LB(64);
  }
#undef ACSR_LB
  return ACSR_ERR_UNSUPPORTED;
}

// sequences sorted by the number of keys in play, longest first (any L <= 1024; the L <= 64 kernel lives in attn_fwd.cu)
__global__ void __launch_bounds__(256) seq_order_long_kernel(const int64_t* __restrict__ item_seq, int B, int L, int32_t* __restrict__ order) {
  __shared__ int cnt[1026], off[1026];
  __shared__ unsigned short key[2048];
  for (int i = threadIdx.x; i < L + 2; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long* row = reinterpret_cast<const long long*>(item_seq) + (long long)b * L;
    int nkey = 0;
#pragma unroll 8
    for (int j = 0; j < L; ++j) if (__ldg(row + j) != 0) nkey = j + 1;
    key[b] = (unsigned short)nkey;
    atomicAdd(cnt + nkey, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = L + 1; k >= 0; --k) { off[k] = run; run += cnt[k]; }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) order[atomicAdd(off + key[b], 1)] = b;
}

int seq_order_long(const int64_t* item_seq, int B, int L, int32_t* order, cudaStream_t st) {
  seq_order_long_kernel<<<dim3(1), dim3(256), 0, st>>>(item_seq, B, L, order);
  return check_launch("seq_order");
}

}  // namespace acsr
