// one instantiation per translation unit: the long-sequence attention kernels are big (attn_long_impl.cuh)
#include "attn_long_impl.cuh"
namespace acsr {
int attn_long_bwd1_dh32(const AttnParams& p, cudaStream_t st) { return launch_long_bwd<32, 1>(p, st); }
}
