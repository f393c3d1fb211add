// one instantiation per translation unit: the long-sequence attention kernels are big (attn_long_impl.cuh)
#include "attn_long_impl.cuh"
namespace acsr {
int attn_long_bwd2_dh16(const AttnParams& p, cudaStream_t st) { return launch_long_bwd<16, 2>(p, st); }
}
