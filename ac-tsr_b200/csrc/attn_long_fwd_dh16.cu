// one instantiation per translation unit: the long-sequence attention kernels are big (attn_long_impl.cuh)
#include "attn_long_impl.cuh"
namespace acsr {
int attn_long_fwd_dh16(const AttnParams& p, cudaStream_t st) { return launch_long_fwd<16>(p, st); }
}
