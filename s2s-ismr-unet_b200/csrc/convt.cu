// convt.cu — translation unit of the Conv2DTranspose forward kernels.
#define S2S_KERNEL_IMPL
#include "convt.cuh"

namespace s2s {

int convt_fwd(const ConvTArgs& a, int k, cudaStream_t st) { return convt_fwd_impl(a, k, st); }

}  // namespace s2s
