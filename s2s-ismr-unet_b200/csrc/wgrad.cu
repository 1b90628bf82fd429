// wgrad.cu — translation unit of the weight-gradient kernels (Conv2D and Conv2DTranspose).
#define S2S_KERNEL_IMPL
#include "wgrad.cuh"

namespace s2s {

int wgrad_run(int K, int S, const WgradArgs& a, int nslots, cudaStream_t st) {
    if (S == 1 && K == 3) return wgrad_dispatch<3, 1>(a, nslots, st);
    if (S == 2 && K == 2) return wgrad_dispatch<2, 2>(a, nslots, st);
    if (S == 2 && K == 3) return wgrad_dispatch<3, 2>(a, nslots, st);
    if (S == 2 && K == 5) return wgrad_dispatch<5, 2>(a, nslots, st);
    return fail(S2S_ERR_INVALID, "wgrad: unsupported kernel %d / stride %d", K, S);
}

int reduce_partials(const float* part, float* out, int64_t P, int nslots, cudaStream_t st) {
    return reduce_partials_impl(part, out, P, nslots, st);
}

}  // namespace s2s
