// gconv.cu — translation unit of the gather-convolution kernels (conv fwd / dgrad / transposed-conv dgrad).
#define S2S_KERNEL_IMPL
#include "gconv.cuh"

namespace s2s {

int gconv_run(int K, int S, bool allow_co4, const GConvArgs& a, cudaStream_t st) {
    if (S == 1 && K == 3) return allow_co4 ? gconv_dispatch<3, 1, true>(a, st) : gconv_dispatch<3, 1, false>(a, st);
    if (S == 2 && K == 2) return gconv_dispatch<2, 2, false>(a, st);
    if (S == 2 && K == 3) return gconv_dispatch<3, 2, false>(a, st);
    if (S == 2 && K == 5) return gconv_dispatch<5, 2, false>(a, st);
    return fail(S2S_ERR_INVALID, "gconv: unsupported kernel %d / stride %d", K, S);
}

}  // namespace s2s
