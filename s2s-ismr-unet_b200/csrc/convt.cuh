// convt.cuh — Conv2DTranspose(k, strides=2, padding='same') forward, gather form, NHWC fp32.
//
// Keras semantics (deep_nn_models.py:154): kernel (kh,kw,Cout,Cin), output 2h x 2w,
//   y[n, oy, ox, co] = b[co] + sum_{ky,kx} [ (oy+pb-ky) even, (ox+pb-kx) even, in range ]
//                                x[n, (oy+pb-ky)/2, (ox+pb-kx)/2, ci] * W[ky,kx,co,ci],   pb = (k-2)/2
// The result is written into a channel slice of a wider NHWC buffer (ldy, y_coff): the second half
// of the skip-concat buffer, which fuses Concatenate()([c, u]) (deep_nn_models.py:156).
// Thread = one output pixel x 4 output channels; x rows and weight rows are contiguous over ci
// (float4 dot products, L1-resident weights).
#pragma once
#include "common.cuh"

namespace s2s {

struct ConvTArgs {
    const float* x; int ldx, x_coff, h, w, Cin;
    const float* wgt; const float* bias;
    float* y; int ldy, y_coff, Cout;
    int N;
};

template <int K>
__global__ void __launch_bounds__(256) convt_fwd_kernel(const ConvTArgs a) {
    constexpr int PB = (K - 2) / 2;
    const int H2 = 2 * a.h, W2 = 2 * a.w;
    const int CQ = a.Cout >> 2;
    const int64_t total = (int64_t)a.N * H2 * W2 * CQ;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int coq = (int)(idx % CQ);
    const int64_t pix = idx / CQ;
    const int ox = (int)(pix % W2);
    const int oy = (int)((pix / W2) % H2);
    const int n = (int)(pix / ((int64_t)W2 * H2));
    const int co = 4 * coq;

    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = a.bias ? __ldg(a.bias + co + j) : 0.f;

    const int ty = oy + PB, tx = ox + PB;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const int ry = ty - ky;
        if (ry < 0 || (ry & 1)) continue;
        const int iy = ry >> 1;
        if (iy >= a.h) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const int rx = tx - kx;
            if (rx < 0 || (rx & 1)) continue;
            const int ix = rx >> 1;
            if (ix >= a.w) continue;
            const float* xp = a.x + (((size_t)n * a.h + iy) * a.w + ix) * a.ldx + a.x_coff;
            const float* wp = a.wgt + ((size_t)(ky * K + kx) * a.Cout + co) * a.Cin;
            for (int ci = 0; ci < a.Cin; ci += 4) {
                const float4 xv = ld4(xp + ci);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * a.Cin + ci));
                    acc[j] = fmaf(xv.x, wv.x, acc[j]);
                    acc[j] = fmaf(xv.y, wv.y, acc[j]);
                    acc[j] = fmaf(xv.z, wv.z, acc[j]);
                    acc[j] = fmaf(xv.w, wv.w, acc[j]);
                }
            }
        }
    }
    st4(a.y + (size_t)pix * a.ldy + a.y_coff + co, make_float4(acc[0], acc[1], acc[2], acc[3]));
}

static inline int convt_fwd(const ConvTArgs& a, int k, cudaStream_t st) {
    S2S_REQUIRE((a.Cin & 3) == 0 && (a.Cout & 3) == 0 && (a.ldx & 3) == 0 && (a.ldy & 3) == 0 &&
                (a.y_coff & 3) == 0 && (a.x_coff & 3) == 0,
                "convt: channel counts/strides must be multiples of 4");
    const int64_t total = (int64_t)a.N * 4 * a.h * a.w * (a.Cout / 4);
    const unsigned grid = (unsigned)cdiv64(total, 256);
    prof_begin(st, "convT_fwd", 4.0 * a.N * ((double)a.h * a.w * a.Cin + 4.0 * a.h * a.w * a.Cout),
               2.0 * k * k * (double)a.Cin * a.Cout * a.N * a.h * a.w);
    if (k == 2) convt_fwd_kernel<2><<<grid, 256, 0, st>>>(a);
    else if (k == 3) convt_fwd_kernel<3><<<grid, 256, 0, st>>>(a);
    else if (k == 5) convt_fwd_kernel<5><<<grid, 256, 0, st>>>(a);
    else return fail(S2S_ERR_INVALID, "convt: ct_kernel must be 2, 3 or 5 (got %d)", k);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2s
