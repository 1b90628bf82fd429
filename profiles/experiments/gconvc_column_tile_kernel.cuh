// gconvc.cuh — 3x3 stride-1 gather convolution for THROUGHPUT sizes (large batches / 256x256 grids), strict fp32.
//
// Same operator and arguments as gconv.cuh (Conv2D forward with bias+ELU(+BN partials), Conv2D dgrad with ELU'),
// different register tile.  ncu on the latency-oriented kernel at batch 128 (profiles/r1c_*): FFMA was 47 % of the
// issued instructions and shared-memory wavefronts ran at ~75 % of the FFMA issue rate — every (tap, channel quad)
// re-loaded its 4 pixels with 4 LDS.128 for 128 FFMA.  Here a thread owns a COLUMN of 8 output pixels x 8 output
// channels (64 accumulators) and lanes run along x:
//   * for a channel quad and a kernel column kx the thread loads the 10 input rows of its column once (10 LDS.128,
//     conflict-free: lane stride = CS floats with CS/4 odd) and re-uses them for the 3 kernel rows ky:
//     34 LDS (10 inputs + 24 warp-broadcast weight quads) per 768 FFMA instead of 72;
//   * CTA = 2 warps = 16 rows x 32 columns (the tile gconv_tile() promises to the BatchNorm partial workspace);
//   * 8-channel chunks of the contraction are staged with cp.async (zero-fill = 'same' padding), double-buffered when
//     Cb > 8; 32 KB per buffer -> 7 CTAs / SM;
//   * epilogue straight from registers: bias, ELU | ELU'(aux), two 16-byte stores per pixel (a warp writes 1 KB
//     contiguous for Ca = 8), BatchNorm (sum, sumsq) by warp butterfly -> one partial per CTA in fixed order.
#pragma once
#include "gconv.cuh"

#ifdef S2S_KERNEL_IMPL
namespace s2s {

constexpr int GC_TH = 16, GC_TW = 32, GC_RP = 8;            // CTA tile, rows per thread
constexpr int GC_ITH = GC_TH + 2, GC_ITW = GC_TW + 2;       // haloed input tile
constexpr int GC_CBC = 8, GC_CS = GC_CBC + 4;               // channels per chunk, padded pixel stride (CS/4 odd)
constexpr int GC_IN_FLOATS = GC_ITH * GC_ITW * GC_CS;
constexpr int GC_W_FLOATS = GC_CBC * 9 * 8;
constexpr int GC_BUF_FLOATS = GC_IN_FLOATS + GC_W_FLOATS;

template <bool STATS>
__global__ void __launch_bounds__(64, 7) gconvc_kernel(const GConvArgs a) {
    extern __shared__ float4 gc_smem4[];
    float* smem = reinterpret_cast<float*>(gc_smem4);
    const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
    const int tile = blockIdx.x;
    const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
    const int ca0 = blockIdx.y * 8;
    const int n = blockIdx.z;
    const int oy0 = tile_y * GC_TH, ox0 = tile_x * GC_TW;
    const int iy0 = oy0 - 1, ix0 = ox0 - 1;
    const int nchunk = (a.Cb + GC_CBC - 1) / GC_CBC;
    const float* in_n = a.in + (size_t)n * a.Hin * a.Win * a.ldin + a.in_coff;

    float acc[GC_RP][8];
#pragma unroll
    for (int p = 0; p < GC_RP; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;

    auto stage = [&](int chunk, int b) {
        float* sIn = smem + b * GC_BUF_FLOATS;
        float* sW = sIn + GC_IN_FLOATS;
        const int cb0 = chunk * GC_CBC;
        const int nq = min(GC_CBC, a.Cb - cb0) >> 2;            // 1 or 2 quads (Cb % 4 == 0)
        // input: warp rg stages tile rows rg, rg+2, ...; lanes walk (pixel, quad)
        for (int r = rg; r < GC_ITH; r += 2) {
            const int iy = iy0 + r;
            const bool rok = iy >= 0 && iy < a.Hin;
            const float* grow = in_n + (iy * a.Win + ix0) * a.ldin + cb0;
            float* srow = sIn + r * GC_ITW * GC_CS;
            for (int idx = lane; idx < GC_ITW * 2; idx += 32) {
                const int c = idx >> 1, q = idx & 1;
                const bool ok = rok && q < nq && (unsigned)(ix0 + c) < (unsigned)a.Win;
                cp_async16(srow + c * GC_CS + 4 * q, ok ? grow + c * a.ldin + 4 * q : a.in, ok);
            }
        }
        // weights [cbl][tap][8]: 8 channels x 9 taps x 2 float4
        for (int i = tid; i < GC_CBC * 9 * 2; i += 64) {
            const int h4 = i & 1, row = i >> 1;
            const int tap = row % 9, cbl = row / 9;
            const bool ok = cbl < 4 * nq && (ca0 + 4 * h4) < a.Ca;
            const float* src = ok ? a.w + ((size_t)tap * a.Cb + cb0 + cbl) * a.Ca + ca0 + 4 * h4 : a.w;
            cp_async16(sW + row * 8 + 4 * h4, src, ok);
        }
        cp_async_commit();
    };

    auto compute = [&](int chunk, int b) {
        const float* sIn = smem + b * GC_BUF_FLOATS;
        const float* sW = sIn + GC_IN_FLOATS;
        const int nq = min(GC_CBC, a.Cb - chunk * GC_CBC) >> 2;
        const float* sCol = sIn + (GC_RP * rg * GC_ITW + lane) * GC_CS;      // haloed row 8*rg, column lane (+kx)
        for (int q = 0; q < nq; ++q) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float4 iv[GC_RP + 2];
#pragma unroll
                for (int i = 0; i < GC_RP + 2; ++i) iv[i] = ld4(sCol + (i * GC_ITW + kx) * GC_CS + 4 * q);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float* wp = sW + ((4 * q + e) * 9 + ky * 3 + kx) * 8;
                        const float4 w0 = ld4(wp), w1 = ld4(wp + 4);
#pragma unroll
                        for (int p = 0; p < GC_RP; ++p) {
                            const float4 v = iv[p + ky];
                            const float x = e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w;
                            acc[p][0] = fmaf(x, w0.x, acc[p][0]); acc[p][1] = fmaf(x, w0.y, acc[p][1]);
                            acc[p][2] = fmaf(x, w0.z, acc[p][2]); acc[p][3] = fmaf(x, w0.w, acc[p][3]);
                            acc[p][4] = fmaf(x, w1.x, acc[p][4]); acc[p][5] = fmaf(x, w1.y, acc[p][5]);
                            acc[p][6] = fmaf(x, w1.z, acc[p][6]); acc[p][7] = fmaf(x, w1.w, acc[p][7]);
                        }
                    }
                }
            }
        }
    };

    pdl_wait();
    pdl_trigger();
    stage(0, 0);
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) {
            stage(c + 1, (c + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        compute(c, c & 1);
        if (c + 1 < nchunk) __syncthreads();
    }

    // ---- epilogue from registers
    const int ox = ox0 + lane;
    const bool h0 = ca0 < a.Ca, h1 = ca0 + 4 < a.Ca;           // which channel quads exist (Ca % 4 == 0)
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    if (a.bias != nullptr && (a.epi == EPI_BIAS_ELU || a.epi == EPI_BIAS)) {
        if (h0) b0 = __ldg(reinterpret_cast<const float4*>(a.bias + ca0));
        if (h1) b1 = __ldg(reinterpret_cast<const float4*>(a.bias + ca0 + 4));
    }
    float ssum[8], ssq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
#pragma unroll
    for (int p = 0; p < GC_RP; ++p) {
        const int oy = oy0 + GC_RP * rg + p;
        if (oy >= a.Hout || ox >= a.Wout) continue;
        const size_t opix = ((size_t)n * a.Hout + oy) * a.Wout + ox;
        float v[8];
        v[0] = acc[p][0] + b0.x; v[1] = acc[p][1] + b0.y; v[2] = acc[p][2] + b0.z; v[3] = acc[p][3] + b0.w;
        v[4] = acc[p][4] + b1.x; v[5] = acc[p][5] + b1.y; v[6] = acc[p][6] + b1.z; v[7] = acc[p][7] + b1.w;
        if (a.epi == EPI_BIAS_ELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = elu_f(v[j]);
        } else if (a.epi == EPI_ELUGRAD) {
            if (h0) {
                const float4 y = ld4(a.aux + opix * a.ldaux + ca0);
                v[0] *= elu_grad_from_out(y.x); v[1] *= elu_grad_from_out(y.y); v[2] *= elu_grad_from_out(y.z); v[3] *= elu_grad_from_out(y.w);
            }
            if (h1) {
                const float4 y = ld4(a.aux + opix * a.ldaux + ca0 + 4);
                v[4] *= elu_grad_from_out(y.x); v[5] *= elu_grad_from_out(y.y); v[6] *= elu_grad_from_out(y.z); v[7] *= elu_grad_from_out(y.w);
            }
        }
        float* op = a.out + opix * a.ldout + a.out_coff + ca0;
        if (h0) st4(op, make_float4(v[0], v[1], v[2], v[3]));
        if (h1) st4(op + 4, make_float4(v[4], v[5], v[6], v[7]));
        if (STATS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { ssum[j] += v[j]; ssq[j] += v[j] * v[j]; }
        }
    }
    if (STATS) {
        __syncthreads();                       // staged tiles are dead: reuse shared memory
        float* sS = smem;                      // [2 warps][2][8]
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float s = warp_sum(ssum[j]), q = warp_sum(ssq[j]);
            if (lane == 0) { sS[(rg * 2 + 0) * 8 + j] = s; sS[(rg * 2 + 1) * 8 + j] = q; }
        }
        __syncthreads();
        const int slot = n * (a.tiles_x * a.tiles_y) + tile;
        if (tid < 16) {
            const int which = tid >> 3, j = tid & 7;
            const float s = sS[(0 * 2 + which) * 8 + j] + sS[(1 * 2 + which) * 8 + j];
            if (ca0 + j < a.Ca) a.stat_part[((size_t)slot * 2 + which) * a.Ca + ca0 + j] = s;
        }
    }
}

static inline bool gconvc_eligible(const GConvArgs& a) {
    int th, tw;
    gconv_tile(a.Hout, a.Wout, a.N, th, tw);
    return tw == GC_TW && th == GC_TH && (a.Cb & 3) == 0 && (a.ldin & 3) == 0 && (a.in_coff & 3) == 0 && a.pad == 1 &&
           a.Hin == a.Hout && a.Win == a.Wout;
}

static int gconvc_launch(GConvArgs a, cudaStream_t st) {
    a.tiles_x = cdiv(a.Wout, GC_TW);
    a.tiles_y = cdiv(a.Hout, GC_TH);
    const int nbuf = a.Cb > GC_CBC ? 2 : 1;
    const size_t smem = (size_t)nbuf * GC_BUF_FLOATS * sizeof(float);
    dim3 grid(a.tiles_x * a.tiles_y, cdiv(a.Ca, 8), a.N);
    static bool attr = false;
    if (!attr) {
        S2S_CUDA(cudaFuncSetAttribute(gconvc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * GC_BUF_FLOATS * (int)sizeof(float)));
        S2S_CUDA(cudaFuncSetAttribute(gconvc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * GC_BUF_FLOATS * (int)sizeof(float)));
        attr = true;
    }
    prof_begin(st, a.epi == EPI_BIAS_ELU || a.epi == EPI_BIAS ? "conv3x3_fwd" : "conv3x3_dgrad",
               4.0 * a.N * ((double)a.Hin * a.Win * a.Cb + (double)a.Hout * a.Wout * a.Ca),
               18.0 * (double)a.Cb * a.Ca * a.N * a.Hout * a.Wout);
    if (a.stat_part) launch_k(gconvc_kernel<true>, grid, 64, smem, st, a);
    else launch_k(gconvc_kernel<false>, grid, 64, smem, st, a);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2s
#endif  // S2S_KERNEL_IMPL
