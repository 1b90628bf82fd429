// gconv.cuh — generic gather convolution on NHWC fp32 (CUDA-core FFMA path, strict fp32).
//
//   out[n, oy, ox, ca] = epi( sum_{ky,kx<K} sum_{cb<Cb} in[n, S*oy+ky-pad, S*ox+kx-pad, cb] * Wt[ky,kx,cb,ca] )
//
// One kernel covers three reference operators (Keras semantics, deep_nn_models.py:139-163):
//   * Conv2D 3x3 'same' forward          K=3 S=1 pad=1   epi = bias+ELU (+ BatchNorm batch-statistic partials)
//   * Conv2D 3x3 input gradient (dgrad)   K=3 S=1 pad=1   weights pre-flipped/transposed (wprep); epi = none | * ELU'(act)
//   * Conv2DTranspose(k, stride 2) dgrad  K=k S=2 pad=(k-2)/2 (kernel (k,k,Cout,Cin): cb=Cout, ca=Cin)
//
// Design (v2, latency-oriented: the layers are tiny at batch 16, see DESIGN.md "conv"):
//   * a CTA owns a TH x TW tile of output pixels of one image and CO_T = CG*CO_PT output channels;
//   * the input tile (with halo) and the weight slab are staged with cp.async (LDGSTS, 16 B, zero-fill
//     for the padding) in ONE shot when all contracted channels fit in shared memory, otherwise in
//     double-buffered chunks: all global loads are in flight together, one barrier, no register staging;
//   * shared layout is pixel-major [pixel][cb + 4 pad] (NHWC like global memory -> 16 B copies); a thread
//     owns PX pixels strided by TW/PX along x (conflict-free LDS.128 of 4 channels) x CO_PT channels;
//   * KS warp-groups split the channel quads of a chunk (k-split) and are reduced in fixed order;
//   * BatchNorm statistics: per-CTA (sum, sumsq) partials in fixed order; the CONSUMER (bn_apply) finalises
//     them, so there is no election / __threadfence tail in this kernel.
#pragma once
#include "common.cuh"

namespace s2s {

enum { EPI_NONE = 0, EPI_BIAS_ELU = 1, EPI_ELUGRAD = 2, EPI_BIAS = 3 };

struct GConvArgs {
    const float* in;  int ldin, in_coff, Hin, Win, Cb;
    const float* w;                        // [tap][cb][ca]
    const float* bias;                     // [Ca] or null
    const float* aux; int ldaux;           // EPI_ELUGRAD: ELU output at the output positions
    float* out;       int ldout, out_coff, Hout, Wout, Ca;
    int pad, epi, tiles_x, tiles_y, N;
    int act;                               // s2s_act_kind of EPI_BIAS_ELU / EPI_ELUGRAD
    int w_early;                           // weights may be staged before the programmatic-dependency wait
    int early_loads;                       // compile-time plans: the epilogue's global operands (bias, `aux`, stat_aux) are older than the
                                           // preceding kernel and are fetched into registers before the dependency wait
    int CG, KS, cbc, nbuf, c4_shift;       // runtime tiling: channel groups, k-slices, channel chunk, buffers, log2(CO_T/4)
    float* stat_part;                      // [slots][2][Ca] BatchNorm (sum, sumsq) partials, nullable
    // BatchNorm BACKWARD statistics instead (the transposed-conv input gradient IS the gradient dc wrt a BatchNorm output):
    // row 1 of the partials becomes sum dc * xhat, xhat = (stat_aux - mean) * rstd at the output position (bn.cuh, R1)
    const float* stat_aux; int ldstat;     // the BatchNorm layer's input (ELU output), dense [N,Hout,Wout,ldstat]; null = forward statistics
    const float* stat_mean; const float* stat_rstd;   // [Ca] batch statistics of that layer (published by bn_apply)
    // pooled layer (stat_pool != 0): stat_aux is the FULL-resolution [N,2Hout,2Wout,ldstat] input of the BatchNorm layer, `out` the
    // gradient of its 2x2-pooled output; stat_g1 (nullable) the skip-connection gradient at full resolution
    int stat_pool;                         // 0 = not pooled | 1 + S2S_POOL_AVG | 1 + S2S_POOL_MAX
    const float* stat_g1; int stat_ld1, stat_coff1;
    const float* stat_scale; const float* stat_shift;  // max pooling: the BatchNorm output decides which pixel of the window takes the gradient
};

#ifdef S2S_KERNEL_IMPL
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int K, int S, int TH, int TW, int PX>
struct GConvGeo {
    static constexpr int PGX = TW / PX;
    static constexpr int PG = TH * PGX;
    static constexpr int IN_TH = S * (TH - 1) + K;
    static constexpr int IN_TW = S * (TW - 1) + K;
    static constexpr int NPIX = IN_TH * IN_TW;
    static constexpr int K2 = K * K;
    static_assert(TW % PX == 0 && PG % 32 == 0, "pixel groups must fill warps");
};

// CBC / KST > 0: compile-time plan (the whole contraction in one chunk of CBC channels, KST k-slices, one channel group):
// every shared-memory offset of the inner loop and of the k-slice reduction becomes an immediate (-1/3 of the executed
// instructions at batch 16, profiles/experiments/conv_compile_time_tiling_static_analysis.md).  0 = runtime plan.
// STATS: 0 none | 1 BatchNorm forward statistics (sum v, sum v^2) | 2 BatchNorm backward statistics of an un-pooled layer
// (sum dc, sum dc*xhat with dc = v) | 3 of a pooled layer (v is the gradient of the 2x2-pooled output: dc = g1 + unpool(v))
template <int K, int S, int TH, int TW, int PX, int CO_PT, int STATS, int CBC = 0, int KST = 0>
__global__ void __launch_bounds__(256) gconv_kernel(const GConvArgs a) {
    using G = GConvGeo<K, S, TH, TW, PX>;
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);

    const int tid = threadIdx.x;
    const int NT = CBC ? G::PG * KST : blockDim.x;
    const int CG = CBC ? 1 : a.CG, KS = CBC ? KST : a.KS;
    const int CO_T = CG * CO_PT;
    const int pg = tid % G::PG;
    const int cg = (tid / G::PG) % CG;
    const int ks = tid / (G::PG * CG);
    const int ty = pg / G::PGX, tx = pg % G::PGX;

    const int tile = blockIdx.x;
    const int tile_y = tile / a.tiles_x, tile_x = tile % a.tiles_x;
    const int ca0 = blockIdx.y * CO_T;
    const int n = blockIdx.z;
    const int oy0 = tile_y * TH, ox0 = tile_x * TW;
    const int iy0 = S * oy0 - a.pad, ix0 = S * ox0 - a.pad;

    const int cbc = CBC ? CBC : a.cbc;     // channels per chunk (multiple of 4)
    const int CS = cbc + 4;                // padded pixel stride in shared memory
    const int in_floats = G::NPIX * CS;
    const int buf_floats = in_floats + cbc * G::K2 * CO_T;
    const int nchunk = CBC ? 1 : (a.Cb + cbc - 1) / cbc;
    const bool vec_in = ((a.Cb & 3) == 0) && ((a.ldin & 3) == 0) && ((a.in_coff & 3) == 0);
    const float* in_n = a.in + (size_t)n * a.Hin * a.Win * a.ldin + a.in_coff;

    float acc[PX][CO_PT];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int j = 0; j < CO_PT; ++j) acc[p][j] = 0.f;

    // ---- staging of one channel chunk into buffer `b`.  Index math is kept to shifts and compile-time
    // divisions: a thread owns channel quad (tid & 3) [+4, +8, ...] of pixels tid/4, tid/4 + NT/4, ...
    auto stage = [&](int chunk, int b, int what = 3) {       // what: bit 0 = input tile, bit 1 = weight slab
        float* sIn = smem + b * buf_floats;
        float* sW = sIn + in_floats;
        const int cb0 = chunk * cbc;
        const int cbn = min(cbc, a.Cb - cb0);                  // real channels of the chunk (the rest is zero padding)
        const int nq = CBC ? CBC / 4 : (cbn + 3) >> 2;
        if (!(what & 1)) {
        } else if (vec_in) {
            // row-wise: warp w stages tile rows w, w+nwarps, ...; lanes walk the (pixel, quad) chunks of a row
            const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
            const bool pow2 = (nq & (nq - 1)) == 0;
            const int nqs = __ffs(nq) - 1;
            const int nchunks = G::IN_TW * nq;
            for (int r = warp; r < G::IN_TH; r += nwarps) {
                const int iy = iy0 + r;
                const bool rok = iy >= 0 && iy < a.Hin;
                const float* grow = in_n + (iy * a.Win + ix0) * a.ldin + cb0;      // 32-bit offsets (tensors < 2^31 elements)
                float* srow = sIn + r * G::IN_TW * CS;
                for (int idx = lane; idx < nchunks; idx += 32) {
                    const int c = pow2 ? (idx >> nqs) : (idx / nq);
                    const int q = idx - c * nq;
                    const bool ok = rok && (unsigned)(ix0 + c) < (unsigned)a.Win;
                    cp_async16(srow + c * CS + 4 * q, ok ? grow + c * a.ldin + 4 * q : a.in, ok);
                }
            }
        } else if (nq == 1) {
            // thin first layer (Cin = 1, 2, 3): scalar loads, channels zero-padded to one quad.  The loads of U pixels are
            // issued together before the first shared-memory store (one global round trip per U pixels instead of one per
            // pixel: the first conv of the net spent most of its time in these dependent loads).
            constexpr int U = 8;
            const int cl = tid & 3, pstep = NT >> 2;
            const bool cok = cl < cbn;
            for (int pix0 = tid >> 2; pix0 < G::NPIX; pix0 += U * pstep) {
                float v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pix = pix0 + u * pstep;
                    const int c = pix % G::IN_TW, r = pix / G::IN_TW;
                    const int iy = iy0 + r, ix = ix0 + c;
                    const bool ok = cok && pix < G::NPIX && iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win;
                    v[u] = ok ? __ldg(in_n + (iy * a.Win + ix) * a.ldin + cb0 + cl) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pix = pix0 + u * pstep;
                    if (pix < G::NPIX) sIn[pix * CS + cl] = v[u];
                }
            }
        } else {   // channel counts that are no multiple of 4 (not produced by the U-Net itself): scalar loads
            const int cl = tid & 3;
            for (int pix = tid >> 2; pix < G::NPIX; pix += NT >> 2) {
                const int c = pix % G::IN_TW, r = pix / G::IN_TW;
                const int iy = iy0 + r, ix = ix0 + c;
                const bool ok = iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win;
                for (int ch = cl; ch < 4 * nq; ch += 4) {
                    float v = 0.f;
                    if (ok && ch < cbn) v = __ldg(in_n + ((size_t)iy * a.Win + ix) * a.ldin + cb0 + ch);
                    sIn[pix * CS + ch] = v;
                }
            }
        }
        // weights [cbl][tap][ca_l]; rows of padded channels are zero.  c4n = CO_T/4 is a power of two.
        const int c4s = CBC ? (CO_PT == 8 ? 1 : 0) : a.c4_shift, c4n = 1 << c4s;
        const int c4 = tid & (c4n - 1);
        const bool cok = (ca0 + 4 * c4) < a.Ca;
        for (int row = tid >> c4s; (what & 2) && row < 4 * nq * G::K2; row += NT >> c4s) {
            const int tap = row % G::K2, cbl = row / G::K2;
            const bool ok = cok && cbl < cbn;
            const float* src = ok ? a.w + ((size_t)tap * a.Cb + cb0 + cbl) * a.Ca + ca0 + 4 * c4 : a.w;
            cp_async16(sW + row * CO_T + 4 * c4, src, ok);
        }
        if (what & 1) cp_async_commit();
    };

    // ---- accumulate one chunk from buffer `b`
    auto compute = [&](int chunk, int b) {
        const float* sIn = smem + b * buf_floats;
        const float* sW = sIn + in_floats;
        const int cbn = CBC ? CBC : min(cbc, a.Cb - chunk * cbc);
        const int nq = (cbn + 3) >> 2;
        const float* sInT = sIn + ((S * ty) * G::IN_TW + S * tx) * CS + (CBC ? 4 * ks : 0);
        const float* sWt = sW + cg * CO_PT + (CBC ? 4 * ks * G::K2 * CO_T : 0);
        auto quad = [&](int q) {
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    float4 iv[PX];
#pragma unroll
                    for (int p = 0; p < PX; ++p)
                        iv[p] = ld4(sInT + (ky * G::IN_TW + S * G::PGX * p + kx) * CS + 4 * q);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float wv[CO_PT];
#pragma unroll
                        for (int j4 = 0; j4 < CO_PT / 4; ++j4) {
                            const float4 t = ld4(sWt + ((4 * q + e) * G::K2 + ky * K + kx) * CO_T + 4 * j4);
                            wv[4 * j4 + 0] = t.x; wv[4 * j4 + 1] = t.y; wv[4 * j4 + 2] = t.z; wv[4 * j4 + 3] = t.w;
                        }
#pragma unroll
                        for (int p = 0; p < PX; ++p) {
                            const float x = e == 0 ? iv[p].x : e == 1 ? iv[p].y : e == 2 ? iv[p].z : iv[p].w;
#pragma unroll
                            for (int j = 0; j < CO_PT; ++j) acc[p][j] = fmaf(x, wv[j], acc[p][j]);
                        }
                    }
                }
            }
        };
        if constexpr (CBC > 0) {
            // compile-time plan: q = ks + qi * KST with the ks part folded into the base pointers above
            constexpr int QN = (CBC / 4) / (KST ? KST : 1);
#pragma unroll
            for (int qi = 0; qi < QN; ++qi) quad(qi * KST);
        } else {
            for (int q = ks; q < nq; q += KS) quad(q);
        }
    };

    // Programmatic dependent launch: the weight slab does not depend on the preceding kernel (w_early: set by the caller
    // when the weights were last written at least two kernels ago), so its copies are in flight while that kernel drains;
    // only the input tile waits.  The next kernel is released after the main loop, when this one is down to its epilogue
    // (releasing it right after the staging copies were issued measured 407 vs 388 us per step at batch 16).
    if (a.w_early) stage(0, 0, 2);

    // Epilogue operands fetched up front (compile-time plans = the latency regime, where the registers are free): the bias, the
    // forward activation of the ELU' factor and of the BatchNorm-backward statistics are all older than the preceding kernel, so
    // their global round trip runs under the dependency wait and the main loop instead of between the k-slice reduction and the
    // stores (SASS: every item path had its own LDG -> MUFU -> STG chain at the very end of the kernel).
    constexpr int NITEMS = PX * (CO_PT / 4);
    constexpr bool PRE = CBC > 0;
    const int oy = oy0 + ty;
    const int cab = ca0 + cg * CO_PT;
    float4 pre[PRE ? NITEMS : 1], pre2[(PRE && STATS == 2) ? NITEMS : 1];
    const bool pre_on = PRE && a.early_loads != 0;
    if constexpr (PRE) {
        if (pre_on) {
#pragma unroll
            for (int it0 = 0; it0 < NITEMS; ++it0) {
                pre[it0] = make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (STATS == 2) pre2[it0] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (KS > 1 && (it0 & (KS - 1)) != ks) continue;            // the items this thread will emit
                const int ox = ox0 + tx + G::PGX * (it0 / (CO_PT / 4));
                const int ca = cab + 4 * (it0 % (CO_PT / 4));
                if (oy >= a.Hout || ox >= a.Wout || ca >= a.Ca) continue;
                const size_t opix = ((size_t)n * a.Hout + oy) * a.Wout + ox;
                if (a.epi == EPI_ELUGRAD) pre[it0] = ld4(a.aux + opix * a.ldaux + ca);
                else if (a.bias != nullptr && (a.epi == EPI_BIAS_ELU || a.epi == EPI_BIAS)) pre[it0] = __ldg(reinterpret_cast<const float4*>(a.bias + ca));
                if constexpr (STATS == 2) {      // xhat itself: mean / rstd are forward-pass values too
                    const float4 y = ld4(a.stat_aux + opix * a.ldstat + ca);
                    const float4 mu = __ldg(reinterpret_cast<const float4*>(a.stat_mean + ca));
                    const float4 rs = __ldg(reinterpret_cast<const float4*>(a.stat_rstd + ca));
                    pre2[it0] = make_float4((y.x - mu.x) * rs.x, (y.y - mu.y) * rs.y, (y.z - mu.z) * rs.z, (y.w - mu.w) * rs.w);
                }
            }
        }
    }
    pdl_wait();
    stage(0, 0, a.w_early ? 1 : 3);
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) {
            stage(c + 1, (c + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        compute(c, c & 1);
        if (c + 1 < nchunk) __syncthreads();    // buffer (c & 1) is re-filled by stage(c + 2)
    }
    pdl_trigger();

    // ---- epilogue.  An "item" is (pixel p, channel quad j4).  With KS > 1 every k-slice publishes its
    // accumulators to shared memory and the items are dealt round-robin to the KS slices, so the fixed-order
    // reduction, bias / ELU and the stores are spread over all threads instead of the ks == 0 warps only.
    float ssum[CO_PT], ssq[CO_PT];
#pragma unroll
    for (int j = 0; j < CO_PT; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }

    auto emit = [&](int p, int j4, float4 accv) {
        const int ox = ox0 + tx + G::PGX * p;
        const int ca = cab + 4 * j4;
        if (oy >= a.Hout || ox >= a.Wout || ca >= a.Ca) return;
        const size_t opix = ((size_t)n * a.Hout + oy) * a.Wout + ox;
        const int it = PRE ? p * (CO_PT / 4) + j4 : 0;
        float v[4] = {accv.x, accv.y, accv.z, accv.w};
        if (a.bias != nullptr && (a.epi == EPI_BIAS_ELU || a.epi == EPI_BIAS)) {
            const float4 b = pre_on ? pre[it] : __ldg(reinterpret_cast<const float4*>(a.bias + ca));
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        if (a.epi == EPI_BIAS_ELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = act_f(v[e], a.act);
        } else if (a.epi == EPI_ELUGRAD) {
            const float4 y = pre_on ? pre[it] : ld4(a.aux + opix * a.ldaux + ca);
            v[0] *= act_grad_from_out(y.x, a.act); v[1] *= act_grad_from_out(y.y, a.act);
            v[2] *= act_grad_from_out(y.z, a.act); v[3] *= act_grad_from_out(y.w, a.act);
        }
        st4(a.out + opix * a.ldout + a.out_coff + ca, make_float4(v[0], v[1], v[2], v[3]));
        if constexpr (STATS == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
#pragma unroll
                for (int jj = 0; jj < CO_PT / 4; ++jj)
                    if (jj == j4) { ssum[4 * jj + e] += v[e]; ssq[4 * jj + e] += v[e] * v[e]; }
            }
        } else if constexpr (STATS == 2) {
            float xh[4];
            if (pre_on) {
                xh[0] = pre2[it].x; xh[1] = pre2[it].y; xh[2] = pre2[it].z; xh[3] = pre2[it].w;
            } else {
                const float4 y = ld4(a.stat_aux + opix * a.ldstat + ca);
                const float4 mu = __ldg(reinterpret_cast<const float4*>(a.stat_mean + ca));
                const float4 rs = __ldg(reinterpret_cast<const float4*>(a.stat_rstd + ca));
                xh[0] = (y.x - mu.x) * rs.x; xh[1] = (y.y - mu.y) * rs.y; xh[2] = (y.z - mu.z) * rs.z; xh[3] = (y.w - mu.w) * rs.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
#pragma unroll
                for (int jj = 0; jj < CO_PT / 4; ++jj)
                    if (jj == j4) { ssum[4 * jj + e] += v[e]; ssq[4 * jj + e] = fmaf(v[e], xh[e], ssq[4 * jj + e]); }
            }
        } else if constexpr (STATS == 3) {
            // the four full-resolution pixels of this output's 2x2 window: dc_i = g1_i + unpool(v)_i  (bn.cuh, BnUnit<true>)
            const float4 mu = __ldg(reinterpret_cast<const float4*>(a.stat_mean + ca));
            const float4 rs = __ldg(reinterpret_cast<const float4*>(a.stat_rstd + ca));
            const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, rsv[4] = {rs.x, rs.y, rs.z, rs.w};
            float av[4][4], dcv[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const size_t pix = ((size_t)n * (2 * a.Hout) + 2 * oy + (i >> 1)) * (2 * a.Wout) + 2 * ox + (i & 1);
                const float4 y = ld4(a.stat_aux + pix * a.ldstat + ca);
                av[i][0] = y.x; av[i][1] = y.y; av[i][2] = y.z; av[i][3] = y.w;
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.stat_g1 != nullptr) g = ld4(a.stat_g1 + pix * a.stat_ld1 + a.stat_coff1 + ca);
                dcv[i][0] = g.x; dcv[i][1] = g.y; dcv[i][2] = g.z; dcv[i][3] = g.w;
            }
            if (a.stat_pool == 1 + S2S_POOL_AVG) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) dcv[i][e] = fmaf(v[e], 0.25f, dcv[i][e]);
            } else {      // the first maximum (row-major window order) of the BatchNorm output takes the gradient
                const float4 sc = __ldg(reinterpret_cast<const float4*>(a.stat_scale + ca));
                const float4 sh = __ldg(reinterpret_cast<const float4*>(a.stat_shift + ca));
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int best = 0;
                    float bv = fmaf(av[0][e], scv[e], shv[e]);
#pragma unroll
                    for (int i = 1; i < 4; ++i) {
                        const float t = fmaf(av[i][e], scv[e], shv[e]);
                        if (t > bv) { bv = t; best = i; }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i == best) dcv[i][e] += v[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) { s1 += dcv[i][e]; s2 = fmaf(dcv[i][e], (av[i][e] - muv[e]) * rsv[e], s2); }
#pragma unroll
                for (int jj = 0; jj < CO_PT / 4; ++jj)
                    if (jj == j4) { ssum[4 * jj + e] += s1; ssq[4 * jj + e] += s2; }
            }
        }
    };

    if (KS == 1) {
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int j4 = 0; j4 < CO_PT / 4; ++j4)
                emit(p, j4, make_float4(acc[p][4 * j4], acc[p][4 * j4 + 1], acc[p][4 * j4 + 2], acc[p][4 * j4 + 3]));
    } else {
        __syncthreads();                       // all slices are done with the staged tiles
        const int GRP = G::PG * CG;
        const int g = cg * G::PG + pg;
        float4* sR = reinterpret_cast<float4*>(smem);        // [ks][item][g]
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int j4 = 0; j4 < CO_PT / 4; ++j4)
                sR[(ks * NITEMS + p * (CO_PT / 4) + j4) * GRP + g] =
                    make_float4(acc[p][4 * j4], acc[p][4 * j4 + 1], acc[p][4 * j4 + 2], acc[p][4 * j4 + 3]);
        __syncthreads();
#pragma unroll
        for (int it0 = 0; it0 < NITEMS; ++it0) {
            if ((it0 & (KS - 1)) != ks) continue;   // items dealt round-robin (KS is a power of two; ks is per warp)
            float4 s = sR[(0 * NITEMS + it0) * GRP + g];
            for (int k2 = 1; k2 < KS; ++k2) {
                const float4 t = sR[(k2 * NITEMS + it0) * GRP + g];
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            emit(it0 / (CO_PT / 4), it0 % (CO_PT / 4), s);
        }
    }

    if (STATS) {
        // per-CTA (sum, sumsq) per channel: butterfly over the warp (fixed order), then over all warps of a
        // channel group through shared memory, written as one partial per CTA.  Finalised by bn_apply.
        __syncthreads();
        constexpr int WPG = G::PG / 32;      // warps per (channel group, k-slice)
        float* sS = smem;                    // [KS][CG][WPG][2][CO_PT]
        const int lane = tid & 31, wip = pg >> 5;
#pragma unroll
        for (int j = 0; j < CO_PT; ++j) {
            const float s = warp_sum(ssum[j]), q = warp_sum(ssq[j]);
            if (lane == 0) {
                sS[(((ks * CG + cg) * WPG + wip) * 2 + 0) * CO_PT + j] = s;
                sS[(((ks * CG + cg) * WPG + wip) * 2 + 1) * CO_PT + j] = q;
            }
        }
        __syncthreads();
        const int slot = n * (a.tiles_x * a.tiles_y) + tile;
        if (tid < 2 * CO_T) {
            const int which = tid / CO_T, c = tid % CO_T;
            const int g2 = c / CO_PT, j = c % CO_PT;
            float s = 0.f;
            for (int k2 = 0; k2 < KS; ++k2)
#pragma unroll
                for (int w = 0; w < WPG; ++w) s += sS[(((k2 * CG + g2) * WPG + w) * 2 + which) * CO_PT + j];
            if (ca0 + c < a.Ca) a.stat_part[((size_t)slot * 2 + which) * a.Ca + ca0 + c] = s;
        }
    }
}

#endif  // S2S_KERNEL_IMPL

// ------------------------------------------------------------------ host-side planning / dispatch
// The tile shape must be a pure function of the output shape: the caller sizes the BatchNorm statistics
// workspace from it (slots = N * tiles).
struct GConvPlan { int th, tw, px, copt, cg, ks, cbc, nbuf; size_t smem; };

// 8x8 for the deep levels, 8x16 by default (parallelism at batch 16), 16x32 once that tile still gives >= 1.5 CTAs per
// SM (large batches, 256x256 grids: fewer, fatter CTAs amortise the halo and the staging; +4..8 % per layer,
// tools/conv_bench.py).  S2S_BIGTILE_MIN_CTAS overrides the threshold (experiments).
static inline long gconv_bigtile_min() {
    static const long v = [] { const char* e = getenv("S2S_BIGTILE_MIN_CTAS"); return e ? atol(e) : 222L; }();
    return v;
}
static inline void gconv_tile(int Hout, int Wout, int N, int& th, int& tw) {
    th = 8;
    tw = (Hout <= 8 && Wout <= 8) ? 8 : 16;
    if (Hout >= 32 && Wout >= 32 && (long)N * cdiv(Hout, 16) * cdiv(Wout, 32) >= gconv_bigtile_min()) { th = 16; tw = 32; }
}
static inline int gconv_stat_slots(int Hout, int Wout, int N) {
    int th, tw;
    gconv_tile(Hout, Wout, N, th, tw);
    return N * cdiv(Hout, th) * cdiv(Wout, tw);
}
// upper bound over every batch size <= N (the small tile gives the most slots): sizes the workspace
static inline int gconv_stat_slots_max(int Hout, int Wout, int N) {
    const int tw = (Hout <= 8 && Wout <= 8) ? 8 : 16;
    return N * cdiv(Hout, 8) * cdiv(Wout, tw);
}

static inline GConvPlan gconv_plan(int K, int S, int Hout, int Wout, int Ca, int Cb, int N) {
    GConvPlan p;
    gconv_tile(Hout, Wout, N, p.th, p.tw);
    if (S != 1 && p.tw == 32) { p.th = 8; p.tw = 16; }      // strided gather (convT dgrad): the 16x32 input tile would not fit; no BN partials here
    const bool small = p.tw == 8;
    p.copt = (Ca % 8 == 0) ? 8 : 4;
    const int tiles = cdiv(Hout, p.th) * cdiv(Wout, p.tw) * N;
    // spread a layer over >= 8 warps per SM (k-split, 2 pixels per thread).  The fat-thread alternative (4 pixels, no
    // k-split) was measured slower even with 8-16 concurrent fits sharing the GPU (59.6k vs 73.9k samples/s).
    const long target_warps = 148 * 8;
    // pixels per thread: 4 when there is plenty of work, else 2 (small tiles always 2)
    p.px = small ? 2 : 4;
    p.cg = 1;
    p.ks = 1;
    auto warps = [&]() { return (long)tiles * cdiv(Ca, p.cg * p.copt) * (p.th * p.tw / p.px / 32) * p.cg * p.ks; };
    if (!small && warps() < target_warps) p.px = 2;
    const int pg = p.th * p.tw / p.px;
    const int nq = (Cb + 3) / 4;
    while (warps() < target_warps && p.ks < 8 && pg * p.cg * p.ks * 2 <= 256 && nq >= p.ks * 2) p.ks *= 2;
    // plenty of CTAs: give each more output channels (input tile reuse)
    while ((long)tiles * cdiv(Ca, p.cg * 2 * p.copt) >= 4 * 148 && p.cg * 2 * p.copt <= Ca && pg * p.cg * 2 * p.ks <= 256) p.cg *= 2;
    // channel chunk: everything at once when it fits in ~96 KB, else double-buffered chunks
    const int in_th = S * (p.th - 1) + K, in_tw = S * (p.tw - 1) + K;
    const int npix = in_th * in_tw, cot = p.cg * p.copt;
    auto bytes = [&](int cbc, int nbuf) { return (size_t)nbuf * ((size_t)npix * (cbc + 4) + (size_t)cbc * K * K * cot) * 4; };
    const int cb4 = (Cb + 3) / 4 * 4;
    if (bytes(cb4, 1) <= 96 * 1024) { p.cbc = cb4; p.nbuf = 1; }
    else {
        int cbc = 64;
        while (cbc > 8 && bytes(cbc, 2) > 96 * 1024) cbc >>= 1;
        p.cbc = cbc; p.nbuf = 2;
    }
    p.smem = bytes(p.cbc, p.nbuf);
    const size_t red = (size_t)p.ks * pg * p.cg * p.px * p.copt * 4;
    if (red > p.smem) p.smem = red;
    if (p.smem < 4096) p.smem = 4096;
    return p;
}

// Defined in gconv.cu (its own translation unit): K in {2,3,5}, S in {1,2}; allow_co4 admits Ca % 8 != 0.
int gconv_run(int K, int S, bool allow_co4, const GConvArgs& a, cudaStream_t st);

#ifdef S2S_KERNEL_IMPL
template <int K, int S, int TH, int TW, int PX, int CO_PT, int CBC = 0, int KST = 0>
static int gconv_launch_cfg(GConvArgs a, const GConvPlan& p, cudaStream_t st) {
    using G = GConvGeo<K, S, TH, TW, PX>;
    a.tiles_x = cdiv(a.Wout, TW);
    a.tiles_y = cdiv(a.Hout, TH);
    a.CG = p.cg; a.KS = p.ks; a.cbc = p.cbc; a.nbuf = p.nbuf;
    { int c4n = p.cg * CO_PT / 4, sh = 0; while ((1 << sh) < c4n) ++sh; a.c4_shift = sh; }
    const int NT = G::PG * p.cg * p.ks;
    dim3 grid(a.tiles_x * a.tiles_y, cdiv(a.Ca, p.cg * CO_PT), a.N);
    // statistics variant: forward (S = 1 only), un-pooled backward (the stride-2 gather = transposed-conv input gradient), pooled
    // backward (3x3 input gradient of the first conv of the next level)
    constexpr int SV_BWD = (S == 1) ? 3 : 2;
    const int sv = !a.stat_part ? 0 : (a.stat_aux ? SV_BWD : 1);
    if (sv == 1 && S != 1) return fail(S2S_ERR_INVALID, "gconv: forward statistics need stride 1");
    if (sv == 3 && !a.stat_pool) return fail(S2S_ERR_INVALID, "gconv: un-pooled backward statistics ride on the stride-2 kernel only");
    static DevOnce once_0, once_1, once_b;
    constexpr int SMEM_ATTR = 100 * 1024;
    prof_begin(st, S == 2 ? "convT_dgrad" : (a.epi == EPI_BIAS_ELU || a.epi == EPI_BIAS ? "conv3x3_fwd" : "conv3x3_dgrad"),
               4.0 * a.N * ((double)a.Hin * a.Win * a.Cb + (double)a.Hout * a.Wout * a.Ca),
               2.0 * K * K * (double)a.Cb * a.Ca * a.N * a.Hout * a.Wout);
    if (sv == 0) {
        S2S_CUDA(once_0.run([] { return cudaFuncSetAttribute(gconv_kernel<K, S, TH, TW, PX, CO_PT, 0, CBC, KST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ATTR); }));
        launch_k(gconv_kernel<K, S, TH, TW, PX, CO_PT, 0, CBC, KST>, grid, NT, p.smem, st, a);
    } else if (sv == SV_BWD) {
        S2S_CUDA(once_b.run([] { return cudaFuncSetAttribute(gconv_kernel<K, S, TH, TW, PX, CO_PT, SV_BWD, CBC, KST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ATTR); }));
        launch_k(gconv_kernel<K, S, TH, TW, PX, CO_PT, SV_BWD, CBC, KST>, grid, NT, p.smem, st, a);
    } else {
        if constexpr (S == 1) {
            S2S_CUDA(once_1.run([] { return cudaFuncSetAttribute(gconv_kernel<K, S, TH, TW, PX, CO_PT, 1, CBC, KST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ATTR); }));
            launch_k(gconv_kernel<K, S, TH, TW, PX, CO_PT, 1, CBC, KST>, grid, NT, p.smem, st, a);
        }
    }
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

template <int K, int S, bool ALLOW_CO4>
static int gconv_dispatch(const GConvArgs& a, cudaStream_t st) {
    S2S_REQUIRE((a.Ca & 3) == 0 && (a.ldout & 3) == 0 && (a.out_coff & 3) == 0,
                "gconv: output channels/stride must be multiples of 4 (Ca=%d ld=%d off=%d)", a.Ca, a.ldout, a.out_coff);
    S2S_REQUIRE(a.epi != EPI_ELUGRAD || (a.aux != nullptr && (a.ldaux & 3) == 0), "gconv: bad aux");
    const GConvPlan p = gconv_plan(K, S, a.Hout, a.Wout, a.Ca, a.Cb, a.N);
    if constexpr (K == 3 && S == 1) {
        // compile-time plans of the latency regime (batch 16): the whole contraction in one chunk, one channel group
        static const bool spec_on = getenv("S2S_NO_GCONV_SPEC") == nullptr;
        if (spec_on && p.copt == 8 && p.px == 2 && p.cg == 1 && p.nbuf == 1 && (a.Cb + 3) / 4 * 4 == p.cbc) {
#define S2S_GSPEC(CBCV, KSV)                                                                         \
            if (p.cbc == CBCV && p.ks == KSV) {                                                      \
                if (p.tw == 16) return gconv_launch_cfg<3, 1, 8, 16, 2, 8, CBCV, KSV>(a, p, st);     \
                if (p.tw == 8) return gconv_launch_cfg<3, 1, 8, 8, 2, 8, CBCV, KSV>(a, p, st);       \
            }
            S2S_GSPEC(4, 1) S2S_GSPEC(8, 1) S2S_GSPEC(8, 2) S2S_GSPEC(16, 2) S2S_GSPEC(16, 4) S2S_GSPEC(32, 2) S2S_GSPEC(32, 4)
            S2S_GSPEC(32, 8) S2S_GSPEC(64, 4) S2S_GSPEC(64, 8)
            // (compile-time plans for the filled-GPU regime - 16x32 tiles of 4 pixels per thread, one k-slice - were measured
            // and rejected: the fully unrolled bodies run the batch-128 step 5 % SLOWER, 1538 vs 1460 us on the same box,
            // profiles/r2_summary.md; those launches are throughput-bound and the compact runtime-plan loop is kinder to the
            // instruction cache)
#undef S2S_GSPEC
        }
    }
    if constexpr (K == 3 && S == 2) {
        // transposed-conv input gradient of the default ct_kernel = (3, 3) at batch 16: the same compile-time plans
        static const bool spec2_on = getenv("S2S_NO_GCONV_SPEC") == nullptr;
        if (spec2_on && p.copt == 8 && p.px == 2 && p.cg == 1 && p.nbuf == 1 && (a.Cb + 3) / 4 * 4 == p.cbc) {
            if (p.tw == 16 && p.cbc == 8 && p.ks == 2) return gconv_launch_cfg<3, 2, 8, 16, 2, 8, 8, 2>(a, p, st);
            if (p.tw == 16 && p.cbc == 16 && p.ks == 4) return gconv_launch_cfg<3, 2, 8, 16, 2, 8, 16, 4>(a, p, st);
            if (p.tw == 8 && p.cbc == 32 && p.ks == 8) return gconv_launch_cfg<3, 2, 8, 8, 2, 8, 32, 8>(a, p, st);
        }
    }
    if (p.copt == 8) {
        if (p.tw == 32) return gconv_launch_cfg<K, S, 16, 32, 4, 8>(a, p, st);
        if (p.tw == 16 && p.px == 4) return gconv_launch_cfg<K, S, 8, 16, 4, 8>(a, p, st);
        if (p.tw == 16 && p.px == 2) return gconv_launch_cfg<K, S, 8, 16, 2, 8>(a, p, st);
        if (p.tw == 8 && p.px == 2) return gconv_launch_cfg<K, S, 8, 8, 2, 8>(a, p, st);
    }
    if constexpr (ALLOW_CO4) {
        if (p.copt == 4) {
            if (p.tw == 32) return gconv_launch_cfg<K, S, 16, 32, 4, 4>(a, p, st);
            if (p.tw == 16 && p.px == 4) return gconv_launch_cfg<K, S, 8, 16, 4, 4>(a, p, st);
            if (p.tw == 16 && p.px == 2) return gconv_launch_cfg<K, S, 8, 16, 2, 4>(a, p, st);
            if (p.tw == 8 && p.px == 2) return gconv_launch_cfg<K, S, 8, 8, 2, 4>(a, p, st);
        }
    }
    return fail(S2S_ERR_INVALID, "gconv: no kernel for plan tw=%d px=%d copt=%d (K=%d S=%d Ca=%d)", p.tw, p.px, p.copt, K, S, a.Ca);
}

#endif  // S2S_KERNEL_IMPL

}  // namespace s2s
