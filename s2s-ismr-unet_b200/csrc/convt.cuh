// convt.cuh — Conv2DTranspose(k, strides=2, padding='same') forward on NHWC fp32.
//
// Keras semantics (deep_nn_models.py:154): kernel (kh,kw,Cout,Cin), output 2h x 2w,
//   y[n, oy, ox, co] = b[co] + sum_{ky,kx} [ (oy+pb-ky) even, (ox+pb-kx) even, in range ]
//                                x[n, (oy+pb-ky)/2, (ox+pb-kx)/2, ci] * W[ky,kx,co,ci],   pb = (k-2)/2
// The result is written into a channel slice of a wider NHWC buffer (ldy, y_coff): the second half
// of the skip-concat buffer, which fuses Concatenate()([c, u]) (deep_nn_models.py:156).
//
// Parity decomposition: a thread owns PX input positions (a, b) and produces their 2x2 output blocks
// (oy = 2a+dy, ox = 2b+dx) for CO_PT channels.  Tap ky feeds output row parity dy = (ky+pb)&1 from input
// row a + shift - ky/2 (shift = 1 only for k=5, even ky), all compile-time, so every thread runs the same
// k*k tap loop with warp-broadcast weights ([tap][ci][co], transposed once per step by wprep).
// Staging / tiling follow gconv.cuh: one-shot (or double-buffered) cp.async of the pixel-major input tile
// and the weight slab, optional k-split with a fixed-order reduction.
#pragma once
#include "common.cuh"
#include "gconv.cuh"

namespace s2s {

struct ConvTArgs {
    const float* x; int ldx, x_coff, h, w, Cin;
    const float* wt;               // [tap][ci][co]  (wprep of the Keras (k,k,Cout,Cin) kernel)
    const float* bias;
    float* y; int ldy, y_coff, Cout;
    int N, tiles_x, tiles_y, CG, KS, cbc, c4_shift;
};

#ifdef S2S_KERNEL_IMPL
template <int K>
struct ConvTGeo {
    static constexpr int PB = (K - 2) / 2;
    static constexpr int LO = (K == 2) ? 0 : 1;      // halo rows/cols before the position
    static constexpr int HI = (K == 5) ? 1 : 0;      // halo after
    static constexpr int WIN = LO + HI + 1;
    static constexpr int K2 = K * K;
    // tap k -> window index and output parity
    __host__ __device__ static constexpr int win(int k) { return (((k & 1) == 0 && PB == 1) ? 1 : 0) - k / 2 + LO; }
    __host__ __device__ static constexpr int par(int k) { return ((k & 1) + PB) & 1; }
};

template <int K, int TH, int TW, int PX, int CO_PT>
__global__ void __launch_bounds__(256) convt_fwd_kernel(const ConvTArgs a) {
    using T = ConvTGeo<K>;
    constexpr int PGX = TW / PX, PG = TH * PGX;
    constexpr int IN_TH = TH + T::LO + T::HI, IN_TW = TW + T::LO + T::HI, NPIX = IN_TH * IN_TW;
    static_assert(PG % 32 == 0, "pixel groups must fill warps");
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);

    const int tid = threadIdx.x, NT = blockDim.x;
    const int CG = a.CG, KS = a.KS, CO_T = CG * CO_PT;
    const int pg = tid % PG, cg = (tid / PG) % CG, ks = tid / (PG * CG);
    const int ty = pg / PGX, tx = pg % PGX;
    const int tile = blockIdx.x, tile_y = tile / a.tiles_x, tile_x = tile % a.tiles_x;
    const int co0 = blockIdx.y * CO_T, n = blockIdx.z;
    const int a0 = tile_y * TH, b0 = tile_x * TW;

    const int cbc = a.cbc, CS = cbc + 4;
    const int in_floats = NPIX * CS, buf_floats = in_floats + cbc * T::K2 * CO_T;
    const int nchunk = (a.Cin + cbc - 1) / cbc;
    const float* x_n = a.x + (size_t)n * a.h * a.w * a.ldx + a.x_coff;

    float acc[PX][4][CO_PT];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int j = 0; j < CO_PT; ++j) acc[p][o][j] = 0.f;

    auto stage = [&](int chunk, int b, int what = 3) {       // what: bit 0 = input tile, bit 1 = weight slab
        float* sIn = smem + b * buf_floats;
        float* sW = sIn + in_floats;
        const int cb0 = chunk * cbc, cbn = min(cbc, a.Cin - cb0), nq = cbn >> 2;
        if (what & 1) {
            const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
            const bool pow2 = (nq & (nq - 1)) == 0;
            const int nqs = __ffs(nq) - 1, nchunks = IN_TW * nq;
            const int ix0 = b0 - T::LO;
            for (int r = warp; r < IN_TH; r += nwarps) {
                const int iy = a0 - T::LO + r;
                const bool rok = iy >= 0 && iy < a.h;
                const float* grow = x_n + (iy * a.w + ix0) * a.ldx + cb0;
                float* srow = sIn + r * IN_TW * CS;
                for (int idx = lane; idx < nchunks; idx += 32) {
                    const int c = pow2 ? (idx >> nqs) : (idx / nq);
                    const int q = idx - c * nq;
                    const bool ok = rok && (unsigned)(ix0 + c) < (unsigned)a.w;
                    cp_async16(srow + c * CS + 4 * q, ok ? grow + c * a.ldx + 4 * q : a.x, ok);
                }
            }
        }
        const int c4s = a.c4_shift, c4n = 1 << c4s;
        const int c4 = tid & (c4n - 1);
        const bool cok = (co0 + 4 * c4) < a.Cout;
        for (int row = tid >> c4s; (what & 2) && row < cbn * T::K2; row += NT >> c4s) {
            const int tap = row % T::K2, cbl = row / T::K2;
            const float* src = cok ? a.wt + ((size_t)tap * a.Cin + cb0 + cbl) * a.Cout + co0 + 4 * c4 : a.wt;
            cp_async16(sW + row * CO_T + 4 * c4, src, cok);
        }
        if (what & 1) cp_async_commit();
    };

    auto compute = [&](int chunk, int b) {
        const float* sIn = smem + b * buf_floats;
        const float* sW = sIn + in_floats + cg * CO_PT;
        const int cbn = min(cbc, a.Cin - chunk * cbc), nq = cbn >> 2;
        const float* sInT = sIn + (ty * IN_TW + tx) * CS;
        for (int q = ks; q < nq; q += KS) {
            float4 xv[PX][T::WIN][T::WIN];
#pragma unroll
            for (int p = 0; p < PX; ++p)
#pragma unroll
                for (int r = 0; r < T::WIN; ++r)
#pragma unroll
                    for (int c = 0; c < T::WIN; ++c) xv[p][r][c] = ld4(sInT + (r * IN_TW + PGX * p + c) * CS + 4 * q);
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    constexpr int dummy = 0; (void)dummy;
                    const int o = T::par(ky) * 2 + T::par(kx);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float wv[CO_PT];
#pragma unroll
                        for (int j4 = 0; j4 < CO_PT / 4; ++j4) {
                            const float4 t = ld4(sW + ((4 * q + e) * T::K2 + ky * K + kx) * CO_T + 4 * j4);
                            wv[4 * j4 + 0] = t.x; wv[4 * j4 + 1] = t.y; wv[4 * j4 + 2] = t.z; wv[4 * j4 + 3] = t.w;
                        }
#pragma unroll
                        for (int p = 0; p < PX; ++p) {
                            const float4 v = xv[p][T::win(ky)][T::win(kx)];
                            const float xs = e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w;
#pragma unroll
                            for (int j = 0; j < CO_PT; ++j) acc[p][o][j] = fmaf(xs, wv[j], acc[p][o][j]);
                        }
                    }
                }
        }
    };

    // programmatic dependent launch: the transposed weights (written by wprep, joined through a full event edge) are staged
    // before the wait on the preceding kernel; dependents are released after the main loop (see gconv.cuh)
    stage(0, 0, 2);
    // the bias is a parameter (older than the preceding kernel): in registers before the wait instead of a global load per item
    // between the k-slice reduction and the stores
    constexpr int J4 = CO_PT / 4;
    const int cob = co0 + cg * CO_PT;
    float4 bpre[J4];
#pragma unroll
    for (int j4 = 0; j4 < J4; ++j4) {
        const int co = cob + 4 * j4;
        bpre[j4] = (a.bias && co < a.Cout) ? __ldg(reinterpret_cast<const float4*>(a.bias + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pdl_wait();
    stage(0, 0, 1);
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) { stage(c + 1, (c + 1) & 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        compute(c, c & 1);
        if (c + 1 < nchunk) __syncthreads();
    }
    pdl_trigger();

    // ---- epilogue: item = (position p, output parity o, channel quad j4), dealt round-robin to the k-slices
    constexpr int NITEMS = PX * 4 * J4;
    const int ay = a0 + ty;
    const int H2 = 2 * a.h, W2 = 2 * a.w;
    auto emit = [&](int p, int o, int j4, float4 v) {
        const int bx = b0 + tx + PGX * p;
        const int co = cob + 4 * j4;
        if (ay >= a.h || bx >= a.w || co >= a.Cout) return;
        {
            const float4 b = bpre[j4];
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        const size_t opix = ((size_t)n * H2 + 2 * ay + (o >> 1)) * W2 + 2 * bx + (o & 1);
        st4(a.y + opix * a.ldy + a.y_coff + co, v);
    };
    if (KS == 1) {
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int j4 = 0; j4 < J4; ++j4)
                    emit(p, o, j4, make_float4(acc[p][o][4 * j4], acc[p][o][4 * j4 + 1], acc[p][o][4 * j4 + 2], acc[p][o][4 * j4 + 3]));
    } else {
        __syncthreads();
        const int GRP = PG * CG, g = cg * PG + pg;
        float4* sR = reinterpret_cast<float4*>(smem);        // [ks][item][g]
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int j4 = 0; j4 < J4; ++j4)
                    sR[(ks * NITEMS + (p * 4 + o) * J4 + j4) * GRP + g] =
                        make_float4(acc[p][o][4 * j4], acc[p][o][4 * j4 + 1], acc[p][o][4 * j4 + 2], acc[p][o][4 * j4 + 3]);
        __syncthreads();
#pragma unroll
        for (int it0 = 0; it0 < NITEMS; ++it0) {
            if ((it0 & (KS - 1)) != ks) continue;
            float4 s = sR[(0 * NITEMS + it0) * GRP + g];
            for (int k2 = 1; k2 < KS; ++k2) {
                const float4 t = sR[(k2 * NITEMS + it0) * GRP + g];
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            emit(it0 / (4 * J4), (it0 / J4) & 3, it0 % J4, s);
        }
    }
}

#endif  // S2S_KERNEL_IMPL

struct ConvTPlan { int tw, copt, cg, ks, cbc; size_t smem; };

static inline ConvTPlan convt_plan(int k, int h, int w, int Cin, int Cout, int N) {
    ConvTPlan p;
    p.tw = (w <= 8) ? 8 : 16;
    p.copt = (Cout % 8 == 0) ? 8 : 4;
    const int pg = 8 * p.tw / 2;
    const int tiles = cdiv(h, 8) * cdiv(w, p.tw) * N;
    // fewer than one CTA per SM with 8 channels per thread (deep levels at batch 16: 64 CTAs): halve the channel tile
    if (p.copt == 8 && (long)tiles * cdiv(Cout, 8) < 148) p.copt = 4;
    p.cg = 1; p.ks = 1;
    auto warps = [&]() { return (long)tiles * cdiv(Cout, p.cg * p.copt) * (pg / 32) * p.cg * p.ks; };
    while (warps() < 148 * 6 && p.ks < 8 && pg * p.cg * p.ks * 2 <= 256 && Cin / 4 >= p.ks * 2) p.ks *= 2;
    while ((long)tiles * cdiv(Cout, p.cg * 2 * p.copt) >= 4 * 148 && p.cg * 2 * p.copt <= Cout && pg * p.cg * 2 * p.ks <= 256) p.cg *= 2;
    const int lo = k == 2 ? 0 : 1, hi = k == 5 ? 1 : 0;
    const int npix = (8 + lo + hi) * (p.tw + lo + hi), cot = p.cg * p.copt;
    auto bytes = [&](int cbc, int nbuf) { return (size_t)nbuf * ((size_t)npix * (cbc + 4) + (size_t)cbc * k * k * cot) * 4; };
    if (bytes(Cin, 1) <= 96 * 1024) p.cbc = Cin;
    else { int cbc = 64; while (cbc > 8 && bytes(cbc, 2) > 96 * 1024) cbc >>= 1; p.cbc = cbc; }
    p.smem = bytes(p.cbc, p.cbc == Cin ? 1 : 2);
    const size_t red = (size_t)p.ks * pg * p.cg * 2 * 4 * p.copt * 4;
    if (red > p.smem) p.smem = red;
    if (p.smem < 4096) p.smem = 4096;
    return p;
}

int convt_fwd(const ConvTArgs& a, int k, cudaStream_t st);     // defined in convt.cu

#ifdef S2S_KERNEL_IMPL
template <int K, int TW, int CO_PT>
static int convt_launch_cfg(ConvTArgs a, const ConvTPlan& p, cudaStream_t st) {
    constexpr int PX = 2, PG = 8 * TW / PX;
    a.tiles_x = cdiv(a.w, TW); a.tiles_y = cdiv(a.h, 8);
    a.CG = p.cg; a.KS = p.ks; a.cbc = p.cbc;
    { int c4n = p.cg * CO_PT / 4, sh = 0; while ((1 << sh) < c4n) ++sh; a.c4_shift = sh; }
    dim3 grid(a.tiles_x * a.tiles_y, cdiv(a.Cout, p.cg * CO_PT), a.N);
    static DevOnce once;
    S2S_CUDA(once.run([] { return cudaFuncSetAttribute(convt_fwd_kernel<K, 8, TW, PX, CO_PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); }));
    prof_begin(st, "convT_fwd", 4.0 * a.N * ((double)a.h * a.w * a.Cin + 4.0 * a.h * a.w * a.Cout),
               2.0 * K * K * (double)a.Cin * a.Cout * a.N * a.h * a.w);
    launch_k(convt_fwd_kernel<K, 8, TW, PX, CO_PT>, grid, PG * p.cg * p.ks, p.smem, st, a);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

template <int K>
static int convt_fwd_k(const ConvTArgs& a, cudaStream_t st) {
    const ConvTPlan p = convt_plan(K, a.h, a.w, a.Cin, a.Cout, a.N);
    if (p.tw == 8 && p.copt == 8) return convt_launch_cfg<K, 8, 8>(a, p, st);
    if (p.tw == 16 && p.copt == 8) return convt_launch_cfg<K, 16, 8>(a, p, st);
    if (p.tw == 8 && p.copt == 4) return convt_launch_cfg<K, 8, 4>(a, p, st);
    return convt_launch_cfg<K, 16, 4>(a, p, st);
}

static inline int convt_fwd_impl(const ConvTArgs& a, int k, cudaStream_t st) {
    S2S_REQUIRE((a.Cin & 3) == 0 && (a.Cout & 3) == 0 && (a.ldx & 3) == 0 && (a.ldy & 3) == 0 &&
                (a.y_coff & 3) == 0 && (a.x_coff & 3) == 0,
                "convt: channel counts/strides must be multiples of 4");
    if (k == 2) return convt_fwd_k<2>(a, st);
    if (k == 3) return convt_fwd_k<3>(a, st);
    if (k == 5) return convt_fwd_k<5>(a, st);
    return fail(S2S_ERR_INVALID, "convt: ct_kernel must be 2, 3 or 5 (got %d)", k);
}

#endif  // S2S_KERNEL_IMPL

}  // namespace s2s
