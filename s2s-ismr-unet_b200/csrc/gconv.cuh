// gconv.cuh — generic gather convolution on NHWC fp32 (CUDA-core FFMA path, strict fp32).
//
//   out[n, oy, ox, ca] = epi( sum_{ky,kx<K} sum_{cb<Cb} in[n, S*oy+ky-pad, S*ox+kx-pad, cb] * Wt[ky,kx,cb,ca] )
//
// One kernel covers three reference operators (Keras semantics, deep_nn_models.py:139-163):
//   * Conv2D 3x3 'same' forward          K=3 S=1 pad=1 wmode=0  epi = bias+ELU (+ BatchNorm batch statistics)
//   * Conv2D 3x3 input gradient (dgrad)   K=3 S=1 pad=1 wmode=1  epi = none | * ELU'(act)
//   * Conv2DTranspose(k, stride 2) dgrad  K=k S=2 pad=(k-2)/2 wmode=0 (kernel (k,k,Cout,Cin): cb=Cout, ca=Cin)
//
// Tiling: a CTA owns a TH x TW tile of output pixels of one image and CO_T = CG*CO_PT output
// channels; the contracted channels are streamed through shared memory in chunks of CI_T
// (input tile with halo as channel planes [cb][row][col], weights as [cb][tap][ca]).  A thread
// owns 4 consecutive pixels along W x CO_PT channels (register tile) so that input taps are read
// with LDS.128 and weights are warp-broadcast LDS.128.  KS > 1 splits each channel chunk over KS
// thread groups (for the deep, spatially tiny layers) with a fixed-order shared-memory reduction.
#pragma once
#include "common.cuh"

namespace s2s {

enum { EPI_NONE = 0, EPI_BIAS_ELU = 1, EPI_ELUGRAD = 2, EPI_BIAS = 3 };

struct GConvArgs {
    const float* in;  int ldin, in_coff, Hin, Win, Cb;
    const float* w;   int wmode;           // 0: [tap][cb][ca]   1: flipped taps, [tap][ca][cb]
    const float* bias;                     // [Ca] or null
    const float* aux; int ldaux;           // EPI_ELUGRAD: ELU output at the output positions
    float* out;       int ldout, out_coff, Hout, Wout, Ca;
    int pad, epi, tiles_x, tiles_y, N;
    // BatchNorm batch statistics of the output (training forward), nullable
    float* stat_part;                      // [slots][2][Ca]
    unsigned int* counter;
    // finalize (done by the last CTA)
    const float* gamma; const float* beta; // [Ca]
    float* mov_mean; float* mov_var;       // [Ca] updated in place when update_moving
    float* bn_mean; float* bn_rstd; float* bn_scale; float* bn_shift;  // [Ca] outputs
    float bn_eps, bn_momentum; int update_moving;
};

template <int K, int S, int TH, int TW, int CG, int CO_PT, int KS, int CI_T>
struct GConvCfg {
    static constexpr int PX = 4;
    static constexpr int PGX = TW / PX;
    static constexpr int PG = TH * PGX;
    static constexpr int NT = PG * CG * KS;
    static constexpr int CO_T = CG * CO_PT;
    static constexpr int IN_TH = S * (TH - 1) + K;
    static constexpr int IN_TW = S * (TW - 1) + K;
    static constexpr int SPAN = S * (PX - 1) + K;
    static constexpr int SPANV = (SPAN + 3) / 4;
    static constexpr int RP0 = (IN_TW + 3) / 4 * 4;
    static constexpr int RP1 = S * PX * (PGX - 1) + 4 * SPANV;
    static constexpr int RP = RP0 > RP1 ? RP0 : RP1;
    static constexpr int PS = IN_TH * RP;          // plane stride (multiple of 4)
    static constexpr int K2 = K * K;
    static constexpr int SIN = CI_T * PS;
    static constexpr int SW = CI_T * K2 * CO_T;
    static constexpr int SRED_KS = (KS - 1) * PG * CG * PX * CO_PT;
    static constexpr int SRED_ST = 2 * PG * CO_T;
    static constexpr int SMAIN = SIN + SW;
    static constexpr int SMEM0 = SMAIN > SRED_KS ? SMAIN : SRED_KS;
    static constexpr int SMEM = SMEM0 > SRED_ST ? SMEM0 : SRED_ST;
    static constexpr int CPK = CI_T / KS;          // channels of a chunk per k-slice
    static_assert(TW % PX == 0, "TW must be a multiple of 4");
    static_assert(CI_T % KS == 0, "CI_T must split evenly over KS");
    static_assert(CO_PT % 4 == 0, "CO_PT must be a multiple of 4");
    static_assert(NT <= 1024 && NT % 32 == 0, "bad thread count");
    static_assert(SMEM * 4 <= 48 * 1024, "static shared memory budget exceeded");
};

template <int K, int S, int TH, int TW, int CG, int CO_PT, int KS, int CI_T, bool STATS>
__global__ void __launch_bounds__(GConvCfg<K, S, TH, TW, CG, CO_PT, KS, CI_T>::NT)
gconv_kernel(const GConvArgs a) {
    using C = GConvCfg<K, S, TH, TW, CG, CO_PT, KS, CI_T>;
    __shared__ __align__(16) float smem[C::SMEM];
    float* sIn = smem;
    float* sW = smem + C::SIN;

    const int tid = threadIdx.x;
    const int pg = tid % C::PG;
    const int cg = (tid / C::PG) % CG;
    const int ks = tid / (C::PG * CG);
    const int ty = pg / C::PGX, tx = pg % C::PGX;

    const int tile = blockIdx.x;
    const int tile_y = tile / a.tiles_x, tile_x = tile % a.tiles_x;
    const int ca0 = blockIdx.y * C::CO_T;
    const int n = blockIdx.z;
    const int oy0 = tile_y * TH, ox0 = tile_x * TW;
    const int iy0 = S * oy0 - a.pad, ix0 = S * ox0 - a.pad;

    float acc[C::PX][CO_PT];
#pragma unroll
    for (int p = 0; p < C::PX; ++p)
#pragma unroll
        for (int j = 0; j < CO_PT; ++j) acc[p][j] = 0.f;

    const float* in_n = a.in + (size_t)n * a.Hin * a.Win * a.ldin + a.in_coff;
    const bool vec_in = ((a.Cb & 3) == 0) && ((a.ldin & 3) == 0) && ((a.in_coff & 3) == 0);

    for (int cb0 = 0; cb0 < a.Cb; cb0 += CI_T) {
        const int cbn = min(CI_T, a.Cb - cb0);
        __syncthreads();   // previous chunk fully consumed
        // ---- stage the input tile (with halo) as channel planes
        if (vec_in) {
            constexpr int Q = CI_T / 4;
            for (int idx = tid; idx < C::IN_TH * C::IN_TW * Q; idx += C::NT) {
                const int q = idx % Q;
                const int pix = idx / Q;
                const int c = pix % C::IN_TW, r = pix / C::IN_TW;
                const int iy = iy0 + r, ix = ix0 + c;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win && (cb0 + 4 * q) < a.Cb)
                    v = ld4(in_n + ((size_t)iy * a.Win + ix) * a.ldin + cb0 + 4 * q);
                float* d = sIn + (4 * q) * C::PS + r * C::RP + c;
                d[0] = v.x; d[C::PS] = v.y; d[2 * C::PS] = v.z; d[3 * C::PS] = v.w;
            }
        } else {
            for (int idx = tid; idx < C::IN_TH * C::IN_TW * CI_T; idx += C::NT) {
                const int cbl = idx % CI_T;
                const int pix = idx / CI_T;
                const int c = pix % C::IN_TW, r = pix / C::IN_TW;
                const int iy = iy0 + r, ix = ix0 + c;
                float v = 0.f;
                if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win && cbl < cbn)
                    v = __ldg(in_n + ((size_t)iy * a.Win + ix) * a.ldin + cb0 + cbl);
                sIn[cbl * C::PS + r * C::RP + c] = v;
            }
        }
        // ---- stage the weights as [cbl][tap][ca_l]
        if (a.wmode == 0) {
            for (int idx = tid; idx < CI_T * C::K2 * C::CO_T; idx += C::NT) {
                const int cal = idx % C::CO_T;
                const int tap = (idx / C::CO_T) % C::K2;
                const int cbl = idx / (C::CO_T * C::K2);
                float v = 0.f;
                if (cbl < cbn && ca0 + cal < a.Ca)
                    v = __ldg(a.w + ((size_t)tap * a.Cb + cb0 + cbl) * a.Ca + ca0 + cal);
                sW[idx] = v;
            }
        } else {
            for (int idx = tid; idx < CI_T * C::K2 * C::CO_T; idx += C::NT) {
                const int cbl = idx % CI_T;
                const int cal = (idx / CI_T) % C::CO_T;
                const int tap = idx / (CI_T * C::CO_T);
                float v = 0.f;
                if (cbl < cbn && ca0 + cal < a.Ca)
                    v = __ldg(a.w + ((size_t)(C::K2 - 1 - tap) * a.Ca + ca0 + cal) * a.Cb + cb0 + cbl);
                sW[(cbl * C::K2 + tap) * C::CO_T + cal] = v;
            }
        }
        __syncthreads();
        // ---- accumulate
        const float* sInT = sIn + (S * ty) * C::RP + S * C::PX * tx;
        const int cb_end = min((ks + 1) * C::CPK, cbn);
        for (int cbl = ks * C::CPK; cbl < cb_end; ++cbl) {
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                float iv[4 * C::SPANV];
#pragma unroll
                for (int v4 = 0; v4 < C::SPANV; ++v4) {
                    const float4 t = ld4(sInT + cbl * C::PS + ky * C::RP + 4 * v4);
                    iv[4 * v4 + 0] = t.x; iv[4 * v4 + 1] = t.y; iv[4 * v4 + 2] = t.z; iv[4 * v4 + 3] = t.w;
                }
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    float wv[CO_PT];
#pragma unroll
                    for (int j4 = 0; j4 < CO_PT / 4; ++j4) {
                        const float4 t = ld4(sW + (cbl * C::K2 + ky * K + kx) * C::CO_T + cg * CO_PT + 4 * j4);
                        wv[4 * j4 + 0] = t.x; wv[4 * j4 + 1] = t.y; wv[4 * j4 + 2] = t.z; wv[4 * j4 + 3] = t.w;
                    }
#pragma unroll
                    for (int p = 0; p < C::PX; ++p)
#pragma unroll
                        for (int j = 0; j < CO_PT; ++j) acc[p][j] = fmaf(iv[S * p + kx], wv[j], acc[p][j]);
                }
            }
        }
    }

    // ---- fixed-order reduction over the k-slices
    if (KS > 1) {
        __syncthreads();
        constexpr int GRP = C::PG * CG;
        const int g = cg * C::PG + pg;
        if (ks > 0) {
#pragma unroll
            for (int p = 0; p < C::PX; ++p)
#pragma unroll
                for (int j = 0; j < CO_PT; ++j)
                    smem[((p * CO_PT + j) * (KS - 1) + (ks - 1)) * GRP + g] = acc[p][j];
        }
        __syncthreads();
        if (ks == 0) {
#pragma unroll
            for (int p = 0; p < C::PX; ++p)
#pragma unroll
                for (int j = 0; j < CO_PT; ++j) {
                    float s = acc[p][j];
                    for (int k2 = 0; k2 < KS - 1; ++k2) s += smem[((p * CO_PT + j) * (KS - 1) + k2) * GRP + g];
                    acc[p][j] = s;
                }
        }
    }

    // ---- epilogue
    const int oy = oy0 + ty;
    const int cab = ca0 + cg * CO_PT;
    float ssum[CO_PT], ssq[CO_PT];
#pragma unroll
    for (int j = 0; j < CO_PT; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
    if (ks == 0 && oy < a.Hout) {
        float bv[CO_PT];
#pragma unroll
        for (int j = 0; j < CO_PT; ++j)
            bv[j] = (a.bias != nullptr && (a.epi == EPI_BIAS_ELU || a.epi == EPI_BIAS) && cab + j < a.Ca)
                        ? __ldg(a.bias + cab + j) : 0.f;
#pragma unroll
        for (int p = 0; p < C::PX; ++p) {
            const int ox = ox0 + C::PX * tx + p;
            if (ox >= a.Wout) continue;
            const size_t opix = ((size_t)n * a.Hout + oy) * a.Wout + ox;
#pragma unroll
            for (int j4 = 0; j4 < CO_PT / 4; ++j4) {
                const int ca = cab + 4 * j4;
                if (ca >= a.Ca) continue;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = acc[p][4 * j4 + e] + bv[4 * j4 + e];
                if (a.epi == EPI_BIAS_ELU) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = elu_f(v[e]);
                } else if (a.epi == EPI_ELUGRAD) {
                    const float4 y = ld4(a.aux + opix * a.ldaux + ca);
                    v[0] *= elu_grad_from_out(y.x); v[1] *= elu_grad_from_out(y.y);
                    v[2] *= elu_grad_from_out(y.z); v[3] *= elu_grad_from_out(y.w);
                }
                st4(a.out + opix * a.ldout + a.out_coff + ca, make_float4(v[0], v[1], v[2], v[3]));
                if (STATS) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) { ssum[4 * j4 + e] += v[e]; ssq[4 * j4 + e] += v[e] * v[e]; }
                }
            }
        }
    }

    if (STATS) {
        // CTA-level fixed-order reduction of (sum, sumsq) per channel, then last-CTA finalize
        __syncthreads();
        float* sS = smem;                       // [PG][CO_T]
        float* sQ = smem + C::PG * C::CO_T;     // [PG][CO_T]
        if (ks == 0) {
#pragma unroll
            for (int j = 0; j < CO_PT; ++j) {
                sS[pg * C::CO_T + cg * CO_PT + j] = ssum[j];
                sQ[pg * C::CO_T + cg * CO_PT + j] = ssq[j];
            }
        }
        __syncthreads();
        const int slot = n * (a.tiles_x * a.tiles_y) + tile;
        if (tid < 2 * C::CO_T) {
            const int which = tid / C::CO_T, c = tid % C::CO_T;
            const float* src = which ? sQ : sS;
            float s = 0.f;
            for (int g = 0; g < C::PG; ++g) s += src[g * C::CO_T + c];
            if (ca0 + c < a.Ca) a.stat_part[((size_t)slot * 2 + which) * a.Ca + ca0 + c] = s;
        }
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        if (cta_is_last(a.counter, total)) {
            const int nslots = a.N * a.tiles_x * a.tiles_y;
            const double M = (double)a.N * a.Hout * a.Wout;
            for (int c = tid; c < a.Ca; c += C::NT) {
                double s = 0.0, q = 0.0;
                for (int sl = 0; sl < nslots; ++sl) {
                    s += (double)__ldcg(a.stat_part + ((size_t)sl * 2 + 0) * a.Ca + c);
                    q += (double)__ldcg(a.stat_part + ((size_t)sl * 2 + 1) * a.Ca + c);
                }
                const double mean = s / M;
                double var = q / M - mean * mean;
                if (var < 0.0) var = 0.0;
                const float meanf = (float)mean, varf = (float)var;
                const float rstd = rsqrtf(varf + a.bn_eps);
                const float sc = a.gamma[c] * rstd;
                a.bn_mean[c] = meanf;
                a.bn_rstd[c] = rstd;
                a.bn_scale[c] = sc;
                a.bn_shift[c] = a.beta[c] - meanf * sc;
                if (a.update_moving) {
                    a.mov_mean[c] = a.mov_mean[c] * a.bn_momentum + meanf * (1.f - a.bn_momentum);
                    a.mov_var[c] = a.mov_var[c] * a.bn_momentum + varf * (1.f - a.bn_momentum);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ host-side dispatch
template <int K, int S, int TH, int TW, int CG, int CO_PT, int KS, int CI_T>
static int gconv_launch_cfg(GConvArgs a, cudaStream_t st) {
    using C = GConvCfg<K, S, TH, TW, CG, CO_PT, KS, CI_T>;
    a.tiles_x = cdiv(a.Wout, TW);
    a.tiles_y = cdiv(a.Hout, TH);
    dim3 grid(a.tiles_x * a.tiles_y, cdiv(a.Ca, C::CO_T), a.N);
    prof_begin(st, S == 2 ? "convT_dgrad" : (a.wmode == 1 ? "conv3x3_dgrad" : "conv3x3_fwd"),
               4.0 * a.N * ((double)a.Hin * a.Win * a.Cb + (double)a.Hout * a.Wout * a.Ca),
               2.0 * K * K * (double)a.Cb * a.Ca * a.N * a.Hout * a.Wout);
    if (a.stat_part)
        gconv_kernel<K, S, TH, TW, CG, CO_PT, KS, CI_T, true><<<grid, C::NT, 0, st>>>(a);
    else
        gconv_kernel<K, S, TH, TW, CG, CO_PT, KS, CI_T, false><<<grid, C::NT, 0, st>>>(a);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

// Tile / thread-group plan.  It must be a pure function of the shape because the caller sizes
// the BatchNorm statistics workspace from it (slots = N * tiles).
//   spatial tile 8x16 (8x8 for the deep levels); CTA = 128 threads (256 when the grid is thin):
//   CG channel groups x KS k-slices.  Prefer few channels per CTA until the grid covers 148 SMs.
struct GConvPlan { int th, tw, cg, copt, ks; };

static inline GConvPlan gconv_plan(int Hout, int Wout, int Ca, int Cb, int N) {
    GConvPlan p;
    const bool small = (Hout <= 8 && Wout <= 8);
    p.th = 8;
    p.tw = small ? 8 : 16;
    p.copt = (Ca % 8 == 0) ? 8 : 4;
    const int tiles = cdiv(Hout, p.th) * cdiv(Wout, p.tw) * N;
    const int budget = small ? 8 : 4;   // cg * ks for 128 threads
    int cg = 4;
    while (cg > 1 && (cg * p.copt > Ca || tiles * cdiv(Ca, cg * p.copt) < 148)) cg >>= 1;
    int ks = budget / cg;
    if (small && tiles * cdiv(Ca, cg * p.copt) < 148 && Cb >= 64 && ks * 2 <= 8) ks *= 2;
    p.cg = cg;
    p.ks = ks;
    return p;
}

static inline int gconv_stat_slots(int Hout, int Wout, int Ca, int Cb, int N) {
    const GConvPlan p = gconv_plan(Hout, Wout, Ca, Cb, N);
    return N * cdiv(Hout, p.th) * cdiv(Wout, p.tw);
}

#define S2S_GCONV_CASE(TH_, TW_, CG_, COPT_, KS_)                                              \
    if (p.th == TH_ && p.tw == TW_ && p.cg == CG_ && p.copt == COPT_ && p.ks == KS_)           \
        return gconv_launch_cfg<K, S, TH_, TW_, CG_, COPT_, KS_, 8>(a, st);

template <int K, int S, bool ALLOW_CO4>
static int gconv_dispatch(const GConvArgs& a, cudaStream_t st) {
    S2S_REQUIRE((a.Ca & 3) == 0 && (a.ldout & 3) == 0 && (a.out_coff & 3) == 0,
                "gconv: output channels/stride must be multiples of 4 (Ca=%d ld=%d off=%d)", a.Ca, a.ldout, a.out_coff);
    S2S_REQUIRE(a.epi != EPI_ELUGRAD || (a.aux != nullptr && (a.ldaux & 3) == 0), "gconv: bad aux");
    const GConvPlan p = gconv_plan(a.Hout, a.Wout, a.Ca, a.Cb, a.N);
    S2S_GCONV_CASE(8, 16, 4, 8, 1) S2S_GCONV_CASE(8, 16, 2, 8, 2) S2S_GCONV_CASE(8, 16, 1, 8, 4)
    S2S_GCONV_CASE(8, 8, 4, 8, 2) S2S_GCONV_CASE(8, 8, 2, 8, 4) S2S_GCONV_CASE(8, 8, 1, 8, 8)
    S2S_GCONV_CASE(8, 8, 4, 8, 4) S2S_GCONV_CASE(8, 8, 2, 8, 8)
    if constexpr (ALLOW_CO4) {
        S2S_GCONV_CASE(8, 16, 4, 4, 1) S2S_GCONV_CASE(8, 16, 2, 4, 2) S2S_GCONV_CASE(8, 16, 1, 4, 4)
        S2S_GCONV_CASE(8, 8, 4, 4, 2) S2S_GCONV_CASE(8, 8, 2, 4, 4) S2S_GCONV_CASE(8, 8, 1, 4, 8)
        S2S_GCONV_CASE(8, 8, 4, 4, 4) S2S_GCONV_CASE(8, 8, 2, 4, 8)
    }
    return fail(S2S_ERR_INVALID, "gconv: no kernel for plan th=%d tw=%d cg=%d copt=%d ks=%d (K=%d S=%d Ca=%d)",
                p.th, p.tw, p.cg, p.copt, p.ks, K, S, a.Ca);
}

}  // namespace s2s
