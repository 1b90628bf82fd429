// wgrad.cuh — generic weight-gradient reduction on NHWC fp32.
//
//   part[slot][tap][cb][ca] = sum over the slot's pixels p of  A[n,p,ca] * B[n, S*p + tap - pad, cb]
//
//   * Conv2D 3x3 wgrad:          A = dZ (ca = Cout), B = layer input X (cb = Cin), S=1, pad=1
//                                 -> Keras kernel layout (kh,kw,Cin,Cout) = [tap][cb][ca]
//   * Conv2DTranspose wgrad:     A = layer input x (ca = Cin), B = dY (cb = Cout), S=2, pad=(k-2)/2
//                                 -> Keras kernel layout (kh,kw,Cout,Cin) = [tap][cb][ca]
// Bias gradient (conv only): bias_part[slot][ca] = sum_p A[n,p,ca].
//
// Deterministic two-stage reduction: each CTA (slot) walks its share of pixel tiles, keeps
// K*K*4 accumulators per thread (thread = one cb x one quad of ca, one tile row per warp),
// reduces its warps in fixed order through shared memory and writes a partial; the partials are
// summed in slot order by reduce_partials_kernel / the fused Adam kernel (optim.cuh).
#pragma once
#include "common.cuh"
#include "gconv.cuh"

namespace s2s {

struct WgradArgs {
    const float* A; int ldA, coffA, HA, WA, Ca;
    const float* B; int ldB, coffB, HB, WB, Cb;
    int pad, N, nslots, tiles_x, tiles_y;
    float* part;        // [nslots][K*K*Cb*Ca]
    float* bias_part;   // [nslots][Ca] or null
};

#ifdef S2S_KERNEL_IMPL
template <int K, int S, int TH, int TW, int CB_T, int CAQ, int PH = 1>
struct WgradCfg {
    static constexpr int NW = TH;
    static constexpr int NT = 32 * NW;
    static constexpr int CA_T = 4 * CAQ;
    static constexpr int IN_TH = S * (TH - 1) + K;
    static constexpr int IN_TW = S * (TW - 1) + K;
    static constexpr int NPIXB = IN_TH * IN_TW;
    static constexpr int CSB = CB_T + 4;                            // padded pixel stride of the B tile
    static constexpr int SA = TH * TW * CA_T;
    static constexpr int SB = NPIXB * CSB;
    static constexpr int BUF = SA + SB;                             // one pipeline stage
    static constexpr int K2 = K * K;
    static constexpr int TCH = K2 < 9 ? K2 : 9;                     // taps reduced per round
    static constexpr int ROUNDS = (K2 + TCH - 1) / TCH;
    static constexpr int SRED = NW * TCH * 4 * 32;
    static constexpr int SMEM0 = (2 * BUF) > SRED ? (2 * BUF) : SRED;
    static constexpr int SMEM = SMEM0 + NW * PH * CA_T;             // + bias scratch
    static constexpr int OWN = 32 / PH;                             // (cb, ca quad) owners per warp
    static constexpr int TWP = TW / PH;                             // columns of a tile row per pixel-split lane group
    static_assert(CB_T * CAQ * PH == 32 && TW % PH == 0, "a warp covers CB_T x CAQ owners x PH column groups");
    static_assert(SMEM * 4 <= 200 * 1024, "shared memory budget exceeded");
};

// v2: both tiles are pixel-major in shared memory ([pixel][channel], like NHWC global memory) so that they are
// staged with 16-byte cp.async (zero-fill at the borders) and double-buffered across the CTA's tiles.
// PH > 1 (thin layers, Cb * Ca < 128): the lanes that would idle split the columns of the tile row instead
// (lane = ph * OWN + caq * CB_T + cbl) and are summed in the fixed-order reduction.
template <int K, int S, int TH, int TW, int CB_T, int CAQ, int PH, bool BIAS>
__global__ void __launch_bounds__(WgradCfg<K, S, TH, TW, CB_T, CAQ, PH>::NT)
wgrad_kernel(const WgradArgs a) {
    using C = WgradCfg<K, S, TH, TW, CB_T, CAQ, PH>;
    extern __shared__ float4 wg_smem4[];
    float* smem = reinterpret_cast<float*>(wg_smem4);
    float* sBias = smem + C::SMEM0;  // [NW][CA_T]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int cbl = lane % CB_T, caq = (lane / CB_T) % CAQ, ph = lane / C::OWN;
    const int c0 = ph * C::TWP;
    const int slot = blockIdx.x;
    const int cb0 = blockIdx.y * CB_T;
    const int ca0 = blockIdx.z * C::CA_T;
    const int tiles = a.tiles_x * a.tiles_y;
    const int total_tiles = a.N * tiles;

    float acc[C::K2][4];
#pragma unroll
    for (int t = 0; t < C::K2; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};

    const bool vecA = ((a.ldA & 3) == 0) && ((a.coffA & 3) == 0) && ((a.Ca & 3) == 0);
    const bool vecB = ((a.ldB & 3) == 0) && ((a.coffB & 3) == 0) && ((a.Cb & 3) == 0);

    auto stage = [&](int t, int b) {
        float* sA = smem + b * C::BUF;       // [TH*TW][CA_T]
        float* sB = sA + C::SA;              // [NPIXB][CSB]
        const int n = t / tiles, tl = t % tiles;
        const int py0 = (tl / a.tiles_x) * TH, px0 = (tl % a.tiles_x) * TW;
        const int by0 = S * py0 - a.pad, bx0 = S * px0 - a.pad;
        const float* An = a.A + (size_t)n * a.HA * a.WA * a.ldA + a.coffA;
        const float* Bn = a.B + (size_t)n * a.HB * a.WB * a.ldB + a.coffB;
        if (vecA) {
            for (int idx = tid; idx < TH * TW * CAQ; idx += C::NT) {
                const int q = idx % CAQ, pix = idx / CAQ;
                const int c = pix % TW, r = pix / TW;
                const int y = py0 + r, x = px0 + c;
                const bool ok = y < a.HA && x < a.WA && ca0 + 4 * q < a.Ca;
                cp_async16(sA + pix * C::CA_T + 4 * q, ok ? An + ((size_t)y * a.WA + x) * a.ldA + ca0 + 4 * q : a.A, ok);
            }
        } else {
            for (int idx = tid; idx < TH * TW * C::CA_T; idx += C::NT) {
                const int cl = idx % C::CA_T, pix = idx / C::CA_T;
                const int c = pix % TW, r = pix / TW;
                const int y = py0 + r, x = px0 + c;
                float v = 0.f;
                if (y < a.HA && x < a.WA && ca0 + cl < a.Ca) v = __ldg(An + ((size_t)y * a.WA + x) * a.ldA + ca0 + cl);
                sA[pix * C::CA_T + cl] = v;
            }
        }
        if (vecB) {
            constexpr int Q = CB_T / 4;
            for (int idx = tid; idx < C::NPIXB * Q; idx += C::NT) {
                const int q = idx % Q, pix = idx / Q;
                const int c = pix % C::IN_TW, r = pix / C::IN_TW;
                const int y = by0 + r, x = bx0 + c;
                const bool ok = y >= 0 && y < a.HB && x >= 0 && x < a.WB && cb0 + 4 * q < a.Cb;
                cp_async16(sB + pix * C::CSB + 4 * q, ok ? Bn + ((size_t)y * a.WB + x) * a.ldB + cb0 + 4 * q : a.B, ok);
            }
        } else {
            // first layer of the net (Cb = 1, 2, 3): scalar loads, U of them in flight per thread before the first
            // shared-memory store (this is the last weight-gradient kernel of a step, right in front of Adam)
            constexpr int U = 4;
            for (int idx0 = tid; idx0 < C::NPIXB * CB_T; idx0 += U * C::NT) {
                float v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = idx0 + u * C::NT;
                    const int cl = idx % CB_T, pix = idx / CB_T;
                    const int c = pix % C::IN_TW, r = pix / C::IN_TW;
                    const int y = by0 + r, x = bx0 + c;
                    const bool ok = idx < C::NPIXB * CB_T && y >= 0 && y < a.HB && x >= 0 && x < a.WB && cb0 + cl < a.Cb;
                    v[u] = ok ? __ldg(Bn + ((size_t)y * a.WB + x) * a.ldB + cb0 + cl) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = idx0 + u * C::NT;
                    if (idx < C::NPIXB * CB_T) sB[(idx / CB_T) * C::CSB + idx % CB_T] = v[u];
                }
            }
        }
        cp_async_commit();
    };

    int it = 0;
    if (slot < total_tiles) stage(slot, 0);
    for (int t = slot; t < total_tiles; t += a.nslots, ++it) {
        if (t + a.nslots < total_tiles) {
            stage(t + a.nslots, (it + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        // ---- accumulate: warp = tile row, lane = (cb, ca quad)
        const float* sA = smem + (it & 1) * C::BUF;
        const float* sBt = sA + C::SA + ((S * warp) * C::IN_TW + S * c0) * C::CSB + cbl;
        const float* sAt = sA + (warp * TW + c0) * C::CA_T + 4 * caq;
        // sliding K x K window of B along the row: each step loads only S new columns (K*S LDS instead of K*K)
        float win[K][K];
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K - S; ++kx) win[ky][kx + S] = sBt[(ky * C::IN_TW + kx) * C::CSB];
#pragma unroll
        for (int c = 0; c < C::TWP; ++c) {
            const float4 av = ld4(sAt + c * C::CA_T);
            if (BIAS) { bsum[0] += av.x; bsum[1] += av.y; bsum[2] += av.z; bsum[3] += av.w; }
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
#pragma unroll
                for (int kx = 0; kx < K - S; ++kx) win[ky][kx] = win[ky][kx + S];
#pragma unroll
                for (int kx = (K - S > 0 ? K - S : 0); kx < K; ++kx) win[ky][kx] = sBt[(ky * C::IN_TW + S * c + kx) * C::CSB];
            }
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float b = win[ky][kx];
                    acc[ky * K + kx][0] = fmaf(b, av.x, acc[ky * K + kx][0]);
                    acc[ky * K + kx][1] = fmaf(b, av.y, acc[ky * K + kx][1]);
                    acc[ky * K + kx][2] = fmaf(b, av.z, acc[ky * K + kx][2]);
                    acc[ky * K + kx][3] = fmaf(b, av.w, acc[ky * K + kx][3]);
                }
        }
        __syncthreads();       // the buffer just consumed is re-filled two iterations later
    }

    // ---- fixed-order reduction across the NW warps, TCH taps per round
    const size_t P = (size_t)C::K2 * a.Cb * a.Ca;
    float* part = a.part + (size_t)slot * P;
#pragma unroll
    for (int rd = 0; rd < C::ROUNDS; ++rd) {
        __syncthreads();
#pragma unroll
        for (int tl = 0; tl < C::TCH; ++tl) {
            const int tap = rd * C::TCH + tl;
            if (tap < C::K2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) smem[((warp * C::TCH + tl) * 4 + j) * 32 + lane] = acc[tap < C::K2 ? tap : 0][j];
            }
        }
        __syncthreads();
        for (int v = tid; v < C::TCH * 4 * C::OWN; v += C::NT) {
            const int ln = v % C::OWN, j = (v / C::OWN) & 3, tl = v / (4 * C::OWN);
            const int tap = rd * C::TCH + tl;
            if (tap >= C::K2) continue;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < C::NW; ++w)
#pragma unroll
                for (int p = 0; p < PH; ++p) s += smem[((w * C::TCH + tl) * 4 + j) * 32 + p * C::OWN + ln];
            const int cb = cb0 + ln % CB_T, ca = ca0 + 4 * (ln / CB_T) + j;
            if (cb < a.Cb && ca < a.Ca) part[((size_t)tap * a.Cb + cb) * a.Ca + ca] = s;
        }
    }
    if (BIAS) {
        if (blockIdx.y == 0) {
            if (cbl == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) sBias[(warp * PH + ph) * C::CA_T + 4 * caq + j] = bsum[j];
            }
            __syncthreads();
            if (tid < C::CA_T) {
                float s = 0.f;
                for (int w = 0; w < C::NW * PH; ++w) s += sBias[w * C::CA_T + tid];
                if (ca0 + tid < a.Ca) a.bias_part[(size_t)slot * a.Ca + ca0 + tid] = s;
            }
        }
    }
}

// out[i] = sum_{s < nslots} part[s*P + i]   (fixed slot order -> bit-reproducible gradients)
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t P, int nslots) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
    int sl = 0;
    for (; sl + 4 <= nslots; sl += 4) {
        const float v0 = __ldcg(part + (int64_t)(sl + 0) * P + i);
        const float v1 = __ldcg(part + (int64_t)(sl + 1) * P + i);
        const float v2 = __ldcg(part + (int64_t)(sl + 2) * P + i);
        const float v3 = __ldcg(part + (int64_t)(sl + 3) * P + i);
        s += v0; s += v1; s += v2; s += v3;
    }
    for (; sl < nslots; ++sl) s += __ldcg(part + (int64_t)sl * P + i);
    out[i] = s;
}

#endif  // S2S_KERNEL_IMPL

// ------------------------------------------------------------------ host-side dispatch
struct WgradPlan { int th, tw, cbt, caq, ph, nslots, ychunks, zchunks; };

static inline WgradPlan wgrad_plan(int HA, int WA, int Ca, int Cb, int N) {
    WgradPlan p;
    p.th = 8;
    p.tw = (WA <= 8) ? 8 : 16;
    p.ph = 1;
    if (Ca <= 8) { p.cbt = 16; p.caq = 2; } else { p.cbt = 8; p.caq = 4; }
    // thin layers: lanes without a (cb, ca) owner split the tile row's columns instead (PH)
    if (p.tw == 16) {
        if (Ca <= 8 && Cb <= 4) { p.cbt = 4; p.caq = 2; p.ph = 4; }
        else if (Ca <= 8 && Cb <= 8) { p.cbt = 8; p.caq = 2; p.ph = 2; }
        else if (Ca > 8 && Cb <= 4) { p.cbt = 4; p.caq = 4; p.ph = 2; }
    }
    p.ychunks = cdiv(Cb, p.cbt);
    p.zchunks = cdiv(Ca, 4 * p.caq);
    const int total_tiles = N * cdiv(HA, p.th) * cdiv(WA, p.tw);
    // ~4 CTAs per SM over the whole grid; the slot partials are reduced 8 warps wide by grad_reduce_adam
    int ns = cdiv(4 * 148, p.ychunks * p.zchunks);
    if (ns > (total_tiles + 1) / 2) ns = (total_tiles + 1) / 2;      // >= 2 tiles per CTA: the second hides behind the first
    if (ns > 256) ns = 256;
    if (ns < 1) ns = 1;
    p.nslots = ns;
    return p;
}

// Defined in wgrad.cu: nslots is fixed by the caller (it sized the partial workspace with wgrad_plan at max batch).
int wgrad_run(int K, int S, const WgradArgs& a, int nslots, cudaStream_t st);
int reduce_partials(const float* part, float* out, int64_t P, int nslots, cudaStream_t st);

#ifdef S2S_KERNEL_IMPL
template <int K, int S, int TH, int TW, int CB_T, int CAQ, int PH = 1>
static int wgrad_launch_cfg(WgradArgs a, const WgradPlan& p, cudaStream_t st) {
    using C = WgradCfg<K, S, TH, TW, CB_T, CAQ, PH>;
    a.tiles_x = cdiv(a.WA, TW);
    a.tiles_y = cdiv(a.HA, TH);
    a.nslots = p.nslots;
    dim3 grid(p.nslots, p.ychunks, p.zchunks);
    constexpr size_t smem_bytes = (size_t)C::SMEM * sizeof(float);
    static DevOnce once_b, once_n;
    if (a.bias_part)
        S2S_CUDA(once_b.run([] { return cudaFuncSetAttribute(wgrad_kernel<K, S, TH, TW, CB_T, CAQ, PH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); }));
    else
        S2S_CUDA(once_n.run([] { return cudaFuncSetAttribute(wgrad_kernel<K, S, TH, TW, CB_T, CAQ, PH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); }));
    prof_begin(st, S == 2 ? "convT_wgrad" : "conv3x3_wgrad",
               4.0 * a.N * ((double)a.HA * a.WA * a.Ca + (double)a.HB * a.WB * a.Cb),
               2.0 * K * K * (double)a.Cb * a.Ca * a.N * a.HA * a.WA);
    if (a.bias_part)
        wgrad_kernel<K, S, TH, TW, CB_T, CAQ, PH, true><<<grid, C::NT, smem_bytes, st>>>(a);
    else
        wgrad_kernel<K, S, TH, TW, CB_T, CAQ, PH, false><<<grid, C::NT, smem_bytes, st>>>(a);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

// nslots is fixed by the caller (it sized the partial workspace with wgrad_plan at max batch);
// slots that receive no tile simply write zeros.
template <int K, int S>
static int wgrad_dispatch(const WgradArgs& a, int nslots, cudaStream_t st) {
    WgradPlan p = wgrad_plan(a.HA, a.WA, a.Ca, a.Cb, a.N);
    p.nslots = nslots;
    if constexpr (S == 1) {
        if (p.tw == 16 && p.ph == 4) return wgrad_launch_cfg<K, S, 8, 16, 4, 2, 4>(a, p, st);
        if (p.tw == 16 && p.ph == 2 && p.cbt == 8) return wgrad_launch_cfg<K, S, 8, 16, 8, 2, 2>(a, p, st);
        if (p.tw == 16 && p.ph == 2 && p.cbt == 4) return wgrad_launch_cfg<K, S, 8, 16, 4, 4, 2>(a, p, st);
    } else if (p.ph != 1) {      // transposed-conv layers are never thin (Ca = 2 * Cb >= 8): keep the owner-only mapping
        p.ph = 1;
        if (a.Ca <= 8) { p.cbt = 16; p.caq = 2; } else { p.cbt = 8; p.caq = 4; }
        p.ychunks = cdiv(a.Cb, p.cbt); p.zchunks = cdiv(a.Ca, 4 * p.caq);
    }
    if (p.tw == 16 && p.cbt == 16) return wgrad_launch_cfg<K, S, 8, 16, 16, 2>(a, p, st);
    if (p.tw == 16 && p.cbt == 8) return wgrad_launch_cfg<K, S, 8, 16, 8, 4>(a, p, st);
    if (p.tw == 8 && p.cbt == 16) return wgrad_launch_cfg<K, S, 8, 8, 16, 2>(a, p, st);
    if (p.tw == 8 && p.cbt == 8) return wgrad_launch_cfg<K, S, 8, 8, 8, 4>(a, p, st);
    return fail(S2S_ERR_INVALID, "wgrad: no kernel for plan");
}

static inline int reduce_partials_impl(const float* part, float* out, int64_t P, int nslots, cudaStream_t st) {
    prof_begin(st, "reduce_partials", 4.0 * P * (nslots + 1), 0.0);
    reduce_partials_kernel<<<(unsigned)cdiv64(P, 256), 256, 0, st>>>(part, out, P, nslots);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

#endif  // S2S_KERNEL_IMPL

}  // namespace s2s
