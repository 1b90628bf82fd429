// bn.cuh — BatchNormalization apply (+ 2x2 pooling + skip write), BN backward, channel sums.
//
// Keras semantics (deep_nn_models.py:92,147-148,162): BN over (N,H,W) per
// channel with eps 1e-3, biased variance; y = x*scale + shift with scale = gamma*rsqrt(var+eps),
// shift = beta - mean*scale; AveragePooling2D(2) (default) or MaxPooling2D(2).
// Batch statistics themselves are produced by the conv epilogue (gconv.cuh, STATS).
//
// All kernels here are HBM/L2-bound element-wise or per-channel-reduction passes over NHWC fp32:
// thread = (pixel or 2x2 window) x 4 channels, float4 accesses, fixed-order reductions.
#pragma once
#include "common.cuh"
#include "dp_dev.cuh"

namespace s2s {

// ------------------------------------------------------------------ inference-mode BN folding
struct BnFoldEntry { int64_t gamma_off, beta_off, mm_off, mv_off, out_off; int C; };

__global__ void bn_fold_kernel(const BnFoldEntry* __restrict__ tab, const float* __restrict__ params,
                               const float* __restrict__ state, float* __restrict__ scale,
                               float* __restrict__ shift, float eps) {
    const BnFoldEntry e = tab[blockIdx.x];
    for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
        const float sc = params[e.gamma_off + c] * rsqrtf(state[e.mv_off + c] + eps);
        scale[e.out_off + c] = sc;
        shift[e.out_off + c] = params[e.beta_off + c] - state[e.mm_off + c] * sc;
    }
}

// ------------------------------------------------------------------ forward apply (+ pool)
// Training mode: every CTA first finalises the batch statistics from the conv epilogue's per-CTA
// partials (fixed order, identical in every CTA -> deterministic, no election / fence in the producer);
// CTA 0 publishes mean / rstd / scale / shift (needed by the backward) and updates the moving statistics.
struct BnApplyArgs {
    const float* a;                 // ELU output, dense [N,h,w,C]
    const float* scale; const float* shift;   // [C] (inference: folded moving statistics)
    float* c_out; int ldc, coffc;   // BN output (skip half of the concat buffer, or dense)
    float* p_out;                   // pooled BN output dense [N,h/2,w/2,C] (POOLED only)
    int pool_kind, N, h, w, C;
    // training-mode finalize (stat_part != null)
    const float* stat_part; int nslots;       // [nslots][2][C]
    const float* gamma; const float* beta;
    float* mov_mean; float* mov_var;
    float* bn_mean; float* bn_rstd; float* bn_scale; float* bn_shift;
    float eps, momentum; int update_moving;
    double M_total;                 // > 0: element count of the GLOBAL batch (sync-BN, dp.cuh); 0: N*h*w
    int sync_id;                    // >= 0: exchange the per-channel sums with the peer ranks inside this kernel
    int early_loads;                // issue the first item's activation loads before the statistics are finalised
    DpDev dp;
};

constexpr int BN_MAXC = 512;        // filters*4*2^n_blocks <= 4*4*32

template <bool POOLED>
__global__ void __launch_bounds__(256) bn_apply_kernel(const BnApplyArgs a) {
    __shared__ float s_scale[BN_MAXC], s_shift[BN_MAXC];
    __shared__ double sd_tmp[1024], sd_out[2 * BN_MAXC];
    const int tid = threadIdx.x;
    pdl_wait();
    // The activation loads of the thread's first item do not depend on the statistics: they are issued BEFORE the finalize, so
    // their global round trip overlaps the one of the partials (one dependent round trip less per BatchNorm on the chain).
    const int CQ = a.C >> 2;
    const int h2 = a.h >> 1, w2 = a.w >> 1;
    const int64_t total = POOLED ? (int64_t)a.N * h2 * w2 * CQ : (int64_t)a.N * a.h * a.w * CQ;
    const int64_t idx0 = (int64_t)blockIdx.x * 256 + tid;
    constexpr int NPX = POOLED ? 4 : 1;
    float4 pre[NPX];
    auto pixel_of = [&](int64_t idx, int i) -> size_t {
        if (POOLED) {
            const int64_t win = idx / CQ;
            const int px = (int)(win % w2), py = (int)((win / w2) % h2), n = (int)(win / ((int64_t)w2 * h2));
            return ((size_t)n * a.h + 2 * py + (i >> 1)) * a.w + 2 * px + (i & 1);
        }
        return (size_t)(idx / CQ);
    };
    const bool early = a.early_loads != 0 && idx0 < total;
    if (early) {
        const int cq = (int)(idx0 % CQ);
#pragma unroll
        for (int i = 0; i < NPX; ++i) pre[i] = ld4(a.a + pixel_of(idx0, i) * a.C + 4 * cq);
    }
    if (a.stat_part) {
        const double M = a.M_total > 0.0 ? a.M_total : (double)a.N * a.h * a.w;
        // gamma / beta of the (at most two) channels this thread finalises: parameters, fetched before the partials are reduced
        float gam[2] = {0.f, 0.f}, bet[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (tid + 256 * k < a.C) { gam[k] = a.gamma[tid + 256 * k]; bet[k] = a.beta[tid + 256 * k]; }
        auto finalize = [&](int c, double sum, double sumsq) {
            const double mean = sum / M;
            double var = sumsq / M - mean * mean;
            if (var < 0.0) var = 0.0;
            const float meanf = (float)mean, varf = (float)var;
            const float rstd = rsqrtf(varf + a.eps);
            const float sc = gam[c >> 8] * rstd;
            const float sh = bet[c >> 8] - meanf * sc;
            s_scale[c] = sc;
            s_shift[c] = sh;
            if (blockIdx.x == 0) {
                a.bn_mean[c] = meanf; a.bn_rstd[c] = rstd; a.bn_scale[c] = sc; a.bn_shift[c] = sh;
                if (a.update_moving) {
                    a.mov_mean[c] = a.mov_mean[c] * a.momentum + meanf * (1.f - a.momentum);
                    a.mov_var[c] = a.mov_var[c] * a.momentum + varf * (1.f - a.momentum);
                }
            }
        };
        // A: this rank's per-channel (sum, sumsq) -> sd_out[0..C), sd_out[C..2C)
        if (cta_reduce_block_ok(2 * a.C)) {        // [nslots][2][C] is one contiguous block: single pass (common.cuh)
            cta_reduce_block256(a.stat_part, a.nslots, 2 * a.C, sd_tmp, sd_out, tid);
        } else {
            __shared__ double sd_chunk[256];
            for (int c0 = 0; c0 < a.C; c0 += 128) {
                const int nc = min(128, a.C - c0);
                cta_reduce_slots<256>(a.stat_part + c0, a.nslots, (size_t)2 * a.C, nc, sd_tmp, sd_chunk, tid);
                cta_reduce_slots<256>(a.stat_part + a.C + c0, a.nslots, (size_t)2 * a.C, nc, sd_tmp, sd_chunk + nc, tid);
                if (tid < nc) { sd_out[c0 + tid] = sd_chunk[tid]; sd_out[a.C + c0 + tid] = sd_chunk[nc + tid]; }
                __syncthreads();
            }
        }
        // B: sync-BN — add the peers' sums over NVLink peer memory (dp_dev.cuh)
        if (a.sync_id >= 0) dp_exchange_sums(a.dp, a.sync_id, sd_out, 2 * a.C, tid, 256);
        // C: finalize
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (tid + 256 * k < a.C) finalize(tid + 256 * k, sd_out[tid + 256 * k], sd_out[a.C + tid + 256 * k]);
        __syncthreads();
    } else {
        for (int c = tid; c < a.C; c += 256) { s_scale[c] = a.scale[c]; s_shift[c] = a.shift[c]; }
        __syncthreads();
    }
    pdl_trigger();          // the next kernel's prologue overlaps the apply loop
    for (int64_t idx = idx0; idx < total; idx += (int64_t)gridDim.x * 256) {
        const int cq = (int)(idx % CQ);
        const float4 sc = ld4(s_scale + 4 * cq), sh = ld4(s_shift + 4 * cq);
        float4 y[NPX];
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const size_t pix = pixel_of(idx, i);
            const float4 v = (early && idx == idx0) ? pre[i] : ld4(a.a + pix * a.C + 4 * cq);
            y[i] = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
            st4(a.c_out + pix * a.ldc + a.coffc + 4 * cq, y[i]);
        }
        if (POOLED) {
            float4 p;
            if (a.pool_kind == S2S_POOL_AVG) {
                p.x = ((y[0].x + y[NPX > 1 ? 1 : 0].x) + (y[NPX > 2 ? 2 : 0].x + y[NPX > 3 ? 3 : 0].x)) * 0.25f;
                p.y = ((y[0].y + y[NPX > 1 ? 1 : 0].y) + (y[NPX > 2 ? 2 : 0].y + y[NPX > 3 ? 3 : 0].y)) * 0.25f;
                p.z = ((y[0].z + y[NPX > 1 ? 1 : 0].z) + (y[NPX > 2 ? 2 : 0].z + y[NPX > 3 ? 3 : 0].z)) * 0.25f;
                p.w = ((y[0].w + y[NPX > 1 ? 1 : 0].w) + (y[NPX > 2 ? 2 : 0].w + y[NPX > 3 ? 3 : 0].w)) * 0.25f;
            } else {
                p.x = fmaxf(fmaxf(y[0].x, y[NPX > 1 ? 1 : 0].x), fmaxf(y[NPX > 2 ? 2 : 0].x, y[NPX > 3 ? 3 : 0].x));
                p.y = fmaxf(fmaxf(y[0].y, y[NPX > 1 ? 1 : 0].y), fmaxf(y[NPX > 2 ? 2 : 0].y, y[NPX > 3 ? 3 : 0].y));
                p.z = fmaxf(fmaxf(y[0].z, y[NPX > 1 ? 1 : 0].z), fmaxf(y[NPX > 2 ? 2 : 0].z, y[NPX > 3 ? 3 : 0].z));
                p.w = fmaxf(fmaxf(y[0].w, y[NPX > 1 ? 1 : 0].w), fmaxf(y[NPX > 2 ? 2 : 0].w, y[NPX > 3 ? 3 : 0].w));
            }
            st4(a.p_out + (size_t)(idx / CQ) * a.C + 4 * cq, p);
        }
    }
}

static inline int bn_apply(const BnApplyArgs& a, bool pooled, cudaStream_t st) {
    S2S_REQUIRE((a.C & 3) == 0 && (a.ldc & 3) == 0 && (a.coffc & 3) == 0, "bn_apply: C must be a multiple of 4");
    S2S_REQUIRE(a.C <= BN_MAXC, "bn_apply: C=%d exceeds %d", a.C, BN_MAXC);
    prof_begin(st, pooled ? "bn_apply_pool" : "bn_apply", 4.0 * a.N * a.h * a.w * a.C * (pooled ? 2.25 : 2.0), 0.0);
    if (pooled) S2S_REQUIRE((a.h & 1) == 0 && (a.w & 1) == 0, "pooling needs even H and W (got %dx%d)", a.h, a.w);
    const int64_t total = pooled ? (int64_t)a.N * (a.h / 2) * (a.w / 2) * (a.C / 4) : (int64_t)a.N * a.h * a.w * (a.C / 4);
    // each CTA repeats the statistics finalize: keep the grid at <= 2 CTAs per SM and grid-stride the pixels
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv64(total, 256), 2 * 148);
    if (pooled) launch_k(bn_apply_kernel<true>, grid, 256, 0, st, a);
    else launch_k(bn_apply_kernel<false>, grid, 256, 0, st, a);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ backward
// dc   = g1 (full-res gradient source, channel slice of a wider buffer)  +  unpool(g2)
// R1   : s1 = sum dc, s2 = sum dc * xhat          (xhat = (act - mean) * rstd)  -> dbeta, dgamma, m1, m2
// R2   : dz = scale * (dc - m1 - xhat*m2) * ELU'(act)         (training-mode BN)
//        dz = scale * dc * ELU'(act)                          (inference-mode BN / no BN: m1=m2 unused)
struct BnBwdArgs {
    const float* act;                       // ELU output (pre-BN), dense [N,h,w,C]
    const float* g1; int ld1, coff1;        // nullable
    const float* g2; int pool_kind;         // nullable, dense [N,h/2,w/2,C]
    const float* mean; const float* rstd; const float* scale; const float* shift;
    float* part; int nslots;                // [nslots][2][C] partials of (sum dc, sum dc*xhat), written by bn_bwd_reduce or by the
                                            // epilogue of the kernel that produced g1 (gconv.cuh, stat_aux)
    float* dbeta; float* dgamma;            // [C] in the dense gradient arena: this rank's (sum dc, sum dc*xhat), written by CTA 0 of the
                                            // apply kernel from the partials it finalises anyway (nullable)
    float* dz;                              // dense [N,h,w,C]
    int N, h, w, C, batch_stats, apply_elugrad, act_kind;
    double M_total;                         // > 0: element count of the global batch (sync-BN)
    int sync_id;                            // >= 0: exchange (sum dc, sum dc*xhat) with the peer ranks inside bn_bwd_apply
    int early_loads;                        // bn_bwd_apply: load the thread's first unit before the statistics are finalised
    DpDev dp;
};

// gradient wrt the BN output for the 4 pixels of a 2x2 window (POOLED) or 1 pixel, 4 channels
template <bool POOLED>
struct BnUnit {
    static constexpr int NP = POOLED ? 4 : 1;
    float4 a[NP];
    float4 dc[NP];
    size_t pix[NP];
    __device__ __forceinline__ void load(const BnBwdArgs& g, int64_t unit, int cq) {
        if (POOLED) {
            const int h2 = g.h >> 1, w2 = g.w >> 1;
            const int px = (int)(unit % w2), py = (int)((unit / w2) % h2), n = (int)(unit / ((int64_t)w2 * h2));
#pragma unroll
            for (int i = 0; i < NP; ++i) pix[i] = ((size_t)n * g.h + 2 * py + (i >> 1)) * g.w + 2 * px + (i & 1);
        } else {
            pix[0] = (size_t)unit;
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            a[i] = ld4(g.act + pix[i] * g.C + 4 * cq);
            dc[i] = g.g1 ? ld4(g.g1 + pix[i] * g.ld1 + g.coff1 + 4 * cq) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (POOLED && g.g2) {
            const float4 gp = ld4(g.g2 + (size_t)unit * g.C + 4 * cq);
            if (g.pool_kind == S2S_POOL_AVG) {
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    dc[i].x = fmaf(gp.x, 0.25f, dc[i].x); dc[i].y = fmaf(gp.y, 0.25f, dc[i].y);
                    dc[i].z = fmaf(gp.z, 0.25f, dc[i].z); dc[i].w = fmaf(gp.w, 0.25f, dc[i].w);
                }
            } else {
                // route to the first maximum (row-major window order) of the BN output
                const float4 sc = ld4(g.scale + 4 * cq), sh = ld4(g.shift + 4 * cq);
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
                const float gpv[4] = {gp.x, gp.y, gp.z, gp.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int best = 0;
                    float bv = fmaf(reinterpret_cast<const float*>(&a[0])[e], scv[e], shv[e]);
#pragma unroll
                    for (int i = 1; i < NP; ++i) {
                        const float v = fmaf(reinterpret_cast<const float*>(&a[i])[e], scv[e], shv[e]);
                        if (v > bv) { bv = v; best = i; }
                    }
#pragma unroll
                    for (int i = 0; i < NP; ++i)
                        if (i == best) reinterpret_cast<float*>(&dc[i])[e] += gpv[e];
                }
            }
        }
    }
};

template <bool POOLED>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwdArgs g, int64_t units) {
    __shared__ __align__(16) float sred[256 * 8];
    const int CQ = g.C >> 2;
    const int cq = blockIdx.y * blockDim.x + threadIdx.x;
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0.f;
    pdl_wait();
    if (cq < CQ) {
        const float4 mu = ld4(g.mean + 4 * cq), rs = ld4(g.rstd + 4 * cq);
        for (int64_t u = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; u < units; u += (int64_t)gridDim.x * blockDim.y) {
            BnUnit<POOLED> un;
            un.load(g, u, cq);
#pragma unroll
            for (int i = 0; i < BnUnit<POOLED>::NP; ++i) {
                s[0] += un.dc[i].x; s[1] += un.dc[i].y; s[2] += un.dc[i].z; s[3] += un.dc[i].w;
                s[4] = fmaf(un.dc[i].x, (un.a[i].x - mu.x) * rs.x, s[4]);
                s[5] = fmaf(un.dc[i].y, (un.a[i].y - mu.y) * rs.y, s[5]);
                s[6] = fmaf(un.dc[i].z, (un.a[i].z - mu.z) * rs.z, s[6]);
                s[7] = fmaf(un.dc[i].w, (un.a[i].w - mu.w) * rs.w, s[7]);
            }
        }
    }
    pdl_trigger();
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; ++k) sred[t * 8 + k] = s[k];
    __syncthreads();
    if (t < (int)blockDim.x * 8) {
        const int tx = t >> 3, k = t & 7;
        float acc = 0.f;
        for (int ty = 0; ty < (int)blockDim.y; ++ty) acc += sred[(ty * blockDim.x + tx) * 8 + k];
        const int c = 4 * (blockIdx.y * blockDim.x + tx) + (k & 3);
        if (c < g.C) g.part[((size_t)blockIdx.x * 2 + (k >> 2)) * g.C + c] = acc;
    }
}

template <bool POOLED>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdArgs g, int64_t units) {
    __shared__ __align__(16) float s_m1[BN_MAXC], s_m2[BN_MAXC];
    __shared__ double sd_tmp[1024], sd_out[2 * BN_MAXC];
    const int tid = threadIdx.x;
    pdl_wait();
    // the thread's first unit is loaded BEFORE the statistics are finalised (independent global round trips overlap)
    const int CQ = g.C >> 2;
    const int64_t idx0 = (int64_t)blockIdx.x * 256 + tid;
    BnUnit<POOLED> first;
    const bool early = g.early_loads != 0 && idx0 < units * CQ;
    float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), mu0 = make_float4(0.f, 0.f, 0.f, 0.f), rs0 = mu0;
    if (early) {
        const int cq0 = (int)(idx0 % CQ);
        first.load(g, idx0 / CQ, cq0);
        if (g.scale) sc0 = ld4(g.scale + 4 * cq0);
        if (g.batch_stats) { mu0 = ld4(g.mean + 4 * cq0); rs0 = ld4(g.rstd + 4 * cq0); }
    }
    if (g.batch_stats) {   // finalise (sum dc, sum dc*xhat) / M from the reduce kernel's partials, in every CTA
        const double M = g.M_total > 0.0 ? g.M_total : (double)g.N * g.h * g.w;
        if (cta_reduce_block_ok(2 * g.C)) {
            cta_reduce_block256(g.part, g.nslots, 2 * g.C, sd_tmp, sd_out, tid);
        } else {
            __shared__ double sd_chunk[256];
            for (int c0 = 0; c0 < g.C; c0 += 128) {
                const int nc = min(128, g.C - c0);
                cta_reduce_slots<256>(g.part + c0, g.nslots, (size_t)2 * g.C, nc, sd_tmp, sd_chunk, tid);
                cta_reduce_slots<256>(g.part + g.C + c0, g.nslots, (size_t)2 * g.C, nc, sd_tmp, sd_chunk + nc, tid);
                if (tid < nc) { sd_out[c0 + tid] = sd_chunk[tid]; sd_out[g.C + c0 + tid] = sd_chunk[nc + tid]; }
                __syncthreads();
            }
        }
        // d beta / d gamma of THIS rank (the data-parallel exchange of the gradient arena sums the ranks): before the sync-BN exchange
        if (blockIdx.x == 0 && g.dbeta != nullptr)
            for (int c = tid; c < g.C; c += 256) { g.dbeta[c] = (float)sd_out[c]; g.dgamma[c] = (float)sd_out[g.C + c]; }
        if (g.sync_id >= 0) dp_exchange_sums(g.dp, g.sync_id, sd_out, 2 * g.C, tid, 256);
        for (int c = tid; c < g.C; c += 256) { s_m1[c] = (float)(sd_out[c] / M); s_m2[c] = (float)(sd_out[g.C + c] / M); }
        __syncthreads();
    }
    pdl_trigger();
    for (int64_t idx = idx0; idx < units * CQ; idx += (int64_t)gridDim.x * 256) {
    const int cq = (int)(idx % CQ);
    const int64_t u = idx / CQ;
    BnUnit<POOLED> un;
    const bool pre = early && idx == idx0;
    if (pre) un = first; else un.load(g, u, cq);
    const float4 sc = pre ? sc0 : (g.scale ? ld4(g.scale + 4 * cq) : make_float4(1.f, 1.f, 1.f, 1.f));
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu, m1 = mu, m2 = mu;
    if (g.batch_stats) {
        if (pre) { mu = mu0; rs = rs0; } else { mu = ld4(g.mean + 4 * cq); rs = ld4(g.rstd + 4 * cq); }
        m1 = ld4(s_m1 + 4 * cq); m2 = ld4(s_m2 + 4 * cq);
    }
#pragma unroll
    for (int i = 0; i < BnUnit<POOLED>::NP; ++i) {
        const float4 a = un.a[i], dc = un.dc[i];
        float4 r;
        r.x = sc.x * (dc.x - m1.x - (a.x - mu.x) * rs.x * m2.x);
        r.y = sc.y * (dc.y - m1.y - (a.y - mu.y) * rs.y * m2.y);
        r.z = sc.z * (dc.z - m1.z - (a.z - mu.z) * rs.z * m2.z);
        r.w = sc.w * (dc.w - m1.w - (a.w - mu.w) * rs.w * m2.w);
        if (g.apply_elugrad) {
            r.x *= act_grad_from_out(a.x, g.act_kind); r.y *= act_grad_from_out(a.y, g.act_kind);
            r.z *= act_grad_from_out(a.z, g.act_kind); r.w *= act_grad_from_out(a.w, g.act_kind);
        }
        st4(g.dz + un.pix[i] * g.C + 4 * cq, r);
    }
    }
}

// ------------------------------------------------------------------ fused backward (latency regime)
// bn_bwd_reduce + bn_bwd_apply in ONE launch for grids that are co-resident on the GPU (the reference's batch sizes):
// every thread keeps the (activation, gradient) values of its first item in registers, the CTAs publish their (sum dc,
// sum dc*xhat) partials, meet at a software grid barrier, then every CTA finalises the sums in the same fixed order and
// applies them to the values it still holds — the second pass over HBM/L2 and one kernel boundary of the critical path go
// away (6 launches per step).  The partial layout [slot][2][C] is unchanged (= d beta / d gamma partials for the fused Adam).
// Launched cooperatively (co-residency guaranteed by the driver) with the grid capped at half of what fits the device, so
// that the weight-gradient kernels of the side streams keep running beside it.
struct GridBarrier { unsigned int arrive, leave; };

__device__ __forceinline__ void grid_barrier_once(GridBarrier* gb, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        atomicAdd(&gb->arrive, 1u);
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned int*>(&gb->arrive) < nblocks) {
            if (++spins > (1u << 23)) __trap();          // a placement bug must not hang the GPU
        }
        __threadfence();
        if (atomicAdd(&gb->leave, 1u) == nblocks - 1u) { gb->arrive = 0u; gb->leave = 0u; }   // the last one out resets for the next launch
    }
    __syncthreads();
}

template <bool POOLED>
__global__ void __launch_bounds__(256) bn_bwd_fused_kernel(const BnBwdArgs g, int64_t units, GridBarrier* gb) {
    __shared__ __align__(16) float sred[256 * 8];
    __shared__ __align__(16) float s_m1[BN_MAXC], s_m2[BN_MAXC];
    __shared__ double sd_tmp[1024], sd_out[2 * BN_MAXC];
    const int CQ = g.C >> 2;
    const int cq = blockIdx.y * blockDim.x + threadIdx.x;
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    const int64_t u0 = (int64_t)blockIdx.x * blockDim.y + threadIdx.y, ustep = (int64_t)gridDim.x * blockDim.y;
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0.f;
    BnUnit<POOLED> keep;                     // the thread's first item stays in registers across the barrier
    const bool have = cq < CQ && u0 < units;
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu;
    if (cq < CQ) { mu = ld4(g.mean + 4 * cq); rs = ld4(g.rstd + 4 * cq); }
    auto accumulate = [&](const BnUnit<POOLED>& un) {
#pragma unroll
        for (int i = 0; i < BnUnit<POOLED>::NP; ++i) {
            s[0] += un.dc[i].x; s[1] += un.dc[i].y; s[2] += un.dc[i].z; s[3] += un.dc[i].w;
            s[4] = fmaf(un.dc[i].x, (un.a[i].x - mu.x) * rs.x, s[4]);
            s[5] = fmaf(un.dc[i].y, (un.a[i].y - mu.y) * rs.y, s[5]);
            s[6] = fmaf(un.dc[i].z, (un.a[i].z - mu.z) * rs.z, s[6]);
            s[7] = fmaf(un.dc[i].w, (un.a[i].w - mu.w) * rs.w, s[7]);
        }
    };
    if (have) {
        keep.load(g, u0, cq);
        accumulate(keep);
        for (int64_t u = u0 + ustep; u < units; u += ustep) {
            BnUnit<POOLED> un;
            un.load(g, u, cq);
            accumulate(un);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sred[t * 8 + k] = s[k];
    __syncthreads();
    if (t < (int)blockDim.x * 8) {
        const int tx = t >> 3, k = t & 7;
        float acc = 0.f;
        for (int ty = 0; ty < (int)blockDim.y; ++ty) acc += sred[(ty * blockDim.x + tx) * 8 + k];
        const int c = 4 * (blockIdx.y * blockDim.x + tx) + (k & 3);
        if (c < g.C) g.part[((size_t)blockIdx.x * 2 + (k >> 2)) * g.C + c] = acc;
    }
    grid_barrier_once(gb, gridDim.x * gridDim.y);
    // finalise (sum dc, sum dc*xhat) / M in every CTA, same order as bn_bwd_apply_kernel
    {
        const double M = g.M_total > 0.0 ? g.M_total : (double)g.N * g.h * g.w;
        if (cta_reduce_block_ok(2 * g.C)) {
            cta_reduce_block256(g.part, g.nslots, 2 * g.C, sd_tmp, sd_out, t);
        } else {
            __shared__ double sd_chunk[256];
            for (int c0 = 0; c0 < g.C; c0 += 128) {
                const int nc = min(128, g.C - c0);
                cta_reduce_slots<256>(g.part + c0, g.nslots, (size_t)2 * g.C, nc, sd_tmp, sd_chunk, t);
                cta_reduce_slots<256>(g.part + g.C + c0, g.nslots, (size_t)2 * g.C, nc, sd_tmp, sd_chunk + nc, t);
                if (t < nc) { sd_out[c0 + t] = sd_chunk[t]; sd_out[g.C + c0 + t] = sd_chunk[nc + t]; }
                __syncthreads();
            }
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && g.dbeta != nullptr)
            for (int c = t; c < g.C; c += 256) { g.dbeta[c] = (float)sd_out[c]; g.dgamma[c] = (float)sd_out[g.C + c]; }
        for (int c = t; c < g.C; c += 256) { s_m1[c] = (float)(sd_out[c] / M); s_m2[c] = (float)(sd_out[g.C + c] / M); }
        __syncthreads();
    }
    if (!have) return;
    const float4 sc = g.scale ? ld4(g.scale + 4 * cq) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 m1 = ld4(s_m1 + 4 * cq), m2 = ld4(s_m2 + 4 * cq);
    auto apply = [&](const BnUnit<POOLED>& un) {
#pragma unroll
        for (int i = 0; i < BnUnit<POOLED>::NP; ++i) {
            const float4 a = un.a[i], dc = un.dc[i];
            float4 r;
            r.x = sc.x * (dc.x - m1.x - (a.x - mu.x) * rs.x * m2.x);
            r.y = sc.y * (dc.y - m1.y - (a.y - mu.y) * rs.y * m2.y);
            r.z = sc.z * (dc.z - m1.z - (a.z - mu.z) * rs.z * m2.z);
            r.w = sc.w * (dc.w - m1.w - (a.w - mu.w) * rs.w * m2.w);
            if (g.apply_elugrad) {
                r.x *= act_grad_from_out(a.x, g.act_kind); r.y *= act_grad_from_out(a.y, g.act_kind);
                r.z *= act_grad_from_out(a.z, g.act_kind); r.w *= act_grad_from_out(a.w, g.act_kind);
            }
            st4(g.dz + un.pix[i] * g.C + 4 * cq, r);
        }
    };
    apply(keep);
    for (int64_t u = u0 + ustep; u < units; u += ustep) {
        BnUnit<POOLED> un;
        un.load(g, u, cq);
        apply(un);
    }
}

// nslots the reduce kernel will use (sizes the workspace)
static inline int bn_bwd_slots(int64_t units, int py) {
    // 64 CTAs at the reference's batch sizes (latency-tuned); grows to 4 CTAs per SM for large batches, where 64 CTAs
    // left the reduction at ~1 TB/s (profiles/r1_summary.md)
    static const int lo_cap = [] { const char* e = getenv("S2S_BN_SLOTS_LO"); return e ? atoi(e) : 64; }();
    int64_t lo = cdiv64(units, py);
    if (lo > lo_cap) lo = lo_cap;
    int64_t s = cdiv64(units, 4 * (int64_t)py);
    if (s < lo) s = lo;
    if (s > 4 * 148) s = 4 * 148;
    return (int)s;
}
static inline int bn_cqb(int C) {
    const int cq = C / 4;
    int b = 1;
    while (b < cq && b < 32) b <<= 1;
    return b;
}

static inline int bn_bwd_reduce(const BnBwdArgs& g, cudaStream_t st) {
    S2S_REQUIRE((g.C & 3) == 0 && g.C <= BN_MAXC, "bn_bwd: C must be a multiple of 4 and <= %d", BN_MAXC);
    const bool pooled = g.g2 != nullptr;
    const int64_t units = pooled ? (int64_t)g.N * (g.h / 2) * (g.w / 2) : (int64_t)g.N * g.h * g.w;
    const int cqb = bn_cqb(g.C);
    const int py = 256 / cqb;
    dim3 block(cqb, py);
    dim3 grid(g.nslots, cdiv(g.C / 4, cqb));      // nslots is fixed at handle creation; idle slots write zeros
    prof_begin(st, "bn_bwd_reduce", 4.0 * g.N * g.h * g.w * g.C * ((g.g1 ? 2.0 : 1.0) + (pooled ? 0.25 : 0.0)), 0.0);
    if (pooled) launch_k(bn_bwd_reduce_kernel<true>, grid, block, 0, st, g, units);
    else launch_k(bn_bwd_reduce_kernel<false>, grid, block, 0, st, g, units);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

static inline int bn_bwd_apply(const BnBwdArgs& g, cudaStream_t st) {
    const bool pooled = g.g2 != nullptr;
    const int64_t units = pooled ? (int64_t)g.N * (g.h / 2) * (g.w / 2) : (int64_t)g.N * g.h * g.w;
    const int64_t total = units * (g.C / 4);
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv64(total, 256), 2 * 148);
    prof_begin(st, "bn_bwd_apply", 4.0 * g.N * g.h * g.w * g.C * ((g.g1 ? 3.0 : 2.0) + (pooled ? 0.25 : 0.0)), 0.0);
    if (pooled) launch_k(bn_bwd_apply_kernel<true>, grid, 256, 0, st, g, units);
    else launch_k(bn_bwd_apply_kernel<false>, grid, 256, 0, st, g, units);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

// One-launch variant: usable when training-mode statistics are needed, no sync-BN exchange is pending and the whole grid
// is co-resident (checked against the occupancy of this kernel on the current device).
static inline bool bn_bwd_fused_ok(const BnBwdArgs& g) {
    // measured on B200 (batch 16): 430-436 us per step fused vs 429 us with the two kernels: the fused grid (64 reduction
    // slots) applies with fewer CTAs than bn_bwd_apply's 296 and pays the barrier; kept as an opt-in experiment
    static const bool on = [] { const char* e = getenv("S2S_BN_FUSE"); return e && e[0] == '1'; }();
    if (!on || !g.batch_stats || g.sync_id >= 0) return false;
    const int cqb = bn_cqb(g.C);
    const int gy = cdiv(g.C / 4, cqb);
    static int cap_pooled = -1, cap_plain = -1;
    int& cap = g.g2 ? cap_pooled : cap_plain;
    if (cap < 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (g.g2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_fused_kernel<true>, 256, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_fused_kernel<false>, 256, 0);
        cap = sms * per_sm / 2;          // half of the device: the weight-gradient kernels of the side streams keep running
    }
    return (int64_t)g.nslots * gy <= cap;
}
static inline int bn_bwd_fused(const BnBwdArgs& g, GridBarrier* gb, cudaStream_t st) {
    S2S_REQUIRE((g.C & 3) == 0 && g.C <= BN_MAXC, "bn_bwd: C must be a multiple of 4 and <= %d", BN_MAXC);
    const bool pooled = g.g2 != nullptr;
    const int64_t units = pooled ? (int64_t)g.N * (g.h / 2) * (g.w / 2) : (int64_t)g.N * g.h * g.w;
    const int cqb = bn_cqb(g.C);
    dim3 block(cqb, 256 / cqb);
    dim3 grid(g.nslots, cdiv(g.C / 4, cqb));
    prof_begin(st, "bn_bwd_fused", 4.0 * g.N * g.h * g.w * g.C * ((g.g1 ? 3.0 : 2.0) + (pooled ? 0.25 : 0.0)), 0.0);
    // cooperative launch: the driver guarantees (or refuses) co-residency of the whole grid, also inside a captured graph
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    static const bool coop = [] { const char* e = getenv("S2S_BN_FUSE_COOP"); return !e || e[0] != '0'; }();
    cfg.attrs = at; cfg.numAttrs = coop ? 1 : 0;
    if (pooled) cudaLaunchKernelEx(&cfg, bn_bwd_fused_kernel<true>, g, units, gb);
    else cudaLaunchKernelEx(&cfg, bn_bwd_fused_kernel<false>, g, units, gb);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ per-channel sum of a channel slice
// out[c] = sum over pixels of g[pix*ld + coff + c]      (Conv2DTranspose bias gradient)
struct ChanSumArgs { const float* g; int ld, coff, C; int64_t npix; float* part; int nslots; };   // part [nslots][C]

__global__ void __launch_bounds__(256) chansum_kernel(const ChanSumArgs g) {
    __shared__ __align__(16) float sred[256 * 4];
    const int CQ = g.C >> 2;
    const int cq = blockIdx.y * blockDim.x + threadIdx.x;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (cq < CQ) {
        const int64_t step = (int64_t)gridDim.x * blockDim.y;
        int64_t p = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
        for (; p + 3 * step < g.npix; p += 4 * step) {      // 4 loads in flight, summed in pixel order
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld4(g.g + (size_t)(p + u * step) * g.ld + g.coff + 4 * cq);
#pragma unroll
            for (int u = 0; u < 4; ++u) { s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w; }
        }
        for (; p < g.npix; p += step) {
            const float4 v = ld4(g.g + (size_t)p * g.ld + g.coff + 4 * cq);
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        }
    }
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) sred[t * 4 + k] = s[k];
    __syncthreads();
    if (t < (int)blockDim.x * 4) {
        const int tx = t >> 2, k = t & 3;
        float acc = 0.f;
        for (int ty = 0; ty < (int)blockDim.y; ++ty) acc += sred[(ty * blockDim.x + tx) * 4 + k];
        const int c = 4 * (blockIdx.y * blockDim.x + tx) + k;
        if (c < g.C) g.part[(size_t)blockIdx.x * g.C + c] = acc;
    }
}

static inline int chansum(const ChanSumArgs& g, cudaStream_t st) {
    S2S_REQUIRE((g.C & 3) == 0 && (g.ld & 3) == 0 && (g.coff & 3) == 0, "chansum: C must be a multiple of 4");
    const int cqb = bn_cqb(g.C);
    const int py = 256 / cqb;
    dim3 block(cqb, py);
    dim3 grid(g.nslots, cdiv(g.C / 4, cqb));
    prof_begin(st, "chansum", 4.0 * g.npix * g.C, 0.0);
    chansum_kernel<<<grid, block, 0, st>>>(g);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2s
