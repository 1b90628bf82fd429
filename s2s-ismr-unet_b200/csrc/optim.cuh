// optim.cuh — single-pass Keras-form Adam over the flat parameter arena, fused with the
// fixed-order reduction of the per-layer gradient partials.
//
// Keras-3 Adam (optimizers.Adam(learning_rate=lr), training.py:66,95):
//   t = iterations + 1;  alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)
//   m += (g - m)(1 - b1);  v += (g^2 - v)(1 - b2);  w -= alpha * m / (sqrt(v) + eps)      eps = 1e-7
// HBM-bound: reads g (or its partials), m, v, w and writes m, v, w once: 7*P*4 bytes per step.
#pragma once
#include "common.cuh"

namespace s2s {

struct AdamHyper {          // lives in device memory so that a captured graph sees updates
    double lr, beta1, beta2;
    float eps, omb1, omb2;  // eps, 1-beta1, 1-beta2 rounded once from the double values (as Keras does)
    float alpha;            // lr * sqrt(1 - b2^t) / (1 - b1^t) for the current step (set when step is bumped)
    long long step;         // number of completed + current optimiser steps (1-based when Adam runs)
};

// bias-corrected step size in double (Keras evaluates it in fp32; see DESIGN.md "Adam")
__host__ __device__ inline float adam_alpha(double lr, double b1, double b2, long long step) {
    const double t = (double)step;
    return (float)(lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t)));
}
inline AdamHyper make_hyper(double lr, double b1, double b2, double eps, long long step) {
    AdamHyper h;
    h.lr = lr; h.beta1 = b1; h.beta2 = b2;
    h.eps = (float)eps; h.omb1 = (float)(1.0 - b1); h.omb2 = (float)(1.0 - b2);
    h.alpha = step > 0 ? adam_alpha(lr, b1, b2, step) : 0.f;
    h.step = step;
    return h;
}
// bump the step counter and refresh alpha (called by exactly one thread per optimiser step)
__device__ inline void adam_bump(AdamHyper* hy) {
    hy->step += 1;
    hy->alpha = adam_alpha(hy->lr, hy->beta1, hy->beta2, hy->step);
}

// One entry per 32-element block of the parameter arena (GRAD_BLK consecutive parameters = one coalesced row).
constexpr int GRAD_BLK = 32;
struct GradBlock {
    int64_t param_off;      // first parameter element of this block
    int32_t count;          // <= GRAD_BLK
    int32_t nslots;         // 0: the dense gradient is already in grads[]; >0: sum nslots partials
    int64_t part_off;       // offset of partial slot 0 for element param_off
    int64_t part_stride;    // elements between slots
};

// A CTA (8 warps) handles `nblk` consecutive blocks with W = 8 / 4 / 2 / 1 warps each (nblk * W <= 8): W grows with the
// number of partial slots so that no thread issues more than ~32 loads (thin layers: 256 slots -> 8 warps; deep layers:
// 8-16 slots -> 1 warp, 8 blocks per CTA).  The step's whole reduction + Adam is then ~640 CTAs = one wave instead of
// 4218 CTAs of which 7 warps idled through the Adam update (ncu: 16 us at the tail of the critical path).
struct GradCta { int32_t first, nblk, W, pad; };
static inline int grad_block_warps(int nslots) { return nslots > 64 ? 8 : nslots > 32 ? 4 : nslots > 16 ? 2 : 1; }

// lane = parameter; warp (bi, wi) sums slots wi, wi+W, ... of block bi (8 loads in flight), the W warp sums are added
// in warp order by warp wi == 0, which then applies Adam.  Fixed order -> bit-reproducible.
__device__ __forceinline__ bool grad_block_reduce(const GradCta c, const GradBlock* __restrict__ blocks, const float* __restrict__ part,
                                                  float (*sred)[GRAD_BLK], GradBlock& b, float& g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = warp / c.W, wi = warp - bi * c.W;
    const bool have = bi < c.nblk;
    if (have) b = blocks[c.first + bi];
    const bool live = have && lane < b.count;
    float s = 0.f;
    if (live && b.nslots > 0) {
        const float* src = part + b.part_off + lane;
        const int W = c.W;
        int sl = wi;
        for (; sl + 7 * W < b.nslots; sl += 8 * W) {
            float t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldcg(src + (int64_t)(sl + u * W) * b.part_stride);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += t[u];
        }
        for (; sl < b.nslots; sl += W) s += __ldcg(src + (int64_t)sl * b.part_stride);
    }
    sred[warp][lane] = s;
    __syncthreads();
    if (!live || wi != 0) return false;
    g = 0.f;
    for (int k = 0; k < c.W; ++k) g += sred[bi * c.W + k][lane];
    return true;
}

template <bool ADAM>
__global__ void __launch_bounds__(256) grad_reduce_adam_kernel(const GradCta* __restrict__ ctas, const GradBlock* __restrict__ blocks,
                                                               const float* __restrict__ part,
                                                               float* __restrict__ grads, float* __restrict__ p,
                                                               float* __restrict__ m, float* __restrict__ v,
                                                               const AdamHyper* __restrict__ hy, int early_loads) {
    __shared__ float sred[8][GRAD_BLK];
    // Before the programmatic-dependency wait: the block tables (constant) and the parameter / moment values this thread will
    // update (last written by the PREVIOUS step's Adam) — three dependent round trips leave the tail of the step.  Only the
    // partials and the step size (bumped by this step's head kernel) have to wait.
    const GradCta c = ctas[blockIdx.x];
    float mm = 0.f, vv = 0.f, pp = 0.f;
    if (ADAM && early_loads) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int bi = warp / c.W, wi = warp - bi * c.W;
        if (bi < c.nblk && wi == 0) {
            const GradBlock b0 = blocks[c.first + bi];
            if (lane < b0.count) { const int64_t e0 = b0.param_off + lane; mm = m[e0]; vv = v[e0]; pp = p[e0]; }
        }
    }
    pdl_wait();
    pdl_trigger();
    float alpha = 0.f, omb1 = 0.f, omb2 = 0.f, eps = 0.f;       // (bumped by this step's head kernel: after the wait, before the partials)
    if (ADAM) { alpha = hy->alpha; omb1 = hy->omb1; omb2 = hy->omb2; eps = hy->eps; }
    GradBlock b;
    float g;
    if (!grad_block_reduce(c, blocks, part, sred, b, g)) return;
    const int64_t e = b.param_off + (threadIdx.x & 31);
    if (b.nslots > 0) grads[e] = g;
    else g = grads[e];
    if (ADAM) {
        if (!early_loads) { mm = m[e]; vv = v[e]; pp = p[e]; }
        mm += (g - mm) * omb1;
        vv += (g * g - vv) * omb2;
        m[e] = mm;
        v[e] = vv;
        p[e] = pp - alpha * mm / (sqrtf(vv) + eps);
    }
}

// plain Adam on caller arenas (s2s_adam_step and the data-parallel path after the all-reduce)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, size_t n,
                                                   const AdamHyper* __restrict__ hy_dev, AdamHyper hy_val, int use_dev) {
    const AdamHyper hy = use_dev ? *hy_dev : hy_val;
    const float alpha = hy.alpha;
    const float omb1 = hy.omb1, omb2 = hy.omb2;
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 4 <= n && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0)) {
        const float4 gg = ld4(g + i4);
        float4 mm = ld4(m + i4), vv = ld4(v + i4), pp = ld4(p + i4);
        mm.x += (gg.x - mm.x) * omb1; mm.y += (gg.y - mm.y) * omb1; mm.z += (gg.z - mm.z) * omb1; mm.w += (gg.w - mm.w) * omb1;
        vv.x += (gg.x * gg.x - vv.x) * omb2; vv.y += (gg.y * gg.y - vv.y) * omb2;
        vv.z += (gg.z * gg.z - vv.z) * omb2; vv.w += (gg.w * gg.w - vv.w) * omb2;
        pp.x -= alpha * mm.x / (sqrtf(vv.x) + hy.eps); pp.y -= alpha * mm.y / (sqrtf(vv.y) + hy.eps);
        pp.z -= alpha * mm.z / (sqrtf(vv.z) + hy.eps); pp.w -= alpha * mm.w / (sqrtf(vv.w) + hy.eps);
        st4(m + i4, mm); st4(v + i4, vv); st4(p + i4, pp);
    } else {
        for (size_t i = i4; i < n && i < i4 + 4; ++i) {
            const float gg = g[i];
            float mm = m[i], vv = v[i];
            mm += (gg - mm) * omb1;
            vv += (gg * gg - vv) * omb2;
            m[i] = mm; v[i] = vv;
            p[i] -= alpha * mm / (sqrtf(vv) + hy.eps);
        }
    }
}

static inline int adam_launch(float* p, const float* g, float* m, float* v, size_t n, const AdamHyper* hy_dev,
                              const AdamHyper* hy_val, cudaStream_t st) {
    if (n == 0) return 0;
    AdamHyper hv = hy_val ? *hy_val : make_hyper(0.0, 0.9, 0.999, 1e-7, 1);
    prof_begin(st, "adam", 28.0 * n, 0.0);
    adam_kernel<<<(unsigned)cdiv64((int64_t)cdiv64((int64_t)n, 4), 256), 256, 0, st>>>(p, g, m, v, n, hy_dev, hv, hy_dev != nullptr);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2s
