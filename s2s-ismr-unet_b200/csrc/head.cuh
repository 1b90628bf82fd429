// head.cuh — 1x1 output convolution fused with softmax, the Keras loss, its gradient and accuracy.
//
// Reference: Conv2D(3,(1,1),softmax) | Conv2D(1,(1,1),relu)  (deep_nn_models.py:102-105) and
// loss="categorical_crossentropy", metrics=['accuracy']        (training.py:67,96).
// Keras-3 CCE: p <- p / sum(p); p <- clip(p, 1e-7, 1-1e-7); l = -sum_c t_c log p_c;
// loss = mean over N*H*W.  The gradient below is the exact derivative of that formulation
// (clip passes gradient only inside [1e-7, 1-1e-7]), not the fused softmax-CE shortcut.
// The "deterministic" head (relu, 1 channel) is paired with a masked MSE: sum m (relu(z)-t)^2 / sum m.
//
// Thread = one pixel.  Per-CTA partials of {loss, correct, d bias, d kernel} are reduced in fixed
// order; the last CTA writes the head gradients straight into the grad arena, the batch statistics
// and the running epoch accumulators, and bumps the optimiser step counter.
#pragma once
#include "common.cuh"
#include "optim.cuh"

namespace s2s {

struct HeadArgs {
    const float* u; int ldu;           // head input [npix][C0] (ELU output of up_conv1_3, no BN)
    const float* wh; const float* bh;  // kernel (1,1,C0,NC) = [c][k]; bias [NC]
    const float* y;                    // targets [npix][NC] (one-hot for CCE)
    const uint8_t* mask; int hw;       // MASKED_MSE: mask[pix % hw]; nullable
    float mask_norm;                   // MASKED_MSE: 1 / (N * sum(mask))
    float* probs;                      // [npix][NC] or null
    float* dz_out;                     // [npix][C0] gradient wrt the pre-activation of the head input layer, or null
    int apply_elugrad, act;            // multiply dz_out by act'(u)
    float* part; unsigned int* counter;
    float* dwh; float* dbh;            // grad arena (train) or null
    float* stats;                      // [2] batch mean loss, accuracy
    double* stats_acc;                 // [3] running sum loss, correct, pixels
    AdamHyper* hyper;                  // optimiser step counter + alpha (bumped when training)
    float grad_scale;                  // multiplies d loss (1/world under DP)
    const float* gscale_dev;           // if non-null overrides grad_scale (graph-replay safe)
    int64_t npix;
    int loss_kind, train;
    int cam_cls;                       // >=0: Grad-CAM seed instead of a loss: S = mean_hw p[cam_cls]
    float cam_norm;                    // 1/(H*W)
    int early_loads;                   // stage the head weights before the programmatic-dependency wait
    int defer_final;                   // the per-CTA partials are finalised by head_final_kernel (side stream) instead of the last CTA
};

// Fixed-order sum of the per-CTA partials -> head gradients, batch statistics, epoch accumulators, optimiser step counter.
// Called by all 256 threads of ONE CTA: the elected last CTA of head_kernel, or head_final_kernel.
template <int C0, int NC>
__device__ __forceinline__ void head_finalize(const HeadArgs& a, int nparts, int tid) {
    constexpr int NV = C0 * NC + NC + 2;
    __shared__ double sfin[NV];
    __shared__ double stmp[256];
    cta_reduce_slots<256>(a.part, nparts, (size_t)NV, NV, stmp, sfin, tid);
    if (tid < NV) {
        const double s = sfin[tid];
        if (a.train && a.dwh) {
            if (tid < C0 * NC) a.dwh[tid] = (float)s;
            else if (tid < C0 * NC + NC) a.dbh[tid - C0 * NC] = (float)s;
        }
    }
    __syncthreads();
    if (tid == 0 && a.cam_cls < 0) {
        const double lsum = sfin[C0 * NC + NC], csum = sfin[C0 * NC + NC + 1];
        const double norm = (a.loss_kind == S2S_LOSS_MASKED_MSE) ? (double)a.mask_norm : 1.0 / (double)a.npix;
        if (a.stats) { a.stats[0] = (float)(lsum * norm); a.stats[1] = (float)(csum / (double)a.npix); }
        if (a.stats_acc) { a.stats_acc[0] += lsum * norm * (double)a.npix; a.stats_acc[1] += csum; a.stats_acc[2] += (double)a.npix; }
        if (a.train && a.hyper) adam_bump(a.hyper);
    }
}

// The training step runs this on a side stream: the gradient chain that follows the head kernel needs dz_out only, so the
// election fence, the 256-row partial sum and the statistics leave the critical path (joined again in front of Adam).
template <int C0, int NC>
__global__ void __launch_bounds__(256) head_final_kernel(const HeadArgs a, int nparts) {
    head_finalize<C0, NC>(a, nparts, threadIdx.x);
}

template <int C0, int NC>
__global__ void __launch_bounds__(256) head_kernel(const HeadArgs a) {
    constexpr int NV = C0 * NC + NC + 2;   // d kernel, d bias, loss, correct
    __shared__ float swh[C0 * NC + NC];
    __shared__ float sred[8][NV];
    const int tid = threadIdx.x;
    // the head weights were last written by the previous step's Adam: staged before the programmatic-dependency wait
    if (!a.early_loads) pdl_wait();
    for (int i = tid; i < C0 * NC; i += 256) swh[i] = a.wh[i];
    for (int i = tid; i < NC; i += 256) swh[C0 * NC + i] = a.bh[i];
    __syncthreads();
    if (a.early_loads) pdl_wait();

    const int64_t pix = (int64_t)blockIdx.x * 256 + tid;
    float vals[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) vals[i] = 0.f;

    if (pix < a.npix) {
        float u[C0];
#pragma unroll
        for (int c4 = 0; c4 < C0 / 4; ++c4) {
            const float4 v = ld4(a.u + (size_t)pix * a.ldu + 4 * c4);
            u[4 * c4 + 0] = v.x; u[4 * c4 + 1] = v.y; u[4 * c4 + 2] = v.z; u[4 * c4 + 3] = v.w;
        }
        // targets and the loss-gradient scale are fetched together with the activations (a store to `probs` further down
        // would otherwise pin these loads behind the softmax arithmetic: one more dependent round trip on the chain)
        float t_pre[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) t_pre[k] = (a.y != nullptr && a.cam_cls < 0) ? a.y[(size_t)pix * NC + k] : 0.f;
        const float gs_pre = a.gscale_dev ? __ldg(a.gscale_dev) : a.grad_scale;
        float z[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float s = swh[C0 * NC + k];
#pragma unroll
            for (int c = 0; c < C0; ++c) s = fmaf(u[c], swh[c * NC + k], s);
            z[k] = s;
        }
        float dz[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) dz[k] = 0.f;
        float loss = 0.f, correct = 0.f;
        if constexpr (NC == 3) {
            const float mx = fmaxf(z[0], fmaxf(z[1], z[NC - 1]));
            float p[NC], S = 0.f;
#pragma unroll
            for (int k = 0; k < NC; ++k) { p[k] = expf(z[k] - mx); S += p[k]; }
            const float inv = 1.f / S;
#pragma unroll
            for (int k = 0; k < NC; ++k) p[k] *= inv;
            if (a.probs) {
#pragma unroll
                for (int k = 0; k < NC; ++k) a.probs[(size_t)pix * NC + k] = p[k];
            }
            if (a.cam_cls >= 0) {
                // seed: S_c = mean_hw p[c]  ->  dS/dz_j = p_c (delta_jc - p_j) / (H*W)
                const float pc = p[a.cam_cls];
#pragma unroll
                for (int k = 0; k < NC; ++k) dz[k] = pc * ((k == a.cam_cls ? 1.f : 0.f) - p[k]) * a.cam_norm;
            } else if (a.y) {
                float t[NC];
#pragma unroll
                for (int k = 0; k < NC; ++k) t[k] = t_pre[k];
                // Keras CCE on probabilities: renormalise, clip, -sum t log q
                float Sp = 0.f;
#pragma unroll
                for (int k = 0; k < NC; ++k) Sp += p[k];
                const float invS = 1.f / Sp;
                float gq[NC], q[NC], gdot = 0.f;
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    q[k] = p[k] * invS;
                    const bool inside = (q[k] >= 1e-7f) && (q[k] <= 1.f - 1e-7f);
                    const float qc = fminf(fmaxf(q[k], 1e-7f), 1.f - 1e-7f);
                    loss -= t[k] * logf(qc);
                    gq[k] = inside ? -t[k] / qc : 0.f;     // d l / d q_k
                    gdot = fmaf(gq[k], q[k], gdot);
                }
                // d l / d p_j = (gq_j - sum_k gq_k q_k) / Sp ;  d l / d z_j = p_j (gp_j - sum_k p_k gp_k)
                float gp[NC], pdot = 0.f;
#pragma unroll
                for (int k = 0; k < NC; ++k) { gp[k] = (gq[k] - gdot) * invS; pdot = fmaf(p[k], gp[k], pdot); }
                const float gs = gs_pre / (float)a.npix;
#pragma unroll
                for (int k = 0; k < NC; ++k) dz[k] = p[k] * (gp[k] - pdot) * gs;
                // accuracy: argmax(pred) == argmax(target), first maximum wins (np.argmax)
                int ap = 0, at = 0;
#pragma unroll
                for (int k = 1; k < NC; ++k) { if (p[k] > p[ap]) ap = k; if (t[k] > t[at]) at = k; }
                correct = (ap == at) ? 1.f : 0.f;
            }
        } else {
            const float pred = fmaxf(z[0], 0.f);
            if (a.probs) a.probs[pix] = pred;
            if (a.cam_cls >= 0) {
                dz[0] = (z[0] > 0.f ? 1.f : 0.f) * a.cam_norm;
            } else if (a.y) {
                const float m = a.mask ? (a.mask[pix % a.hw] ? 1.f : 0.f) : 1.f;
                const float d = pred - t_pre[0];
                loss = m * d * d;                                   // normalised by mask_norm in the finalize
                dz[0] = (z[0] > 0.f) ? 2.f * m * d * a.mask_norm * gs_pre : 0.f;
                correct = 0.f;
            }
        }
        // gradient wrt the head input, then through ELU of the producing layer
        if (a.dz_out) {
            float du[C0];
#pragma unroll
            for (int c = 0; c < C0; ++c) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < NC; ++k) s = fmaf(dz[k], swh[c * NC + k], s);
                du[c] = a.apply_elugrad ? s * act_grad_from_out(u[c], a.act) : s;
            }
#pragma unroll
            for (int c4 = 0; c4 < C0 / 4; ++c4)
                st4(a.dz_out + (size_t)pix * C0 + 4 * c4, make_float4(du[4 * c4], du[4 * c4 + 1], du[4 * c4 + 2], du[4 * c4 + 3]));
        }
#pragma unroll
        for (int c = 0; c < C0; ++c)
#pragma unroll
            for (int k = 0; k < NC; ++k) vals[c * NC + k] = u[c] * dz[k];
#pragma unroll
        for (int k = 0; k < NC; ++k) vals[C0 * NC + k] = dz[k];
        vals[C0 * NC + NC] = loss;
        vals[C0 * NC + NC + 1] = correct;
    }

    pdl_trigger();          // the first dgrad kernel's prologue (weight slab) overlaps the reduction tail
    // CTA reduction: warp shuffles, then fixed-order over the 8 warps
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float s = warp_sum(vals[i]);
        if (lane == 0) sred[warp][i] = s;
    }
    __syncthreads();
    if (tid < NV) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sred[w][tid];
        a.part[(size_t)blockIdx.x * NV + tid] = s;
    }
    if (!a.defer_final && cta_is_last(a.counter, gridDim.x)) head_finalize<C0, NC>(a, (int)gridDim.x, tid);
}

static inline int head_part_floats(int C0, int NC, int64_t npix) {
    return (int)(cdiv64(npix, 256) * (C0 * NC + NC + 2));
}

// plain launch (no programmatic dependency: the kernel has no griddepcontrol.wait) on the stream the caller forked after head_launch
static inline int head_final_launch(const HeadArgs& a, int C0, int NC, cudaStream_t st) {
    const int nparts = (int)cdiv64(a.npix, 256);
    prof_begin(st, "head_finalize", 4.0 * nparts * (C0 * NC + NC + 2), 0.0);
    if (C0 == 8 && NC == 3) head_final_kernel<8, 3><<<1, 256, 0, st>>>(a, nparts);
    else if (C0 == 12 && NC == 3) head_final_kernel<12, 3><<<1, 256, 0, st>>>(a, nparts);
    else if (C0 == 8 && NC == 1) head_final_kernel<8, 1><<<1, 256, 0, st>>>(a, nparts);
    else if (C0 == 12 && NC == 1) head_final_kernel<12, 1><<<1, 256, 0, st>>>(a, nparts);
    else if (C0 == 4 && NC == 3) head_final_kernel<4, 3><<<1, 256, 0, st>>>(a, nparts);
    else if (C0 == 16 && NC == 3) head_final_kernel<16, 3><<<1, 256, 0, st>>>(a, nparts);
    else if (C0 == 4 && NC == 1) head_final_kernel<4, 1><<<1, 256, 0, st>>>(a, nparts);
    else if (C0 == 16 && NC == 1) head_final_kernel<16, 1><<<1, 256, 0, st>>>(a, nparts);
    else return fail(S2S_ERR_INVALID, "head: unsupported filters*4=%d / classes=%d", C0, NC);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

static inline int head_launch(const HeadArgs& a, int C0, int NC, cudaStream_t st) {
    const unsigned grid = (unsigned)cdiv64(a.npix, 256);
    prof_begin(st, "head_softmax_loss", 4.0 * a.npix * (C0 + (a.y ? NC : 0) + (a.dz_out ? C0 : 0) + (a.probs ? NC : 0)),
               2.0 * a.npix * C0 * NC * (a.dz_out ? 3.0 : 1.0));
    if (C0 == 8 && NC == 3) launch_k(head_kernel<8, 3>, grid, 256, 0, st, a);
    else if (C0 == 12 && NC == 3) launch_k(head_kernel<12, 3>, grid, 256, 0, st, a);
    else if (C0 == 8 && NC == 1) launch_k(head_kernel<8, 1>, grid, 256, 0, st, a);
    else if (C0 == 12 && NC == 1) launch_k(head_kernel<12, 1>, grid, 256, 0, st, a);
    else if (C0 == 4 && NC == 3) launch_k(head_kernel<4, 3>, grid, 256, 0, st, a);
    else if (C0 == 16 && NC == 3) launch_k(head_kernel<16, 3>, grid, 256, 0, st, a);
    else if (C0 == 4 && NC == 1) launch_k(head_kernel<4, 1>, grid, 256, 0, st, a);
    else if (C0 == 16 && NC == 1) launch_k(head_kernel<16, 1>, grid, 256, 0, st, a);
    else return fail(S2S_ERR_INVALID, "head: unsupported filters*4=%d / classes=%d", C0, NC);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2s
