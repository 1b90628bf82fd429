// dp.cuh — batch-sharded data parallelism over NVLink peer memory, one process per GPU.
//
// The reference trains on one device (model.fit, training.py:102); splitting its batch over G GPUs needs one
// exchange per optimiser step (gradients) and, for exact parity with the single-device batch, one per
// BatchNormalization layer in each direction (batch statistics).  Both are done by OUR kernels with loads /
// stores on peer memory mapped through CUDA IPC — no NCCL call on the data path:
//
//   * every rank owns one "exchange" allocation (flags | BN sums | 2 x dense gradient arena | stats) whose
//     cudaIpcMemHandle_t is exchanged once by the host (torch.distributed is only the rendezvous);
//   * dp_sum_adam_kernel = all-reduce fused with the optimiser: the local fixed-order slot reduction
//     (grad_reduce, optim.cuh) writes the rank's dense gradient into its own exchange buffer; this kernel
//     publishes a flag to every peer (st.release.sys over NVLink), waits for all peers' flags
//     (ld.acquire.sys on local memory), then every thread loads its float4 of the gradient from EVERY
//     rank's buffer (peer loads through NVSwitch), adds them in rank order — identical bits on all replicas —
//     and applies the Keras-form Adam update to its own replica.  One launch replaces ncclAllReduce + Adam.
//   * sync-BN is fused into the BatchNorm CONSUMER kernels (bn_apply / bn_bwd_apply, bn.cuh + dp_exchange_sums in
//     dp_dev.cuh): every CTA reduces the rank's per-CTA partials to per-channel sums in double, CTA 0 publishes and
//     flags them (pushed into every peer's buffer), every CTA waits for the peers' flags and adds the sums in rank order, then
//     finalises with the global element count — a batch split over G GPUs normalises exactly like the reference's
//     single-device batch, with no extra launch.
//   * buffers are double-buffered by step parity: a rank can only overwrite buffer (t & 1) at step t + 2,
//     after the step t + 1 barrier, which every peer signals after it finished reading step t.
//   * every spin has a wall-clock timeout (globaltimer) that raises an error flag instead of hanging the GPU.
#pragma once
#include "common.cuh"
#include "optim.cuh"
#include "dp_dev.cuh"

namespace s2s {

// ---------------------------------------------------------------------------------------
// local slot reduction into the exchange buffer of the current step parity (kernel "A")
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dp_grad_reduce_kernel(const GradCta* __restrict__ ctas, const GradBlock* __restrict__ blocks,
                                                             const float* __restrict__ part, const float* __restrict__ dense, const DpDev d,
                                                             const float* __restrict__ stats_local, float n_local) {
    __shared__ float sred[8][GRAD_BLK];
    const unsigned long long e = *d.epoch + 1;
    float* out = d.grads[d.rank] + (e & 1) * d.n_pad;
    if (blockIdx.x == 0 && threadIdx.x == 0) {     // this rank's loss / accuracy, weighted by its sample count
        float* s = d.stats[d.rank] + (e & 1) * 4;
        s[0] = stats_local[0] * n_local; s[1] = stats_local[1] * n_local; s[2] = n_local;
    }
    GradBlock b;
    float g;
    if (!grad_block_reduce(ctas[blockIdx.x], blocks, part, sred, b, g)) return;
    const int64_t el = b.param_off + (threadIdx.x & 31);
    out[el] = b.nslots > 0 ? g : dense[el];        // head gradients are written densely by the head kernel
}

// ---------------------------------------------------------------------------------------
// all-reduce (peer loads, fixed rank order) fused with Keras-form Adam (kernel "B")
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dp_sum_adam_kernel(const DpDev d, float* __restrict__ grads_local, float* __restrict__ p,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          const AdamHyper* __restrict__ hy, float* __restrict__ stats_global, int apply) {
    const unsigned long long e = *d.epoch + 1;
    const int buf = (int)(e & 1);
    if (blockIdx.x == 0) dp_signal(d, 0, e);       // kernel A of this step has completed (stream order)
    dp_wait(d, 0, e);
    // A peer that timed out here or at any earlier BatchNorm sync point of this step left partial / stale buffers:
    // do NOT touch the weights, the moments or the step counter; the sticky flag is reported through stats_global[2]
    // (read by the host entry points) and s2s_dp_error.
    __shared__ int s_err;
    if (threadIdx.x == 0) s_err = *reinterpret_cast<volatile int*>(d.error);
    __syncthreads();
    const int err = s_err;                            // uniform over the CTA
    const size_t i4 = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (err != 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && stats_global) stats_global[2] = (float)err;
        return;
    }
    if (i4 < d.n_pad) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < d.world; ++r) {
            const float4 t = ld_peer4(d.grads[r] + (size_t)buf * d.n_pad + i4);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        st4(grads_local + i4, g);
        if (apply) {
            const float alpha = hy->alpha, omb1 = hy->omb1, omb2 = hy->omb2, eps = hy->eps;
            float4 mm = ld4(m + i4), vv = ld4(v + i4), pp = ld4(p + i4);
            mm.x += (g.x - mm.x) * omb1; mm.y += (g.y - mm.y) * omb1; mm.z += (g.z - mm.z) * omb1; mm.w += (g.w - mm.w) * omb1;
            vv.x += (g.x * g.x - vv.x) * omb2; vv.y += (g.y * g.y - vv.y) * omb2;
            vv.z += (g.z * g.z - vv.z) * omb2; vv.w += (g.w * g.w - vv.w) * omb2;
            pp.x -= alpha * mm.x / (sqrtf(vv.x) + eps); pp.y -= alpha * mm.y / (sqrtf(vv.y) + eps);
            pp.z -= alpha * mm.z / (sqrtf(vv.z) + eps); pp.w -= alpha * mm.w / (sqrtf(vv.w) + eps);
            st4(m + i4, mm); st4(v + i4, vv); st4(p + i4, pp);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats_global) {     // global (sample-weighted) loss / accuracy
        float sl = 0.f, sa = 0.f, sn = 0.f;
        for (int r = 0; r < d.world; ++r) {
            const float4 t = ld_peer4(d.stats[r] + buf * 4);
            sl += t.x; sa += t.y; sn += t.z;
        }
        stats_global[0] = sl / sn; stats_global[1] = sa / sn; stats_global[2] = 0.f;
    }
    // the last CTA closes the step
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(d.counter, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) { *d.counter = 0u; *d.epoch = e; }
    }
}

// ------------------------------------------------------------------ host-side communicator
struct DpLayout {
    size_t flags_off, bn_off, stats_off, grads_off, total;
};
static inline DpLayout dp_layout(size_t n_pad) {
    DpLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
    L.flags_off = take(sizeof(unsigned long long) * (1 + DP_MAXSYNC) * DP_MAXW);
    L.bn_off = take(sizeof(double) * DP_MAXSYNC * 2 * DP_MAXW * 2 * DP_BN_MAXC);     // [sync][parity][source rank][2C]
    L.stats_off = take(sizeof(float) * 2 * 4);
    L.grads_off = take(sizeof(float) * 2 * n_pad);
    L.total = o;
    return L;
}

}  // namespace s2s

struct s2s_dp {
    int rank = 0, world = 1;
    size_t n_pad = 0;
    s2s::DpLayout lay{};
    char* local = nullptr;                       // this rank's exchange allocation
    char* mapped[s2s::DP_MAXW] = {};             // peers' allocations (mapped[rank] == local)
    bool opened[s2s::DP_MAXW] = {};
    char* state = nullptr;                       // local: epoch | error | counter
    bool connected = false;
    s2s::DpDev dev{};
};
