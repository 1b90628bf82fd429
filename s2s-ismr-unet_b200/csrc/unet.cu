// unet.cu — model handle, step orchestration (CUDA-graph replay) and the C ABI of libs2s_unet.so.
//
// Mirrors the graph built by Unet.build_model (utils/deep_nn_models.py:73-163) and the per-step
// work of model.fit / model.predict (utils/training.py:95-103,133-135).  See include/s2s_unet.h.
#include <algorithm>
#include <map>
#include <vector>
#include <string>
#include <math.h>
#include <stdlib.h>
#include <mutex>

#include "common.cuh"
#include "gconv.cuh"
#include "wgrad.cuh"
#include "convt.cuh"
#include "bn.cuh"
#include "optim.cuh"
#include "head.cuh"
#include "skill.cuh"
#include "tcconv.cuh"
#include "tc3conv.cuh"
#include "tcwgrad.cuh"
#include "micro.cuh"
#include "dp.cuh"

using namespace s2s;

namespace {

constexpr int MAXB = 5;

struct ConvL {
    int Cin = 0, Cout = 0, H = 0, W = 0;
    int64_t w_off = 0, b_off = 0;        // parameter arena
    int64_t part_off = 0, bpart_off = 0; // wgrad partial workspace
    int nslots = 0;
    std::string name;
    // tensor-core inference path (precision = BF16_TC): TMA maps over the shared bf16 input scratch / bf16 weights
    bool tc = false;
    int64_t wb_off = 0;
    CUtensorMap map_a, map_b;
    // tcgen05 tf32 training / inference path (tc3conv.cuh): forward (input x) and input-gradient (input dz) plans, their
    // per-step weight blocks inside h->wq, and TMA maps built lazily for the tensor each direction actually reads
    bool t3f = false, t3d = false;
    Tc3Plan pf{}, pd{};
    int64_t wqf_off = 0, wqd_off = 0;
    mutable CUtensorMap t3map_f, t3map_d;
    mutable const float *t3f_in = nullptr, *t3d_in = nullptr;
    // tcgen05 tf32 weight gradient (tcwgrad.cuh): plan + TMA maps over (x, dz), rebuilt when the tensors or the batch change
    bool twg = false;
    TcWgPlan pw{};
    mutable CUtensorMap wgmap_x, wgmap_z;
    mutable const float *wg_x = nullptr, *wg_z = nullptr;
    mutable int wg_N = 0;
};
struct ConvTL {
    int Cin = 0, Cout = 0, h = 0, w = 0, k = 0;   // input grid h x w, output 2h x 2w
    int64_t w_off = 0, b_off = 0, part_off = 0;
    int nslots = 0;
    int64_t cs_part_off = 0;             // bias-gradient partials [cs_slots][Cout] inside gpart
    int cs_slots = 0;
    std::string name;
    // tcgen05 tf32 path (tc3conv.cuh, precision = tf32): forward = four parity 3x3 convs on the input grid, input gradient =
    // one conv over the four stride-2 parity planes of dy; weight blocks inside h->wq
    bool t3f = false, t3d = false;
    Tc3Plan pf{}, pd{};
    int64_t wqf_off = 0, wqd_off = 0;
    mutable Tc3Maps t3maps_f, t3maps_d;
    mutable const float *t3f_in = nullptr, *t3d_in = nullptr;
    // tcgen05 tf32 weight gradient (tcwgrad.cuh): x is the no-halo operand, the parity planes of dy the halo operands
    bool twg = false;
    TcWgPlan pw{};
    mutable Tc3Maps wgmaps_x;
    mutable CUtensorMap wgmap_z;
    mutable const float *wg_x = nullptr, *wg_dy = nullptr;
    mutable int wg_N = 0;
};
struct BnL {
    bool on = false;
    int C = 0;
    int64_t g_off = 0, be_off = 0;       // parameter arena
    int64_t mm_off = 0, mv_off = 0;      // state arena
    int64_t ch_off = 0;                  // offset into the per-channel buffers
    int64_t part_off = 0;                // backward partials [bwd_slots][2][C] inside gpart (= beta/gamma grad partials)
    int bwd_slots = 0;
    std::string name;
};

struct Bump {   // carve one device allocation
    size_t off = 0;
    size_t take(size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; }
};

enum GraphKind { GK_TRAIN = 0, GK_BWD = 1, GK_EVAL = 2, GK_FWD_INFER = 3, GK_FWD_TRAIN = 4, GK_DP = 5 };
enum { GK_HOST_FWD = 3 + 100, GK_HOST_BWD = 4 + 100, GK_HOST_DP_FWD = 6, GK_HOST_DP_BWD = 7 };   // halves of the end-to-end step (DP kinds >= GK_DP are dropped on re-attach)

}  // namespace

struct s2s_unet {
    s2s_unet_cfg cfg;
    int nb = 0, NC = 3, C0 = 8, pbk = 0;
    ConvL dconv[MAXB][2], bconv[2], uconv[MAXB][2];
    ConvTL upT[MAXB];
    BnL dbn[MAXB], bbn, ubn[MAXB];
    int64_t head_w = 0, head_b = 0;
    std::vector<s2s_tensor_desc> descs;
    size_t n_params = 0, n_state = 0, n_bnch = 0, gpart_floats = 0;

    char* pool = nullptr;   // single device allocation
    int device = 0;         // CUDA device the handle (pool, streams, graphs, TMA maps) lives on
    size_t pool_bytes = 0;
    float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr, *state = nullptr;
    AdamHyper* hyper = nullptr;
    AdamHyper hyper_host{};
    float* gscale = nullptr;            // device grad scale (1 float)
    float gscale_host = 1.f;
    float* stats = nullptr;             // [2]
    double* stats_acc = nullptr;        // [3]
    float *x_in = nullptr, *y_in = nullptr, *probs = nullptr;
    float *a1[MAXB] = {}, *a2[MAXB] = {}, *cat[MAXB] = {}, *pl[MAXB] = {};
    float *ab1 = nullptr, *ab2 = nullptr, *cb = nullptr;
    float *ua1[MAXB] = {}, *ua2[MAXB] = {}, *uo[MAXB] = {};
    float *dz_a1[MAXB] = {}, *dz_a2[MAXB] = {}, *dcat[MAXB] = {}, *dpl[MAXB] = {};
    float *dz_ab1 = nullptr, *dz_ab2 = nullptr, *dcb = nullptr;
    float *dz_ua1[MAXB] = {}, *dz_ua2[MAXB] = {}, *duo[MAXB] = {};
    float *bn_scale = nullptr, *bn_shift = nullptr, *bn_mean = nullptr, *bn_rstd = nullptr;
    float* wt = nullptr;                // flipped + transposed Conv2D kernels for dgrad (rebuilt every step)
    __nv_bfloat16* xb = nullptr;        // bf16 copy of the current thick layer's input (tensor-core path)
    __nv_bfloat16* wb = nullptr;        // bf16 [tap][co][ci] kernels of the tensor-core layers
    bool tc_mode = false;
    int t3_npass = 0;                   // 0 = off; 1 = precision TF32 (single pass); 3 = 3xTF32 split (fp32 parity, S2S_TC3_FP32=1)
    float* wq = nullptr;                // per-step tf32 weight blocks of the tcgen05 path (tc3_wprep_kernel)
    int t3prep_maxcount = 0;            // elements of the largest weight-block set (grid sizing)
    void* t3prep_tab = nullptr;
    int n_t3prep = 0, n_t3prep_fwd = 0;  // weight-block table: [0, n_t3prep_fwd) forward-direction entries, then the gradient ones
    float *ones = nullptr, *zeros = nullptr;
    float *stat_part = nullptr, *head_part = nullptr, *gpart = nullptr;
    void* wprep_tab = nullptr;
    int n_wprep = 0, wprep_maxcount = 0;
    cudaEvent_t ev_wprep = nullptr;
    bool wprep_pending = false;
    cudaEvent_t ev_wq = nullptr;         // tf32 weight blocks (h->wq) ready: joined before the FIRST tensor-core conv of the forward pass
    bool wq_record = false, wq_pending = false;
    float* cam_grad = nullptr;
    unsigned int* counters = nullptr;
    int n_counters = 0;
    GradBlock* blocks_dev = nullptr;
    int nblocks = 0;
    GradCta* ctas_dev = nullptr;        // CTA table of the fused reduce + Adam kernel (optim.cuh)
    int nctas = 0;
    BnFoldEntry* fold_dev = nullptr;
    int nfold = 0;

    int loss_kind = S2S_LOSS_CCE;
    bool compiled = false;
    bool use_graphs = true;
    bool last_forward_training = false;
    int last_N = 0;
    int64_t launches = 0;
    std::map<long long, std::pair<cudaGraphExec_t, int>> graphs;   // key -> (exec, kernels per replay)
    float mask_norm_cache = 0.f, mask_count = 0.f;
    // side streams: weight-gradient kernels run concurrently with the dgrad chain (forked / joined with
    // events, which also works under stream capture -> parallel branches of the CUDA graph)
    static constexpr int NSIDE = 2, NEV = 96;
    cudaStream_t side[NSIDE] = {};
    cudaEvent_t ev[NEV] = {};
    int ev_next = 0, side_next = 0;
    bool use_side = true, side_used[NSIDE] = {};
    bool early_loads = true;                       // loads that do not depend on the preceding kernel are issued before the statistics
                                                   // finalize / the programmatic-dependency wait (bn.cuh, optim.cuh, head.cuh)
    bool bn_fold = true, bn_fold_pool = true;      // BatchNorm backward statistics in the epilogue of the kernel that produces dc
    int64_t bn_fold_max = 3 << 19;                 // ... for layers of at most this many elements (the latency regime: at batch 128 the
                                                   // separate reduction kernel is the faster one, 1454 vs 1475 us per step)
    const uint8_t* mask_cache = nullptr;
    // data parallelism over peer memory (dp.cuh): attached communicator, sync-BN workspace
    s2s_dp* dp = nullptr;
    bool dp_sync_bn = false;
    cudaStream_t copy_stream = nullptr; // end-to-end step: H2D of the targets overlaps the forward pass
    cudaEvent_t ev_y = nullptr;
    // streamed end-to-end steps (s2s_unet_train_steps_host): two device staging slots filled by the copy stream one step ahead
    float* stage_x[2] = {nullptr, nullptr};
    float* stage_y[2] = {nullptr, nullptr};
    cudaEvent_t ev_staged[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    float* stream_stats_pinned = nullptr;   // [STREAM_CHUNK][2] pinned
    // pageable host batches (the reference hands model.fit ordinary NumPy arrays): a ring of pinned staging slots, filled by the
    // calling thread while the GPU computes, each guarded by the event of the H2D copy that last read it
    static constexpr int NPIN = 3;
    float* pin_x[NPIN] = {nullptr, nullptr, nullptr};
    float* pin_y[NPIN] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_pin[NPIN] = {nullptr, nullptr, nullptr};
    bool pin_busy[NPIN] = {false, false, false};
    int pin_next = 0;
    bool dp_in_step = false;            // true only while s2s_unet_dp_train_step enqueues / captures its sequence
    int dp_n_global = 0, dp_sync_next = 0;
    float* stats_global = nullptr;      // [4] sample-weighted {loss, accuracy} over all ranks, exchange error code
    float* dp_stats_host = nullptr;     // pinned mirror of stats_global for the host entry point
};

namespace {

int levelH(const s2s_unet* h, int b) { return h->cfg.H >> b; }
int levelW(const s2s_unet* h, int b) { return h->cfg.W >> b; }
int levelC(const s2s_unet* h, int b) { return h->cfg.filters * 4 * (1 << b); }

// Device-pool cache: the tuning loops create and destroy hundreds of handles (training.py:87-93); cudaMalloc /
// cudaFree of a few hundred MB cost tens of milliseconds each, so freed pools are kept (up to 4, <= 8 GB) and
// re-used by the next handle that fits.  Pools are zero-filled on (re)use.
struct PoolEntry { char* p; size_t bytes; int dev; };
struct PoolCache {
    std::mutex mu;
    std::vector<PoolEntry> free_list;       // keyed by CUDA device: a pool is only handed back to a handle on ITS device
    size_t bytes = 0;
};
PoolCache& pool_cache() {
    static PoolCache c;
    return c;
}
int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}
cudaError_t pool_acquire(size_t need, char** out, size_t* got) {
    PoolCache& c = pool_cache();
    const int dev = current_device();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        int best = -1;
        for (int i = 0; i < (int)c.free_list.size(); ++i)
            if (c.free_list[i].dev == dev && c.free_list[i].bytes >= need && c.free_list[i].bytes <= 4 * need + (64u << 20) &&
                (best < 0 || c.free_list[i].bytes < c.free_list[best].bytes)) best = i;
        if (best >= 0) {
            *out = c.free_list[best].p; *got = c.free_list[best].bytes;
            c.bytes -= *got;
            c.free_list.erase(c.free_list.begin() + best);
            return cudaSuccess;
        }
    }
    *got = need;
    cudaError_t e = cudaMalloc((void**)out, need);
    if (e == cudaErrorMemoryAllocation) {      // make room on this device and retry once
        cudaGetLastError();
        PoolCache& cc = pool_cache();
        std::lock_guard<std::mutex> lk(cc.mu);
        for (auto it = cc.free_list.begin(); it != cc.free_list.end();) {
            if (it->dev == dev) { cudaFree(it->p); cc.bytes -= it->bytes; it = cc.free_list.erase(it); }
            else ++it;
        }
        e = cudaMalloc((void**)out, need);
    }
    return e;
}
void pool_release(char* p, size_t bytes, int dev) {
    if (!p) return;
    PoolCache& c = pool_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    int on_dev = 0;
    for (const auto& e : c.free_list) on_dev += e.dev == dev;
    if (on_dev < 4 && c.bytes + bytes <= ((size_t)8 << 30)) {
        c.free_list.push_back(PoolEntry{p, bytes, dev});
        c.bytes += bytes;
    } else {
        int cur = current_device();
        if (cur != dev) cudaSetDevice(dev);
        cudaFree(p);
        if (cur != dev) cudaSetDevice(cur);
    }
}

void add_desc(s2s_unet* h, const std::string& name, int arena, int ndim, const int* shape, int64_t off, int64_t count) {
    s2s_tensor_desc d;
    memset(&d, 0, sizeof d);
    snprintf(d.name, sizeof d.name, "%s", name.c_str());
    d.arena = arena; d.ndim = ndim;
    for (int i = 0; i < ndim; ++i) d.shape[i] = shape[i];
    d.offset = off; d.count = count;
    h->descs.push_back(d);
}

// parameter offsets are kept multiples of 4 floats so that every tensor is float4-aligned
int64_t take_param(s2s_unet* h, int64_t count) {
    int64_t off = (int64_t)h->n_params;
    h->n_params += (size_t)((count + 3) / 4 * 4);
    return off;
}
int64_t take_state(s2s_unet* h, int64_t count) {
    int64_t off = (int64_t)h->n_state;
    h->n_state += (size_t)((count + 3) / 4 * 4);
    return off;
}

void def_conv(s2s_unet* h, ConvL& L, const std::string& name, int Cin, int Cout, int H, int W) {
    L.name = name; L.Cin = Cin; L.Cout = Cout; L.H = H; L.W = W;
    L.w_off = take_param(h, (int64_t)9 * Cin * Cout);
    L.b_off = take_param(h, Cout);
    const int ks[4] = {3, 3, Cin, Cout};
    add_desc(h, name + "/kernel", 0, 4, ks, L.w_off, (int64_t)9 * Cin * Cout);
    const int bs[1] = {Cout};
    add_desc(h, name + "/bias", 0, 1, bs, L.b_off, Cout);
}
void def_convt(s2s_unet* h, ConvTL& L, const std::string& name, int Cin, int Cout, int hh, int ww, int k) {
    L.name = name; L.Cin = Cin; L.Cout = Cout; L.h = hh; L.w = ww; L.k = k;
    L.w_off = take_param(h, (int64_t)k * k * Cin * Cout);
    L.b_off = take_param(h, Cout);
    const int ks[4] = {k, k, Cout, Cin};
    add_desc(h, name + "/kernel", 0, 4, ks, L.w_off, (int64_t)k * k * Cin * Cout);
    const int bs[1] = {Cout};
    add_desc(h, name + "/bias", 0, 1, bs, L.b_off, Cout);
}
void def_bn(s2s_unet* h, BnL& L, int& bn_index, int C, bool on) {
    L.on = on; L.C = C;
    L.ch_off = (int64_t)h->n_bnch;
    h->n_bnch += (size_t)C;
    if (!on) return;
    L.name = bn_index == 0 ? std::string("batch_normalization") : "batch_normalization_" + std::to_string(bn_index);
    ++bn_index;
    L.g_off = take_param(h, C);
    L.be_off = take_param(h, C);
    L.mm_off = take_state(h, C);
    L.mv_off = take_state(h, C);
    const int s[1] = {C};
    add_desc(h, L.name + "/gamma", 0, 1, s, L.g_off, C);
    add_desc(h, L.name + "/beta", 0, 1, s, L.be_off, C);
    add_desc(h, L.name + "/moving_mean", 1, 1, s, L.mm_off, C);
    add_desc(h, L.name + "/moving_variance", 1, 1, s, L.mv_off, C);
}

// ---------------------------------------------------------------------------------------
// dgrad weight preparation: Wt[tap'][co][ci] = W[8 - tap'][ci][co] for every Conv2D 3x3 kernel
// (flip + transpose once per step, so that dgrad is the same gather convolution as forward)
// ---------------------------------------------------------------------------------------
// Per-step weight preparation (one launch, overlapped with the forward pass on a side stream):
//   Conv2D 3x3 (flip=1): Wt[tap'][co][ci] = W[8-tap'][ci][co]   -> dgrad becomes the same gather conv as forward
//   Conv2DTranspose (flip=0): Wt[tap][ci][co] = W[tap][co][ci]  -> forward reads warp-broadcast rows over co
// Generic form: dst[tap][r][c] = src[flip ? K2-1-tap : tap][c][r], r < R, c < Cc.
struct WPrepEntry { int64_t w_off; int R, Cc, K2, flip; };

__global__ void wprep_kernel(const WPrepEntry* __restrict__ tab, const float* __restrict__ params, float* __restrict__ wt) {
    const WPrepEntry e = tab[blockIdx.y];
    const int total = e.K2 * e.R * e.Cc;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % e.Cc;
        const int r = (i / e.Cc) % e.R;
        const int tap = i / (e.Cc * e.R);
        const int ts = e.flip ? e.K2 - 1 - tap : tap;
        wt[e.w_off + i] = __ldg(params + e.w_off + ((size_t)ts * e.Cc + c) * e.R + r);
    }
}

int run_wprep(s2s_unet* h, cudaStream_t st) {
    if (h->n_wprep == 0 && h->n_t3prep == 0) return 0;
    if (h->n_t3prep) {      // first: the forward pass needs these blocks at its first tensor-core conv
        if (h->n_t3prep_fwd > 0) prof_begin(st, "wprep_tf32", 12.0 * h->n_params, 0.0);
        // grid.x sized for the largest entry (>= 4 elements per thread): 32 CTAs per layer left the 1.3 M-element kernels of the
        // deep grid points at 350 GB/s on the forward pass's critical path (grid_max: 218 us)
        const int gx = std::max(32, std::min(cdiv(std::max(h->t3prep_maxcount, 1), 1024), 592));
        const Tc3WPrep* tab = reinterpret_cast<const Tc3WPrep*>(h->t3prep_tab);
        // two launches: the forward-direction blocks (what the forward pass waits for), then the dgrad ones behind them
        if (h->n_t3prep_fwd > 0) {
            tc3_wprep_kernel<<<dim3(gx, h->n_t3prep_fwd), 256, 0, st>>>(tab, h->params, h->wq);
            prof_end(st);
            S2S_LAUNCH_CHECK();
        }
        if (h->wq_record) { S2S_CUDA(cudaEventRecord(h->ev_wq, st)); h->wq_pending = true; }
        if (h->n_t3prep > h->n_t3prep_fwd) {
            prof_begin(st, "wprep_tf32_grad", 12.0 * h->n_params, 0.0);
            tc3_wprep_kernel<<<dim3(gx, h->n_t3prep - h->n_t3prep_fwd), 256, 0, st>>>(tab + h->n_t3prep_fwd, h->params, h->wq);
            prof_end(st);
            S2S_LAUNCH_CHECK();
        }
    }
    if (h->n_wprep) {
        dim3 grid(std::max(std::min(cdiv(std::max(h->wprep_maxcount, 1), 256), 32), std::min(cdiv(std::max(h->wprep_maxcount, 1), 1024), 592)), h->n_wprep);
        prof_begin(st, "wprep_dgrad", 8.0 * h->n_params, 0.0);
        wprep_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const WPrepEntry*>(h->wprep_tab), h->params, h->wt);
        prof_end(st);
        S2S_LAUNCH_CHECK();
    }
    if (h->tc_mode) {
        auto one = [&](const ConvL& L) -> int {
            if (!L.tc) return 0;
            const int total = 9 * L.Cin * L.Cout;
            prof_begin(st, "wprep_bf16", 6.0 * total, 0.0);
            wprep_bf16_kernel<<<std::min(cdiv(total, 256), 64), 256, 0, st>>>(h->params + L.w_off, h->wb + L.wb_off, L.Cin, L.Cout);
            prof_end(st);
            S2S_LAUNCH_CHECK();
            return 0;
        };
        for (int b = 0; b < h->nb; ++b) { S2S_CHECK(one(h->dconv[b][0])); S2S_CHECK(one(h->dconv[b][1])); S2S_CHECK(one(h->uconv[b][0])); S2S_CHECK(one(h->uconv[b][1])); }
        S2S_CHECK(one(h->bconv[0])); S2S_CHECK(one(h->bconv[1]));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// launch helpers bound to a handle
// ---------------------------------------------------------------------------------------
int run_conv_fwd(s2s_unet* h, const ConvL& L, const float* in, float* out, int N, const BnL* bn, bool training,
                 cudaStream_t st) {
    if (h->tc_mode && L.tc && !training) {
        // tensor-core inference: cast the input to bf16, then tcgen05 implicit GEMM (weights cast by run_wprep)
        const int64_t nx = (int64_t)N * L.H * L.W * L.Cin;
        prof_begin(st, "cast_bf16", 6.0 * nx, 0.0);
        cast_bf16_kernel<<<(unsigned)cdiv64(cdiv64(nx, 4), 256), 256, 0, st>>>(in, h->xb, nx);
        prof_end(st);
        S2S_LAUNCH_CHECK();
        TcConvArgs t;
        memset(&t, 0, sizeof t);
        t.bias = h->params + L.b_off; t.out = out; t.ldout = L.Cout;
        t.N = N; t.H = L.H; t.W = L.W; t.Cin = L.Cin; t.Cout = L.Cout; t.apply_elu = 1; t.act = h->cfg.act;
        return tcconv_launch(L.map_a, L.map_b, t, st);
    }
    if (h->t3_npass && L.t3f) {
        // tcgen05 tf32 implicit GEMM on the fp32 activations (training and inference)
        if (h->wq_pending) { S2S_CUDA(cudaStreamWaitEvent(st, h->ev_wq, 0)); h->wq_pending = false; }   // weight blocks of this step
        if (L.t3f_in != in) {
            S2S_CHECK(tc3_make_map_any(in, h->cfg.max_batch, L.H, L.W, L.Cin, L.Cin, L.pf.CK, &L.t3map_f));
            L.t3f_in = in;
        }
        Tc3Args t;
        memset(&t, 0, sizeof t);
        t.wq = h->wq + L.wqf_off; t.bias = h->params + L.b_off;
        t.out = out; t.ldout = L.Cout; t.in = in; t.ldin = L.Cin;
        t.N = N; t.H = L.H; t.W = L.W; t.Cin = L.Cin; t.Cout = L.Cout; t.epi = T3_EPI_BIAS_ACT; t.act = h->cfg.act;
        t.w_early = (&L != &h->dconv[0][0]) ? 1 : 0;       // wq is written by wprep (side stream, joined with a full event edge)
        t.early = (h->early_loads && (int64_t)N * L.H * L.W * std::max(L.Cin, L.Cout) <= h->bn_fold_max) ? 1 : 0;
        if (bn && bn->on && training) t.stat_part = h->stat_part;
        return tc3_launch(L.t3map_f, t, L.pf, h->t3_npass, 0, "conv3x3_fwd_tf32", st);
    }
    GConvArgs a;
    memset(&a, 0, sizeof a);
    a.in = in; a.ldin = L.Cin; a.in_coff = 0; a.Hin = L.H; a.Win = L.W; a.Cb = L.Cin;
    a.w = h->params + L.w_off; a.bias = h->params + L.b_off;
    a.out = out; a.ldout = L.Cout; a.out_coff = 0; a.Hout = L.H; a.Wout = L.W; a.Ca = L.Cout;
    a.pad = 1; a.epi = EPI_BIAS_ELU; a.N = N; a.act = h->cfg.act;
    a.w_early = (&L != &h->dconv[0][0]) ? 1 : 0;      // the first conv directly follows the previous step's Adam kernel
    a.early_loads = (h->early_loads && a.w_early) ? 1 : 0;      // (its bias too)
    if (bn && bn->on && training) a.stat_part = h->stat_part;     // finalised by the following bn_apply
    return gconv_run(3, 1, true, a, st);
}

// Backward statistics of a POOLED BatchNorm layer riding on the kernel that produces the gradient of its pooled output
// (run_conv_dgrad of the first conv of the next level): dc = g1 + unpool(dx), partials (sum dc, sum dc*xhat) per CTA.
struct BnFold {
    const BnL* bn = nullptr;          // the layer (full resolution 2H x 2W of this kernel's output)
    const float* act = nullptr;       // its input (ELU output), dense
    const float* g1 = nullptr; int ld1 = 0, coff1 = 0;    // skip-connection gradient at full resolution (nullable)
    int slots = 0;                    // out: partial slots written (0 = not folded, run bn_bwd_reduce)
};

// dx = dgrad(dz) [* ELU'(act)]; output may be a plain dense tensor
int run_conv_dgrad(s2s_unet* h, const ConvL& L, const float* dz, const float* act, float* dx, int N, cudaStream_t st,
                   BnFold* fold = nullptr) {
    if (fold) fold->slots = 0;
    if (h->t3_npass && L.t3d) {
        if (L.t3d_in != dz) {
            S2S_CHECK(tc3_make_map_any(dz, h->cfg.max_batch, L.H, L.W, L.Cout, L.Cout, L.pd.CK, &L.t3map_d));
            L.t3d_in = dz;
        }
        Tc3Args t;
        memset(&t, 0, sizeof t);
        t.wq = h->wq + L.wqd_off;
        t.out = dx; t.ldout = L.Cin; t.in = dz; t.ldin = L.Cout;
        t.N = N; t.H = L.H; t.W = L.W; t.Cin = L.Cout; t.Cout = L.Cin; t.act = h->cfg.act; t.w_early = 1;
        t.early = (h->early_loads && (int64_t)N * L.H * L.W * std::max(L.Cin, L.Cout) <= h->bn_fold_max) ? 1 : 0;
        if (act) { t.epi = T3_EPI_ACTGRAD; t.aux = act; t.ldaux = L.Cin; } else t.epi = T3_EPI_NONE;
        return tc3_launch(L.t3map_d, t, L.pd, h->t3_npass, 0, "conv3x3_dgrad_tf32", st);
    }
    GConvArgs a;
    memset(&a, 0, sizeof a);
    a.in = dz; a.ldin = L.Cout; a.Hin = L.H; a.Win = L.W; a.Cb = L.Cout;
    a.w = h->wt + L.w_off;                 // flipped + transposed by run_wprep
    a.out = dx; a.ldout = L.Cin; a.Hout = L.H; a.Wout = L.W; a.Ca = L.Cin;
    a.pad = 1; a.N = N;
    a.act = h->cfg.act; a.w_early = 1;                 // wt was prepared by wprep during the forward pass
    a.early_loads = h->early_loads ? 1 : 0;            // `act` (and the fold's operands) are forward activations / older gradients
    if (act) { a.epi = EPI_ELUGRAD; a.aux = act; a.ldaux = L.Cin; } else a.epi = EPI_NONE;
    if (h->bn_fold_pool && fold && fold->bn && fold->bn->on && fold->act && !act &&
        (int64_t)N * 4 * L.H * L.W * L.Cin <= h->bn_fold_max) {
        const GConvPlan p = gconv_plan(3, 1, L.H, L.W, L.Cin, L.Cout, N);
        const int slots = N * cdiv(L.H, p.th) * cdiv(L.W, p.tw);
        if (slots <= fold->bn->bwd_slots && fold->bn->C == L.Cin) {
            const BnL& bn = *fold->bn;
            a.stat_part = h->gpart + bn.part_off;
            a.stat_aux = fold->act; a.ldstat = bn.C;
            a.stat_mean = h->bn_mean + bn.ch_off; a.stat_rstd = h->bn_rstd + bn.ch_off;
            a.stat_pool = 1 + h->cfg.pool;
            a.stat_g1 = fold->g1; a.stat_ld1 = fold->ld1; a.stat_coff1 = fold->coff1;
            a.stat_scale = h->bn_scale + bn.ch_off; a.stat_shift = h->bn_shift + bn.ch_off;
            fold->slots = slots;
        }
    }
    return gconv_run(3, 1, true, a, st);
}

int run_conv_wgrad(s2s_unet* h, const ConvL& L, const float* x, int ldx, const float* dz, int N, cudaStream_t st) {
    if (L.twg) {
        // tcgen05 tf32 pixel-contraction GEMM (precision = tf32); the maps bake in the batch size (flat geometry)
        if (L.wg_x != x || L.wg_z != dz || L.wg_N != N) {
            S2S_CHECK(tcwg_make_maps(L.pw, x, ldx, dz, L.Cout, N, L.H, L.W, L.Cin, L.Cout, &L.wgmap_x, &L.wgmap_z));
            L.wg_x = x; L.wg_z = dz; L.wg_N = N;
        }
        return tcwg_launch(L.wgmap_x, L.wgmap_z, L.pw, h->gpart + L.part_off, h->gpart + L.bpart_off, N, L.H, L.W, L.Cin, L.Cout, L.nslots, st);
    }
    WgradArgs a;
    memset(&a, 0, sizeof a);
    a.A = dz; a.ldA = L.Cout; a.HA = L.H; a.WA = L.W; a.Ca = L.Cout;
    a.B = x; a.ldB = ldx; a.HB = L.H; a.WB = L.W; a.Cb = L.Cin;
    a.pad = 1; a.N = N;
    a.part = h->gpart + L.part_off;
    a.bias_part = h->gpart + L.bpart_off;
    return wgrad_run(3, 1, a, L.nslots, st);
}

int run_convt_fwd(s2s_unet* h, const ConvTL& L, const float* x, float* y, int ldy, int coff, int N, cudaStream_t st) {
    if (L.t3f) {
        if (h->wq_pending) { S2S_CUDA(cudaStreamWaitEvent(st, h->ev_wq, 0)); h->wq_pending = false; }
        if (L.t3f_in != x) {
            S2S_CHECK(tc3_make_map_any(x, h->cfg.max_batch, L.h, L.w, L.Cin, L.Cin, L.pf.CK, &L.t3maps_f.m[0]));
            for (int i = 1; i < 4; ++i) L.t3maps_f.m[i] = L.t3maps_f.m[0];
            L.t3f_in = x;
        }
        Tc3Args t;
        memset(&t, 0, sizeof t);
        t.wq = h->wq + L.wqf_off; t.bias = h->params + L.b_off;
        t.out = y; t.ldout = ldy; t.out_coff = coff; t.in = x; t.ldin = L.Cin;
        t.N = N; t.H = L.h; t.W = L.w; t.Cin = L.Cin; t.Cout = L.Cout; t.epi = T3_EPI_BIAS; t.act = h->cfg.act; t.w_early = 1;
        t.up = 1;
        for (int par = 0; par < 4; ++par) t.tapmask[par] = tc3_convt_tapmask(L.k, par, false);
        return tc3_launch_maps(L.t3maps_f, t, L.pf, h->t3_npass, 0, "convT_fwd_tf32", st);
    }
    ConvTArgs a;
    memset(&a, 0, sizeof a);
    a.x = x; a.ldx = L.Cin; a.h = L.h; a.w = L.w; a.Cin = L.Cin;
    a.wt = h->wt + L.w_off; a.bias = h->params + L.b_off;
    a.y = y; a.ldy = ldy; a.y_coff = coff; a.Cout = L.Cout; a.N = N;
    return convt_fwd(a, L.k, st);
}

// dx [N,h,w,Cin] from dy = channel slice of dcat
// bnst / bn_act / bn_slots (optional): dx is the gradient wrt the output of BatchNorm layer *bnst, whose input is bn_act — the
// CUDA-core kernel then leaves that layer's backward statistics (sum dc, sum dc*xhat per CTA) in its partial area and reports
// the number of slots it wrote, and the caller skips bn_bwd_reduce (one launch less on the gradient chain per such layer)
int run_convt_dgrad(s2s_unet* h, const ConvTL& L, const float* dy, int ldy, int coff, float* dx, int N, cudaStream_t st,
                    const BnL* bnst = nullptr, const float* bn_act = nullptr, int* bn_slots = nullptr) {
    if (bn_slots) *bn_slots = 0;
    if (L.t3d) {
        if (L.t3d_in != dy) {
            S2S_CHECK(tc3_make_maps_parity(dy + coff, h->cfg.max_batch, L.h, L.w, L.Cout, ldy, L.pd.CK, &L.t3maps_d));
            L.t3d_in = dy;
        }
        Tc3Args t;
        memset(&t, 0, sizeof t);
        t.wq = h->wq + L.wqd_off;
        t.out = dx; t.ldout = L.Cin; t.in = dy; t.ldin = ldy;
        t.N = N; t.H = L.h; t.W = L.w; t.Cin = 4 * L.Cout; t.Cout = L.Cin; t.epi = T3_EPI_NONE; t.act = h->cfg.act; t.w_early = 1;
        t.kpp = L.pd.kchunks / 4;
        for (int par = 0; par < 4; ++par) t.tapmask[par] = tc3_convt_tapmask(L.k, par, true);
        return tc3_launch_maps(L.t3maps_d, t, L.pd, h->t3_npass, 0, "convT_dgrad_tf32", st);
    }
    GConvArgs a;
    memset(&a, 0, sizeof a);
    a.in = dy; a.ldin = ldy; a.in_coff = coff; a.Hin = 2 * L.h; a.Win = 2 * L.w; a.Cb = L.Cout;
    a.w = h->params + L.w_off;
    a.out = dx; a.ldout = L.Cin; a.Hout = L.h; a.Wout = L.w; a.Ca = L.Cin;
    a.pad = (L.k - 2) / 2; a.epi = EPI_NONE; a.N = N;
    a.early_loads = h->early_loads ? 1 : 0;
    if (h->bn_fold && bnst && bnst->on && bn_act && bn_slots && (L.Cin & 3) == 0 && (int64_t)N * L.h * L.w * L.Cin <= h->bn_fold_max) {
        const GConvPlan p = gconv_plan(L.k, 2, L.h, L.w, L.Cin, L.Cout, N);
        const int slots = N * cdiv(L.h, p.th) * cdiv(L.w, p.tw);
        if (slots <= bnst->bwd_slots) {
            a.stat_part = h->gpart + bnst->part_off;
            a.stat_aux = bn_act; a.ldstat = bnst->C;
            a.stat_mean = h->bn_mean + bnst->ch_off; a.stat_rstd = h->bn_rstd + bnst->ch_off;
            *bn_slots = slots;
        }
    }
    if (L.k == 2) return gconv_run(2, 2, false, a, st);
    if (L.k == 3) return gconv_run(3, 2, false, a, st);
    return gconv_run(5, 2, false, a, st);
}

int run_convt_wgrad(s2s_unet* h, const ConvTL& L, const float* x, const float* dy, int ldy, int coff, int N, cudaStream_t st) {
    if (L.twg) {
        if (L.wg_x != x || L.wg_dy != dy || L.wg_N != N) {
            S2S_CHECK(tcwg_make_maps_convt(L.pw, x, L.Cin, dy + coff, ldy, N, L.h, L.w, L.Cin, L.Cout, &L.wgmaps_x, &L.wgmap_z));
            L.wg_x = x; L.wg_dy = dy; L.wg_N = N;
        }
        return tcwg_launch_maps(L.wgmaps_x, L.wgmap_z, L.pw, h->gpart + L.part_off, nullptr, N, L.h, L.w, L.Cout, L.Cin, L.nslots, L.k, st);
    }
    WgradArgs a;
    memset(&a, 0, sizeof a);
    a.A = x; a.ldA = L.Cin; a.HA = L.h; a.WA = L.w; a.Ca = L.Cin;
    a.B = dy; a.ldB = ldy; a.coffB = coff; a.HB = 2 * L.h; a.WB = 2 * L.w; a.Cb = L.Cout;
    a.pad = (L.k - 2) / 2; a.N = N;
    a.part = h->gpart + L.part_off;
    a.bias_part = nullptr;
    if (L.k == 2) return wgrad_run(2, 2, a, L.nslots, st);
    if (L.k == 3) return wgrad_run(3, 2, a, L.nslots, st);
    return wgrad_run(5, 2, a, L.nslots, st);
}

int run_bn_apply(s2s_unet* h, const BnL& bn, const ConvL& producer, const float* act, float* c_out, int ldc, int coffc,
                 float* p_out, int N, int hh, int ww, bool training, cudaStream_t st) {
    BnApplyArgs a;
    memset(&a, 0, sizeof a);
    a.sync_id = -1;
    a.early_loads = h->early_loads ? 1 : 0;
    a.a = act;
    a.scale = bn.on ? h->bn_scale + bn.ch_off : h->ones;
    a.shift = bn.on ? h->bn_shift + bn.ch_off : h->zeros;
    a.c_out = c_out; a.ldc = ldc; a.coffc = coffc; a.p_out = p_out;
    a.pool_kind = h->cfg.pool; a.N = N; a.h = hh; a.w = ww; a.C = bn.C;
    if (bn.on && training) {
        a.stat_part = h->stat_part;
        a.nslots = (h->t3_npass && producer.t3f) ? tc3_stat_slots(producer.H, producer.W, N) : gconv_stat_slots(producer.H, producer.W, N);
        a.gamma = h->params + bn.g_off; a.beta = h->params + bn.be_off;
        a.mov_mean = h->state + bn.mm_off; a.mov_var = h->state + bn.mv_off;
        a.bn_mean = h->bn_mean + bn.ch_off; a.bn_rstd = h->bn_rstd + bn.ch_off;
        a.bn_scale = h->bn_scale + bn.ch_off; a.bn_shift = h->bn_shift + bn.ch_off;
        a.eps = h->cfg.bn_eps; a.momentum = h->cfg.bn_momentum; a.update_moving = 1;
        if (h->dp && h->dp_sync_bn && h->dp_in_step) {      // global batch statistics: the sums are exchanged inside bn_apply
            S2S_REQUIRE(h->dp_sync_next < DP_MAXSYNC, "too many BN sync points");
            a.sync_id = h->dp_sync_next++;
            a.dp = h->dp->dev;
            a.M_total = (double)h->dp_n_global * hh * ww;
        }
    }
    return bn_apply(a, p_out != nullptr, st);
}

// BN backward (+ pool backward + skip add + ELU'):  dz = f(g1 + unpool(g2))
// producer_slots > 0: the kernel that produced g1 already left the [producer_slots][2][C] statistics partials (run_convt_dgrad)
int run_bn_bwd(s2s_unet* h, const BnL& bn, const float* act, const float* g1, int ld1, int coff1, const float* g2,
               float* dz, int N, int hh, int ww, bool batch_stats, bool elugrad, cudaStream_t st, int producer_slots = 0) {
    BnBwdArgs g;
    memset(&g, 0, sizeof g);
    g.sync_id = -1;
    g.early_loads = h->early_loads ? 1 : 0;
    g.act = act; g.g1 = g1; g.ld1 = ld1; g.coff1 = coff1; g.g2 = g2; g.pool_kind = h->cfg.pool;
    g.scale = bn.on ? h->bn_scale + bn.ch_off : h->ones;
    g.shift = bn.on ? h->bn_shift + bn.ch_off : h->zeros;
    g.mean = h->bn_mean + bn.ch_off; g.rstd = h->bn_rstd + bn.ch_off;
    g.part = h->gpart + bn.part_off; g.nslots = bn.bwd_slots;
    g.dz = dz; g.N = N; g.h = hh; g.w = ww; g.C = bn.C;
    g.apply_elugrad = elugrad ? 1 : 0; g.act_kind = h->cfg.act;
    g.batch_stats = (bn.on && batch_stats) ? 1 : 0;
    if (g.batch_stats) { g.dbeta = h->grads + bn.be_off; g.dgamma = h->grads + bn.g_off; }
    if (g.batch_stats && producer_slots == 0 && !(h->dp && h->dp_sync_bn && h->dp_in_step) && bn_bwd_fused_ok(g))
        return bn_bwd_fused(g, reinterpret_cast<GridBarrier*>(h->counters + 2), st);      // one launch: reduce | grid barrier | apply
    if (g.batch_stats) {
        if (producer_slots > 0) g.nslots = producer_slots;
        else S2S_CHECK(bn_bwd_reduce(g, st));
        if (h->dp && h->dp_sync_bn && h->dp_in_step) {
            S2S_REQUIRE(h->dp_sync_next < DP_MAXSYNC, "too many BN sync points");
            g.sync_id = h->dp_sync_next++;
            g.dp = h->dp->dev;
            g.M_total = (double)h->dp_n_global * hh * ww;
        }
    }
    return bn_bwd_apply(g, st);
}

const float* up_input(const s2s_unet* h, int b) { return (b == h->nb - 1) ? h->cb : h->uo[b + 1]; }
float* up_input_grad(s2s_unet* h, int b) { return (b == h->nb - 1) ? h->dcb : h->duo[b + 1]; }
// output of up block b (post-BN for b>0, raw ELU output for b==0 or bn off is still routed through uo when b>0)
const float* up_output(const s2s_unet* h, int b) { return b > 0 ? h->uo[b] : h->ua2[0]; }

int run_fold_bn(s2s_unet* h, cudaStream_t st) {
    if (h->nfold == 0) return 0;
    prof_begin(st, "bn_fold", 24.0 * h->n_bnch, 0.0);
    bn_fold_kernel<<<h->nfold, 128, 0, st>>>(h->fold_dev, h->params, h->state, h->bn_scale, h->bn_shift, h->cfg.bn_eps);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------
// forward: everything up to (excluding) the head
// ---------------------------------------------------------------------------------------
int run_forward_body(s2s_unet* h, int N, bool training, cudaStream_t st) {
    const int nb = h->nb;
    if (!training) S2S_CHECK(run_fold_bn(h, st));
    const float* cur = h->x_in;
    for (int b = 0; b < nb; ++b) {
        const int hh = levelH(h, b), ww = levelW(h, b), C = levelC(h, b);
        S2S_CHECK(run_conv_fwd(h, h->dconv[b][0], cur, h->a1[b], N, nullptr, training, st));
        S2S_CHECK(run_conv_fwd(h, h->dconv[b][1], h->a1[b], h->a2[b], N, &h->dbn[b], training, st));
        S2S_CHECK(run_bn_apply(h, h->dbn[b], h->dconv[b][1], h->a2[b], h->cat[b], 2 * C, 0, h->pl[b], N, hh, ww, training, st));
        cur = h->pl[b];
    }
    {
        const int hh = levelH(h, nb), ww = levelW(h, nb);
        S2S_CHECK(run_conv_fwd(h, h->bconv[0], cur, h->ab1, N, nullptr, training, st));
        S2S_CHECK(run_conv_fwd(h, h->bconv[1], h->ab1, h->ab2, N, &h->bbn, training, st));
        S2S_CHECK(run_bn_apply(h, h->bbn, h->bconv[1], h->ab2, h->cb, h->bbn.C, 0, nullptr, N, hh, ww, training, st));
    }
    if (h->wprep_pending) { S2S_CUDA(cudaStreamWaitEvent(st, h->ev_wprep, 0)); h->wprep_pending = false; }
    for (int b = nb - 1; b >= 0; --b) {
        const int hh = levelH(h, b), ww = levelW(h, b), C = levelC(h, b);
        S2S_CHECK(run_convt_fwd(h, h->upT[b], up_input(h, b), h->cat[b], 2 * C, C, N, st));
        S2S_CHECK(run_conv_fwd(h, h->uconv[b][0], h->cat[b], h->ua1[b], N, nullptr, training, st));
        S2S_CHECK(run_conv_fwd(h, h->uconv[b][1], h->ua1[b], h->ua2[b], N, b > 0 ? &h->ubn[b] : nullptr, training, st));
        if (b > 0) S2S_CHECK(run_bn_apply(h, h->ubn[b], h->uconv[b][1], h->ua2[b], h->uo[b], C, 0, nullptr, N, hh, ww, training, st));
    }
    return 0;
}

cudaStream_t side_after(s2s_unet* h, cudaStream_t st);

// defer_final: the sum of the per-CTA partials (head gradients, loss / accuracy, optimiser step counter) runs as its own
// small kernel on a side stream, off the gradient chain; the caller joins the side streams before Adam (run_backward does)
int run_head(s2s_unet* h, int N, float* probs, const float* y, const uint8_t* mask, float* dz_out, bool train,
             int cam_cls, cudaStream_t st, bool defer_final = false) {
    HeadArgs a;
    memset(&a, 0, sizeof a);
    a.u = h->ua2[0]; a.ldu = h->C0;
    a.wh = h->params + h->head_w; a.bh = h->params + h->head_b;
    a.y = y; a.mask = mask; a.hw = h->cfg.H * h->cfg.W; a.mask_norm = h->mask_norm_cache;
    a.probs = probs; a.dz_out = dz_out; a.apply_elugrad = 1; a.act = h->cfg.act;
    a.part = h->head_part; a.counter = h->counters + 0;
    a.dwh = train ? h->grads + h->head_w : nullptr; a.dbh = train ? h->grads + h->head_b : nullptr;
    a.stats = h->stats; a.stats_acc = h->stats_acc;
    a.hyper = train ? h->hyper : nullptr;
    a.gscale_dev = h->gscale; a.grad_scale = 1.f;
    a.npix = (int64_t)N * h->cfg.H * h->cfg.W;
    a.loss_kind = h->loss_kind; a.train = train ? 1 : 0;
    a.cam_cls = cam_cls; a.cam_norm = 1.f / (float)(h->cfg.H * h->cfg.W);
    a.defer_final = defer_final ? 1 : 0;
    a.early_loads = h->early_loads ? 1 : 0;
    S2S_CHECK(head_launch(a, h->C0, h->NC, st));
    if (defer_final) S2S_CHECK(head_final_launch(a, h->C0, h->NC, side_after(h, st)));
    return 0;
}

// ---------------------------------------------------------------------------------------
// backward: from dz_ua2[0] (written by the head kernel) to all parameter-gradient partials.
// Grad-CAM mode (cam != nullptr): inference-mode BN, no weight gradients; the walk stops at the
// named layer and reports where the gradient wrt that layer's OUTPUT (post-ELU) was left.
// ---------------------------------------------------------------------------------------
struct CamTarget {
    std::string layer;
    const float* grad = nullptr; int ld = 0;      // gradient wrt the layer output
    const float* act = nullptr; int lda = 0;      // the layer output itself
    int H = 0, W = 0, C = 0;
    bool found = false;
};

// fork: the returned stream has waited for everything enqueued on `st` so far
cudaStream_t side_after(s2s_unet* h, cudaStream_t st) {
    if (!h->use_side || h->ev_next >= s2s_unet::NEV) return st;
    cudaEvent_t e = h->ev[h->ev_next++];
    if (cudaEventRecord(e, st) != cudaSuccess) return st;
    const int k = h->side_next;
    h->side_next = (k + 1) % s2s_unet::NSIDE;
    if (cudaStreamWaitEvent(h->side[k], e, 0) != cudaSuccess) return st;
    h->side_used[k] = true;
    return h->side[k];
}
// join: `st` waits for all side work enqueued since the last join
int side_join(s2s_unet* h, cudaStream_t st) {
    for (int k = 0; k < s2s_unet::NSIDE; ++k) {
        if (!h->side_used[k]) continue;
        S2S_REQUIRE(h->ev_next < s2s_unet::NEV, "event pool exhausted");
        cudaEvent_t e = h->ev[h->ev_next++];
        S2S_CUDA(cudaEventRecord(e, h->side[k]));
        S2S_CUDA(cudaStreamWaitEvent(st, e, 0));
        h->side_used[k] = false;
    }
    h->ev_next = 0;
    return 0;
}

int run_backward(s2s_unet* h, int N, CamTarget* cam, cudaStream_t st) {
    const int nb = h->nb;
    const bool train = cam == nullptr;
    auto is = [&](const std::string& nm) { return cam && cam->layer == nm; };
    auto hit = [&](const float* g, int ld, const float* a, int lda, int hh, int ww, int C) {
        cam->grad = g; cam->ld = ld; cam->act = a; cam->lda = lda; cam->H = hh; cam->W = ww; cam->C = C; cam->found = true;
        return 0;
    };
    int bott_slots = 0;              // statistics partials the deepest transposed-conv input gradient left for the bottleneck BatchNorm
    BnFold down_fold;                // same for the BatchNorm of the down block whose pooled-output gradient was produced last
    auto make_down_fold = [&](int b) {
        BnFold f;
        f.bn = &h->dbn[b]; f.act = h->a2[b];
        f.g1 = h->dcat[b]; f.ld1 = 2 * levelC(h, b); f.coff1 = 0;
        return f;
    };
    for (int b = 0; b < nb; ++b) {   // up blocks, shallow -> deep
        const int C = levelC(h, b), hh = levelH(h, b), ww = levelW(h, b);
        const std::string n = "up_conv" + std::to_string(b + 1);
        const ConvL& c3 = h->uconv[b][1];
        const ConvL& c2 = h->uconv[b][0];
        if (is(n + "_3")) return hit(h->dz_ua2[b], C, h->ua2[b], C, hh, ww, C);
        if (train) S2S_CHECK(run_conv_wgrad(h, c3, h->ua1[b], C, h->dz_ua2[b], N, side_after(h, st)));
        S2S_CHECK(run_conv_dgrad(h, c3, h->dz_ua2[b], is(n + "_2") ? nullptr : h->ua1[b], h->dz_ua1[b], N, st));
        if (is(n + "_2")) return hit(h->dz_ua1[b], C, h->ua1[b], C, hh, ww, C);
        if (train) S2S_CHECK(run_conv_wgrad(h, c2, h->cat[b], 2 * C, h->dz_ua1[b], N, side_after(h, st)));
        S2S_CHECK(run_conv_dgrad(h, c2, h->dz_ua1[b], nullptr, h->dcat[b], N, st));
        if (is(n + "_1")) return hit(h->dcat[b] + C, 2 * C, h->cat[b] + C, 2 * C, hh, ww, C);
        const ConvTL& T = h->upT[b];
        if (train) {   // Conv2DTranspose bias gradient: channel sums of the upper half of dcat
            ChanSumArgs cs;
            memset(&cs, 0, sizeof cs);
            cs.g = h->dcat[b]; cs.ld = 2 * C; cs.coff = C; cs.C = C;
            cs.npix = (int64_t)N * hh * ww;
            cs.part = h->gpart + T.cs_part_off; cs.nslots = T.cs_slots;
            cudaStream_t ss = side_after(h, st);
            S2S_CHECK(chansum(cs, ss));
            S2S_CHECK(run_convt_wgrad(h, T, up_input(h, b), h->dcat[b], 2 * C, C, N, ss));
        }
        // the layer below ends in a BatchNorm whose output gradient this kernel produces: its backward statistics ride along
        const BnL& bn_below = (b + 1 < nb) ? h->ubn[b + 1] : h->bbn;
        const float* act_below = (b + 1 < nb) ? h->ua2[b + 1] : h->ab2;
        int pslots = 0;
        S2S_CHECK(run_convt_dgrad(h, T, h->dcat[b], 2 * C, C, up_input_grad(h, b), N, st, train ? &bn_below : nullptr, act_below, &pslots));
        if (b + 1 < nb) {
            const std::string n1 = "up_conv" + std::to_string(b + 2) + "_3";
            S2S_CHECK(run_bn_bwd(h, h->ubn[b + 1], h->ua2[b + 1], h->duo[b + 1], levelC(h, b + 1), 0, nullptr,
                                 h->dz_ua2[b + 1], N, levelH(h, b + 1), levelW(h, b + 1), train, !is(n1), st, pslots));
        } else {
            bott_slots = pslots;
        }
    }
    {   // bottleneck
        const int C = levelC(h, nb), hh = levelH(h, nb), ww = levelW(h, nb);
        S2S_CHECK(run_bn_bwd(h, h->bbn, h->ab2, h->dcb, C, 0, nullptr, h->dz_ab2, N, hh, ww, train, !is("conv2d"), st, bott_slots));
        if (is("conv2d")) return hit(h->dz_ab2, C, h->ab2, C, hh, ww, C);
        if (train) S2S_CHECK(run_conv_wgrad(h, h->bconv[1], h->ab1, C, h->dz_ab2, N, side_after(h, st)));
        S2S_CHECK(run_conv_dgrad(h, h->bconv[1], h->dz_ab2, is("bottleneck") ? nullptr : h->ab1, h->dz_ab1, N, st));
        if (is("bottleneck")) return hit(h->dz_ab1, C, h->ab1, C, hh, ww, C);
        if (train) S2S_CHECK(run_conv_wgrad(h, h->bconv[0], h->pl[nb - 1], h->bconv[0].Cin, h->dz_ab1, N, side_after(h, st)));
        // dpl is the gradient of the pooled output of down block nb-1's BatchNorm: its backward statistics ride along
        down_fold = make_down_fold(nb - 1);
        S2S_CHECK(run_conv_dgrad(h, h->bconv[0], h->dz_ab1, nullptr, h->dpl[nb - 1], N, st, train ? &down_fold : nullptr));
    }
    for (int b = nb - 1; b >= 0; --b) {   // down blocks, deep -> shallow
        const int C = levelC(h, b), hh = levelH(h, b), ww = levelW(h, b);
        const std::string n = "down_conv" + std::to_string(b + 1);
        S2S_CHECK(run_bn_bwd(h, h->dbn[b], h->a2[b], h->dcat[b], 2 * C, 0, h->dpl[b], h->dz_a2[b], N, hh, ww, train,
                             !is(n + "_2"), st, train ? down_fold.slots : 0));
        if (is(n + "_2")) return hit(h->dz_a2[b], C, h->a2[b], C, hh, ww, C);
        if (train) S2S_CHECK(run_conv_wgrad(h, h->dconv[b][1], h->a1[b], C, h->dz_a2[b], N, side_after(h, st)));
        S2S_CHECK(run_conv_dgrad(h, h->dconv[b][1], h->dz_a2[b], is(n + "_1") ? nullptr : h->a1[b], h->dz_a1[b], N, st));
        if (is(n + "_1")) return hit(h->dz_a1[b], C, h->a1[b], C, hh, ww, C);
        const float* xin = b > 0 ? h->pl[b - 1] : h->x_in;
        if (train) S2S_CHECK(run_conv_wgrad(h, h->dconv[b][0], xin, h->dconv[b][0].Cin, h->dz_a1[b], N, side_after(h, st)));
        if (b > 0) {
            down_fold = make_down_fold(b - 1);
            S2S_CHECK(run_conv_dgrad(h, h->dconv[b][0], h->dz_a1[b], nullptr, h->dpl[b - 1], N, st, train ? &down_fold : nullptr));
        }
    }
    S2S_CHECK(side_join(h, st));
    return 0;
}

// Grad-CAM combine: alpha_k = mean_hw dA[n,:,:,k]; cam[n,h,w] = relu(sum_k alpha_k A[n,h,w,k]).  One CTA per sample.
__global__ void __launch_bounds__(256) gradcam_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ act, int lda,
                                                      int HW, int C, float* __restrict__ cam) {
    extern __shared__ float s_alpha[];     // [C]
    __shared__ float s_red[8];
    const int n = blockIdx.x, tid = threadIdx.x;
    const float* gn = g + (size_t)n * HW * ldg;
    const float* an = act + (size_t)n * HW * lda;
    for (int c = 0; c < C; ++c) {
        float s = 0.f;
        for (int p = tid; p < HW; p += 256) s += gn[(size_t)p * ldg + c];
        s = warp_sum(s);
        if ((tid & 31) == 0) s_red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += s_red[w];
            s_alpha[c] = t / (float)HW;
        }
        __syncthreads();
    }
    for (int p = tid; p < HW; p += 256) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(s_alpha[c], an[(size_t)p * lda + c], s);
        cam[(size_t)n * HW + p] = fmaxf(s, 0.f);
    }
}

// dst[i, :] = src[idx ? idx[i] : i, :]   (mini-batch assembly from the device-resident dataset)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst,
                                   int64_t row, int n) {
    const int64_t total = (int64_t)n * row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / row);
        const int64_t c = i % row;
        const int64_t sr = idx ? idx[r] : r;
        dst[i] = __ldg(src + sr * row + c);
    }
}
int gather_rows(const float* src, const int* idx, float* dst, int64_t row, int n, cudaStream_t st) {
    const int64_t total = (int64_t)n * row;
    prof_begin(st, "gather_rows", 8.0 * total, 0.0);
    gather_rows_kernel<<<(unsigned)std::min<int64_t>(cdiv64(total, 256), 148 * 8), 256, 0, st>>>(src, idx, dst, row, n);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

// data-parallel step end: local slot reduction into the exchange buffer, then ONE kernel that all-reduces over
// peer memory (fixed rank order) and applies Adam (dp.cuh)
int run_grad_finish_dp(s2s_unet* h, int n_local, bool adam, cudaStream_t st) {
    s2s_dp* dp = h->dp;
    prof_begin(st, "dp_grad_reduce", 4.0 * ((double)h->gpart_floats + h->n_params), 0.0);
    dp_grad_reduce_kernel<<<h->nctas, 256, 0, st>>>(h->ctas_dev, h->blocks_dev, h->gpart, h->grads, dp->dev, h->stats, (float)n_local);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    const unsigned grid = (unsigned)cdiv64((int64_t)cdiv64((int64_t)dp->n_pad, 4), 256);
    prof_begin(st, "dp_allreduce_adam", 4.0 * h->n_params * (dp->world + 7.0), 0.0);
    dp_sum_adam_kernel<<<grid, 256, 0, st>>>(dp->dev, h->grads, h->params, h->m, h->v, h->hyper, h->stats_global, adam ? 1 : 0);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

int run_grad_finish(s2s_unet* h, bool adam, cudaStream_t st) {
    prof_begin(st, adam ? "grad_reduce_adam" : "grad_reduce", 4.0 * ((double)h->gpart_floats + (adam ? 7.0 : 1.0) * h->n_params), 0.0);
    if (adam)
        launch_k(grad_reduce_adam_kernel<true>, h->nctas, 256, 0, st, h->ctas_dev, h->blocks_dev, h->gpart, h->grads, h->params, h->m, h->v, h->hyper, h->early_loads ? 1 : 0);
    else
        launch_k(grad_reduce_adam_kernel<false>, h->nctas, 256, 0, st, h->ctas_dev, h->blocks_dev, h->gpart, h->grads, h->params, h->m, h->v, h->hyper, 0);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

// full sequences -------------------------------------------------------------------------
// The training sequence in two halves, so that the end-to-end entry point can overlap the H2D copy of the targets
// (first needed by the head kernel) with the forward pass: each half is its own CUDA graph there.
int seq_train_fwd(s2s_unet* h, int N, cudaStream_t st) {
    h->ev_next = 0;
    h->dp_sync_next = 0;
    {   // dgrad weight preparation overlaps the forward pass on a side stream
        cudaStream_t ss = side_after(h, st);
        h->wq_record = ss != st;
        const int wrc = run_wprep(h, ss);
        h->wq_record = false;
        S2S_CHECK(wrc);
        if (ss != st) {
            S2S_CUDA(cudaEventRecord(h->ev_wprep, ss));
            h->wprep_pending = true;      // joined by the forward pass before its first transposed conv
        }
    }
    S2S_CHECK(run_forward_body(h, N, true, st));
    if (h->wprep_pending) { S2S_CUDA(cudaStreamWaitEvent(st, h->ev_wprep, 0)); h->wprep_pending = false; }
    if (h->wq_pending) { S2S_CUDA(cudaStreamWaitEvent(st, h->ev_wq, 0)); h->wq_pending = false; }
    return 0;
}
int seq_train_bwd(s2s_unet* h, int N, bool adam, const uint8_t* mask, cudaStream_t st, bool dp = false) {
    S2S_CHECK(run_head(h, N, nullptr, h->y_in, mask, h->dz_ua2[0], true, -1, st, h->use_side));
    S2S_CHECK(run_backward(h, N, nullptr, st));
    if (dp) S2S_CHECK(run_grad_finish_dp(h, N, adam, st));
    else S2S_CHECK(run_grad_finish(h, adam, st));
    return 0;
}
int seq_train(s2s_unet* h, int N, bool adam, const uint8_t* mask, cudaStream_t st, bool dp = false) {
    S2S_CHECK(seq_train_fwd(h, N, st));
    return seq_train_bwd(h, N, adam, mask, st, dp);
}
int seq_eval(s2s_unet* h, int N, const uint8_t* mask, cudaStream_t st) {
    S2S_CHECK(run_wprep(h, st));
    S2S_CHECK(run_forward_body(h, N, false, st));
    S2S_CHECK(run_head(h, N, nullptr, h->y_in, mask, nullptr, false, -1, st));
    return 0;
}
int seq_forward(s2s_unet* h, int N, bool training, cudaStream_t st) {
    S2S_CHECK(run_wprep(h, st));
    S2S_CHECK(run_forward_body(h, N, training, st));
    S2S_CHECK(run_head(h, N, h->probs, nullptr, nullptr, nullptr, false, -1, st));
    return 0;
}

// Run a sequence either eagerly or through a cached CUDA graph (captured on first use).
template <typename F>
int run_cached(s2s_unet* h, int kind, int N, cudaStream_t st, F&& body) {
    const bool graphable = h->use_graphs && st != nullptr && !prof().on;
    if (!graphable) {
        const int64_t before = launch_counter();
        S2S_CHECK(body(st));
        h->launches += launch_counter() - before;
        return 0;
    }
    const long long key = ((long long)kind << 40) | (long long)N;      // N may be a composite (n_local << 16 | n_global)
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        const int64_t before = launch_counter();
        cudaGraph_t g = nullptr;
        S2S_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = body(st);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(S2S_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
        cudaGraphExec_t ex = nullptr;
        e = cudaGraphInstantiate(&ex, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(S2S_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
        const int nk = (int)(launch_counter() - before);
        it = h->graphs.emplace(key, std::make_pair(ex, nk)).first;
    }
    S2S_CUDA(cudaGraphLaunch(it->second.first, st));
    h->launches += it->second.second;
    return 0;
}

int stage_inputs(s2s_unet* h, const float* x, const float* y, int N, cudaStream_t st) {
    const size_t xb = (size_t)N * h->cfg.H * h->cfg.W * h->cfg.Cin * sizeof(float);
    const size_t yb = (size_t)N * h->cfg.H * h->cfg.W * h->NC * sizeof(float);
    if (x && x != h->x_in) S2S_CUDA(cudaMemcpyAsync(h->x_in, x, xb, cudaMemcpyDeviceToDevice, st));
    if (y && y != h->y_in) S2S_CUDA(cudaMemcpyAsync(h->y_in, y, yb, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int set_gscale(s2s_unet* h, float gs, cudaStream_t st) {
    if (gs != h->gscale_host) {
        h->gscale_host = gs;
        S2S_CUDA(cudaMemcpyAsync(h->gscale, &h->gscale_host, sizeof(float), cudaMemcpyHostToDevice, st));
    }
    return 0;
}

int check_N(const s2s_unet* h, int N) {
    S2S_REQUIRE(h != nullptr, "null handle");
    S2S_REQUIRE(N >= 1 && N <= h->cfg.max_batch, "batch %d outside [1, max_batch=%d]", N, h->cfg.max_batch);
    return 0;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int s2s_version(void) { return S2S_ABI_VERSION; }
const char* s2s_last_error(void) { return last_error_ref().c_str(); }

int s2s_device_count(int* n) { S2S_REQUIRE(n, "null"); S2S_CUDA(cudaGetDeviceCount(n)); return 0; }
int s2s_set_device(int dev) { S2S_CUDA(cudaSetDevice(dev)); return 0; }
int s2s_get_device(int* dev) { S2S_REQUIRE(dev, "null"); S2S_CUDA(cudaGetDevice(dev)); return 0; }
int s2s_stream_create(void** stream) {
    S2S_REQUIRE(stream, "null");
    cudaStream_t s;
    S2S_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return 0;
}
int s2s_stream_destroy(void* stream) { S2S_CUDA(cudaStreamDestroy((cudaStream_t)stream)); return 0; }
int s2s_stream_sync(void* stream) { S2S_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return 0; }
// Host-layer device buffers (runtime.DeviceBuffer: the data sets fit / predict upload, the prediction outputs): cudaFree is a
// device-wide synchronisation plus an unmap that costs milliseconds in a process holding many graphs and large arenas (measured:
// predict at batch 32 fell from 61k to 6k samples/s inside bench.py), and the tuning loops allocate the same sizes over and
// over.  Freed blocks are kept per device in size classes (1/8-octave steps) up to S2S_DEV_CACHE_MB (default 2048); like
// cudaFree, s2s_dev_free returns only after the device has drained, so a block is never handed out while work still reads it.
namespace {
struct DevBlockCache {
    std::mutex mu;
    std::map<void*, std::pair<size_t, int>> live;                     // ptr -> (class size, device)
    std::map<std::pair<int, size_t>, std::vector<void*>> idle;        // (device, class size) -> blocks
    size_t cached = 0;
};
DevBlockCache& dev_cache() { static DevBlockCache c; return c; }
size_t dev_class(size_t b) {
    if (b < 512) return 512;
    size_t p = 512;
    while (p < b) p <<= 1;
    const size_t step = p >> 4;                                       // classes at 1/8 steps of the octave below p
    return (b + step - 1) / step * step;
}
size_t dev_cache_cap() {
    static const size_t cap = [] { const char* e = getenv("S2S_DEV_CACHE_MB"); return (size_t)(e ? atoll(e) : 2048) << 20; }();
    return cap;
}
void dev_cache_flush_locked(DevBlockCache& c) {
    for (auto& kv : c.idle)
        for (void* q : kv.second) cudaFree(q);
    c.idle.clear();
    c.cached = 0;
}
}  // namespace
int s2s_dev_alloc(void** p, size_t bytes) {
    S2S_REQUIRE(p, "null");
    const size_t sz = dev_class(bytes ? bytes : 1);
    int dev = 0;
    S2S_CUDA(cudaGetDevice(&dev));
    DevBlockCache& c = dev_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.idle.find({dev, sz});
    if (it != c.idle.end() && !it->second.empty()) {
        *p = it->second.back();
        it->second.pop_back();
        c.cached -= sz;
    } else {
        cudaError_t e = cudaMalloc(p, sz);
        if (e == cudaErrorMemoryAllocation) {                         // give the cached blocks back and retry once
            cudaGetLastError();
            dev_cache_flush_locked(c);
            e = cudaMalloc(p, sz);
        }
        if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? S2S_ERR_NOMEM : S2S_ERR_CUDA, "cudaMalloc(%zu bytes): %s", sz, cudaGetErrorString(e));
    }
    c.live[*p] = {sz, dev};
    return 0;
}
int s2s_dev_free(void* p) {
    if (!p) return 0;
    DevBlockCache& c = dev_cache();
    size_t sz = 0;
    int dev = -1;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.live.find(p);
        if (it != c.live.end()) { sz = it->second.first; dev = it->second.second; c.live.erase(it); }
    }
    if (dev < 0 || sz > dev_cache_cap() / 2) { S2S_CUDA(cudaFree(p)); return 0; }
    // No device-wide synchronisation here: the tuning loops run several trials on their own threads and streams, and a
    // cudaDeviceSynchronize from one thread invalidates another thread's stream capture (seen in tools/sweep_demo.py: "operation
    // failed due to a previous error during capture").  Contract instead (include/s2s_unet.h): the caller has synchronised the
    // stream(s) that used the block — the host layer does (DeviceBuffer.free synchronises the stream the buffer was last used on).
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.cached + sz > dev_cache_cap()) dev_cache_flush_locked(c);
    c.idle[{dev, sz}].push_back(p);
    c.cached += sz;
    return 0;
}
int s2s_host_alloc(void** p, size_t bytes) { S2S_REQUIRE(p, "null"); S2S_CUDA(cudaMallocHost(p, bytes ? bytes : 1)); return 0; }
int s2s_host_free(void* p) { S2S_CUDA(cudaFreeHost(p)); return 0; }
int s2s_memcpy_h2d(void* d, const void* s, size_t b, void* st) { S2S_CUDA(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, (cudaStream_t)st)); return 0; }
int s2s_memcpy_d2h(void* d, const void* s, size_t b, void* st) { S2S_CUDA(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, (cudaStream_t)st)); return 0; }
int s2s_memcpy_d2d(void* d, const void* s, size_t b, void* st) { S2S_CUDA(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, (cudaStream_t)st)); return 0; }
int s2s_memset_dev(void* d, int byte, size_t b, void* st) { S2S_CUDA(cudaMemsetAsync(d, byte, b, (cudaStream_t)st)); return 0; }
int s2s_event_create(void** ev) { S2S_REQUIRE(ev, "null"); cudaEvent_t e; S2S_CUDA(cudaEventCreate(&e)); *ev = (void*)e; return 0; }
int s2s_event_destroy(void* ev) { S2S_CUDA(cudaEventDestroy((cudaEvent_t)ev)); return 0; }
int s2s_event_record(void* ev, void* st) { S2S_CUDA(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)st)); return 0; }
int s2s_event_elapsed_ms(void* a, void* b, float* ms) {
    S2S_REQUIRE(ms, "null");
    S2S_CUDA(cudaEventSynchronize((cudaEvent_t)b));
    S2S_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return 0;
}
int s2s_l2_flush(void* scratch, size_t bytes, void* st) {
    S2S_CUDA(cudaMemsetAsync(scratch, 0, bytes, (cudaStream_t)st));
    return 0;
}

// ---- per-launch profiler -----------------------------------------------------------------
int s2s_prof_enable(int on) {
    Profiler& p = prof();
    for (ProfRec& r : p.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    p.recs.clear();
    p.on = on != 0;
    return 0;
}
// Event-bracket overhead calibration: mean microseconds measured by the profiler's own bracket around an
// EMPTY kernel (its true in-stream cost is ~1 us on B200, tools/graph_floor.cu); bench.py subtracts the excess.
__global__ void prof_null_kernel() {}
int s2s_prof_null_us(float* us_out, void* stream) {
    S2S_REQUIRE(us_out, "null");
    cudaStream_t st = (cudaStream_t)stream;
    const int R = 64;
    std::vector<cudaEvent_t> ev(2 * R);
    for (auto& e : ev) S2S_CUDA(cudaEventCreate(&e));
    for (int i = 0; i < R; ++i) {
        cudaEventRecord(ev[2 * i], st);
        prof_null_kernel<<<1, 32, 0, st>>>();
        cudaEventRecord(ev[2 * i + 1], st);
    }
    S2S_CUDA(cudaStreamSynchronize(st));
    double tot = 0;
    for (int i = 8; i < R; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]); tot += ms; }
    for (auto& e : ev) cudaEventDestroy(e);
    *us_out = (float)(tot * 1000.0 / (R - 8));
    return 0;
}

// Aggregates the records per tag into "tag,launches,total_ms,bytes,flops\n" lines.
int s2s_prof_report(char* buf, size_t buflen) {
    S2S_REQUIRE(buf && buflen > 0, "null buffer");
    S2S_CUDA(cudaDeviceSynchronize());
    struct Agg { int n = 0; double ms = 0, bytes = 0, flops = 0; };
    std::map<std::string, Agg> agg;
    std::vector<std::string> order;
    for (ProfRec& r : prof().recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
        if (!agg.count(r.tag)) order.push_back(r.tag);
        Agg& a = agg[r.tag];
        a.n++; a.ms += ms; a.bytes += r.bytes; a.flops += r.flops;
    }
    std::string out;
    char line[256];
    for (const std::string& t : order) {
        const Agg& a = agg[t];
        snprintf(line, sizeof line, "%s,%d,%.6f,%.0f,%.0f\n", t.c_str(), a.n, a.ms, a.bytes, a.flops);
        out += line;
    }
    S2S_REQUIRE(out.size() + 1 <= buflen, "report needs %zu bytes", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

// ---------------------------------------------------------------------------------------
int s2s_unet_create(const s2s_unet_cfg* cfg, s2s_unet** out) {
    S2S_REQUIRE(cfg && out, "null argument");
    // the reference graph always builds three down / up blocks and adds the 4th / 5th on request (deep_nn_models.py:82-86)
    S2S_REQUIRE(cfg->n_blocks >= 3 && cfg->n_blocks <= MAXB, "n_blocks must be in [3,%d] (got %d)", MAXB, cfg->n_blocks);
    S2S_REQUIRE(cfg->filters >= 1 && cfg->filters <= 4, "filters must be in [1,4] (got %d)", cfg->filters);
    S2S_REQUIRE(cfg->ct_kernel == 2 || cfg->ct_kernel == 3 || cfg->ct_kernel == 5, "ct_kernel must be 2, 3 or 5 (got %d)", cfg->ct_kernel);
    S2S_REQUIRE(cfg->H > 0 && cfg->W > 0 && cfg->Cin > 0 && cfg->max_batch > 0, "bad shape");
    const int div = 1 << cfg->n_blocks;
    // Keras raises on the Concatenate shape mismatch when H or W is not divisible by 2^n_blocks
    // (comment at tune_ECMWF_com.py:26); fail with a clear message instead.
    S2S_REQUIRE(cfg->H % div == 0 && cfg->W % div == 0,
                "input %dx%d is not divisible by 2^n_blocks=%d: the skip concatenation shapes would not match",
                cfg->H, cfg->W, div);
    S2S_REQUIRE(cfg->head == S2S_HEAD_SOFTMAX3 || cfg->head == S2S_HEAD_RELU1, "bad head kind");
    S2S_REQUIRE(cfg->act == S2S_ACT_ELU || cfg->act == S2S_ACT_RELU, "bad activation kind %d", cfg->act);
    S2S_REQUIRE(cfg->precision == S2S_PREC_FP32 || cfg->precision == S2S_PREC_BF16_TC || cfg->precision == S2S_PREC_TF32,
                "bad precision %d", cfg->precision);

    s2s_unet* h = new s2s_unet();
    h->cfg = *cfg;
    if (h->cfg.bn_eps <= 0.f) h->cfg.bn_eps = 1e-3f;
    if (h->cfg.bn_momentum <= 0.f) h->cfg.bn_momentum = 0.99f;
    h->nb = cfg->n_blocks;
    h->NC = cfg->head == S2S_HEAD_SOFTMAX3 ? 3 : 1;
    h->C0 = cfg->filters * 4;
    const int nb = h->nb;
    const bool bn = cfg->bn != 0;

    // ---- layer table in Keras creation order (deep_nn_models.py:82-103)
    int bn_index = 0;
    int cin = cfg->Cin;
    for (int b = 0; b < nb; ++b) {
        const std::string n = std::to_string(b + 1);
        def_conv(h, h->dconv[b][0], "down_conv" + n + "_1", cin, levelC(h, b), levelH(h, b), levelW(h, b));
        def_conv(h, h->dconv[b][1], "down_conv" + n + "_2", levelC(h, b), levelC(h, b), levelH(h, b), levelW(h, b));
        def_bn(h, h->dbn[b], bn_index, levelC(h, b), bn);
        cin = levelC(h, b);
    }
    def_conv(h, h->bconv[0], "bottleneck", cin, levelC(h, nb), levelH(h, nb), levelW(h, nb));
    def_conv(h, h->bconv[1], "conv2d", levelC(h, nb), levelC(h, nb), levelH(h, nb), levelW(h, nb));
    def_bn(h, h->bbn, bn_index, levelC(h, nb), bn);
    for (int b = nb - 1; b >= 0; --b) {
        const std::string n = std::to_string(b + 1);
        const int C = levelC(h, b);
        def_convt(h, h->upT[b], "up_conv" + n + "_1", 2 * C, C, levelH(h, b + 1), levelW(h, b + 1), cfg->ct_kernel);
        def_conv(h, h->uconv[b][0], "up_conv" + n + "_2", 2 * C, C, levelH(h, b), levelW(h, b));
        def_conv(h, h->uconv[b][1], "up_conv" + n + "_3", C, C, levelH(h, b), levelW(h, b));
        def_bn(h, h->ubn[b], bn_index, C, bn && b > 0);   // no normalisation directly before softmax (:99)
    }
    h->head_w = take_param(h, (int64_t)h->C0 * h->NC);
    h->head_b = take_param(h, h->NC);
    {
        const int ks[4] = {1, 1, h->C0, h->NC};
        add_desc(h, "conv2d_1/kernel", 0, 4, ks, h->head_w, (int64_t)h->C0 * h->NC);
        const int bs[1] = {h->NC};
        add_desc(h, "conv2d_1/bias", 0, 1, bs, h->head_b, h->NC);
    }

    // ---- workspaces
    const int NB = cfg->max_batch;
    std::vector<GradBlock> blocks;
    std::vector<WPrepEntry> wprep;
    size_t gpart_floats = 0;
    const bool tf32_mode = cfg->precision == S2S_PREC_TF32;
    // fp32 precision: the THICK layers (Cin * Cout >= 96 * 96 — every one of them sits on an 8x8 or smaller grid of the deep nets,
    // where the FFMA kernels are weight-streaming bound at 2-4 TFLOP/s) run the same tensor-core kernels with the error-compensated
    // 3xTF32 split: fp32 parity (forward rel-L2 ~5e-7, weight gradients ~9e-7 against the fp64 oracle; bars 1e-5 / 2e-4).
    // S2S_TC3_FP32=0 turns that off, =1 also takes every >= 16-row layer (bring-up).  The reference's default net has no such layer.
    const char* e3 = getenv("S2S_TC3_FP32");
    const int fp32_policy = cfg->precision == S2S_PREC_FP32 ? (e3 ? (e3[0] == '0' ? 0 : 2) : 1) : 0;
    auto thick = [](int ci, int co) { return (int64_t)ci * co >= 9216; };
    auto plan_conv = [&](ConvL& L) {
        WgradPlan p = wgrad_plan(L.H, L.W, L.Cout, L.Cin, NB);
        if ((tf32_mode && tcwg_wanted(L.H, L.W, L.Cin, L.Cout, NB)) || (fp32_policy && thick(L.Cin, L.Cout))) {
            const TcWgPlan pw = tcwg_plan(L.H, L.W, L.Cin, L.Cout, NB, 0, 1, tf32_mode ? 1 : 3);
            if (pw.ok) { L.twg = true; L.pw = pw; p.nslots = pw.nslots; }
        }
        L.nslots = p.nslots;
        const int64_t P = (int64_t)9 * L.Cin * L.Cout;
        if (L.Cin % 4 == 0) {   // layers that need a dgrad (everything but the input layer)
            wprep.push_back(WPrepEntry{L.w_off, L.Cout, L.Cin, 9, 1});      // dst [tap][co][ci]
            h->wprep_maxcount = std::max(h->wprep_maxcount, (int)P);
        }
        L.part_off = (int64_t)gpart_floats; gpart_floats += (size_t)p.nslots * P;
        L.bpart_off = (int64_t)gpart_floats; gpart_floats += (size_t)p.nslots * L.Cout;
        for (int64_t o = 0; o < P; o += GRAD_BLK)
            blocks.push_back(GradBlock{L.w_off + o, (int32_t)std::min<int64_t>(GRAD_BLK, P - o), p.nslots, L.part_off + o, P});
        for (int64_t o = 0; o < L.Cout; o += GRAD_BLK)
            blocks.push_back(GradBlock{L.b_off + o, (int32_t)std::min<int64_t>(GRAD_BLK, L.Cout - o), p.nslots, L.bpart_off + o, (int64_t)L.Cout});
    };
    auto plan_direct = [&](int64_t off, int64_t count) {
        for (int64_t o = 0; o < count; o += GRAD_BLK)
            blocks.push_back(GradBlock{off + o, (int32_t)std::min<int64_t>(GRAD_BLK, count - o), 0, 0, 0});
    };
    int n_counters = 4;   // counter 0: head; 2-3: grid barrier of the fused BatchNorm backward (bn.cuh)
    size_t stat_floats = 0;
    auto plan_bn = [&](BnL& B, const ConvL& producer, bool pooled) {
        if (!B.on) return;
        const int slots = gconv_stat_slots_max(producer.H, producer.W, NB);
        stat_floats = std::max(stat_floats, (size_t)slots * 2 * B.C);
        // backward partials [bwd_slots][2][C]: row 0 = sum dc (d beta), row 1 = sum dc*xhat (d gamma)
        const int64_t units = (int64_t)NB * producer.H * producer.W / (pooled ? 4 : 1);
        B.bwd_slots = bn_bwd_slots(units, 256 / bn_cqb(B.C));
        // un-pooled layers: the transposed-conv input gradient that produces dc may write the partials itself (one per CTA)
        if (!pooled) B.bwd_slots = std::max(B.bwd_slots, slots);
        else B.bwd_slots = std::max(B.bwd_slots, gconv_stat_slots_max(producer.H / 2, producer.W / 2, NB));   // ... or the 3x3 input gradient of the next level
        B.part_off = (int64_t)gpart_floats; gpart_floats += (size_t)B.bwd_slots * 2 * B.C;
        // d beta / d gamma: written densely into the gradient arena by bn_bwd_apply (CTA 0 sums the partials anyway)
        plan_direct(B.be_off, B.C);
        plan_direct(B.g_off, B.C);
    };
    for (int b = 0; b < nb; ++b) {
        plan_conv(h->dconv[b][0]); plan_conv(h->dconv[b][1]); plan_bn(h->dbn[b], h->dconv[b][1], true);
    }
    plan_conv(h->bconv[0]); plan_conv(h->bconv[1]); plan_bn(h->bbn, h->bconv[1], false);
    for (int b = nb - 1; b >= 0; --b) {
        ConvTL& T = h->upT[b];
        WgradPlan p = wgrad_plan(T.h, T.w, T.Cin, T.Cout, NB);
        if ((tf32_mode && tcwg_wanted_convt(T.h, T.w, T.Cin, T.Cout, T.k, NB)) || (fp32_policy && (int64_t)T.k * T.k * T.Cin * T.Cout >= 75000)) {
            const TcWgPlan pw = tcwg_plan(T.h, T.w, T.Cout, T.Cin, NB, 0, 4, tf32_mode ? 1 : 3);
            if (pw.ok) { T.twg = true; T.pw = pw; p.nslots = pw.nslots; }
        }
        T.nslots = p.nslots;
        const int64_t P = (int64_t)T.k * T.k * T.Cin * T.Cout;
        T.part_off = (int64_t)gpart_floats; gpart_floats += (size_t)p.nslots * P;
        for (int64_t o = 0; o < P; o += GRAD_BLK)
            blocks.push_back(GradBlock{T.w_off + o, (int32_t)std::min<int64_t>(GRAD_BLK, P - o), p.nslots, T.part_off + o, P});
        wprep.push_back(WPrepEntry{T.w_off, T.Cin, T.Cout, T.k * T.k, 0});   // dst [tap][ci][co]
        h->wprep_maxcount = std::max(h->wprep_maxcount, (int)P);
        T.cs_slots = bn_bwd_slots((int64_t)NB * 4 * T.h * T.w, 256 / bn_cqb(T.Cout));
        T.cs_part_off = (int64_t)gpart_floats; gpart_floats += (size_t)T.cs_slots * T.Cout;
        for (int64_t o = 0; o < T.Cout; o += GRAD_BLK)
            blocks.push_back(GradBlock{T.b_off + o, (int32_t)std::min<int64_t>(GRAD_BLK, T.Cout - o), T.cs_slots, T.cs_part_off + o, (int64_t)T.Cout});
        plan_conv(h->uconv[b][0]); plan_conv(h->uconv[b][1]); plan_bn(h->ubn[b], h->uconv[b][1], false);
    }
    plan_direct(h->head_w, (int64_t)h->C0 * h->NC);
    plan_direct(h->head_b, h->NC);
    h->nblocks = (int)blocks.size();
    h->gpart_floats = gpart_floats;
    // CTA table of the reduce + Adam kernel: consecutive blocks with the same warp count share a CTA (optim.cuh)
    std::vector<GradCta> gctas;
    for (int i = 0; i < (int)blocks.size();) {
        const int W = grad_block_warps(blocks[i].nslots);
        int n = 1;
        while (n < 8 / W && i + n < (int)blocks.size() && grad_block_warps(blocks[i + n].nslots) == W) ++n;
        gctas.push_back(GradCta{i, n, W, 0});
        i += n;
    }
    h->nctas = (int)gctas.size();
    // tensor-core inference layers
    h->tc_mode = cfg->precision == S2S_PREC_BF16_TC;
    size_t xb_elems = 0, wb_elems = 0;
    std::vector<ConvL*> tc_layers;
    if (h->tc_mode) {
        auto consider = [&](ConvL& L) {
            if (!tcconv_eligible(L.Cin, L.Cout)) return;
            L.tc = true;
            L.wb_off = (int64_t)wb_elems; wb_elems += ((size_t)9 * L.Cin * L.Cout + 63) / 64 * 64;
            xb_elems = std::max(xb_elems, (size_t)NB * L.H * L.W * L.Cin);
            tc_layers.push_back(&L);
        };
        for (int b = 0; b < nb; ++b) { consider(h->dconv[b][0]); consider(h->dconv[b][1]); consider(h->uconv[b][0]); consider(h->uconv[b][1]); }
        consider(h->bconv[0]); consider(h->bconv[1]);
    }
    h->n_counters = n_counters;
    // tcgen05 tf32 path: forward + dgrad plans of every eligible 3x3 layer
    h->t3_npass = tf32_mode ? 1 : (fp32_policy ? 3 : 0);
    size_t wq_floats = 0;
    std::vector<Tc3WPrep> t3prep;
    if (h->t3_npass) {
        const int minH = h->t3_npass == 1 ? 8 : 16;      // the 3-pass (fp32 parity) variant only pays on >= 16-row grids
        auto consider3 = [&](ConvL& L, bool need_dgrad) {
            // small images (H*W < 64) run the flat geometry in the single-pass mode: there the FFMA kernel is weight-streaming
            // bound at 2-3 TFLOP/s (grid_max: 64 us per layer)
            // (3-pass / fp32: the thick layers only, flat or tiled)
            if (h->t3_npass == 3 && fp32_policy == 1 && !thick(L.Cin, L.Cout)) return;
            const bool flat = tc3_flat(L.H, L.W) && (h->t3_npass == 1 || thick(L.Cin, L.Cout));
            if (!flat && !(h->t3_npass == 3 && thick(L.Cin, L.Cout)) && (L.H < minH || L.W < 8)) return;
            const Tc3Plan pf = tc3_plan_for(L.H, L.W, NB, L.Cin, L.Cout, h->t3_npass);
            if (pf.ok) {
                L.pf = pf; L.t3f = true; L.wqf_off = (int64_t)wq_floats; wq_floats += pf.wq_floats;
                t3prep.push_back(Tc3WPrep{L.w_off, L.wqf_off, L.Cin, L.Cout, 0, pf.NT, pf.nchunks_n, pf.CK, pf.kchunks, 0, h->t3_npass});
                h->t3prep_maxcount = std::max(h->t3prep_maxcount, pf.nchunks_n * pf.kchunks * 9 * pf.CK * pf.NT);
                stat_floats = std::max(stat_floats, (size_t)tc3_stat_slots(L.H, L.W, NB) * 2 * L.Cout);
            }
            const Tc3Plan pd = need_dgrad ? tc3_plan_for(L.H, L.W, NB, L.Cout, L.Cin, h->t3_npass) : Tc3Plan{};
            if (pd.ok) {
                L.pd = pd; L.t3d = true; L.wqd_off = (int64_t)wq_floats; wq_floats += pd.wq_floats;
                t3prep.push_back(Tc3WPrep{L.w_off, L.wqd_off, L.Cout, L.Cin, 0, pd.NT, pd.nchunks_n, pd.CK, pd.kchunks, 1, h->t3_npass});
                h->t3prep_maxcount = std::max(h->t3prep_maxcount, pd.nchunks_n * pd.kchunks * 9 * pd.CK * pd.NT);
            }
        };
        for (int b = 0; b < nb; ++b) {
            consider3(h->dconv[b][0], b > 0); consider3(h->dconv[b][1], true);
            consider3(h->uconv[b][0], true); consider3(h->uconv[b][1], true);
        }
        consider3(h->bconv[0], true); consider3(h->bconv[1], true);
        // Conv2DTranspose layers: forward as four parity convs, input gradient over four parity planes
        static const bool convt_tc = [] { const char* e = getenv("S2S_TC3_CONVT"); return !e || e[0] != '0'; }();
        if (convt_tc) {
            const int np = h->t3_npass;
            for (int b = 0; b < nb; ++b) {
                ConvTL& T = h->upT[b];
                if (T.Cin % 4 != 0 || T.Cout % 4 != 0 || T.Cin < 8 || T.Cout < 8) continue;
                // thin transposed convs stay on the FFMA kernels (default net at batch 128: 70 vs 89 us forward, 118 vs 164 us dgrad)
                static const bool convt_all = [] { const char* e = getenv("S2S_TC3_CONVT"); return e && e[0] == '2'; }();
                if (!convt_all && (int64_t)T.k * T.k * T.Cin * T.Cout < (np == 1 ? 25000 : 75000)) continue;
                const Tc3Plan pf = tc3_plan_convt_fwd(tc3_plan_for(T.h, T.w, NB, T.Cin, T.Cout, np));
                if (pf.ok) {
                    T.pf = pf; T.t3f = true; T.wqf_off = (int64_t)wq_floats; wq_floats += pf.wq_floats;
                    t3prep.push_back(Tc3WPrep{T.w_off, T.wqf_off, T.Cin, T.Cout, T.k, pf.NT, pf.nchunks_n, pf.CK, pf.kchunks, 2, np});
                    h->t3prep_maxcount = std::max(h->t3prep_maxcount, 4 * pf.nchunks_n * pf.kchunks * 9 * pf.CK * pf.NT);
                }
                const Tc3Plan pd = tc3_plan_convt_dgrad(tc3_plan_for(T.h, T.w, NB, T.Cout, T.Cin, np));
                if (pd.ok) {
                    T.pd = pd; T.t3d = true; T.wqd_off = (int64_t)wq_floats; wq_floats += pd.wq_floats;
                    t3prep.push_back(Tc3WPrep{T.w_off, T.wqd_off, T.Cout, T.Cin, T.k, pd.NT, pd.nchunks_n, pd.CK, pd.kchunks, 3, np});
                    h->t3prep_maxcount = std::max(h->t3prep_maxcount, pd.nchunks_n * pd.kchunks * 9 * pd.CK * pd.NT);
                }
            }
        }
    }
    // forward-direction blocks first: the forward pass waits only for those (run_wprep launches the two groups separately)
    std::stable_partition(t3prep.begin(), t3prep.end(), [](const Tc3WPrep& e) { return e.flip == 0 || e.flip == 2; });
    h->n_t3prep = (int)t3prep.size();
    h->n_t3prep_fwd = (int)std::count_if(t3prep.begin(), t3prep.end(), [](const Tc3WPrep& e) { return e.flip == 0 || e.flip == 2; });

    std::vector<BnFoldEntry> fold;
    auto add_fold = [&](const BnL& B) { if (B.on) fold.push_back(BnFoldEntry{B.g_off, B.be_off, B.mm_off, B.mv_off, B.ch_off, B.C}); };
    for (int b = 0; b < nb; ++b) add_fold(h->dbn[b]);
    add_fold(h->bbn);
    for (int b = nb - 1; b >= 0; --b) add_fold(h->ubn[b]);
    h->nfold = (int)fold.size();

    // ---- carve the device pool
    Bump bp;
    const size_t F = sizeof(float);
    const size_t P = h->n_params ? h->n_params : 4;
    const size_t o_params = bp.take(P * F), o_grads = bp.take(P * F), o_m = bp.take(P * F), o_v = bp.take(P * F);
    const size_t o_state = bp.take((h->n_state ? h->n_state : 4) * F);
    const size_t o_hyper = bp.take(sizeof(AdamHyper)), o_gscale = bp.take(F), o_stats = bp.take(2 * F), o_sacc = bp.take(3 * sizeof(double));
    const size_t HW = (size_t)cfg->H * cfg->W;
    const size_t o_x = bp.take(NB * HW * cfg->Cin * F), o_y = bp.take(NB * HW * h->NC * F), o_probs = bp.take(NB * HW * h->NC * F);
    size_t o_a1[MAXB], o_a2[MAXB], o_cat[MAXB], o_pl[MAXB], o_ua1[MAXB], o_ua2[MAXB], o_uo[MAXB];
    size_t o_dza1[MAXB], o_dza2[MAXB], o_dcat[MAXB], o_dpl[MAXB], o_dzua1[MAXB], o_dzua2[MAXB], o_duo[MAXB];
    size_t max_act = 0;
    for (int b = 0; b < nb; ++b) {
        const size_t px = (size_t)NB * levelH(h, b) * levelW(h, b), C = (size_t)levelC(h, b);
        max_act = std::max(max_act, px * 2 * C);
        o_a1[b] = bp.take(px * C * F); o_a2[b] = bp.take(px * C * F); o_cat[b] = bp.take(px * 2 * C * F);
        o_pl[b] = bp.take(px / 4 * C * F);
        o_ua1[b] = bp.take(px * C * F); o_ua2[b] = bp.take(px * C * F); o_uo[b] = bp.take(px * C * F);
        o_dza1[b] = bp.take(px * C * F); o_dza2[b] = bp.take(px * C * F); o_dcat[b] = bp.take(px * 2 * C * F);
        o_dpl[b] = bp.take(px / 4 * C * F);
        o_dzua1[b] = bp.take(px * C * F); o_dzua2[b] = bp.take(px * C * F); o_duo[b] = bp.take(px * C * F);
    }
    const size_t pxb = (size_t)NB * levelH(h, nb) * levelW(h, nb), Cb = (size_t)levelC(h, nb);
    const size_t o_ab1 = bp.take(pxb * Cb * F), o_ab2 = bp.take(pxb * Cb * F), o_cb = bp.take(pxb * Cb * F);
    const size_t o_dzab1 = bp.take(pxb * Cb * F), o_dzab2 = bp.take(pxb * Cb * F), o_dcb = bp.take(pxb * Cb * F);
    const size_t nch = h->n_bnch ? h->n_bnch : 4;
    const size_t o_sc = bp.take(nch * F), o_sh = bp.take(nch * F), o_mu = bp.take(nch * F), o_rs = bp.take(nch * F);
    const size_t o_wt = bp.take(P * F), o_wprep = bp.take(std::max<size_t>(wprep.size(), 1) * sizeof(WPrepEntry));
    const size_t maxC = (size_t)levelC(h, nb);
    const size_t o_ones = bp.take(maxC * F), o_zeros = bp.take(maxC * F);
    const size_t o_statp = bp.take(std::max<size_t>(stat_floats, 4) * F);
    const size_t o_headp = bp.take((size_t)head_part_floats(h->C0, h->NC, (int64_t)NB * HW) * F);
    const size_t o_gpart = bp.take(std::max<size_t>(gpart_floats, 4) * F);
    const size_t o_cam = bp.take(std::max(max_act, pxb * Cb) * F);
    const size_t o_xb = bp.take(std::max<size_t>(xb_elems, 8) * 2), o_wb = bp.take(std::max<size_t>(wb_elems, 8) * 2);
    const size_t o_wq = bp.take(std::max<size_t>(wq_floats, 4) * F), o_t3prep = bp.take(std::max<size_t>(t3prep.size(), 1) * sizeof(Tc3WPrep));
    const size_t o_cnt = bp.take((size_t)n_counters * sizeof(unsigned int));
    const size_t o_blocks = bp.take(blocks.size() * sizeof(GradBlock));
    const size_t o_gctas = bp.take(gctas.size() * sizeof(GradCta));
    const size_t o_fold = bp.take(std::max<size_t>(fold.size(), 1) * sizeof(BnFoldEntry));
    h->pool_bytes = bp.off;
    const size_t used_bytes = bp.off;
    h->device = current_device();
    cudaError_t e = pool_acquire(used_bytes, &h->pool, &h->pool_bytes);
    if (e != cudaSuccess) {
        const size_t want = h->pool_bytes;
        delete h;
        return fail(e == cudaErrorMemoryAllocation ? S2S_ERR_NOMEM : S2S_ERR_CUDA, "cudaMalloc(%zu bytes): %s", want, cudaGetErrorString(e));
    }
    e = cudaMemset(h->pool, 0, used_bytes);
    if (e != cudaSuccess) { cudaFree(h->pool); delete h; return fail(S2S_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e)); }
    auto FP = [&](size_t o) { return reinterpret_cast<float*>(h->pool + o); };
    h->params = FP(o_params); h->grads = FP(o_grads); h->m = FP(o_m); h->v = FP(o_v); h->state = FP(o_state);
    h->hyper = reinterpret_cast<AdamHyper*>(h->pool + o_hyper); h->gscale = FP(o_gscale); h->stats = FP(o_stats);
    h->stats_acc = reinterpret_cast<double*>(h->pool + o_sacc);
    h->x_in = FP(o_x); h->y_in = FP(o_y); h->probs = FP(o_probs);
    for (int b = 0; b < nb; ++b) {
        h->a1[b] = FP(o_a1[b]); h->a2[b] = FP(o_a2[b]); h->cat[b] = FP(o_cat[b]); h->pl[b] = FP(o_pl[b]);
        h->ua1[b] = FP(o_ua1[b]); h->ua2[b] = FP(o_ua2[b]); h->uo[b] = FP(o_uo[b]);
        h->dz_a1[b] = FP(o_dza1[b]); h->dz_a2[b] = FP(o_dza2[b]); h->dcat[b] = FP(o_dcat[b]); h->dpl[b] = FP(o_dpl[b]);
        h->dz_ua1[b] = FP(o_dzua1[b]); h->dz_ua2[b] = FP(o_dzua2[b]); h->duo[b] = FP(o_duo[b]);
    }
    h->ab1 = FP(o_ab1); h->ab2 = FP(o_ab2); h->cb = FP(o_cb);
    h->dz_ab1 = FP(o_dzab1); h->dz_ab2 = FP(o_dzab2); h->dcb = FP(o_dcb);
    h->bn_scale = FP(o_sc); h->bn_shift = FP(o_sh); h->bn_mean = FP(o_mu); h->bn_rstd = FP(o_rs);
    h->ones = FP(o_ones); h->zeros = FP(o_zeros);
    h->wt = FP(o_wt); h->wprep_tab = h->pool + o_wprep; h->n_wprep = (int)wprep.size();
    h->stat_part = FP(o_statp); h->head_part = FP(o_headp); h->gpart = FP(o_gpart);
    h->cam_grad = FP(o_cam);
    h->xb = reinterpret_cast<__nv_bfloat16*>(h->pool + o_xb);
    h->wb = reinterpret_cast<__nv_bfloat16*>(h->pool + o_wb);
    for (ConvL* L : tc_layers) {
        const int mrc = tcconv_make_maps(h->xb, h->wb + L->wb_off, NB, L->H, L->W, L->Cin, L->Cout, &L->map_a, &L->map_b);
        if (mrc != 0) { cudaFree(h->pool); delete h; return mrc; }
    }
    h->wq = FP(o_wq); h->t3prep_tab = h->pool + o_t3prep;
    h->counters = reinterpret_cast<unsigned int*>(h->pool + o_cnt);
    h->blocks_dev = reinterpret_cast<GradBlock*>(h->pool + o_blocks);
    h->ctas_dev = reinterpret_cast<GradCta*>(h->pool + o_gctas);
    h->fold_dev = reinterpret_cast<BnFoldEntry*>(h->pool + o_fold);

    // ---- constant tables / Keras default initial state (gamma=1, moving_var=1)
    int rc = 0;
    auto up = [&](void* d, const void* s, size_t bytes) { if (bytes && cudaMemcpy(d, s, bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = 1; };
    up(h->blocks_dev, blocks.data(), blocks.size() * sizeof(GradBlock));
    up(h->ctas_dev, gctas.data(), gctas.size() * sizeof(GradCta));
    up(h->fold_dev, fold.data(), fold.size() * sizeof(BnFoldEntry));
    up(h->wprep_tab, wprep.data(), wprep.size() * sizeof(WPrepEntry));
    up(h->t3prep_tab, t3prep.data(), t3prep.size() * sizeof(Tc3WPrep));
    std::vector<float> onesv(maxC, 1.f);
    up(h->ones, onesv.data(), maxC * F);
    std::vector<float> onesn(nch, 1.f);
    up(h->bn_scale, onesn.data(), nch * F);
    up(h->bn_rstd, onesn.data(), nch * F);
    for (const BnFoldEntry& fe : fold) {
        up(h->params + fe.gamma_off, onesv.data(), (size_t)fe.C * F);
        up(h->state + fe.mv_off, onesv.data(), (size_t)fe.C * F);
    }
    h->hyper_host = make_hyper(1e-3, 0.9, 0.999, 1e-7, 0);
    up(h->hyper, &h->hyper_host, sizeof(AdamHyper));
    h->use_side = getenv("S2S_NO_SIDE") == nullptr;
    auto env_on = [](const char* name) { const char* e = getenv(name); return e && e[0] && e[0] != '0'; };
    h->early_loads = !env_on("S2S_NO_EARLY_LOADS");
    h->bn_fold = !env_on("S2S_NO_BN_FOLD");                                            // read per handle (tests compare both paths)
    h->bn_fold_pool = h->bn_fold && !env_on("S2S_NO_BN_FOLD_POOL");
    if (const char* e = getenv("S2S_BN_FOLD_MAX")) h->bn_fold_max = atoll(e);
    for (int k = 0; k < s2s_unet::NSIDE; ++k) if (cudaStreamCreateWithFlags(&h->side[k], cudaStreamNonBlocking) != cudaSuccess) rc = 1;
    for (int k = 0; k < s2s_unet::NEV; ++k) if (cudaEventCreateWithFlags(&h->ev[k], cudaEventDisableTiming) != cudaSuccess) rc = 1;
    if (cudaEventCreateWithFlags(&h->ev_wprep, cudaEventDisableTiming) != cudaSuccess) rc = 1;
    if (cudaEventCreateWithFlags(&h->ev_wq, cudaEventDisableTiming) != cudaSuccess) rc = 1;
    const float one = 1.f;
    up(h->gscale, &one, F);
    if (rc) { cudaFree(h->pool); delete h; return fail(S2S_ERR_CUDA, "initial upload failed: %s", cudaGetErrorString(cudaGetLastError())); }
    *out = h;
    return 0;
}

int s2s_unet_destroy(s2s_unet* h) {
    if (!h) return 0;
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.first);
    for (int k = 0; k < s2s_unet::NSIDE; ++k) if (h->side[k]) { cudaStreamSynchronize(h->side[k]); cudaStreamDestroy(h->side[k]); }
    for (int k = 0; k < s2s_unet::NEV; ++k) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    if (h->ev_wprep) cudaEventDestroy(h->ev_wprep);
    if (h->ev_wq) cudaEventDestroy(h->ev_wq);
    // The caller must have drained the stream(s) it ran this handle on (Model.close does); the handle's own side
    // streams are drained here, so nothing can still touch the pool when the next handle re-uses it.
    pool_release(h->pool, h->pool_bytes, h->device);
    if (h->stats_global) cudaFree(h->stats_global);
    if (h->dp_stats_host) cudaFreeHost(h->dp_stats_host);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->ev_y) cudaEventDestroy(h->ev_y);
    for (int b = 0; b < 2; ++b) {
        if (h->stage_x[b]) cudaFree(h->stage_x[b]);
        if (h->stage_y[b]) cudaFree(h->stage_y[b]);
        if (h->ev_staged[b]) cudaEventDestroy(h->ev_staged[b]);
        if (h->ev_consumed[b]) cudaEventDestroy(h->ev_consumed[b]);
    }
    if (h->stream_stats_pinned) cudaFreeHost(h->stream_stats_pinned);
    for (int k = 0; k < s2s_unet::NPIN; ++k) {
        if (h->pin_x[k]) cudaFreeHost(h->pin_x[k]);
        if (h->pin_y[k]) cudaFreeHost(h->pin_y[k]);
        if (h->ev_pin[k]) cudaEventDestroy(h->ev_pin[k]);
    }
    delete h;
    return 0;
}

int s2s_unet_param_layout(const s2s_unet* h, s2s_tensor_desc* descs, int* n) {
    S2S_REQUIRE(h && n, "null argument");
    if (descs) {
        S2S_REQUIRE(*n >= (int)h->descs.size(), "descs array too small (%d < %zu)", *n, h->descs.size());
        memcpy(descs, h->descs.data(), h->descs.size() * sizeof(s2s_tensor_desc));
    }
    *n = (int)h->descs.size();
    return 0;
}
int s2s_unet_params(s2s_unet* h, float** p, size_t* n) { S2S_REQUIRE(h, "null"); if (p) *p = h->params; if (n) *n = h->n_params; return 0; }
int s2s_unet_state(s2s_unet* h, float** p, size_t* n) { S2S_REQUIRE(h, "null"); if (p) *p = h->state; if (n) *n = h->n_state; return 0; }
int s2s_unet_grad_arena(s2s_unet* h, float** p, size_t* n) { S2S_REQUIRE(h, "null"); if (p) *p = h->grads; if (n) *n = h->n_params; return 0; }
int s2s_unet_opt_state(s2s_unet* h, float** m, float** v, int64_t** step) {
    S2S_REQUIRE(h, "null");
    if (m) *m = h->m;
    if (v) *v = h->v;
    if (step) *step = reinterpret_cast<int64_t*>(&h->hyper->step);
    return 0;
}
int s2s_unet_io_buffers(s2s_unet* h, float** x, float** y) { S2S_REQUIRE(h, "null"); if (x) *x = h->x_in; if (y) *y = h->y_in; return 0; }
int s2s_unet_stats_buffers(s2s_unet* h, float** stats, double** stats_acc) {
    S2S_REQUIRE(h, "null");
    if (stats) *stats = h->stats;
    if (stats_acc) *stats_acc = h->stats_acc;
    return 0;
}
int s2s_unet_launch_count(const s2s_unet* h, int64_t* n) { S2S_REQUIRE(h && n, "null"); *n = h->launches; return 0; }
int s2s_unet_set_graphs(s2s_unet* h, int enable) { S2S_REQUIRE(h, "null"); h->use_graphs = enable != 0; return 0; }

int s2s_unet_activation(s2s_unet* h, const char* name, float** act, int* Hl, int* Wl, int* Cl, int* ld) {
    S2S_REQUIRE(h && name, "null argument");
    const std::string s(name);
    float* p = nullptr; int hh = 0, ww = 0, C = 0, l = 0;
    if (s == "bottleneck") { p = h->ab1; hh = levelH(h, h->nb); ww = levelW(h, h->nb); C = l = levelC(h, h->nb); }
    else if (s == "conv2d") { p = h->ab2; hh = levelH(h, h->nb); ww = levelW(h, h->nb); C = l = levelC(h, h->nb); }
    else {
        for (int b = 0; b < h->nb && !p; ++b) {
            const std::string n = std::to_string(b + 1);
            const int c = levelC(h, b);
            if (s == "down_conv" + n + "_1") { p = h->a1[b]; C = l = c; }
            else if (s == "down_conv" + n + "_2") { p = h->a2[b]; C = l = c; }
            else if (s == "up_conv" + n + "_1") { p = h->cat[b] + c; C = c; l = 2 * c; }
            else if (s == "up_conv" + n + "_2") { p = h->ua1[b]; C = l = c; }
            else if (s == "up_conv" + n + "_3") { p = h->ua2[b]; C = l = c; }
            if (p) { hh = levelH(h, b); ww = levelW(h, b); }
        }
    }
    S2S_REQUIRE(p != nullptr, "unknown layer name '%s'", name);
    if (act) *act = p;
    if (Hl) *Hl = hh;
    if (Wl) *Wl = ww;
    if (Cl) *Cl = C;
    if (ld) *ld = l;
    return 0;
}

int s2s_unet_compile(s2s_unet* h, const s2s_adam_cfg* adam, int loss_kind) {
    S2S_REQUIRE(h && adam, "null argument");
    S2S_REQUIRE(loss_kind == S2S_LOSS_CCE || loss_kind == S2S_LOSS_MASKED_MSE, "bad loss kind %d", loss_kind);
    S2S_REQUIRE((loss_kind == S2S_LOSS_CCE) == (h->cfg.head == S2S_HEAD_SOFTMAX3),
                "categorical_crossentropy needs the softmax head, masked MSE the relu head");
    h->loss_kind = loss_kind;
    h->hyper_host = make_hyper(adam->lr, adam->beta1, adam->beta2, adam->eps, 0);
    S2S_CUDA(cudaMemcpy(h->hyper, &h->hyper_host, sizeof(AdamHyper), cudaMemcpyHostToDevice));
    S2S_CUDA(cudaMemset(h->m, 0, h->n_params * sizeof(float)));
    S2S_CUDA(cudaMemset(h->v, 0, h->n_params * sizeof(float)));
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy stream; the model's streams are
    // non-blocking (not ordered against it), so drain it before any upload of restored moments can be enqueued
    S2S_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    h->compiled = true;
    return 0;
}

int s2s_unet_set_lr(s2s_unet* h, double lr, void* stream) {
    S2S_REQUIRE(h, "null");
    h->hyper_host.lr = lr;
    S2S_CUDA(cudaMemcpyAsync(&h->hyper->lr, &h->hyper_host.lr, sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}

int s2s_unet_reset_epoch_stats(s2s_unet* h, void* stream) {
    S2S_REQUIRE(h, "null");
    S2S_CUDA(cudaMemsetAsync(h->stats_acc, 0, 3 * sizeof(double), (cudaStream_t)stream));
    return 0;
}

int s2s_unet_forward(s2s_unet* h, const float* x, int N, float* probs, int training, void* stream) {
    S2S_CHECK(check_N(h, N));
    cudaStream_t st = (cudaStream_t)stream;
    S2S_CHECK(stage_inputs(h, x, nullptr, N, st));
    const bool tr = training != 0;
    S2S_CHECK(run_cached(h, tr ? GK_FWD_TRAIN : GK_FWD_INFER, N, st, [&](cudaStream_t s) { return seq_forward(h, N, tr, s); }));
    h->last_forward_training = tr; h->last_N = N;
    if (probs && probs != h->probs)
        S2S_CUDA(cudaMemcpyAsync(probs, h->probs, (size_t)N * h->cfg.H * h->cfg.W * h->NC * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// MASKED_MSE: number of land points of the [H,W] mask (one-off synchronous count per mask pointer)
static int prep_mask(s2s_unet* h, const uint8_t* mask, cudaStream_t st) {
    const int hw = h->cfg.H * h->cfg.W;
    if (!mask) { h->mask_count = (float)hw; h->mask_cache = nullptr; return 0; }
    if (mask != h->mask_cache) {
        std::vector<uint8_t> hm(hw);
        S2S_CUDA(cudaStreamSynchronize(st));
        S2S_CUDA(cudaMemcpy(hm.data(), mask, hw, cudaMemcpyDeviceToHost));
        double c = 0;
        for (int i = 0; i < hw; ++i) c += hm[i] ? 1.0 : 0.0;
        h->mask_cache = mask;
        h->mask_count = c > 0 ? (float)c : 1.f;
    }
    return 0;
}

static int train_like(s2s_unet* h, const float* x, const float* y, const uint8_t* mask, int N, float gscale, bool adam,
                      float* stats_dev, cudaStream_t st) {
    S2S_CHECK(check_N(h, N));
    S2S_REQUIRE(h->compiled, "call s2s_unet_compile before training");
    S2S_CHECK(stage_inputs(h, x, y, N, st));
    S2S_CHECK(set_gscale(h, gscale, st));
    if (h->loss_kind == S2S_LOSS_MASKED_MSE) {
        S2S_CHECK(prep_mask(h, mask, st));
        h->mask_norm_cache = 1.f / ((float)N * h->mask_count);
        const int64_t before = launch_counter();
        const int rc = seq_train(h, N, adam, mask, st);   // eager: mask pointer / norm are by-value arguments
        h->launches += launch_counter() - before;
        S2S_CHECK(rc);
    } else {
        S2S_CHECK(run_cached(h, adam ? GK_TRAIN : GK_BWD, N, st, [&](cudaStream_t s) { return seq_train(h, N, adam, nullptr, s); }));
    }
    h->last_forward_training = true; h->last_N = N;
    if (stats_dev && stats_dev != h->stats) S2S_CUDA(cudaMemcpyAsync(stats_dev, h->stats, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int s2s_unet_train_step(s2s_unet* h, const float* x, const float* y, const uint8_t* mask, int N, float* stats_dev, void* stream) {
    return train_like(h, x, y, mask, N, 1.f, true, stats_dev, (cudaStream_t)stream);
}
// End-to-end step from host batches; n_global > 0 = data-parallel step on this rank's shard (communicator attached).
static int train_step_host_impl(s2s_unet* h, const float* x_host, const float* y_host, int N, int n_global, float* stats_host,
                                cudaStream_t st) {
    S2S_CHECK(check_N(h, N));
    S2S_REQUIRE(x_host && y_host, "null host batch");
    S2S_REQUIRE(h->compiled, "call s2s_unet_compile before training");
    const bool dp = n_global > 0;
    if (dp) {
        S2S_REQUIRE(h->dp, "attach a communicator (s2s_unet_attach_dp) first");
        S2S_REQUIRE(h->loss_kind == S2S_LOSS_CCE, "the data-parallel step supports the categorical cross-entropy path");
        S2S_REQUIRE(n_global >= N && n_global < 65536, "bad global batch %d (local %d)", n_global, N);
    }
    const size_t xb = (size_t)N * h->cfg.H * h->cfg.W * h->cfg.Cin * sizeof(float);
    const size_t yb = (size_t)N * h->cfg.H * h->cfg.W * h->NC * sizeof(float);
    const bool split = h->loss_kind == S2S_LOSS_CCE && h->use_graphs && st != nullptr && !prof().on;
    const float gs = dp ? (float)N / (float)n_global : 1.f;
    if (!split) {
        S2S_CUDA(cudaMemcpyAsync(h->x_in, x_host, xb, cudaMemcpyHostToDevice, st));
        S2S_CUDA(cudaMemcpyAsync(h->y_in, y_host, yb, cudaMemcpyHostToDevice, st));
        if (dp) S2S_CHECK(s2s_unet_dp_train_step(h, h->x_in, h->y_in, N, n_global, nullptr, st));
        else S2S_CHECK(train_like(h, h->x_in, h->y_in, nullptr, N, 1.f, true, nullptr, st));
    } else {
        // the targets are first read by the head kernel: their copy runs on a second stream behind the forward graph.
        // (The previous call ended with a synchronisation, so y_in is free.)
        if (!h->copy_stream) {
            S2S_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            S2S_CUDA(cudaEventCreateWithFlags(&h->ev_y, cudaEventDisableTiming));
        }
        S2S_CUDA(cudaMemcpyAsync(h->x_in, x_host, xb, cudaMemcpyHostToDevice, st));
        S2S_CUDA(cudaMemcpyAsync(h->y_in, y_host, yb, cudaMemcpyHostToDevice, h->copy_stream));
        S2S_CUDA(cudaEventRecord(h->ev_y, h->copy_stream));
        S2S_CHECK(set_gscale(h, gs, st));
        const int key = dp ? (N << 16) | n_global : N;
        h->dp_n_global = n_global;
        h->dp_in_step = dp;
        int rc = run_cached(h, dp ? GK_HOST_DP_FWD : GK_HOST_FWD, key, st, [&](cudaStream_t s) { return seq_train_fwd(h, N, s); });
        if (rc == 0) {
            cudaStreamWaitEvent(st, h->ev_y, 0);
            rc = run_cached(h, dp ? GK_HOST_DP_BWD : GK_HOST_BWD, key, st, [&](cudaStream_t s) { return seq_train_bwd(h, N, true, nullptr, s, dp); });
        }
        h->dp_in_step = false;
        S2S_CHECK(rc);
        h->last_forward_training = true; h->last_N = N;
    }
    if (dp) {
        // the global statistics travel with the exchange's error code: a peer that timed out must fail the step
        S2S_CUDA(cudaMemcpyAsync(h->dp_stats_host, h->stats_global, 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
        S2S_CUDA(cudaStreamSynchronize(st));
        if (stats_host) { stats_host[0] = h->dp_stats_host[0]; stats_host[1] = h->dp_stats_host[1]; }
        if (h->dp_stats_host[2] != 0.f)
            return fail(S2S_ERR_STATE, "data-parallel exchange timed out at sync group %d: a peer is slow or dead; this step was not applied",
                        (int)h->dp_stats_host[2] - 1);
        return 0;
    }
    if (stats_host) S2S_CUDA(cudaMemcpyAsync(stats_host, h->stats, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    S2S_CUDA(cudaStreamSynchronize(st));
    return 0;
}
int s2s_unet_train_step_host(s2s_unet* h, const float* x_host, const float* y_host, int N, float* stats_host, void* stream) {
    return train_step_host_impl(h, x_host, y_host, N, 0, stats_host, (cudaStream_t)stream);
}
// A stream of end-to-end steps (what model.fit does with host arrays, training.py:102-103): every step copies ITS batch host ->
// device and returns ITS {loss, accuracy} device -> host, but the copy stream stages batch i + 1 into a second device slot while
// step i computes (x is read by the first conv AND by its weight gradient at the very end of the step, so the step's own input
// buffer is busy for the whole step: double buffering + a 1 us device-to-device copy instead of an exposed 30 us H2D).
constexpr int STREAM_CHUNK = 256;
static int train_steps_host_impl(s2s_unet* h, const float* const* x_hosts, const float* const* y_hosts, int nsteps, int N, int n_global,
                                 float* stats_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    S2S_CHECK(check_N(h, N));
    const bool dp = n_global > 0;
    if (dp) {
        S2S_REQUIRE(h->dp, "attach a communicator (s2s_unet_attach_dp) first");
        S2S_REQUIRE(h->loss_kind == S2S_LOSS_CCE, "the data-parallel step supports the categorical cross-entropy path");
        S2S_REQUIRE(n_global >= N && n_global < 65536, "bad global batch %d (local %d)", n_global, N);
    }
    S2S_REQUIRE(x_hosts && y_hosts && nsteps >= 0, "null host batch list");
    S2S_REQUIRE(h->compiled, "call s2s_unet_compile before training");
    S2S_REQUIRE(st != nullptr, "the streamed steps need an explicit (non-default) stream");
    if (nsteps == 0) return 0;
    const size_t xcap = (size_t)h->cfg.max_batch * h->cfg.H * h->cfg.W * h->cfg.Cin * sizeof(float);
    const size_t ycap = (size_t)h->cfg.max_batch * h->cfg.H * h->cfg.W * h->NC * sizeof(float);
    const size_t xb = (size_t)N * h->cfg.H * h->cfg.W * h->cfg.Cin * sizeof(float);
    const size_t yb = (size_t)N * h->cfg.H * h->cfg.W * h->NC * sizeof(float);
    if (!h->copy_stream) {
        S2S_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        S2S_CUDA(cudaEventCreateWithFlags(&h->ev_y, cudaEventDisableTiming));
    }
    if (!h->stage_x[0]) {
        for (int b = 0; b < 2; ++b) {
            S2S_CUDA(cudaMalloc((void**)&h->stage_x[b], xcap));
            S2S_CUDA(cudaMalloc((void**)&h->stage_y[b], ycap));
            S2S_CUDA(cudaEventCreateWithFlags(&h->ev_staged[b], cudaEventDisableTiming));
            S2S_CUDA(cudaEventCreateWithFlags(&h->ev_consumed[b], cudaEventDisableTiming));
        }
        S2S_CUDA(cudaMallocHost((void**)&h->stream_stats_pinned, (size_t)STREAM_CHUNK * 3 * sizeof(float)));
    }
    cudaStream_t cs = h->copy_stream;
    // A pageable source would make cudaMemcpyAsync synchronous (the driver stages it itself and blocks the enqueueing thread):
    // such batches are copied into a pinned ring slot by THIS thread — while the GPU is busy with the steps already enqueued —
    // and travel from there.  Pinned / registered sources are copied directly.
    auto is_pageable = [](const void* p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
        return at.type == cudaMemoryTypeUnregistered;
    };
    auto stage = [&](int i) -> int {       // copy stream: batch i -> slot i & 1 (once the step that last used the slot has read it)
        const int b = i & 1;
        const float* sx = x_hosts[i];
        const float* sy = y_hosts[i];
        int k = -1;
        if (is_pageable(sx) || is_pageable(sy)) {
            k = h->pin_next;
            h->pin_next = (k + 1) % s2s_unet::NPIN;
            if (!h->pin_x[k]) {
                S2S_CUDA(cudaMallocHost((void**)&h->pin_x[k], xcap));
                S2S_CUDA(cudaMallocHost((void**)&h->pin_y[k], ycap));
                S2S_CUDA(cudaEventCreateWithFlags(&h->ev_pin[k], cudaEventDisableTiming));
            }
            if (h->pin_busy[k]) S2S_CUDA(cudaEventSynchronize(h->ev_pin[k]));      // the H2D copy that last read this slot is done
            memcpy(h->pin_x[k], sx, xb);
            memcpy(h->pin_y[k], sy, yb);
            sx = h->pin_x[k]; sy = h->pin_y[k];
        }
        if (i >= 2) S2S_CUDA(cudaStreamWaitEvent(cs, h->ev_consumed[b], 0));
        S2S_CUDA(cudaMemcpyAsync(h->stage_x[b], sx, xb, cudaMemcpyHostToDevice, cs));
        S2S_CUDA(cudaMemcpyAsync(h->stage_y[b], sy, yb, cudaMemcpyHostToDevice, cs));
        if (k >= 0) { S2S_CUDA(cudaEventRecord(h->ev_pin[k], cs)); h->pin_busy[k] = true; }
        S2S_CUDA(cudaEventRecord(h->ev_staged[b], cs));
        return 0;
    };
    S2S_CHECK(stage(0));
    for (int c0 = 0; c0 < nsteps; c0 += STREAM_CHUNK) {
        const int c1 = std::min(nsteps, c0 + STREAM_CHUNK);
        for (int i = c0; i < c1; ++i) {
            const int b = i & 1;
            if (i + 1 < nsteps) S2S_CHECK(stage(i + 1));
            S2S_CUDA(cudaStreamWaitEvent(st, h->ev_staged[b], 0));
            S2S_CUDA(cudaMemcpyAsync(h->x_in, h->stage_x[b], xb, cudaMemcpyDeviceToDevice, st));
            S2S_CUDA(cudaMemcpyAsync(h->y_in, h->stage_y[b], yb, cudaMemcpyDeviceToDevice, st));
            S2S_CUDA(cudaEventRecord(h->ev_consumed[b], st));
            if (dp) {
                // global {loss, accuracy} and the exchange's error code (a peer that timed out leaves the step unapplied)
                S2S_CHECK(s2s_unet_dp_train_step(h, h->x_in, h->y_in, N, n_global, nullptr, st));
                S2S_CUDA(cudaMemcpyAsync(h->stream_stats_pinned + 3 * (i - c0), h->stats_global, 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
            } else {
                S2S_CHECK(train_like(h, h->x_in, h->y_in, nullptr, N, 1.f, true, nullptr, st));
                S2S_CUDA(cudaMemcpyAsync(h->stream_stats_pinned + 3 * (i - c0), h->stats, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
            }
        }
        S2S_CUDA(cudaStreamSynchronize(st));
        for (int i = c0; i < c1; ++i) {
            const float* r = h->stream_stats_pinned + 3 * (i - c0);
            if (stats_host) { stats_host[2 * (size_t)i] = r[0]; stats_host[2 * (size_t)i + 1] = r[1]; }
            if (dp && r[2] != 0.f)
                return fail(S2S_ERR_STATE, "data-parallel exchange timed out at sync group %d in streamed step %d: a peer is slow or dead; "
                                           "that step and the ones after it were not applied", (int)r[2] - 1, i);
        }
    }
    return 0;
}
int s2s_unet_train_steps_host(s2s_unet* h, const float* const* x_hosts, const float* const* y_hosts, int nsteps, int N, float* stats_host,
                              void* stream) {
    return train_steps_host_impl(h, x_hosts, y_hosts, nsteps, N, 0, stats_host, stream);
}
int s2s_unet_dp_train_steps_host(s2s_unet* h, const float* const* x_hosts, const float* const* y_hosts, int nsteps, int n_local, int n_global,
                                 float* stats_host, void* stream) {
    S2S_REQUIRE(n_global > 0, "n_global must be positive");
    return train_steps_host_impl(h, x_hosts, y_hosts, nsteps, n_local, n_global, stats_host, stream);
}
int s2s_unet_dp_train_step_host(s2s_unet* h, const float* x_host, const float* y_host, int n_local, int n_global, float* stats_host,
                                void* stream) {
    S2S_REQUIRE(n_global > 0, "n_global must be positive");
    return train_step_host_impl(h, x_host, y_host, n_local, n_global, stats_host, (cudaStream_t)stream);
}
int s2s_unet_backward_only(s2s_unet* h, const float* x, const float* y, const uint8_t* mask, int N, float grad_scale,
                           float* stats_dev, void* stream) {
    return train_like(h, x, y, mask, N, grad_scale, false, stats_dev, (cudaStream_t)stream);
}
int s2s_unet_apply_adam(s2s_unet* h, void* stream) {
    S2S_REQUIRE(h && h->compiled, "call s2s_unet_compile first");
    const int64_t before = launch_counter();
    S2S_CHECK(adam_launch(h->params, h->grads, h->m, h->v, h->n_params, h->hyper, nullptr, (cudaStream_t)stream));
    h->launches += launch_counter() - before;
    return 0;
}

// ---- data parallelism over peer memory (dp.cuh) ---------------------------------------------
int s2s_dp_create(int rank, int world, size_t n_floats, s2s_dp** out) {
    S2S_REQUIRE(out, "null");
    S2S_REQUIRE(world >= 1 && world <= DP_MAXW && rank >= 0 && rank < world, "bad rank %d / world %d (max %d)", rank, world, DP_MAXW);
    s2s_dp* d = new s2s_dp();
    d->rank = rank; d->world = world;
    d->n_pad = (n_floats + 3) / 4 * 4;
    d->push = dp_use_push(d->n_pad, world);       // the same decision on every rank (same n_floats, world, environment)
    d->lay = dp_layout(d->n_pad, world, d->push);
    cudaError_t e = cudaMalloc((void**)&d->local, d->lay.total);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->state, 256);
    if (e == cudaSuccess) e = cudaMemset(d->local, 0, d->lay.total);
    if (e == cudaSuccess) e = cudaMemset(d->state, 0, 256);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(d->local); cudaFree(d->state);
        delete d;
        return fail(S2S_ERR_CUDA, "s2s_dp_create: %s", cudaGetErrorString(e));
    }
    d->mapped[rank] = d->local;
    if (world == 1) {   // nothing to connect
        DpDev& v = d->dev;
        v.rank = 0; v.world = 1; v.n_pad = d->n_pad; v.push = d->push ? 1 : 0;
        v.epoch = (unsigned long long*)d->state; v.error = (int*)(d->state + 64); v.counter = (unsigned int*)(d->state + 128);
        v.flags[0] = (unsigned long long*)(d->local + d->lay.flags_off);
        v.bn[0] = (double*)(d->local + d->lay.bn_off);
        v.grads[0] = (float*)(d->local + d->lay.grads_off);
        v.stats[0] = (float*)(d->local + d->lay.stats_off);
        d->connected = true;
    }
    *out = d;
    return 0;
}
int s2s_dp_ipc_handle(s2s_dp* d, void* handle64) {
    S2S_REQUIRE(d && handle64, "null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hnd;
    S2S_CUDA(cudaIpcGetMemHandle(&hnd, d->local));
    memcpy(handle64, &hnd, 64);
    return 0;
}
int s2s_dp_connect(s2s_dp* d, const void* handles) {
    S2S_REQUIRE(d && handles, "null");
    for (int p = 0; p < d->world; ++p) {
        if (p == d->rank || d->mapped[p]) continue;
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, (const char*)handles + 64 * p, 64);
        void* ptr = nullptr;
        S2S_CUDA(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
        d->mapped[p] = (char*)ptr; d->opened[p] = true;
    }
    DpDev& v = d->dev;
    v.rank = d->rank; v.world = d->world; v.n_pad = d->n_pad; v.push = d->push ? 1 : 0;
    v.epoch = (unsigned long long*)d->state; v.error = (int*)(d->state + 64); v.counter = (unsigned int*)(d->state + 128);
    for (int p = 0; p < d->world; ++p) {
        v.flags[p] = (unsigned long long*)(d->mapped[p] + d->lay.flags_off);
        v.bn[p] = (double*)(d->mapped[p] + d->lay.bn_off);
        v.grads[p] = (float*)(d->mapped[p] + d->lay.grads_off);
        v.stats[p] = (float*)(d->mapped[p] + d->lay.stats_off);
    }
    d->connected = true;
    return 0;
}
int s2s_dp_error(s2s_dp* d, int* err) {      // synchronous: != 0 after a peer timed out (1 + sync group)
    S2S_REQUIRE(d && err, "null");
    S2S_CUDA(cudaMemcpy(err, d->state + 64, sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}
int s2s_dp_destroy(s2s_dp* d) {
    if (!d) return 0;
    cudaDeviceSynchronize();
    for (int p = 0; p < d->world; ++p)
        if (d->opened[p]) cudaIpcCloseMemHandle(d->mapped[p]);
    cudaFree(d->local);
    cudaFree(d->state);
    delete d;
    return 0;
}
int s2s_unet_attach_dp(s2s_unet* h, s2s_dp* d, int sync_bn) {
    S2S_REQUIRE(h, "null handle");
    for (auto& kv : h->graphs) if ((kv.first >> 40) >= GK_DP) { cudaGraphExecDestroy(kv.second.first); kv.second.first = nullptr; }
    for (auto it = h->graphs.begin(); it != h->graphs.end();) it = it->second.first ? std::next(it) : h->graphs.erase(it);
    if (!d) { h->dp = nullptr; h->dp_sync_bn = false; return 0; }
    S2S_REQUIRE(d->connected, "s2s_dp_connect must succeed on every rank before attaching");
    S2S_REQUIRE(d->n_pad >= h->n_params, "communicator sized for %zu floats, model has %zu", d->n_pad, h->n_params);
    if (!h->stats_global) {
        S2S_CUDA(cudaMalloc((void**)&h->stats_global, 16));
        S2S_CUDA(cudaMemset(h->stats_global, 0, 16));
        S2S_CUDA(cudaHostAlloc((void**)&h->dp_stats_host, 16, cudaHostAllocDefault));
        memset(h->dp_stats_host, 0, 16);
    }
    h->dp = d; h->dp_sync_bn = sync_bn != 0;
    return 0;
}
int s2s_unet_dp_train_step(s2s_unet* h, const float* x, const float* y, int n_local, int n_global, float* stats_dev, void* stream) {
    S2S_CHECK(check_N(h, n_local));
    S2S_REQUIRE(h->compiled && h->dp, "compile the model and attach a communicator (s2s_unet_attach_dp) first");
    S2S_REQUIRE(h->loss_kind == S2S_LOSS_CCE, "the data-parallel step supports the categorical cross-entropy path");
    S2S_REQUIRE(n_global >= n_local && n_global < 65536, "bad global batch %d (local %d)", n_global, n_local);
    cudaStream_t st = (cudaStream_t)stream;
    S2S_CHECK(stage_inputs(h, x, y, n_local, st));
    S2S_CHECK(set_gscale(h, (float)n_local / (float)n_global, st));
    h->dp_n_global = n_global;
    h->dp_in_step = true;
    const int rc = run_cached(h, GK_DP, (n_local << 16) | n_global, st, [&](cudaStream_t s) { return seq_train(h, n_local, true, nullptr, s, true); });
    h->dp_in_step = false;
    S2S_CHECK(rc);
    h->last_forward_training = true; h->last_N = n_local;
    if (stats_dev) S2S_CUDA(cudaMemcpyAsync(stats_dev, h->stats_global, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}
int s2s_unet_eval_batch(s2s_unet* h, const float* x, const float* y, const uint8_t* mask, int N, float* stats_dev, void* stream) {
    S2S_CHECK(check_N(h, N));
    cudaStream_t st = (cudaStream_t)stream;
    S2S_CHECK(stage_inputs(h, x, y, N, st));
    if (h->loss_kind == S2S_LOSS_MASKED_MSE) {
        S2S_CHECK(prep_mask(h, mask, st));
        h->mask_norm_cache = 1.f / ((float)N * h->mask_count);
        const int64_t before = launch_counter();
        const int rc = seq_eval(h, N, mask, st);
        h->launches += launch_counter() - before;
        S2S_CHECK(rc);
    } else {
        S2S_CHECK(run_cached(h, GK_EVAL, N, st, [&](cudaStream_t s) { return seq_eval(h, N, nullptr, s); }));
    }
    h->last_forward_training = false; h->last_N = N;
    if (stats_dev && stats_dev != h->stats) S2S_CUDA(cudaMemcpyAsync(stats_dev, h->stats, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int s2s_unet_gradcam(s2s_unet* h, const float* x, int N, const char* layer_name, int cls, float* cam, void* stream) {
    S2S_CHECK(check_N(h, N));
    S2S_REQUIRE(layer_name && cam, "null argument");
    S2S_REQUIRE(cls >= 0 && cls < h->NC, "class %d outside [0,%d)", cls, h->NC);
    cudaStream_t st = (cudaStream_t)stream;
    float* probe = nullptr;
    S2S_CHECK(s2s_unet_activation(h, layer_name, &probe, nullptr, nullptr, nullptr, nullptr));   // validates the name
    S2S_CHECK(stage_inputs(h, x, nullptr, N, st));
    const int64_t before = launch_counter();
    CamTarget tgt;
    tgt.layer = layer_name;
    h->ev_next = 0;
    S2S_CHECK(run_wprep(h, st));
    S2S_CHECK(run_forward_body(h, N, false, st));
    {
        HeadArgs a;
        memset(&a, 0, sizeof a);
        a.u = h->ua2[0]; a.ldu = h->C0;
        a.wh = h->params + h->head_w; a.bh = h->params + h->head_b;
        a.dz_out = h->dz_ua2[0];
        a.apply_elugrad = (tgt.layer == "up_conv1_3") ? 0 : 1;
        a.part = h->head_part; a.counter = h->counters + 0;
        a.grad_scale = 1.f;
        a.npix = (int64_t)N * h->cfg.H * h->cfg.W;
        a.loss_kind = h->loss_kind; a.train = 0;
        a.cam_cls = cls; a.cam_norm = 1.f / (float)(h->cfg.H * h->cfg.W);
        S2S_CHECK(head_launch(a, h->C0, h->NC, st));
    }
    S2S_CHECK(run_backward(h, N, &tgt, st));
    S2S_REQUIRE(tgt.found, "layer '%s' not reached by the backward walk", layer_name);
    prof_begin(st, "gradcam_combine", 4.0 * N * tgt.H * tgt.W * (2.0 * tgt.C + 1.0), 0.0);
    gradcam_kernel<<<N, 256, (size_t)tgt.C * sizeof(float), st>>>(tgt.grad, tgt.ld, tgt.act, tgt.lda, tgt.H * tgt.W, tgt.C, cam);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    h->launches += launch_counter() - before;
    h->last_forward_training = false; h->last_N = N;
    return 0;
}

// ---- on-device fit protocol: whole epochs without host round trips (training.py:102-103) ----
int s2s_unet_fit_epoch(s2s_unet* h, const float* x_all, const float* y_all, const int32_t* perm, int T, int batch_size,
                       const uint8_t* mask, void* stream) {
    S2S_REQUIRE(h && x_all && y_all && T > 0, "bad argument");
    S2S_CHECK(check_N(h, batch_size));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t xrow = (int64_t)h->cfg.H * h->cfg.W * h->cfg.Cin, yrow = (int64_t)h->cfg.H * h->cfg.W * h->NC;
    for (int i = 0; i < T; i += batch_size) {
        const int n = std::min(batch_size, T - i);          // the partial last batch is kept (Keras)
        const int64_t before = launch_counter();
        S2S_CHECK(gather_rows(perm ? x_all : x_all + (int64_t)i * xrow, perm ? perm + i : nullptr, h->x_in, xrow, n, st));
        S2S_CHECK(gather_rows(perm ? y_all : y_all + (int64_t)i * yrow, perm ? perm + i : nullptr, h->y_in, yrow, n, st));
        h->launches += launch_counter() - before;
        S2S_CHECK(train_like(h, h->x_in, h->y_in, mask, n, 1.f, true, nullptr, st));
    }
    return 0;
}
int s2s_unet_eval_dataset(s2s_unet* h, const float* x_all, const float* y_all, int T, int batch_size, const uint8_t* mask,
                          void* stream) {
    S2S_REQUIRE(h && x_all && y_all && T > 0, "bad argument");
    S2S_CHECK(check_N(h, batch_size));
    const int64_t xrow = (int64_t)h->cfg.H * h->cfg.W * h->cfg.Cin, yrow = (int64_t)h->cfg.H * h->cfg.W * h->NC;
    for (int i = 0; i < T; i += batch_size) {
        const int n = std::min(batch_size, T - i);
        S2S_CHECK(s2s_unet_eval_batch(h, x_all + (int64_t)i * xrow, y_all + (int64_t)i * yrow, mask, n, nullptr, stream));
    }
    return 0;
}
int s2s_unet_predict_dataset(s2s_unet* h, const float* x_all, int T, int batch_size, float* probs_all, void* stream) {
    S2S_REQUIRE(h && x_all && probs_all && T > 0, "bad argument");
    S2S_CHECK(check_N(h, batch_size));
    const int64_t xrow = (int64_t)h->cfg.H * h->cfg.W * h->cfg.Cin, yrow = (int64_t)h->cfg.H * h->cfg.W * h->NC;
    for (int i = 0; i < T; i += batch_size) {
        const int n = std::min(batch_size, T - i);
        S2S_CHECK(s2s_unet_forward(h, x_all + (int64_t)i * xrow, n, probs_all + (int64_t)i * yrow, 0, stream));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
int s2s_adam_step(float* p, const float* g, float* m, float* v, size_t n, const s2s_adam_cfg* cfg, int64_t step, void* stream) {
    S2S_REQUIRE(p && g && m && v && cfg, "null argument");
    S2S_REQUIRE(step >= 1, "step is 1-based (got %lld)", (long long)step);
    AdamHyper hv = make_hyper(cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, step);
    return adam_launch(p, g, m, v, n, nullptr, &hv, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------
int s2s_rps_map(const float* p, const float* o, int T, int Y, int X, float* out, void* stream) {
    S2S_REQUIRE(p && o && out && T > 0 && Y > 0 && X > 0, "bad argument");
    const int64_t YX = (int64_t)Y * X;
    prof_begin((cudaStream_t)stream, "rps_map", 4.0 * (6.0 * T + 1.0) * YX, 0.0);
    rps_kernel<false><<<(unsigned)cdiv64(YX, 32), dim3(32, SK_TS), 0, (cudaStream_t)stream>>>(p, nullptr, o, T, YX, out);
    prof_end((cudaStream_t)stream);
    S2S_LAUNCH_CHECK();
    return 0;
}
int s2s_rpss_map(const float* f, const float* r, const float* o, int T, int Y, int X, float* out, void* stream) {
    S2S_REQUIRE(f && r && o && out && T > 0 && Y > 0 && X > 0, "bad argument");
    const int64_t YX = (int64_t)Y * X;
    prof_begin((cudaStream_t)stream, "rpss_map", 4.0 * (9.0 * T + 1.0) * YX, 0.0);
    rps_kernel<true><<<(unsigned)cdiv64(YX, 32), dim3(32, SK_TS), 0, (cudaStream_t)stream>>>(f, r, o, T, YX, out);
    prof_end((cudaStream_t)stream);
    S2S_LAUNCH_CHECK();
    return 0;
}
int s2s_acc_map(const float* x, const float* y, const int32_t* order, const int32_t* gstart, int n_groups, int T, int Y, int X,
                float* acc, float* cc, void* stream) {
    S2S_REQUIRE(x && y && order && gstart && n_groups > 0 && T > 0 && Y > 0 && X > 0, "bad argument");
    const int64_t YX = (int64_t)Y * X;
    prof_begin((cudaStream_t)stream, "acc_map", 4.0 * (2.0 * T + 2.0) * YX, 0.0);
    acc_kernel<<<(unsigned)cdiv64(YX, 32), dim3(32, SK_TS), 0, (cudaStream_t)stream>>>(x, y, order, gstart, n_groups, YX, acc, cc);
    prof_end((cudaStream_t)stream);
    S2S_LAUNCH_CHECK();
    return 0;
}
int s2s_ensemble_mean(const float* x, int T, int M, int Y, int X, float* out, void* stream) {
    S2S_REQUIRE(x && out && T > 0 && M > 0 && Y > 0 && X > 0, "bad argument");
    const int64_t YX = (int64_t)Y * X, total = (int64_t)T * YX;
    prof_begin((cudaStream_t)stream, "ensemble_mean", 4.0 * (M + 1.0) * total, 0.0);
    ensemble_mean_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, (cudaStream_t)stream>>>(x, M, YX, total, out);
    prof_end((cudaStream_t)stream);
    S2S_LAUNCH_CHECK();
    return 0;
}
int s2s_mme_combine(const float* probs, int n_models, int64_t n_points, float* out, void* stream) {
    S2S_REQUIRE(probs && out && n_models > 0 && n_points > 0, "bad argument");
    prof_begin((cudaStream_t)stream, "mme_combine", 12.0 * (n_models + 1.0) * n_points, 0.0);
    mme_combine_kernel<<<(unsigned)cdiv64(n_points, 256), 256, 0, (cudaStream_t)stream>>>(probs, n_models, n_points, out);
    prof_end((cudaStream_t)stream);
    S2S_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------
// single-operator entry points (parity tests / profiling).  They allocate scratch with
// cudaMalloc and are NOT part of the hot path.
// ---------------------------------------------------------------------------------------
int s2s_op_conv3x3_fwd(const float* x, const float* w, const float* b, float* y, int N, int H, int W, int Cin, int Cout,
                       int apply_elu, void* stream) {
    GConvArgs a;
    memset(&a, 0, sizeof a);
    a.in = x; a.ldin = Cin; a.Hin = H; a.Win = W; a.Cb = Cin;
    a.w = w; a.bias = b;
    a.out = y; a.ldout = Cout; a.Hout = H; a.Wout = W; a.Ca = Cout;
    a.pad = 1; a.epi = apply_elu ? EPI_BIAS_ELU : EPI_BIAS; a.N = N;
    return gconv_run(3, 1, true, a, (cudaStream_t)stream);
}
// tensor-core variant (bf16 operands, fp32 accumulate in TMEM): casts x and w, builds the TMA maps, launches
int s2s_op_conv3x3_fwd_tc(const float* x, const float* w, const float* b, float* y, int N, int H, int W, int Cin, int Cout,
                          int apply_elu, void* stream) {
    S2S_REQUIRE(tcconv_eligible(Cin, Cout), "tensor-core conv needs Cin %% 64 == 0, Cout %% 16 == 0, Cout <= 256 (got %d -> %d)", Cin, Cout);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nx = (int64_t)N * H * W * Cin, nw = (int64_t)9 * Cin * Cout;
    __nv_bfloat16* tmp = nullptr;
    S2S_CUDA(cudaMalloc((void**)&tmp, (size_t)(nx + nw + 64) * sizeof(__nv_bfloat16)));
    __nv_bfloat16* xb = tmp;
    __nv_bfloat16* wb = tmp + ((nx + 63) / 64) * 64;
    cast_bf16_kernel<<<(unsigned)cdiv64(cdiv64(nx, 4), 256), 256, 0, st>>>(x, xb, nx);
    wprep_bf16_kernel<<<std::min(cdiv((int)nw, 256), 64), 256, 0, st>>>(w, wb, Cin, Cout);
    CUtensorMap ma, mb;
    int rc = tcconv_make_maps(xb, wb, N, H, W, Cin, Cout, &ma, &mb);
    if (rc == 0) {
        TcConvArgs a;
        memset(&a, 0, sizeof a);
        a.bias = b; a.out = y; a.ldout = Cout; a.out_coff = 0;
        a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.apply_elu = apply_elu;
        rc = tcconv_launch(ma, mb, a, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (rc == 0 && e != cudaSuccess) return fail(S2S_ERR_CUDA, "tcconv: %s", cudaGetErrorString(e));
    return rc;
}

// tcgen05 tf32 forward / input gradient (tc3conv.cuh): prepares the weight blocks, builds the TMA map, launches, syncs
static int op_tc3(const float* in, const float* w, const float* bias, const float* aux, float* out, int N, int H, int W,
                  int Kc, int Nc, int flip, int epi, int npass, cudaStream_t st) {
    S2S_REQUIRE(npass == 1 || npass == 3, "npass must be 1 or 3");
    const Tc3Plan p = tc3_plan_for(H, W, N, Kc, Nc, npass);
    S2S_REQUIRE(p.ok, "tf32 tensor-core conv needs channel counts %% 4 == 0 and >= 8 contracted channels (got %d -> %d)", Kc, Nc);
    char* tmp = nullptr;
    S2S_CUDA(cudaMalloc((void**)&tmp, p.wq_floats * sizeof(float) + 256));
    float* wq = reinterpret_cast<float*>(tmp + 256);
    Tc3WPrep e{0, 0, Kc, Nc, 0, p.NT, p.nchunks_n, p.CK, p.kchunks, flip, npass};
    cudaMemcpyAsync(tmp, &e, sizeof e, cudaMemcpyHostToDevice, st);
    tc3_wprep_kernel<<<dim3(32, 1), 256, 0, st>>>(reinterpret_cast<const Tc3WPrep*>(tmp), w, wq);
    CUtensorMap map;
    int rc = tc3_make_map_any(in, N, H, W, Kc, Kc, p.CK, &map);
    if (rc == 0) {
        Tc3Args a;
        memset(&a, 0, sizeof a);
        a.wq = wq; a.bias = bias; a.aux = aux; a.ldaux = Nc; a.out = out; a.ldout = Nc; a.in = in; a.ldin = Kc;
        a.N = N; a.H = H; a.W = W; a.Cin = Kc; a.Cout = Nc; a.epi = epi; a.act = S2S_ACT_ELU;
        rc = tc3_launch(map, a, p, npass, 0, flip ? "conv3x3_dgrad_tf32" : "conv3x3_fwd_tf32", st);
    }
    const cudaError_t ce = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (rc == 0 && ce != cudaSuccess) return fail(S2S_ERR_CUDA, "tc3conv: %s", cudaGetErrorString(ce));
    return rc;
}
int s2s_op_conv3x3_fwd_tf32(const float* x, const float* w, const float* b, float* y, int N, int H, int W, int Cin, int Cout,
                            int apply_elu, int npass, void* stream) {
    return op_tc3(x, w, b, nullptr, y, N, H, W, Cin, Cout, 0, apply_elu ? T3_EPI_BIAS_ACT : T3_EPI_BIAS, npass, (cudaStream_t)stream);
}
int s2s_op_conv3x3_dgrad_tf32(const float* dz, const float* w, const float* act, float* dx, int N, int H, int W, int Cin, int Cout,
                              int npass, void* stream) {
    return op_tc3(dz, w, nullptr, act, dx, N, H, W, Cout, Cin, 1, act ? T3_EPI_ACTGRAD : T3_EPI_NONE, npass, (cudaStream_t)stream);
}
int s2s_ffma_peak(float* scalar_tflops, float* packed_tflops, void* stream) {
    S2S_REQUIRE(scalar_tflops && packed_tflops, "null");
    S2S_CHECK(ffma_peak_measure(0, scalar_tflops, (cudaStream_t)stream));
    return ffma_peak_measure(1, packed_tflops, (cudaStream_t)stream);
}

int s2s_op_conv3x3_dgrad(const float* dz, const float* w, const float* act, float* dx, int N, int H, int W, int Cin, int Cout, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P = (size_t)9 * Cin * Cout;
    char* tmp = nullptr;
    S2S_CUDA(cudaMalloc((void**)&tmp, P * sizeof(float) + sizeof(WPrepEntry)));
    float* wt = reinterpret_cast<float*>(tmp);
    WPrepEntry e{0, Cout, Cin, 9, 1};
    WPrepEntry* e_dev = reinterpret_cast<WPrepEntry*>(tmp + P * sizeof(float));
    cudaMemcpyAsync(e_dev, &e, sizeof e, cudaMemcpyHostToDevice, st);
    wprep_kernel<<<dim3(std::min(cdiv((int)P, 256), 32), 1), 256, 0, st>>>(e_dev, w, wt);
    GConvArgs a;
    memset(&a, 0, sizeof a);
    a.in = dz; a.ldin = Cout; a.Hin = H; a.Win = W; a.Cb = Cout;
    a.w = wt;
    a.out = dx; a.ldout = Cin; a.Hout = H; a.Wout = W; a.Ca = Cin;
    a.pad = 1; a.N = N;
    if (act) { a.epi = EPI_ELUGRAD; a.aux = act; a.ldaux = Cin; } else a.epi = EPI_NONE;
    const int rc = gconv_run(3, 1, true, a, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    return rc;
}
static int op_wgrad_finish(float* part, float* bpart, const WgradPlan& p, int64_t P, int Cbias, float* dw, float* db, cudaStream_t st) {
    int rc = reduce_partials(part, dw, P, p.nslots, st);
    if (rc == 0 && db) rc = reduce_partials(bpart, db, Cbias, p.nslots, st);
    cudaStreamSynchronize(st);
    cudaFree(part);
    return rc;
}
int s2s_op_conv3x3_wgrad_tf32(const float* x, const float* dz, float* dw, float* db, int N, int H, int W, int Cin, int Cout, int n_max,
                              void* stream) {
    // n_max >= N sizes the plan (tile geometry, slots) like a handle created for max_batch = n_max; the tensors hold n_max images
    const TcWgPlan pw = tcwg_plan(H, W, Cin, Cout, n_max > N ? n_max : N);
    S2S_REQUIRE(pw.ok, "conv3x3_wgrad_tf32: needs Cin %% 4 == 0 and Cout %% 4 == 0 (got %d -> %d)", Cin, Cout);
    const int64_t P = (int64_t)9 * Cin * Cout;
    float* part = nullptr;
    S2S_CUDA(cudaMalloc((void**)&part, (size_t)pw.nslots * (P + Cout) * sizeof(float)));
    float* bpart = part + (size_t)pw.nslots * P;
    CUtensorMap mx, mz;
    int rc = tcwg_make_maps(pw, x, Cin, dz, Cout, N, H, W, Cin, Cout, &mx, &mz);
    if (rc == 0) rc = tcwg_launch(mx, mz, pw, part, bpart, N, H, W, Cin, Cout, pw.nslots, (cudaStream_t)stream);
    if (rc) { cudaFree(part); return rc; }
    WgradPlan p{};
    p.nslots = pw.nslots;
    return op_wgrad_finish(part, bpart, p, P, Cout, dw, db, (cudaStream_t)stream);
}
int s2s_op_conv3x3_wgrad(const float* x, const float* dz, float* dw, float* db, int N, int H, int W, int Cin, int Cout, void* stream) {
    const WgradPlan p = wgrad_plan(H, W, Cout, Cin, N);
    const int64_t P = (int64_t)9 * Cin * Cout;
    float* part = nullptr;
    S2S_CUDA(cudaMalloc((void**)&part, (size_t)p.nslots * (P + Cout) * sizeof(float)));
    float* bpart = part + (size_t)p.nslots * P;
    WgradArgs a;
    memset(&a, 0, sizeof a);
    a.A = dz; a.ldA = Cout; a.HA = H; a.WA = W; a.Ca = Cout;
    a.B = x; a.ldB = Cin; a.HB = H; a.WB = W; a.Cb = Cin;
    a.pad = 1; a.N = N; a.part = part; a.bias_part = bpart;
    int rc = wgrad_run(3, 1, a, p.nslots, (cudaStream_t)stream);
    if (rc) { cudaFree(part); return rc; }
    return op_wgrad_finish(part, bpart, p, P, Cout, dw, db, (cudaStream_t)stream);
}
int s2s_op_convt_fwd(const float* x, const float* w, const float* b, float* y, int N, int hh, int ww, int Cin, int Cout, int k, void* stream) {
    S2S_REQUIRE(k == 2 || k == 3 || k == 5, "ct_kernel must be 2, 3 or 5");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P = (size_t)k * k * Cin * Cout;
    char* tmp = nullptr;
    S2S_CUDA(cudaMalloc((void**)&tmp, P * sizeof(float) + sizeof(WPrepEntry)));
    float* wt = reinterpret_cast<float*>(tmp);
    WPrepEntry e{0, Cin, Cout, k * k, 0};
    WPrepEntry* e_dev = reinterpret_cast<WPrepEntry*>(tmp + P * sizeof(float));
    cudaMemcpyAsync(e_dev, &e, sizeof e, cudaMemcpyHostToDevice, st);
    wprep_kernel<<<dim3(std::min(cdiv((int)P, 256), 32), 1), 256, 0, st>>>(e_dev, w, wt);
    ConvTArgs a;
    memset(&a, 0, sizeof a);
    a.x = x; a.ldx = Cin; a.h = hh; a.w = ww; a.Cin = Cin; a.wt = wt; a.bias = b;
    a.y = y; a.ldy = Cout; a.y_coff = 0; a.Cout = Cout; a.N = N;
    const int rc = convt_fwd(a, k, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    return rc;
}
int s2s_op_convt_dgrad(const float* dy, const float* w, float* dx, int N, int hh, int ww, int Cin, int Cout, int k, void* stream) {
    GConvArgs a;
    memset(&a, 0, sizeof a);
    a.in = dy; a.ldin = Cout; a.Hin = 2 * hh; a.Win = 2 * ww; a.Cb = Cout;
    a.w = w;
    a.out = dx; a.ldout = Cin; a.Hout = hh; a.Wout = ww; a.Ca = Cin;
    a.pad = (k - 2) / 2; a.epi = EPI_NONE; a.N = N;
    if (k == 2) return gconv_run(2, 2, false, a, (cudaStream_t)stream);
    if (k == 3) return gconv_run(3, 2, false, a, (cudaStream_t)stream);
    if (k == 5) return gconv_run(5, 2, false, a, (cudaStream_t)stream);
    return fail(S2S_ERR_INVALID, "ct_kernel must be 2, 3 or 5");
}
static int op_tc3_convt(const float* in, const float* w, const float* bias, float* out, int N, int hh, int ww, int Cin, int Cout, int k, bool dgrad,
                        cudaStream_t st) {
    S2S_REQUIRE(k == 2 || k == 3 || k == 5, "ct_kernel must be 2, 3 or 5");
    S2S_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0 && Cin >= 8 && Cout >= 8, "tf32 tensor-core transposed conv needs channel counts %% 4 == 0 and >= 8 (got %d -> %d)", Cin, Cout);
    const Tc3Plan p = dgrad ? tc3_plan_convt_dgrad(tc3_plan_for(hh, ww, N, Cout, Cin, 1)) : tc3_plan_convt_fwd(tc3_plan_for(hh, ww, N, Cin, Cout, 1));
    S2S_REQUIRE(p.ok, "tf32 tensor-core transposed conv: no plan for %d -> %d", Cin, Cout);
    char* tmp = nullptr;
    S2S_CUDA(cudaMalloc((void**)&tmp, p.wq_floats * sizeof(float) + 256));
    float* wq = reinterpret_cast<float*>(tmp + 256);
    Tc3WPrep e{0, 0, dgrad ? Cout : Cin, dgrad ? Cin : Cout, k, p.NT, p.nchunks_n, p.CK, p.kchunks, dgrad ? 3 : 2, 1};
    cudaMemcpyAsync(tmp, &e, sizeof e, cudaMemcpyHostToDevice, st);
    tc3_wprep_kernel<<<dim3(64, 1), 256, 0, st>>>(reinterpret_cast<const Tc3WPrep*>(tmp), w, wq);
    Tc3Maps ms;
    int rc = dgrad ? tc3_make_maps_parity(in, N, hh, ww, Cout, Cout, p.CK, &ms) : tc3_make_map_any(in, N, hh, ww, Cin, Cin, p.CK, &ms.m[0]);
    if (rc == 0) {
        if (!dgrad) for (int i = 1; i < 4; ++i) ms.m[i] = ms.m[0];
        Tc3Args a;
        memset(&a, 0, sizeof a);
        a.wq = wq; a.bias = bias; a.in = in; a.out = out; a.N = N; a.H = hh; a.W = ww; a.act = S2S_ACT_ELU;
        if (dgrad) { a.ldin = Cout; a.ldout = Cin; a.Cin = 4 * Cout; a.Cout = Cin; a.epi = T3_EPI_NONE; a.kpp = p.kchunks / 4; }
        else { a.ldin = Cin; a.ldout = Cout; a.Cin = Cin; a.Cout = Cout; a.epi = T3_EPI_BIAS; a.up = 1; }
        for (int par = 0; par < 4; ++par) a.tapmask[par] = tc3_convt_tapmask(k, par, dgrad);
        rc = tc3_launch_maps(ms, a, p, 1, 0, dgrad ? "convT_dgrad_tf32" : "convT_fwd_tf32", st);
    }
    const cudaError_t ce = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (rc == 0 && ce != cudaSuccess) return fail(S2S_ERR_CUDA, "tc3conv (transposed): %s", cudaGetErrorString(ce));
    return rc;
}
int s2s_op_convt_fwd_tf32(const float* x, const float* w, const float* b, float* y, int N, int hh, int ww, int Cin, int Cout, int k, void* stream) {
    return op_tc3_convt(x, w, b, y, N, hh, ww, Cin, Cout, k, false, (cudaStream_t)stream);
}
int s2s_op_convt_dgrad_tf32(const float* dy, const float* w, float* dx, int N, int hh, int ww, int Cin, int Cout, int k, void* stream) {
    return op_tc3_convt(dy, w, nullptr, dx, N, hh, ww, Cin, Cout, k, true, (cudaStream_t)stream);
}
int s2s_op_convt_wgrad_tf32(const float* x, const float* dy, float* dw, int N, int hh, int ww, int Cin, int Cout, int k, int n_max, void* stream) {
    S2S_REQUIRE(k == 2 || k == 3 || k == 5, "ct_kernel must be 2, 3 or 5");
    const TcWgPlan pw = tcwg_plan(hh, ww, Cout, Cin, n_max > N ? n_max : N, 0, 4);
    S2S_REQUIRE(pw.ok, "convt_wgrad_tf32: needs Cin %% 4 == 0 and Cout %% 4 == 0 (got %d -> %d)", Cin, Cout);
    const int64_t P = (int64_t)k * k * Cin * Cout;
    float* part = nullptr;
    S2S_CUDA(cudaMalloc((void**)&part, (size_t)pw.nslots * P * sizeof(float)));
    Tc3Maps mx;
    CUtensorMap mz;
    int rc = tcwg_make_maps_convt(pw, x, Cin, dy, Cout, N, hh, ww, Cin, Cout, &mx, &mz);
    if (rc == 0) rc = tcwg_launch_maps(mx, mz, pw, part, nullptr, N, hh, ww, Cout, Cin, pw.nslots, k, (cudaStream_t)stream);
    if (rc) { cudaFree(part); return rc; }
    WgradPlan p{};
    p.nslots = pw.nslots;
    return op_wgrad_finish(part, nullptr, p, P, 0, dw, nullptr, (cudaStream_t)stream);
}
int s2s_op_convt_wgrad(const float* x, const float* dy, float* dw, float* db, int N, int hh, int ww, int Cin, int Cout, int k, void* stream) {
    S2S_REQUIRE(k == 2 || k == 3 || k == 5, "ct_kernel must be 2, 3 or 5");
    cudaStream_t st = (cudaStream_t)stream;
    const WgradPlan p = wgrad_plan(hh, ww, Cin, Cout, N);
    const int64_t P = (int64_t)k * k * Cin * Cout;
    float* part = nullptr;
    S2S_CUDA(cudaMalloc((void**)&part, ((size_t)p.nslots * P + 128 * (size_t)Cout + 64) * sizeof(float)));
    WgradArgs a;
    memset(&a, 0, sizeof a);
    a.A = x; a.ldA = Cin; a.HA = hh; a.WA = ww; a.Ca = Cin;
    a.B = dy; a.ldB = Cout; a.HB = 2 * hh; a.WB = 2 * ww; a.Cb = Cout;
    a.pad = (k - 2) / 2; a.N = N; a.part = part; a.bias_part = nullptr;
    int rc = k == 2 ? wgrad_run(2, 2, a, p.nslots, st) : k == 3 ? wgrad_run(3, 2, a, p.nslots, st) : wgrad_run(5, 2, a, p.nslots, st);
    if (rc == 0 && db) {
        ChanSumArgs cs;
        memset(&cs, 0, sizeof cs);
        cs.g = dy; cs.ld = Cout; cs.coff = 0; cs.C = Cout; cs.npix = (int64_t)N * 4 * hh * ww;
        cs.part = part + (size_t)p.nslots * P;
        cs.nslots = bn_bwd_slots(cs.npix, 256 / bn_cqb(Cout));
        rc = chansum(cs, st);
        if (rc == 0) rc = reduce_partials(cs.part, db, Cout, cs.nslots, st);
    }
    if (rc) { cudaFree(part); return rc; }
    return op_wgrad_finish(part, nullptr, p, P, 0, dw, nullptr, st);
}

}  // extern "C"
