// common.cuh — shared device/host helpers for the sm_100a U-Net hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <atomic>
#include <mutex>
#include <vector>
#include "../../include/s2s_unet.h"

namespace s2s {

// ---------------------------------------------------------------- error plumbing
inline std::string& last_error_ref() {
    static thread_local std::string e;
    return e;
}
inline int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}
#define S2S_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::s2s::fail(S2S_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,     \
                               cudaGetErrorString(_e));                                       \
    } while (0)
#define S2S_CHECK(expr)                                                                       \
    do {                                                                                      \
        int _s = (expr);                                                                      \
        if (_s != 0) return _s;                                                               \
    } while (0)
#define S2S_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) return ::s2s::fail(S2S_ERR_INVALID, __VA_ARGS__);                        \
    } while (0)
#define S2S_LAUNCH_CHECK() S2S_CUDA(cudaGetLastError())

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// One-shot-per-device guard for cudaFuncSetAttribute: the attribute is PER DEVICE and one process may drive several GPUs
// (Model(device=k)); a process-wide `static bool` left launches that need > 48 KB of shared memory failing on the second
// device.  Usage:  static DevOnce once;  S2S_CUDA(once.run([&] { return cudaFuncSetAttribute(...); }));
struct DevOnce {
    std::mutex mu;
    uint64_t done = 0;
    template <typename F>
    cudaError_t run(F&& f) {
        int d = 0;
        cudaGetDevice(&d);
        const uint64_t bit = 1ull << (d & 63);
        std::lock_guard<std::mutex> lk(mu);
        if (done & bit) return cudaSuccess;
        const cudaError_t e = f();
        if (e == cudaSuccess) done |= bit;
        return e;
    }
};

// launch counter (bench: gpu_launches); one per process is enough for a claim
inline std::atomic<int64_t>& launch_counter() {      // atomic: tuning trials run on several host threads
    static std::atomic<int64_t> n{0};
    return n;
}

// ---------------------------------------------------------------- per-launch profiler (bench.py roofline)
// When enabled every kernel launch is bracketed by CUDA events on its own stream and tagged with
// its ALGORITHMIC bytes / flops; graphs are bypassed while it is on.
struct ProfRec { const char* tag; double bytes, flops; cudaEvent_t e0, e1; };
struct Profiler {
    bool on = false;
    std::vector<ProfRec> recs;
};
inline Profiler& prof() {
    static Profiler p;
    return p;
}
inline void prof_begin(cudaStream_t st, const char* tag, double bytes, double flops) {
    Profiler& p = prof();
    if (!p.on) return;
    ProfRec r{tag, bytes, flops, nullptr, nullptr};
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    p.recs.push_back(r);
}
inline void prof_end(cudaStream_t st) {
    launch_counter()++;
    Profiler& p = prof();
    if (!p.on || p.recs.empty()) return;
    cudaEventRecord(p.recs.back().e1, st);
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The kernels of a step form one long dependent chain of sub-wave grids (batch 16): with a plain stream / graph edge
// kernel N+1 is not even scheduled before kernel N has drained.  Critical-path kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (also captured into the CUDA graphs as programmatic edges) and are
// written in three parts:
//   prologue   everything that does not depend on the preceding kernel — address arithmetic, barrier / TMEM set-up and the
//              WEIGHT slab (cp.async / cp.async.bulk), which was last written at least two kernels earlier;
//   pdl_wait() griddepcontrol.wait: the preceding grid has completed and its writes are visible; only now the input tile
//              (the preceding kernel's output) is read;
//   pdl_trigger() griddepcontrol.launch_dependents AFTER the main loop: the next kernel is scheduled while this one is
//              down to its epilogue, so its prologue overlaps this kernel's tail.
// Measured on B200 (profiles/r2_summary.md): batch 16 fp32 428 -> 390 us per step, tf32 385 -> 344 us; batch 128 fp32
// 1659 -> 1489 us.  (Round 1 triggered at the very top of every kernel and waited before any load: no useful overlap and
// -5 %, which is why it was off then.)  S2S_PDL=0 disables it.
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("S2S_PDL"); return !e || e[0] != '0'; }();
    return on;
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);       // errors surface through S2S_LAUNCH_CHECK()
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------- device helpers
// ELU(alpha=1): x > 0 ? x : exp(x) - 1      (Keras activation='elu', deep_nn_models.py:142)
// Branch-free: a degree-5 Taylor polynomial for x > -1/8 (truncation error x^6/720 < 6e-9 relative there, so the
// small outputs keep their relative accuracy) and exp2(x * log2 e) - 1 on the SFU (MUFU.EX2, relative error 2^-22
// on a value in (0, 1] -> absolute error <= 2.4e-7) below it.  expm1f's divergent slow path cost 3x the convolution
// arithmetic of the thin layers (profiles/r1_summary.md) and the full-range expf still 11 % of the conv kernel's
// instructions at batch 128 (profiles/r1c_*).
__device__ __forceinline__ float elu_f(float x) {
    const float xn = fminf(x, 0.f);
    const float big = __expf(xn) - 1.f;
    const float small = xn * (1.f + xn * (0.5f + xn * (0.16666667f + xn * (0.041666668f + xn * 0.0083333338f))));
    const float neg = xn > -0.125f ? small : big;
    return x > 0.f ? x : neg;
}
// dELU/dx expressed through the OUTPUT y = ELU(x):  y > 0 ? 1 : y + 1   (exp(x) = y + 1)
__device__ __forceinline__ float elu_grad_from_out(float y) { return y > 0.f ? 1.f : y + 1.f; }

// hidden-layer activation selected by s2s_unet_cfg.act (S2S_ACT_ELU | S2S_ACT_RELU) and its derivative from the OUTPUT
__device__ __forceinline__ float act_f(float x, int act) { return act == S2S_ACT_ELU ? elu_f(x) : fmaxf(x, 0.f); }
__device__ __forceinline__ float act_grad_from_out(float y, int act) { return y > 0.f ? 1.f : (act == S2S_ACT_ELU ? y + 1.f : 0.f); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// "Last CTA done" election (deterministic two-stage reductions without a second launch).
// Every thread must have issued its global partial writes before calling.  Returns true in
// exactly one CTA of the grid: the one that arrives last.  That CTA may then read all
// partials (with __ldcg) and must not rely on L1.  The counter is reset for the next launch.
__device__ __forceinline__ bool cta_is_last(unsigned int* counter, unsigned int total_ctas) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    if (tid == 0) {
        unsigned int prev = atomicAdd(counter, 1u);
        s_last = (prev == total_ctas - 1u);
        if (s_last) *counter = 0u;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// Fixed-order reduction of `nslots` partial vectors by the whole CTA (used by the elected last CTA).
// Element v of slot s lives at part[s * slot_stride + v].  The nvals totals (double) are left in
// sd_out[0..nvals) (shared memory, nvals <= NT); sd_tmp needs NT doubles.  All NT threads must call.
// Thread (g, v) sums slots g, g+G, ... (G = NT / nvals groups), then thread v adds the G group sums
// in group order: the summation order depends only on (nslots, nvals, NT) -> bit-reproducible.
template <int NT>
__device__ __forceinline__ void cta_reduce_slots(const float* part, int nslots, size_t slot_stride, int nvals,
                                                 double* sd_tmp, double* sd_out, int tid) {
    const int G = NT / nvals;
    const int g = tid / nvals, v = tid - g * nvals;
    double s = 0.0;
    if (g < G) {
        int sl = g;
        for (; sl + 7 * G < nslots; sl += 8 * G) {       // 8 independent L2 loads in flight, summed in slot order
            float t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = __ldcg(part + (size_t)(sl + k * G) * slot_stride + v);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += (double)t[k];
        }
        for (; sl < nslots; sl += G) s += (double)__ldcg(part + (size_t)sl * slot_stride + v);
        sd_tmp[g * nvals + v] = s;
    }
    __syncthreads();
    if (tid < nvals) {
        double t = 0.0;
        for (int k = 0; k < G; ++k) t += sd_tmp[k * nvals + tid];
        sd_out[tid] = t;
    }
    __syncthreads();
}

// Fixed-order reduction of a CONTIGUOUS block of partials [nslots][V] (V % 4 == 0, Q = V/4 a power of two <= 256) by a
// 256-thread CTA: thread t owns the float4 chunks t, t+256, ... (always the same channel quad, since Q divides 256) with
// 8 independent 16-byte loads in flight, lanes sharing a quad are combined by an xor butterfly, the 8 warps through
// shared memory.  One L2 round trip instead of one per 8 slots and value (cta_reduce_slots): the BatchNorm finalize
// sits on the critical path of every training step.  sd4 needs 1024 doubles, sd_out V doubles.  All threads call.
__device__ __forceinline__ bool cta_reduce_block_ok(int V) {
    const int Q = V >> 2;
    return (V & 3) == 0 && Q >= 1 && Q <= 256 && (Q & (Q - 1)) == 0;
}
__device__ __forceinline__ void cta_reduce_block256(const float* part, int nslots, int V, double* sd4, double* sd_out, int tid) {
    const int Q = V >> 2;
    const int total4 = nslots * Q;
    const float4* p4 = reinterpret_cast<const float4*>(part);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = tid;
    for (; i + 7 * 256 < total4; i += 8 * 256) {
        float4 t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = __ldcg(p4 + i + k * 256);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s0 += (double)t[k].x; s1 += (double)t[k].y; s2 += (double)t[k].z; s3 += (double)t[k].w; }
    }
    for (; i < total4; i += 256) {
        const float4 t = __ldcg(p4 + i);
        s0 += (double)t.x; s1 += (double)t.y; s2 += (double)t.z; s3 += (double)t.w;
    }
    for (int o = 16; o >= Q; o >>= 1) {            // lanes l, l + Q, l + 2Q, ... hold the same channel quad
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    sd4[tid * 4 + 0] = s0; sd4[tid * 4 + 1] = s1; sd4[tid * 4 + 2] = s2; sd4[tid * 4 + 3] = s3;
    __syncthreads();
    for (int v = tid; v < V; v += 256) {
        const int q = v >> 2, c = v & 3;
        double t = 0.0;
        if (Q < 32) {
#pragma unroll 1
            for (int w = 0; w < 8; ++w) t += sd4[(w * 32 + q) * 4 + c];
        } else {
#pragma unroll 1
            for (int k = q; k < 256; k += Q) t += sd4[k * 4 + c];
        }
        sd_out[v] = t;
    }
    __syncthreads();
}

}  // namespace s2s
