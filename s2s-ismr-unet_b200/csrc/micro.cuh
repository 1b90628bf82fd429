// micro.cuh — fp32 CUDA-core peak micro-benchmarks: the denominators of the "ffma" roofline in bench.py
// (BASELINE.md §2 asks for the MEASURED FFMA peak rather than the nominal 148 SMs x 128 lanes x 2 x clock).
//   variant 0: scalar FFMA   (fma.rn.f32, 3 register operands)
//   variant 1: packed FFMA2  (fma.rn.f32x2, sm_100: two fp32 FMAs per issued instruction)
// 16 independent accumulator chains per thread (latency 4 x issue 2..4 well covered with 8 warps per scheduler).
#pragma once
#include "common.cuh"

namespace s2s {

template <int VARIANT>
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* __restrict__ out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (VARIANT == 1) {
                    acc[i] = __ffma2_rn(acc[i], a2, b2);
                } else {
                    acc[i].x = fmaf(acc[i].x, a2.x, b2.x);
                    acc[i].y = fmaf(acc[i].y, a2.y, b2.y);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    if (s == 123.456f) out[0] = s;      // never true: keeps the chains alive
}

// returns TFLOP/s (2 flop per FMA) of one variant, timed with CUDA events over `reps` launches after a warm-up
static inline int ffma_peak_measure(int variant, float* tflops, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* d = nullptr;
    S2S_CUDA(cudaMalloc((void**)&d, 16));
    const int iters = 2048, grid = sms * 8, reps = 5;
    cudaEvent_t e0, e1;
    S2S_CUDA(cudaEventCreate(&e0));
    S2S_CUDA(cudaEventCreate(&e1));
    float best = 0.f;
    for (int r = 0; r < reps + 1; ++r) {
        S2S_CUDA(cudaEventRecord(e0, st));
        if (variant == 1) ffma_peak_kernel<1><<<grid, 256, 0, st>>>(d, iters, 0.999f, 0.001f);
        else ffma_peak_kernel<0><<<grid, 256, 0, st>>>(d, iters, 0.999f, 0.001f);
        S2S_CUDA(cudaEventRecord(e1, st));
        S2S_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        S2S_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 16.0 * 8.0 * (double)iters * 256.0 * grid;     // 16 FMAs x 8 rounds per iteration and thread
        const float t = (float)(fl / (ms * 1e-3) * 1e-12);
        if (r > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    S2S_LAUNCH_CHECK();
    *tflops = best;
    return 0;
}

}  // namespace s2s
