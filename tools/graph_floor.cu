// Micro-benchmark: per-node cost of a dependent kernel chain inside a CUDA graph on this GPU,
// with and without programmatic dependent launch (PDL), and with cp.async/smem-heavy configs.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_plain(float* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = p[i] * 1.0001f + 1.f; }
__global__ void k_pdl(float* p, int n) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = p[i] * 1.0001f + 1.f;
}
template <bool PDL>
float run(int nodes, int grid, int block, size_t smem, float* d, int n) {
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaGraph_t g; cudaGraphExec_t ex;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < nodes; ++i) {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
        if (PDL) cudaLaunchKernelEx(&cfg, k_pdl, d, n); else cudaLaunchKernelEx(&cfg, k_plain, d, n);
    }
    cudaStreamEndCapture(st, &g);
    cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
    if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) cudaGraphLaunch(ex, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < 50; ++i) cudaGraphLaunch(ex, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / 50 / nodes;
}
int main() {
    int n = 1 << 20; float* d; cudaMalloc(&d, n * 4); cudaMemset(d, 0, n * 4);
    cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_pdl, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    printf("us per node (70-node chain)\n");
    int grids[] = {1, 128, 512, 4096};
    for (int gi = 0; gi < 4; ++gi) {
        int gr = grids[gi];
        printf("grid %5d blk 256 smem 0    : plain %.2f  pdl %.2f\n", gr, run<false>(70, gr, 256, 0, d, n), run<true>(70, gr, 256, 0, d, n));
        printf("grid %5d blk 256 smem 64K  : plain %.2f  pdl %.2f\n", gr, run<false>(70, gr, 256, 64 * 1024, d, n), run<true>(70, gr, 256, 64 * 1024, d, n));
    }
    cudaError_t e = cudaDeviceSynchronize(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
