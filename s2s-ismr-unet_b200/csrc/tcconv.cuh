// tcconv.cuh — Conv2D 3x3 'same' + bias + ELU as an implicit GEMM on the 5th-generation tensor cores
// (tcgen05.mma, accumulator in TMEM, operands staged by TMA) for the THICK layers (Cin % 64 == 0).
//
//   D[m, co] = sum_{tap, ci} A_tap[m, ci] * B_tap[co, ci]      m = pixel of a 16 x 8 output tile (M = 128)
//
//   * A_tap is not materialised: for every tap one TMA tiled load of the box (C=64, W=8, H=16, N=1) at the
//     tap-shifted coordinates (x0+kx-1, y0+ky-1) lands in shared memory as 128 rows x 128 B, which IS the
//     canonical K-major SWIZZLE_128B UMMA operand layout (8-row atoms of 1024 B); out-of-image rows/cols are
//     zero-filled by the TMA unit = the 'same' padding.  Inputs are bf16 NHWC (cast kernel below).
//   * B_tap = weights as bf16 [tap][co][ci] (K-major), box (64, Cout, 1).
//   * 4-stage mbarrier ring: warp 0 / lane 0 issues TMA, warp 1 / lane 0 issues 4 x tcgen05.mma (K = 16 each) per
//     stage and tcgen05.commit's the stage back to the producer; after the last k-block the accumulator
//     (128 lanes x Cout fp32 columns of TMEM) is committed to the epilogue barrier.
//   * epilogue: all 4 warps tcgen05.ld their 32 TMEM lanes (32x32b.x8), add bias, ELU, store fp32 NHWC.
// Precision: inputs/weights rounded to bf16, fp32 accumulation and output -> rel-L2 ~2e-3 (BASELINE bf16
// tolerance 1e-2).  The strict-fp32 path (gconv.cuh) stays the parity path for training.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace s2s {

constexpr int TC_TH = 16, TC_TW = 8, TC_M = 128, TC_STAGES = 4;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major swizzled shared-memory matrix descriptor, version 1 (Blackwell).  KC = 64 bf16: 128-byte rows,
// SWIZZLE_128B (layout type 2), 8-row atoms 1024 B apart; KC = 32: 64-byte rows, SWIZZLE_64B (type 4), atoms 512 B.
template <int KC>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
    constexpr uint64_t sbo = (8 * KC * 2) >> 4, lt = (KC == 64) ? 2 : 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (lt << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TcConvArgs {
    const float* bias;
    float* out; int ldout, out_coff;
    int N, H, W, Cin, Cout, tiles_x, tiles_y, apply_elu, act;
};

template <int NCOLS, int TC_KC>   // TMEM columns = padded Cout (32, 64, 128, 256); channels per k-block (64 | 32)
__global__ void __launch_bounds__(128) tcconv_kernel(const __grid_constant__ CUtensorMap map_a,
                                                     const __grid_constant__ CUtensorMap map_b, const TcConvArgs a) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    constexpr int A_BYTES = TC_M * TC_KC * 2;                 // 16 KB
    const int b_bytes = a.Cout * TC_KC * 2;
    const int stage_bytes = A_BYTES + ((b_bytes + 1023) / 1024) * 1024;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, n = blockIdx.y;
    const int y0 = (tile / a.tiles_x) * TC_TH, x0 = (tile % a.tiles_x) * TC_TW;
    const int kchunks = a.Cin / TC_KC;
    const int kit = 9 * kchunks;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(NCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer
        for (int it = 0; it < kit; ++it) {
            const int s = it % TC_STAGES;
            mbar_wait(&empty_bar[s], ((it / TC_STAGES) & 1) ^ 1);
            const int tap = it / kchunks, kc = it % kchunks;
            const int ky = tap / 3, kx = tap % 3;
            uint8_t* sa = base + s * stage_bytes;
            mbar_expect_tx(&full_bar[s], A_BYTES + b_bytes);
            tma_load_4d(sa, &map_a, &full_bar[s], kc * TC_KC, x0 + kx - 1, y0 + ky - 1, n);
            tma_load_3d(sa + A_BYTES, &map_b, &full_bar[s], kc * TC_KC, 0, tap);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer: instruction descriptor = F32 accum, BF16 x BF16, K-major A and B, N = Cout, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.Cout >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        for (int it = 0; it < kit; ++it) {
            const int s = it % TC_STAGES;
            mbar_wait(&full_bar[s], (it / TC_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(base + s * stage_bytes), sb = sa + A_BYTES;
#pragma unroll
            for (int k = 0; k < TC_KC / 16; ++k)
                umma_bf16(tmem_base, umma_desc_kmajor<TC_KC>(sa + k * 32), umma_desc_kmajor<TC_KC>(sb + k * 32), idesc, (it | k) != 0);
            umma_commit(&empty_bar[s]);      // implies tcgen05.fence::before_thread_sync
        }
        umma_commit(&acc_bar);
    }

    // ===== epilogue: every warp reads its 32 TMEM lanes (rows 32*warp .. 32*warp+31 of the tile)
    __syncwarp();
    mbar_wait(&acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = 32 * warp + lane;
    const int oy = y0 + m / TC_TW, ox = x0 + m % TC_TW;
    const bool inside = oy < a.H && ox < a.W;
    float* orow = a.out + (((size_t)n * a.H + (inside ? oy : 0)) * a.W + (inside ? ox : 0)) * a.ldout + a.out_coff;
    for (int c0 = 0; c0 < a.Cout; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (inside) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = __uint_as_float(r[j]) + (a.bias ? __ldg(a.bias + c0 + j) : 0.f);
                if (a.apply_elu) v[j] = act_f(v[j], a.act);
            }
            st4(orow + c0, make_float4(v[0], v[1], v[2], v[3]));
            st4(orow + c0 + 4, make_float4(v[4], v[5], v[6], v[7]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NCOLS) : "memory");
}

// fp32 -> bf16 casts (activations; weights transposed to [tap][co][ci])
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 4 <= n) {
        const float4 v = ld4(in + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<__nv_bfloat162*>(out + i)[0] = lo;
        reinterpret_cast<__nv_bfloat162*>(out + i)[1] = hi;
    } else {
        for (int64_t k = i; k < n; ++k) out[k] = __float2bfloat16(in[k]);
    }
}
__global__ void wprep_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt, int Cin, int Cout) {
    const int total = 9 * Cin * Cout;                  // dst [tap][co][ci]  <-  src [tap][ci][co]
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % Cin, co = (i / Cin) % Cout, tap = i / (Cin * Cout);
        wt[i] = __float2bfloat16(__ldg(w + ((size_t)tap * Cin + ci) * Cout + co));
    }
}

// ------------------------------------------------------------------ host: tensor maps + launch
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline PFN_tmapEncodeTiled tmap_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

static inline bool tcconv_eligible(int Cin, int Cout) { return Cin % 32 == 0 && Cout % 16 == 0 && Cout >= 16 && Cout <= 256; }
static inline int tcconv_kc(int Cin) { return Cin % 64 == 0 ? 64 : 32; }

static inline int tcconv_make_maps(const __nv_bfloat16* x, const __nv_bfloat16* wt, int N, int H, int W, int Cin, int Cout,
                                   CUtensorMap* ma, CUtensorMap* mb) {
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint32_t TC_KC = (cuuint32_t)tcconv_kc(Cin);
    const CUtensorMapSwizzle swz = TC_KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {TC_KC, TC_TW, TC_TH, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 9};
        cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
        cuuint32_t box[3] = {TC_KC, (cuuint32_t)Cout, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)wt, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
    return 0;
}

static inline int tcconv_launch(const CUtensorMap& ma, const CUtensorMap& mb, TcConvArgs a, cudaStream_t st) {
    S2S_REQUIRE(tcconv_eligible(a.Cin, a.Cout), "tcconv: needs Cin %% 32 == 0 and Cout %% 16 == 0 (<= 256), got %d -> %d", a.Cin, a.Cout);
    const int TC_KC = tcconv_kc(a.Cin);
    S2S_REQUIRE((a.ldout & 3) == 0 && (a.out_coff & 3) == 0, "tcconv: output stride must be a multiple of 4");
    a.tiles_x = cdiv(a.W, TC_TW); a.tiles_y = cdiv(a.H, TC_TH);
    const int b_bytes = a.Cout * TC_KC * 2;
    const size_t smem = 1024 + (size_t)TC_STAGES * (TC_M * TC_KC * 2 + ((b_bytes + 1023) / 1024) * 1024);
    dim3 grid(a.tiles_x * a.tiles_y, a.N);
    const int ncols = a.Cout <= 32 ? 32 : a.Cout <= 64 ? 64 : a.Cout <= 128 ? 128 : 256;
    prof_begin(st, "conv3x3_fwd_tcgen05", 2.0 * a.N * a.H * a.W * a.Cin + 4.0 * a.N * a.H * a.W * a.Cout,
               18.0 * (double)a.Cin * a.Cout * a.N * a.H * a.W);
#define S2S_TC_LAUNCH(NC, KCV)                                                                                     \
    {                                                                                                              \
        static DevOnce once;                                                                                       \
        S2S_CUDA(once.run([] { return cudaFuncSetAttribute(tcconv_kernel<NC, KCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); })); \
        tcconv_kernel<NC, KCV><<<grid, 128, smem, st>>>(ma, mb, a);                                                \
    }
#define S2S_TC_BY_N(KCV)                                                                                           \
    if (ncols == 32) S2S_TC_LAUNCH(32, KCV) else if (ncols == 64) S2S_TC_LAUNCH(64, KCV) else if (ncols == 128) S2S_TC_LAUNCH(128, KCV) else S2S_TC_LAUNCH(256, KCV)
    if (TC_KC == 64) { S2S_TC_BY_N(64) } else { S2S_TC_BY_N(32) }
#undef S2S_TC_BY_N
#undef S2S_TC_LAUNCH
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2s
