// tc3conv.cuh — Conv2D 3x3 'same' (forward and input-gradient) as an implicit GEMM on the 5th-generation tensor
// cores for the TRAINING path: tcgen05.mma.kind::tf32 on the fp32 NHWC activations themselves, accumulators in TMEM,
// operands staged by TMA, with an error-compensated 3xTF32 split (NPASS = 3) that keeps fp32 parity
// (forward rel-L2 <= 1e-5 against the fp64 oracle) and a single-pass mode (NPASS = 1, ~5e-4) for precision="tf32".
//
//   D[m, co] = sum_{tap, ci} A[pixel(m) + tap, ci] * Wq[tap][co][ci]       m = y*8 + x of an 8 (x) x 16 (y) output tile
//
// One halo tile, nine taps.  The input tile with its halo (10 x 18 pixels) is loaded ONCE per channel chunk by a
// single 5-D TMA box (c4 = 4 channels, x = 10, y = 18, n = 1, cq = CK/4 channel quads) into the PLANAR layout
// [cq][y][x][4 floats]: out-of-image pixels are zero-filled by the TMA unit (= 'same' padding).  In that layout eight
// x-consecutive pixels of one channel quad are 128 contiguous bytes = one 8-row x 16-byte core matrix of the
// no-swizzle ("interleaved") K-major UMMA operand layout:
//     A descriptor:  start = tile + ((2j*18 + ky)*10 + kx)*16,  LBO (next K quad) = 180*16 B,  SBO (next 8 rows = next
//     image row) = 10*16 B
// so the nine filter taps are nine start addresses into the same tile — no im2col, no per-tap reload.
// Weights are pre-arranged per step (tc3_wprep_kernel) as [tap][cq][co][4] blocks (core matrix = 8 co x 16 B) and
// fetched with one cp.async.bulk per chunk.  (Conv2D layers: deep_nn_models.py:142,145,157,160.)
//
// 3xTF32: a = hi + lo with hi = rna_tf32(a), lo = a - hi (exact in fp32).  D0 += Ahi*Bhi, D1 += Alo*Bhi + Ahi*Blo in a
// second TMEM accumulator (the small terms are not rounded against the large sum), epilogue adds D0 + D1.  The
// transform warps split the staged activation tile in shared memory (hi in place, lo to a second buffer) and publish
// it to the async proxy; weights are split by the wprep kernel.
//
// Warp roles (192 threads): warp 0 = TMA / bulk-copy producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = operand transform, then epilogue (tcgen05.ld -> bias + activation | * act'(aux) -> NHWC stores,
// BatchNorm (sum, sumsq) partials per tile in fixed order).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "tcconv.cuh"

namespace s2s {

constexpr int T3_TH = 16, T3_TW = 8, T3_HH = 18, T3_HW = 10, T3_NPIX = T3_HH * T3_HW;
constexpr int T3_THREADS = 192, T3_MAXSTAGE = 4;
constexpr int T3_NISS = 3, T3_THREADS_V1 = 256;      // v1 kernel: three MMA-issuing warps (1, 6, 7)

enum { T3_EPI_BIAS_ACT = 0, T3_EPI_ACTGRAD = 1, T3_EPI_NONE = 2, T3_EPI_BIAS = 3 };

struct Tc3Args {
    const float* wq;               // [nchunks_n][kchunks][NPASS == 3 ? 2 : 1][9][CK/4][NT][4]   (tc3_wprep_kernel)
    const float* bias;             // [Cout] or null
    const float* aux; int ldaux;   // ACTGRAD: activation OUTPUT at the output positions
    float* out; int ldout, out_coff;
    float* stat_part;              // [slots][2][Cout] BatchNorm (sum, sumsq) partials, nullable
    int early;                     // epilogue operands (bias / aux) are older than the preceding kernel: fetch them ahead (latency regime)
    const float* in; int ldin, in_coff;   // only read by the LOADER = 1 (cooperative ld.global) variant
    int N, H, W, Cin, Cout, NT, kchunks, nstage, tmem_cols, tiles_x, tiles_y, epi, act;
    int w_early;                   // the weight blocks do not depend on the preceding kernel (programmatic dependent launch)
    // flat geometry (small images, H*W < 64): a tile = nimg whole zero-padded images, M row m = flat padded position
    // (img, py, px) = (m / (BX*BY), ..): the tap shift is the flat offset ky*BX + kx, rows whose (py, px) fall on the padding
    // are junk and never stored.  One 5-D TMA box (4, BX, BY, nimg, CK/4) at (x, y, n) = (-1, -1, n0).
    int flat, BX, BY, nimg;
    // Conv2DTranspose(k, strides 2, 'same') on the same kernel (single pass only):
    //   up = 1  forward: four stride-1 3x3 convolutions on the INPUT grid, one per output parity (a, b) = blockIdx.y /
    //           nchn; output pixel (2y + a, 2x + b) of the 2H x 2W image; weights [parity][n chunk][k chunk] blocks;
    //   kpp > 0 input gradient: the contraction runs over (parity plane, output channel): k chunk kc reads the TMA map of
    //           plane kc / kpp (the stride-2 sub-image dy[2y + a, 2x + b]) at channel quad (kc % kpp) * CK/4.
    // tapmask[parity]: the taps of the 3x3 window that exist for that parity (the others have all-zero weights: skipped).
    int up, kpp, nchn, tapmask[4];
};
struct Tc3Maps { CUtensorMap m[4]; };

// ------------------------------------------------------------------ PTX wrappers (beyond tcconv.cuh)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    // try_wait suspends in hardware; the counter only bounds a protocol bug to a trap instead of a hung GPU
    uint32_t done = 0;
    for (uint32_t it = 0; !done; ++it) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && it > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// no-swizzle K-major shared-memory matrix descriptor (Blackwell version 1): core matrix = 8 rows x 16 B contiguous;
// lbo = byte distance between the two 16-byte K chunks of one MMA (K = 8 tf32), sbo = byte distance between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// ------------------------------------------------------------------ kernel
// GEN = 0: the plain tiled 3x3 conv (tile geometry, tap set and accumulate flags are compile-time constants: the single MMA-issuing
// thread is the bottleneck of the thin layers, every runtime multiply in its loop shows — tf32 inference at 256x256 fell from 31.2k
// to 28.4k samples/s when the geometry became runtime); GEN = 1: flat geometry and the transposed-conv variants
// PRE: the epilogue's global operands are fetched one group ahead (latency regime; its own instantiation: the extra live values
// raise the kernel from 40-48 to 64-76 registers, which cost the throughput-bound launches 3-12 %)
template <int CK, int NPASS, int LOADER, int GEN, bool PRE = false>     // channels per chunk (8 | 16 | 32); 1 | 3 passes; 0 = TMA, 1 = ld.global
__global__ void __launch_bounds__(T3_THREADS_V1) tc3conv_kernel(const __grid_constant__ Tc3Maps maps, const Tc3Args a) {
    extern __shared__ __align__(128) uint8_t t3_smem[];
    constexpr int KQ = CK / 4;
    constexpr int A_BYTES = KQ * T3_NPIX * 16;
    constexpr int F = NPASS == 3 ? 2 : 1;
    const int NT = a.NT;
    const int b_bytes = 36 * CK * NT;                         // 9 taps x CK x NT x 4 B
    const int stage_bytes = F * (A_BYTES + b_bytes);
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(t3_smem) + 127) & ~(uintptr_t)127);
    __shared__ uint64_t full_bar[T3_MAXSTAGE], ready_bar[T3_MAXSTAGE], empty_bar[T3_MAXSTAGE], acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool g_flat = GEN && a.flat, g_up = GEN && a.up;
    const int g_kpp = GEN ? a.kpp : 0;
    const int tile = blockIdx.x, ycoord = blockIdx.y;
    const int nc = g_up ? ycoord % a.nchn : ycoord, par = g_up ? ycoord / a.nchn : 0;     // output-channel chunk, output parity
    const int n = g_flat ? tile * a.nimg : blockIdx.z;                       // (first) image of the tile
    const int y0 = g_flat ? 0 : (tile / a.tiles_x) * T3_TH, x0 = g_flat ? 0 : (tile % a.tiles_x) * T3_TW;
    const int kchunks = a.kchunks, nstage = a.nstage;
    constexpr bool kTransform = (NPASS == 3) || (LOADER == 1);
    const int ppi = a.BX * a.BY;                                             // flat: padded positions per image
    const int qpos = g_flat ? a.nimg * ppi : T3_NPIX;                        // positions per channel quad of the staged box
    const int a_tx = LOADER == 0 ? KQ * qpos * 16 : 0;                       // bytes the activation box delivers

    if (tid == 0) {
        for (int s = 0; s < T3_MAXSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], 128); mbar_init(&empty_bar[s], T3_NISS); }
        mbar_init(&acc_bar, T3_NISS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer: one TMA box (activations) + one bulk copy (weights, hi and lo) per channel chunk
            if (LOADER == 0) {
                for (int mi = 0; mi < (g_kpp ? 4 : 1); ++mi) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[mi]) : "memory");
            }
            // programmatic dependent launch: everything up to here (barriers, TMEM, descriptor) and the first weight block ran
            // while the preceding kernel drained; the activations are only touched after the wait
            if (a.w_early) {
                mbar_expect_tx(&full_bar[0], a_tx + F * b_bytes);
                bulk_g2s(base + F * A_BYTES, a.wq + (size_t)(ycoord * kchunks) * F * 9 * CK * NT, F * b_bytes, &full_bar[0]);
            }
            pdl_wait();
            for (int kc = 0; kc < kchunks; ++kc) {
                const int s = kc % nstage;
                mbar_wait_bounded(&empty_bar[s], ((kc / nstage) & 1) ^ 1);
                uint8_t* sa = base + s * stage_bytes;
                const bool w_done = a.w_early && kc == 0;
                if (!w_done) mbar_expect_tx(&full_bar[s], a_tx + F * b_bytes);
                if (LOADER == 0) {
                    const int mi = g_kpp ? kc / g_kpp : 0;                   // transposed-conv dgrad: parity plane of this k chunk
                    tma_load_5d(sa, &maps.m[mi], &full_bar[s], 0, x0 - 1, y0 - 1, n, (g_kpp ? kc - mi * g_kpp : kc) * KQ);
                }
                if (!w_done) bulk_g2s(sa + F * A_BYTES, a.wq + (size_t)(ycoord * kchunks + kc) * F * 9 * CK * NT, F * b_bytes, &full_bar[s]);
            }
        }
    } else if (warp == 1 || warp >= 6) {
        const int me = warp == 1 ? 0 : warp - 5;     // issuer 0, 1, 2
        if (lane == 0) {
            // ===== MMA issuers: THREE single-thread issuers, each with its OWN TMEM accumulator(s), so the order of the sums is
            // fixed whatever the interleaving of the three instruction streams (the epilogue adds the accumulators in a fixed
            // order).  One thread issuing all MMAs of a thin or small-image layer was the bottleneck (~14 SASS instructions and
            // ~90 cycles per MMA against ~50 cycles in the pipe).
            //   1 pass : issuer i takes the taps t with t % 3 == i into accumulator i                    (columns i NT)
            //   3 pass : issuer 0 hi*hi (even taps -> d0, odd taps -> d2), issuer 1 lo*hi -> d1a, issuer 2 hi*lo -> d1b: the tensor
            //            core's fp32 accumulation truncates, so its error grows with the number of sequential accumulations into
            //            one accumulator (measured rel-L2 ~2.2e-9 x K) — four short chains instead of one long one
            // Instruction descriptor: D = F32, A = B = TF32, K-major both, N = NT, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t d0 = tmem_base, d1a = tmem_base + (uint32_t)NT, d2 = tmem_base + 2u * (uint32_t)NT, d1b = tmem_base + 3u * (uint32_t)NT;
            const uint32_t dme = tmem_base + (uint32_t)(me * NT);              // 1 pass: this issuer's accumulator
            // A operand: K quads are qpos*16 B apart; 8-row groups = the next image row of the halo tile, or (flat) the next
            // eight flat positions
            const int rs = g_flat ? a.BX : T3_HW;
            const uint32_t a_lbo = (uint32_t)qpos * 16, a_sbo = g_flat ? 128u : (uint32_t)T3_HW * 16;
            uint32_t started = 0u, started2 = 0u;   // GEN: the first MMA issued into an accumulator overwrites it
            for (int kc = 0; kc < kchunks; ++kc) {
                const int s = kc % nstage;
                const int mask = g_up ? a.tapmask[par] : (g_kpp ? a.tapmask[kc / g_kpp] : 0x1FF);
                mbar_wait_bounded(kTransform ? &ready_bar[s] : &full_bar[s], (kc / nstage) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa_hi = smem_u32(base + s * stage_bytes), sa_lo = sa_hi + A_BYTES;
                const uint32_t sb_hi = sa_hi + F * A_BYTES, sb_lo = sb_hi + b_bytes;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap % 3;
                    if (GEN && !((mask >> tap) & 1)) continue;
                    if (NPASS == 1 && tap % T3_NISS != me) continue;
#pragma unroll
                    for (int j = 0; j < CK / 8; ++j) {
                        const uint32_t aoff = (uint32_t)((2 * j * qpos + ky * rs + kx) * 16);
                        const uint32_t boff = (uint32_t)((tap * KQ + 2 * j) * NT * 16);
                        if (NPASS == 1) {
                            // plain conv: positional (compile-time) flag — the first tap of issuer i is tap i; GEN: running flag
                            const uint32_t first = GEN ? started : ((kc | j) == 0 && tap == me ? 0u : 1u);
                            started = 1u;
                            umma_tf32(dme, umma_desc_nosw(sa_hi + aoff, a_lbo, a_sbo), umma_desc_nosw(sb_hi + boff, (uint32_t)NT * 16, 128), idesc, first);
                        } else if (me == 0) {
                            const uint64_t ah = umma_desc_nosw(sa_hi + aoff, a_lbo, a_sbo);
                            const uint64_t bh = umma_desc_nosw(sb_hi + boff, (uint32_t)NT * 16, 128);
                            if (tap & 1) { umma_tf32(d2, ah, bh, idesc, GEN ? started2 : ((kc | j) == 0 && tap == 1 ? 0u : 1u)); started2 = 1u; }
                            else { umma_tf32(d0, ah, bh, idesc, GEN ? started : ((kc | tap | j) == 0 ? 0u : 1u)); started = 1u; }
                        } else {
                            const uint32_t first = GEN ? started : ((kc | tap | j) == 0 ? 0u : 1u);
                            started = 1u;
                            if (me == 1)
                                umma_tf32(d1a, umma_desc_nosw(sa_lo + aoff, a_lbo, a_sbo), umma_desc_nosw(sb_hi + boff, (uint32_t)NT * 16, 128), idesc, first);
                            else
                                umma_tf32(d1b, umma_desc_nosw(sa_hi + aoff, a_lbo, a_sbo), umma_desc_nosw(sb_lo + boff, (uint32_t)NT * 16, 128), idesc, first);
                        }
                    }
                }
                umma_commit(&empty_bar[s]);          // implies tcgen05.fence::before_thread_sync
            }
            umma_commit(&acc_bar);
            if (me == 0) pdl_trigger();              // all loads of this CTA are consumed: release the dependent kernel
        }
    } else {
        const int et = tid - 64;                     // 0 .. 127
        pdl_wait();                                  // aux / stat buffers / direct input loads belong to the preceding kernels
        if (kTransform) {
            // ===== operand transform: split the staged tile into hi / lo (3xTF32), or stage it from global memory
            for (int kc = 0; kc < kchunks; ++kc) {
                const int s = kc % nstage;
                float4* Ah = reinterpret_cast<float4*>(base + s * stage_bytes);
                float4* Al = Ah + A_BYTES / 16;
                if (LOADER == 1) {
                    mbar_wait_bounded(&empty_bar[s], ((kc / nstage) & 1) ^ 1);
                    const float* in_n = a.in + (size_t)n * a.H * a.W * a.ldin + a.in_coff + kc * CK;
                    for (int i = et; i < KQ * T3_NPIX; i += 128) {
                        const int cq = i / T3_NPIX, p = i - cq * T3_NPIX;
                        const int iy = y0 - 1 + p / T3_HW, ix = x0 - 1 + p % T3_HW;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) v = ld4(in_n + ((size_t)iy * a.W + ix) * a.ldin + 4 * cq);
                        if (NPASS == 3) {
                            const float4 h = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
                            Ah[i] = h;
                            Al[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                        } else {
                            Ah[i] = v;
                        }
                    }
                    mbar_wait_bounded(&full_bar[s], (kc / nstage) & 1);      // the weights of this chunk have landed
                } else {
                    mbar_wait_bounded(&full_bar[s], (kc / nstage) & 1);
                    for (int i = et; i < KQ * T3_NPIX; i += 128) {
                        const float4 v = Ah[i];
                        const float4 h = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
                        Ah[i] = h;
                        Al[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
                mbar_arrive(&ready_bar[s]);
            }
        }
        // ===== epilogue: thread = output pixel m = 32*q + lane of the tile (TMEM lane m), q = warp % 4
        const int q = warp & 3;
        const int m = 32 * q + lane;
        int oy = y0 + (m >> 3), ox = x0 + (m & 7), on = n;
        bool inside = oy < a.H && ox < a.W;
        if (g_flat) {
            const int img = m / ppi, r = m - img * ppi;
            oy = r / a.BX; ox = r - oy * a.BX; on = n + img;
            inside = img < a.nimg && on < a.N && oy < a.H && ox < a.W;
        }
        size_t opix = inside ? ((size_t)on * a.H + oy) * a.W + ox : 0;
        if (g_up && inside) opix = ((size_t)on * 2 * a.H + 2 * oy + (par >> 1)) * (2 * a.W) + 2 * ox + (par & 1);
        float* orow = a.out + opix * a.ldout + a.out_coff;
        const float* arow = a.aux ? a.aux + opix * a.ldaux : nullptr;
        const int n0 = nc * NT;
        const int nvalid = min(NT, a.Cout - n0);
        const int NTP = NT + 1;
        float* sT = reinterpret_cast<float*>(base);          // [128][NT + 1] activations of the tile (stats only)
        const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16);
        // an accumulator exists only if its issuer had a tap (transposed conv, k = 2: the centre tap only)
        const int allmask = g_up ? a.tapmask[par] : (g_kpp ? (a.tapmask[0] | a.tapmask[1] | a.tapmask[2] | a.tapmask[3]) : 0x1FF);
        const bool use_d2 = (allmask & 0xAA) != 0;                                   // 3 pass: odd taps
        const bool live0 = (allmask & 0x049) != 0, live1 = (allmask & 0x092) != 0, live2 = (allmask & 0x124) != 0;   // 1 pass: taps % 3
        // The epilogue's global operands (bias; the forward activation of the act' factor) are older than the preceding kernel:
        // the first group of 8 channels is fetched BEFORE the wait for the accumulators, every further group one iteration ahead
        // (a dependent L2 round trip after each tcgen05.ld cost 0.7 us per 8 output channels: 8 us on a 96-channel layer).
        const bool epi_bias = a.epi == T3_EPI_BIAS_ACT || a.epi == T3_EPI_BIAS;
        const bool epi_grad = a.epi == T3_EPI_ACTGRAD && inside;
        float bs_n[8];
        float4 ya_n = make_float4(0.f, 0.f, 0.f, 0.f), yb_n = ya_n;
        auto fetch = [&](int c0) {
            const int ca = n0 + c0;
            const bool second = c0 + 4 < nvalid;
            if (epi_bias) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bs_n[j] = (j < 4 || second) ? __ldg(a.bias + ca + j) : 0.f;
            } else if (epi_grad) {
                ya_n = ld4(arow + ca);
                if (second) yb_n = ld4(arow + ca + 4);
            }
        };
        // (latency regime only, a.early: at 256x256 / batch 128 the launches are throughput-bound and fetching ahead measured 3-8 %
        // slower — C5 tf32 inference 32.0k -> 29.4k samples/s —, there the operands are loaded after the tcgen05.ld as before)
        if constexpr (PRE) { if (nvalid > 0) fetch(0); }
        mbar_wait_bounded(&acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < nvalid; c0 += 8) {
            float bs[8];
            float4 ya = make_float4(0.f, 0.f, 0.f, 0.f), yb = ya;
            if constexpr (PRE) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bs[j] = bs_n[j];
                ya = ya_n; yb = yb_n;
                if (c0 + 8 < nvalid) fetch(c0 + 8);
            }
            uint32_t r[8], r1[8], r2[8], r3[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(trow + (uint32_t)c0));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]) : "r"(trow + (uint32_t)(NT + c0)));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]) : "r"(trow + (uint32_t)(2 * NT + c0)));
            if (NPASS == 3) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(r3[0]), "=r"(r3[1]), "=r"(r3[2]), "=r"(r3[3]), "=r"(r3[4]), "=r"(r3[5]), "=r"(r3[6]), "=r"(r3[7]) : "r"(trow + (uint32_t)(3 * NT + c0)));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (NPASS == 3)      // (d0 + d2) + (d1a + d1b): hi*hi of the even and odd taps, then the two cross terms
                    v[j] = (__uint_as_float(r[j]) + (use_d2 ? __uint_as_float(r2[j]) : 0.f)) + (__uint_as_float(r1[j]) + __uint_as_float(r3[j]));
                else                 // the three issuers' tap groups in a fixed order
                    v[j] = ((live0 ? __uint_as_float(r[j]) : 0.f) + (live1 ? __uint_as_float(r1[j]) : 0.f)) + (live2 ? __uint_as_float(r2[j]) : 0.f);
            }
            const int ca = n0 + c0;
            const bool second = c0 + 4 < nvalid;             // Cout % 4 == 0: a group of 8 holds 4 or 8 valid channels
            if (epi_bias) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < 4 || second) {
                        if constexpr (PRE) v[j] += bs[j];
                        else v[j] += __ldg(a.bias + ca + j);
                        if (a.epi == T3_EPI_BIAS_ACT) v[j] = act_f(v[j], a.act);
                    }
            } else if (epi_grad) {
                if constexpr (!PRE) ya = ld4(arow + ca);
                v[0] *= act_grad_from_out(ya.x, a.act); v[1] *= act_grad_from_out(ya.y, a.act);
                v[2] *= act_grad_from_out(ya.z, a.act); v[3] *= act_grad_from_out(ya.w, a.act);
                if (second) {
                    if constexpr (!PRE) yb = ld4(arow + ca + 4);
                    v[4] *= act_grad_from_out(yb.x, a.act); v[5] *= act_grad_from_out(yb.y, a.act);
                    v[6] *= act_grad_from_out(yb.z, a.act); v[7] *= act_grad_from_out(yb.w, a.act);
                }
            }
            if (inside) {
                st4(orow + ca, make_float4(v[0], v[1], v[2], v[3]));
                if (second) st4(orow + ca + 4, make_float4(v[4], v[5], v[6], v[7]));
            }
            if (a.stat_part) {
#pragma unroll
                for (int j = 0; j < 8; ++j) sT[m * NTP + c0 + j] = inside ? v[j] : 0.f;
            }
        }
        if (a.stat_part) {
            // per-tile (sum, sumsq) per channel in a fixed order: G row groups, then the groups in order
            asm volatile("bar.sync 1, 128;" ::: "memory");
            float* sP = sT + 128 * NTP;                       // [G][2][NT]
            const int G = NT <= 128 ? 128 / NT : 1;
            const int c = et % NT, g = et / NT;
            if (g < G && c < nvalid) {
                const int R = 128 / G;
                float s = 0.f, sq = 0.f;
                for (int rr = g * R; rr < (g + 1) * R; ++rr) { const float t = sT[rr * NTP + c]; s += t; sq = fmaf(t, t, sq); }
                sP[(g * 2 + 0) * NT + c] = s;
                sP[(g * 2 + 1) * NT + c] = sq;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int slot = g_flat ? tile : n * (a.tiles_x * a.tiles_y) + tile;
            for (int i = et; i < 2 * nvalid; i += 128) {
                const int which = i / nvalid, cc = i - which * nvalid;
                float s = 0.f;
                for (int gg = 0; gg < G; ++gg) s += sP[(gg * 2 + which) * NT + cc];
                a.stat_part[((size_t)slot * 2 + which) * a.Cout + n0 + cc] = s;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
}

// ------------------------------------------------------------------ kernel v2: persistent + pipelined
// One CTA per SM slot loops over output tiles (static round robin over the flattened (n, tile, n-chunk) space):
//   * the producer runs ahead through an nstage ring of (activation chunk [+ weight chunk]) stages, so the TMA loads of the
//     next tiles are in flight while the current one is multiplied; single-chunk layers (Cin <= CK) load their weight
//     block once per stage instead of once per tile;
//   * two TMEM accumulator buffers: the MMA warp starts tile i+1 as soon as its operands are ready while the epilogue warps
//     drain tile i (tcgen05.ld -> bias / activation -> stores), then hand the buffer back through acc_empty;
//   * barriers, TMEM allocation and the descriptor prefetch are paid once per CTA, not once per tile;
//   * BatchNorm partials by warp shuffles (fixed order) — no staging buffer aliases the in-flight stages.
// Same arithmetic, argument block and weight layout as v1 (bit-identical outputs).
struct Tc3Sched { int total, tiles_per_img, nchunks_n; };

template <int CK, int NPASS, int LOADER>
__global__ void __launch_bounds__(T3_THREADS) tc3conv2_kernel(const __grid_constant__ CUtensorMap map_a, const Tc3Args a, const Tc3Sched sc) {
    extern __shared__ __align__(128) uint8_t t3_smem[];
    constexpr int KQ = CK / 4;
    constexpr int A_BYTES = KQ * T3_NPIX * 16;
    constexpr int F = NPASS == 3 ? 2 : 1;
    constexpr int NACC = NPASS == 3 ? 3 : 1;
    const int NT = a.NT;
    const int b_bytes = 36 * CK * NT;
    const int stage_bytes = F * (A_BYTES + b_bytes);
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(t3_smem) + 127) & ~(uintptr_t)127);
    __shared__ uint64_t full_bar[T3_MAXSTAGE], ready_bar[T3_MAXSTAGE], empty_bar[T3_MAXSTAGE], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_stat[4][2][128];                       // per-warp (sum, sumsq) of the current tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kchunks = a.kchunks, nstage = a.nstage;
    constexpr bool kTransform = (NPASS == 3) || (LOADER == 1);
    const int my_tiles = (sc.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles t = blockIdx.x + i * gridDim.x
    auto decode = [&](int i, int& n, int& y0, int& x0, int& nc) {
        const int t = (int)blockIdx.x + i * (int)gridDim.x;
        nc = t % sc.nchunks_n;
        const int r = t / sc.nchunks_n;
        const int tile = r % sc.tiles_per_img;
        n = r / sc.tiles_per_img;
        y0 = (tile / a.tiles_x) * T3_TH; x0 = (tile % a.tiles_x) * T3_TW;
    };

    if (tid == 0) {
        for (int s = 0; s < T3_MAXSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], 128); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int nit = my_tiles * kchunks;                       // flat (tile, chunk) iterations of this CTA
    const bool b_once = kchunks == 1 && sc.nchunks_n == 1;    // the weight block is the same for every iteration

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer
            if (LOADER == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            int pre = 0;                                       // weight blocks issued before the dependency wait
            if (a.w_early && nit > 0) {
                const int lim = b_once ? min(nstage, nit) : 1;
                for (; pre < lim; ++pre) {
                    int n, y0, x0, nc;
                    decode(pre / kchunks, n, y0, x0, nc);
                    mbar_expect_tx(&full_bar[pre], (LOADER == 0 ? A_BYTES : 0) + F * b_bytes);
                    bulk_g2s(base + pre * stage_bytes + F * A_BYTES, a.wq + (size_t)(nc * kchunks + pre % kchunks) * F * 9 * CK * NT, F * b_bytes, &full_bar[pre]);
                }
            }
            pdl_wait();
            for (int it = 0; it < nit; ++it) {
                const int s = it % nstage, kc = it % kchunks;
                int n, y0, x0, nc;
                decode(it / kchunks, n, y0, x0, nc);
                mbar_wait_bounded(&empty_bar[s], ((it / nstage) & 1) ^ 1);
                uint8_t* sa = base + s * stage_bytes;
                const bool need_b = !(b_once && it >= nstage);            // single-block layers keep it in every stage
                if (it >= pre) mbar_expect_tx(&full_bar[s], (LOADER == 0 ? A_BYTES : 0) + (need_b ? F * b_bytes : 0));
                if (LOADER == 0) tma_load_5d(sa, &map_a, &full_bar[s], 0, x0 - 1, y0 - 1, n, kc * KQ);
                if (need_b && it >= pre) bulk_g2s(sa + F * A_BYTES, a.wq + (size_t)(nc * kchunks + kc) * F * 9 * CK * NT, F * b_bytes, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                mbar_wait_bounded(&acc_empty[buf], ((i >> 1) & 1) ^ 1);        // the epilogue has drained this accumulator buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d0 = tmem_base + (uint32_t)(buf * NACC * NT), d1 = d0 + (uint32_t)NT, d2 = d0 + 2u * (uint32_t)NT;
                for (int kc = 0; kc < kchunks; ++kc) {
                    const int it = i * kchunks + kc, s = it % nstage;
                    mbar_wait_bounded(kTransform ? &ready_bar[s] : &full_bar[s], (it / nstage) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa_hi = smem_u32(base + s * stage_bytes), sa_lo = sa_hi + A_BYTES;
                    const uint32_t sb_hi = sa_hi + F * A_BYTES, sb_lo = sb_hi + b_bytes;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
#pragma unroll
                        for (int j = 0; j < CK / 8; ++j) {
                            const uint32_t aoff = (uint32_t)(((2 * j * T3_HH + ky) * T3_HW + kx) * 16);
                            const uint32_t boff = (uint32_t)((tap * KQ + 2 * j) * NT * 16);
                            const uint32_t first = (kc | tap | j) == 0 ? 0u : 1u;
                            const uint64_t ah = umma_desc_nosw(sa_hi + aoff, T3_NPIX * 16, T3_HW * 16);
                            const uint64_t bh = umma_desc_nosw(sb_hi + boff, (uint32_t)NT * 16, 128);
                            if (NPASS == 3 && (tap & 1)) umma_tf32(d2, ah, bh, idesc, (kc | j) == 0 && tap == 1 ? 0u : 1u);
                            else umma_tf32(d0, ah, bh, idesc, first);
                            if (NPASS == 3) {
                                const uint64_t al = umma_desc_nosw(sa_lo + aoff, T3_NPIX * 16, T3_HW * 16);
                                const uint64_t bl = umma_desc_nosw(sb_lo + boff, (uint32_t)NT * 16, 128);
                                umma_tf32(d1, al, bh, idesc, first);
                                umma_tf32(d1, ah, bl, idesc, 1u);
                            }
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&acc_full[buf]);
            }
            pdl_trigger();
        }
    } else {
        const int et = tid - 64;
        const int q = warp & 3;
        const int m = 32 * q + lane;
        pdl_wait();
        auto transform = [&](int i) {        // split / stage the operand chunks of tile i
            for (int kc = 0; kc < kchunks; ++kc) {
                const int it = i * kchunks + kc, s = it % nstage;
                float4* Ah = reinterpret_cast<float4*>(base + s * stage_bytes);
                float4* Al = Ah + A_BYTES / 16;
                if (LOADER == 1) {
                    int n, y0, x0, nc;
                    decode(i, n, y0, x0, nc);
                    mbar_wait_bounded(&empty_bar[s], ((it / nstage) & 1) ^ 1);
                    const float* in_n = a.in + (size_t)n * a.H * a.W * a.ldin + a.in_coff + kc * CK;
                    for (int k = et; k < KQ * T3_NPIX; k += 128) {
                        const int cq = k / T3_NPIX, p = k - cq * T3_NPIX;
                        const int iy = y0 - 1 + p / T3_HW, ix = x0 - 1 + p % T3_HW;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) v = ld4(in_n + ((size_t)iy * a.W + ix) * a.ldin + 4 * cq);
                        if (NPASS == 3) {
                            const float4 h = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
                            Ah[k] = h;
                            Al[k] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                        } else {
                            Ah[k] = v;
                        }
                    }
                    mbar_wait_bounded(&full_bar[s], (it / nstage) & 1);
                } else {
                    mbar_wait_bounded(&full_bar[s], (it / nstage) & 1);
                    for (int k = et; k < KQ * T3_NPIX; k += 128) {
                        const float4 v = Ah[k];
                        const float4 h = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
                        Ah[k] = h;
                        Al[k] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&ready_bar[s]);
            }
        };
        if (kTransform && my_tiles > 0) transform(0);
        for (int i = 0; i < my_tiles; ++i) {
            if (kTransform && i + 1 < my_tiles) transform(i + 1);            // tile i+1 can be multiplied while tile i is drained
            const int buf = i & 1;
            int n, y0, x0, nc;
            decode(i, n, y0, x0, nc);
            mbar_wait_bounded(&acc_full[buf], (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int oy = y0 + (m >> 3), ox = x0 + (m & 7);
            const bool inside = oy < a.H && ox < a.W;
            const size_t opix = ((size_t)n * a.H + (inside ? oy : 0)) * a.W + (inside ? ox : 0);
            float* orow = a.out + opix * a.ldout + a.out_coff;
            const float* arow = a.aux ? a.aux + opix * a.ldaux : nullptr;
            const int n0 = nc * NT;
            const int nvalid = min(NT, a.Cout - n0);
            const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * NACC * NT);
            for (int c0 = 0; c0 < nvalid; c0 += 8) {
                uint32_t r[8], r1[8], r2[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(trow + (uint32_t)c0));
                if (NPASS == 3) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]) : "r"(trow + (uint32_t)(NT + c0)));
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]) : "r"(trow + (uint32_t)(2 * NT + c0)));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + 8 >= nvalid) {       // last read of this accumulator buffer: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&acc_empty[buf]);
                }
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    v[j] = NPASS == 3 ? (__uint_as_float(r[j]) + __uint_as_float(r2[j])) + __uint_as_float(r1[j]) : __uint_as_float(r[j]);
                const int ca = n0 + c0;
                const bool second = c0 + 4 < nvalid;
                if (a.epi == T3_EPI_BIAS_ACT || a.epi == T3_EPI_BIAS) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < 4 || second) {
                            v[j] += __ldg(a.bias + ca + j);
                            if (a.epi == T3_EPI_BIAS_ACT) v[j] = act_f(v[j], a.act);
                        }
                } else if (a.epi == T3_EPI_ACTGRAD && inside) {
                    const float4 ya = ld4(arow + ca);
                    v[0] *= act_grad_from_out(ya.x, a.act); v[1] *= act_grad_from_out(ya.y, a.act);
                    v[2] *= act_grad_from_out(ya.z, a.act); v[3] *= act_grad_from_out(ya.w, a.act);
                    if (second) {
                        const float4 yb = ld4(arow + ca + 4);
                        v[4] *= act_grad_from_out(yb.x, a.act); v[5] *= act_grad_from_out(yb.y, a.act);
                        v[6] *= act_grad_from_out(yb.z, a.act); v[7] *= act_grad_from_out(yb.w, a.act);
                    }
                }
                if (inside) {
                    st4(orow + ca, make_float4(v[0], v[1], v[2], v[3]));
                    if (second) st4(orow + ca + 4, make_float4(v[4], v[5], v[6], v[7]));
                }
                if (a.stat_part) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float t = inside ? v[j] : 0.f;
                        const float su = warp_sum(t), sq = warp_sum(t * t);
                        if (lane == 0) { s_stat[q][0][c0 + j] = su; s_stat[q][1][c0 + j] = sq; }
                    }
                }
            }
            if (a.stat_part) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int slot = n * sc.tiles_per_img + ((y0 / T3_TH) * a.tiles_x + x0 / T3_TW);
                for (int k = et; k < 2 * nvalid; k += 128) {
                    const int which = k / nvalid, cc = k - which * nvalid;
                    const float sv = ((s_stat[0][which][cc] + s_stat[1][which][cc]) + s_stat[2][which][cc]) + s_stat[3][which][cc];
                    a.stat_part[((size_t)slot * 2 + which) * a.Cout + n0 + cc] = sv;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");      // s_stat is rewritten by the next tile
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
}

// ------------------------------------------------------------------ weight preparation
// dst block (nc, kc): [hl][tap][kq][n][e] with k = kc*CK + 4*kq + e the contracted channel, ng = nc*NT + n the output one.
//   forward (flip = 0): src = W[tap][k][ng]           (Keras kernel (3,3,Cin,Cout): contraction over Cin)
//   dgrad   (flip = 1): src = W[8 - tap][ng][k]       (contraction over Cout, output channel = Cin index)
// flip: 0 Conv2D forward, 1 Conv2D dgrad, 2 Conv2DTranspose forward (4 parity sets), 3 Conv2DTranspose dgrad (kchunks = 4 planes
// x Cout chunks); ldw_k = the transposed conv's kernel size (2 | 3 | 5)
struct Tc3WPrep { int64_t w_off, dst_off; int Kc, Nc, ldw_k, NT, nchunks_n, CK, kchunks, flip, npass; };

__global__ void tc3_wprep_kernel(const Tc3WPrep* __restrict__ tab, const float* __restrict__ params, float* __restrict__ dst) {
    // one thread = one (block, tap, channel quad, output channel): four contracted channels -> one float4 store.  Consecutive
    // threads take consecutive output channels, so the stores are contiguous and the loads are either contiguous rows (forward,
    // transposed-conv dgrad: W[..][k][n]) or 16-byte pieces one weight row apart (dgrad, transposed-conv forward: W[..][n][k]) —
    // the element-wise version gathered single floats a weight row apart and ran at 350-500 GB/s on the step's critical path
    const Tc3WPrep e = tab[blockIdx.y];
    const int KQ = e.CK / 4;
    const int per_block = 9 * e.CK * e.NT;
    const int nblk = (e.flip == 2 ? 4 : 1) * e.nchunks_n * e.kchunks;
    const int total4 = nblk * 9 * KQ * e.NT;
    const int F = e.npass == 3 ? 2 : 1;
    const int ksz = e.ldw_k, pb = (ksz - 2) / 2;               // transposed conv: kernel size, 'same' crop offset
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const int nn = i % e.NT;
        const int kq = (i / e.NT) % KQ;
        const int tap = (i / (e.NT * KQ)) % 9;
        const int blk = i / (e.NT * KQ * 9);
        const int kc = blk % e.kchunks, nc = (blk / e.kchunks) % e.nchunks_n;
        const int ng = nc * e.NT + nn;
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        // src index of (k, ng) = base + k * sk + ng * sn; valid = the tap exists and ng < Nc
        int64_t base = -1;
        int sk = 0, k0 = 0;
        if (e.flip == 0) {              // Conv2D forward, W (3,3,Cin = Kc,Cout = Nc): [tap][k][ng]
            k0 = kc * e.CK + 4 * kq; base = e.w_off + (int64_t)tap * e.Kc * e.Nc + ng; sk = e.Nc;
        } else if (e.flip == 1) {       // Conv2D dgrad (contraction over Cout = Kc): W[8 - tap][ng][k]
            k0 = kc * e.CK + 4 * kq; base = e.w_off + ((int64_t)(8 - tap) * e.Nc + ng) * e.Kc; sk = 1;
        } else if (e.flip == 2) {
            // Conv2DTranspose forward, W (k, k, Cout = Nc, Cin = Kc): parity (a, b), window offset d = tap - 1 reads x[i + d]
            // with the tap ky = a + pb - 2 d   (out[2i + a] = sum x[i'] W[2 (i - i') + a + pb])
            const int par = blk / (e.kchunks * e.nchunks_n);
            const int ky = (par >> 1) + pb - 2 * (tap / 3 - 1), kx = (par & 1) + pb - 2 * (tap % 3 - 1);
            k0 = kc * e.CK + 4 * kq; sk = 1;
            if (ky >= 0 && ky < ksz && kx >= 0 && kx < ksz) base = e.w_off + ((int64_t)(ky * ksz + kx) * e.Nc + ng) * e.Kc;
        } else {
            // Conv2DTranspose input gradient, W (k, k, Cout = Kc, Cin = Nc): k chunk = (parity plane, Cout chunk); plane (a, b)
            // at window offset e = tap - 1 carries the tap ky = 2 e + a + pb   (dx[i] = sum dy[2 (i + e) + a] W[2 e + a + pb])
            const int kpp = e.kchunks / 4, par = kc / kpp;
            const int ky = 2 * (tap / 3 - 1) + (par >> 1) + pb, kx = 2 * (tap % 3 - 1) + (par & 1) + pb;
            k0 = (kc - par * kpp) * e.CK + 4 * kq; sk = e.Nc;
            if (ky >= 0 && ky < ksz && kx >= 0 && kx < ksz) base = e.w_off + (int64_t)(ky * ksz + kx) * e.Kc * e.Nc + ng;
        }
        if (base >= 0 && ng < e.Nc) {
#pragma unroll
            for (int el = 0; el < 4; ++el)
                if (k0 + el < e.Kc) w[el] = __ldg(params + base + (int64_t)(k0 + el) * sk);
        }
        float4 hi = make_float4(rna_tf32(w[0]), rna_tf32(w[1]), rna_tf32(w[2]), rna_tf32(w[3]));
        float* d = dst + e.dst_off + (size_t)blk * F * per_block + ((size_t)(tap * KQ + kq) * e.NT + nn) * 4;
        *reinterpret_cast<float4*>(d) = hi;
        if (F == 2) *reinterpret_cast<float4*>(d + per_block) = make_float4(w[0] - hi.x, w[1] - hi.y, w[2] - hi.z, w[3] - hi.w);
    }
}

// taps of the 3x3 window that exist for output parity / parity plane p of a transposed conv with kernel ksz (see above)
static inline int tc3_convt_tapmask(int ksz, int par, bool dgrad) {
    const int pb = (ksz - 2) / 2;
    int mask = 0;
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int ky = dgrad ? 2 * dy + (par >> 1) + pb : (par >> 1) + pb - 2 * dy;
        const int kx = dgrad ? 2 * dx + (par & 1) + pb : (par & 1) + pb - 2 * dx;
        if (ky >= 0 && ky < ksz && kx >= 0 && kx < ksz) mask |= 1 << tap;
    }
    return mask;
}

// ------------------------------------------------------------------ host: plan, tensor map, launch
struct Tc3Plan { bool ok; int CK, NT, nchunks_n, kchunks, nstage, tmem_cols; size_t smem, wq_floats; int ns_max; size_t stage_bytes;
                 int nstage2, tmem_cols2, ctas_per_sm2; size_t smem2; };      // *2: the persistent kernel (tc3conv2_kernel)

static inline Tc3Plan tc3_plan(int Cin, int Cout, int npass, int nt_cap = 0) {
    Tc3Plan p;
    memset(&p, 0, sizeof p);
    // contracted channels: any multiple of 4; a count that is no multiple of 8 (12, 20, ...) is padded to the next one — the
    // missing channel quad of the last chunk is zero-filled by the TMA unit and has zero weights (tc3_wprep_kernel)
    if (Cin % 4 != 0 || Cout % 4 != 0 || Cin < 8 || Cout < 4) return p;
    const int Cp = (Cin + 7) / 8 * 8;
    const int F = npass == 3 ? 2 : 1;
    const int npad = (Cout + 15) / 16 * 16;
    int ntmax = npass == 3 ? 64 : 128;
    if (nt_cap >= 16 && nt_cap < ntmax) ntmax = nt_cap;      // weight-streaming layers: more, narrower output-channel chunks
    p.nchunks_n = cdiv(npad, ntmax);
    p.NT = (cdiv(npad, p.nchunks_n) + 15) / 16 * 16;
    const int cks[3] = {32, 16, 8};
    for (int i = 0; i < 3; ++i) {
        const int ck = cks[i];
        if (Cp % ck) continue;
        const size_t stage = (size_t)F * ((size_t)ck * 720 + (size_t)36 * ck * p.NT);
        if (stage > 72 * 1024 && ck > 8) continue;
        p.CK = ck;
        p.kchunks = Cp / ck;
        int ns = (int)std::min<size_t>((size_t)T3_MAXSTAGE - 1, (180 * 1024) / stage);
        if (ns < 1) return p;
        p.nstage = std::min(p.kchunks, ns);
        p.ns_max = ns; p.stage_bytes = stage;
        const size_t stats = (size_t)(128 * (p.NT + 1) + 2 * 128) * 4;
        p.smem = std::max((size_t)p.nstage * stage, stats) + 128;
        break;
    }
    if (!p.CK) return p;
    const int cols = (npass == 3 ? 4 : 3) * p.NT;              // v1 kernel: one accumulator per MMA issuer (3), four for the 3-pass split
    p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    p.wq_floats = (size_t)p.nchunks_n * p.kchunks * F * 9 * p.CK * p.NT;
    {   // persistent kernel: a deeper ring (tiles in flight), two accumulator buffers
        const size_t stage = (size_t)F * ((size_t)p.CK * 720 + (size_t)36 * p.CK * p.NT);
        int ns = (int)std::min<size_t>(T3_MAXSTAGE, (96 * 1024) / stage);
        if (ns < 2) ns = (int)std::min<size_t>(T3_MAXSTAGE, (200 * 1024) / stage);
        if (ns < 2) ns = 1;
        p.nstage2 = ns;
        p.smem2 = (size_t)ns * stage + 128;
        const int cols2 = 2 * (npass == 3 ? 3 : 1) * p.NT;
        p.tmem_cols2 = cols2 <= 32 ? 32 : cols2 <= 64 ? 64 : cols2 <= 128 ? 128 : cols2 <= 256 ? 256 : 512;
        int r = (int)std::min<size_t>(4, (200 * 1024) / (p.smem2 + 6 * 1024));
        r = std::min(r, 512 / p.tmem_cols2);
        p.ctas_per_sm2 = std::max(r, 1);
    }
    p.ok = true;
    return p;
}

// transposed-conv variants of a plan: forward = 4 parity weight sets; dgrad = 4 parity planes in the contraction
static inline Tc3Plan tc3_plan_convt_fwd(Tc3Plan p) { p.wq_floats *= 4; return p; }
static inline Tc3Plan tc3_plan_convt_dgrad(Tc3Plan p) {
    if (!p.ok) return p;
    p.kchunks *= 4; p.wq_floats *= 4;
    p.nstage = std::min(p.kchunks, p.ns_max);
    const size_t stats = (size_t)(128 * (p.NT + 1) + 2 * 128) * 4;
    p.smem = std::max((size_t)p.nstage * p.stage_bytes, stats) + 128;
    return p;
}

static inline bool tc3_flat(int H, int W) { return H * W < 64; }
static inline int tc3_flat_nimg(int H, int W) { return 128 / ((H + 2) * (W + 2)); }
// plan of a layer on an H x W grid at batch <= Nmax: small-image (flat) layers are weight-streaming bound — a handful of
// tiles, megabytes of weights — so their output channels are cut into more, narrower chunks (~64 CTAs share the stream)
static inline Tc3Plan tc3_plan_for(int H, int W, int Nmax, int Cin, int Cout, int npass) {
    if (!tc3_flat(H, W)) return tc3_plan(Cin, Cout, npass);
    const int tiles = cdiv(Nmax, tc3_flat_nimg(H, W));
    const int cap = std::max(16, (Cout * tiles / 64 + 15) / 16 * 16);
    return tc3_plan(Cin, Cout, npass, cap);
}

// activations [Nmax, H, W, ld] fp32 (x points at the first contracted channel): 5-D view (c4, x, y, n, cq)
static inline int tc3_make_map(const float* x, int Nmax, int H, int W, int C, int ld, int CK, CUtensorMap* m) {
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[5] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nmax, (cuuint64_t)(C / 4)};
    cuuint64_t strides[4] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4, 16};
    cuuint32_t box[5] = {4, T3_HW, T3_HH, 1, (cuuint32_t)(CK / 4)}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled(tc3 A: C=%d ld=%d H=%d W=%d) failed: %d", C, ld, H, W, (int)r);
    return 0;
}

// flat geometry of small images: whole zero-padded images, as many as fit the 128 rows of a tile
static inline int tc3_make_map_flat(const float* x, int Nmax, int H, int W, int C, int ld, int CK, CUtensorMap* m) {
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[5] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nmax, (cuuint64_t)(C / 4)};
    cuuint64_t strides[4] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4, 16};
    cuuint32_t box[5] = {4, (cuuint32_t)(W + 2), (cuuint32_t)(H + 2), (cuuint32_t)tc3_flat_nimg(H, W), (cuuint32_t)(CK / 4)}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled(tc3 flat A: C=%d ld=%d H=%d W=%d) failed: %d", C, ld, H, W, (int)r);
    return 0;
}
// 5-D view with explicit pixel strides (in floats): the stride-2 parity planes of a transposed conv's output gradient
static inline int tc3_make_map_strided(const float* x, int Nmax, int H, int W, int C, int64_t sx, int64_t sy, int64_t sn, int CK, CUtensorMap* m) {
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const bool flat = tc3_flat(H, W);
    cuuint64_t dims[5] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nmax, (cuuint64_t)(C / 4)};
    cuuint64_t strides[4] = {(cuuint64_t)sx * 4, (cuuint64_t)sy * 4, (cuuint64_t)sn * 4, 16};
    cuuint32_t box[5] = {4, (cuuint32_t)(flat ? W + 2 : T3_HW), (cuuint32_t)(flat ? H + 2 : T3_HH), (cuuint32_t)(flat ? tc3_flat_nimg(H, W) : 1),
                         (cuuint32_t)(CK / 4)}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled(tc3 strided A: C=%d H=%d W=%d) failed: %d", C, H, W, (int)r);
    return 0;
}
// the four parity planes (a, b) of dy [Nmax, 2H, 2W, ld] (dy points at the first channel of the slice)
static inline int tc3_make_maps_parity(const float* dy, int Nmax, int H, int W, int C, int ld, int CK, Tc3Maps* ms) {
    for (int par = 0; par < 4; ++par) {
        const float* base = dy + ((size_t)(par >> 1) * 2 * W + (par & 1)) * ld;
        S2S_CHECK(tc3_make_map_strided(base, Nmax, H, W, C, 2 * (int64_t)ld, 4 * (int64_t)W * ld, 4 * (int64_t)H * W * ld, CK, &ms->m[par]));
    }
    return 0;
}
// one map builder for both geometries
static inline int tc3_make_map_any(const float* x, int Nmax, int H, int W, int C, int ld, int CK, CUtensorMap* m) {
    return tc3_flat(H, W) ? tc3_make_map_flat(x, Nmax, H, W, C, ld, CK, m) : tc3_make_map(x, Nmax, H, W, C, ld, CK, m);
}

static inline int tc3_stat_slots(int H, int W, int N) {
    return tc3_flat(H, W) ? cdiv(N, tc3_flat_nimg(H, W)) : N * cdiv(H, T3_TH) * cdiv(W, T3_TW);
}

template <int CK, int NPASS, int LOADER, int GEN>
static int tc3_launch_inst_g(const Tc3Maps& map, const Tc3Args& a, const Tc3Plan& p, cudaStream_t st) {
    static DevOnce once, once_pre;
    const int gy = (a.up ? 4 : 1) * p.nchunks_n;
    dim3 grid(a.tiles_x * a.tiles_y, gy, a.N);
    if (a.flat) grid = dim3(cdiv(a.N, a.nimg), gy, 1);
    if (a.early) {
        S2S_CUDA(once_pre.run([] { return cudaFuncSetAttribute(tc3conv_kernel<CK, NPASS, LOADER, GEN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
        launch_k(tc3conv_kernel<CK, NPASS, LOADER, GEN, true>, grid, dim3(T3_THREADS_V1), p.smem, st, map, a);
        return 0;
    }
    S2S_CUDA(once.run([] { return cudaFuncSetAttribute(tc3conv_kernel<CK, NPASS, LOADER, GEN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
    launch_k(tc3conv_kernel<CK, NPASS, LOADER, GEN, false>, grid, dim3(T3_THREADS_V1), p.smem, st, map, a);
    return 0;
}
template <int CK, int NPASS, int LOADER>
static int tc3_launch_inst(const Tc3Maps& map, const Tc3Args& a, const Tc3Plan& p, cudaStream_t st) {
    if (a.flat || a.up || a.kpp) return tc3_launch_inst_g<CK, NPASS, LOADER, 1>(map, a, p, st);
    return tc3_launch_inst_g<CK, NPASS, LOADER, 0>(map, a, p, st);
}

template <int CK, int NPASS, int LOADER>
static int tc3_launch_inst2(const Tc3Maps& maps, const Tc3Args& a, const Tc3Plan& p, cudaStream_t st) {
    const CUtensorMap& map = maps.m[0];
    static DevOnce once;
    S2S_CUDA(once.run([] { return cudaFuncSetAttribute(tc3conv2_kernel<CK, NPASS, LOADER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    Tc3Sched sc;
    sc.tiles_per_img = a.tiles_x * a.tiles_y; sc.nchunks_n = p.nchunks_n;
    sc.total = a.N * sc.tiles_per_img * p.nchunks_n;
    const int grid = std::min(sc.total, sms * p.ctas_per_sm2);
    launch_k(tc3conv2_kernel<CK, NPASS, LOADER>, dim3(grid), dim3(T3_THREADS), p.smem2, st, map, a, sc);
    return 0;
}

static inline bool tc3_use_v2() {
    // opt-in (S2S_TC3_V2=1): measured on B200 the persistent kernel is no faster inside a train step (batch 16: 363.6 vs
    // 349.3 us per step, batch 128: 1073.6 vs 1083.6 us) — the layers are short and the per-tile ring adds latency.
    static const bool v2 = [] { const char* e = getenv("S2S_TC3_V2"); return e && e[0] == '1'; }();
    return v2;
}

static inline int tc3_launch_maps(const Tc3Maps& map, Tc3Args a, const Tc3Plan& p, int npass, int loader, const char* tag, cudaStream_t st) {
    S2S_REQUIRE(p.ok, "tc3conv: no plan for %d -> %d", a.Cin, a.Cout);
    S2S_REQUIRE((a.ldout & 3) == 0 && (a.out_coff & 3) == 0, "tc3conv: output stride must be a multiple of 4");
    a.flat = tc3_flat(a.H, a.W) ? 1 : 0;
    if (a.flat) {
        S2S_REQUIRE(loader == 0, "tc3conv: the flat (small-image) geometry needs the TMA loader");
        a.BX = a.W + 2; a.BY = a.H + 2; a.nimg = tc3_flat_nimg(a.H, a.W);
    }
    S2S_REQUIRE(loader == 0 || a.Cin % p.CK == 0, "tc3conv: the ld.global loader needs Cin %% CK == 0");
    S2S_REQUIRE(!(a.up || a.kpp) || loader == 0, "tc3conv: the transposed-conv variants are TMA-loaded");
    a.nchn = p.nchunks_n;
    const bool v2 = tc3_use_v2() && p.nstage2 >= 2 && !a.flat && !a.up && !a.kpp;
    a.NT = p.NT; a.kchunks = p.kchunks; a.nstage = v2 ? p.nstage2 : p.nstage; a.tmem_cols = v2 ? p.tmem_cols2 : p.tmem_cols;
    a.tiles_x = cdiv(a.W, T3_TW); a.tiles_y = cdiv(a.H, T3_TH);
    prof_begin(st, tag, 4.0 * a.N * a.H * a.W * ((double)a.Cin + a.Cout), 18.0 * (double)a.Cin * a.Cout * a.N * a.H * a.W);
    int rc = -1;
#define S2S_T3(CKV)                                                                                        \
    if (p.CK == CKV && v2) {                                                                               \
        if (npass == 3) rc = loader ? tc3_launch_inst2<CKV, 3, 1>(map, a, p, st) : tc3_launch_inst2<CKV, 3, 0>(map, a, p, st); \
        else rc = loader ? tc3_launch_inst2<CKV, 1, 1>(map, a, p, st) : tc3_launch_inst2<CKV, 1, 0>(map, a, p, st);            \
    } else if (p.CK == CKV) {                                                                              \
        if (npass == 3) rc = loader ? tc3_launch_inst<CKV, 3, 1>(map, a, p, st) : tc3_launch_inst<CKV, 3, 0>(map, a, p, st); \
        else rc = loader ? tc3_launch_inst<CKV, 1, 1>(map, a, p, st) : tc3_launch_inst<CKV, 1, 0>(map, a, p, st);            \
    }
    S2S_T3(8) S2S_T3(16) S2S_T3(32)
#undef S2S_T3
    prof_end(st);
    if (rc != 0) return rc < 0 ? fail(S2S_ERR_INVALID, "tc3conv: no instantiation for CK=%d", p.CK) : rc;
    S2S_LAUNCH_CHECK();
    return 0;
}

static inline int tc3_launch(const CUtensorMap& map, const Tc3Args& a, const Tc3Plan& p, int npass, int loader, const char* tag, cudaStream_t st) {
    Tc3Maps ms;
    for (int i = 0; i < 4; ++i) ms.m[i] = map;
    return tc3_launch_maps(ms, a, p, npass, loader, tag, st);
}

}  // namespace s2s
