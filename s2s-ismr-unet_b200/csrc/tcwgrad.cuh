// tcwgrad.cuh — Conv2D 3x3 'same' WEIGHT gradient (+ bias gradient) as a pixel-contraction GEMM on the 5th-generation
// tensor cores (tcgen05.mma.kind::tf32, accumulators in TMEM, operands staged by TMA).  precision = "tf32" training.
//
//   dW[tap][ci][co] = sum_{n,y,x} X[n, y+ky-1, x+kx-1, ci] * dZ[n, y, x, co]          (Keras kernel layout (3,3,Cin,Cout))
//   db[co]          = sum_{n,y,x} dZ[n, y, x, co]
//
// GEMM view per filter tap:  D_tap[co (M <= 128), ci (N)] += A[co, p] * B_tap[p, ci]  with the PIXELS p as the contracted
// dimension.  Both operands are read from the NHWC activations exactly as the forward kernel (tc3conv.cuh) stages them:
// one 5-D TMA box (c4 = 4 channels, x, y, n, channel quad) lands in shared memory as the planar layout
// [quad][position][4 floats].  Eight consecutive positions of one channel quad are 128 contiguous bytes = one core matrix
// (8 K-rows x 16 B) of the no-swizzle MN-MAJOR UMMA operand layout, quads are SBO bytes apart — so
//     A descriptor (dZ tile):      start = Z + g*128,                          SBO = positions_Z*16
//     B descriptor (X halo tile):  start = X + (g*GS + ky*RS + kx)*16,         SBO = positions_X*16
// give, for K-group g (8 pixels) and tap (ky, kx), one tcgen05.mma of M = 128 x N = NCI x K = 8: the nine taps are nine
// start addresses into the same halo tile and accumulate into nine column blocks of TMEM (9*NCI + 16 <= 512 columns).
// The bias gradient rides along as a tenth MMA per group against a tile of ones.
//
// Two tile geometries:
//   * normal (H*W >= 64): 8 (x) x 16 (y) output tile of one image, halo box 10 x 18 at (x0-1, y0-1): group g = tile row,
//     GS = RS = 10.  Out-of-image positions are zero-filled by the TMA unit ('same' padding; zero dZ rows add nothing).
//   * flat (small images, H*W < 64): NIMG whole zero-padded images per tile, box (BX >= W+2, BY = H+2, NIMG) for X at
//     (-1,-1) and for dZ at (0,0): position p = (img, py, px) flat, group g = positions 8g..8g+7, GS = 8, RS = BX.  A dZ
//     position outside the image is zero (TMA fill), so the junk pairings of the flat shift contribute nothing.
// Deterministic: CTA (slot, ci chunk, co chunk) walks tiles slot, slot+nslots, ... in order and writes one partial
// [slot][tap][ci][co]; the partials are summed in slot order by the fused reduce + Adam kernel (optim.cuh), like wgrad.cuh.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 =
// epilogue (TMEM lane = output channel co: tcgen05.ld 16 columns = 16 ci of one tap, stores coalesced over co).
// (Conv2D layers: deep_nn_models.py:142,145,157,160; the gradient Keras' fit computes for them, training.py:102-103.)
#pragma once
#include "tc3conv.cuh"

namespace s2s {

constexpr int TWG_THREADS = 256, TWG_MAXSTAGE = 4, TWG_NCI = 32;

struct TcWgArgs {
    float* part;             // [nslots][9][Cin][Cout]
    float* bias_part;        // [nslots][Cout] or null
    int N, H, W, Cin, Cout;
    int flat, nimg, tiles_x, tiles_y, ntiles;
    int groups, GS, RS;      // K groups of 8 positions per tile; B start = X + (g*GS + ky*RS + kx)*128
    int zpos, xpos;          // positions of the dZ / X boxes
    int z_stride, x_stride;  // bytes between 32-channel chunks (dZ) / between the kx tiles (X)
    int x_off, stage_bytes;  // X region offset inside a stage, stage stride
    int half_bytes;          // 3-pass: a stage = [hi half | lo half], each [dZ region | X region] of half_bytes
    int kx_tiles;            // 1: one halo tile, taps = start-address offsets; 3: one tile per kx (all starts 512-B aligned)
    int bo_mode;             // descriptor base-offset rule for unaligned starts (kx_tiles == 1): 0 none, 1 (addr>>7)&3, 2 (addr>>7)&7
    int nstage, tmem_cols;
    // Conv2DTranspose(k, strides 2) weight gradient on the same kernel: the no-halo operand is the layer INPUT x (M = Cin_T),
    // the halo operand one of the four stride-2 parity planes of dy (N = Cout_T), blockIdx.y = parity * nxc + chunk; window
    // offset `tap` of plane (a, b) is the kernel tap (ky, kx) = (2 e_y + a + pb, 2 e_x + b + pb) = tapdst[parity][tap] (or -1:
    // no such tap, MMA skipped).  The partial is [ktaps][N channels][M channels] = Keras' (k, k, Cout, Cin).  npar = 1, ktaps =
    // 9, tapdst = identity for Conv2D.
    int npar, nxc, ktaps, tapdst[4][9];
    int dbg;                 // bring-up only (S2S_TCWG_DBG): 1 = set-up and tear-down only, 2 = no epilogue stores, 3 = no MMAs
    int nissue;              // MMA-issuing warps (1..3): the taps are dealt round robin (warps 1, 6, 7)
};

// MN-major shared-memory matrix descriptor, SWIZZLE_128B_BASE32B (layout type 1): rows of 128 B = 32 tf32 along M/N, atoms of
// 4 K-rows (512 B, 32-byte chunks XOR-swizzled by the row), lbo = bytes between 32-wide M/N chunks, sbo = bytes between atoms
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)(base_off & 7) << 49) | (1ull << 61);
}
// tcgen05.mma.kind::tf32 with the two 64-bit descriptors passed as 32-bit halves (the start-address arithmetic stays 32-bit)
__device__ __forceinline__ void umma_tf32_split(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_load_4d_sw(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int NPASS, int NISSUE>       // passes (1 = tf32); MMA-issuing warps (1..3)
__global__ void __launch_bounds__(TWG_THREADS) tcwgrad_kernel(const __grid_constant__ Tc3Maps maps_x, const __grid_constant__ CUtensorMap map_z,
                                                              const TcWgArgs a) {
    extern __shared__ __align__(1024) uint8_t twg_smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(twg_smem) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[TWG_MAXSTAGE], empty_bar[TWG_MAXSTAGE], ready_bar[TWG_MAXSTAGE], acc_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(1024) float ones[256];      // B operand of the bias MMA: 8 K-rows x 32 of 1.0f

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = blockIdx.x, par = blockIdx.y / a.nxc, cic = blockIdx.y - par * a.nxc, coc = blockIdx.z;
    const CUtensorMap& map_x = maps_x.m[par];
    const int nslots = gridDim.x;
    const int my_tiles = slot < a.ntiles ? (a.ntiles - slot + nslots - 1) / nslots : 0;
    const int nstage = a.nstage;
    constexpr int NCI = TWG_NCI;
    const int mco = min(128, a.Cout - 128 * coc);            // output channels of this chunk
    const int nci = min(NCI, a.Cin - NCI * cic);             // input channels of this chunk
    const int zc = (mco + 31) >> 5;                          // 32-channel dZ boxes per tile
    const bool do_bias = a.bias_part != nullptr && cic == 0;

    if (tid == 0) {
        for (int s = 0; s < TWG_MAXSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NISSUE); mbar_init(&ready_bar[s], 128); }
        mbar_init(&acc_bar, NISSUE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 256; i += TWG_THREADS) ones[i] = 1.f;
    if (a.flat) {
        // the flat shift reads up to 2*RS + 2 positions past the X box: memory no TMA box writes.  It is multiplied by zero dZ
        // rows only, but 0 * NaN garbage would poison the sum: clear the X regions once (the boxes overwrite their part).
        for (int s = 0; s < nstage * (NPASS == 3 ? 2 : 1); ++s) {            // 3-pass: the lo halves as well (0 * NaN)
            uint8_t* x0 = base + s * a.half_bytes + a.x_off;
            for (int i = tid * 16; i < a.half_bytes - a.x_off; i += TWG_THREADS * 16)
                *reinterpret_cast<float4*>(x0 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (a.dbg == 1) {
    } else if (warp == 0) {
        if (lane == 0 && my_tiles > 0) {
            // ===== producer: per pixel tile zc boxes of dZ (32 channels each) and the X halo box(es) of this ci chunk
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_z) : "memory");
            const int tiles_per_img = a.tiles_x * a.tiles_y;
            const uint32_t tx = (uint32_t)(zc * a.zpos + a.kx_tiles * a.xpos) * 128u;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nstage;
                const int t = slot + i * nslots;
                mbar_wait_bounded(&empty_bar[s], ((i / nstage) & 1) ^ 1);
                uint8_t* sz = base + s * a.stage_bytes;
                mbar_expect_tx(&full_bar[s], tx);
                int n0, y0, x0;
                if (a.flat) { n0 = t * a.nimg; y0 = 0; x0 = 0; }
                else {
                    n0 = t / tiles_per_img;
                    const int r = t - n0 * tiles_per_img;
                    y0 = (r / a.tiles_x) * T3_TH; x0 = (r % a.tiles_x) * T3_TW;
                }
                for (int c = 0; c < zc; ++c) tma_load_4d_sw(sz + c * a.z_stride, &map_z, &full_bar[s], 128 * coc + 32 * c, x0, y0, n0);
                for (int k = 0; k < a.kx_tiles; ++k) tma_load_4d_sw(sz + a.x_off + k * a.x_stride, &map_x, &full_bar[s], NCI * cic, x0 - 1 + k, y0 - 1, n0);
            }
        }
    } else if (warp == 1 || warp >= 6) {
        const int me = warp == 1 ? 0 : warp - 5;            // issuer index
        if (lane == 0 && my_tiles > 0 && me < NISSUE) {
            // ===== MMA issuer(s): one thread per issuing warp; the ten MMAs of a K group (nine taps + bias) are dealt round robin
            // over the issuers (every tap has its own TMEM columns, so issue order across warps does not matter).
            // Instruction descriptor: D = F32, A = B = TF32, both MN-major (bits 15, 16), N = 32, M = 128.
            // Descriptor words: lo = start >> 4 | (LBO >> 4) << 16, hi = SBO >> 4 | version 1 (bit 46) | layout type 1 (bit 61).
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NCI >> 3) << 17) | (8u << 24);
            const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
            const uint32_t a_lbo = ((uint32_t)a.z_stride >> 4) << 16, b_lbo = (1024u >> 4) << 16;
            const uint32_t ones_lo = (smem_u32(ones) >> 4) | b_lbo;
            const uint32_t dbias = tmem_base + 9u * (uint32_t)NCI;
            uint32_t toff[9];                                   // tap offsets inside the X region, in 16-byte units
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap % 3;
                toff[tap] = a.kx_tiles == 3 ? (uint32_t)(kx * a.x_stride + ky * a.RS * 128) >> 4 : (uint32_t)((ky * a.RS + kx) * 128) >> 4;
            }
            const uint32_t gstep = (uint32_t)a.GS * 8;         // GS positions x 128 B, in 16-byte units
            int tmask = 0;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) tmask |= (a.tapdst[par][tap] >= 0 ? 1 : 0) << tap;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nstage;
                mbar_wait_bounded(NPASS == 3 ? &ready_bar[s] : &full_bar[s], (i / nstage) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sz = smem_u32(base + s * a.stage_bytes);
                uint32_t a_lo = (sz >> 4) | a_lbo, b_lo = ((sz + (uint32_t)a.x_off) >> 4) | b_lbo;
                const uint32_t lo_off = (uint32_t)a.half_bytes >> 4;        // 3-pass: the lo operands, same layout one half further
                uint32_t acc = i == 0 ? 0u : 1u;
                for (int g = 0; g < (a.dbg == 3 ? 0 : a.groups); ++g) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
                        if ((NISSUE == 1 || tap % NISSUE == me) && ((tmask >> tap) & 1)) {
                            const uint32_t dt = tmem_base + (uint32_t)(tap * NCI);
                            umma_tf32_split(dt, a_lo, hi, b_lo + toff[tap], hi, idesc, acc);
                            if (NPASS == 3) {      // 3xTF32: + dZ_lo X_hi + dZ_hi X_lo into the same fp32 accumulator
                                umma_tf32_split(dt, a_lo + lo_off, hi, b_lo + toff[tap], hi, idesc, 1u);
                                umma_tf32_split(dt, a_lo, hi, b_lo + toff[tap] + lo_off, hi, idesc, 1u);
                            }
                        }
                    if (do_bias && 9 % NISSUE == me) {
                        umma_tf32_split(dbias, a_lo, hi, ones_lo, hi, idesc, acc);
                        if (NPASS == 3) umma_tf32_split(dbias, a_lo + lo_off, hi, ones_lo, hi, idesc, 1u);
                    }
                    acc = 1u;
                    a_lo += 1024u >> 4;
                    b_lo += gstep;
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&acc_bar);
        }
    } else if (warp >= 2 && warp < 6) {
        // ===== epilogue: thread = output channel co (TMEM lane), 16 columns = 16 input channels of one tap per tcgen05.ld
        const int q = warp & 3;
        const int col = 32 * q + lane;
        const int co = 128 * coc + col;
        const bool co_ok = col < mco;
        const bool warp_ok = 32 * q < mco;
        float* part = a.part + (size_t)slot * a.ktaps * a.Cin * a.Cout;
        if (NPASS == 3 && a.dbg != 1) {
            // ===== operand split for 3xTF32: v = hi + lo with hi = rna_tf32(v) (in place), lo = v - hi (exact in fp32) one half
            // further; element-wise on the swizzled bytes, so both halves keep the operand layout.  Then publish to the async proxy.
            const int et = tid - 64;
            const int nz = (zc * a.z_stride) >> 4, nx0 = a.x_off >> 4, nx1 = (a.x_off + a.kx_tiles * a.x_stride) >> 4;
            for (int i = 0; i < my_tiles; ++i) {
                const int s = i % nstage;
                mbar_wait_bounded(&full_bar[s], (i / nstage) & 1);
                float4* H = reinterpret_cast<float4*>(base + s * a.stage_bytes);
                float4* L = reinterpret_cast<float4*>(base + s * a.stage_bytes + a.half_bytes);
                for (int k = et; k < nz + (nx1 - nx0); k += 128) {
                    const int idx = k < nz ? k : nx0 + (k - nz);
                    const float4 v = H[idx];
                    const float4 h = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
                    H[idx] = h;
                    L[idx] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&ready_bar[s]);
            }
        }
        if (my_tiles > 0) {
            mbar_wait_bounded(&acc_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16);
        if (warp_ok) {
            for (int c0 = 0; c0 < 9 * NCI; c0 += 16) {
                const int tap = c0 / NCI, cil = c0 - tap * NCI;
                const int dtap = a.tapdst[par][tap];
                if (cil >= nci || dtap < 0) continue;           // padded columns of the chunk / no such kernel tap for this parity
                uint32_t r[16];
                if (my_tiles > 0) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                                 : "r"(trow + (uint32_t)c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = 0u;     // a slot without tiles (batch smaller than planned) contributes zeros
                }
                if (co_ok && a.dbg != 2) {
                    float* dst = part + ((size_t)dtap * a.Cin + (size_t)cic * NCI + cil) * a.Cout + co;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (cil + j < nci) dst[(size_t)j * a.Cout] = __uint_as_float(r[j]);
                }
            }
            if (do_bias) {
                uint32_t r[8];
                if (my_tiles > 0) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                                 : "r"(trow + (uint32_t)(9 * NCI)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
                    r[0] = 0u;
                }
                if (co_ok) a.bias_part[(size_t)slot * a.Cout + co] = __uint_as_float(r[0]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
}

// ------------------------------------------------------------------ host: plan, tensor maps, launch
struct TcWgPlan {
    bool ok;
    int flat, nimg, BX, BY, ci_chunks, co_chunks, groups, GS, RS, zpos, xpos, xbox_w;
    int z_stride, x_stride, x_off, stage_bytes, half_bytes, npass, kx_tiles, nstage, tmem_cols, ntiles_max, nslots;
    size_t smem;
};

static inline int tcwg_ntiles(const TcWgPlan& p, int H, int W, int N) {
    return p.flat ? cdiv(N, p.nimg) : N * cdiv(H, T3_TH) * cdiv(W, T3_TW);
}

static inline int tcwg_default_kx_tiles() {
    static const int v = [] { const char* e = getenv("S2S_TCWG_KXTILES"); return e && e[0] == '3' ? 3 : 1; }();
    return v;
}

// Where the tensor-core kernel beats the FFMA one (wgrad.cuh), measured on B200 (profiles/r2d_tcwgrad_harness.log).  Every
// MMA contracts only 8 pixels (K = 32 bytes) and re-reads its 128-row A operand from shared memory (~50 cycles in the pipe),
// so a 128-pixel tile costs ~160 MMAs = ~4 us whatever the channel counts, on top of ~10 us of set-up / TMA / epilogue:
//     t_tc   ~ 10 us + 4 us x tile jobs per CTA            t_ffma ~ flops / (5 TFLOP/s small grids | 12 TFLOP/s large ones)
// Thin layers (8 -> 8 at 64x64: 24.6 vs 13 us) stay on the FFMA kernel; thick layers (96 -> 96 at 64x64: 60 us = 182 TFLOP/s
// vs > 2 ms) and the small-image layers of deep nets move here.  S2S_TCWG=0 | 1 forces the choice.
static inline bool tcwg_wanted(int H, int W, int Cin, int Cout, int Nmax) {
    static const int force = [] { const char* e = getenv("S2S_TCWG"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    if (force >= 0) return force == 1;
    const bool flat = H * W < 64;
    const int64_t pixels = (int64_t)Nmax * H * W;
    const double flops = 18.0 * Cin * Cout * (double)pixels;
    const double t_ffma = flops / (pixels < 262144 ? 5.0e6 : 12.0e6);                                   // us
    const int chunks = cdiv(Cin, 32) * cdiv(Cout, 128);
    const int64_t tiles = flat ? cdiv(Nmax, 4) : (int64_t)Nmax * cdiv(H, 16) * cdiv(W, 8);
    const double t_tc = 10.0 + 4.0 * (double)cdiv64(tiles * chunks, 148);
    // measured inside a step (batch 128, default net): a tensor-core wgrad CTA owns its SM (200 KB of shared memory, 512 TMEM
    // columns) and delays the tcgen05 dgrad chain of the main stream, so the middle layers (16..64 channels) stay on the FFMA
    // kernel even where the isolated kernel wins (1105 -> 1223 us per step with them on the tensor cores)
    return t_ffma > 2.0 * t_tc && ((int64_t)Cin * Cout >= 4096 || flat);
}

// transposed conv (kernel ksz, input grid h x w, Cin_T -> Cout_T): same model, four parity planes, FFMA at ~3.4 TFLOP/s
static inline bool tcwg_wanted_convt(int h, int w, int CinT, int CoutT, int ksz, int Nmax) {
    static const int force = [] { const char* e = getenv("S2S_TCWG"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    if (force >= 0) return force == 1;
    const bool flat = h * w < 64;
    const int64_t pixels = (int64_t)Nmax * h * w;
    const double t_ffma = 2.0 * ksz * ksz * CinT * CoutT * (double)pixels / (pixels < 65536 ? 3.4e6 : 12.0e6);
    const int chunks = 4 * cdiv(CoutT, 32) * cdiv(CinT, 128);
    const int64_t tiles = flat ? cdiv(Nmax, 4) : (int64_t)Nmax * cdiv(h, 16) * cdiv(w, 8);
    const double t_tc = 10.0 + 4.0 * (double)cdiv64(tiles * chunks, 148);
    return t_ffma > 1.25 * t_tc && (int64_t)ksz * ksz * CinT * CoutT >= 25000;
}

static inline int tcwg_default_nissue() {
    static const int v = [] { const char* e = getenv("S2S_TCWG_NISSUE"); return e && e[0] >= '1' && e[0] <= '3' ? e[0] - '0' : 3; }();
    return v;
}

static inline TcWgPlan tcwg_plan(int H, int W, int Cin, int Cout, int Nmax, int kx_tiles = 0, int npar = 1, int npass = 1) {
    TcWgPlan p;
    memset(&p, 0, sizeof p);
    if (Cin % 4 != 0 || Cout % 4 != 0 || Cin < 4 || Cout < 4) return p;
    if (!kx_tiles) kx_tiles = tcwg_default_kx_tiles();
    p.kx_tiles = kx_tiles;
    p.npass = npass;
    const size_t F = npass == 3 ? 2 : 1;                      // 3-pass stages hold a hi and a lo half
    p.co_chunks = cdiv(Cout, 128);
    p.ci_chunks = cdiv(Cin, TWG_NCI);
    p.tmem_cols = 512;                                       // 9 taps x 32 columns + 32 (bias)
    const int zc = std::min(4, cdiv(Cout, 32));              // 32-channel dZ boxes per stage
    const size_t budget = 218 * 1024;                        // dynamic shared memory we allow ourselves (227 KB per CTA exist, ~2 KB static)
    auto stage_of = [&](int zpos, int xpos, int over, int* x_stride) {
        *x_stride = ((xpos + over) * 128 + 1023) / 1024 * 1024;
        return (size_t)zc * zpos * 128 + (size_t)kx_tiles * *x_stride;
    };
    p.flat = (H * W < 64) ? 1 : 0;
    if (p.flat) {
        p.BY = H + 2;
        bool found = false;
        for (int nimg = 16; nimg >= 1 && !found; nimg >>= 1) {
            for (int bx = W + 2; bx <= W + 9; ++bx) {
                const int pos = nimg * p.BY * bx;
                if (pos % 8 || (kx_tiles == 3 && bx % 4)) continue;
                int xs;
                const size_t st = stage_of(pos, pos, 2 * bx + 2 + 8, &xs);
                const size_t over = (size_t)4 * pos * 128 > st ? (size_t)4 * pos * 128 - st : 0;
                if ((npass == 3 ? 1 : 2) * F * st + over > budget || bx > 256 || p.BY > 256 || nimg > 256) break;
                p.nimg = nimg; p.BX = bx;
                found = true;
                break;
            }
        }
        if (!found) return p;
        p.zpos = p.xpos = p.nimg * p.BY * p.BX;
        p.xbox_w = p.BX;
        p.groups = p.zpos / 8; p.GS = 8; p.RS = p.BX;
    } else {
        p.nimg = 1; p.BY = T3_HH;
        p.zpos = T3_TH * T3_TW;
        p.xbox_w = kx_tiles == 3 ? T3_TW : T3_HW;            // three 8-wide boxes shifted by kx, or one 10-wide halo box
        p.BX = p.xbox_w;
        p.xpos = T3_HH * p.xbox_w;
        p.groups = T3_TH; p.GS = p.xbox_w; p.RS = p.xbox_w;
    }
    p.z_stride = p.zpos * 128;
    const size_t st = stage_of(p.zpos, p.xpos, p.flat ? 2 * p.RS + 2 + 8 : 0, &p.x_stride);
    p.x_off = zc * p.z_stride;
    p.half_bytes = (int)st;
    p.stage_bytes = (int)(F * st);
    // M = 128 rows (four 32-channel chunks) are always read from the dZ region: keep the over-read of the last stage in bounds
    const size_t slack = (size_t)4 * p.z_stride > st ? (size_t)4 * p.z_stride - st : 0;
    int ns = (int)std::min<size_t>(TWG_MAXSTAGE, (budget - slack) / (F * st));
    if (ns < 1) return p;
    p.nstage = ns;
    p.smem = (size_t)ns * F * st + slack + 1024;
    p.ntiles_max = tcwg_ntiles(p, H, W, Nmax);
    int sl = 148 / (npar * p.ci_chunks * p.co_chunks);
    if (sl < 1) sl = 1;
    p.nslots = std::min(sl, p.ntiles_max);
    p.ok = true;
    return p;
}

// 4-D view (c, x, y, n) of an NHWC tensor [N, H, W, ld] with 32-channel boxes, swizzle 128B with 32-byte atoms: a box lands
// in shared memory as [position][32 channels] = the MN-major SWIZZLE_128B_BASE32B operand layout of tcgen05.mma.kind::tf32
static inline int tcwg_make_map(const float* t, int N, int H, int W, int C, int ld, int bx, int by, int bn, CUtensorMap* m) {
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bn}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)t, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled(tcwgrad: C=%d ld=%d H=%d W=%d box %d x %d x %d) failed: %d", C, ld, H, W, bx, by, bn, (int)r);
    return 0;
}

// maps for N images (N is baked in: images past N must read as zero in the flat geometry)
static inline int tcwg_make_maps(const TcWgPlan& p, const float* x, int ldx, const float* dz, int lddz, int N, int H, int W, int Cin, int Cout,
                                 CUtensorMap* mx, CUtensorMap* mz) {
    S2S_CHECK(tcwg_make_map(x, N, H, W, Cin, ldx, p.xbox_w, p.BY, p.nimg, mx));
    if (p.flat) return tcwg_make_map(dz, N, H, W, Cout, lddz, p.BX, p.BY, p.nimg, mz);
    return tcwg_make_map(dz, N, H, W, Cout, lddz, T3_TW, T3_TH, 1, mz);
}

// strided variant: the parity planes of a transposed conv's dy
static inline int tcwg_make_map_strided(const float* t, int N, int H, int W, int C, int64_t sx, int64_t sy, int64_t sn, int bx, int by, int bn,
                                        CUtensorMap* m) {
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sx * 4, (cuuint64_t)sy * 4, (cuuint64_t)sn * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bn}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)t, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled(tcwgrad strided: C=%d H=%d W=%d) failed: %d", C, H, W, (int)r);
    return 0;
}
// transposed-conv wgrad: x [N, h, w, Cin_T] is the no-halo (M) operand, the four parity planes of dy [N, 2h, 2w, ld] (dy points
// at the first channel of the slice) the halo (N) operands; the plan was made with (Cin = Cout_T, Cout = Cin_T)
static inline int tcwg_make_maps_convt(const TcWgPlan& p, const float* x, int ldx, const float* dy, int lddy, int N, int h, int w, int CinT, int CoutT,
                                       Tc3Maps* mx, CUtensorMap* mz) {
    for (int par = 0; par < 4; ++par) {
        const float* base = dy + ((size_t)(par >> 1) * 2 * w + (par & 1)) * lddy;
        S2S_CHECK(tcwg_make_map_strided(base, N, h, w, CoutT, 2 * (int64_t)lddy, 4 * (int64_t)w * lddy, 4 * (int64_t)h * w * lddy, p.xbox_w, p.BY,
                                        p.nimg, &mx->m[par]));
    }
    if (p.flat) return tcwg_make_map(x, N, h, w, CinT, ldx, p.BX, p.BY, p.nimg, mz);
    return tcwg_make_map(x, N, h, w, CinT, ldx, T3_TW, T3_TH, 1, mz);
}

// ksz = 0: Conv2D 3x3 (mx.m[0] only); ksz = 2 | 3 | 5: Conv2DTranspose (four parity maps, Cin = Cout_T, Cout = Cin_T, no bias)
static inline int tcwg_launch_maps(const Tc3Maps& mx, const CUtensorMap& mz, const TcWgPlan& p, float* part, float* bias_part, int N, int H, int W,
                                   int Cin, int Cout, int nslots, int ksz, cudaStream_t st, int bo_mode = 0, int nissue = 0) {
    S2S_REQUIRE(p.ok, "tcwgrad: no plan for %d -> %d", Cin, Cout);
    TcWgArgs a;
    memset(&a, 0, sizeof a);
    a.part = part; a.bias_part = bias_part; a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    a.flat = p.flat; a.nimg = p.nimg;
    a.tiles_x = cdiv(W, T3_TW); a.tiles_y = cdiv(H, T3_TH); a.ntiles = tcwg_ntiles(p, H, W, N);
    a.groups = p.groups; a.GS = p.GS; a.RS = p.RS; a.zpos = p.zpos; a.xpos = p.xpos;
    a.z_stride = p.z_stride; a.x_stride = p.x_stride; a.x_off = p.x_off; a.stage_bytes = p.stage_bytes; a.half_bytes = p.half_bytes;
    a.kx_tiles = p.kx_tiles; a.bo_mode = bo_mode;
    a.npar = ksz ? 4 : 1; a.nxc = p.ci_chunks; a.ktaps = ksz ? ksz * ksz : 9;
    for (int par = 0; par < 4; ++par)
        for (int tap = 0; tap < 9; ++tap) {
            int d = tap;
            if (ksz) {
                const int pb = (ksz - 2) / 2;
                const int ky = 2 * (tap / 3 - 1) + (par >> 1) + pb, kx = 2 * (tap % 3 - 1) + (par & 1) + pb;
                d = (ky >= 0 && ky < ksz && kx >= 0 && kx < ksz) ? ky * ksz + kx : -1;
            }
            a.tapdst[par][tap] = d;
        }
    a.nstage = p.nstage; a.tmem_cols = p.tmem_cols;
    a.nissue = nissue ? nissue : tcwg_default_nissue();
    { static const int dbg = [] { const char* e = getenv("S2S_TCWG_DBG"); return e ? atoi(e) : 0; }(); a.dbg = dbg; }
    if (ksz) prof_begin(st, p.npass == 3 ? "convT_wgrad_3xtf32" : "convT_wgrad_tf32", 4.0 * N * H * W * (4.0 * Cin + Cout), 2.0 * ksz * ksz * (double)Cin * Cout * N * H * W);
    else prof_begin(st, p.npass == 3 ? "conv3x3_wgrad_3xtf32" : "conv3x3_wgrad_tf32", 4.0 * N * H * W * ((double)Cin + Cout), 18.0 * (double)Cin * Cout * N * H * W);
    const dim3 grid(nslots, a.npar * p.ci_chunks, p.co_chunks);
#define S2S_TWG(NI)                                                                                                             \
    {                                                                                                                           \
        static DevOnce once;                                                                                                    \
        S2S_CUDA(once.run([] { return cudaFuncSetAttribute(tcwgrad_kernel<1, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024); })); \
        tcwgrad_kernel<1, NI><<<grid, TWG_THREADS, p.smem, st>>>(mx, mz, a);                                                    \
    }
    if (p.npass == 3) {
        static DevOnce once3;
        S2S_CUDA(once3.run([] { return cudaFuncSetAttribute(tcwgrad_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024); }));
        tcwgrad_kernel<3, 3><<<grid, TWG_THREADS, p.smem, st>>>(mx, mz, a);
    } else if (a.nissue == 1) S2S_TWG(1) else if (a.nissue == 2) S2S_TWG(2) else S2S_TWG(3)
#undef S2S_TWG
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

static inline int tcwg_launch(const CUtensorMap& mx, const CUtensorMap& mz, const TcWgPlan& p, float* part, float* bias_part, int N, int H, int W,
                              int Cin, int Cout, int nslots, cudaStream_t st, int bo_mode = 0, int nissue = 0) {
    Tc3Maps ms;
    for (int i = 0; i < 4; ++i) ms.m[i] = mx;
    return tcwg_launch_maps(ms, mz, p, part, bias_part, N, H, W, Cin, Cout, nslots, 0, st, bo_mode, nissue);
}

}  // namespace s2s
