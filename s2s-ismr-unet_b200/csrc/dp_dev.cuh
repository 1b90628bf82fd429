// dp_dev.cuh — device-side primitives of the peer-memory data-parallel exchange (see dp.cuh): the per-rank view of
// the mapped exchange buffers, system-scope flag signal / wait with a wall-clock timeout, coherent peer loads.
#pragma once
#include "common.cuh"

namespace s2s {

constexpr int DP_MAXW = 8;            // ranks per node (8 x B200)
constexpr int DP_MAXSYNC = 24;        // BN sync points per step: (2*MAXB+1) layers x {forward, backward}
constexpr int DP_BN_MAXC = 512;
constexpr unsigned long long DP_TIMEOUT_NS = 8000000000ull;

struct DpDev {
    int rank, world;
    unsigned long long* epoch;            // local: completed steps (bumped by the last CTA of dp_sum_adam)
    int* error;                           // local: != 0 after a timeout
    unsigned int* counter;                // local: last-CTA election
    unsigned long long* flags[DP_MAXW];   // flags[p] -> rank p's flag array [(1 + DP_MAXSYNC)][DP_MAXW]
    double* bn[DP_MAXW];                  // bn[p]    -> rank p's BN sums  [DP_MAXSYNC][2][DP_MAXW][2 * DP_BN_MAXC]
    float* grads[DP_MAXW];                // grads[p] -> rank p's gradient exchange area: pull [2][n_pad], push [2][world][n_pad]
    float* stats[DP_MAXW];                // stats[p] -> rank p's {loss * n, correct-fraction * n, n}: pull [2][4], push [2][DP_MAXW][4]
    size_t n_pad;
    int push;                             // 1: every rank WRITES its gradient into slot [parity][rank] of every peer (posted NVLink
                                          // stores from the reduction kernel); the sum + Adam kernel then reads local memory only
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long dp_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {      // system-coherent 16 B load (peer or local)
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Publish "group `grp` of step `e` is ready on this rank" to every rank.  One CTA only; the caller's data writes
// must precede (stream order for earlier kernels, __syncthreads for this CTA's own writes).
__device__ __forceinline__ void dp_signal(const DpDev& d, int grp, unsigned long long e) {
    if ((int)threadIdx.x < d.world) {
        __threadfence_system();
        st_release_sys(d.flags[threadIdx.x] + grp * DP_MAXW + d.rank, e);
    }
}
// Every calling CTA waits until all ranks have published group `grp` of step `e`.
__device__ __forceinline__ void dp_wait(const DpDev& d, int grp, unsigned long long e) {
    if ((int)threadIdx.x < d.world) {
        const unsigned long long* f = d.flags[d.rank] + grp * DP_MAXW + threadIdx.x;
        const unsigned long long t0 = dp_now_ns();
        while (ld_acquire_sys(f) < e) {
            if (dp_now_ns() - t0 > DP_TIMEOUT_NS) { *d.error = 1 + grp; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

// Exchange of per-channel sums between the ranks from INSIDE a consumer kernel (sync-BN): `sums[0..V)` (shared memory,
// identical in every CTA of this rank) are this rank's local sums; on return they hold the sums over all ranks, added
// in rank order (bit-identical on every rank and CTA).  PUSH model: CTA 0 stores the rank's sums into slot [rank] of
// EVERY peer's buffer (posted NVLink writes), fences, and releases the flag; every CTA then waits on local flags and
// reads local memory only — one one-way NVLink latency per exchange instead of a flag plus a remote-load round trip.
// Buffer layout per rank: [sync][parity][source rank][2 * DP_BN_MAXC] doubles.  blockDim.x >= world; all threads call.
__device__ __forceinline__ size_t dp_bn_off(int sync_id, unsigned long long e, int src) {
    return (((size_t)sync_id * 2 + (size_t)(e & 1)) * DP_MAXW + (size_t)src) * (2 * DP_BN_MAXC);
}
__device__ __forceinline__ void dp_exchange_sums(const DpDev& d, int sync_id, double* sums, int V, int tid, int nthreads) {
    const unsigned long long e = *d.epoch + 1;
    if (blockIdx.x == 0) {
        const size_t off = dp_bn_off(sync_id, e, d.rank);
        for (int i = tid; i < V * d.world; i += nthreads) {
            const int r = i / V, c = i - r * V;
            if (r != d.rank) d.bn[r][off + c] = sums[c];
        }
        __syncthreads();
        dp_signal(d, 1 + sync_id, e);
    }
    dp_wait(d, 1 + sync_id, e);
    const double* mine = d.bn[d.rank];
    for (int c = tid; c < V; c += nthreads) {
        double s = 0.0;
        for (int r = 0; r < d.world; ++r) s += (r == d.rank) ? sums[c] : ld_peer_f64(mine + dp_bn_off(sync_id, e, r) + c);
        sums[c] = s;
    }
    __syncthreads();
}

}  // namespace s2s
