// skill.cuh — per-gridpoint skill reductions over the start-date axis T (HBM-bound).
//
//   RPS / RPSS   utils/performance_metrics.py:26-45 (xskillscore.rps, input_distributions='p')
//   CC / ACC     ACCs.ipynb:362-388 (xr.corr over T of raw fields / ISO-week anomalies)
//   ensemble mean over members, MME probability combine (preprocessing.py:21-23, training.py:344-350)
//
// Layout: lanes of a warp are 32 consecutive gridpoints (coalesced rows of the [T,Y,X(,3)] arrays),
// the 8 warps of a CTA split T (RPS) or the ISO-week groups (ACC); the slices are combined in
// fixed order through shared memory, so results are bit-reproducible.
#pragma once
#include "common.cuh"

namespace s2s {

constexpr int SK_TS = 8;   // T slices (warps) per CTA

template <bool SS>
__global__ void __launch_bounds__(32 * SK_TS) rps_kernel(const float* __restrict__ f, const float* __restrict__ r,
                                                         const float* __restrict__ o, int T, int64_t YX,
                                                         float* __restrict__ out) {
    __shared__ double sred[SK_TS][3][32];
    const int lane = threadIdx.x, ts = threadIdx.y;
    const int64_t gp = (int64_t)blockIdx.x * 32 + lane;
    double sf = 0.0, sr = 0.0, cnt = 0.0;
    if (gp < YX) {
        for (int t = ts; t < T; t += SK_TS) {
            const size_t base = ((size_t)t * YX + gp) * 3;
            const float o0 = __ldg(o + base), o1 = __ldg(o + base + 1), o2 = __ldg(o + base + 2);
            if (isnan(o0) || isnan(o1) || isnan(o2)) continue;
            {
                const float p0 = __ldg(f + base), p1 = __ldg(f + base + 1), p2 = __ldg(f + base + 2);
                const float c1 = p0 - o0, c2 = (p0 + p1) - (o0 + o1), c3 = ((p0 + p1) + p2) - ((o0 + o1) + o2);
                sf += (double)(c1 * c1 + c2 * c2 + c3 * c3);
            }
            if (SS) {
                const float p0 = __ldg(r + base), p1 = __ldg(r + base + 1), p2 = __ldg(r + base + 2);
                const float c1 = p0 - o0, c2 = (p0 + p1) - (o0 + o1), c3 = ((p0 + p1) + p2) - ((o0 + o1) + o2);
                sr += (double)(c1 * c1 + c2 * c2 + c3 * c3);
            }
            cnt += 1.0;
        }
    }
    sred[ts][0][lane] = sf; sred[ts][1][lane] = sr; sred[ts][2][lane] = cnt;
    __syncthreads();
    if (ts == 0 && gp < YX) {
        double a = 0.0, b = 0.0, n = 0.0;
        for (int k = 0; k < SK_TS; ++k) { a += sred[k][0][lane]; b += sred[k][1][lane]; n += sred[k][2][lane]; }
        float res;
        if (n == 0.0) res = __int_as_float(0x7fc00000);
        else if (SS) res = (float)(1.0 - (a / n) / (b / n));
        else res = (float)(a / n);
        out[gp] = res;
    }
}

// CC and ACC in ONE pass over the data.  order[T]: start indices sorted by ISO-week group; gstart[G+1]: offsets of
// each group in order[].  Anomalies are value - mean over the starts of the same ISO week (ACCs.ipynb:369-376), each
// variable's mean over its own valid starts (xarray mean skips NaN), correlation over pairwise-valid starts
// (xr.corr).  Per group the kernel accumulates {n_x, S x, n_y, S y} and the pairwise sums {n, S x, S y, S xx, S yy,
// S xy} of pivot-shifted values and expands the anomaly sums algebraically at the end of the group:
//   S a = S_p x - n mx,  S aa = S_p xx - 2 mx S_p x + n mx^2,  S ab = S_p xy - mx S_p y - my S_p x + n mx my,
// so every element is read once (HBM-bound: 8 B per start and gridpoint) with 4 independent loads in flight.
__global__ void __launch_bounds__(32 * SK_TS, 3) acc_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         const int* __restrict__ order, const int* __restrict__ gstart,
                                                         int G, int64_t YX, float* __restrict__ acc_out,
                                                         float* __restrict__ cc_out) {
    __shared__ double sred[SK_TS][12][32];
    const int lane = threadIdx.x, ts = threadIdx.y;
    const int64_t gp = (int64_t)blockIdx.x * 32 + lane;
    // the 12 running sums live in this thread's own shared-memory slots (not registers: the kernel ran at 171 registers
    // = 1 CTA per SM and 21 % of the HBM peak with them; bytes in flight per SM are what this kernel needs)
    double* s = &sred[ts][0][lane];
    constexpr int SS_ = 32;            // stride between the 12 slots of a thread
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k * SS_] = 0.0;
    if (gp < YX) {
#pragma unroll 1
        for (int g = ts; g < G; g += SK_TS) {
            const int t0 = gstart[g], t1 = gstart[g + 1];
            if (t1 <= t0) continue;
            // fp32 arithmetic on values shifted by a per-group pivot (the group's first start), flushed to double every
            // 8 starts: products of shifted values keep ~1e-7 relative accuracy without the mean^2/variance
            // cancellation, and the double-precision pipe sees 8x fewer operations
            const size_t e0 = (size_t)order[t0] * YX + gp;
            float xp = __ldg(x + e0), yp = __ldg(y + e0);
            if (!(xp == xp)) xp = 0.f;
            if (!(yp == yp)) yp = 0.f;
            int nx = 0, ny = 0, n = 0;
            double sx = 0.0, sy = 0.0, px = 0.0, py = 0.0, pxx = 0.0, pyy = 0.0, pxy = 0.0;
            int i = t0;
#pragma unroll 1
            for (; i + 8 <= t1; i += 8) {
                float xv[8], yv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const size_t e = (size_t)order[i + u] * YX + gp;
                    xv[u] = __ldg(x + e); yv[u] = __ldg(y + e);
                }
                float csx = 0.f, csy = 0.f, cpx = 0.f, cpy = 0.f, cxx = 0.f, cyy = 0.f, cxy = 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool vx = xv[u] == xv[u], vy = yv[u] == yv[u], vp = vx && vy;
                    const float xs = vx ? xv[u] - xp : 0.f, ys = vy ? yv[u] - yp : 0.f;
                    nx += vx; ny += vy; n += vp;
                    csx += xs; csy += ys;
                    const float xq = vp ? xs : 0.f, yq = vp ? ys : 0.f;
                    cpx += xq; cpy += yq;
                    cxx = fmaf(xq, xq, cxx); cyy = fmaf(yq, yq, cyy); cxy = fmaf(xq, yq, cxy);
                }
                sx += (double)csx; sy += (double)csy; px += (double)cpx; py += (double)cpy;
                pxx += (double)cxx; pyy += (double)cyy; pxy += (double)cxy;
            }
#pragma unroll 1
            for (; i < t1; ++i) {
                const size_t e = (size_t)order[i] * YX + gp;
                const float xv = __ldg(x + e), yv = __ldg(y + e);
                const bool vx = xv == xv, vy = yv == yv, vp = vx && vy;
                const float xs = vx ? xv - xp : 0.f, ys = vy ? yv - yp : 0.f;
                nx += vx; ny += vy; n += vp;
                sx += (double)xs; sy += (double)ys;
                if (vp) { px += (double)xs; py += (double)ys; pxx += (double)(xs * xs); pyy += (double)(ys * ys); pxy += (double)(xs * ys); }
            }
            // shifted group means (each variable over its own valid starts), anomaly sums over the valid pairs
            const double dn = (double)n;
            const double mx = nx > 0 ? sx / (double)nx : 0.0, my = ny > 0 ? sy / (double)ny : 0.0;
            s[0 * SS_] += dn;
            s[6 * SS_] += px - dn * mx;
            s[7 * SS_] += py - dn * my;
            s[8 * SS_] += pxx - 2.0 * mx * px + dn * mx * mx;
            s[9 * SS_] += pyy - 2.0 * my * py + dn * my * my;
            s[10 * SS_] += pxy - mx * py - my * px + dn * mx * my;
            // raw sums for CC: undo the pivot shift
            const double dxp = (double)xp, dyp = (double)yp;
            s[1 * SS_] += px + dn * dxp;
            s[2 * SS_] += py + dn * dyp;
            s[3 * SS_] += pxx + 2.0 * dxp * px + dn * dxp * dxp;
            s[4 * SS_] += pyy + 2.0 * dyp * py + dn * dyp * dyp;
            s[5 * SS_] += pxy + dxp * py + dyp * px + dn * dxp * dyp;
        }
    }
    __syncthreads();
    if (ts == 0 && gp < YX) {
        double r[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            double v = 0.0;
#pragma unroll 1
            for (int w = 0; w < SK_TS; ++w) v += sred[w][k][lane];      // not unrolled: 96 hoisted loads cost 170 registers
            r[k] = v;
        }
        const double n = r[0];
        float cc = __int_as_float(0x7fc00000), ac = cc;
        if (n > 0.0) {
            {
                const double mx = r[1] / n, my = r[2] / n;
                const double cov = r[5] / n - mx * my, vx = r[3] / n - mx * mx, vy = r[4] / n - my * my;
                cc = (float)(cov / sqrt(vx * vy));
            }
            {
                const double mx = r[6] / n, my = r[7] / n;
                const double cov = r[10] / n - mx * my, vx = r[8] / n - mx * mx, vy = r[9] / n - my * my;
                ac = (float)(cov / sqrt(vx * vy));
            }
        }
        if (cc_out) cc_out[gp] = cc;
        if (acc_out) acc_out[gp] = ac;
    }
}

// x [T,M,YX] -> out [T,YX], NaN members skipped (xarray .mean('M'))
__global__ void ensemble_mean_kernel(const float* __restrict__ x, int M, int64_t YX, int64_t total, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t t = i / YX, g = i % YX;
    float s = 0.f, n = 0.f;
    for (int m = 0; m < M; ++m) {
        const float v = __ldg(x + ((size_t)t * M + m) * YX + g);
        if (!isnan(v)) { s += v; n += 1.f; }
    }
    out[i] = n > 0.f ? s / n : __int_as_float(0x7fc00000);
}

// probs [n_models][n_points][3] -> mean over models, renormalised over the category axis
__global__ void mme_combine_kernel(const float* __restrict__ p, int n_models, int64_t n_points, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int m = 0; m < n_models; ++m) {
        const float* q = p + ((size_t)m * n_points + i) * 3;
        a0 += __ldg(q); a1 += __ldg(q + 1); a2 += __ldg(q + 2);
    }
    const float inv = 1.f / (float)n_models;
    a0 *= inv; a1 *= inv; a2 *= inv;
    const float s = (a0 + a1) + a2;
    out[i * 3 + 0] = a0 / s; out[i * 3 + 1] = a1 / s; out[i * 3 + 2] = a2 / s;
}

}  // namespace s2s
