// elr.cu — extended logistic regression (ELR) baseline, one IRLS fit per gridpoint, batched on the GPU.
//
// Replaces the Python Y x X loop around statsmodels in train_single_bootstrap_ELR (utils/training.py:402-530):
//   per gridpoint, rows = the 2T pairs (start t, threshold q in {33, 67}); design [1, x_t, q] with x = ensemble-mean
//   forecast (:413-416, 439-446); response = [y_t <= edge_q(t)] from rolling_labeler_ELR (preprocessing.py:270-333),
//   rows dropped where the week's edges are NaN, e0 == 0 or e0 == e1 (:304-308); sm.GLM(..., Binomial()).fit()
//   (:487-489); P(below) = p33, P(normal) = p67 - p33, P(above) = 1 - p67 (:503-505, 519-521); the other starts of a
//   fitted gridpoint get 1/3 (:507, 522); gridpoints with a NaN in ytrain, no valid row, a NaN predictor or <= 2 rows
//   stay NaN (:431-432, 463-480).
//
// statsmodels 0.14.4 GLM._fit_irls, restated: mu0 = (y + 0.5)/2, eta = logit(mu); repeat (<= 100): w = mu(1-mu)
//   (clipped to [eps, 1-eps]), z = eta + (y - mu)/(mu(1-mu)), beta = WLS(z ~ X, w), eta = X beta, mu = 1/(1+exp(-eta)),
//   deviance = 2 sum [y log(clip(y/(mu+1e-20))) + (1-y) log(clip((1-y)/(1-mu+1e-20)))]; stop when |dev - dev_prev| <= 1e-8.
//   The WLS is solved through 3x3 normal equations in a centred / scaled basis (x -> (x - mean)/sd, q -> +-1): the
//   fitted linear predictor is basis-invariant, and the basis keeps the normal equations well conditioned in double.
//
// thread = gridpoint; every pass streams the gridpoint's column of x / y ([T, YX] layout -> coalesced over the warp).
// HBM/L2-bound: (8 + sizeof(y)) * T * YX bytes per IRLS pass; double-precision math is negligible.
#include <float.h>
#include "common.cuh"

namespace s2s {

struct ElrArgs {
    const double* x_train; const void* y_train; const int32_t* slot_train;
    const double* x_test; const int32_t* slot_test;
    const double* edges;           // [nW][2][YX]
    int T, Tt; int64_t YX;
    double* p_train; double* p_test;       // [T, YX, 3], [Tt, YX, 3]
    int32_t* iters;                // [YX] IRLS iterations (0: skipped gridpoint), nullable
    int maxiter; double tol;
};

__device__ __forceinline__ double elr_clip(double p) { return fmin(fmax(p, DBL_EPSILON), 1.0 - DBL_EPSILON); }

// solve the symmetric 3x3 system A b = r (A given by its 6 upper entries) by Gaussian elimination with partial pivoting
__device__ __forceinline__ bool elr_solve3(const double a[6], const double r[3], double b[3]) {
    double m[3][4] = {{a[0], a[1], a[2], r[0]}, {a[1], a[3], a[4], r[1]}, {a[2], a[4], a[5], r[2]}};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        double best = fabs(m[c][c]);
#pragma unroll
        for (int k = c + 1; k < 3; ++k)
            if (fabs(m[k][c]) > best) { best = fabs(m[k][c]); piv = k; }
        if (best < 1e-300) return false;
        if (piv != c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { const double t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
        }
#pragma unroll
        for (int k = c + 1; k < 3; ++k) {
            const double f = m[k][c] / m[c][c];
#pragma unroll
            for (int j = c; j < 4; ++j) m[k][j] -= f * m[c][j];
        }
    }
    b[2] = m[2][3] / m[2][2];
    b[1] = (m[1][3] - m[1][2] * b[2]) / m[1][1];
    b[0] = (m[0][3] - m[0][1] * b[1] - m[0][2] * b[2]) / m[0][0];
    return true;
}

template <typename TY>
__global__ void __launch_bounds__(128) elr_kernel(const ElrArgs a) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.YX) return;
    const TY* y = reinterpret_cast<const TY*>(a.y_train);
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    const int64_t YX = a.YX;

    auto masked = [&](int slot) {
        const double e0 = a.edges[((int64_t)slot * 2 + 0) * YX + g], e1 = a.edges[((int64_t)slot * 2 + 1) * YX + g];
        return (e0 != e0) || (e1 != e1) || e0 == 0.0 || e0 == e1;
    };
    auto fill = [&](double* p, int n, double v) {
        for (int t = 0; t < n; ++t) {
            p[((int64_t)t * YX + g) * 3 + 0] = v; p[((int64_t)t * YX + g) * 3 + 1] = v; p[((int64_t)t * YX + g) * 3 + 2] = v;
        }
    };

    // ---- eligibility (training.py:431-432, 463-480) and the centred / scaled basis of x
    bool skip = false;
    int nv = 0, nvt = 0;
    double sx = 0.0;
    for (int t = 0; t < a.T; ++t) {
        const double yv = (double)y[(int64_t)t * YX + g];
        if (yv != yv) skip = true;
        if (!masked(a.slot_train[t])) {
            const double xv = a.x_train[(int64_t)t * YX + g];
            if (xv != xv) skip = true;
            sx += xv; ++nv;
        }
    }
    for (int t = 0; t < a.Tt; ++t)
        if (!masked(a.slot_test[t])) {
            const double xv = a.x_test[(int64_t)t * YX + g];
            if (xv != xv) skip = true;
            ++nvt;
        }
    if (skip || nv == 0 || 2 * nv <= 2 || 2 * nvt <= 2) {
        fill(a.p_train, a.T, NaN); fill(a.p_test, a.Tt, NaN);
        if (a.iters) a.iters[g] = 0;
        return;
    }
    const double xm = sx / nv;
    double sxx = 0.0;
    for (int t = 0; t < a.T; ++t)
        if (!masked(a.slot_train[t])) { const double d = a.x_train[(int64_t)t * YX + g] - xm; sxx += d * d; }
    const double xs = sxx > 0.0 ? sqrt(sxx / nv) : 1.0;

    // ---- IRLS (statsmodels GLM._fit_irls).  beta is in the basis [1, (x - xm)/xs, +-1]
    double beta[3] = {0.0, 0.0, 0.0};
    double dev_prev = 0.0;
    bool have_beta = false;
    int it = 0;
    for (; it <= a.maxiter; ++it) {
        double A[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0}, dev = 0.0;
        for (int t = 0; t < a.T; ++t) {
            const int slot = a.slot_train[t];
            if (masked(slot)) continue;
            const double yv = (double)y[(int64_t)t * YX + g];
            const double z1 = (a.x_train[(int64_t)t * YX + g] - xm) / xs;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double e = a.edges[((int64_t)slot * 2 + q) * YX + g];
                const double yy = yv <= e ? 1.0 : 0.0;
                const double z2 = q == 0 ? -1.0 : 1.0;
                double mu, eta;
                if (!have_beta) {             // starting values: mu = (y + 0.5)/2, eta = logit(mu)
                    mu = (yy + 0.5) * 0.5;
                    const double pc = elr_clip(mu);
                    eta = log(pc / (1.0 - pc));
                } else {
                    eta = beta[0] + beta[1] * z1 + beta[2] * z2;
                    mu = 1.0 / (1.0 + exp(-eta));
                }
                // deviance of the CURRENT mu (statsmodels appends it after each update)
                const double c1 = fmax(yy / (mu + 1e-20), DBL_EPSILON), c0 = fmax((1.0 - yy) / (1.0 - mu + 1e-20), DBL_EPSILON);
                dev += 2.0 * (yy * log(c1) + (1.0 - yy) * log(c0));
                const double pc = elr_clip(mu);
                const double w = pc * (1.0 - pc);                 // 1 / (deriv^2 * variance)
                const double zz = eta + (yy - mu) / (pc * (1.0 - pc));
                A[0] += w; A[1] += w * z1; A[2] += w * z2; A[3] += w * z1 * z1; A[4] += w * z1 * z2; A[5] += w * z2 * z2;
                r[0] += w * zz; r[1] += w * z1 * zz; r[2] += w * z2 * zz;
            }
        }
        // `dev` belongs to the parameters fitted in the previous pass: convergence test of that update
        if (it >= 1 && fabs(dev - dev_prev) <= a.tol) break;
        if (it == a.maxiter) break;
        dev_prev = dev;
        double nb[3];
        bool ok = elr_solve3(A, r, nb);
        if (!ok || sxx == 0.0) {                // constant predictor: drop its column (min-norm solution has beta_x = 0)
            const double det = A[0] * A[5] - A[2] * A[2];
            nb[1] = 0.0;
            if (fabs(det) > 1e-300) { nb[0] = (r[0] * A[5] - r[2] * A[2]) / det; nb[2] = (A[0] * r[2] - A[2] * r[0]) / det; }
            else { nb[0] = r[0] / A[0]; nb[2] = 0.0; }
        }
        beta[0] = nb[0]; beta[1] = nb[1]; beta[2] = nb[2];
        have_beta = true;
    }
    if (a.iters) a.iters[g] = it;      // WLS updates performed

    // ---- predictions (training.py:491-522)
    auto predict = [&](const double* xs_, const int32_t* slots, int n, double* p) {
        for (int t = 0; t < n; ++t) {
            double p0 = 1.0 / 3.0, p1 = 1.0 / 3.0, p2 = 1.0 / 3.0;
            if (!masked(slots[t])) {
                const double z1 = (xs_[(int64_t)t * YX + g] - xm) / xs;
                const double q33 = 1.0 / (1.0 + exp(-(beta[0] + beta[1] * z1 - beta[2])));
                const double q67 = 1.0 / (1.0 + exp(-(beta[0] + beta[1] * z1 + beta[2])));
                p0 = q33; p1 = q67 - q33; p2 = 1.0 - q67;
            }
            p[((int64_t)t * YX + g) * 3 + 0] = p0; p[((int64_t)t * YX + g) * 3 + 1] = p1; p[((int64_t)t * YX + g) * 3 + 2] = p2;
        }
    };
    predict(a.x_train, a.slot_train, a.T, a.p_train);
    predict(a.x_test, a.slot_test, a.Tt, a.p_test);
}

}  // namespace s2s

using namespace s2s;

extern "C" int s2s_elr_fit_predict(const double* x_train_dev, const void* y_train_dev, int y_is_f64, const int32_t* slot_train_dev,
                                   const double* x_test_dev, const int32_t* slot_test_dev, const double* edges_dev, int T, int Tt,
                                   int64_t YX, double* p_train_dev, double* p_test_dev, int32_t* iters_dev, void* stream) {
    S2S_REQUIRE(x_train_dev && y_train_dev && slot_train_dev && x_test_dev && slot_test_dev && edges_dev && p_train_dev && p_test_dev,
                "null argument");
    S2S_REQUIRE(T >= 1 && Tt >= 1 && YX >= 1, "bad sizes");
    ElrArgs a;
    a.x_train = x_train_dev; a.y_train = y_train_dev; a.slot_train = slot_train_dev;
    a.x_test = x_test_dev; a.slot_test = slot_test_dev; a.edges = edges_dev;
    a.T = T; a.Tt = Tt; a.YX = YX; a.p_train = p_train_dev; a.p_test = p_test_dev; a.iters = iters_dev;
    a.maxiter = 100; a.tol = 1e-8;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)cdiv64(YX, 128);
    prof_begin(st, "elr_irls", 0.0, 0.0);
    if (y_is_f64) elr_kernel<double><<<grid, 128, 0, st>>>(a);
    else elr_kernel<float><<<grid, 128, 0, st>>>(a);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}
