// prep.cu — GPU pre-processing of the predictand: ISO-week rolling tercile edges and labels.
//
// Replaces the per-ISO-week xarray loops of utils/preprocessing.py:
//   rolling_labeler (preprocessing.py:53-167): for every ISO week w of the TRAINING period the tercile edges are the
//   1/3 and 2/3 quantiles over all training starts whose week lies in [w-window, w+window] (wrapping at 53), per
//   gridpoint (:112-126, xarray .quantile = numpy nanquantile, method 'linear'); the labeler then assigns
//   0 (y < e0) / 2 (y > e1) / 1, NaN where an edge is NaN (:137-158), and preprocess() one-hot encodes the labels with
//   to_categorical(., 3) (:426-428).
//
// Quantile arithmetic follows numpy bit for bit (numpy/lib/_function_base_impl.py, method 'linear'):
//   virtual index v = (n - 1) * q (double), lo = floor(v), g = v - lo,
//   d = b - a in the INPUT dtype, r = a + d*g in double, and r = b - d*(1-g) where g >= 0.5.
//
// Kernel 1 (edges): thread = (gridpoint, week); the window's values are gathered into a shared-memory column
//   [k][thread] (coalesced global reads: consecutive threads = consecutive gridpoints), NaNs dropped, insertion-sorted
//   in place, and the two interpolated order statistics written as double.  Algorithmic traffic: every start belongs
//   to 2*window+1 windows -> (2*window+1) * T * YX * sizeof(T) bytes read, 16 * nW * YX written.
// Kernel 2 (labels): element-wise over [T, YX]: one read of y and of the two edges, writes label and/or one-hot.
#include "common.cuh"

namespace s2s {

template <typename T>
__global__ void tercile_edges_kernel(const T* __restrict__ y, const int32_t* __restrict__ win_start,
                                     const int32_t* __restrict__ win_idx, int64_t YX, int nmax, double* __restrict__ edges) {
    extern __shared__ __align__(16) unsigned char te_smem[];
    T* col = reinterpret_cast<T*>(te_smem);            // [nmax][blockDim.x]
    const int w = blockIdx.y;
    const int nt = blockDim.x, tid = threadIdx.x;
    const int64_t g = (int64_t)blockIdx.x * nt + tid;
    const int s0 = win_start[w], s1 = win_start[w + 1];
    if (g >= YX) return;
    int n = 0;
    for (int k = s0; k < s1; ++k) {
        const T v = y[(int64_t)win_idx[k] * YX + g];
        if (v == v) {                                   // nanquantile: NaNs are dropped
            // insertion into the sorted prefix col[0..n)
            int j = n;
            while (j > 0 && col[(j - 1) * nt + tid] > v) { col[j * nt + tid] = col[(j - 1) * nt + tid]; --j; }
            col[j * nt + tid] = v;
            ++n;
        }
    }
    (void)nmax;
    const double qs[2] = {1.0 / 3.0, 2.0 / 3.0};
#pragma unroll
    for (int qi = 0; qi < 2; ++qi) {
        double r = __longlong_as_double(0x7ff8000000000000LL);       // all-NaN slice -> NaN edge
        if (n > 0) {
            const double q = qs[qi];
            const double v = __dmul_rn((double)(n - 1), q);     // numpy 'linear': get_virtual_index = (n - 1) * q
            double lo = floor(v);
            double gam = __dadd_rn(v, -lo);
            int ilo = (int)lo, ihi = ilo + 1;
            if (ilo < 0) { ilo = 0; ihi = 0; gam = 0.0; }
            if (ihi > n - 1) ihi = n - 1;
            if (ilo > n - 1) ilo = n - 1;
            const T a = col[ilo * nt + tid], b = col[ihi * nt + tid];
            const T d = b - a;                          // numpy subtracts in the input dtype
            r = __dadd_rn((double)a, __dmul_rn((double)d, gam));
            if (gam >= 0.5) r = __dadd_rn((double)b, -__dmul_rn((double)d, __dadd_rn(1.0, -gam)));
        }
        edges[((int64_t)w * 2 + qi) * YX + g] = r;
    }
}

template <typename T>
__global__ void tercile_label_kernel(const T* __restrict__ y, const int32_t* __restrict__ week_slot, const double* __restrict__ edges,
                                     int Tn, int64_t YX, float* __restrict__ labels, float* __restrict__ onehot) {
    const int64_t total = (int64_t)Tn * YX;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / YX);
        const int64_t g = i - (int64_t)t * YX;
        const int w = week_slot[t];
        const double e0 = edges[((int64_t)w * 2 + 0) * YX + g], e1 = edges[((int64_t)w * 2 + 1) * YX + g];
        const double v = (double)y[i];
        // xr.where(y < e0, 0, xr.where(y > e1, 2, 1)).where(~mask)   (preprocessing.py:152-155)
        float lab = v < e0 ? 0.f : (v > e1 ? 2.f : 1.f);
        const bool bad = (e0 != e0) || (e1 != e1);
        if (bad) lab = __int_as_float(0x7fc00000);
        if (labels) labels[i] = lab;
        if (onehot) {
            // to_categorical(labels, 3); a NaN label (edge missing) has no class: all-NaN row
            const float nanf_ = __int_as_float(0x7fc00000);
            onehot[3 * i + 0] = bad ? nanf_ : (lab == 0.f ? 1.f : 0.f);
            onehot[3 * i + 1] = bad ? nanf_ : (lab == 1.f ? 1.f : 0.f);
            onehot[3 * i + 2] = bad ? nanf_ : (lab == 2.f ? 1.f : 0.f);
        }
    }
}

template <typename T>
static int tercile_edges_launch(const T* y, const int32_t* ws, const int32_t* wi, int nW, int64_t YX, int nmax, double* edges,
                                cudaStream_t st) {
    int nt = 128;
    while (nt > 32 && (size_t)nt * nmax * sizeof(T) > 200 * 1024) nt >>= 1;
    const size_t smem = (size_t)nt * nmax * sizeof(T);
    S2S_REQUIRE(smem <= 200 * 1024, "tercile_edges: window of %d starts does not fit in shared memory", nmax);
    static DevOnce once;
    S2S_CUDA(once.run([] { return cudaFuncSetAttribute(tercile_edges_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
    dim3 grid((unsigned)cdiv64(YX, nt), nW);
    prof_begin(st, "tercile_edges", 0.0, 0.0);
    tercile_edges_kernel<T><<<grid, nt, smem, st>>>(y, ws, wi, YX, nmax, edges);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2s

using namespace s2s;

extern "C" {

int s2s_tercile_edges(const void* y_dev, int is_f64, const int32_t* win_start_dev, const int32_t* win_idx_dev, int n_weeks,
                      int64_t YX, int max_window_len, double* edges_dev, void* stream) {
    S2S_REQUIRE(y_dev && win_start_dev && win_idx_dev && edges_dev, "null argument");
    S2S_REQUIRE(n_weeks >= 1 && YX >= 1 && max_window_len >= 1, "bad sizes (weeks %d, points %lld, window %d)", n_weeks, (long long)YX, max_window_len);
    cudaStream_t st = (cudaStream_t)stream;
    if (is_f64) return tercile_edges_launch<double>((const double*)y_dev, win_start_dev, win_idx_dev, n_weeks, YX, max_window_len, edges_dev, st);
    return tercile_edges_launch<float>((const float*)y_dev, win_start_dev, win_idx_dev, n_weeks, YX, max_window_len, edges_dev, st);
}

int s2s_tercile_label(const void* y_dev, int is_f64, const int32_t* week_slot_dev, const double* edges_dev, int T, int64_t YX,
                      float* labels_dev, float* onehot_dev, void* stream) {
    S2S_REQUIRE(y_dev && week_slot_dev && edges_dev && (labels_dev || onehot_dev), "null argument");
    S2S_REQUIRE(T >= 1 && YX >= 1, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)T * YX;
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv64(total, 256), 148 * 16);
    prof_begin(st, "tercile_label", 0.0, 0.0);
    if (is_f64) tercile_label_kernel<double><<<grid, 256, 0, st>>>((const double*)y_dev, week_slot_dev, edges_dev, T, YX, labels_dev, onehot_dev);
    else tercile_label_kernel<float><<<grid, 256, 0, st>>>((const float*)y_dev, week_slot_dev, edges_dev, T, YX, labels_dev, onehot_dev);
    prof_end(st);
    S2S_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
