// tc3_test.cu — standalone bring-up harness of the tcgen05 3xTF32 convolution (csrc/tc3conv.cuh) against a
// double-precision CPU convolution.  Not part of the library; built by tools/build_tc3_test.sh.
//   tc3_test <loader 0|1> <mode: diag | cases | time>
#include <vector>
#include <random>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include "../s2s-ismr-unet_b200/csrc/tc3conv.cuh"

using namespace s2s;

#define CK_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(3); } } while (0)

struct Case { const char* name; int N, H, W, Cin, Cout, ldin, coff, epi, stats, flip; };

static double elu_d(double x) { return x > 0 ? x : std::expm1(x); }

// reference: out[n,y,x,co] = epi(sum in[n,y+ky-1,x+kx-1,ci] * Weff[tap][ci][co]); Weff from W by flip rule
static void cpu_conv(const Case& c, const std::vector<float>& in, const std::vector<float>& W, const std::vector<float>& bias,
                     const std::vector<float>& aux, std::vector<double>& out) {
    // W layout: forward [tap][Cin][Cout]; flip: stored as Keras kernel of the FORWARD layer [tap][Nc][Kc] where the op contracts Kc=c.Cin
    out.assign((size_t)c.N * c.H * c.W * c.Cout, 0.0);
    for (int n = 0; n < c.N; ++n)
        for (int y = 0; y < c.H; ++y)
            for (int x = 0; x < c.W; ++x)
                for (int co = 0; co < c.Cout; ++co) {
                    double s = 0;
                    for (int ky = 0; ky < 3; ++ky)
                        for (int kx = 0; kx < 3; ++kx) {
                            const int iy = y + ky - 1, ix = x + kx - 1;
                            if (iy < 0 || iy >= c.H || ix < 0 || ix >= c.W) continue;
                            const int tap = ky * 3 + kx;
                            for (int ci = 0; ci < c.Cin; ++ci) {
                                const double a = in[(((size_t)n * c.H + iy) * c.W + ix) * c.ldin + c.coff + ci];
                                const double w = c.flip ? W[((size_t)(8 - tap) * c.Cout + co) * c.Cin + ci] : W[((size_t)tap * c.Cin + ci) * c.Cout + co];
                                s += a * w;
                            }
                        }
                    const size_t o = (((size_t)n * c.H + y) * c.W + x) * c.Cout + co;
                    if (c.epi == T3_EPI_BIAS_ACT) s = elu_d(s + bias[co]);
                    else if (c.epi == T3_EPI_ACTGRAD) { const double ya = aux[o]; s *= ya > 0 ? 1.0 : ya + 1.0; }
                    out[o] = s;
                }
}

static int run_case(const Case& c, int npass, int loader, bool diag, int time_iters) {
    std::mt19937 rng(1234 + c.Cin * 7 + c.Cout);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    const size_t nin = (size_t)c.N * c.H * c.W * c.ldin, nout = (size_t)c.N * c.H * c.W * c.Cout;
    std::vector<float> in(nin), W((size_t)9 * c.Cin * c.Cout), bias(c.Cout), aux(nout);
    for (auto& v : in) v = U(rng) * 2.f;
    const float ws = std::sqrt(6.f / (9.f * (c.Cin + c.Cout)));
    for (auto& v : W) v = U(rng) * ws * 3.f;
    for (auto& v : bias) v = U(rng) * 0.1f;
    for (auto& v : aux) v = U(rng);
    if (diag) {   // identity on the centre tap: out[.., co] = in[.., co]
        std::fill(W.begin(), W.end(), 0.f);
        for (int ch = 0; ch < std::min(c.Cin, c.Cout); ++ch) W[((size_t)4 * c.Cin + ch) * c.Cout + ch] = 1.f;
        for (size_t i = 0; i < nin; ++i) in[i] = (float)(i % 4096) / 8.f;      // exactly representable in tf32
    }
    const Tc3Plan p = tc3_plan_for(c.H, c.W, c.N, c.Cin, c.Cout, npass);
    if (!p.ok) { printf("%-28s no plan\n", c.name); return 1; }
    float *d_in, *d_W, *d_b, *d_aux, *d_out, *d_wq, *d_stat;
    Tc3WPrep* d_tab;
    const int slots = tc3_stat_slots(c.H, c.W, c.N);
    CK_(cudaMalloc(&d_in, nin * 4)); CK_(cudaMalloc(&d_W, W.size() * 4)); CK_(cudaMalloc(&d_b, c.Cout * 4));
    CK_(cudaMalloc(&d_aux, nout * 4)); CK_(cudaMalloc(&d_out, nout * 4)); CK_(cudaMalloc(&d_wq, p.wq_floats * 4));
    CK_(cudaMalloc(&d_stat, (size_t)slots * 2 * c.Cout * 4)); CK_(cudaMalloc(&d_tab, sizeof(Tc3WPrep)));
    CK_(cudaMemcpy(d_in, in.data(), nin * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemcpy(d_W, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemcpy(d_b, bias.data(), c.Cout * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemcpy(d_aux, aux.data(), nout * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemset(d_out, 0xff, nout * 4));
    Tc3WPrep e;
    e.w_off = 0; e.dst_off = 0; e.Kc = c.Cin; e.Nc = c.Cout; e.ldw_k = 0; e.NT = p.NT; e.nchunks_n = p.nchunks_n; e.CK = p.CK;
    e.kchunks = p.kchunks; e.flip = c.flip; e.npass = npass;
    CK_(cudaMemcpy(d_tab, &e, sizeof e, cudaMemcpyHostToDevice));
    tc3_wprep_kernel<<<dim3(32, 1), 256>>>(d_tab, d_W, d_wq);
    CK_(cudaGetLastError());
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    if (loader == 0 && tc3_make_map_any(d_in + c.coff, c.N, c.H, c.W, c.Cin, c.ldin, p.CK, &map) != 0) { printf("%-28s map: %s\n", c.name, last_error_ref().c_str()); return 2; }
    Tc3Args a;
    memset(&a, 0, sizeof a);
    a.wq = d_wq; a.bias = d_b; a.aux = c.epi == T3_EPI_ACTGRAD ? d_aux : nullptr; a.ldaux = c.Cout;
    a.out = d_out; a.ldout = c.Cout; a.out_coff = 0; a.stat_part = c.stats ? d_stat : nullptr;
    a.in = d_in; a.ldin = c.ldin; a.in_coff = c.coff;
    a.N = c.N; a.H = c.H; a.W = c.W; a.Cin = c.Cin; a.Cout = c.Cout; a.epi = c.epi; a.act = S2S_ACT_ELU;
    if (tc3_launch(map, a, p, npass, loader, "tc3", 0) != 0) { printf("%-28s launch: %s\n", c.name, last_error_ref().c_str()); return 2; }
    cudaError_t se = cudaDeviceSynchronize();
    if (se != cudaSuccess) { printf("%-28s npass %d loader %d: kernel FAILED: %s\n", c.name, npass, loader, cudaGetErrorString(se)); exit(4); }
    if (time_iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 5; ++i) tc3_launch(map, a, p, npass, loader, "tc3", 0);
        cudaEventRecord(e0);
        for (int i = 0; i < time_iters; ++i) tc3_launch(map, a, p, npass, loader, "tc3", 0);
        cudaEventRecord(e1);
        CK_(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / time_iters;
        const double fl = 18.0 * c.Cin * c.Cout * c.N * c.H * c.W;
        printf("%-28s npass %d loader %d: %8.2f us  %7.2f TFLOP/s  (CK %d NT %d kch %d nst %d/%d smem %zu/%zu cta/sm %d tiles %dx%dx%d)\n", c.name, npass, loader, us,
               fl / us * 1e-6, p.CK, p.NT, p.kchunks, p.nstage, p.nstage2, p.smem, p.smem2, p.ctas_per_sm2, cdiv(c.W, 8) * cdiv(c.H, 16), p.nchunks_n, c.N);
    }
    if (time_iters > 0 && (size_t)c.N * c.H * c.W * c.Cin * c.Cout > (size_t)1 << 26) return 0;   // too slow to check on the CPU
    std::vector<float> out(nout), stat((size_t)slots * 2 * c.Cout);
    CK_(cudaMemcpy(out.data(), d_out, nout * 4, cudaMemcpyDeviceToHost));
    CK_(cudaMemcpy(stat.data(), d_stat, stat.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> ref;
    cpu_conv(c, in, W, bias, aux, ref);
    double num = 0, den = 0, mx = 0;
    size_t bad = 0, first_bad = (size_t)-1;
    for (size_t i = 0; i < nout; ++i) {
        const double d = (double)out[i] - ref[i];
        num += d * d; den += ref[i] * ref[i];
        if (!(std::fabs(d) <= 1e-3 * (1 + std::fabs(ref[i])))) { ++bad; if (first_bad == (size_t)-1) first_bad = i; }
        mx = std::max(mx, std::fabs(d));
    }
    double stat_err = 0;
    if (c.stats) {
        std::vector<double> s1(c.Cout, 0.0), s2(c.Cout, 0.0), r1(c.Cout, 0.0), r2(c.Cout, 0.0);
        for (int sl = 0; sl < slots; ++sl)
            for (int ch = 0; ch < c.Cout; ++ch) { s1[ch] += stat[((size_t)sl * 2 + 0) * c.Cout + ch]; s2[ch] += stat[((size_t)sl * 2 + 1) * c.Cout + ch]; }
        for (size_t i = 0; i < nout; ++i) { r1[i % c.Cout] += ref[i]; r2[i % c.Cout] += ref[i] * ref[i]; }
        for (int ch = 0; ch < c.Cout; ++ch) {
            stat_err = std::max(stat_err, std::fabs(s1[ch] - r1[ch]) / (std::fabs(r1[ch]) + 1.0));
            stat_err = std::max(stat_err, std::fabs(s2[ch] - r2[ch]) / (std::fabs(r2[ch]) + 1.0));
        }
    }
    printf("%-28s npass %d loader %d: rel-L2 %.3e  max|d| %.3e  bad %zu/%zu  stat_err %.2e  (CK %d NT %d kch %d nst %d)\n", c.name, npass, loader,
           std::sqrt(num / std::max(den, 1e-300)), mx, bad, nout, stat_err, p.CK, p.NT, p.kchunks, p.nstage);
    if (bad && (diag || bad < nout)) {
        // decode the first mismatches: (n, y, x, co) got / want
        int shown = 0;
        for (size_t i = first_bad; i < nout && shown < 12; ++i) {
            const double d = (double)out[i] - ref[i];
            if (std::fabs(d) <= 1e-3 * (1 + std::fabs(ref[i]))) continue;
            const int co = i % c.Cout; const size_t pix = i / c.Cout;
            printf("    n %zu y %zu x %zu co %d: got %.6g want %.6g\n", pix / ((size_t)c.H * c.W), (pix / c.W) % c.H, pix % c.W, co, out[i], ref[i]);
            ++shown;
        }
    }
    cudaFree(d_in); cudaFree(d_W); cudaFree(d_b); cudaFree(d_aux); cudaFree(d_out); cudaFree(d_wq); cudaFree(d_stat); cudaFree(d_tab);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    const int loader = argc > 1 ? atoi(argv[1]) : 0;
    const std::string mode = argc > 2 ? argv[2] : "cases";
    const int iters = argc > 3 ? atoi(argv[3]) : 50;
    int fails = 0;
    if (mode == "diag") {
        const Case d1{"diag 16x8 8->8", 1, 16, 8, 8, 8, 8, 0, T3_EPI_NONE, 0, 0};
        const Case d2{"diag 32x16 16->16", 1, 32, 16, 16, 16, 16, 0, T3_EPI_NONE, 0, 0};
        fails += run_case(d1, 1, loader, true, 0);
        fails += run_case(d1, 3, loader, true, 0);
        fails += run_case(d2, 1, loader, true, 0);
        const Case r1{"rand 16x8 8->8", 1, 16, 8, 8, 8, 8, 0, T3_EPI_NONE, 0, 0};
        fails += run_case(r1, 1, loader, false, 0);
        fails += run_case(r1, 3, loader, false, 0);
    } else if (mode == "cases") {
        const Case cs[] = {
            {"64x64 8->8 fwd stats", 2, 64, 64, 8, 8, 8, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"32x32 16->16 fwd stats", 2, 32, 32, 16, 16, 16, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"32x32 8->16 fwd", 2, 32, 32, 8, 16, 8, 0, T3_EPI_BIAS_ACT, 0, 0},
            {"16x16 64->32 fwd (cat)", 3, 16, 16, 64, 32, 64, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"16x16 32->32 fwd", 3, 16, 16, 32, 32, 32, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"24x24 32->16 fwd ragged", 2, 24, 24, 32, 16, 32, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"32x32 16->8 dgrad elugrad", 2, 32, 32, 16, 8, 16, 0, T3_EPI_ACTGRAD, 0, 1},
            {"64x64 8->8 dgrad none", 2, 64, 64, 8, 8, 8, 0, T3_EPI_NONE, 0, 1},
            {"16x16 48->96 fwd nchunks", 2, 16, 16, 48, 96, 48, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"16x16 24->12 fwd", 2, 16, 16, 24, 12, 24, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"32x32 8->8 slice of ld16", 2, 32, 32, 8, 8, 16, 8, T3_EPI_BIAS_ACT, 0, 0},
            {"16x16 192->192 fwd thick", 1, 16, 16, 192, 192, 192, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 2x2 32->64 flat fwd", 16, 2, 2, 32, 64, 32, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b5 4x4 32->32 flat fwd", 5, 4, 4, 32, 32, 32, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 4x4 96->192 flat fwd", 16, 4, 4, 96, 192, 96, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b3 6x6 16->16 flat dgrad", 3, 6, 6, 16, 16, 16, 0, T3_EPI_ACTGRAD, 0, 1},
            {"b16 1x1 64->64 flat fwd", 16, 1, 1, 64, 64, 64, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b7 3x3 24->12 flat fwd", 7, 3, 3, 24, 12, 24, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 2x2 192->384 flat dgrad", 16, 2, 2, 192, 384, 192, 0, T3_EPI_NONE, 0, 1},
        };
        for (const Case& c : cs) { fails += run_case(c, 3, loader, false, 0) != 0; fails += run_case(c, 1, loader, false, 0) > 1; }
    } else {
        const Case ts[] = {
            {"b16 64x64 8->8", 16, 64, 64, 8, 8, 8, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 32x32 16->16", 16, 32, 32, 16, 16, 16, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 16x16 32->32", 16, 16, 16, 32, 32, 32, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 16x16 64->32", 16, 16, 16, 64, 32, 64, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 64x64 8->8 dgrad", 16, 64, 64, 8, 8, 8, 0, T3_EPI_ACTGRAD, 0, 1},
            {"b128 64x64 8->8", 128, 64, 64, 8, 8, 8, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b128 32x32 16->16", 128, 32, 32, 16, 16, 16, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b128 16x16 32->32", 128, 16, 16, 32, 32, 32, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b64 256x256 8->8", 64, 256, 256, 8, 8, 8, 0, T3_EPI_BIAS_ACT, 0, 0},
            {"b16 64x64 96->96", 16, 64, 64, 96, 96, 96, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 16x16 192->192", 16, 16, 16, 192, 192, 192, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 4x4 192->192 flat", 16, 4, 4, 192, 192, 192, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 2x2 192->384 flat", 16, 2, 2, 192, 384, 192, 0, T3_EPI_BIAS_ACT, 1, 0},
            {"b16 2x2 384->384 flat", 16, 2, 2, 384, 384, 384, 0, T3_EPI_BIAS_ACT, 1, 0},
        };
        for (const Case& c : ts) { run_case(c, 3, loader, false, iters); run_case(c, 1, loader, false, iters); }
    }
    printf("tc3_test loader %d mode %s kernel %s: %d failing\n", loader, mode.c_str(), tc3_use_v2() ? "v2 (persistent)" : "v1", fails);
    return fails ? 1 : 0;
}
