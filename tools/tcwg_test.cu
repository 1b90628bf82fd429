// tcwg_test.cu — standalone bring-up harness of the tcgen05 weight-gradient kernel (csrc/tcwgrad.cuh) against a
// double-precision CPU reduction.  Not part of the library; built by tools/build_harness.sh.
//   tcwg_test <mode: diag | cases | time> [kx_tiles 0|1|3] [bo_mode 0|1|2] [nissue 0..3] [npass 1|3]
#include <vector>
#include <random>
#include <cmath>
#include <cstdio>
#include <string>
#include <algorithm>
#include "../s2s-ismr-unet_b200/csrc/tcwgrad.cuh"

using namespace s2s;

#define CK_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(3); } } while (0)

struct Case { const char* name; int N, H, W, Cin, Cout, Nmax; };

__global__ void reduce_slots(const float* part, float* out, int64_t P, int nslots) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
    for (int sl = 0; sl < nslots; ++sl) s += part[(int64_t)sl * P + i];
    out[i] = s;
}

static int g_kx = 0, g_bo = 0, g_ni = 0, g_np = 1;
static int run_case(const Case& c, bool diag, int time_iters) {
    std::mt19937 rng(99 + c.Cin * 7 + c.Cout + c.H);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    const size_t nx = (size_t)c.Nmax * c.H * c.W * c.Cin, nz = (size_t)c.Nmax * c.H * c.W * c.Cout;
    const int64_t P = (int64_t)9 * c.Cin * c.Cout;
    std::vector<float> x(nx), z(nz);
    for (auto& v : x) v = U(rng) * 2.f;
    for (auto& v : z) v = U(rng);
    if (diag) {     // small integers: exact in tf32, sums exact in fp32
        for (size_t i = 0; i < nx; ++i) x[i] = (float)((int)(i * 7 % 13) - 6);
        for (size_t i = 0; i < nz; ++i) z[i] = (float)((int)(i * 5 % 7) - 3);
    }
    const TcWgPlan p = tcwg_plan(c.H, c.W, c.Cin, c.Cout, c.Nmax, g_kx, 1, g_np);
    if (!p.ok) { printf("%-30s no plan\n", c.name); return 1; }
    float *d_x, *d_z, *d_part, *d_bpart, *d_dw, *d_db;
    CK_(cudaMalloc(&d_x, nx * 4)); CK_(cudaMalloc(&d_z, nz * 4));
    CK_(cudaMalloc(&d_part, (size_t)p.nslots * P * 4)); CK_(cudaMalloc(&d_bpart, (size_t)p.nslots * c.Cout * 4));
    CK_(cudaMalloc(&d_dw, P * 4)); CK_(cudaMalloc(&d_db, c.Cout * 4));
    CK_(cudaMemcpy(d_x, x.data(), nx * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemcpy(d_z, z.data(), nz * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemset(d_part, 0xff, (size_t)p.nslots * P * 4));
    CK_(cudaMemset(d_bpart, 0xff, (size_t)p.nslots * c.Cout * 4));
    CUtensorMap mx, mz;
    if (tcwg_make_maps(p, d_x, c.Cin, d_z, c.Cout, c.N, c.H, c.W, c.Cin, c.Cout, &mx, &mz) != 0) { printf("%-30s map: %s\n", c.name, last_error_ref().c_str()); return 2; }
    if (tcwg_launch(mx, mz, p, d_part, d_bpart, c.N, c.H, c.W, c.Cin, c.Cout, p.nslots, 0, g_bo, g_ni) != 0) { printf("%-30s launch: %s\n", c.name, last_error_ref().c_str()); return 2; }
    cudaError_t se = cudaDeviceSynchronize();
    if (se != cudaSuccess) { printf("%-30s kernel FAILED: %s\n", c.name, cudaGetErrorString(se)); exit(4); }
    char geo[160];
    snprintf(geo, sizeof geo, "%s nimg %d BX %d ci x%d co x%d groups %d kxt %d nst %d smem %zu slots %d tiles %d",
             p.flat ? "flat" : "tile", p.nimg, p.BX, p.ci_chunks, p.co_chunks, p.groups, p.kx_tiles, p.nstage, p.smem, p.nslots,
             tcwg_ntiles(p, c.H, c.W, c.N));
    if (time_iters > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 5; ++i) tcwg_launch(mx, mz, p, d_part, d_bpart, c.N, c.H, c.W, c.Cin, c.Cout, p.nslots, 0, g_bo, g_ni);
        cudaEventRecord(e0);
        for (int i = 0; i < time_iters; ++i) tcwg_launch(mx, mz, p, d_part, d_bpart, c.N, c.H, c.W, c.Cin, c.Cout, p.nslots, 0, g_bo, g_ni);
        cudaEventRecord(e1);
        CK_(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / time_iters;
        const double fl = 18.0 * c.Cin * c.Cout * c.N * c.H * c.W;
        printf("%-30s %8.2f us  %7.2f TFLOP/s  %7.1f GB/s  (%s)\n", c.name, us, fl / us * 1e-6,
               4.0 * c.N * c.H * c.W * (c.Cin + c.Cout) / us * 1e-3, geo);
        if ((double)c.N * c.H * c.W * c.Cin * c.Cout > (double)(1 << 27)) return 0;      // too slow to check on the CPU
    }
    reduce_slots<<<(unsigned)((P + 255) / 256), 256>>>(d_part, d_dw, P, p.nslots);
    reduce_slots<<<(c.Cout + 255) / 256, 256>>>(d_bpart, d_db, c.Cout, p.nslots);
    std::vector<float> dw(P), db(c.Cout);
    CK_(cudaMemcpy(dw.data(), d_dw, P * 4, cudaMemcpyDeviceToHost));
    CK_(cudaMemcpy(db.data(), d_db, c.Cout * 4, cudaMemcpyDeviceToHost));
    // CPU reference (images n >= N of the buffers must not contribute)
    std::vector<double> rw(P, 0.0), rb(c.Cout, 0.0);
    for (int n = 0; n < c.N; ++n)
        for (int y = 0; y < c.H; ++y)
            for (int xx = 0; xx < c.W; ++xx) {
                const float* zp = &z[(((size_t)n * c.H + y) * c.W + xx) * c.Cout];
                for (int co = 0; co < c.Cout; ++co) rb[co] += zp[co];
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) {
                        const int iy = y + ky - 1, ix = xx + kx - 1;
                        if (iy < 0 || iy >= c.H || ix < 0 || ix >= c.W) continue;
                        const float* xp = &x[(((size_t)n * c.H + iy) * c.W + ix) * c.Cin];
                        double* r = &rw[(size_t)(ky * 3 + kx) * c.Cin * c.Cout];
                        for (int ci = 0; ci < c.Cin; ++ci) {
                            const double xv = xp[ci];
                            for (int co = 0; co < c.Cout; ++co) r[(size_t)ci * c.Cout + co] += xv * zp[co];
                        }
                    }
            }
    double num = 0, den = 0, mx_d = 0, bnum = 0, bden = 0;
    int bad = 0, shown = 0;
    const double tol = diag ? 1e-6 : (g_np == 3 ? 2e-5 : 5e-3);
    double scale = 0;
    for (int64_t i = 0; i < P; ++i) scale = std::max(scale, std::fabs(rw[i]));
    for (int64_t i = 0; i < P; ++i) {
        const double d = (double)dw[i] - rw[i];
        num += d * d; den += rw[i] * rw[i]; mx_d = std::max(mx_d, std::fabs(d));
        if (!(std::fabs(d) <= tol * std::max(scale, 1.0))) {
            ++bad;
            if (shown < 10) {
                const int co = (int)(i % c.Cout), ci = (int)((i / c.Cout) % c.Cin), tap = (int)(i / ((int64_t)c.Cin * c.Cout));
                printf("    tap %d ci %d co %d: got %g want %g\n", tap, ci, co, dw[i], rw[i]);
                ++shown;
            }
        }
    }
    for (int co = 0; co < c.Cout; ++co) { const double d = (double)db[co] - rb[co]; bnum += d * d; bden += rb[co] * rb[co]; }
    const double rel = std::sqrt(num / std::max(den, 1e-300)), brel = std::sqrt(bnum / std::max(bden, 1e-300));
    const double rtol = diag ? 1e-6 : (g_np == 3 ? 5e-6 : 2e-3);
    const bool fail_ = bad > 0 || !(rel < rtol) || !(brel < rtol);
    printf("%-30s %s rel-L2 %.3e  max|d| %.3e (scale %.3g) bad %d/%lld  bias rel %.3e  (%s)\n", c.name, fail_ ? "FAIL" : "ok  ", rel, mx_d, scale, bad,
           (long long)P, brel, geo);
    cudaFree(d_x); cudaFree(d_z); cudaFree(d_part); cudaFree(d_bpart); cudaFree(d_dw); cudaFree(d_db);
    return fail_ ? 1 : 0;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "cases";
    g_kx = argc > 2 ? atoi(argv[2]) : 0;       // 0 = library default, 1 = one halo tile, 3 = one tile per kx
    g_bo = argc > 3 ? atoi(argv[3]) : 0;
    g_ni = argc > 4 ? atoi(argv[4]) : 0;       // MMA-issuing warps (0 = library default)
    g_np = argc > 5 ? atoi(argv[5]) : 1;       // passes: 1 = tf32, 3 = 3xTF32 (fp32 parity)       // descriptor base-offset rule (only matters for kx tiles = 1)
    int fails = 0;
    if (mode == "diag" || mode == "cases") {
        const bool diag = mode == "diag";
        const Case cs[] = {
            {"b2 16x16 8->8", 2, 16, 16, 8, 8, 2},
            {"b2 16x8 4->8", 2, 16, 8, 4, 8, 2},
            {"b3 64x64 8->8", 3, 64, 64, 8, 8, 3},
            {"b2 32x32 16->8", 2, 32, 32, 16, 8, 2},
            {"b2 24x24 12->24", 2, 24, 24, 12, 24, 2},
            {"b2 16x16 32->32", 2, 16, 16, 32, 32, 2},
            {"b2 8x8 64->32", 2, 8, 8, 64, 32, 2},
            {"b2 16x16 96->96", 2, 16, 16, 96, 96, 2},
            {"b1 16x16 48->192", 1, 16, 16, 48, 192, 1},
            {"b4 4x4 32->64 flat", 4, 4, 4, 32, 64, 4},
            {"b5 4x4 64->64 flat Nmax8", 5, 4, 4, 64, 64, 8},
            {"b16 2x2 96->192 flat", 16, 2, 2, 96, 192, 16},
            {"b5 2x2 32->32 flat Nmax16", 5, 2, 2, 32, 32, 16},
            {"b3 3x3 16->16 flat", 3, 3, 3, 16, 16, 3},
            {"b16 1x1 64->64 flat", 16, 1, 1, 64, 64, 16},
            {"b3 6x6 16->16 flat", 3, 6, 6, 16, 16, 3},
            {"b5 64x64 8->8 Nmax16", 5, 64, 64, 8, 8, 16},
        };
        for (const Case& c : cs) fails += run_case(c, diag, 0) != 0;
    } else {
        const int iters = 50;
        const Case ts[] = {
            {"b1 8x8 8->8", 1, 8, 8, 8, 8, 1},
            {"b16 64x64 8->8", 16, 64, 64, 8, 8, 16},
            {"b16 64x64 16->8", 16, 64, 64, 16, 8, 16},
            {"b16 32x32 16->16", 16, 32, 32, 16, 16, 16},
            {"b16 16x16 32->32", 16, 16, 16, 32, 32, 16},
            {"b16 8x8 64->64", 16, 8, 8, 64, 64, 16},
            {"b128 64x64 8->8", 128, 64, 64, 8, 8, 128},
            {"b128 64x64 16->8", 128, 64, 64, 16, 8, 128},
            {"b128 32x32 32->16", 128, 32, 32, 32, 16, 128},
            {"b128 16x16 32->32", 128, 16, 16, 32, 32, 128},
            {"b16 64x64 12->12", 16, 64, 64, 12, 12, 16},
            {"b16 64x64 24->12", 16, 64, 64, 24, 12, 16},
            {"b16 32x32 48->24", 16, 32, 32, 48, 24, 16},
            {"b16 16x16 96->48", 16, 16, 16, 96, 48, 16},
            {"b16 8x8 192->96", 16, 8, 8, 192, 96, 16},
            {"b16 4x4 192->192", 16, 4, 4, 192, 192, 16},
            {"b16 4x4 384->192", 16, 4, 4, 384, 192, 16},
            {"b16 2x2 192->384", 16, 2, 2, 192, 384, 16},
            {"b16 2x2 384->384", 16, 2, 2, 384, 384, 16},
            {"b16 64x64 96->96", 16, 64, 64, 96, 96, 16},
        };
        for (const Case& c : ts) fails += run_case(c, false, iters) != 0;
    }
    printf("tcwg_test mode %s kx_tiles %d bo_mode %d nissue %d npass %d: %d failing\n", mode.c_str(), g_kx, g_bo, g_ni, g_np, fails);
    return fails ? 1 : 0;
}
