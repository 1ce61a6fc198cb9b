// CPU emulation of the one-warp transform of fftw512.cuh: the same phase functions the CUDA kernels call, executed for the
// 32 logical lanes with the two shuffle exchanges emulated.  Checks
//   (1) forward == conj(phi) * (naive twisted DFT) in the W layout (bin map wbin, phase wphase_exp),
//   (2) inverse(forward(x)) == 512 x,
//   (3) forward(a) .* TRUE spectrum(b) in the W layout -> inverse == negacyclic product a * b mod X^1024 + 1.
// Exit code 0 on success.  Built and run by tests/test_oracle_cpu.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fft_tables.h"
using namespace cbs;

static std::vector<double> g_tab;
struct HostTab {
    const double *p;  // this lane's table
    void get4(int first, cplx *w) const
    {
        for (int k = 0; k < 4; k++) w[k] = cplx{p[(first + k) * 2], p[(first + k) * 2 + 1]};
    }
};
static HostTab tab_of(int lane) { return HostTab{g_tab.data() + (size_t)lane * kWTabCplx * 2}; }

static void exch_a(cplx v[32][16], int dir)
{
    static cplx w[32][16];
    memcpy(w, v, sizeof(w));
    for (int t = 0; t < 32; t++)
        for (int x = 1; x < 8; x++)
            for (int u = 0; u < 2; u++) v[t][u + 2 * x] = w[wsrc_a(t, dir, x)][u + 2 * x];
}
static void exch_b(cplx v[32][16], int dir)
{
    static cplx w[32][16];
    memcpy(w, v, sizeof(w));
    for (int t = 0; t < 32; t++)
        for (int x = 1; x < 4; x++)
            for (int uv = 0; uv < 4; uv++) v[t][uv + 4 * x] = w[wsrc_b(t, dir, x)][uv + 4 * x];
}
static void forward(const std::vector<int64_t> &p, cplx v[32][16])
{
    for (int t = 0; t < 32; t++)
        for (int m = 0; m < 16; m++) v[t][m] = {(double)p[t + 32 * m], (double)p[t + 32 * m + 512]};
    for (int t = 0; t < 32; t++) wfwd_s1(v[t], tab_of(t));
    exch_a(v, -1);
    for (int t = 0; t < 32; t++) wfwd_s2(v[t], tab_of(t));
    exch_b(v, -1);
    for (int t = 0; t < 32; t++) wfwd_s3(v[t]);
}
static void inverse(cplx in[32][16], std::vector<double> &coef)
{
    static cplx v[32][16];
    memcpy(v, in, sizeof(v));
    for (int t = 0; t < 32; t++) winv_s3(v[t]);
    exch_b(v, +1);
    for (int t = 0; t < 32; t++) winv_s2(v[t], tab_of(t));
    exch_a(v, +1);
    for (int t = 0; t < 32; t++) winv_s1(v[t], tab_of(t));
    coef.assign(1024, 0.0);
    for (int t = 0; t < 32; t++)
        for (int m = 0; m < 16; m++) {
            coef[t + 32 * m] = v[t][m].x;
            coef[t + 32 * m + 512] = v[t][m].y;
        }
}
// naive twisted DFT bin k of the fold of p (long double)
static void naive_bin(const std::vector<int64_t> &p, int k, long double &re, long double &im)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    re = im = 0;
    for (int j = 0; j < 512; j++) {
        long double ang = pi * j / 1024.0L - 2.0L * pi * (long double)((j * k) % 512) / 512.0L;
        long double cr = cosl(ang), ci = sinl(ang);
        re += p[j] * cr - p[j + 512] * ci;
        im += p[j] * ci + p[j + 512] * cr;
    }
}
static cplx phase(int e)  // W4^e = exp(-2 pi i e / 4)
{
    const cplx w[4] = {{1, 0}, {0, -1}, {-1, 0}, {0, 1}};
    return w[e & 3];
}

int main()
{
    g_tab = make_w_tables();
    srand(11);
    std::vector<int64_t> a(1024), b(1024);
    for (int i = 0; i < 1024; i++) {
        a[i] = (rand() % 2001) - 1000;
        b[i] = (rand() % 2001) - 1000;
    }
    static cplx fa[32][16], fc[32][16], kb[32][16];
    forward(a, fa);
    // (1) bin map + phase, all 512 bins; also every bin is hit exactly once
    double e1 = 0;
    std::vector<int> hit(512, 0);
    for (int t = 0; t < 32; t++)
        for (int r = 0; r < 16; r++) {
            const int k = wbin(t, r);
            hit[k]++;
            long double xr, xi;
            naive_bin(a, k, xr, xi);
            // fa = conj(phi) * X
            cplx want = cmul_conj(cplx{(double)xr, (double)xi}, phase(wphase_exp(t, r)));
            e1 = fmax(e1, fmax(fabs(want.x - fa[t][r].x), fabs(want.y - fa[t][r].y)));
        }
    int bad_hit = 0;
    for (int k = 0; k < 512; k++) bad_hit += hit[k] != 1;
    // (2) round trip
    std::vector<double> back;
    inverse(fa, back);
    double e2 = 0;
    for (int i = 0; i < 1024; i++) e2 = fmax(e2, fabs(back[i] / 512.0 - (double)a[i]));
    // (3) product with the TRUE spectrum of b in the W layout
    for (int t = 0; t < 32; t++)
        for (int r = 0; r < 16; r++) {
            long double xr, xi;
            naive_bin(b, wbin(t, r), xr, xi);
            kb[t][r] = cplx{(double)xr, (double)xi};
            fc[t][r] = cmul(fa[t][r], kb[t][r]);
        }
    std::vector<double> conv;
    inverse(fc, conv);
    std::vector<double> ref(1024, 0.0);
    for (int i = 0; i < 1024; i++)
        for (int j = 0; j < 1024; j++) {
            const double p = (double)a[i] * (double)b[j];
            if (i + j >= 1024) ref[i + j - 1024] -= p;
            else ref[i + j] += p;
        }
    double e3 = 0;
    for (int i = 0; i < 1024; i++) e3 = fmax(e3, fabs(conv[i] / 512.0 - ref[i]));
    printf("w_bin_err=%.3e w_bins_bad=%d w_roundtrip_err=%.3e w_conv_err=%.3e\n", e1, bad_hit, e2, e3);
    return (e1 < 1e-6 && bad_hit == 0 && e2 < 1e-9 && e3 < 1e-3) ? 0 : 1;
}
