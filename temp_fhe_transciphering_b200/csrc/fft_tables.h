// fft_tables.h — host-side construction of the per-thread twiddle table used by fft512.cuh.
#pragma once
#include <cmath>
#include <vector>
#include "fft512.cuh"

namespace cbs {

// [64][8] t1 (thread t, k1): exp(i*pi*t/1024) * exp(-2*pi*i*t*k1/512) = exp(-2*pi*i*t*(4*k1-1)/2048)
// [8][8]  t2 (t', k2)      : exp(-2*pi*i*t'*k2/64)
// followed by the two tables of the shuffle-exchange variant (t1x, t2x)
inline std::vector<double> make_twiddle_table()
{
    std::vector<double> tab(kTwiddleTableDoubles);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int t = 0; t < 64; t++)
        for (int k1 = 0; k1 < 8; k1++) {
            long double ang = -two_pi * (long double)(t * (4 * k1 - 1)) / 2048.0L;
            tab[(t * 8 + k1) * 2] = (double)cosl(ang);
            tab[(t * 8 + k1) * 2 + 1] = (double)sinl(ang);
        }
    for (int tp = 0; tp < 8; tp++)
        for (int k2 = 0; k2 < 8; k2++) {
            long double ang = -two_pi * (long double)(tp * k2) / 64.0L;
            tab[(512 + tp * 8 + k2) * 2] = (double)cosl(ang);
            tab[(512 + tp * 8 + k2) * 2 + 1] = (double)sinl(ang);
        }
    // shuffle-exchange variant (fft512.cuh "x"): t1x = t1 * W8^((t>>3)*(t&7)); t2x[a][r] = W64^(a * ((a + r) & 7))
    double *x = tab.data() + kTwiddleXOffset;
    for (int t = 0; t < 64; t++)
        for (int k1 = 0; k1 < 8; k1++) {
            long double ang = -two_pi * (long double)(t * (4 * k1 - 1)) / 2048.0L - two_pi * (long double)(((t >> 3) * (t & 7)) & 7) / 8.0L;
            x[(t * 8 + k1) * 2] = (double)cosl(ang);
            x[(t * 8 + k1) * 2 + 1] = (double)sinl(ang);
        }
    for (int a = 0; a < 8; a++)
        for (int r = 0; r < 8; r++) {
            long double ang = -two_pi * (long double)(a * ((a + r) & 7)) / 64.0L;
            x[(512 + a * 8 + r) * 2] = (double)cosl(ang);
            x[(512 + a * 8 + r) * 2 + 1] = (double)sinl(ang);
        }
    return tab;
}

}  // namespace cbs

#include "fftw512.cuh"
namespace cbs {

// Per-lane tables of the one-warp transform (fftw512.cuh): [32 lanes][kWTabCplx] complex doubles, lane t = b + 4a:
//   C[m]    = exp(i pi 32 m / 1024) * W8^(a m)                          (twist part of register m, output rotation of pass 1)
//   T1[k]   = exp(i pi t / 1024) * W512^(t k2) * i^(b x),  k = u + 2x,  k2 = (k + 2a) mod 16
//   T2[k]   = W8^(a ka) * W32^(b ka),                      ka = (k + 2b) mod 8
inline std::vector<double> make_w_tables()
{
    std::vector<double> tab((size_t)32 * kWTabCplx * 2);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    auto put = [&](int lane, int idx, long double turns) {  // exp(2 pi i * turns)
        tab[((size_t)lane * kWTabCplx + idx) * 2] = (double)cosl(two_pi * turns);
        tab[((size_t)lane * kWTabCplx + idx) * 2 + 1] = (double)sinl(two_pi * turns);
    };
    for (int t = 0; t < 32; t++) {
        const int b = t & 3, a = t >> 2;
        for (int m = 0; m < 16; m++) put(t, kWTabC + m, (long double)(32 * m) / 2048.0L - (long double)((a * m) & 7) / 8.0L);
        for (int k = 0; k < 16; k++) {
            const int x = k >> 1, k2 = (k + 2 * a) & 15;
            put(t, kWTabT1 + k, (long double)t / 2048.0L - (long double)((t * k2) & 511) / 512.0L + (long double)((b * x) & 3) / 4.0L);
        }
        for (int k = 0; k < 8; k++) {
            const int ka = (k + 2 * b) & 7;
            put(t, kWTabT2 + k, -(long double)((a * ka) & 7) / 8.0L - (long double)((b * ka) & 31) / 32.0L);
        }
    }
    return tab;
}

}  // namespace cbs
