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
