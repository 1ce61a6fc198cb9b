// fftw512.cuh — negacyclic FP64 transform for N = 1024 (512 complex points) on ONE WARP, 16 points per thread.
//
// Why a second transform (round 2): ncu on the 64-thread transform of fft512.cuh (profiles/r01b_*.csv, r02_br_v5*.csv) shows
// the blind rotation limited by three co-saturating resources - FP64 pipe, shared-memory data pipe, issue slots, all at
// 45-58 % - behind a per-group latency chain with two named barriers per transform.  With 16 points per thread a 512-point
// transform fits ONE warp (512 = 16 x 32 lanes): all three passes exchange data with warp shuffles only - no shared-memory
// transpose tile, no barrier, 104 shuffle wavefronts per polynomial instead of 184 shared-memory + shuffle wavefronts - and
// the per-thread twiddle tables (40 complex = 160 words, far too many for registers) live in tensor memory, which
// tools/bench_tmem.cu measured at 360-470 B/clk/SM beside an unaffected shared-memory pipe.
//
//   lane t = b + 4a (b = t & 3, a = t >> 2), register m = 0..15 holds the folded input point j = t + 32 m
//   pass 1  DFT-16 over m (registers)            x twiddle T1      [exchange A: 8 lanes of equal b, 2 sets of 8 registers]
//   pass 2  2 x DFT-8 over a (registers u + 2x)  x twiddle T2      [exchange B: 4 lanes of equal a, 4 sets of 4 registers]
//   pass 3  4 x DFT-4 over b (registers u + 2v + 4x')
//   result: lane (a, b), register rho = u + 2v + 4kb  <->  DFT bin 16 (v + 2b + 8 kb) + u + 2a   ("W layout")
//
// Exchanges use the static-register rotation trick of fft512.cuh: round x moves register x of lane s to lane (s + x) mod R.
// For that the producing DFT must leave its outputs rotated by the lane's own coordinate (= a modulation of its inputs,
// folded into the table of the preceding multiply: C carries W8^(a m), T1 carries i^(b x)) and the consuming DFT sees its
// inputs reversed and rotated (= the opposite-direction DFT followed by a per-lane phase, folded into T2; after the last
// pass the phase phi = W4^(b kb) is simply not applied).  So the forward transform yields conj(phi) * X in the W layout,
// key spectra are stored as TRUE values in the W layout, products carry conj(phi), and the mirrored inverse consumes
// exactly that.  tests/cpu_emul/fftw_emul.cpp executes these very phase functions on the CPU (bin map, round trip,
// negacyclic convolution against true spectra).
//
// Arithmetic: tfhe's fft64 wrapper semantics (fold N reals into N/2 complex, twist exp(i pi j / N), forward sign "-";
// SURVEY.md Appendix A), identical to fft512.cuh up to the order of the bins.
#pragma once
#include "fft512.cuh"

namespace cbs {

constexpr int kWRegs = 16;                                // complex points per thread
constexpr int kWTabC = 0, kWTabT1 = 16, kWTabT2 = 32;     // per-lane table: C[16], T1[16], T2[8] (complex)
constexpr int kWTabCplx = 40;

// 4-point DFT, natural order.  INV = false: W = exp(-2 pi i / 4)
template <bool INV>
CBS_HD void dft4(cplx &x0, cplx &x1, cplx &x2, cplx &x3)
{
    const cplx s02 = cadd(x0, x2), d02 = csub(x0, x2), s13 = cadd(x1, x3), d13 = csub(x1, x3);
    const cplx r = INV ? cplx{-d13.y, d13.x} : cplx{d13.y, -d13.x};  // d13 * (+-i)
    x0 = cadd(s02, s13);
    x2 = csub(s02, s13);
    x1 = cadd(d02, r);
    x3 = csub(d02, r);
}

// 16-point DFT, natural order in and out (radix 4 x 4).
template <bool INV>
CBS_HD void dft16(cplx v[16])
{
    // stage 1: four DFT-4 over n1 of x[n2 + 4 n1]; y[n2][q] kept at v[n2 + 4q]
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4<INV>(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
    // twiddles W16^(n2 q)
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = kSqrtHalf;
    const double sg = INV ? 1.0 : -1.0;  // sign of the imaginary part of W16^e, e in (0, 8)
    const cplx w1{c1, sg * s1}, w2{h, sg * h}, w3{s1, sg * c1}, w6{-h, sg * h}, w9{-c1, -sg * s1};
    v[1 + 4] = cmul(v[1 + 4], w1);
    v[2 + 4] = cmul(v[2 + 4], w2);
    v[3 + 4] = cmul(v[3 + 4], w3);
    v[1 + 8] = cmul(v[1 + 8], w2);
    v[2 + 8] = INV ? cplx{-v[2 + 8].y, v[2 + 8].x} : cplx{v[2 + 8].y, -v[2 + 8].x};  // W16^4 = -+i
    v[3 + 8] = cmul(v[3 + 8], w6);
    v[1 + 12] = cmul(v[1 + 12], w3);
    v[2 + 12] = cmul(v[2 + 12], w6);
    v[3 + 12] = cmul(v[3 + 12], w9);
    // stage 2: for each q a DFT-4 over n2 of y[n2][q] -> X[q + 4p]
    cplx o[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        cplx y0 = v[4 * q], y1 = v[4 * q + 1], y2 = v[4 * q + 2], y3 = v[4 * q + 3];
        dft4<INV>(y0, y1, y2, y3);
        o[q] = y0;
        o[q + 4] = y1;
        o[q + 8] = y2;
        o[q + 12] = y3;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = o[k];
}

// ---- forward: v[m] = (coef[j], coef[j + 512]) as doubles, j = lane + 32 m ---------------------------------------------
// Tab::get4(first, w) fetches the per-lane table entries [first, first + 4) (device: tcgen05.ld from tensor memory)
template <class Tab>
CBS_HD void wfwd_s1(cplx v[16], const Tab &tab)
{
    cplx w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        tab.get4(kWTabC + 4 * q, w);
#pragma unroll
        for (int k = 0; k < 4; k++) v[4 * q + k] = cmul(v[4 * q + k], w[k]);
    }
    dft16<false>(v);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        tab.get4(kWTabT1 + 4 * q, w);
#pragma unroll
        for (int k = 0; k < 4; k++) v[4 * q + k] = cmul(v[4 * q + k], w[k]);
    }
}
// [exchange A, DIR = -1]
template <class Tab>
CBS_HD void wfwd_s2(cplx v[16], const Tab &tab)
{
    cplx w[8];
    tab.get4(kWTabT2, w);
    tab.get4(kWTabT2 + 4, w + 4);
#pragma unroll
    for (int u = 0; u < 2; u++) {
        cplx e[8];
#pragma unroll
        for (int x = 0; x < 8; x++) e[x] = v[u + 2 * x];
        dft8<true>(e);
#pragma unroll
        for (int x = 0; x < 8; x++) v[u + 2 * x] = cmul(e[x], w[x]);
    }
}
// [exchange B, DIR = -1]
CBS_HD void wfwd_s3(cplx v[16])
{
#pragma unroll
    for (int uv = 0; uv < 4; uv++) dft4<true>(v[uv], v[uv + 4], v[uv + 8], v[uv + 12]);
}

// ---- inverse (unnormalised): v[rho] = conj(phi) * bin, W layout -> v[m] = (coef[j], coef[j + 512]) scaled -------------------
CBS_HD void winv_s3(cplx v[16])
{
#pragma unroll
    for (int uv = 0; uv < 4; uv++) dft4<false>(v[uv], v[uv + 4], v[uv + 8], v[uv + 12]);
}
// [exchange B, DIR = +1]
template <class Tab>
CBS_HD void winv_s2(cplx v[16], const Tab &tab)
{
    cplx w[8];
    tab.get4(kWTabT2, w);
    tab.get4(kWTabT2 + 4, w + 4);
#pragma unroll
    for (int u = 0; u < 2; u++) {
        cplx e[8];
#pragma unroll
        for (int x = 0; x < 8; x++) e[x] = cmul_conj(v[u + 2 * x], w[x]);
        dft8<false>(e);
#pragma unroll
        for (int x = 0; x < 8; x++) v[u + 2 * x] = e[x];
    }
}
// [exchange A, DIR = +1]
template <class Tab>
CBS_HD void winv_s1(cplx v[16], const Tab &tab)
{
    cplx w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        tab.get4(kWTabT1 + 4 * q, w);
#pragma unroll
        for (int k = 0; k < 4; k++) v[4 * q + k] = cmul_conj(v[4 * q + k], w[k]);
    }
    dft16<true>(v);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        tab.get4(kWTabC + 4 * q, w);
#pragma unroll
        for (int k = 0; k < 4; k++) v[4 * q + k] = cmul_conj(v[4 * q + k], w[k]);
    }
}

// exchange source lanes (the same for registers of every set): A among the 8 lanes of equal b, B among the 4 of equal a
CBS_HD int wsrc_a(int lane, int dir, int x) { return (lane & 3) | ((((lane >> 2) + dir * x) & 7) << 2); }
CBS_HD int wsrc_b(int lane, int dir, int x) { return (lane & ~3) | (((lane & 3) + dir * x) & 3); }

// W layout: DFT bin and unapplied phase exponent (phi = W4^e) of (lane, register rho)
CBS_HD int wbin(int lane, int rho) { return 16 * (((rho >> 1) & 1) + 2 * (lane & 3) + 8 * (rho >> 2)) + (rho & 1) + 2 * (lane >> 2); }
CBS_HD int wphase_exp(int lane, int rho) { return ((lane & 3) * (rho >> 2)) & 3; }

#ifdef __CUDACC__
template <int DIR>
__device__ __forceinline__ void wexchange_a(cplx v[16], int lane)
{
#pragma unroll
    for (int x = 1; x < 8; x++) {
        const int src = wsrc_a(lane, DIR, x);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            v[u + 2 * x].x = __shfl_sync(0xffffffffu, v[u + 2 * x].x, src);
            v[u + 2 * x].y = __shfl_sync(0xffffffffu, v[u + 2 * x].y, src);
        }
    }
}
template <int DIR>
__device__ __forceinline__ void wexchange_b(cplx v[16], int lane)
{
#pragma unroll
    for (int x = 1; x < 4; x++) {
        const int src = wsrc_b(lane, DIR, x);
#pragma unroll
        for (int uv = 0; uv < 4; uv++) {
            v[uv + 4 * x].x = __shfl_sync(0xffffffffu, v[uv + 4 * x].x, src);
            v[uv + 4 * x].y = __shfl_sync(0xffffffffu, v[uv + 4 * x].y, src);
        }
    }
}
#endif

}  // namespace cbs
