// Register-resident Stockham passes for the half-spectrum transforms of fft.cu (rfft_rows16_kernel / rfft_cols16_kernel).
//
// A transform of length n = 2^logn (256 <= n <= 4096 here) is shared by n / 16 threads; a thread holds 16 complex values
// in registers per pass and does radix-16 butterflies there (the last pass: 16 / R butterflies of radix R = 2, 4, 8 or 16
// when logn is not a multiple of four).  Per pass (Stockham autosort, kernel e^{+2 pi i jk/n} like fft_plus_rows_r4):
//     butterfly j in [0, n/R):  v[k] = x[j + k n/R] * e^{+2 pi i k (j % Ns) / (Ns R)},  V = DFT_R(v),
//                               y[(j / Ns) Ns R + (j % Ns) + m Ns] = V[m],            Ns = product of the earlier radices
// so every pass reads with unit stride across the threads, the result comes out in natural order, the first pass can take
// its input straight from global memory and the last one can hand its output to the caller from registers: a 1024-point
// transform exchanges its data through shared memory twice instead of the five read-modify-write sweeps (plus the
// bit-reversed fill and the read-out) of the in-place radix-4 routine.
//
// Everything here is __host__ __device__ so that the index arithmetic and the butterflies are tested on the CPU
// (tests/test_fft_r16_host.py compiles tests/fft_r16_host.cpp with g++ and compares with numpy.fft).
#pragma once
#ifdef __CUDACC__
#define R16_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
#else
#define R16_HD inline
#include <cmath>
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
#endif

namespace r16 {

// shared-memory slot of element i of a transform: one pad slot per 16 elements, so that the first pass's stores
// (thread j writes 16 j + m) and the unit-stride reads both spread over the bank groups
R16_HD int slot(int i) { return i + (i >> 4); }
R16_HD constexpr int rowlen(int n) { return n + (n >> 4); }

R16_HD double2 cmul(double2 a, double2 b)
{
#ifdef __CUDA_ARCH__
    return make_double2(fma(a.x, b.x, -__dmul_rn(a.y, b.y)), fma(a.x, b.y, __dmul_rn(a.y, b.x)));
#else
    return make_double2(std::fma(a.x, b.x, -(a.y * b.y)), std::fma(a.x, b.y, a.y * b.x));
#endif
}
R16_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
R16_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// e^{+2 pi i idx / 16}, idx in [0, 8): compile-time constants once the butterfly loops are unrolled
#define R16_C1 0.92387953251128674      /* cos(pi / 8) */
#define R16_S1 0.38268343236508977      /* sin(pi / 8) */
#define R16_C2 0.70710678118654752      /* cos(pi / 4) */
R16_HD constexpr double w16_re(int i)
{
    return i == 0 ? 1.0 : i == 1 ? R16_C1 : i == 2 ? R16_C2 : i == 3 ? R16_S1 : i == 4 ? 0.0 : i == 5 ? -R16_S1 : i == 6 ? -R16_C2 : -R16_C1;
}
R16_HD constexpr double w16_im(int i)
{
    return i == 0 ? 0.0 : i == 1 ? R16_S1 : i == 2 ? R16_C2 : i == 3 ? R16_C1 : i == 4 ? 1.0 : i == 5 ? R16_C1 : i == 6 ? R16_C2 : R16_S1;
}

template <int R> struct Log2;
template <> struct Log2<2> { static constexpr int v = 1; };
template <> struct Log2<4> { static constexpr int v = 2; };
template <> struct Log2<8> { static constexpr int v = 3; };
template <> struct Log2<16> { static constexpr int v = 4; };

R16_HD void swp(double2 &a, double2 &b)
{
    const double2 t = a;
    a = b;
    b = t;
}

// radix-2 stage S of the in-register transform: R / 2 butterflies, one flat loop with compile-time trip count
template <int R, int S>
R16_HD void stage(double2 *v)
{
    constexpr int half = 1 << (S - 1);
#pragma unroll
    for (int q = 0; q < R / 2; q++) {
        const int k = q & (half - 1), g = (q >> (S - 1)) << S;
        const int idx = k * (16 >> S);                               // e^{2 pi i k / 2^S} as a 16th root
        const double2 b = v[g + k + half];
        double2 t;
        if (idx == 0) t = b;
        else if (idx == 4) t = make_double2(-b.y, b.x);
        else t = cmul(make_double2(w16_re(idx), w16_im(idx)), b);
        const double2 a = v[g + k];
        v[g + k] = cadd(a, t);
        v[g + k + half] = csub(a, t);
    }
}

// V[m] = sum_k v[k] e^{+2 pi i k m / R}, in place, natural order in and out; every index is a compile-time constant
// after unrolling, so v stays in registers
template <int R>
R16_HD void dft_regs(double2 *v)
{
    constexpr int LOG = Log2<R>::v;
    // bit-reversal of the register indices, spelled out (static indices keep v in registers)
    if (R == 4) swp(v[1], v[2]);
    if (R == 8) {
        swp(v[1], v[4]);
        swp(v[3], v[6]);
    }
    if (R == 16) {
        swp(v[1], v[8]);
        swp(v[2], v[4]);
        swp(v[3], v[12]);
        swp(v[5], v[10]);
        swp(v[7], v[14]);
        swp(v[11], v[13]);
    }
    stage<R, 1>(v);
    if (LOG >= 2) stage<R, 2>(v);
    if (LOG >= 3) stage<R, 3>(v);
    if (LOG >= 4) stage<R, 4>(v);
}

// ---- one thread's share of a pass: 16 / R butterflies j = t + i T, T = n / 16 threads per transform ----
// read phase (from a padded shared-memory row)
template <int R>
R16_HD void pass_load(const double2 *x, int n, int t, double2 *v)
{
    const int T = n >> 4, nR = n / R;
#pragma unroll
    for (int i = 0; i < 16 / R; i++) {
        const int j = t + i * T;
#pragma unroll
        for (int k = 0; k < R; k++) v[i * R + k] = x[slot(j + k * nR)];
    }
}

// twiddles and butterflies; tw[q] = e^{+2 pi i q / n}, q < n / 2.  The k = 1 factor comes from the table, the others
// are its powers (binary products: at most four roundings deep).
template <int R>
R16_HD void pass_compute(const double2 *tw, int n, int Ns, int t, double2 *v)
{
    const int T = n >> 4;
#pragma unroll
    for (int i = 0; i < 16 / R; i++) {
        const int j = t + i * T;
        double2 *b = v + i * R;
        if (Ns > 1) {
            const int q = (j & (Ns - 1)) * (n / (Ns * R));           // < n / R <= n / 2
            double2 w[R];
            w[1] = tw[q];
#pragma unroll
            for (int k = 2; k < R; k++) w[k] = (k & 1) ? cmul(w[k - 1], w[1]) : cmul(w[k >> 1], w[k >> 1]);
#pragma unroll
            for (int k = 1; k < R; k++) b[k] = cmul(b[k], w[k]);
        }
        dft_regs<R>(b);
    }
}

// natural-order index of output m of butterfly j
R16_HD int out_index(int j, int m, int Ns, int R) { return (j / Ns) * Ns * R + (j & (Ns - 1)) + m * Ns; }

template <int R>
R16_HD void pass_store(double2 *y, int n, int Ns, int t, const double2 *v)
{
    const int T = n >> 4;
#pragma unroll
    for (int i = 0; i < 16 / R; i++) {
        const int j = t + i * T;
#pragma unroll
        for (int m = 0; m < R; m++) y[slot(out_index(j, m, Ns, R))] = v[i * R + m];
    }
}

// radix of the last pass: what is left of logn after the radix-16 passes (logn >= 8: two or three passes in all)
R16_HD constexpr int last_radix(int logn) { return (logn & 3) == 0 ? 16 : 1 << (logn & 3); }
R16_HD constexpr int full_passes(int logn) { return (logn & 3) == 0 ? logn / 4 - 1 : logn / 4; }      // radix-16 passes before the last

}  // namespace r16
