// Shared device helpers of the gridding code (grid.cu: ordered mode, re-weighting, average, center;
// grid_fast.cu: the fast sorted-tile mode).  Reference: pdspy/interferometry/libinterferometry.pyx:313-585.
#pragma once
#include "common.cuh"

namespace pdsb {

constexpr uint32_t KEY_DEAD = 0xffffffffu;

// ---- convolution kernels, term for term as libinterferometry.pyx:547-585 -----------------------
__device__ __forceinline__ double k_sinc(double x)
{
    const double xp = x * 3.14159265358979323846;
    const double x2 = xp * xp, x4 = x2 * x2, x6 = x4 * x2, x8 = x4 * x4, x10 = x8 * x2, x12 = x8 * x4,
                 x14 = x8 * x6, x16 = x8 * x8;
    // divisions by the literal factorials as reciprocal multiplies: the reference is built with
    // -ffast-math (setup.py:11), under which gcc does the same, so neither form is "the" bit pattern.
    return 1. - x2 * (1. / 6.) + x4 * (1. / 120.) - x6 * (1. / 5040.) + x8 * (1. / 362880.) -
           x10 * (1. / 39916800.) + x12 * (1. / 6227020800.) - x14 * (1. / 1307674368000.) +
           x16 * (1. / 355687428096000.);
}
__device__ __forceinline__ double k_exp(double x)
{
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2, x5 = x4 * x;
    return 1 + x + x2 * 0.5 + x3 * (1. / 6.) + x4 * (1. / 24.) + x5 * (1. / 120.);
}
__device__ __forceinline__ double k_exp_sinc(double u, double v)
{
    const double inv_alpha1 = 1. / 1.55, inv_alpha2 = 1. / 2.52, norm = 2.350016262343186;
    if (fabs(u) >= 3.0 || fabs(v) >= 3.0) return 0.;
    const double a = u * inv_alpha2, b = v * inv_alpha2;
    return k_sinc(u * inv_alpha1) * k_sinc(v * inv_alpha1) * k_exp(-1 * (a * a)) * k_exp(-1 * (b * b)) *
           (1. / norm);
}
// exp_sinc is separable: exp_sinc(u,v) = g(u) g(v) / norm with g(x) = sinc(x/1.55) exp(-(x/2.52)^2)
// inside |x| < 3.  The fast mode evaluates 6+6 one-dimensional factors per visibility instead of
// 36 two-dimensional values (the product order differs from the reference's at the 1e-16 level,
// which is below the fast mode's own summation-order noise).
__device__ __forceinline__ double k_exp_sinc_1d(double x)
{
    if (fabs(x) >= 3.0) return 0.;
    const double a = x * (1. / 2.52);
    return k_sinc(x * (1. / 1.55)) * k_exp(-1 * (a * a));
}
__device__ __forceinline__ double k_ones(double u, double v)
{
    if (fabs(u) >= 0.5 || fabs(v) >= 0.5) return 0.;
    return 1.0;
}

// numpy's float64 -> uint32 cast on x86-64: cvttsd2si to int64 (NaN / out of range ->
// 0x8000000000000000), then the low 32 bits.
__device__ __forceinline__ uint32_t np_f64_to_u32(double x)
{
    if (!(x > -9.2233720368547758e18 && x < 9.2233720368547758e18)) return 0u;
    return (uint32_t)(unsigned long long)__double2ll_rz(x);
}

struct GridParams {
    const double *u, *v, *freq, *re, *im, *w_in;
    double *w;              // [nuv, nf] working weights (clamped, then re-weighted)
    uint32_t *gi, *gj;      // [nuv, nf] index maps
    uint8_t *good;
    int64_t nuv;
    int nf, G, nch, spectral, conv;
    double binsize, inv_binsize, inv_freq, half;   // half = G/2. or (G-1)/2.
    const double *uu, *vv;
    uint32_t nmin, nmax;    // footprint half-widths of the main scatter
    uint32_t row_lo, row_hi;   // main scatter keeps output rows [row_lo, row_hi) only (multi-GPU row bands)
};


// grid_fast.cu: fast-mode scatter of the main sums (smode 0) or the box sums of the weights (smode 1) into
// t_re / t_im / t_w [G*G*nch].  w_src: the working weights [nuv*nf] when the prep kernel ran (re-weighting or
// map outputs requested), else nullptr (weights clamped on the fly from P.w_in as libinterferometry.pyx:351-353).
// n_outside (device counter, may be null) receives the number of visibilities off the grid.
int grid_fast_scatter(const GridParams &P, int smode, uint32_t lo, uint32_t hi, const double *w_src,
                      double *t_re, double *t_im, double *t_w, unsigned long long *n_outside);

}  // namespace pdsb
