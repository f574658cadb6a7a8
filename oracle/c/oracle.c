/* CPU oracle, plain C (fp64).  TEST INFRASTRUCTURE ONLY: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path never links or calls it.
 *
 * Each function cites the reference lines it follows (paths under /root/reference).
 *
 * PARITY STATUS
 *   oracle_grid_*      pinned: checked against the live reference module
 *                      (oracle/_ref, built from pdspy/interferometry/libinterferometry.pyx)
 *                      and the golden vectors under tests/golden/.
 *   oracle_chisq, oracle_loglike   pinned the same way (chisq) / against the verbatim
 *                      numpy expression of pdspy/utils/emcee.py:31-43.
 *   oracle_dft         PARITY UNPINNED: the reference delegates this arithmetic to the
 *                      un-vendored, un-pinned third-party package galario
 *                      (interpolate_model.py:5,23-24); see oracle/dft.py header.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* Exact DFT in the reference's final convention (oracle/dft.py header):
 *   V_i(u,v) = sum_{j,c} image[j,c,i] * exp(+2 pi i ( u*(x_c - dRA) + v*(y_j - dDec) ))
 *   x_c = dxy*(c - nx/2),  y_j = dxy*(ny/2 - 1 - j)
 * following pdspy/interferometry/interpolate_model.py:20-27 (row flip, dxy from
 * model.x, arcsec-scaled offsets, imag -> -imag).  One sincos per (pixel, uv) pair:
 * literal, no separability, no symmetry.  image is [ny,nx,nf] (the reference's
 * [ny,nx,nf,1] layout), out_re/out_im are [nuv,nf]. */
void oracle_dft(const double *u, const double *v, int64_t nuv,
                const double *image, int ny, int nx, int nf,
                double dxy, double dRA, double dDec,
                double *out_re, double *out_im)
{
    const double twopi = 2.0 * M_PI;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t k = 0; k < nuv; k++) {
        double *accr = (double *)calloc((size_t)nf, sizeof(double));
        double *acci = (double *)calloc((size_t)nf, sizeof(double));
        for (int j = 0; j < ny; j++) {
            double yj = dxy * (double)(ny / 2 - 1 - j) - dDec;
            for (int c = 0; c < nx; c++) {
                double xc = dxy * (double)(c - nx / 2) - dRA;
                double ph = twopi * (u[k] * xc + v[k] * yj);
                double s = sin(ph), co = cos(ph);
                const double *px = image + ((size_t)j * nx + c) * nf;
                for (int i = 0; i < nf; i++) {
                    accr[i] += px[i] * co;
                    acci[i] += px[i] * s;
                }
            }
        }
        for (int i = 0; i < nf; i++) {
            out_re[k * nf + i] = accr[i];
            out_im[k * nf + i] = acci[i];
        }
        free(accr);
        free(acci);
    }
}

/* ------------------------------------------------------------------------- */
/* chisq(): pdspy/interferometry/libinterferometry.pyx:610-633.  Sum over
 * i < nuv of ((d.re-m.re)^2+(d.im-m.im)^2)*w on CHANNEL 0 ONLY ([i,0] indexing,
 * row stride nf), accumulated in double, RETURNED AS C float (:616 `cdef float`). */
float oracle_chisq(const double *dre, const double *dim, const double *w,
                   const double *mre, const double *mim, int64_t nuv, int nf)
{
    double chisq = 0;
    for (int64_t i = 0; i < nuv; i++) {
        double d1 = dre[i * nf] - mre[i * nf];
        double d2 = dim[i * nf] - mim[i * nf];
        chisq += (d1 * d1 + d2 * d2) * w[i * nf];
    }
    return (float)chisq;
}

/* Visibility term of the log-likelihood: pdspy/utils/emcee.py:31-43 (identical
 * copy pdspy/utils/dynesty.py:47-59), over all uv and all channels:
 *   -0.5*sum((d.re-m.re)^2*w) - sum(log(w[w>0]/2pi))
 *   -0.5*sum((d.im-m.im)^2*w) - sum(log(w[w>0]/2pi))
 * (the log term enters twice with a minus sign: replicated verbatim).
 * Also returns the two chi^2 sums and the log sum separately. */
double oracle_loglike(const double *dre, const double *dim, const double *w,
                      const double *mre, const double *mim, int64_t n,
                      double *chi2_re, double *chi2_im, double *logsum)
{
    double sr = 0, si = 0, sl = 0;
    const double twopi = 2.0 * M_PI;
    for (int64_t i = 0; i < n; i++) {
        double d1 = dre[i] - mre[i], d2 = dim[i] - mim[i];
        sr += d1 * d1 * w[i];
        si += d2 * d2 * w[i];
        if (w[i] > 0) sl += log(w[i] / twopi);
    }
    if (chi2_re) *chi2_re = sr;
    if (chi2_im) *chi2_im = si;
    if (logsum) *logsum = sl;
    return -0.5 * sr - sl + -0.5 * si - sl;
}

/* ------------------------------------------------------------------------- */
/* Convolution kernels: libinterferometry.pyx:543-585.  sinc is the degree-16
 * Taylor polynomial in pi*x (:547-553), exp the degree-5 Taylor polynomial
 * (:555-557), replicated term for term (they are NOT libm's).  Powers are
 * evaluated as repeated products, which is what Cython emits for `x**k` with an
 * integer literal exponent under cdef double (pow(x,k) calls); the live reference is
 * built with -ffast-math so its last bits are compiler-dependent: expsinc is
 * compared at 1e-12 relative, pillbox bit for bit. */
static size_t binned_index(uint32_t j, uint32_t i, int n, int G, int nch, int spectral)
{
    size_t q = ((size_t)j * G + i) * nch + (size_t)n;     /* what [l,m,n] addresses */
    if (spectral) return q;
    if (q >= (size_t)G * G) q = (size_t)j * G + i;
    return q;
}

static double o_sinc(double x)
{
    double xp = x * M_PI;
    return 1. - pow(xp, 2) / 6. + pow(xp, 4) / 120. - pow(xp, 6) / 5040. + pow(xp, 8) / 362880. -
           pow(xp, 10) / 39916800. + pow(xp, 12) / 6227020800. - pow(xp, 14) / 1307674368000. +
           pow(xp, 16) / 355687428096000.;
}
static double o_exp(double x)
{
    return 1 + x + pow(x, 2) / 2. + pow(x, 3) / 6. + pow(x, 4) / 24. + pow(x, 5) / 120.;
}
static double o_exp_sinc(double u, double v)
{
    double inv_alpha1 = 1. / 1.55, inv_alpha2 = 1. / 2.52, norm = 2.350016262343186;
    int m = 6;
    if (fabs(u) >= m * 0.5 || fabs(v) >= m * 0.5) return 0.;
    return o_sinc(u * inv_alpha1) * o_sinc(v * inv_alpha1) *
           o_exp(-1 * pow(u * inv_alpha2, 2)) * o_exp(-1 * pow(v * inv_alpha2, 2)) / norm;
}
static double o_ones(double u, double v)
{
    int m = 1;
    if (fabs(u) >= m * 0.5 || fabs(v) >= m * 0.5) return 0.;
    return 1.0;
}
double oracle_kernel(int conv, double u, double v) { return conv ? o_exp_sinc(u, v) : o_ones(u, v); }

/* grid(): libinterferometry.pyx:313-541, steps after the numpy preamble.
 * The caller (oracle/grid.py) does the numpy parts exactly as the reference does
 * (weights clamp/zeroing :351-353, linspace cell centres :370-381, index maps
 * :388-403, good mask :421-423) and passes the results in, because those depend on
 * numpy's own operation order.  This function does :429-533.
 *   conv: 0 pillbox (ninclude 3), 1 expsinc (ninclude 6)           :405-417
 *   weighting: 0 natural, 1 uniform, 2 superuniform, 3 robust      :429-485
 *   spectral: 0 continuum (all channels collapse to 0), 1 spectralline
 *   weights [nuv,nf] is MODIFIED in place by the re-weighting, as in the reference.
 * new_* are [G,G,nch] zero-initialised by the caller.
 * The reference indexes binned_weights[l,m,n] with the data channel n even in
 * continuum mode where the last dim is 1 (:472,:485; boundscheck is off, :313).  In the
 * contiguous [G,G,1] array that reads flat element (l*G+m)+n, i.e. the cell n columns
 * further on: replicated here (checked against the live module); past the end of the
 * array, where the reference reads foreign memory, channel 0 of the home cell is used. */
void oracle_grid_core(const double *u, const double *v, const double *freq,
                      const double *real, const double *imag, double *weights,
                      int64_t nuv, int nf, const uint32_t *ii, const uint32_t *jj,
                      const uint8_t *good, int G, double binsize,
                      const double *uu, const double *vv,
                      int conv, int weighting, double robust, int npixels,
                      int spectral, int imaging,
                      double *new_real, double *new_imag, double *new_weights)
{
    int nch = spectral ? nf : 1;
    double inv_binsize = 1. / binsize;
    double mean_freq = 0;
    for (int n = 0; n < nf; n++) mean_freq += freq[n];
    mean_freq /= nf;                       /* numpy.mean: pairwise == sequential for nf<8; see grid.py */
    double inv_freq = 1. / mean_freq;
    int ninclude = conv ? 6 : 3;
    uint32_t nmin, nmax;
    if (ninclude % 2 == 0) { nmin = (uint32_t)(ninclude * 0.5 - 1); nmax = (uint32_t)(ninclude * 0.5); }
    else { nmin = nmax = (uint32_t)((ninclude - 1) * 0.5); }
    size_t GG = (size_t)G * G;

    if (weighting > 0) {
        double *binned = (double *)malloc(GG * nch * sizeof(double));
        for (size_t q = 0; q < GG * nch; q++) binned[q] = 1.0;          /* numpy.ones :430 */
        uint32_t npix = (uint32_t)npixels;
        if (weighting == 2) npix = 3;
        for (int64_t k = 0; k < nuv; k++)
            for (int n = 0; n < nf; n++) {
                if (!good[k * nf + n]) continue;
                uint32_t j = jj[k * nf + n], i = ii[k * nf + n];
                uint32_t lmin = npix > j ? 0 : j - npix;
                uint32_t lmax = (uint32_t)((int)(j + npix + 1) < G ? (int)(j + npix + 1) : G);
                uint32_t mmin = npix > i ? 0 : i - npix;
                uint32_t mmax = (uint32_t)((int)(i + npix + 1) < G ? (int)(i + npix + 1) : G);
                int c = spectral ? n : 0;
                for (uint32_t l = lmin; l < lmax; l++)
                    for (uint32_t m = mmin; m < mmax; m++)
                        binned[((size_t)l * G + m) * nch + c] += weights[k * nf + n];
            }
        if (weighting == 1 || weighting == 2) {
            for (int64_t k = 0; k < nuv; k++)
                for (int n = 0; n < nf; n++) {
                    if (!good[k * nf + n]) continue;
                    weights[k * nf + n] /= binned[binned_index(jj[k * nf + n], ii[k * nf + n], n, G, nch, spectral)];
                }
        } else {
            /* f2 = (5*10**(-robust))**2 / ((binned**2).sum(axis=(0,1)) / weights.sum(axis=0))  :476-477
             * numpy reductions are pairwise; the caller passes sums computed by numpy
             * when bit-exactness is needed (grid.py); here plain sums (<=1e-15 rel). */
            double *f2 = (double *)malloc(nf * sizeof(double));
            for (int n = 0; n < nf; n++) {
                int c = spectral ? n : 0;
                double sb = 0, sw = 0;
                for (size_t q = 0; q < GG; q++) sb += binned[q * nch + c] * binned[q * nch + c];
                for (int64_t k = 0; k < nuv; k++) sw += weights[k * nf + n];
                double a = 5 * pow(10., -robust);
                f2[n] = a * a / (sb / sw);
            }
            for (int64_t k = 0; k < nuv; k++)
                for (int n = 0; n < nf; n++) {
                    if (!good[k * nf + n]) continue;
                    weights[k * nf + n] /=
                        (1 + f2[n] * binned[binned_index(jj[k * nf + n], ii[k * nf + n], n, G, nch, spectral)]);
                }
            free(f2);
        }
        free(binned);
    }

    /* main scatter loop :489-521 (k, then n, then l, m: this order fixes the
     * floating-point summation order of every cell) */
    for (int64_t k = 0; k < nuv; k++)
        for (int n = 0; n < nf; n++) {
            if (!good[k * nf + n]) continue;
            uint32_t j = jj[k * nf + n], i = ii[k * nf + n];
            uint32_t lmin = nmin > j ? 0 : j - nmin;
            uint32_t lmax = (uint32_t)((int)(j + nmax + 1) < G ? (int)(j + nmax + 1) : G);
            uint32_t mmin = nmin > i ? 0 : i - nmin;
            uint32_t mmax = (uint32_t)((int)(i + nmax + 1) < G ? (int)(i + nmax + 1) : G);
            int c = spectral ? n : 0;
            double us = u[k] * freq[n] * inv_freq, vs = v[k] * freq[n] * inv_freq;
            double w = weights[k * nf + n], re = real[k * nf + n], im = imag[k * nf + n];
            for (uint32_t l = lmin; l < lmax; l++)
                for (uint32_t m = mmin; m < mmax; m++) {
                    /* new_u[l,m] = uu[m], new_v[l,m] = vv[l] (numpy.meshgrid :383) */
                    double du = (us - uu[m]) * inv_binsize, dv = (vs - vv[l]) * inv_binsize;
                    double cv = conv ? o_exp_sinc(du, dv) : o_ones(du, dv);
                    size_t q = ((size_t)l * G + m) * nch + c;
                    new_real[q] += re * w * cv;
                    new_imag[q] += im * w * cv;
                    new_weights[q] += w * cv;
                }
        }

    /* normalisation :525-533.  numpy's .sum() over a strided [G,G] slice is pairwise;
     * the caller redoes the imaging normalisation with numpy when bit-exactness
     * against the live reference is asserted (grid.py, normalise=False). */
    if (imaging == 1) {
        for (int c = 0; c < nch; c++) {
            double s = 0;
            for (size_t q = 0; q < GG; q++) s += new_weights[q * nch + c];
            for (size_t q = 0; q < GG; q++) {
                new_real[q * nch + c] /= s;
                new_imag[q * nch + c] /= s;
                new_weights[q * nch + c] /= s;
            }
        }
    } else if (imaging == 0) {
        for (size_t q = 0; q < GG * nch; q++)
            if (new_weights[q] > 0) {
                new_real[q] = new_real[q] / new_weights[q];
                new_imag[q] = new_imag[q] / new_weights[q];
            }
    }   /* imaging == 2: leave raw sums, caller normalises with numpy */
}
