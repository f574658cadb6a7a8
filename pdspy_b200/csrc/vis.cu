// Dataset handles, the visibility epilogue (centre phase, Hermitian expansion, [nuv,nf] layout),
// the fused chi^2 / log-likelihood reductions (fp64, deterministic), and the C entry points of the
// image -> visibility -> likelihood path.
//
// Reference lines: pdspy/interferometry/interpolate_model.py:11-57 (output layout [nuv,nf],
// imag sign), pdspy/utils/emcee.py:31-43 (likelihood term), pdspy/interferometry/
// libinterferometry.pyx:610-633 (chisq).
#include "dft.cuh"
#include <algorithm>
#include <cstdlib>
#include <climits>

struct pdsb_dataset {
    int64_t nuv = 0, nuvh = 0;
    int hermitian = 0;
    double *u = nullptr, *v = nullptr;           // device, unique points [nuvh]
    int nf = 0;
    double *re = nullptr, *im = nullptr, *w = nullptr;   // device [nuv, nf]
    double logsum = 0.0;                         // sum log(w/2pi) over w>0
    bool has_data = false;
    int *order = nullptr;                        // device [nuvh]: the unique points along a Morton curve of the uv plane
};

// NUFFT variant of run_dft: S(u, v) of every channel as partial sums part[channel][unique uv] (defined with the NUFFT
// path at the end of this file, at global scope like the entry points)
extern "C" {
static int nufft_partials(pdsb_dataset *ds, const double *img_dev, int ny, int nx, int nf, double dxy, double2 *part);
// chi^2 per channel straight from the NUFFT sampler (no partial sums); *done = 0 when the shape is outside its domain
static int nufft_chi2_channels(pdsb_dataset *ds, const double *img_dev, int ny, int nx, int nf, double dxy, double dRA,
                               double dDec, double *chi2_dev, int *done);
}

namespace pdsb {

constexpr double kTwoPi = 6.283185307179586476925286766559;

// ---------------------------------------------------------------------------------
// Block-level fp64 sum (fixed tree: deterministic for a fixed launch geometry).
template <int THREADS>
__device__ __forceinline__ double block_sum(double x, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = x;
    __syncthreads();
    double t = 0.0;
    if (wid == 0) {
        t = lane < THREADS / 32 ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;   // valid in warp 0
}

// out[c] = sum_b part[b*ncol + c], b in order: one block per column, fixed tree.
__global__ void __launch_bounds__(256) reduce_columns_kernel(const double *__restrict__ part, int nb, int ncol,
                                                             double *__restrict__ out)
{
    __shared__ double sh[8];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nb; b += 256) s += part[(size_t)b * ncol + c];
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) out[c] = s;
}

// ---------------------------------------------------------------------------------
// Epilogue tile: 32 unique uv points x 32 channels per block, block = (32, 8).
// Loads the split partial sums coalesced over k, applies the centre phase
// G_k = exp(2 pi i (u (xcen - dRA) + v (ycen - dDec))) in fp64, transposes through shared
// memory so that the [nuv, nf] side is coalesced over channels.
struct EpiParams {
    const double2 *part;     // [nsplit][nf][nuvh]
    int nsplit, nf;
    int64_t nuvh, nuv;
    int hermitian;
    const double *u, *v;     // unique
    double xs, ys;           // xcen - dRA, ycen - dDec (radians)
    // optional model modifiers applied here instead of in host passes over the image / visibilities:
    const double *chan_scale;   // [nf] device or null: V_i *= chan_scale[i]  (flux_unc, exp(-tau_i))
    const double *plane_unscale;   // [nf] device or null: undo the tensor-core variant's operand scaling
    double ff_flux, ff_x0, ff_y0;   // point source added to the REAL part only (run_disk_model.py:329-334)
};

__device__ __forceinline__ void epi_load_tile(const EpiParams &P, double2 (*tile)[33])
{
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t kh = (int64_t)blockIdx.x * 32 + tx;
    double gs = 0.0, gc = 1.0, ff = 0.0;
    if (kh < P.nuvh) {
        double a = P.u[kh] * P.xs + P.v[kh] * P.ys;
        sincospi(2.0 * (a - rint(a)), &gs, &gc);
        // free-free term: real part of flux*exp(-2*3.14159*(0+1j*(u*x0+v*y0)))  (model.py:102-104)
        if (P.ff_flux != 0.0) ff = P.ff_flux * cos(-2 * 3.14159 * (P.u[kh] * P.ff_x0 + P.v[kh] * P.ff_y0));
    }
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int il = ty + 8 * m;
        const int i = blockIdx.y * 32 + il;
        double2 acc = make_double2(0.0, 0.0);
        if (kh < P.nuvh && i < P.nf) {
            for (int sp = 0; sp < P.nsplit; sp++) {
                const double2 p = P.part[((size_t)sp * P.nf + i) * (size_t)P.nuvh + kh];
                acc.x += p.x;
                acc.y += p.y;
            }
            acc = make_double2(acc.x * gc - acc.y * gs, acc.x * gs + acc.y * gc);
            if (P.plane_unscale) {
                const double us = P.plane_unscale[i];
                acc.x *= us;
                acc.y *= us;
            }
            if (P.chan_scale) {
                const double sc = P.chan_scale[i];
                acc.x *= sc;
                acc.y *= sc;
            }
            acc.x += ff;
        }
        tile[il][tx] = acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) finish_vis_kernel(const EpiParams P, double *__restrict__ out_re,
                                                         double *__restrict__ out_im)
{
    __shared__ double2 tile[32][33];
    epi_load_tile(P, tile);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.y * 32 + tx;
    if (i >= P.nf) return;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int kl = ty + 8 * m;
        const int64_t kh = (int64_t)blockIdx.x * 32 + kl;
        if (kh >= P.nuvh) continue;
        const double2 val = tile[tx][kl];
        out_re[kh * P.nf + i] = val.x;
        out_im[kh * P.nf + i] = val.y;
        if (P.hermitian) {
            out_re[(kh + P.nuvh) * P.nf + i] = val.x;
            out_im[(kh + P.nuvh) * P.nf + i] = -val.y;
        }
    }
}

// Fused likelihood epilogue: per block and channel, sum_k w ((d.re-m.re)^2 + (d.im-m.im)^2),
// written to blockpart[blockIdx.x][nf]; reduce_columns_kernel sums over blocks.
__global__ void __launch_bounds__(256) loglike_epi_kernel(const EpiParams P, const double *__restrict__ d_re,
                                                          const double *__restrict__ d_im,
                                                          const double *__restrict__ d_w,
                                                          double *__restrict__ blockpart)
{
    __shared__ double2 tile[32][33];
    __shared__ double red[8][33];
    epi_load_tile(P, tile);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.y * 32 + tx;
    double s = 0.0;
    if (i < P.nf) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int kl = ty + 8 * m;
            const int64_t kh = (int64_t)blockIdx.x * 32 + kl;
            if (kh >= P.nuvh) continue;
            const double2 val = tile[tx][kl];
            {
                const int64_t o = kh * P.nf + i;
                const double a = d_re[o] - val.x, b = d_im[o] - val.y;
                s += (a * a + b * b) * d_w[o];
            }
            if (P.hermitian) {
                const int64_t o = (kh + P.nuvh) * P.nf + i;
                const double a = d_re[o] - val.x, b = d_im[o] + val.y;
                s += (a * a + b * b) * d_w[o];
            }
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < P.nf) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; r++) t += red[r][tx];
        blockpart[(size_t)blockIdx.x * P.nf + i] = t;
    }
}

// ---------------------------------------------------------------------------------
// Likelihood pieces on caller-supplied model arrays (flat [n]).
// blockpart[b][0..2] = sum (d.re-m.re)^2 w, sum (d.im-m.im)^2 w, sum log(w/2pi) [w>0]
__global__ void __launch_bounds__(256) chi2_flat_kernel(const double *__restrict__ dre, const double *__restrict__ dim,
                                                        const double *__restrict__ w, const double *__restrict__ mre,
                                                        const double *__restrict__ mim, int64_t n,
                                                        double *__restrict__ blockpart)
{
    __shared__ double sh[8];
    double sr = 0.0, si = 0.0, sl = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) {
        const double ww = w[k];
        const double a = dre[k] - mre[k], b = dim[k] - mim[k];
        sr += a * a * ww;
        si += b * b * ww;
        if (ww > 0.0) sl += log(ww / kTwoPi);
    }
    sr = block_sum<256>(sr, sh);
    si = block_sum<256>(si, sh);
    sl = block_sum<256>(sl, sh);
    if (threadIdx.x == 0) {
        blockpart[(size_t)blockIdx.x * 3 + 0] = sr;
        blockpart[(size_t)blockIdx.x * 3 + 1] = si;
        blockpart[(size_t)blockIdx.x * 3 + 2] = sl;
    }
}

__global__ void __launch_bounds__(256) logsum_kernel(const double *__restrict__ w, int64_t n,
                                                     double *__restrict__ blockpart)
{
    __shared__ double sh[8];
    double sl = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) {
        const double ww = w[k];
        if (ww > 0.0) sl += log(ww / kTwoPi);
    }
    sl = block_sum<256>(sl, sh);
    if (threadIdx.x == 0) blockpart[blockIdx.x] = sl;
}

// chisq(): channel 0 of [nuv, nf]; combined sum, as chisq_calc does.
__global__ void __launch_bounds__(256) chisq_ch0_kernel(const double *__restrict__ dre, const double *__restrict__ dim,
                                                        const double *__restrict__ w, const double *__restrict__ mre,
                                                        const double *__restrict__ mim, int64_t nuv, int nf,
                                                        double *__restrict__ blockpart)
{
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < nuv; k += (int64_t)gridDim.x * 256) {
        const int64_t o = k * nf;
        const double a = dre[o] - mre[o], b = dim[o] - mim[o];
        s += (a * a + b * b) * w[o];
    }
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) blockpart[blockIdx.x] = s;
}

// ---------------------------------------------------------------------------------
// galario's algorithm, as the reference calls it (interpolate_model.py:23-27): bilinear interpolation of
// F = fftshift_rows(rfft2(fftshift(A))), A = image[::-1, :, i, 0], at (row n/2 + v/du, column |u|/du),
// du = 1/(n dxy); mirrored row and conjugate for u < 0; times exp(+2 pi i (u dRA + v dDec)); then the
// reference's imag -> -imag.  Ysh is rfft2_planes' output (fft.cu; the half spectrum, channel fastest):
// F[R][b] = conj(Ysh[(R (n/2 + 1) + b) nf + i]), b <= n/2; row n of F is the periodic copy of row 0.  One thread per (visibility, channel), channel fastest.
struct FftSampleArgs {
    const double2 *Ysh;
    const double *u, *v;
    const int *order;          // visiting order of the unique points (Morton curve of the uv plane), or null
    int64_t nuv, nuvh;
    int n, nf;
    double dxy, dRA, dDec;
};
// per-visibility part (shared by all channels): the four corners' offsets and weights, the conjugation flag, the
// phase factor
struct FftCorner {
    int64_t o00, o01, o10, o11;        // element offsets of the corners' channel 0 in Ysh
    double w00, w01, w10, w11;
    double pc, ps;
    bool uneg;
};
__device__ __forceinline__ FftCorner fft_corner(const FftSampleArgs &P, int64_t k)
{
    const int n = P.n, h = n / 2;
    const double uu = k < P.nuvh ? P.u[k] : -P.u[k - P.nuvh], vv = k < P.nuvh ? P.v[k] : -P.v[k - P.nuvh];   // Hermitian half
    const double du = 1.0 / ((double)n * P.dxy);
    FftCorner q;
    q.uneg = uu < 0.0;
    const double indu = fabs(uu) / du, indv = (double)n / 2.0 + (q.uneg ? -vv : vv) / du;
    double fu = floor(indu), fv = floor(indv);
    const double t = indu - fu, sfrac = indv - fv;
    // baselines beyond the image's Nyquist range have no cell: clamp (galario refuses them)
    fu = fmin(fmax(fu, 0.0), (double)h);
    fv = fmin(fmax(fv, 0.0), (double)n);
    const int cu0 = (int)fu, cu1 = cu0 + 1 < h ? cu0 + 1 : h;
    const int rv0 = (int)fv, rv1 = rv0 + 1 < n ? rv0 + 1 : n;
    auto off = [&](int R, int b) -> int64_t { return ((int64_t)(R % n) * (h + 1) + b) * P.nf; };
    q.o00 = off(rv0, cu0);
    q.o01 = off(rv0, cu1);
    q.o10 = off(rv1, cu0);
    q.o11 = off(rv1, cu1);
    q.w00 = (1 - t) * (1 - sfrac);
    q.w01 = t * (1 - sfrac);
    q.w10 = (1 - t) * sfrac;
    q.w11 = t * sfrac;
    sincos(kTwoPi * (uu * P.dRA + vv * P.dDec), &q.ps, &q.pc);
    return q;
}
// channel i: F = conj(Ysh), bilinear sum, conjugate for u < 0, phase shift, the reference's imag -> -imag
__device__ __forceinline__ double2 fft_sample_channel(const FftSampleArgs &P, const FftCorner &q, int i)
{
    const double2 f00 = P.Ysh[q.o00 + i], f01 = P.Ysh[q.o01 + i], f10 = P.Ysh[q.o10 + i], f11 = P.Ysh[q.o11 + i];
    const double vr = q.w00 * f00.x + q.w01 * f01.x + q.w10 * f10.x + q.w11 * f11.x;
    double vi = -(q.w00 * f00.y + q.w01 * f01.y + q.w10 * f10.y + q.w11 * f11.y);
    if (q.uneg) vi = -vi;
    return make_double2(vr * q.pc - vi * q.ps, -(vr * q.ps + vi * q.pc));
}

// A group of gs lanes (a power of two <= 32, <= nf) shares one visibility: the per-visibility part is evaluated once
// per group instruction instead of once per (visibility, channel); the lanes take the channels gs apart, so the
// corner reads and the data reads are contiguous runs of the channel-fastest arrays.
// A lane group evaluates one UNIQUE uv point; the second half of a Hermitian-doubled list gets the conjugate (galario's
// own rule for u < 0: mirrored interpolation point, conjugated value and phase - the same numbers).  The unique
// points are visited along the Morton curve of the uv plane (P.order), so that neighbouring groups read neighbouring
// corners of the transformed cube.
__global__ void __launch_bounds__(256) fft_sample_kernel(const FftSampleArgs P, int gs, double *__restrict__ out_re,
                                                         double *__restrict__ out_im)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t kk = t / gs;
    if (kk >= P.nuvh) return;
    const int64_t k = P.order ? P.order[kk] : kk;
    const FftCorner q = fft_corner(P, k);
    const bool twin = P.nuv > P.nuvh;
    for (int i = (int)(t % gs); i < P.nf; i += gs) {
        const double2 m = fft_sample_channel(P, q, i);
        out_re[k * P.nf + i] = m.x;
        out_im[k * P.nf + i] = m.y;
        if (twin) {
            out_re[(k + P.nuvh) * P.nf + i] = m.x;
            out_im[(k + P.nuvh) * P.nf + i] = -m.y;
        }
    }
}

// the same sample fused with the chi^2 sums of chi2_flat_kernel: the model visibilities never reach memory
__global__ void __launch_bounds__(256) fft_chi2_kernel(const FftSampleArgs P, int gs, const double *__restrict__ dre,
                                                       const double *__restrict__ dim, const double *__restrict__ w,
                                                       double *__restrict__ blockpart)
{
    __shared__ double sh[8];
    double sr = 0.0, si = 0.0;
    const int lg = threadIdx.x % gs;
    const bool twin = P.nuv > P.nuvh;
    const int64_t kstep = (int64_t)gridDim.x * (256 / gs);
    int64_t kk = (int64_t)blockIdx.x * (256 / gs) + threadIdx.x / gs;
    int64_t knext = kk < P.nuvh ? (P.order ? P.order[kk] : kk) : 0;      // the visiting order is read one point ahead
    for (; kk < P.nuvh; kk += kstep) {
        const int64_t k = knext;
        if (kk + kstep < P.nuvh) knext = P.order ? P.order[kk + kstep] : kk + kstep;
        const FftCorner q = fft_corner(P, k);
#pragma unroll 2
        for (int i = lg; i < P.nf; i += gs) {
            const int64_t idx = k * P.nf + i, id2 = idx + P.nuvh * P.nf;
            // the 1.5 GB of data stream through once: evict-first, so that they do not push the transformed cube
            // (the gathers' working set, about the size of the L2) out of the cache
            const double w0 = __ldcs(w + idx), a0 = __ldcs(dre + idx), b0 = __ldcs(dim + idx);
            double w1 = 0.0, a1 = 0.0, b1 = 0.0;
            if (twin) {
                w1 = __ldcs(w + id2);
                a1 = __ldcs(dre + id2);
                b1 = __ldcs(dim + id2);
            }
            const double2 m = fft_sample_channel(P, q, i);
            double a = a0 - m.x, b = b0 - m.y;
            sr += a * a * w0;
            si += b * b * w0;
            a = a1 - m.x;
            b = b1 + m.y;
            sr += a * a * w1;
            si += b * b * w1;
        }
    }
    sr = block_sum<256>(sr, sh);
    si = block_sum<256>(si, sh);
    if (threadIdx.x == 0) {
        blockpart[(size_t)blockIdx.x * 2 + 0] = sr;
        blockpart[(size_t)blockIdx.x * 2 + 1] = si;
    }
}

// ---------------------------------------------------------------------------------
// Type-2 NUFFT: the EXACT transform of interpolate_model to ~1e-7 of max|V| in O(n^2 log n + nuv w^2) operations
// instead of the direct sum's O(n^2 nuv) - an extension (code="nufft"): galario's own FFT + bilinear scheme is the
// w = 2 member of this family with a kernel that makes a 1e-3..4e-2 error; here the image is divided by the
// transform of the "exponential of semicircle" kernel (Barnett, Magland & af Klinteberg 2019)
//     psi(t) = exp(beta (sqrt(1 - (2t/w)^2) - 1)),  |t| <= w/2,   beta = 2.30 w,
// zero-padded to N = 2n (rfft2_planes_padded: the pixel grid's frequencies a = u dxy, b = v dxy live on the unit
// torus, the padded transform samples it at k / N), and a visibility is the w x w sum
//     S(a, b) = sum_{tx, ty} psi(aN - kx) psi(bN - ky) G[ky][kx],   V = S exp(-2 pi i (u dRA + v dDec)),
// G[ky][kx] = sum_{j,c} I'[j,c] exp(+2 pi i (ky Y_j + kx X_c) / N), X_c = c - n/2, Y_j = n/2 - 1 - j (the reference's
// flip and centre).  G is held as the half spectrum kx in [0, N/2] (real image: G[-ky][-kx] = conj G[ky][kx]) and is
// periodic in both indices.  w = 8: 4e-8 of max|V| against the exact oracle, chi^2 to 2e-10 (tests/test_gpu_nufft.py).
constexpr int NUFFT_W = 8;
constexpr double NUFFT_BETA = 2.30 * NUFFT_W;

struct NufftArgs {
    const double2 *Yh;         // rfft2_planes_padded output: [N][N/2 + 1][nf]
    const double *u, *v;       // unique half of the uv list
    const int *order;          // visiting order of the unique points (Morton curve: neighbours share taps in L2)
    int64_t nuv, nuvh;
    int N, nf;
    double dxy, dRA, dDec;
};

__device__ __forceinline__ double nufft_psi(double t)
{
    const double z = t * (2.0 / NUFFT_W), q = 1.0 - z * z;
    return q > 0.0 ? exp(NUFFT_BETA * (sqrt(q) - 1.0)) : (q == 0.0 ? exp(-NUFFT_BETA) : 0.0);
}

// per-visibility part: tap offsets / weights along both axes, conjugation of the taps that fall on the mirrored half
struct NufftPoint {
    int64_t rowp[NUFFT_W], rown[NUFFT_W];      // element offset of stored row ky / -ky (column 0, channel 0)
    int colo[NUFFT_W];                         // element offset of column kx (or its mirror) inside a row
    double wx[NUFFT_W], wxs[NUFFT_W], wy[NUFFT_W];   // wxs: wx with the sign of the conjugation for the imaginary part
    unsigned conj;                             // bit tx: the tap reads conj G[-ky][-kx]
    double pc, ps;
};
__device__ __forceinline__ NufftPoint nufft_point(const NufftArgs &P, int64_t k)
{
    const int N = P.N, h = N / 2;
    const double uu = P.u[k], vv = P.v[k];
    const double a = uu * P.dxy * (double)N, b = vv * P.dxy * (double)N;
    const double kx0 = ceil(a - 0.5 * NUFFT_W), ky0 = ceil(b - 0.5 * NUFFT_W);
    NufftPoint q;
    q.conj = 0u;
#pragma unroll
    for (int t = 0; t < NUFFT_W; t++) {
        q.wx[t] = nufft_psi(a - (kx0 + t));
        q.wy[t] = nufft_psi(b - (ky0 + t));
        // N is a power of two: mod N of a (possibly negative) index is a mask
        int kxm = ((int)kx0 + t) & (N - 1);
        const int kyp = ((int)ky0 + t) & (N - 1), kyn = (N - kyp) & (N - 1);
        const bool cj = kxm > h;
        if (cj) {
            kxm = N - kxm;
            q.conj |= 1u << t;
        }
        q.wxs[t] = cj ? -q.wx[t] : q.wx[t];
        q.colo[t] = kxm * P.nf;
        q.rowp[t] = (int64_t)((kyp + h) & (N - 1)) * (h + 1) * P.nf;
        q.rown[t] = (int64_t)((kyn + h) & (N - 1)) * (h + 1) * P.nf;
    }
    sincos(kTwoPi * (uu * P.dRA + vv * P.dDec), &q.ps, &q.pc);
    return q;
}
// channel i of one unique uv point: V = S exp(-i phi)
__device__ __forceinline__ double2 nufft_channel(const NufftArgs &P, const NufftPoint &q, int i)
{
    double sr = 0.0, si = 0.0;
#pragma unroll
    for (int ty = 0; ty < NUFFT_W; ty++) {
        const double2 *rp = P.Yh + q.rowp[ty] + i, *rn = P.Yh + q.rown[ty] + i;
        double tr = 0.0, ti = 0.0;
#pragma unroll
        for (int tx = 0; tx < NUFFT_W; tx++) {
            const double2 y = ((q.conj >> tx) & 1u ? rn : rp)[q.colo[tx]];
            tr = fma(q.wx[tx], y.x, tr);
            ti = fma(q.wxs[tx], y.y, ti);
        }
        sr = fma(q.wy[ty], tr, sr);
        si = fma(q.wy[ty], ti, si);
    }
    return make_double2(sr * q.pc + si * q.ps, si * q.pc - sr * q.ps);
}

// lane groups as in fft_sample_kernel; a group evaluates one UNIQUE uv point and writes both Hermitian halves
__global__ void __launch_bounds__(256) nufft_sample_kernel(const NufftArgs P, int gs, double *__restrict__ out_re,
                                                           double *__restrict__ out_im, double2 *__restrict__ part)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t kk = t / gs;
    if (kk >= P.nuvh) return;
    const int64_t k = P.order ? P.order[kk] : kk;
    const NufftPoint q = nufft_point(P, k);
    const bool twin = P.nuv > P.nuvh;
    for (int i = (int)(t % gs); i < P.nf; i += gs) {
        const double2 m = nufft_channel(P, q, i);
        if (part) {                                      // S without the phase: V = m (pc + i ps)
            part[(size_t)i * P.nuvh + k] = make_double2(m.x * q.pc - m.y * q.ps, m.y * q.pc + m.x * q.ps);
            continue;
        }
        out_re[k * P.nf + i] = m.x;
        out_im[k * P.nf + i] = m.y;
        if (twin) {
            out_re[(k + P.nuvh) * P.nf + i] = m.x;
            out_im[(k + P.nuvh) * P.nf + i] = -m.y;
        }
    }
}

__global__ void __launch_bounds__(256) nufft_chi2_kernel(const NufftArgs P, int gs, const double *__restrict__ dre,
                                                         const double *__restrict__ dim, const double *__restrict__ w,
                                                         double *__restrict__ blockpart)
{
    __shared__ double sh[8];
    double sr = 0.0, si = 0.0;
    const int lg = threadIdx.x % gs;
    const bool twin = P.nuv > P.nuvh;
    const int64_t kstep = (int64_t)gridDim.x * (256 / gs);
    for (int64_t kk = (int64_t)blockIdx.x * (256 / gs) + threadIdx.x / gs; kk < P.nuvh; kk += kstep) {
        const int64_t k = P.order ? P.order[kk] : kk;
        const NufftPoint q = nufft_point(P, k);
        for (int i = lg; i < P.nf; i += gs) {
            const int64_t idx = k * P.nf + i;
            const double w0 = __ldcs(w + idx), a0 = __ldcs(dre + idx), b0 = __ldcs(dim + idx);
            double w1 = 0.0, a1 = 0.0, b1 = 0.0;
            if (twin) {
                const int64_t id2 = idx + P.nuvh * P.nf;
                w1 = __ldcs(w + id2);
                a1 = __ldcs(dre + id2);
                b1 = __ldcs(dim + id2);
            }
            const double2 m = nufft_channel(P, q, i);
            double a = a0 - m.x, b = b0 - m.y;
            sr += a * a * w0;
            si += b * b * w0;
            a = a1 - m.x;
            b = b1 + m.y;                               // the twin's model is the conjugate
            sr += a * a * w1;
            si += b * b * w1;
        }
    }
    sr = block_sum<256>(sr, sh);
    si = block_sum<256>(si, sh);
    if (threadIdx.x == 0) {
        blockpart[(size_t)blockIdx.x * 2 + 0] = sr;
        blockpart[(size_t)blockIdx.x * 2 + 1] = si;
    }
}

// The same sums with the spectrum patch of a batch of uv points staged in shared memory.  The unique points arrive
// along the Morton curve, so NT_PTS consecutive ones sit in a small box of the oversampled grid; the box (+ 8 taps) is
// loaded once per group of NT_CG channels - wrap around the torus and the mirror of the half spectrum resolved while
// staging - and the 64 taps of every (point, channel) are then 16-byte shared-memory reads (a half-warp = one point, its
// lanes = the channels: 256 contiguous bytes, no bank conflicts).  Batches whose box exceeds NT_CAP cells (the sparse
// outskirts of the uv plane) take the direct path of nufft_chi2_kernel.  L2 traffic per C3 likelihood: 33 GB -> ~6 GB.
#ifndef NT_PTS_N
#define NT_PTS_N 32
#endif
#ifndef NT_CAP_N
#define NT_CAP_N 400
#endif
#ifndef NT_MINB
#define NT_MINB 2
#endif
constexpr int NT_PTS = NT_PTS_N;
constexpr int NT_CG = 16;
constexpr int NT_CAP = NT_CAP_N;
struct NtPoint {
    double wx[NUFFT_W], wy[NUFFT_W];
    double pc, ps;
    int kx0, ky0;
    int64_t k;                                   // index of the unique point (-1: past the end of the list)
};
constexpr size_t NT_SMEM = (size_t)NT_CAP * NT_CG * sizeof(double2) + 2 * NT_PTS * sizeof(NtPoint) + 64;      // two (pts, box) slots

// MODE 0: the two chi^2 sums of pdsb_loglike_nufft (real and imaginary part, all channels together).
// MODE 1: instead of chi^2, write S (without the dRA / dDec phase) as the partial sums part[channel][unique uv] that
//         the epilogues shared with the direct-sum kernels consume (run_dft with the NUFFT variant).
// MODE 2: chi^2 per channel (real + imaginary), what pdsb_loglike* return: blockpart[block][channel] (nf <= 16 NT_MAXG).
constexpr int NT_MAXG = 8;
template <int MODE>
__global__ void __launch_bounds__(256, NT_MINB) nufft_chi2_tiled_kernel(const NufftArgs P, const double *__restrict__ dre,
                                                                  const double *__restrict__ dim,
                                                                  const double *__restrict__ w,
                                                                  double *__restrict__ blockpart,
                                                                  double2 *__restrict__ part)
{
    extern __shared__ __align__(16) unsigned char nt_smem[];
    double2 *patch = reinterpret_cast<double2 *>(nt_smem);
    NtPoint *pts = reinterpret_cast<NtPoint *>(nt_smem + (size_t)NT_CAP * NT_CG * sizeof(double2));
    int *box = reinterpret_cast<int *>(pts + 2 * NT_PTS);        // per slot: min kx0, max kx0, min ky0, max ky0
    __shared__ double sh[8];
    const int tid = threadIdx.x;
    const int N = P.N, h = N / 2;
    constexpr bool PART = MODE == 1;
    const bool twin = P.nuv > P.nuvh;
    const int64_t nbatch = (P.nuvh + NT_PTS - 1) / NT_PTS;
    double sr = 0.0, si = 0.0;
    double accg[NT_MAXG];                        // MODE 2: this thread's channels tid % 16 + 16 g
#pragma unroll
    for (int g = 0; g < NT_MAXG; g++) accg[g] = 0.0;
    // ---- per-point part of batch bb into a (pts, box) slot: thread = (point, tap); the box must have been reset ----
    auto points = [&](int64_t bb, NtPoint *pp, int *bx) {
        for (int p = tid >> 3; p < NT_PTS; p += 32) {
            const int t = tid & 7;
            const int64_t kk = bb * NT_PTS + p;
            NtPoint &q = pp[p];
            if (kk < P.nuvh) {
                const int64_t k = P.order ? P.order[kk] : kk;
                const double uu = P.u[k], vv = P.v[k];
                const double a = uu * P.dxy * (double)N, bb2 = vv * P.dxy * (double)N;
                const double kx0 = ceil(a - 0.5 * NUFFT_W), ky0 = ceil(bb2 - 0.5 * NUFFT_W);
                // wx carries the sign of the conjugation: a tap on the mirrored half (kx mod N > N/2) reads conj G[-ky][-kx];
                // the staged patch holds G unconjugated, the weight is psi >= 0, so |wx| serves the real part
                const double px = nufft_psi(a - (kx0 + t));
                q.wx[t] = ((((int)kx0 + t) & (N - 1)) > h) ? -px : px;
                q.wy[t] = nufft_psi(bb2 - (ky0 + t));
                if (t == 0) {
                    q.kx0 = (int)kx0;
                    q.ky0 = (int)ky0;
                    q.k = k;
                    sincos(kTwoPi * (uu * P.dRA + vv * P.dDec), &q.ps, &q.pc);
                    atomicMin(&bx[0], q.kx0);
                    atomicMax(&bx[1], q.kx0);
                    atomicMin(&bx[2], q.ky0);
                    atomicMax(&bx[3], q.ky0);
                }
            } else if (t == 0) {
                q.k = -1;
            }
        }
    };
    // ---- stage the box (minx, miny, bw x bh cells) of one channel group: thread = (cell, channel), channel fastest; wrap
    // around the torus and the mirror of the half spectrum (row -ky, column N - kx) resolved in the source address, the
    // conjugation of mirrored cells left to the weights' sign; cp.async: no registers involved, one commit group ----
    auto stage = [&](int minx, int miny, int bw, int bh, int cg0, double2 *dst) {
        const int ncell16 = bw * bh * NT_CG;
        const float inv_bw = 1.0f / (float)bw;
        for (int idx = tid; idx < ncell16; idx += 256) {
            const int cell = idx / NT_CG, ch = idx % NT_CG;
            if (cg0 + ch >= P.nf) continue;                          // never read
            const int cy = (int)(((float)cell + 0.5f) * inv_bw), cx = cell - cy * bw;              // cell < 400: exact
            int kxm = (minx + cx) & (N - 1);                         // N is a power of two
            int ky = miny + cy;
            if (kxm > h) {
                kxm = N - kxm;
                ky = -ky;
            }
            const int row = ((ky & (N - 1)) + h) & (N - 1);
            const double2 *src = P.Yh + ((int64_t)row * (h + 1) + kxm) * P.nf + cg0 + ch;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + idx);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double2 *const buf1 = patch + (size_t)(NT_CAP / 2) * NT_CG;
    // Two (pts, box) slots: the per-point part of the NEXT batch of this block runs during the first channel group of the
    // current one, and when both boxes fit half the patch buffer the next batch's first patch is requested before the
    // current batch's last group is summed - the sampler then never waits for a patch except at its very first batch.
    int slot = 0, par = 0;                               // par: which half holds channel group 0 of the current batch
    bool prefetched = false;
    if (blockIdx.x < nbatch) {
        if (tid < 4) box[tid] = (tid & 1) ? INT_MIN : INT_MAX;
        __syncthreads();
        points(blockIdx.x, pts, box);
        __syncthreads();
    }
    for (int64_t b = blockIdx.x; b < nbatch; b += gridDim.x, slot ^= 1) {
        const NtPoint *pcur = pts + slot * NT_PTS;
        int *bnext = box + (slot ^ 1) * 4;
        const int64_t bn = b + gridDim.x;
        const bool has_next = bn < nbatch;
        const int minx = box[slot * 4 + 0], miny = box[slot * 4 + 2];
        const int bw = box[slot * 4 + 1] - minx + NUFFT_W, bh = box[slot * 4 + 3] - miny + NUFFT_W;
        const bool staged = (int64_t)bw * bh <= NT_CAP;
        // A box of at most NT_CAP / 2 cells (three batches in four on C3) leaves room for two patches: the next channel
        // group's patch is then on its way while the current one is summed.
        const bool dbl = staged && 2 * bw * bh <= NT_CAP;
        if (tid < 4) bnext[tid] = (tid & 1) ? INT_MIN : INT_MAX;     // (its old content was consumed a batch ago)
        if (staged && !prefetched) stage(minx, miny, bw, bh, 0, dbl && (par & 1) ? buf1 : patch);
        bool prefetched_next = false;
        const int G = (P.nf + NT_CG - 1) / NT_CG;
        for (int cg0 = 0, g = 0; cg0 < P.nf; cg0 += NT_CG, g++) {
            const double2 *cur = dbl && ((par + g) & 1) ? buf1 : patch;
            if (staged) {
                double2 *const other = ((par + g) & 1) ? patch : buf1;       // free since the barrier that closed group g - 1
                bool more = false;
                if (dbl && g + 1 < G) {
                    stage(minx, miny, bw, bh, cg0 + NT_CG, other);
                    more = true;
                } else if (dbl && has_next && g >= 1) {
                    // last group: the next batch's box is known since the barrier that closed group 0
                    const int nminx = bnext[0], nminy = bnext[2];
                    const int nbw = bnext[1] - nminx + NUFFT_W, nbh = bnext[3] - nminy + NUFFT_W;
                    if (2 * (int64_t)nbw * nbh <= NT_CAP) {
                        stage(nminx, nminy, nbw, nbh, 0, other);
                        more = prefetched_next = true;
                    }
                }
                if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            if (staged || g == 0) __syncthreads();       // patch of group g complete; the reset of the next box visible
            if (g == 0 && has_next) points(bn, pts + (slot ^ 1) * NT_PTS, bnext);
            // ---- (point, channel) pairs: half-warp = point, lane = channel ----
#pragma unroll 1
            for (int it = 0; it < NT_PTS * NT_CG / 256; it++) {
                const int p = it * (256 / NT_CG) + tid / NT_CG, ch = tid % NT_CG, i = cg0 + ch;
                const NtPoint &q = pcur[p];
                if (q.k < 0 || i >= P.nf) continue;
                // the data of this (point, channel) and of its Hermitian twin: on their way while the taps are summed
                const int64_t idx = q.k * P.nf + i, id2 = idx + P.nuvh * P.nf;
                double w0 = 0.0, a0 = 0.0, b0 = 0.0, w1 = 0.0, a1 = 0.0, b1 = 0.0;
                if (!PART) {
                    w0 = __ldcs(w + idx);
                    a0 = __ldcs(dre + idx);
                    b0 = __ldcs(dim + idx);
                }
                if (!PART && twin) {
                    w1 = __ldcs(w + id2);
                    a1 = __ldcs(dre + id2);
                    b1 = __ldcs(dim + id2);
                }
                double2 m;
                if (staged) {
                    double wx[NUFFT_W];
#pragma unroll
                    for (int t = 0; t < NUFFT_W; t++) wx[t] = q.wx[t];
                    const double2 *base = cur + ((size_t)(q.ky0 - miny) * bw + (q.kx0 - minx)) * NT_CG + ch;
                    double s_r = 0.0, s_i = 0.0;
#pragma unroll
                    for (int ty = 0; ty < NUFFT_W; ty++) {
                        const double2 *rowp = base + (size_t)ty * bw * NT_CG;
                        double tr = 0.0, ti = 0.0;
#pragma unroll
                        for (int tx = 0; tx < NUFFT_W; tx++) {
                            const double2 y = rowp[tx * NT_CG];
                            tr = fma(fabs(wx[tx]), y.x, tr);
                            ti = fma(wx[tx], y.y, ti);
                        }
                        const double wyv = q.wy[ty];
                        s_r = fma(wyv, tr, s_r);
                        s_i = fma(wyv, ti, s_i);
                    }
                    m = PART ? make_double2(s_r, s_i) : make_double2(s_r * q.pc + s_i * q.ps, s_i * q.pc - s_r * q.ps);
                } else {
                    // the box of this batch does not fit: the same 64 taps straight from the spectrum (weights from pts)
                    double s_r = 0.0, s_i = 0.0;
#pragma unroll 1
                    for (int ty = 0; ty < NUFFT_W; ty++) {
                        const int kyp = (q.ky0 + ty) & (N - 1), kyn = (N - kyp) & (N - 1);
                        const double2 *rp = P.Yh + (int64_t)((kyp + h) & (N - 1)) * (h + 1) * P.nf + i;
                        const double2 *rn = P.Yh + (int64_t)((kyn + h) & (N - 1)) * (h + 1) * P.nf + i;
                        double tr = 0.0, ti = 0.0;
#pragma unroll
                        for (int tx = 0; tx < NUFFT_W; tx++) {
                            int kxm = (q.kx0 + tx) & (N - 1);
                            const bool cj = kxm > h;
                            if (cj) kxm = N - kxm;
                            const double2 y = (cj ? rn : rp)[(int64_t)kxm * P.nf];
                            const double wv = q.wx[tx];              // signed like cj
                            tr = fma(fabs(wv), y.x, tr);
                            ti = fma(wv, y.y, ti);
                        }
                        const double wyv = q.wy[ty];
                        s_r = fma(wyv, tr, s_r);
                        s_i = fma(wyv, ti, s_i);
                    }
                    m = PART ? make_double2(s_r, s_i) : make_double2(s_r * q.pc + s_i * q.ps, s_i * q.pc - s_r * q.ps);
                }
                if (PART) {
                    part[(size_t)i * P.nuvh + q.k] = m;
                    continue;
                }
                double a = a0 - m.x, bq = b0 - m.y;
                double t_re = a * a * w0, t_im = bq * bq * w0;
                a = a1 - m.x;
                bq = b1 + m.y;                           // the twin's model is the conjugate (w1 = 0 without a twin)
                t_re += a * a * w1;
                t_im += bq * bq * w1;
                if (MODE == 2) accg[cg0 / NT_CG] += t_re + t_im;
                else {
                    sr += t_re;
                    si += t_im;
                }
            }
            __syncthreads();                             // the patch of this group is free again
            if (staged && !dbl && cg0 + NT_CG < P.nf) stage(minx, miny, bw, bh, cg0 + NT_CG, patch);
        }
        par = (par + G) & 1;
        prefetched = prefetched_next;
    }
    if (PART) return;
    if (MODE == 2) {
        // per channel: the 16 threads that share tid % 16 hold the same channels
        double *red = reinterpret_cast<double *>(nt_smem);           // the patch is free now
        for (int g = 0; g * NT_CG < P.nf; g++) {
            __syncthreads();
            red[tid] = accg[g];
            __syncthreads();
            if (tid < NT_CG && g * NT_CG + tid < P.nf) {
                double t = 0.0;
                for (int j = 0; j < 256 / NT_CG; j++) t += red[j * NT_CG + tid];
                blockpart[(size_t)blockIdx.x * P.nf + g * NT_CG + tid] = t;
            }
        }
        return;
    }
    sr = block_sum<256>(sr, sh);
    si = block_sum<256>(si, sh);
    if (threadIdx.x == 0) {
        blockpart[(size_t)blockIdx.x * 2 + 0] = sr;
        blockpart[(size_t)blockIdx.x * 2 + 1] = si;
    }
}

static int reduce_blocks(const double *blockpart, int nb, int ncol, double *out_dev)
{
    LaunchScope ls("reduce_columns");
    reduce_columns_kernel<<<ncol, 256, 0, ctx().stream>>>(blockpart, nb, ncol, out_dev);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

// ---------------------------------------------------------------------------------
// image (host or device, fp64 [ny,nx,nf]) -> split partial sums on the device.
struct DftRun {
    DftGeom g;
    int nsplit;
    const double *plane_unscale = nullptr;   // tensor-core kernel: V_i *= plane_unscale[i]
};

static int run_dft(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                   DftRun *run)
{
    Context &c = ctx();
    PDSB_REQUIRE(ds && image, "dataset/image");
    PDSB_REQUIRE(ny > 0 && nx > 0 && nf > 0, "image shape");
    PDSB_REQUIRE(dxy > 0.0, "dxy");
    const double *img_dev = nullptr;
    PDSB_CHECK(to_device(image, image_kind, (size_t)ny * nx * nf * sizeof(double), c.img64, (const void **)&img_dev));
    // (odd or oversized images are outside the NUFFT path's domain: they take the FP32 direct sum below)
    if (c.dft_variant == DFT_VARIANT_NUFFT && ny % 2 == 0 && nx % 2 == 0 && ny <= 2048 && nx <= 2048) {
        // the exact transform through the 8-point non-uniform FFT; sums relative to pixel (ny/2, nx/2): no centre phase
        PDSB_CHECK(c.partial.ensure((size_t)nf * ds->nuvh * sizeof(double2)));
        PDSB_CHECK(nufft_partials(ds, img_dev, ny, nx, nf, dxy, c.partial.as<double2>()));
        run->g = make_geom(ny, nx, nf, 32, dxy);
        run->g.xcen = run->g.ycen = 0.0;
        run->nsplit = 1;
        run->plane_unscale = nullptr;
        return PDSB_OK;
    }
    if (c.dft_variant == DFT_VARIANT_F64) {
        // fp64 throughout, straight from the fp64 cube (dft_f64.cu)
        PDSB_CHECK(c.partial.ensure((size_t)nf * ds->nuvh * sizeof(double2)));
        PDSB_CHECK(launch_dft_f64(img_dev, ny, nx, nf, ds->u, ds->v, ds->nuvh, dxy, c.partial.as<double2>()));
        run->g = make_geom(ny, nx, nf, 32, dxy);
        run->nsplit = 1;
        run->plane_unscale = nullptr;
        return PDSB_OK;
    }
    if (c.dft_variant >= DFT_VARIANT_TC5 && c.dft_variant < DFT_VARIANT_F64) {
        // tensor-core kernel: lattice-split fp16 operands on tcgen05 (dft_tc5.cu)
        DftGeom g = make_geom(ny, nx, nf, 32, dxy);
        PDSB_CHECK(c.folded.ensure(tc5_operand_bytes(ny, nx, nf)));
        PDSB_CHECK(c.mma_ws.ensure(tc5_ws_bytes(ny, nx, nf)));
        PDSB_CHECK(launch_fold_tc5(img_dev, c.folded.as<unsigned char>(), c.mma_ws.ptr, ny, nx, nf));
        const int nsplit = tc5_auto_split(ds->nuvh, nf, nx);
        PDSB_CHECK(c.partial.ensure((size_t)nsplit * nf * ds->nuvh * sizeof(double2)));
        DftParams p;
        p.F = nullptr;
        p.u = ds->u;
        p.v = ds->v;
        p.nuvh = ds->nuvh;
        p.dxy = dxy;
        p.ntile = 0;
        p.nchunk = g.nchunk;
        p.nsplit = nsplit;
        p.nf = nf;
        p.hx = g.hx2 ? 0.5 : 0.0;
        p.hy = g.hy2 ? 0.5 : 0.0;
        p.part = c.partial.as<double2>();
        PDSB_CHECK(launch_dft_tc5(p, c.folded.as<unsigned char>(), c.mma_ws.ptr, ny, nx));
        run->g = g;
        run->nsplit = nsplit;
        run->plane_unscale = tc5_plane_unscale(c.mma_ws.ptr, ny, nx, nf);
        return PDSB_OK;
    }
    const int variant = dft_pick_variant(nx);
    const int tcp = dft_variant_tcp(variant);
    DftGeom g = make_geom(ny, nx, nf, tcp, dxy);
    PDSB_CHECK(c.folded.ensure(folded_floats(g) * sizeof(float)));
    PDSB_CHECK(launch_fold(img_dev, c.folded.as<float>(), g));
    const int nsplit = dft_auto_split(variant, ds->nuvh, nf, g.ntile);
    PDSB_CHECK(c.partial.ensure((size_t)nsplit * nf * ds->nuvh * sizeof(double2)));
    DftParams p;
    p.F = c.folded.as<float>();
    p.u = ds->u;
    p.v = ds->v;
    p.nuvh = ds->nuvh;
    p.dxy = dxy;
    p.ntile = g.ntile;
    p.nchunk = g.nchunk;
    p.nsplit = nsplit;
    p.nf = nf;
    p.hx = g.hx2 ? 0.5 : 0.0;
    p.hy = g.hy2 ? 0.5 : 0.0;
    p.part = c.partial.as<double2>();
    PDSB_CHECK(launch_dft(p, variant, nullptr));
    run->g = g;
    run->nsplit = nsplit;
    return PDSB_OK;
}

static EpiParams make_epi(const pdsb_dataset *ds, const DftRun &run, double dRA, double dDec)
{
    EpiParams e;
    e.part = ctx().partial.as<double2>();
    e.nsplit = run.nsplit;
    e.nf = run.g.nf;
    e.nuvh = ds->nuvh;
    e.nuv = ds->nuv;
    e.hermitian = ds->hermitian;
    e.u = ds->u;
    e.v = ds->v;
    e.xs = run.g.xcen - dRA;
    e.ys = run.g.ycen - dDec;
    e.chan_scale = nullptr;
    e.plane_unscale = run.plane_unscale;
    e.ff_flux = e.ff_x0 = e.ff_y0 = 0.0;
    return e;
}

// Model modifiers of the *_ex entry points (set for the duration of one call).
struct Mods {
    const double *chan_scale_host = nullptr;
    double ff_flux = 0.0, ff_x0 = 0.0, ff_y0 = 0.0;
};
static Mods g_mods;

static int apply_mods(EpiParams &e, int nf)
{
    Context &c = ctx();
    e.ff_flux = g_mods.ff_flux;
    e.ff_x0 = g_mods.ff_x0;
    e.ff_y0 = g_mods.ff_y0;
    if (g_mods.chan_scale_host) {
        PDSB_CHECK(c.small_dev.ensure((size_t)nf * sizeof(double)));
        PDSB_CUDA(cudaMemcpyAsync(c.small_dev.ptr, g_mods.chan_scale_host, (size_t)nf * sizeof(double),
                                  cudaMemcpyHostToDevice, c.stream));
        e.chan_scale = c.small_dev.as<double>();
    }
    return PDSB_OK;
}

// chi2 per channel for one image into chi2_dev[nf] (device)
static int run_loglike_dev(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind,
                           double dxy, double dRA, double dDec, double *chi2_dev)
{
    Context &c = ctx();
    PDSB_REQUIRE(ds && ds->has_data, "dataset has no data (call pdsb_dataset_set_data)");
    PDSB_REQUIRE(nf == ds->nf, "image channel count != dataset channel count");
    if (c.dft_variant == DFT_VARIANT_NUFFT && !g_mods.chan_scale_host && g_mods.ff_flux == 0.0) {
        // NUFFT kernel, no model modifiers: the sampler sums chi^2 per channel itself (no 16 bytes of partial sum per
        // (uv, channel) written and read back)
        PDSB_REQUIRE(ny > 0 && nx > 0 && nf > 0 && dxy > 0.0, "image shape / dxy");
        const double *img_dev = nullptr;
        int done = 0;
        PDSB_CHECK(to_device(image, image_kind, (size_t)ny * nx * nf * sizeof(double), c.img64, (const void **)&img_dev));
        PDSB_CHECK(nufft_chi2_channels(ds, img_dev, ny, nx, nf, dxy, dRA, dDec, chi2_dev, &done));
        if (done) return PDSB_OK;
        image = img_dev;                                 // already on the device for the general route below
        image_kind = PDSB_DEVICE;
    }
    DftRun run;
    PDSB_CHECK(run_dft(ds, image, ny, nx, nf, image_kind, dxy, &run));
    EpiParams e = make_epi(ds, run, dRA, dDec);
    PDSB_CHECK(apply_mods(e, nf));
    dim3 grid((unsigned)ceil_div(ds->nuvh, 32), (unsigned)ceil_div(nf, 32));
    PDSB_CHECK(c.red.ensure((size_t)grid.x * nf * sizeof(double)));
    {
        LaunchScope ls("loglike_epilogue");
        loglike_epi_kernel<<<grid, dim3(32, 8), 0, c.stream>>>(e, ds->re, ds->im, ds->w, c.red.as<double>());
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(reduce_blocks(c.red.as<double>(), (int)grid.x, nf, chi2_dev));
    return PDSB_OK;
}

}  // namespace pdsb

using namespace pdsb;

extern "C" {

int pdsb_dataset_create(const double *u, const double *v, int64_t nuv, int kind, pdsb_dataset **out)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(out, "out");
    *out = nullptr;
    PDSB_REQUIRE(nuv >= 0 && (nuv == 0 || (u && v)), "u/v/nuv");
    Context &c = ctx();
    std::vector<double> hu, hv;
    const double *pu = u, *pv = v;
    if (kind == PDSB_DEVICE && nuv > 0) {
        hu.resize(nuv);
        hv.resize(nuv);
        PDSB_CUDA(cudaMemcpyAsync(hu.data(), u, nuv * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(hv.data(), v, nuv * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
        pu = hu.data();
        pv = hv.data();
    }
    // Hermitian doubling (readuvfits.py:68-73): second half == -first half, exactly.
    int herm = 0;
    if (nuv >= 2 && nuv % 2 == 0) {
        herm = 1;
        const int64_t h = nuv / 2;
        for (int64_t k = 0; k < h; k++)
            if (pu[k + h] != -pu[k] || pv[k + h] != -pv[k]) {
                herm = 0;
                break;
            }
    }
    pdsb_dataset *ds = new pdsb_dataset();
    ds->nuv = nuv;
    ds->hermitian = herm;
    ds->nuvh = herm ? nuv / 2 : nuv;
    if (ds->nuvh > 0) {
        if (cudaMalloc(&ds->u, ds->nuvh * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&ds->v, ds->nuvh * sizeof(double)) != cudaSuccess) {
            set_error("cudaMalloc of uv arrays failed");
            cudaGetLastError();
            pdsb_dataset_destroy(ds);
            return PDSB_ERR_NOMEM;
        }
        PDSB_CUDA(cudaMemcpyAsync(ds->u, pu, ds->nuvh * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(ds->v, pv, ds->nuvh * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    *out = ds;
    return PDSB_OK;
}

int pdsb_dataset_set_data(pdsb_dataset *ds, const double *real, const double *imag, const double *weights, int nf,
                          int kind)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && real && imag && weights && nf > 0, "dataset/data");
    Context &c = ctx();
    const size_t n = (size_t)ds->nuv * nf;
    if (ds->has_data) {
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
        cudaFree(ds->re);
        cudaFree(ds->im);
        cudaFree(ds->w);
        ds->re = ds->im = ds->w = nullptr;
        ds->has_data = false;
    }
    ds->nf = nf;
    if (n == 0) {
        ds->logsum = 0.0;
        ds->has_data = true;
        return PDSB_OK;
    }
    if (cudaMalloc(&ds->re, n * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&ds->im, n * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&ds->w, n * sizeof(double)) != cudaSuccess) {
        set_error("cudaMalloc of data arrays (%zu doubles each) failed", n);
        cudaGetLastError();
        return PDSB_ERR_NOMEM;
    }
    if (kind == PDSB_DEVICE) {
        PDSB_CUDA(cudaMemcpyAsync(ds->re, real, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(ds->im, imag, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(ds->w, weights, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
    } else {
        PDSB_CHECK(copy_h2d(ds->re, real, n * sizeof(double)));
        PDSB_CHECK(copy_h2d(ds->im, imag, n * sizeof(double)));
        PDSB_CHECK(copy_h2d(ds->w, weights, n * sizeof(double)));
    }
    // data-only constant of the likelihood
    const int nb = (int)std::min<int64_t>((int64_t)c.sm_count * 8, (int64_t)((n + 255) / 256));
    PDSB_CHECK(c.red.ensure((size_t)(nb + 1) * sizeof(double)));
    {
        LaunchScope ls("logsum");
        logsum_kernel<<<nb, 256, 0, c.stream>>>(ds->w, (int64_t)n, c.red.as<double>());
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(reduce_blocks(c.red.as<double>(), nb, 1, c.red.as<double>() + nb));
    PDSB_CUDA(cudaMemcpyAsync(&ds->logsum, c.red.as<double>() + nb, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    ds->has_data = true;
    return PDSB_OK;
}

int pdsb_dataset_info(const pdsb_dataset *ds, int64_t *nuv, int64_t *nuv_unique, int *nf, int *hermitian)
{
    PDSB_REQUIRE(ds, "dataset");
    if (nuv) *nuv = ds->nuv;
    if (nuv_unique) *nuv_unique = ds->nuvh;
    if (nf) *nf = ds->nf;
    if (hermitian) *hermitian = ds->hermitian;
    return PDSB_OK;
}

int pdsb_dataset_destroy(pdsb_dataset *ds)
{
    if (!ds) return PDSB_OK;
    if (ctx().inited) cudaStreamSynchronize(ctx().stream);
    cudaFree(ds->u);
    cudaFree(ds->v);
    cudaFree(ds->re);
    cudaFree(ds->im);
    cudaFree(ds->w);
    cudaFree(ds->order);
    delete ds;
    return PDSB_OK;
}

int pdsb_sample_image(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                      double dRA, double dDec, double *out_real, double *out_imag, int out_kind)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && out_real && out_imag, "dataset/outputs");
    Context &c = ctx();
    if (ds->nuv == 0) return PDSB_OK;
    DftRun run;
    PDSB_CHECK(run_dft(ds, image, ny, nx, nf, image_kind, dxy, &run));
    EpiParams e = make_epi(ds, run, dRA, dDec);
    PDSB_CHECK(apply_mods(e, nf));
    const size_t bytes = (size_t)ds->nuv * nf * sizeof(double);
    double *ore = out_real, *oim = out_imag;
    if (out_kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_a.ensure(bytes));
        PDSB_CHECK(c.stage_b.ensure(bytes));
        ore = c.stage_a.as<double>();
        oim = c.stage_b.as<double>();
    }
    dim3 grid((unsigned)ceil_div(ds->nuvh, 32), (unsigned)ceil_div(nf, 32));
    {
        LaunchScope ls("finish_vis");
        finish_vis_kernel<<<grid, dim3(32, 8), 0, c.stream>>>(e, ore, oim);
        PDSB_CUDA(cudaGetLastError());
    }
    if (out_kind == PDSB_HOST) {
        PDSB_CHECK(copy_d2h(out_real, ore, bytes));
        PDSB_CHECK(copy_d2h(out_imag, oim, bytes));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}

static int loglike_impl(pdsb_dataset *ds, const double *images, int nwalkers, int ny, int nx, int nf,
                        int image_kind, double dxy, const double *dRA, const double *dDec, double *chi2,
                        double *lnlike)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && images && dRA && dDec && nwalkers > 0, "arguments");
    PDSB_REQUIRE(chi2 || lnlike, "no output requested");
    PDSB_REQUIRE(ds->has_data, "dataset has no data (call pdsb_dataset_set_data)");
    PDSB_REQUIRE(nf == ds->nf, "image channel count != dataset channel count");
    Context &c = ctx();
    std::vector<double> h((size_t)nwalkers * nf, 0.0);
    if (ds->nuv > 0) {
        PDSB_CHECK(c.stage_c.ensure((size_t)nwalkers * nf * sizeof(double)));
        double *chi2_dev = c.stage_c.as<double>();
        const size_t cube = (size_t)ny * nx * nf;
        if (nwalkers > 1 && image_kind == PDSB_HOST) {
            // Walker batch from host memory: cube wk + 1 travels on the copy stream into the other of two device
            // buffers while cube wk is evaluated on the compute stream.  cube_ready[b]: buffer b holds its cube
            // (compute waits); cube_free[b]: the evaluation that read buffer b is done (the next copy into b waits).
            const size_t bytes = cube * sizeof(double);
            PDSB_CHECK(c.img64.ensure(bytes));
            PDSB_CHECK(c.img64_b.ensure(bytes));
            double *buf[2] = {c.img64.as<double>(), c.img64_b.as<double>()};
            PDSB_CUDA(cudaEventRecord(c.cube_free[0], c.stream));          // earlier work on the stream may read buffer 0
            PDSB_CUDA(cudaStreamWaitEvent(c.copy_stream, c.cube_free[0], 0));
            int rc = copy_h2d_on(buf[0], images, bytes, c.copy_stream);
            if (rc == PDSB_OK && cudaEventRecord(c.cube_ready[0], c.copy_stream) != cudaSuccess) rc = PDSB_ERR_CUDA;
            for (int wk = 0; wk < nwalkers && rc == PDSB_OK; wk++) {
                const int b = wk & 1;
                if (cudaStreamWaitEvent(c.stream, c.cube_ready[b], 0) != cudaSuccess) rc = PDSB_ERR_CUDA;
                if (rc == PDSB_OK)
                    rc = run_loglike_dev(ds, buf[b], ny, nx, nf, PDSB_DEVICE, dxy, dRA[wk], dDec[wk],
                                         chi2_dev + (size_t)wk * nf);
                if (rc == PDSB_OK && cudaEventRecord(c.cube_free[b], c.stream) != cudaSuccess) rc = PDSB_ERR_CUDA;
                if (rc == PDSB_OK && wk + 1 < nwalkers) {
                    if (wk >= 1 && cudaStreamWaitEvent(c.copy_stream, c.cube_free[b ^ 1], 0) != cudaSuccess)
                        rc = PDSB_ERR_CUDA;
                    if (rc == PDSB_OK) rc = copy_h2d_on(buf[b ^ 1], images + (size_t)(wk + 1) * cube, bytes, c.copy_stream);
                    if (rc == PDSB_OK && cudaEventRecord(c.cube_ready[b ^ 1], c.copy_stream) != cudaSuccess)
                        rc = PDSB_ERR_CUDA;
                }
            }
            if (rc != PDSB_OK) {                                           // leave both streams idle before reporting
                cudaStreamSynchronize(c.copy_stream);
                cudaStreamSynchronize(c.stream);
                if (rc == PDSB_ERR_CUDA) {
                    const cudaError_t e = cudaGetLastError();
                    if (e != cudaSuccess) set_error("walker batch: %s", cudaGetErrorString(e));
                }
                return rc;
            }
        } else {
            for (int wk = 0; wk < nwalkers; wk++)
                PDSB_CHECK(run_loglike_dev(ds, images + (size_t)wk * cube, ny, nx, nf, image_kind, dxy, dRA[wk],
                                           dDec[wk], chi2_dev + (size_t)wk * nf));
        }
        PDSB_CUDA(cudaMemcpyAsync(h.data(), chi2_dev, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    for (int wk = 0; wk < nwalkers; wk++) {
        double s = 0.0;
        for (int i = 0; i < nf; i++) {
            s += h[(size_t)wk * nf + i];
            if (chi2) chi2[(size_t)wk * nf + i] = h[(size_t)wk * nf + i];
        }
        // emcee.py:31-43: -0.5 chi2_re - L - 0.5 chi2_im - L
        if (lnlike) lnlike[wk] = -0.5 * s - 2.0 * ds->logsum;
    }
    return PDSB_OK;
}

int pdsb_loglike(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                 double dRA, double dDec, double *chi2, double *lnlike)
{
    return loglike_impl(ds, image, 1, ny, nx, nf, image_kind, dxy, &dRA, &dDec, chi2, lnlike);
}

int pdsb_sample_image_ex(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                         double dRA, double dDec, const double *chan_scale, double ff_flux, double ff_x0,
                         double ff_y0, double *out_real, double *out_imag, int out_kind)
{
    g_mods.chan_scale_host = chan_scale;
    g_mods.ff_flux = ff_flux;
    g_mods.ff_x0 = ff_x0;
    g_mods.ff_y0 = ff_y0;
    const int rc = pdsb_sample_image(ds, image, ny, nx, nf, image_kind, dxy, dRA, dDec, out_real, out_imag, out_kind);
    g_mods = Mods();
    return rc;
}

int pdsb_loglike_ex(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                    double dRA, double dDec, const double *chan_scale, double ff_flux, double ff_x0, double ff_y0,
                    double *chi2, double *lnlike)
{
    g_mods.chan_scale_host = chan_scale;
    g_mods.ff_flux = ff_flux;
    g_mods.ff_x0 = ff_x0;
    g_mods.ff_y0 = ff_y0;
    const int rc = pdsb_loglike(ds, image, ny, nx, nf, image_kind, dxy, dRA, dDec, chi2, lnlike);
    g_mods = Mods();
    return rc;
}

int pdsb_loglike_device(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                        double dRA, double dDec, double *chi2_dev)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && image && chi2_dev, "arguments");
    PDSB_REQUIRE(ds->has_data, "dataset has no data (call pdsb_dataset_set_data)");
    if (ds->nuv == 0) {
        PDSB_CUDA(cudaMemsetAsync(chi2_dev, 0, (size_t)nf * sizeof(double), ctx().stream));
        return PDSB_OK;
    }
    return run_loglike_dev(ds, image, ny, nx, nf, image_kind, dxy, dRA, dDec, chi2_dev);
}

int pdsb_dataset_logsum(const pdsb_dataset *ds, double *logsum)
{
    PDSB_REQUIRE(ds && logsum, "arguments");
    PDSB_REQUIRE(ds->has_data, "dataset has no data");
    *logsum = ds->logsum;
    return PDSB_OK;
}

int pdsb_loglike_batch(pdsb_dataset *ds, const double *images, int nwalkers, int ny, int nx, int nf, int image_kind,
                       double dxy, const double *dRA, const double *dDec, double *lnlike)
{
    return loglike_impl(ds, images, nwalkers, ny, nx, nf, image_kind, dxy, dRA, dDec, nullptr, lnlike);
}

static int fft_group_size(int nf)
{
    int gs = 1;
    while (gs * 2 <= nf && gs < 32) gs *= 2;
    return gs;
}

static int nufft_order(pdsb_dataset *ds);               // the Morton visiting order (defined with the NUFFT path)

// galario's FFT stage: the transformed cube of every channel (rfft2_planes) and the sampling arguments
static int run_fft_transform(pdsb_dataset *ds, const double *image, int n, int nf, int image_kind, double dxy, double dRA,
                             double dDec, FftSampleArgs *a)
{
    Context &c = ctx();
    PDSB_REQUIRE(ds && image, "dataset/image");
    PDSB_REQUIRE(n >= 2 && n <= 4096 && (n & (n - 1)) == 0, "the FFT path needs a square image, side a power of two <= 4096");
    PDSB_REQUIRE(nf > 0 && dxy > 0.0, "nf/dxy");
    const double *img_dev = nullptr;
    const size_t nn = (size_t)n * n;
    PDSB_CHECK(to_device(image, image_kind, nn * nf * sizeof(double), c.img64, (const void **)&img_dev));
    const size_t nh = (size_t)n * (n / 2 + 1);                             // half spectrum: columns 0 .. n/2
    PDSB_CHECK(c.folded.ensure(2 * nh * nf * sizeof(double2)));            // [T | Y]: the DFT's scratch, free here
    double2 *T = c.folded.as<double2>(), *Y = T + nh * nf;
    PDSB_CHECK(rfft2_planes(img_dev, n, nf, 1, T, Y));
    PDSB_REQUIRE(ds->nuvh < ((int64_t)1 << 31), "more than 2^31 unique uv points");
    PDSB_CHECK(nufft_order(ds));
    *a = FftSampleArgs{Y, ds->u, ds->v, ds->order, ds->nuv, ds->nuvh, n, nf, dxy, dRA, dDec};
    return PDSB_OK;
}

int pdsb_sample_image_fft(pdsb_dataset *ds, const double *image, int n, int nf, int image_kind, double dxy, double dRA,
                          double dDec, double *out_real, double *out_imag, int out_kind)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && out_real && out_imag, "dataset/outputs");
    Context &c = ctx();
    if (ds->nuv == 0) return PDSB_OK;
    const size_t bytes = (size_t)ds->nuv * nf * sizeof(double);
    double *ore = out_real, *oim = out_imag;
    if (out_kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_a.ensure(bytes));
        PDSB_CHECK(c.stage_b.ensure(bytes));
        ore = c.stage_a.as<double>();
        oim = c.stage_b.as<double>();
    }
    FftSampleArgs fa;
    PDSB_CHECK(run_fft_transform(ds, image, n, nf, image_kind, dxy, dRA, dDec, &fa));
    {
        LaunchScope ls("fft_sample");
        const int gs = fft_group_size(nf);
        fft_sample_kernel<<<ceil_div(ds->nuvh * gs, 256), 256, 0, c.stream>>>(fa, gs, ore, oim);
        PDSB_CUDA(cudaGetLastError());
    }
    if (out_kind == PDSB_HOST) {
        PDSB_CHECK(copy_d2h(out_real, ore, bytes));
        PDSB_CHECK(copy_d2h(out_imag, oim, bytes));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}

int pdsb_loglike_fft(pdsb_dataset *ds, const double *image, int n, int nf, int image_kind, double dxy, double dRA,
                     double dDec, double *out)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && out, "dataset/out");
    PDSB_REQUIRE(ds->has_data && ds->nf == nf, "dataset has no data or a different channel count");
    Context &c = ctx();
    const int64_t cnt = ds->nuv * nf;
    if (cnt == 0) {
        out[0] = out[1] = out[2] = 0.0;
        out[3] = -0.0;
        return PDSB_OK;
    }
    FftSampleArgs fa;
    PDSB_CHECK(run_fft_transform(ds, image, n, nf, image_kind, dxy, dRA, dDec, &fa));
    const int gs = fft_group_size(nf);
    const int nb = (int)std::min<int64_t>((int64_t)c.sm_count * 16, (ds->nuvh * gs + 255) / 256);
    PDSB_CHECK(c.red.ensure((size_t)(nb + 1) * 2 * sizeof(double)));
    {
        LaunchScope ls("fft_chi2");
        fft_chi2_kernel<<<nb, 256, 0, c.stream>>>(fa, gs, ds->re, ds->im, ds->w, c.red.as<double>());
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(reduce_blocks(c.red.as<double>(), nb, 2, c.red.as<double>() + (size_t)nb * 2));
    double h[3];
    PDSB_CUDA(cudaMemcpyAsync(h, c.red.as<double>() + (size_t)nb * 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    h[2] = ds->logsum;                           // sum log(w / 2 pi) over w > 0: data only, formed once at upload
    out[0] = h[0];
    out[1] = h[1];
    out[2] = h[2];
    out[3] = -0.5 * h[0] - h[2] + -0.5 * h[1] - h[2];
    return PDSB_OK;
}

// ---- NUFFT path (code="nufft") --------------------------------------------------
// 1 / psi_hat(X / N), X = -n/2 .. n/2 - 1: psi_hat(xi) = (w/2) Int_{-pi/2}^{pi/2} exp(beta (cos th - 1)) cos(pi w xi sin th) cos th dth
// (z = sin th removes the square-root end points; midpoint rule, 400 nodes: 1e-13)
static int nufft_corr_table(int ny, int nx, int N, const double **dev_y, const double **dev_x)
{
    Context &c = ctx();
    PDSB_CHECK(c.nufft_corr.ensure((size_t)(ny + nx) * sizeof(double)));
    if (c.nufft_corr_n != ny * 8192 + nx || c.nufft_corr_N != N) {
        constexpr int M = 400;
        const double pi = 3.14159265358979323846;
        std::vector<double> h((size_t)(ny + nx)), f(M), sn(M);
        for (int m = 0; m < M; m++) {
            const double th = (m + 0.5) * pi / M - 0.5 * pi;
            f[m] = exp(NUFFT_BETA * (cos(th) - 1.0)) * cos(th);
            sn[m] = sin(th);
        }
        for (int x = 0; x < ny + nx; x++) {
            const int n = x < ny ? ny : nx, X = (x < ny ? x : x - ny) - n / 2;
            const double xi = (double)X / (double)N;
            double acc = 0.0;
            for (int m = 0; m < M; m++) acc += f[m] * cos(pi * NUFFT_W * xi * sn[m]);
            h[x] = 1.0 / (0.5 * NUFFT_W * (pi / M) * acc);
        }
        PDSB_CUDA(cudaMemcpyAsync(c.nufft_corr.ptr, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));            // h goes out of scope
        c.nufft_corr_n = ny * 8192 + nx;
        c.nufft_corr_N = N;
    }
    *dev_y = c.nufft_corr.as<double>();
    *dev_x = c.nufft_corr.as<double>() + ny;
    return PDSB_OK;
}

// The unique uv points along a Morton (Z-order) curve of the uv plane, built once per data set: the 8 x 8 taps of
// neighbouring visibilities overlap, so visiting them in this order turns most tap reads into L2 hits.
static int nufft_order(pdsb_dataset *ds)
{
    if (ds->order || ds->nuvh == 0) return PDSB_OK;
    if (getenv("PDSB_NUFFT_NO_ORDER")) return PDSB_OK;            // tuning: measure the sampler without the visiting order
    Context &c = ctx();
    const size_t n = (size_t)ds->nuvh;
    std::vector<double> hu(n), hv(n);
    PDSB_CUDA(cudaMemcpyAsync(hu.data(), ds->u, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaMemcpyAsync(hv.data(), ds->v, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    double lo[2] = {hu[0], hv[0]}, hi[2] = {hu[0], hv[0]};
    for (size_t i = 0; i < n; i++) {
        lo[0] = std::min(lo[0], hu[i]); hi[0] = std::max(hi[0], hu[i]);
        lo[1] = std::min(lo[1], hv[i]); hi[1] = std::max(hi[1], hv[i]);
    }
    const double span = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), 1e-300);
    auto spread = [](uint32_t x) {                       // 16 bits -> every second bit of 32
        x &= 0xffffu;
        x = (x | (x << 8)) & 0x00ff00ffu;
        x = (x | (x << 4)) & 0x0f0f0f0fu;
        x = (x | (x << 2)) & 0x33333333u;
        x = (x | (x << 1)) & 0x55555555u;
        return x;
    };
    std::vector<std::pair<uint32_t, int>> key(n);
    for (size_t i = 0; i < n; i++) {
        const double fx = (hu[i] - lo[0]) / span, fy = (hv[i] - lo[1]) / span;
        const uint32_t qx = (uint32_t)std::min(65535.0, std::max(0.0, fx * 65535.0));
        const uint32_t qy = (uint32_t)std::min(65535.0, std::max(0.0, fy * 65535.0));
        key[i] = {spread(qx) | (spread(qy) << 1), (int)i};
    }
    std::sort(key.begin(), key.end());
    std::vector<int> ord(n);
    for (size_t i = 0; i < n; i++) ord[i] = key[i].second;
    if (cudaMalloc(&ds->order, n * sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        ds->order = nullptr;                             // no memory for it: identity order, slower, still correct
        return PDSB_OK;
    }
    PDSB_CUDA(cudaMemcpyAsync(ds->order, ord.data(), n * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    return PDSB_OK;
}

static int nufft_transform_dev(pdsb_dataset *ds, const double *img_dev, int ny, int nx, int nf, double dxy, double dRA,
                               double dDec, NufftArgs *a);
static int run_nufft_transform(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                               double dRA, double dDec, NufftArgs *a)
{
    Context &c = ctx();
    PDSB_REQUIRE(ds && image, "dataset/image");
    const double *img_dev = nullptr;
    PDSB_REQUIRE(ny > 0 && nx > 0 && nf > 0, "image shape");
    PDSB_CHECK(to_device(image, image_kind, (size_t)ny * nx * nf * sizeof(double), c.img64, (const void **)&img_dev));
    return nufft_transform_dev(ds, img_dev, ny, nx, nf, dxy, dRA, dDec, a);
}

static int nufft_transform_dev(pdsb_dataset *ds, const double *img_dev, int ny, int nx, int nf, double dxy, double dRA,
                               double dDec, NufftArgs *a)
{
    Context &c = ctx();
    PDSB_REQUIRE(ds && img_dev, "dataset/image");
    PDSB_REQUIRE(ny >= 2 && nx >= 2 && ny % 2 == 0 && nx % 2 == 0 && ny <= 2048 && nx <= 2048,
                 "the NUFFT path needs even image sides in [2, 2048] (odd sides: use the direct transform)");
    PDSB_REQUIRE(nf > 0 && dxy > 0.0, "nf/dxy");
    int N = 4;                                           // oversampled grid: a power of two >= 2 max(ny, nx)
    while (N < 2 * std::max(ny, nx)) N *= 2;
    const double *corr_y = nullptr, *corr_x = nullptr;
    PDSB_CHECK(nufft_corr_table(ny, nx, N, &corr_y, &corr_x));
    const size_t nh = (size_t)N * (N / 2 + 1);
    PDSB_CHECK(c.folded.ensure(2 * nh * nf * sizeof(double2)));            // [T | Yh]
    double2 *T = c.folded.as<double2>(), *Y = T + nh * nf;
    PDSB_CHECK(rfft2_planes_padded(img_dev, ny, nx, N, nf, 1, corr_y, corr_x, T, Y));
    PDSB_REQUIRE(ds->nuvh < ((int64_t)1 << 31), "more than 2^31 unique uv points");
    PDSB_CHECK(nufft_order(ds));
    *a = NufftArgs{Y, ds->u, ds->v, ds->order, ds->nuv, ds->nuvh, N, nf, dxy, dRA, dDec};
    return PDSB_OK;
}

static int nufft_partials(pdsb_dataset *ds, const double *img_dev, int ny, int nx, int nf, double dxy, double2 *part)
{
    Context &c = ctx();
    if (ds->nuvh == 0) return PDSB_OK;
    NufftArgs fa;
    PDSB_CHECK(nufft_transform_dev(ds, img_dev, ny, nx, nf, dxy, 0.0, 0.0, &fa));
    LaunchScope ls("nufft_sample");
    if (nf >= NT_CG && !getenv("PDSB_NUFFT_DIRECT")) {
        static bool attr = false;
        if (!attr) {
            PDSB_CUDA(cudaFuncSetAttribute(nufft_chi2_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)NT_SMEM));
            attr = true;
        }
        const int nb = (int)std::min<int64_t>((int64_t)c.sm_count * NT_MINB, (ds->nuvh + NT_PTS - 1) / NT_PTS);
        nufft_chi2_tiled_kernel<1><<<nb, 256, NT_SMEM, c.stream>>>(fa, nullptr, nullptr, nullptr, nullptr, part);
    } else {
        const int gs = fft_group_size(nf);
        nufft_sample_kernel<<<ceil_div(ds->nuvh * gs, 256), 256, 0, c.stream>>>(fa, gs, nullptr, nullptr, part);
    }
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

static int nufft_chi2_channels(pdsb_dataset *ds, const double *img_dev, int ny, int nx, int nf, double dxy, double dRA,
                               double dDec, double *chi2_dev, int *done)
{
    Context &c = ctx();
    *done = 0;
    if (ny % 2 || nx % 2 || ny > 2048 || nx > 2048 || nf < NT_CG || nf > NT_CG * NT_MAXG || ds->nuvh == 0 ||
        getenv("PDSB_NUFFT_DIRECT"))
        return PDSB_OK;
    NufftArgs fa;
    PDSB_CHECK(nufft_transform_dev(ds, img_dev, ny, nx, nf, dxy, dRA, dDec, &fa));
    const int nb = (int)std::min<int64_t>((int64_t)c.sm_count * NT_MINB, (ds->nuvh + NT_PTS - 1) / NT_PTS);
    PDSB_CHECK(c.red.ensure((size_t)nb * nf * sizeof(double)));
    {
        LaunchScope ls("nufft_chi2");
        static bool attr = false;
        if (!attr) {
            PDSB_CUDA(cudaFuncSetAttribute(nufft_chi2_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)NT_SMEM));
            attr = true;
        }
        nufft_chi2_tiled_kernel<2><<<nb, 256, NT_SMEM, c.stream>>>(fa, ds->re, ds->im, ds->w, c.red.as<double>(), nullptr);
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(reduce_blocks(c.red.as<double>(), nb, nf, chi2_dev));
    *done = 1;
    return PDSB_OK;
}

int pdsb_sample_image_nufft(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy, double dRA,
                            double dDec, double *out_real, double *out_imag, int out_kind)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && out_real && out_imag, "dataset/outputs");
    Context &c = ctx();
    if (ds->nuv == 0) return PDSB_OK;
    if (nf >= NT_CG && ny > 0 && nx > 0 && ny % 2 == 0 && nx % 2 == 0 && ny <= 2048 && nx <= 2048 && !getenv("PDSB_NUFFT_DIRECT")) {
        // cubes: the shared-memory-staged sampler (partial sums into the epilogue shared with the direct-sum kernels),
        // the same route set_dft_kernel("nufft") takes
        const int keep = c.dft_variant;
        c.dft_variant = DFT_VARIANT_NUFFT;
        const int rc = pdsb_sample_image(ds, image, ny, nx, nf, image_kind, dxy, dRA, dDec, out_real, out_imag, out_kind);
        c.dft_variant = keep;
        return rc;
    }
    const size_t bytes = (size_t)ds->nuv * nf * sizeof(double);
    double *ore = out_real, *oim = out_imag;
    if (out_kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_a.ensure(bytes));
        PDSB_CHECK(c.stage_b.ensure(bytes));
        ore = c.stage_a.as<double>();
        oim = c.stage_b.as<double>();
    }
    NufftArgs fa;
    PDSB_CHECK(run_nufft_transform(ds, image, ny, nx, nf, image_kind, dxy, dRA, dDec, &fa));
    {
        LaunchScope ls("nufft_sample");
        const int gs = fft_group_size(nf);
        nufft_sample_kernel<<<ceil_div(ds->nuvh * gs, 256), 256, 0, c.stream>>>(fa, gs, ore, oim, nullptr);
        PDSB_CUDA(cudaGetLastError());
    }
    if (out_kind == PDSB_HOST) {
        PDSB_CHECK(copy_d2h(out_real, ore, bytes));
        PDSB_CHECK(copy_d2h(out_imag, oim, bytes));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}

int pdsb_loglike_nufft(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy, double dRA,
                       double dDec, double *out)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && out, "dataset/out");
    PDSB_REQUIRE(ds->has_data && ds->nf == nf, "dataset has no data or a different channel count");
    Context &c = ctx();
    if (ds->nuv * nf == 0) {
        out[0] = out[1] = out[2] = 0.0;
        out[3] = -0.0;
        return PDSB_OK;
    }
    NufftArgs fa;
    PDSB_CHECK(run_nufft_transform(ds, image, ny, nx, nf, image_kind, dxy, dRA, dDec, &fa));
    const int gs = fft_group_size(nf);
    int nb = (int)std::min<int64_t>((int64_t)c.sm_count * 16, (ds->nuvh * gs + 255) / 256);
    const bool tiled = nf >= NT_CG && !getenv("PDSB_NUFFT_DIRECT");            // (tuning switch: the untiled sampler)
    if (tiled) nb = (int)std::min<int64_t>((int64_t)c.sm_count * NT_MINB, (ds->nuvh + NT_PTS - 1) / NT_PTS);       // persistent
    PDSB_CHECK(c.red.ensure((size_t)(nb + 1) * 2 * sizeof(double)));
    {
        LaunchScope ls("nufft_chi2");
        if (tiled) {
            static bool attr = false;
            if (!attr) {
                PDSB_CUDA(cudaFuncSetAttribute(nufft_chi2_tiled_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)NT_SMEM));
                attr = true;
            }
            nufft_chi2_tiled_kernel<0><<<nb, 256, NT_SMEM, c.stream>>>(fa, ds->re, ds->im, ds->w, c.red.as<double>(),
                                                                         nullptr);
        } else {
            nufft_chi2_kernel<<<nb, 256, 0, c.stream>>>(fa, gs, ds->re, ds->im, ds->w, c.red.as<double>());
        }
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(reduce_blocks(c.red.as<double>(), nb, 2, c.red.as<double>() + (size_t)nb * 2));
    double h[2];
    PDSB_CUDA(cudaMemcpyAsync(h, c.red.as<double>() + (size_t)nb * 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    out[0] = h[0];
    out[1] = h[1];
    out[2] = ds->logsum;
    out[3] = -0.5 * h[0] - ds->logsum + -0.5 * h[1] - ds->logsum;
    return PDSB_OK;
}

// code="trift": exact transform of a triangulated scattered-point image (trift.cu); same epilogues as the pixel kernels
int pdsb_sample_triangles(pdsb_dataset *ds, const void *tri_records, int ntri, const double *values, int64_t npts, int nf,
                          int values_kind, double dRA, double dDec, double *out_real, double *out_imag, int out_kind)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && out_real && out_imag, "dataset/outputs");
    PDSB_REQUIRE(ntri >= 0 && npts > 0 && nf > 0 && (ntri == 0 || tri_records) && values, "triangles/values");
    Context &c = ctx();
    if (ds->nuv == 0) return PDSB_OK;
    const size_t tb = (size_t)ntri * trift_record_bytes();
    PDSB_CHECK(c.folded.ensure(tb + 64));
    if (ntri > 0) PDSB_CUDA(cudaMemcpyAsync(c.folded.ptr, tri_records, tb, cudaMemcpyHostToDevice, c.stream));
    const double *vals = nullptr;
    PDSB_CHECK(to_device(values, values_kind, (size_t)npts * nf * sizeof(double), c.img64, (const void **)&vals));
    PDSB_CHECK(c.partial.ensure((size_t)nf * ds->nuvh * sizeof(double2)));
    PDSB_CHECK(launch_trift(c.folded.ptr, ntri, vals, nf, ds->u, ds->v, ds->nuvh, c.partial.as<double2>()));
    DftRun run;
    run.g = make_geom(2, 2, nf, 32, 1.0);
    run.g.xcen = run.g.ycen = 0.0;
    run.nsplit = 1;
    EpiParams e = make_epi(ds, run, dRA, dDec);
    PDSB_CHECK(apply_mods(e, nf));
    const size_t bytes = (size_t)ds->nuv * nf * sizeof(double);
    double *ore = out_real, *oim = out_imag;
    if (out_kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_a.ensure(bytes));
        PDSB_CHECK(c.stage_b.ensure(bytes));
        ore = c.stage_a.as<double>();
        oim = c.stage_b.as<double>();
    }
    dim3 grid((unsigned)ceil_div(ds->nuvh, 32), (unsigned)ceil_div(nf, 32));
    {
        LaunchScope ls("finish_vis");
        finish_vis_kernel<<<grid, dim3(32, 8), 0, c.stream>>>(e, ore, oim);
        PDSB_CUDA(cudaGetLastError());
    }
    if (out_kind == PDSB_HOST) {
        PDSB_CHECK(copy_d2h(out_real, ore, bytes));
        PDSB_CHECK(copy_d2h(out_imag, oim, bytes));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}

int pdsb_triangle_record_bytes(void) { return (int)trift_record_bytes(); }

int pdsb_chi2(const double *d_real, const double *d_imag, const double *weights, const double *m_real,
              const double *m_imag, int64_t n, int kind, double *out)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(out && n >= 0, "out/n");
    Context &c = ctx();
    if (n == 0) {
        out[0] = out[1] = out[2] = 0.0;
        out[3] = -0.0;
        return PDSB_OK;
    }
    PDSB_REQUIRE(d_real && d_imag && weights && m_real && m_imag, "arrays");
    const size_t bytes = (size_t)n * sizeof(double);
    const double *a, *b, *w, *mr, *mi;
    PDSB_CHECK(to_device(d_real, kind, bytes, c.stage_a, (const void **)&a));
    PDSB_CHECK(to_device(d_imag, kind, bytes, c.stage_b, (const void **)&b));
    PDSB_CHECK(to_device(weights, kind, bytes, c.stage_c, (const void **)&w));
    PDSB_CHECK(to_device(m_real, kind, bytes, c.stage_d, (const void **)&mr));
    PDSB_CHECK(to_device(m_imag, kind, bytes, c.stage_e, (const void **)&mi));
    const int nb = (int)std::min<int64_t>((int64_t)c.sm_count * 8, (n + 255) / 256);
    PDSB_CHECK(c.red.ensure((size_t)(nb + 1) * 3 * sizeof(double)));
    {
        LaunchScope ls("chi2_flat");
        chi2_flat_kernel<<<nb, 256, 0, c.stream>>>(a, b, w, mr, mi, n, c.red.as<double>());
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(reduce_blocks(c.red.as<double>(), nb, 3, c.red.as<double>() + (size_t)nb * 3));
    double h[3];
    PDSB_CUDA(cudaMemcpyAsync(h, c.red.as<double>() + (size_t)nb * 3, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    out[0] = h[0];
    out[1] = h[1];
    out[2] = h[2];
    out[3] = -0.5 * h[0] - h[2] + -0.5 * h[1] - h[2];
    return PDSB_OK;
}

int pdsb_chi2_dataset(pdsb_dataset *ds, const double *m_real, const double *m_imag, int kind, double *out)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ds && out, "dataset/out");
    PDSB_REQUIRE(ds->has_data, "dataset has no data (call pdsb_dataset_set_data)");
    Context &c = ctx();
    const int64_t n = ds->nuv * ds->nf;
    if (n == 0) {
        out[0] = out[1] = out[2] = 0.0;
        out[3] = -0.0;
        return PDSB_OK;
    }
    PDSB_REQUIRE(m_real && m_imag, "model arrays");
    const size_t bytes = (size_t)n * sizeof(double);
    const double *mr, *mi;
    PDSB_CHECK(to_device(m_real, kind, bytes, c.stage_d, (const void **)&mr));
    PDSB_CHECK(to_device(m_imag, kind, bytes, c.stage_e, (const void **)&mi));
    const int nb = (int)std::min<int64_t>((int64_t)c.sm_count * 8, (n + 255) / 256);
    PDSB_CHECK(c.red.ensure((size_t)(nb + 1) * 3 * sizeof(double)));
    {
        LaunchScope ls("chi2_flat");
        chi2_flat_kernel<<<nb, 256, 0, c.stream>>>(ds->re, ds->im, ds->w, mr, mi, n, c.red.as<double>());
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(reduce_blocks(c.red.as<double>(), nb, 3, c.red.as<double>() + (size_t)nb * 3));
    double h[3];
    PDSB_CUDA(cudaMemcpyAsync(h, c.red.as<double>() + (size_t)nb * 3, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    out[0] = h[0];
    out[1] = h[1];
    out[2] = h[2];
    out[3] = -0.5 * h[0] - h[2] + -0.5 * h[1] - h[2];
    return PDSB_OK;
}

int pdsb_chisq(const double *d_real, const double *d_imag, const double *weights, const double *m_real,
               const double *m_imag, int64_t nuv, int nf, int kind, float *out)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(out && nuv >= 0 && nf > 0, "out/nuv/nf");
    Context &c = ctx();
    if (nuv == 0) {
        *out = 0.f;
        return PDSB_OK;
    }
    PDSB_REQUIRE(d_real && d_imag && weights && m_real && m_imag, "arrays");
    const size_t bytes = (size_t)nuv * nf * sizeof(double);
    const double *a, *b, *w, *mr, *mi;
    PDSB_CHECK(to_device(d_real, kind, bytes, c.stage_a, (const void **)&a));
    PDSB_CHECK(to_device(d_imag, kind, bytes, c.stage_b, (const void **)&b));
    PDSB_CHECK(to_device(weights, kind, bytes, c.stage_c, (const void **)&w));
    PDSB_CHECK(to_device(m_real, kind, bytes, c.stage_d, (const void **)&mr));
    PDSB_CHECK(to_device(m_imag, kind, bytes, c.stage_e, (const void **)&mi));
    const int nb = (int)std::min<int64_t>((int64_t)c.sm_count * 8, (nuv + 255) / 256);
    PDSB_CHECK(c.red.ensure((size_t)(nb + 1) * sizeof(double)));
    {
        LaunchScope ls("chisq_ch0");
        chisq_ch0_kernel<<<nb, 256, 0, c.stream>>>(a, b, w, mr, mi, nuv, nf, c.red.as<double>());
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(reduce_blocks(c.red.as<double>(), nb, 1, c.red.as<double>() + nb));
    double h;
    PDSB_CUDA(cudaMemcpyAsync(&h, c.red.as<double>() + nb, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    *out = (float)h;      // libinterferometry.pyx:616 `cdef float chisq_calc`
    return PDSB_OK;
}

}  // extern "C"
