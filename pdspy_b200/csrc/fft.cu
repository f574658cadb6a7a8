// Dirty-image synthesis for invert() (pdspy/interferometry/invert.py:63-84): per channel
//   im       = fftshift(ifft2(ifftshift(real + i imag))).real * imsize^2
//   convolve = fftshift(ifft2(ifftshift(conv_func(u, v)))).real
//   image[:, :, i] = (im / convolve)[:, ::-1]
// Hand-written fp64 radix-2 FFT: one CTA per row held in shared memory (n <= 4096), two passes with
// the (i)fftshift index rotations and the transpose fused into the loads/stores.  Sizes that are not a
// power of two (the reference takes any imsize: scipy's ifft2) go through Bluestein's chirp-z form of the
// same row transform on the next power of two >= 2n - 1 (bluestein_rows_kernel).
#include "common.cuh"
#include "fft_r16.cuh"
#include <algorithm>

namespace pdsb {

__device__ __forceinline__ unsigned bitrev(unsigned x, int bits) { return __brev(x) >> (32 - bits); }

// One row of length n (power of two) per block: out[...] = sum_c in[...] exp(+2 pi i b c / n).
//   pass 0: row r' of the shifted input: element c comes from x[(r'+h)%n][(c+h)%n] (real/imag planes with
//           element stride `estride`), result b goes to T[b][r']           (transpose)
//   pass 1: row b of T, result a goes to Y[(a+h)%n][(b+h)%n]              (fftshift on output)
// Batched over blockIdx.y = plane: pass-0 input plane p starts p * in_pstride elements further (1 for a
// channel-fastest cube with estride = nf) and may be read with its rows reversed (flip); the transposed
// intermediate of plane p is T + p n^2; the pass-1 result element (A, B) of plane p goes to
// tout[(A n + B) * o_estride + p * o_pstride].
__global__ void __launch_bounds__(512) ifft_rows_kernel(const double *__restrict__ in_re,
                                                        const double *__restrict__ in_im, int64_t estride,
                                                        const double2 *__restrict__ tin, double2 *__restrict__ tout,
                                                        int n, int logn, int pass, int flip = 0, int64_t in_pstride = 0,
                                                        int64_t o_estride = 1, int64_t o_pstride = 0)
{
    extern __shared__ double2 srow[];
    const int row = blockIdx.x, h = n / 2;
    const int64_t plane = blockIdx.y, nn = (int64_t)n * n;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        double2 val;
        if (pass == 0) {
            const int rs = (row + h) % n;
            const int64_t src = ((int64_t)(flip ? n - 1 - rs : rs) * n + (c + h) % n) * estride + plane * in_pstride;
            val = make_double2(in_re[src], in_im ? in_im[src] : 0.0);
        } else {
            val = tin[plane * nn + (int64_t)row * n + c];
        }
        srow[bitrev((unsigned)c, logn)] = val;
    }
    // twiddle table tw[j] = exp(+2 pi i j / n), j < n/2, once per block (the butterflies of stage s use every
    // (n / 2^s)-th entry) instead of one sincospi per butterfly
    double2 *tw = srow + n;
    for (int j = threadIdx.x; j < n / 2; j += blockDim.x) {
        double ws, wc;
        sincospi(2.0 * (double)j / (double)n, &ws, &wc);
        tw[j] = make_double2(wc, ws);
    }
    __syncthreads();
    for (int s = 1; s <= logn; s++) {
        const int m = 1 << s, half = m >> 1, tstep = n >> s;
        for (int idx = threadIdx.x; idx < n / 2; idx += blockDim.x) {
            const int k = idx & (half - 1);
            const int j = ((idx >> (s - 1)) << s) + k;
            const double2 w = tw[k * tstep];                        // inverse transform: +i
            const double2 a = srow[j], b = srow[j + half];
            const double tr = w.x * b.x - w.y * b.y, ti = w.x * b.y + w.y * b.x;
            srow[j] = make_double2(a.x + tr, a.y + ti);
            srow[j + half] = make_double2(a.x - tr, a.y - ti);
        }
        __syncthreads();
    }
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
        if (pass == 0) tout[plane * nn + (int64_t)b * n + row] = srow[b];
        else tout[((int64_t)((b + h) % n) * n + (row + h) % n) * o_estride + plane * o_pstride] = srow[b];
    }
}

// ---- any length n: Bluestein ------------------------------------------------------------------------
//   X[b] = sum_c x[c] e^{+2 pi i b c / n},  b c = (b^2 + c^2 - (b - c)^2) / 2
//        = ch[b] * sum_c (x[c] ch[c]) conj(ch[b - c]),   ch[k] = e^{+i pi k^2 / n}
// i.e. a circular convolution of length m = 2^logm >= 2n - 1, done with the radix-2 transform above:
// conv = IFFT(FFT(a) FFT(b)), FFT(b) precomputed once per n (FB).  Same I/O conventions as ifft_rows_kernel.
__device__ __forceinline__ double2 chirp(int k, int n)
{
    const long long q = ((long long)k * k) % (2LL * n);        // k^2 mod 2n keeps the argument exact
    double s, c;
    sincospi((double)q / (double)n, &s, &c);
    return make_double2(c, s);
}
// Complex product with the contraction fixed (first product fused, second rounded), so that a result does not depend
// on which product nvcc decides to fuse at a given call site.
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(fma(a.x, b.x, -__dmul_rn(a.y, b.y)), fma(a.x, b.y, __dmul_rn(a.y, b.x)));
}

// in-place radix-2 DIT on bit-reversed input, kernel e^{+2 pi i jk/m}; tw[j] = e^{+2 pi i j/m}, j < m/2
__device__ __forceinline__ void fft_plus_smem(double2 *w, const double2 *tw, int m, int logm)
{
    for (int s = 1; s <= logm; s++) {
        const int half = 1 << (s - 1), tstep = m >> s;
        for (int idx = threadIdx.x; idx < m / 2; idx += blockDim.x) {
            const int k = idx & (half - 1);
            const int j = ((idx >> (s - 1)) << s) + k;
            const double2 t = cmul(tw[k * tstep], w[j + half]), a = w[j];
            w[j] = make_double2(a.x + t.x, a.y + t.y);
            w[j + half] = make_double2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
}

// FB = FFT+(b), b[k] = conj(ch[|k|]) for |k| < n (wrapped into [0, m)), 0 elsewhere.  One block.
__global__ void __launch_bounds__(512) bluestein_setup_kernel(double2 *__restrict__ FB, int n, int m, int logm)
{
    extern __shared__ double2 srow[];
    double2 *work = srow, *tw = srow + m;
    for (int j = threadIdx.x; j < m / 2; j += blockDim.x) {
        double ws, wc;
        sincospi(2.0 * (double)j / (double)m, &ws, &wc);
        tw[j] = make_double2(wc, ws);
    }
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        const int d = k < n ? k : (m - k < n ? m - k : -1);
        if (d >= 0) {
            const double2 c = chirp(d, n);
            v = make_double2(c.x, -c.y);
        }
        work[bitrev((unsigned)k, logm)] = v;
    }
    __syncthreads();
    fft_plus_smem(work, tw, m, logm);
    for (int k = threadIdx.x; k < m; k += blockDim.x) FB[k] = work[k];
}

__global__ void __launch_bounds__(512) bluestein_rows_kernel(const double *__restrict__ in_re,
                                                             const double *__restrict__ in_im, int64_t estride,
                                                             const double2 *__restrict__ tin, double2 *__restrict__ tout,
                                                             int n, int m, int logm, int pass,
                                                             const double2 *__restrict__ FB)
{
    extern __shared__ double2 srow[];
    double2 *work = srow, *tw = srow + m;
    const int row = blockIdx.x, h = n / 2;
    for (int j = threadIdx.x; j < m / 2; j += blockDim.x) {
        double ws, wc;
        sincospi(2.0 * (double)j / (double)m, &ws, &wc);
        tw[j] = make_double2(wc, ws);
    }
    for (int c = threadIdx.x; c < m; c += blockDim.x) {
        double2 val = make_double2(0.0, 0.0);
        if (c < n) {
            if (pass == 0) {
                const int64_t src = ((int64_t)((row + h) % n) * n + (c + h) % n) * estride;
                val = make_double2(in_re[src], in_im ? in_im[src] : 0.0);
            } else {
                val = tin[(int64_t)row * n + c];
            }
            val = cmul(val, chirp(c, n));
        }
        work[bitrev((unsigned)c, logm)] = val;
    }
    __syncthreads();
    fft_plus_smem(work, tw, m, logm);
    // z = conj(FFT(a) FB), permuted to bit-reversed order in place for the second transform
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const int r = (int)bitrev((unsigned)k, logm);
        if (k < r) {
            const double2 zk = cmul(work[k], FB[k]), zr = cmul(work[r], FB[r]);
            work[r] = make_double2(zk.x, -zk.y);
            work[k] = make_double2(zr.x, -zr.y);
        } else if (k == r) {
            const double2 zk = cmul(work[k], FB[k]);
            work[k] = make_double2(zk.x, -zk.y);
        }
    }
    __syncthreads();
    fft_plus_smem(work, tw, m, logm);               // conj(work) / m is the convolution
    const double inv_m = 1.0 / (double)m;
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
        const double2 y = make_double2(work[b].x * inv_m, -work[b].y * inv_m);
        const double2 X = cmul(y, chirp(b, n));
        if (pass == 0) tout[(int64_t)b * n + row] = X;
        else tout[(int64_t)((b + h) % n) * n + (row + h) % n] = X;
    }
}

// image[A, n-1-B, ch] = Re(Y_im[A,B]) / (Re(Y_conv[A,B]) / n^2)        (invert.py:74-79)
__global__ void __launch_bounds__(256) invert_combine_kernel(const double2 *__restrict__ yim,
                                                             const double2 *__restrict__ yconv, int n, int nch, int ch,
                                                             double *__restrict__ image)
{
    const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (q >= (int64_t)n * n) return;
    const int A = (int)(q / n), B = (int)(q % n);
    const double im = yim[q].x;                                   // ifft2 * imsize^2: the 1/n^2 cancels
    const double cv = yconv[q].x / ((double)n * (double)n);
    image[((int64_t)A * n + (n - 1 - B)) * nch + ch] = im / cv;
}

static int fft_attr()
{
    static bool attr = false;
    if (!attr) {
        PDSB_CUDA(cudaFuncSetAttribute(ifft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 24));
        attr = true;
    }
    return PDSB_OK;
}

// Y[(A n + B) nf + i] = fftshift( sum_{r,c} X_i[r, c] exp(+2 pi i (A r + B c) / n) ) over both axes, where
// X_i = ifftshift of plane i of the cube [n, n, nf] (rows reversed first when flip): the unnormalised
// inverse-sign 2-D transform of every channel at once, channel fastest in the output.  T: [nf, n, n] scratch.
int fft2_planes(const double *cube_dev, int n, int nf, int flip, double2 *T, double2 *Y)
{
    Context &c = ctx();
    int logn = 0;
    while ((1 << logn) < n) logn++;
    PDSB_CHECK(fft_attr());
    const int threads = n / 2 < 512 ? (n / 2 < 32 ? 32 : n / 2) : 512;
    const size_t smem = (size_t)n * sizeof(double2) * 3 / 2;          // row + twiddle table
    LaunchScope ls("fft2_planes");
    ifft_rows_kernel<<<dim3(n, nf), threads, smem, c.stream>>>(cube_dev, nullptr, nf, nullptr, T, n, logn, 0, flip, 1);
    ifft_rows_kernel<<<dim3(n, nf), threads, smem, c.stream>>>(nullptr, nullptr, 0, T, Y, n, logn, 1, 0, 0, nf, 1);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}


// ---- the galario path's forward transform: real planes -> half spectrum ----------------------------------------
// rfft2_planes computes what fft2_planes computes, for the columns B = (b + n/2) % n, b = 0 .. n/2, that galario's
// rfft2-based sampler reads (vis.cu: fft_sample_one), and stores them compactly:
//     Yh[(A (n/2 + 1) + b) nf + i] = Ysh_i[A][(b + n/2) % n]
// It does a quarter of the butterflies of the general routine and moves every byte in full 32-byte sectors:
//   * the planes are real, so two adjacent channels (adjacent doubles of the channel-fastest cube) go through ONE
//     complex row transform as z = x_p + i x_{p+1} and are separated afterwards,
//     A_p[b] = (Z[b] + conj Z[n-b]) / 2,  A_{p+1}[b] = (Z[b] - conj Z[n-b]) / (2 i);
//   * only b <= n/2 is kept (the other half is its conjugate), so the column pass transforms n/2 + 1 rows per plane;
//   * a block works on 2 image rows x PB plane pairs (row pass) or PB planes (column pass) at once, so that the
//     transposed stores of the row pass and the channel-fastest stores of the column pass are contiguous runs;
//   * two radix-2 stages per barrier (radix-4 butterflies in registers, same operations as the radix-2 routine),
//     twiddles from a table computed once per size.
__global__ void __launch_bounds__(256) fft_twiddle_kernel(double2 *__restrict__ tw, int n)
{
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n / 2) return;
    // Built from the first octant only, so that the table has the symmetries of the exponential exactly:
    // tw[j + n/4] = i tw[j] (the butterflies rely on it: one table read stands for two twiddles), tw[n/4 - j] = swap(tw[j]).
    int jq = j;
    const bool rot = 4 * jq > n;
    if (rot) jq -= n / 4;
    const bool swp = 8 * jq > n;
    const int jb = swp ? n / 4 - jq : jq;
    double ws, wc;
    sincospi(2.0 * (double)jb / (double)n, &ws, &wc);
    if (swp) {
        const double t = wc;
        wc = ws;
        ws = t;
    }
    tw[j] = rot ? make_double2(-ws, wc) : make_double2(wc, ws);
}

// in-place transforms of nfft rows of length n (bit-reversed input) held at w + f * n, kernel e^{+2 pi i jk/n};
// tpf threads per row, the block's threads are [row][tpf]
// Shared-memory layout of one transform: element i at slot fft_pad(i) = i + i / 4 (row length fft_padlen(n)).  The
// radix-4 butterflies of the first stages touch 4 consecutive / 4-strided elements per thread: without the padding a
// quarter-warp's 16-byte accesses fall on 2 - 4 of the 8 bank groups (4-way / 2-way conflicts), with it they spread
// over all 8; unit-stride accesses pay one extra wavefront in four.
__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 2); }
__host__ __device__ __forceinline__ int fft_padlen(int n) { return n + (n >> 2); }

__device__ __forceinline__ void fft_plus_rows_r4(double2 *w, const double2 *tw, int n, int logn, int nfft, int tpf)
{
    const int f = threadIdx.x / tpf, tl = threadIdx.x % tpf;
    double2 *x = w + (size_t)f * fft_padlen(n);
    int s = 1;
    for (; s + 1 <= logn; s += 2) {
        const int half = 1 << (s - 1), ts2 = n >> (s + 1);
        if (f < nfft)
            for (int idx = tl; idx < n / 4; idx += tpf) {
                const int k = idx & (half - 1);
                const int j = ((idx >> (s - 1)) << (s + 1)) + k;
                const int p0 = fft_pad(j), p1 = fft_pad(j + half), p2 = fft_pad(j + 2 * half), p3 = fft_pad(j + 3 * half);
                double2 a0 = x[p0], a1 = x[p1], a2 = x[p2], a3 = x[p3];
                // one table read per butterfly: w2a = e^{2 pi i k / 4 half}; the inner stage's twiddle is its square
                // (k ts1 = 2 k ts2), the second outer one is i w2a ((k + half) ts2 = k ts2 + n / 4).
                // First pass: half = 1, k = 0: 1, 1, i - no table read.
                const double2 w2a = s == 1 ? make_double2(1.0, 0.0) : tw[k * ts2];
                const double2 w1 = make_double2(w2a.x * w2a.x - w2a.y * w2a.y, 2.0 * w2a.x * w2a.y);
                const double2 w2b = make_double2(-w2a.y, w2a.x);
                double2 t = cmul(w1, a1);
                a1 = make_double2(a0.x - t.x, a0.y - t.y);
                a0 = make_double2(a0.x + t.x, a0.y + t.y);
                t = cmul(w1, a3);
                a3 = make_double2(a2.x - t.x, a2.y - t.y);
                a2 = make_double2(a2.x + t.x, a2.y + t.y);
                t = cmul(w2a, a2);
                x[p0] = make_double2(a0.x + t.x, a0.y + t.y);
                x[p2] = make_double2(a0.x - t.x, a0.y - t.y);
                t = cmul(w2b, a3);
                x[p1] = make_double2(a1.x + t.x, a1.y + t.y);
                x[p3] = make_double2(a1.x - t.x, a1.y - t.y);
            }
        __syncthreads();
    }
    if (s <= logn) {                                       // odd number of stages: one radix-2 stage left
        const int half = 1 << (s - 1), tstep = n >> s;
        if (f < nfft)
            for (int idx = tl; idx < n / 2; idx += tpf) {
                const int k = idx & (half - 1);
                const int j = ((idx >> (s - 1)) << s) + k;
                const int p0 = fft_pad(j), p1 = fft_pad(j + half);
                const double2 t = cmul(tw[k * tstep], x[p1]), a = x[p0];
                x[p0] = make_double2(a.x + t.x, a.y + t.y);
                x[p1] = make_double2(a.x - t.x, a.y - t.y);
            }
        __syncthreads();
    }
}

// flags[p] = 1 when plane p of the channel-fastest cube holds a non-zero value.  A channel pair shares one complex
// transform; its separation returns the rounding residue of the partner (1e-16 of the partner's values) for an
// all-zero channel unless the zero is known: with the flags an empty channel transforms to exact zeros, like it
// does in the direct-sum kernels.
// (the grid is a multiple of nf blocks, so the grid stride is a multiple of nf and a thread stays on one channel)
__global__ void __launch_bounds__(256) plane_nonzero_kernel(const double *__restrict__ cube, int64_t npix, int nf,
                                                            int *__restrict__ flags)
{
    const int64_t total = npix * nf, i0 = (int64_t)blockIdx.x * 256 + threadIdx.x, stride = (int64_t)gridDim.x * 256;
    bool any = false;
    int64_t i = i0;
    for (; i + 3 * stride < total; i += 4 * stride) {    // four loads in flight per thread
        const double a = cube[i], b = cube[i + stride], c = cube[i + 2 * stride], d = cube[i + 3 * stride];
        any |= a != 0.0 || b != 0.0 || c != 0.0 || d != 0.0;
    }
    for (; i < total; i += stride) any |= cube[i] != 0.0;
    if (any) flags[(int)(i0 % nf)] = 1;                  // benign race: every writer stores 1
}

// row pass: block = rows (rho, rho + 1), rho = 2 blockIdx.x, of the (row-flipped) source planes x plane pairs
// [PB blockIdx.y, ...) -> T[plane][b][r'], b <= n/2  (T: [nf][n/2 + 1][n]).
// The transform length n may exceed the source sides nsy x nsx (zero padding for the NUFFT path, nsy = nsx = n otherwise):
// the source pixel with centred coordinates (R, C) = (rho - nsy/2, gamma - nsx/2) sits at (R mod n, C mod n),
// i.e. the image centre is the origin of the transform (the ifftshift of the unpadded case); corr (may be null)
// multiplies pixel (rho, gamma) by corr[rho] corr[gamma] (the NUFFT's deapodisation).  Rows r' that hold no source row
// are never written: the column pass knows they are zero.
__global__ void __launch_bounds__(512) rfft_rows_kernel(const double *__restrict__ cube, const double2 *__restrict__ twg,
                                                        double2 *__restrict__ T, int n, int logn, int nf, int flip,
                                                        int PB, int tpf, int nsy, int nsx, const double *__restrict__ corr_y,
                                                        const double *__restrict__ corr_x, const int *__restrict__ nonzero)
{
    extern __shared__ double2 srow[];
    const int h = n / 2, hsy = nsy / 2, hsx = nsx / 2, npair = (nf + 1) / 2;
    const int pair0 = blockIdx.y * PB;
    const int pb_n = npair - pair0 < PB ? npair - pair0 : PB;          // pairs this block really has
    const int nfft = 2 * pb_n;                                          // [row 0..1][pair]
    const int rho0 = 2 * blockIdx.x;
    const int npad = fft_padlen(n);
    double2 *tw = srow + (size_t)2 * PB * npad;
    for (int j = threadIdx.x; j < h; j += blockDim.x) tw[j] = twg[j];
    // loads: pair fastest (adjacent doubles of the cube), then column, then row.  n is a power of two; the pair count of
    // a block is 1 .. 4: a thread keeps its pair and strides over the columns (no division in the loop) when the
    // block size allows it
    // A thread that stores slot j of the (bit-reversed) work row fetches column bitrev(j): the stores are linear in
    // shared memory, and the column order of the fetches is free (columns are nf * 8 bytes apart in the cube anyway).
    auto load_one = [&](int rb, int pb, int j) {
        const int c = (int)bitrev((unsigned)j, logn);
        const int rho = rho0 + rb, gamma = (c < h ? c : c - n) + hsx;
        double re = 0.0, im = 0.0;
        if (gamma >= 0 && gamma < nsx && rho < nsy) {
            const int64_t src = ((int64_t)(flip ? nsy - 1 - rho : rho) * nsx + gamma) * nf + 2 * (pair0 + pb);
            const double f = corr_y ? corr_y[rho] * corr_x[gamma] : 1.0;
            re = cube[src] * f;
            if (2 * (pair0 + pb) + 1 < nf) im = cube[src + 1] * f;
        }
        srow[(size_t)(rb * pb_n + pb) * npad + fft_pad(j)] = make_double2(re, im);
    };
    if (blockDim.x % pb_n == 0) {
        const int pb = threadIdx.x % pb_n, cstep = blockDim.x / pb_n;
        for (int rb = 0; rb < 2; rb++)
            for (int c = threadIdx.x / pb_n; c < n; c += cstep) load_one(rb, pb, c);
    } else {
        for (int idx = threadIdx.x; idx < nfft * n; idx += blockDim.x)
            load_one(idx / (pb_n * n), idx % pb_n, (idx / pb_n) & (n - 1));
    }
    __syncthreads();
    fft_plus_rows_r4(srow, tw, n, logn, nfft, tpf);
    // separate the two planes of every pair; stores: row fastest (adjacent double2 of T), then b, then pair
    const int64_t ps = (int64_t)(h + 1) * n;                            // plane stride of T
    for (int pb = 0; pb < pb_n; pb++)
        for (int j = threadIdx.x; j < 2 * (h + 1); j += blockDim.x) {
            const int rb = j & 1, b = j >> 1;
            const double2 *x = srow + (size_t)(rb * pb_n + pb) * npad;
            const double2 z = x[fft_pad(b)], zc = x[fft_pad((n - b) & (n - 1))];
            const int plane = 2 * (pair0 + pb);
            const int64_t o = (int64_t)plane * ps + (int64_t)b * n + ((rho0 + rb - hsy + n) & (n - 1));      // adjacent for even nsy / 2
            const bool nz0 = nonzero[plane] != 0, nz1 = plane + 1 < nf && nonzero[plane + 1] != 0;
            // an empty channel is exactly zero, and its partner is then the whole transform (Z, or Z / i)
            T[o] = !nz0 ? make_double2(0.0, 0.0) : !nz1 ? z : make_double2(0.5 * (z.x + zc.x), 0.5 * (z.y - zc.y));
            if (plane + 1 < nf)
                T[o + ps] = !nz1 ? make_double2(0.0, 0.0)
                                 : !nz0 ? make_double2(z.y, -z.x) : make_double2(0.5 * (z.y + zc.y), -0.5 * (z.x - zc.x));
        }
}

// column pass: block = row b of T for planes [PB blockIdx.y, ...) -> Yh[((a + h) % n (h + 1) + b) nf + plane];
// entries r' of a row of T outside the source rows (centred coordinate not in [-nsrc/2, nsrc/2)) read as zero
__global__ void __launch_bounds__(512) rfft_cols_kernel(const double2 *__restrict__ T, const double2 *__restrict__ twg,
                                                        double2 *__restrict__ Yh, int n, int logn, int nf, int PB, int tpf,
                                                        int nsy)
{
    extern __shared__ double2 srow[];
    const int h = n / 2, hs = nsy / 2, b = blockIdx.x;
    const int plane0 = blockIdx.y * PB;
    const int nfft = nf - plane0 < PB ? nf - plane0 : PB;
    const int npad = fft_padlen(n);
    double2 *tw = srow + (size_t)PB * npad;
    for (int j = threadIdx.x; j < h; j += blockDim.x) tw[j] = twg[j];
    const int64_t ps = (int64_t)(h + 1) * n;
    // slot j of the work row <- entry bitrev(j) of the row of T (linear shared-memory stores; the two halves of a
    // 32-byte sector of T are fetched by two threads of the same block)
    for (int idx = threadIdx.x; idx < nfft * n; idx += blockDim.x) {
        const int j = idx & (n - 1), pl = idx >> logn;
        const int r = (int)bitrev((unsigned)j, logn);
        const int R = r < h ? r : r - n;
        double2 val = make_double2(0.0, 0.0);
        if (R >= -hs && R < nsy - hs) val = T[(int64_t)(plane0 + pl) * ps + (int64_t)b * n + r];
        srow[(size_t)pl * npad + fft_pad(j)] = val;
    }
    __syncthreads();
    fft_plus_rows_r4(srow, tw, n, logn, nfft, tpf);
    if (blockDim.x % nfft == 0) {                        // plane fastest (contiguous channels), no division in the loop
        const int pl = threadIdx.x % nfft, astep = blockDim.x / nfft;
        for (int a = threadIdx.x / nfft; a < n; a += astep)
            Yh[((int64_t)((a + h) & (n - 1)) * (h + 1) + b) * nf + plane0 + pl] = srow[(size_t)pl * npad + fft_pad(a)];
    } else {
        for (int idx = threadIdx.x; idx < nfft * n; idx += blockDim.x) {
            const int pl = idx % nfft, a = idx / nfft;
            Yh[((int64_t)((a + h) & (n - 1)) * (h + 1) + b) * nf + plane0 + pl] = srow[(size_t)pl * npad + fft_pad(a)];
        }
    }
}

// ---- the same two passes with the transforms held in registers (fft_r16.cuh), 256 <= n <= 2048 ----
// A block of NT threads works on F = 16 NT / n transforms (n / 16 threads each): 2 image rows x F / 2 channel pairs
// (row pass), F planes (column pass); blocks that share 32-byte sectors of the cube / of Yh (neighbouring channel groups)
// are neighbours in launch order (blockIdx.x), so the halves meet in the L2.  The first pass reads the cube / T straight into registers (16 independent loads
// in flight per thread), the data goes through shared memory once per later pass, and the results leave through the
// shared row once more so that the stores are the same contiguous runs as in the kernels above.
// Rows of the shared buffer are rowlen(n) + 8 / F slots apart, so that the plane-fastest read-out is conflict-free too.
__host__ __device__ __forceinline__ int r16_rowstride(int n, int F) { return r16::rowlen(n) + (F >= 8 ? 1 : 8 / F); }

// all passes but the first read phase: v holds the transform's input (element t + k n/16 in v[k]); on return the
// spectrum is in x (natural order, slot()-padded).  Every thread of the block must call this (barriers inside).
template <int RL>
__device__ __forceinline__ void r16_transform(double2 *x, const double2 *__restrict__ tw, int n, int logn, int t, double2 *v)
{
    int Ns = 1;
    const int full = r16::full_passes(logn);
    for (int p = 0; p < full; p++) {
        if (p > 0) {
            r16::pass_load<16>(x, n, t, v);
            __syncthreads();
        }
        r16::pass_compute<16>(tw, n, Ns, t, v);
        r16::pass_store<16>(x, n, Ns, t, v);
        __syncthreads();
        Ns <<= 4;
    }
    r16::pass_load<RL>(x, n, t, v);
    __syncthreads();
    r16::pass_compute<RL>(tw, n, Ns, t, v);
    r16::pass_store<RL>(x, n, Ns, t, v);
    __syncthreads();
}

// Both kernels are persistent: a block takes the work items blockIdx.x, blockIdx.x + gridDim.x, ... and issues the 16 loads
// of its NEXT item right after the transform of the current one (the registers are free then: the spectrum sits in shared
// memory), so that they are in flight while the current item is separated / stored.
template <int RL, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) rfft_rows16_kernel(const double *__restrict__ cube, const double2 *__restrict__ twg,
                                                             double2 *__restrict__ T, int n, int logn, int nf, int flip,
                                                             int nsy, int nsx, const double *__restrict__ corr_y,
                                                             const double *__restrict__ corr_x,
                                                             const int *__restrict__ nonzero)
{
    extern __shared__ double2 srow[];
    const int Tn = n >> 4, F = NT / Tn, PBp = F >> 1;
    const int h = n / 2, hsy = nsy / 2, hsx = nsx / 2, npair = (nf + 1) / 2;
    const int npg = (npair + PBp - 1) / PBp;                            // pair groups; item = pair group + npg * row pair
    const int nitems = npg * (nsy / 2);
    const int f = threadIdx.x / Tn, t = threadIdx.x % Tn;
    const int rb = f / PBp, pb = f % PBp;                               // transform f = (row rb, pair pb)
    const int rs = r16_rowstride(n, F);
    double2 *x = srow + (size_t)f * rs;
    const bool aligned = (nf & 1) == 0 && (reinterpret_cast<uintptr_t>(cube) & 15) == 0;
    double2 v[16];
    auto load_item = [&](int item) {
        const int pair0 = (item % npg) * PBp, rho = 2 * (item / npg) + rb;
        const int plane = 2 * (pair0 + pb);
        const bool live = pair0 + pb < npair && rho < nsy;
        const bool second = plane + 1 < nf;
        const bool vec = second && aligned;
        const int64_t rowbase = (int64_t)(flip ? nsy - 1 - rho : rho) * nsx;
        const double fy = live && corr_y ? corr_y[rho] : 1.0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int c = t + k * Tn;
            const int gamma = (c < h ? c : c - n) + hsx;
            double2 val = make_double2(0.0, 0.0);
            if (live && gamma >= 0 && gamma < nsx) {
                const int64_t src = (rowbase + gamma) * nf + plane;
                if (vec) val = *reinterpret_cast<const double2 *>(cube + src);
                else {
                    val.x = cube[src];
                    if (second) val.y = cube[src + 1];
                }
                if (corr_y) {
                    const double fc = fy * corr_x[gamma];
                    val.x *= fc;
                    val.y *= fc;
                }
            }
            v[k] = val;
        }
    };
    const int64_t ps = (int64_t)(h + 1) * n;                            // plane stride of T
    if ((int)blockIdx.x < nitems) load_item(blockIdx.x);
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        r16_transform<RL>(x, twg, n, logn, t, v);
        if (item + (int)gridDim.x < nitems) load_item(item + gridDim.x);
        // separate the two planes of every pair; stores: row fastest (adjacent double2 of T), then b, then pair
        const int pair0 = (item % npg) * PBp, rho0 = 2 * (item / npg);
        const int pb_n = npair - pair0 < PBp ? npair - pair0 : PBp;     // pairs this item really has
        for (int q = 0; q < pb_n; q++)
            for (int j = threadIdx.x; j < 2 * (h + 1); j += NT) {
                const int r2 = j & 1, b = j >> 1;
                if (rho0 + r2 >= nsy) continue;
                const double2 *xr = srow + (size_t)(r2 * PBp + q) * rs;
                const double2 z = xr[r16::slot(b)], zc = xr[r16::slot((n - b) & (n - 1))];
                const int plane = 2 * (pair0 + q);
                const int64_t o = (int64_t)plane * ps + (int64_t)b * n + ((rho0 + r2 - hsy + n) & (n - 1));
                const bool nz0 = nonzero[plane] != 0, nz1 = plane + 1 < nf && nonzero[plane + 1] != 0;
                T[o] = !nz0 ? make_double2(0.0, 0.0) : !nz1 ? z : make_double2(0.5 * (z.x + zc.x), 0.5 * (z.y - zc.y));
                if (plane + 1 < nf)
                    T[o + ps] = !nz1 ? make_double2(0.0, 0.0)
                                     : !nz0 ? make_double2(z.y, -z.x) : make_double2(0.5 * (z.y + zc.y), -0.5 * (z.x - zc.x));
            }
        __syncthreads();                                 // the shared rows are free for the next item
    }
}

template <int RL, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) rfft_cols16_kernel(const double2 *__restrict__ T, const double2 *__restrict__ twg,
                                                             double2 *__restrict__ Yh, int n, int logn, int nf, int nsy)
{
    extern __shared__ double2 srow[];
    const int Tn = n >> 4, F = NT / Tn;
    const int h = n / 2, hs = nsy / 2;
    const int ngr = (nf + F - 1) / F;                                   // plane groups; item = plane group + ngr * b
    const int nitems = ngr * (h + 1);
    const int f = threadIdx.x / Tn, t = threadIdx.x % Tn;
    const int rs = r16_rowstride(n, F);
    double2 *x = srow + (size_t)f * rs;
    const int64_t ps = (int64_t)(h + 1) * n;
    double2 v[16];
    auto load_item = [&](int item) {
        const int plane = (item % ngr) * F + f, b = item / ngr;
        const double2 *src = T + (int64_t)plane * ps + (int64_t)b * n;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int r = t + k * Tn;
            const int R = r < h ? r : r - n;
            v[k] = (plane < nf && R >= -hs && R < nsy - hs) ? src[r] : make_double2(0.0, 0.0);
        }
    };
    if ((int)blockIdx.x < nitems) load_item(blockIdx.x);
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        r16_transform<RL>(x, twg, n, logn, t, v);
        if (item + (int)gridDim.x < nitems) load_item(item + gridDim.x);
        const int plane0 = (item % ngr) * F, b = item / ngr;
        const int nfft = nf - plane0 < F ? nf - plane0 : F;
        if (NT % nfft == 0) {                            // plane fastest (contiguous channels), no division in the loop
            const int pl = threadIdx.x % nfft, astep = NT / nfft;
            for (int a = threadIdx.x / nfft; a < n; a += astep)
                Yh[((int64_t)((a + h) & (n - 1)) * (h + 1) + b) * nf + plane0 + pl] = srow[(size_t)pl * rs + r16::slot(a)];
        } else {
            for (int idx = threadIdx.x; idx < nfft * n; idx += NT) {
                const int pl = idx % nfft, a = idx / nfft;
                Yh[((int64_t)((a + h) & (n - 1)) * (h + 1) + b) * nf + plane0 + pl] = srow[(size_t)pl * rs + r16::slot(a)];
            }
        }
        __syncthreads();                                 // the shared rows are free for the next item
    }
}

template <int RL, int NT>
static int launch_rfft16_nt(const double *cube_dev, const double2 *tw, double2 *T, double2 *Yh, int n, int logn, int nf, int flip,
                            int nsy, int nsx, const double *corr_y, const double *corr_x, const int *flags)
{
    Context &c = ctx();
    const int F = 16 * NT / n, npair = (nf + 1) / 2;
    const size_t smem = (size_t)F * r16_rowstride(n, F) * sizeof(double2);
    static bool attr = false;
    if (!attr) {
        PDSB_CUDA(cudaFuncSetAttribute(rfft_rows16_kernel<RL, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        PDSB_CUDA(cudaFuncSetAttribute(rfft_cols16_kernel<RL, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr = true;
    }
    // persistent blocks (one wave, next-item loads under the current item's stores) once there are at least four items per
    // block to balance; fewer items: one block each (measured: C3 transforms 4 - 5 % faster, C2's single plane 4 % slower)
    const int wave = c.sm_count * (512 / NT);
    const int items_r = ceil_div(npair, F / 2) * (nsy / 2), items_c = ceil_div(nf, F) * (n / 2 + 1);
    rfft_rows16_kernel<RL, NT><<<items_r >= 4 * wave ? wave : items_r, NT, smem, c.stream>>>(cube_dev, tw, T, n, logn, nf, flip,
                                                                                           nsy, nsx, corr_y, corr_x, flags);
    rfft_cols16_kernel<RL, NT><<<items_c >= 4 * wave ? wave : items_c, NT, smem, c.stream>>>(T, tw, Yh, n, logn, nf, nsy);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

// threads per block: 128 (four blocks per SM in different phases of load / transform / store) while that still gives
// the row pass its two rows (n <= 1024), else 256
template <int RL>
static int launch_rfft16(const double *cube_dev, const double2 *tw, double2 *T, double2 *Yh, int n, int logn, int nf, int flip,
                         int nsy, int nsx, const double *corr_y, const double *corr_x, const int *flags)
{
    static const char *env = getenv("PDSB_FFT_NT");
    const int nt = env ? atoi(env) : 128;
    if (n <= 1024 && nt == 128)
        return launch_rfft16_nt<RL, 128>(cube_dev, tw, T, Yh, n, logn, nf, flip, nsy, nsx, corr_y, corr_x, flags);
    return launch_rfft16_nt<RL, 256>(cube_dev, tw, T, Yh, n, logn, nf, flip, nsy, nsx, corr_y, corr_x, flags);
}

// T: [nf][n/2 + 1][n] scratch, Yh: [n][n/2 + 1][nf].  nsy x nsx source planes (even sides) transformed at length
// n >= max(nsy, nsx), a power of two (n = nsy = nsx: no padding); corr_y [nsy], corr_x [nsx]: per-axis pixel factors, or null.
int rfft2_planes_padded(const double *cube_dev, int nsy, int nsx, int n, int nf, int flip, const double *corr_y,
                        const double *corr_x, double2 *T, double2 *Yh)
{
    Context &c = ctx();
    int logn = 0;
    while ((1 << logn) < n) logn++;
    PDSB_REQUIRE(n >= 2 && (1 << logn) == n && n <= 4096, "rfft2_planes: side must be a power of two in [2, 4096]");
    PDSB_REQUIRE(nsy >= 2 && nsx >= 2 && nsy <= n && nsx <= n && nsy % 2 == 0 && nsx % 2 == 0,
                 "rfft2_planes: source sides must be even and <= the transform length");
    PDSB_REQUIRE((corr_y == nullptr) == (corr_x == nullptr), "rfft2_planes: both factor tables or none");
    static bool attr = false;
    if (!attr) {
        PDSB_CUDA(cudaFuncSetAttribute(rfft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        PDSB_CUDA(cudaFuncSetAttribute(rfft_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    PDSB_CHECK(c.fft_tw.ensure((size_t)(4096 / 2 + 1) * sizeof(double2)));
    if (c.fft_tw_n != n) {
        LaunchScope ls("fft_twiddle");
        fft_twiddle_kernel<<<ceil_div(n / 2, 256), 256, 0, c.stream>>>(c.fft_tw.as<double2>(), n);
        PDSB_CUDA(cudaGetLastError());
        c.fft_tw_n = n;
    }
    const double2 *tw = c.fft_tw.as<double2>();
    const int tpf = n / 4 < 1 ? 1 : (n / 4 > 256 ? 256 : n / 4);       // threads per transform
    const int npair = (nf + 1) / 2;
    // rows: 2 rows x PB pairs per block; columns: PB planes per block; <= 512 threads, shared memory <= 192 KB
    auto fit = [&](int per_unit, int have) {
        int pb = 512 / (per_unit * tpf);
        pb = pb < 1 ? 1 : pb > 4 ? 4 : pb;
        while (pb > 1 && (size_t)per_unit * pb * fft_padlen(n) * sizeof(double2) > 160 * 1024) pb--;
        return pb > have ? have : pb;
    };
    const int pb0 = fit(2, npair), pb1 = fit(1, nf);
    const int th0 = std::max(32, 2 * pb0 * tpf), th1 = std::max(32, pb1 * tpf);
    const size_t sm0 = ((size_t)2 * pb0 * fft_padlen(n) + n / 2) * sizeof(double2),
                 sm1 = ((size_t)pb1 * fft_padlen(n) + n / 2) * sizeof(double2);
    LaunchScope ls("rfft2_planes");
    PDSB_CHECK(c.fft_flags.ensure((size_t)nf * sizeof(int)));
    PDSB_CUDA(cudaMemsetAsync(c.fft_flags.ptr, 0, (size_t)nf * sizeof(int), c.stream));
    plane_nonzero_kernel<<<nf * ceil_div(c.sm_count * 8, nf), 256, 0, c.stream>>>(cube_dev, (int64_t)nsy * nsx, nf,
                                                                                  c.fft_flags.as<int>());
    if (n >= 256 && n <= 2048 && !getenv("PDSB_FFT_SMEM_PASSES")) {       // register-resident passes (fft_r16.cuh)
        const int *fl = c.fft_flags.as<int>();
        switch (r16::last_radix(logn)) {
        case 2: return launch_rfft16<2>(cube_dev, tw, T, Yh, n, logn, nf, flip, nsy, nsx, corr_y, corr_x, fl);
        case 4: return launch_rfft16<4>(cube_dev, tw, T, Yh, n, logn, nf, flip, nsy, nsx, corr_y, corr_x, fl);
        case 8: return launch_rfft16<8>(cube_dev, tw, T, Yh, n, logn, nf, flip, nsy, nsx, corr_y, corr_x, fl);
        default: return launch_rfft16<16>(cube_dev, tw, T, Yh, n, logn, nf, flip, nsy, nsx, corr_y, corr_x, fl);
        }
    }
    rfft_rows_kernel<<<dim3(nsy / 2, ceil_div(npair, pb0)), th0, sm0, c.stream>>>(cube_dev, tw, T, n, logn, nf, flip, pb0,
                                                                                  tpf, nsy, nsx, corr_y, corr_x,
                                                                                  c.fft_flags.as<int>());
    rfft_cols_kernel<<<dim3(n / 2 + 1, ceil_div(nf, pb1)), th1, sm1, c.stream>>>(T, tw, Yh, n, logn, nf, pb1, tpf, nsy);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

int rfft2_planes(const double *cube_dev, int n, int nf, int flip, double2 *T, double2 *Yh)
{
    return rfft2_planes_padded(cube_dev, n, n, n, nf, flip, nullptr, nullptr, T, Yh);
}

}  // namespace pdsb

using namespace pdsb;

extern "C" int pdsb_invert_image(const double *g_real, const double *g_imag, const double *conv, int imsize, int nch,
                                 int kind, double *image_out)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(g_real && g_imag && conv && image_out, "arrays");
    PDSB_REQUIRE(imsize >= 2 && imsize <= 4096, "imsize must be in [2, 4096]");
    PDSB_REQUIRE(nch >= 1, "nch");
    const int n = imsize;
    int logn = 0;
    while ((1 << logn) < n) logn++;
    const int64_t nn = (int64_t)n * n;
    const double *dre = g_real, *dim = g_imag, *dconv = conv;
    double *dout = image_out;
    if (kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_a.ensure((size_t)(2 * nn * nch + nn) * sizeof(double)));
        double *p = c.stage_a.as<double>();
        PDSB_CHECK(copy_h2d(p, g_real, (size_t)nn * nch * sizeof(double)));
        dre = p;
        p += nn * nch;
        PDSB_CHECK(copy_h2d(p, g_imag, (size_t)nn * nch * sizeof(double)));
        dim = p;
        p += nn * nch;
        PDSB_CHECK(copy_h2d(p, conv, (size_t)nn * sizeof(double)));
        dconv = p;
        PDSB_CHECK(c.stage_b.ensure((size_t)nn * nch * sizeof(double)));
        dout = c.stage_b.as<double>();
    }
    PDSB_CHECK(c.stage_c.ensure((size_t)3 * nn * sizeof(double2)));
    double2 *T = c.stage_c.as<double2>(), *Yc = T + nn, *Yi = Yc + nn;
    const int threads = n / 2 < 512 ? (n / 2 < 32 ? 32 : n / 2) : 512;
    const size_t smem = (size_t)n * sizeof(double2) * 3 / 2;          // row + twiddle table
    PDSB_CHECK(fft_attr());
    const bool pow2 = (n & (n - 1)) == 0;
    int m = 1, logm = 0;
    const double2 *FB = nullptr;
    if (!pow2) {
        while (m < 2 * n - 1) { m <<= 1; logm++; }
        static bool battr = false;
        if (!battr) {
            PDSB_CUDA(cudaFuncSetAttribute(bluestein_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 24));
            PDSB_CUDA(cudaFuncSetAttribute(bluestein_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 24));
            battr = true;
        }
        PDSB_CHECK(c.fft_fb.ensure((size_t)m * sizeof(double2)));
        if (c.fft_fb_n != n) {
            LaunchScope ls("bluestein_setup");
            bluestein_setup_kernel<<<1, 512, (size_t)m * 24, c.stream>>>(c.fft_fb.as<double2>(), n, m, logm);
            PDSB_CUDA(cudaGetLastError());
            c.fft_fb_n = n;
        }
        FB = c.fft_fb.as<double2>();
    }
    auto ifft2 = [&](const double *re, const double *im, int64_t estride, double2 *Y) -> int {
        LaunchScope ls("invert_ifft2");
        if (pow2) {
            ifft_rows_kernel<<<n, threads, smem, c.stream>>>(re, im, estride, nullptr, T, n, logn, 0);
            ifft_rows_kernel<<<n, threads, smem, c.stream>>>(nullptr, nullptr, 0, T, Y, n, logn, 1);
        } else {
            bluestein_rows_kernel<<<n, 512, (size_t)m * 24, c.stream>>>(re, im, estride, nullptr, T, n, m, logm, 0, FB);
            bluestein_rows_kernel<<<n, 512, (size_t)m * 24, c.stream>>>(nullptr, nullptr, 0, T, Y, n, m, logm, 1, FB);
        }
        PDSB_CUDA(cudaGetLastError());
        return PDSB_OK;
    };
    PDSB_CHECK(ifft2(dconv, nullptr, 1, Yc));
    for (int ch = 0; ch < nch; ch++) {
        PDSB_CHECK(ifft2(dre + ch, dim + ch, nch, Yi));          // gridded maps are [G*G, nch]
        LaunchScope ls("invert_combine");
        invert_combine_kernel<<<ceil_div(nn, 256), 256, 0, c.stream>>>(Yi, Yc, n, nch, ch, dout);
        PDSB_CUDA(cudaGetLastError());
    }
    if (kind == PDSB_HOST) {
        PDSB_CHECK(copy_d2h(image_out, dout, (size_t)nn * nch * sizeof(double)));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}
