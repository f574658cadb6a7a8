// Convolutional gridding: grid() / freqcorrect() of the reference
// (pdspy/interferometry/libinterferometry.pyx:313-541, 587-608), re-designed for the GPU.
//
// The reference is one serial loop nest over (k, n, l, m) doing fp64 read-modify-writes.
// Two device strategies:
//
//  deterministic: every contribution (k, n, footprint slot) gets a key = its target cell and an
//      id that increases in the reference's loop order; a hand-written STABLE LSD radix sort
//      (8-bit digits, warp match_any ranking) groups contributions by cell while keeping the
//      reference's order inside each cell; one thread per cell then adds its run sequentially
//      with non-contracted fp64 multiplies/adds.  For the pillbox kernel (weights 0 or 1) the
//      gridded real/imag/weight maps are bit-identical to the reference.
//  fast: contributions are added with fp64 atomics (RED.ADD.F64); visibilities are processed in
//      home-cell order so that concurrent atomics hit neighbouring L2 lines.
//
// Index maps reproduce numpy's left-to-right fp64 arithmetic and its float64->uint32 cast
// (truncation through int64, wrap mod 2^32, NaN -> 0) exactly: __dmul_rn/__ddiv_rn/__dadd_rn keep
// the compiler from contracting the expression into FMAs.
#include "common.cuh"
#include <algorithm>
#include <cmath>

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
};

// :351-353 weights clamp/zeroing, :388-403 index maps, :421-423 good mask
__global__ void __launch_bounds__(256) grid_prep_kernel(GridParams P, unsigned long long *n_outside)
{
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t total = P.nuv * P.nf;
    if (idx >= total) return;
    const int64_t k = idx / P.nf;
    const int n = (int)(idx % P.nf);
    double w = P.w_in[idx];
    w = w < 0 ? 0.0 : w;
    if (P.re[idx] == 0 && P.im[idx] == 0) w = 0.0;
    P.w[idx] = w;
    const double f = P.freq[n];
    const double xi = __dadd_rn(__ddiv_rn(__dmul_rn(__dmul_rn(P.u[k], f), P.inv_freq), P.binsize), P.half);
    const double xj = __dadd_rn(__ddiv_rn(__dmul_rn(__dmul_rn(P.v[k], f), P.inv_freq), P.binsize), P.half);
    const uint32_t i = np_f64_to_u32(xi), j = np_f64_to_u32(xj);
    P.gi[idx] = i;
    P.gj[idx] = j;
    const bool good = i < (uint32_t)P.G && j < (uint32_t)P.G;
    P.good[idx] = good ? 1 : 0;
    if (!good) atomicAdd(n_outside, 1ull);
}

// Footprint of contribution slot f of visibility idx.  lo/hi half-widths, side = lo+hi+1.
// Returns false when the slot is clipped away (range(lmin,lmax) x range(mmin,mmax) of :493-507).
__device__ __forceinline__ bool slot_cell(uint32_t i, uint32_t j, int f, uint32_t lo, uint32_t hi, int G,
                                          uint32_t *l, uint32_t *m)
{
    const int side = (int)(lo + hi + 1);
    const int dl = f / side, dm = f % side;
    const long long ll = (long long)j - lo + dl, mm = (long long)i - lo + dm;
    if (ll < 0 || mm < 0 || ll >= G || mm >= G) return false;
    *l = (uint32_t)ll;
    *m = (uint32_t)mm;
    return true;
}

__device__ __forceinline__ double conv_value(const GridParams &P, int64_t k, int n, uint32_t l, uint32_t m)
{
    // convolve_func((u[k]*freq[n]*inv_freq - new_u[l,m])*inv_binsize, (v...-new_v[l,m])*inv_binsize) :509-511
    const double us = __dmul_rn(__dmul_rn(P.u[k], P.freq[n]), P.inv_freq);
    const double vs = __dmul_rn(__dmul_rn(P.v[k], P.freq[n]), P.inv_freq);
    const double du = __dmul_rn(__dsub_rn(us, P.uu[m]), P.inv_binsize);
    const double dv = __dmul_rn(__dsub_rn(vs, P.vv[l]), P.inv_binsize);
    return P.conv ? k_exp_sinc(du, dv) : k_ones(du, dv);
}

// ---- contribution keys -------------------------------------------------------------------------
// mode 0: main scatter (footprint nmin/nmax, slots with a zero kernel value are dead)
// mode 1: box sum of weights for uniform/robust re-weighting (footprint npix/npix)
__global__ void __launch_bounds__(256) grid_emit_keys_kernel(GridParams P, int mode, uint32_t lo, uint32_t hi,
                                                             int64_t first, int64_t count, int fp,
                                                             uint32_t *keys, uint32_t *ids)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;       // contribution within the batch
    if (t >= count * fp) return;
    const int64_t idx = first + t / fp;                               // (k, n) flat index
    const int f = (int)(t % fp);
    uint32_t key = KEY_DEAD;
    if (P.good[idx]) {
        uint32_t l, m;
        if (slot_cell(P.gi[idx], P.gj[idx], f, lo, hi, P.G, &l, &m)) {
            const int n = (int)(idx % P.nf);
            bool live = true;
            if (mode == 0) live = conv_value(P, idx / P.nf, n, l, m) != 0.0;
            if (live) key = (l * (uint32_t)P.G + m) * (uint32_t)P.nch + (P.spectral ? (uint32_t)n : 0u);
        }
    }
    keys[t] = key;
    ids[t] = (uint32_t)t;
}

// ---- stable LSD radix sort, 8-bit digits -------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift,
                                                             uint32_t *__restrict__ hist, int nblocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int64_t e = base + r * RS_THREADS + threadIdx.x;
        if (e < n) atomicAdd(&h[(keys[e] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];     // digit-major
}

// exclusive scan of `n` uint32 in place, single block (n = 256 * nblocks)
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *__restrict__ data, int64_t n)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024 * 4) {
        const int64_t e0 = base + (int64_t)threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = (e0 + q < n) ? data[e0 + q] : 0u;
        const uint32_t tsum = v[0] + v[1] + v[2] + v[3];
        uint32_t x = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sums[lane] = s;          // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        uint32_t excl = carry + (wid ? warp_sums[wid - 1] : 0u) + (x - tsum);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (e0 + q < n) data[e0 + q] = excl;
            excl += v[q];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                const uint32_t *__restrict__ vals_in,
                                                                uint32_t *__restrict__ keys_out,
                                                                uint32_t *__restrict__ vals_out, int64_t n, int shift,
                                                                const uint32_t *__restrict__ offsets, int nblocks)
{
    __shared__ uint32_t digit_base[256];
    __shared__ uint32_t warp_cnt[RS_THREADS / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    digit_base[tid] = offsets[(size_t)tid * nblocks + blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ITEMS; r++) {
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; w++) warp_cnt[w][tid] = 0;
        __syncthreads();
        const int64_t e = base + r * RS_THREADS + tid;
        const bool valid = e < n;
        uint32_t key = 0, val = 0;
        if (valid) {
            key = keys_in[e];
            val = vals_in[e];
        }
        const uint32_t d = valid ? ((key >> shift) & 255u) : (256u + lane);   // invalid lanes match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_cnt[wid][d] = __popc(peers);
        __syncthreads();
        // thread `tid` owns digit `tid`: exclusive prefix over warps, then advance the running base
        uint32_t run = digit_base[tid];
        const uint32_t start = run;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; w++) {
            const uint32_t c = warp_cnt[w][tid];
            warp_cnt[w][tid] = run;
            run += c;
        }
        (void)start;
        __syncthreads();
        if (valid) {
            const uint32_t pos = warp_cnt[wid][d] + rank;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
        digit_base[tid] = run;
        // (next iteration's zeroing of warp_cnt is ordered by the barrier above)
    }
}

struct SortBufs {
    uint32_t *k0, *v0, *k1, *v1, *hist;
};

// sorts (k0,v0) by key bits [0, nbits); result pointers returned in *ko, *vo
static int radix_sort(SortBufs b, int64_t n, int nbits, uint32_t **ko, uint32_t **vo)
{
    Context &c = ctx();
    const int nblocks = ceil_div(n, RS_TILE);
    uint32_t *ki = b.k0, *vi = b.v0, *kt = b.k1, *vt = b.v1;
    for (int shift = 0; shift < nbits; shift += 8) {
        {
            LaunchScope ls("grid_sort_hist");
            rs_hist_kernel<<<nblocks, RS_THREADS, 0, c.stream>>>(ki, n, shift, b.hist, nblocks);
            PDSB_CUDA(cudaGetLastError());
        }
        {
            LaunchScope ls("grid_sort_scan");
            rs_scan_kernel<<<1, 1024, 0, c.stream>>>(b.hist, (int64_t)256 * nblocks);
            PDSB_CUDA(cudaGetLastError());
        }
        {
            LaunchScope ls("grid_sort_scatter");
            rs_scatter_kernel<<<nblocks, RS_THREADS, 0, c.stream>>>(ki, vi, kt, vt, n, shift, b.hist, nblocks);
            PDSB_CUDA(cudaGetLastError());
        }
        std::swap(ki, kt);
        std::swap(vi, vt);
    }
    *ko = ki;
    *vo = vi;
    return PDSB_OK;
}

// ---- ordered accumulation: one thread per output cell ----------------------------------------
__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t *a, int64_t n, uint32_t key)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// mode 0: out_re/out_im/out_w += {re,im,1} * w * c ; mode 1: out_w += w (binned weights)
__global__ void __launch_bounds__(128) grid_accumulate_kernel(GridParams P, int mode, uint32_t lo, uint32_t hi,
                                                              int64_t first, int fp,
                                                              const uint32_t *__restrict__ keys,
                                                              const uint32_t *__restrict__ ids, int64_t ncontrib,
                                                              double *out_re, double *out_im, double *out_w)
{
    const int64_t cell = (int64_t)blockIdx.x * 128 + threadIdx.x;      // (l*G+m)*nch + c
    const int64_t ncell = (int64_t)P.G * P.G * P.nch;
    if (cell >= ncell) return;
    int64_t p = lower_bound_u32(keys, ncontrib, (uint32_t)cell);
    if (p >= ncontrib || keys[p] != (uint32_t)cell) return;
    const uint32_t lm = (uint32_t)(cell / P.nch);
    const uint32_t l = lm / (uint32_t)P.G, m = lm % (uint32_t)P.G;
    double sr = 0, si = 0, sw = out_w[cell];
    if (mode == 0) {
        sr = out_re[cell];
        si = out_im[cell];
    }
    for (; p < ncontrib && keys[p] == (uint32_t)cell; p++) {
        const int64_t idx = first + ids[p] / fp;
        const double w = P.w[idx];
        if (mode == 0) {
            const double cv = conv_value(P, idx / P.nf, (int)(idx % P.nf), l, m);
            // new_real[l,m,c] += real[k,n]*weights[k,n]*convolve   (left to right, no FMA)  :514-520
            sr = __dadd_rn(sr, __dmul_rn(__dmul_rn(P.re[idx], w), cv));
            si = __dadd_rn(si, __dmul_rn(__dmul_rn(P.im[idx], w), cv));
            sw = __dadd_rn(sw, __dmul_rn(w, cv));
        } else {
            sw = __dadd_rn(sw, w);
        }
    }
    out_w[cell] = sw;
    if (mode == 0) {
        out_re[cell] = sr;
        out_im[cell] = si;
    }
}

// ---- fast (atomic) scatter -------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_scatter_atomic_kernel(GridParams P, int mode, uint32_t lo, uint32_t hi,
                                                                  int fp, const uint32_t *__restrict__ order,
                                                                  double *out_re, double *out_im, double *out_w)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t total = P.nuv * P.nf;
    if (t >= total * fp) return;
    const int64_t idx = order ? (int64_t)order[t / fp] : t / fp;
    const int f = (int)(t % fp);
    if (!P.good[idx]) return;
    uint32_t l, m;
    if (!slot_cell(P.gi[idx], P.gj[idx], f, lo, hi, P.G, &l, &m)) return;
    const int n = (int)(idx % P.nf);
    const int64_t cell = ((int64_t)l * P.G + m) * P.nch + (P.spectral ? n : 0);
    const double w = P.w[idx];
    if (mode == 0) {
        const double cv = conv_value(P, idx / P.nf, n, l, m);
        if (cv == 0.0) return;
        atomicAdd(out_re + cell, P.re[idx] * w * cv);
        atomicAdd(out_im + cell, P.im[idx] * w * cv);
        atomicAdd(out_w + cell, w * cv);
    } else {
        atomicAdd(out_w + cell, w);
    }
}

// ---- re-weighting, normalisation ---------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_kernel(double *a, int64_t n, double v)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) a[i] = v;
}

// uniform/superuniform: w /= binned[j,i,c] ; robust: w /= (1 + f2[n]*binned[j,i,c])   :463-485
__global__ void __launch_bounds__(256) grid_reweight_kernel(GridParams P, const double *__restrict__ binned,
                                                            const double *__restrict__ f2)
{
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= P.nuv * P.nf) return;
    if (!P.good[idx]) return;
    const int n = (int)(idx % P.nf);
    // The reference addresses binned_weights[l,m,n] with the DATA channel n even in continuum
    // mode, where the last dimension is 1 (:472,:485, bounds checks off): in the contiguous array
    // that is flat element (l*G+m)+n.  Replicated; past the end, channel 0 of the home cell.
    int64_t q = ((int64_t)P.gj[idx] * P.G + P.gi[idx]) * P.nch + n;
    if (!P.spectral && q >= (int64_t)P.G * P.G) q = (int64_t)P.gj[idx] * P.G + P.gi[idx];
    const double b = binned[q];
    if (f2) P.w[idx] = P.w[idx] / __dadd_rn(1.0, __dmul_rn(f2[n], b));
    else P.w[idx] = P.w[idx] / b;
}

// column sums with stride: out[c] = sum_q f(a[q*ncol + c]); square != 0 sums a^2.  One block per column.
__global__ void __launch_bounds__(256) strided_sum_kernel(const double *__restrict__ a, int64_t nrow, int ncol,
                                                          int square, double *__restrict__ out)
{
    __shared__ double sh[8];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int64_t q = threadIdx.x; q < nrow; q += 256) {
        const double x = a[q * ncol + c];
        s += square ? x * x : x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        out[c] = t;
    }
}

// f2[n] = (5*10**(-robust))**2 / (sumb2[c] / sumw[n])     :476-477
__global__ void robust_f2_kernel(const double *sumb2, const double *sumw, int nf, int spectral, double a2,
                                 double *f2)
{
    const int n = threadIdx.x;
    if (n >= nf) return;
    f2[n] = a2 / (sumb2[spectral ? n : 0] / sumw[n]);
}

// imaging: all three maps /= sum(weights) per channel (:525-529); else re,im /= w where w>0 (:531-533)
__global__ void __launch_bounds__(256) grid_normalise_kernel(double *re, double *im, double *w, int64_t ncell, int nch,
                                                             int imaging, const double *__restrict__ wsum)
{
    const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (q >= ncell) return;
    if (imaging) {
        const double s = wsum[q % nch];
        re[q] = re[q] / s;
        im[q] = im[q] / s;
        w[q] = w[q] / s;
    } else if (w[q] > 0) {
        re[q] = re[q] / w[q];
        im[q] = im[q] / w[q];
    }
}

__global__ void __launch_bounds__(256) freqcorrect_kernel(const double *__restrict__ u, const double *__restrict__ v,
                                                          const double *__restrict__ freq, int64_t nuv, int nf,
                                                          double inv_freq, double *__restrict__ ou,
                                                          double *__restrict__ ov)
{
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= nuv * nf) return;
    const double scale = __dmul_rn(freq[idx % nf], inv_freq);          // data.freq * inv_freq  :598
    ou[idx] = __dmul_rn(u[idx / nf], scale);
    ov[idx] = __dmul_rn(v[idx / nf], scale);
}

// home-cell key for the fast path ordering (dead for !good)
__global__ void __launch_bounds__(256) grid_home_keys_kernel(GridParams P, uint32_t *keys, uint32_t *ids)
{
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= P.nuv * P.nf) return;
    // coarse 8x8-cell tiles keep neighbouring visibilities together without a full-resolution sort
    const uint32_t tg = ((uint32_t)P.G + 7u) >> 3;
    keys[idx] = P.good[idx] ? ((P.gj[idx] >> 3) * tg + (P.gi[idx] >> 3)) : KEY_DEAD;
    ids[idx] = (uint32_t)idx;
}

static int bits_for(uint64_t maxval)
{
    int b = 1;
    while (b < 32 && (maxval >> b)) b++;
    return b;
}

}  // namespace pdsb

using namespace pdsb;

extern "C" {

int pdsb_grid(const double *u, const double *v, const double *freq, const double *real, const double *imag,
              const double *weights, int64_t nuv, int nf, int in_kind, int gridsize, double binsize,
              const double *uu, const double *vv, int convolution, int weighting, double robust, int npixels,
              int mode, int imaging, int deterministic, double *out_real, double *out_imag, double *out_weights,
              uint32_t *out_i, uint32_t *out_j, double *out_wmod, int out_kind, int64_t *n_outside)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(nuv >= 0 && nf > 0 && gridsize > 0, "sizes");
    PDSB_REQUIRE(binsize > 0, "binsize");
    PDSB_REQUIRE(uu && vv && freq, "uu/vv/freq");
    PDSB_REQUIRE(convolution == PDSB_CONV_PILLBOX || convolution == PDSB_CONV_EXPSINC, "convolution");
    PDSB_REQUIRE(weighting >= PDSB_WT_NATURAL && weighting <= PDSB_WT_ROBUST, "weighting");
    PDSB_REQUIRE(mode == PDSB_MODE_CONTINUUM || mode == PDSB_MODE_SPECTRALLINE, "mode");
    PDSB_REQUIRE(npixels >= 0, "npixels");
    const int G = gridsize;
    const int nch = mode == PDSB_MODE_SPECTRALLINE ? nf : 1;
    const int64_t ncell = (int64_t)G * G * nch;
    PDSB_REQUIRE(ncell < (int64_t)KEY_DEAD, "gridsize^2 * channels must fit 32 bits");
    const int64_t nvis = nuv * nf;
    PDSB_REQUIRE(nvis < (int64_t)1 << 32, "nuv*nf must fit 32 bits");

    // mean_freq = numpy.mean(freq) (:395): numpy's pairwise sum is sequential below 8 elements and
    // 8-way unrolled above; reproduce both cases on the host (freq is tiny).
    std::vector<double> hfreq(nf);
    if (in_kind == PDSB_DEVICE) {
        PDSB_CUDA(cudaMemcpyAsync(hfreq.data(), freq, nf * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    } else {
        memcpy(hfreq.data(), freq, nf * sizeof(double));
    }
    auto np_pairwise = [&](auto &&self, const double *a, int n) -> double {
        if (n < 8) {
            double r = 0.;
            for (int i = 0; i < n; i++) r += a[i];
            return r;
        } else if (n <= 128) {
            double r[8];
            for (int q = 0; q < 8; q++) r[q] = a[q];
            int i;
            for (i = 8; i < n - (n % 8); i += 8)
                for (int q = 0; q < 8; q++) r[q] += a[i + q];
            double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            for (; i < n; i++) res += a[i];
            return res;
        } else {
            int n2 = n / 2;
            n2 -= n2 % 8;
            return self(self, a, n2) + self(self, a + n2, n - n2);
        }
    };
    const double mean_freq = np_pairwise(np_pairwise, hfreq.data(), nf) / nf;
    const double inv_freq = 1. / mean_freq;

    // ---- device inputs ----
    // scratch map: stage_a = inputs (u,v,freq,re,im,w,uu,vv), stage_b = working arrays,
    // stage_c = output maps, stage_d/e = sort buffers
    const size_t in_bytes = (size_t)(2 * nuv + nf + 3 * nvis + 2 * G) * sizeof(double);
    const double *du = u, *dv = v, *dfreq = freq, *dre = real, *dim = imag, *dw = weights, *duu = uu, *dvv = vv;
    PDSB_CHECK(c.stage_a.ensure(in_bytes + 64));
    {
        double *p = c.stage_a.as<double>();
        auto put = [&](const double *src, size_t n, const double **dst, bool always_host) -> int {
            if (n == 0) {
                *dst = p;
                return PDSB_OK;
            }
            if (in_kind == PDSB_DEVICE && !always_host) {
                *dst = src;
                return PDSB_OK;
            }
            PDSB_CUDA(cudaMemcpyAsync(p, src, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
            *dst = p;
            p += n;
            return PDSB_OK;
        };
        if (nvis > 0) PDSB_REQUIRE(u && v && real && imag && weights, "input arrays");
        PDSB_CHECK(put(u, nuv, &du, false));
        PDSB_CHECK(put(v, nuv, &dv, false));
        PDSB_CHECK(put(freq, nf, &dfreq, false));
        PDSB_CHECK(put(real, nvis, &dre, false));
        PDSB_CHECK(put(imag, nvis, &dim, false));
        PDSB_CHECK(put(weights, nvis, &dw, false));
        // uu/vv: same kind as the other inputs
        PDSB_CHECK(put(uu, G, &duu, false));
        PDSB_CHECK(put(vv, G, &dvv, false));
    }

    // ---- working arrays ----
    const size_t work_bytes = (size_t)nvis * (sizeof(double) + 2 * sizeof(uint32_t) + 1) + 4 * 64 +
                              (size_t)(3 * nf + 4) * sizeof(double) + sizeof(unsigned long long);
    PDSB_CHECK(c.stage_b.ensure(work_bytes + 256));
    char *wp = c.stage_b.as<char>();
    auto carve = [&](size_t bytes) {
        char *r = wp;
        wp += (bytes + 63) / 64 * 64;
        return r;
    };
    double *w_work = (double *)carve((size_t)nvis * sizeof(double));
    uint32_t *gi = (uint32_t *)carve((size_t)nvis * sizeof(uint32_t));
    uint32_t *gj = (uint32_t *)carve((size_t)nvis * sizeof(uint32_t));
    double *small = (double *)carve((size_t)(3 * nf + 4) * sizeof(double));     // sumb2[nf], sumw[nf], f2[nf], wsum
    unsigned long long *d_nout = (unsigned long long *)carve(sizeof(unsigned long long));
    uint8_t *good = (uint8_t *)carve((size_t)nvis);

    double *o_re, *o_im, *o_w;
    PDSB_CHECK(c.stage_c.ensure((size_t)ncell * 4 * sizeof(double) + 256));
    o_re = c.stage_c.as<double>();
    o_im = o_re + ncell;
    o_w = o_im + ncell;
    double *binned = o_w + ncell;

    GridParams P;
    P.u = du; P.v = dv; P.freq = dfreq; P.re = dre; P.im = dim; P.w_in = dw;
    P.w = w_work; P.gi = gi; P.gj = gj; P.good = good;
    P.nuv = nuv; P.nf = nf; P.G = G; P.nch = nch; P.spectral = mode == PDSB_MODE_SPECTRALLINE;
    P.conv = convolution;
    P.binsize = binsize; P.inv_binsize = 1. / binsize; P.inv_freq = inv_freq;
    P.half = (G % 2 == 0) ? G / 2. : (G - 1) / 2.;
    P.uu = duu; P.vv = dvv;
    const int ninclude = convolution == PDSB_CONV_EXPSINC ? 6 : 3;
    if (ninclude % 2 == 0) { P.nmin = (uint32_t)(ninclude * 0.5 - 1); P.nmax = (uint32_t)(ninclude * 0.5); }
    else { P.nmin = P.nmax = (uint32_t)((ninclude - 1) * 0.5); }

    PDSB_CUDA(cudaMemsetAsync(d_nout, 0, sizeof(unsigned long long), c.stream));
    PDSB_CUDA(cudaMemsetAsync(o_re, 0, (size_t)ncell * 3 * sizeof(double), c.stream));
    if (nvis > 0) {
        LaunchScope ls("grid_prep");
        grid_prep_kernel<<<ceil_div(nvis, 256), 256, 0, c.stream>>>(P, d_nout);
        PDSB_CUDA(cudaGetLastError());
    }

    // one engine for both scatters (mode 0 main, mode 1 binned weights)
    auto scatter = [&](int smode, uint32_t lo, uint32_t hi, double *t_re, double *t_im, double *t_w) -> int {
        if (nvis == 0) return PDSB_OK;
        const int fp = (int)((lo + hi + 1) * (lo + hi + 1));
        if (!deterministic) {
            // order visibilities by coarse home tile so concurrent atomics share L2 lines
            const uint32_t tg = ((uint32_t)G + 7u) >> 3;
            PDSB_CHECK(c.stage_d.ensure((size_t)nvis * 4 * sizeof(uint32_t) +
                                        (size_t)256 * ceil_div(nvis, RS_TILE) * sizeof(uint32_t) + 1024));
            uint32_t *k0 = c.stage_d.as<uint32_t>(), *v0 = k0 + nvis, *k1 = v0 + nvis, *v1 = k1 + nvis;
            uint32_t *hist = v1 + nvis;
            {
                LaunchScope ls("grid_home_keys");
                grid_home_keys_kernel<<<ceil_div(nvis, 256), 256, 0, c.stream>>>(P, k0, v0);
                PDSB_CUDA(cudaGetLastError());
            }
            uint32_t *ko, *vo;
            SortBufs sb{k0, v0, k1, v1, hist};
            PDSB_CHECK(radix_sort(sb, nvis, bits_for((uint64_t)tg * tg), &ko, &vo));
            LaunchScope ls("grid_scatter_atomic");
            grid_scatter_atomic_kernel<<<ceil_div(nvis * fp, 256), 256, 0, c.stream>>>(P, smode, lo, hi, fp, vo, t_re,
                                                                                       t_im, t_w);
            PDSB_CUDA(cudaGetLastError());
            return PDSB_OK;
        }
        // deterministic: batches of at most 2^28 contributions, processed in (k,n) order
        const int64_t max_contrib = (int64_t)1 << 28;
        int64_t per_batch = std::max<int64_t>(1, max_contrib / fp);
        const int nbits = bits_for((uint64_t)ncell);          // dead keys (all ones) sort last
        const int sort_bits = ((nbits + 7) / 8) * 8 > 32 ? 32 : ((nbits + 7) / 8) * 8;
        for (int64_t first = 0; first < nvis; first += per_batch) {
            const int64_t count = std::min(per_batch, nvis - first);
            const int64_t ncontrib = count * fp;
            PDSB_CHECK(c.stage_d.ensure((size_t)ncontrib * 4 * sizeof(uint32_t) +
                                        (size_t)256 * ceil_div(ncontrib, RS_TILE) * sizeof(uint32_t) + 1024));
            uint32_t *k0 = c.stage_d.as<uint32_t>(), *v0 = k0 + ncontrib, *k1 = v0 + ncontrib, *v1 = k1 + ncontrib;
            uint32_t *hist = v1 + ncontrib;
            {
                LaunchScope ls("grid_emit_keys");
                grid_emit_keys_kernel<<<ceil_div(ncontrib, 256), 256, 0, c.stream>>>(P, smode, lo, hi, first, count, fp,
                                                                                    k0, v0);
                PDSB_CUDA(cudaGetLastError());
            }
            uint32_t *ko, *vo;
            SortBufs sb{k0, v0, k1, v1, hist};
            // dead keys have every bit set, so sorting on sort_bits bits still puts them last only if
            // sort_bits covers a zero bit of every live key above... live keys < 2^nbits <= 2^sort_bits;
            // dead keys share the low sort_bits with 2^sort_bits-1, which is >= any live key: fine.
            PDSB_CHECK(radix_sort(sb, ncontrib, sort_bits, &ko, &vo));
            LaunchScope ls("grid_accumulate");
            grid_accumulate_kernel<<<ceil_div(ncell, 128), 128, 0, c.stream>>>(P, smode, lo, hi, first, fp, ko, vo,
                                                                              ncontrib, t_re, t_im, t_w);
            PDSB_CUDA(cudaGetLastError());
        }
        return PDSB_OK;
    };

    // ---- optional re-weighting (:429-485) ----
    if (weighting != PDSB_WT_NATURAL) {
        uint32_t npix = weighting == PDSB_WT_SUPERUNIFORM ? 3u : (uint32_t)npixels;
        {
            LaunchScope ls("grid_fill");
            fill_kernel<<<ceil_div(ncell, 256), 256, 0, c.stream>>>(binned, ncell, 1.0);     // numpy.ones :430
            PDSB_CUDA(cudaGetLastError());
        }
        PDSB_CHECK(scatter(1, npix, npix, nullptr, nullptr, binned));
        const double *f2 = nullptr;
        if (weighting == PDSB_WT_ROBUST) {
            double *sumb2 = small, *sumw = small + nf, *f2w = small + 2 * nf;
            {
                LaunchScope ls("grid_sum");
                strided_sum_kernel<<<nch, 256, 0, c.stream>>>(binned, (int64_t)G * G, nch, 1, sumb2);
                PDSB_CUDA(cudaGetLastError());
            }
            {
                LaunchScope ls("grid_sum");
                strided_sum_kernel<<<nf, 256, 0, c.stream>>>(w_work, nuv, nf, 0, sumw);
                PDSB_CUDA(cudaGetLastError());
            }
            PDSB_REQUIRE(nf <= 1024, "robust weighting supports at most 1024 channels");
            const double ra = 5 * pow(10.0, -robust);       // host libm, as Python's 5*10**(-robust)
            robust_f2_kernel<<<1, 1024, 0, c.stream>>>(sumb2, sumw, nf, P.spectral, ra * ra, f2w);
            PDSB_CUDA(cudaGetLastError());
            f2 = f2w;
        }
        if (nvis > 0) {
            LaunchScope ls("grid_reweight");
            grid_reweight_kernel<<<ceil_div(nvis, 256), 256, 0, c.stream>>>(P, binned, f2);
            PDSB_CUDA(cudaGetLastError());
        }
    }

    // ---- main scatter (:489-521) ----
    PDSB_CHECK(scatter(0, P.nmin, P.nmax, o_re, o_im, o_w));

    // ---- normalisation (:525-533) ----
    {
        double *wsum = small + 3 * nf;
        if (imaging) {
            PDSB_REQUIRE(nch <= nf, "channels");
            // per-channel sum of the weight map
            LaunchScope ls("grid_sum");
            strided_sum_kernel<<<nch, 256, 0, c.stream>>>(o_w, (int64_t)G * G, nch, 0,
                                                          nch == 1 ? wsum : binned);   // binned is free now
            PDSB_CUDA(cudaGetLastError());
            if (nch != 1) wsum = binned;
        }
        LaunchScope ls("grid_normalise");
        grid_normalise_kernel<<<ceil_div(ncell, 256), 256, 0, c.stream>>>(o_re, o_im, o_w, ncell, nch, imaging ? 1 : 0,
                                                                         wsum);
        PDSB_CUDA(cudaGetLastError());
    }

    // ---- outputs ----
    cudaMemcpyKind ok = out_kind == PDSB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (out_real) PDSB_CUDA(cudaMemcpyAsync(out_real, o_re, (size_t)ncell * sizeof(double), ok, c.stream));
    if (out_imag) PDSB_CUDA(cudaMemcpyAsync(out_imag, o_im, (size_t)ncell * sizeof(double), ok, c.stream));
    if (out_weights) PDSB_CUDA(cudaMemcpyAsync(out_weights, o_w, (size_t)ncell * sizeof(double), ok, c.stream));
    if (nvis > 0) {
        if (out_i) PDSB_CUDA(cudaMemcpyAsync(out_i, gi, (size_t)nvis * sizeof(uint32_t), ok, c.stream));
        if (out_j) PDSB_CUDA(cudaMemcpyAsync(out_j, gj, (size_t)nvis * sizeof(uint32_t), ok, c.stream));
        if (out_wmod) PDSB_CUDA(cudaMemcpyAsync(out_wmod, w_work, (size_t)nvis * sizeof(double), ok, c.stream));
    }
    unsigned long long hn = 0;
    if (n_outside || out_kind == PDSB_HOST) {
        PDSB_CUDA(cudaMemcpyAsync(&hn, d_nout, sizeof(hn), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
        if (n_outside) *n_outside = (int64_t)hn;
    }
    return PDSB_OK;
}

int pdsb_freqcorrect(const double *u, const double *v, const double *freq, int64_t nuv, int nf, double new_freq,
                     int kind, double *out_u, double *out_v)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(nuv >= 0 && nf > 0 && new_freq > 0, "sizes/new_freq");
    if (nuv == 0) return PDSB_OK;
    PDSB_REQUIRE(u && v && freq && out_u && out_v, "arrays");
    const int64_t n = nuv * nf;
    const double *du, *dv, *df;
    PDSB_CHECK(to_device(u, kind, nuv * sizeof(double), c.stage_a, (const void **)&du));
    PDSB_CHECK(to_device(v, kind, nuv * sizeof(double), c.stage_b, (const void **)&dv));
    PDSB_CHECK(to_device(freq, kind, nf * sizeof(double), c.stage_c, (const void **)&df));
    double *ou = out_u, *ov = out_v;
    if (kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_d.ensure((size_t)n * sizeof(double)));
        PDSB_CHECK(c.stage_e.ensure((size_t)n * sizeof(double)));
        ou = c.stage_d.as<double>();
        ov = c.stage_e.as<double>();
    }
    {
        LaunchScope ls("freqcorrect");
        freqcorrect_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(du, dv, df, nuv, nf, 1. / new_freq, ou, ov);
        PDSB_CUDA(cudaGetLastError());
    }
    if (kind == PDSB_HOST) {
        PDSB_CUDA(cudaMemcpyAsync(out_u, ou, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(out_v, ov, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}

}  // extern "C"
