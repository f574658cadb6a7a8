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
//  fast: visibilities sorted by 8x8-cell uv tile (one or two radix passes); one CTA per tile stages
//      them through shared memory and accumulates the tile's footprint region in registers (one
//      cell per thread), then adds the region to the map with one fp64 atomic per cell.
//
// Index maps reproduce numpy's left-to-right fp64 arithmetic and its float64->uint32 cast
// (truncation through int64, wrap mod 2^32, NaN -> 0) exactly: __dmul_rn/__ddiv_rn/__dadd_rn keep
// the compiler from contracting the expression into FMAs.
#include "grid.cuh"
#include <algorithm>
#include <cmath>

namespace pdsb {

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

// Whether slot (l, m) can receive a non-zero kernel value: the support test of the two kernels (:559-585)
// without the polynomial.  (A value that is exactly 0 inside the support would only add 0.)
__device__ __forceinline__ bool conv_live(const GridParams &P, int64_t k, int n, uint32_t l, uint32_t m)
{
    const double us = __dmul_rn(__dmul_rn(P.u[k], P.freq[n]), P.inv_freq);
    const double vs = __dmul_rn(__dmul_rn(P.v[k], P.freq[n]), P.inv_freq);
    const double du = __dmul_rn(__dsub_rn(us, P.uu[m]), P.inv_binsize);
    const double dv = __dmul_rn(__dsub_rn(vs, P.vv[l]), P.inv_binsize);
    const double lim = P.conv ? 3.0 : 0.5;
    return fabs(du) < lim && fabs(dv) < lim;
}

// ---- contribution keys -------------------------------------------------------------------------
// mode 0: main scatter (footprint nmin/nmax, slots with a zero kernel value are dead)
// mode 1: box sum of weights for uniform/robust re-weighting (footprint npix/npix)
// One thread per emitted contribution.  `compact` (pillbox main scatter): the kernel value is 1 for
// at most one of the 9 footprint cells, so one thread per visibility scans its slots and emits only
// that one (fp_emit = 1) - the sort then handles nvis keys instead of 9*nvis.
__global__ void __launch_bounds__(256) grid_emit_keys_kernel(GridParams P, int mode, uint32_t lo, uint32_t hi,
                                                             int64_t first, int64_t count, int fp, int compact,
                                                             uint32_t *keys, uint32_t *ids)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;       // contribution within the batch
    const int fp_emit = compact ? 1 : fp;
    if (t >= count * fp_emit) return;
    const int64_t idx = first + t / fp_emit;                          // (k, n) flat index
    uint32_t key = KEY_DEAD;
    if (P.good[idx]) {
        const int n = (int)(idx % P.nf);
        const int f0 = compact ? 0 : (int)(t % fp), f1 = compact ? fp : f0 + 1;
        for (int f = f0; f < f1; f++) {
            uint32_t l, m;
            if (!slot_cell(P.gi[idx], P.gj[idx], f, lo, hi, P.G, &l, &m)) continue;
            bool live = l >= P.row_lo && l < P.row_hi;                   // (the whole grid unless pdsb_set_grid_band)
            if (live && mode == 0) live = conv_live(P, idx / P.nf, n, l, m);
            if (live) {
                key = (l * (uint32_t)P.G + m) * (uint32_t)P.nch + (P.spectral ? (uint32_t)n : 0u);
                break;
            }
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

// Exclusive scan of `n` uint32 in place, three small kernels: per-segment scan + segment totals,
// single-block scan of the totals, add-back.  SCAN_SEG elements per block.
constexpr int SCAN_SEG = 1024 * 8;

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t x, uint32_t *warp_sums, uint32_t *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_sums[wid] = inc;
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
    const uint32_t base = wid ? warp_sums[wid - 1] : 0u;
    *total = warp_sums[31];
    __syncthreads();
    return base + inc - x;
}

__global__ void __launch_bounds__(1024) rs_scan_seg_kernel(uint32_t *__restrict__ data, int64_t n,
                                                           uint32_t *__restrict__ seg_totals)
{
    __shared__ uint32_t warp_sums[32];
    const int64_t e0 = (int64_t)blockIdx.x * SCAN_SEG + (int64_t)threadIdx.x * 8;
    uint32_t v[8], tsum = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        v[q] = (e0 + q < n) ? data[e0 + q] : 0u;
        tsum += v[q];
    }
    uint32_t total;
    uint32_t excl = block_excl_scan_1024(tsum, warp_sums, &total);
#pragma unroll
    for (int q = 0; q < 8; q++) {
        if (e0 + q < n) data[e0 + q] = excl;
        excl += v[q];
    }
    if (threadIdx.x == 0) seg_totals[blockIdx.x] = total;
}

// single block: exclusive scan of up to any number of segment totals (loops in chunks of 1024)
__global__ void __launch_bounds__(1024) rs_scan_totals_kernel(uint32_t *__restrict__ t, int nseg)
{
    __shared__ uint32_t warp_sums[32];
    uint32_t carry = 0;
    for (int base = 0; base < nseg; base += 1024) {
        const int e = base + threadIdx.x;
        const uint32_t x = e < nseg ? t[e] : 0u;
        uint32_t total;
        const uint32_t excl = block_excl_scan_1024(x, warp_sums, &total);
        if (e < nseg) t[e] = carry + excl;
        carry += total;
    }
}

__global__ void __launch_bounds__(1024) rs_scan_add_kernel(uint32_t *__restrict__ data, int64_t n,
                                                           const uint32_t *__restrict__ seg_offsets)
{
    const uint32_t off = seg_offsets[blockIdx.x];
    const int64_t e0 = (int64_t)blockIdx.x * SCAN_SEG + (int64_t)threadIdx.x * 8;
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (e0 + q < n) data[e0 + q] += off;
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
            const int64_t nh = (int64_t)256 * nblocks;
            const int nseg = ceil_div(nh, SCAN_SEG);
            uint32_t *seg = b.hist + nh;                  // scratch behind the histogram
            rs_scan_seg_kernel<<<nseg, 1024, 0, c.stream>>>(b.hist, nh, seg);
            rs_scan_totals_kernel<<<1, 1024, 0, c.stream>>>(seg, nseg);
            rs_scan_add_kernel<<<nseg, 1024, 0, c.stream>>>(b.hist, nh, seg);
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

// Step 1 (parallel over sorted contributions): the addends, in sorted order.
//   mode 0: val[0..2][p] = real*w*c, imag*w*c, w*c  (left to right, non-contracted  :514-520)
//   mode 1: val[2][p] = w                           (binned weights)
__global__ void __launch_bounds__(256) grid_values_kernel(GridParams P, int mode, int64_t first, int fp_emit,
                                                          const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ ids, int64_t ncontrib,
                                                          double *__restrict__ v_re, double *__restrict__ v_im,
                                                          double *__restrict__ v_w)
{
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= ncontrib) return;
    const uint32_t key = keys[p];
    if (key == KEY_DEAD) return;
    const int64_t idx = first + ids[p] / fp_emit;
    const double w = P.w[idx];
    if (mode == 0) {
        const uint32_t lm = key / (uint32_t)P.nch;
        const double cv = conv_value(P, idx / P.nf, (int)(idx % P.nf), lm / (uint32_t)P.G, lm % (uint32_t)P.G);
        v_re[p] = __dmul_rn(__dmul_rn(P.re[idx], w), cv);
        v_im[p] = __dmul_rn(__dmul_rn(P.im[idx], w), cv);
        v_w[p] = __dmul_rn(w, cv);
    } else {
        v_w[p] = w;
    }
}

// Step 2: ordered sums.  A warp takes 32 consecutive cells; every lane locates its own run.
// Short runs are added by the owning lane.  Long runs (the dense centre of the uv plane) are
// streamed by the whole warp: 32 lanes fetch 32 consecutive addends of each map into shared
// memory (coalesced, independent of the running sum), then the owning lane adds them in order -
// the serial chain per element is one dependent DADD per map and nothing else.
constexpr int OS_LONG = 96;
constexpr int OS_PER = 4;               // addends per lane per iteration
constexpr int OS_CH = 32 * OS_PER;
// Runs longer than OS_XLONG are not summed here: a warp would work through the very long runs of its 32
// cells one after the other (the dense centre of the uv plane puts several multi-million-addend cells
// next to each other).  They go on a list and grid_ordered_long_kernel gives each its own warp.
constexpr int OS_XLONG = 8192;
struct LongRun {
    long long cell, p, e;
};

// the in-order sum of one run, streamed by a whole warp (see above); `owner` adds
__device__ __forceinline__ void ordered_stream(int mode, int lane, int owner, int64_t rp, int64_t re_,
                                               const double *__restrict__ v_re, const double *__restrict__ v_im,
                                               const double *__restrict__ v_w, double (*sb)[OS_CH], double &sr,
                                               double &si, double &sw)
{
    // OS_CH addends per iteration; the next iteration's loads are in flight (registers) while
    // the owner adds the current ones out of shared memory
    double nr[OS_PER], ni[OS_PER], nw[OS_PER];
#pragma unroll
    for (int t = 0; t < OS_PER; t++) {
        const int64_t q = rp + t * 32 + lane;
        nr[t] = (mode == 0 && q < re_) ? v_re[q] : 0.0;
        ni[t] = (mode == 0 && q < re_) ? v_im[q] : 0.0;
        nw[t] = q < re_ ? v_w[q] : 0.0;
    }
    for (int64_t base = rp; base < re_; base += OS_CH) {
#pragma unroll
        for (int t = 0; t < OS_PER; t++) {
            sb[0][t * 32 + lane] = nr[t];
            sb[1][t * 32 + lane] = ni[t];
            sb[2][t * 32 + lane] = nw[t];
        }
#pragma unroll
        for (int t = 0; t < OS_PER; t++) {
            const int64_t q = base + OS_CH + t * 32 + lane;
            nr[t] = (mode == 0 && q < re_) ? v_re[q] : 0.0;
            ni[t] = (mode == 0 && q < re_) ? v_im[q] : 0.0;
            nw[t] = q < re_ ? v_w[q] : 0.0;
        }
        __syncwarp();
        if (lane == owner) {
            const int n = (int)(re_ - base < OS_CH ? re_ - base : OS_CH);
            if (mode == 0) {
#pragma unroll 16
                for (int t = 0; t < n; t++) {
                    sr = __dadd_rn(sr, sb[0][t]);
                    si = __dadd_rn(si, sb[1][t]);
                    sw = __dadd_rn(sw, sb[2][t]);
                }
            } else {
#pragma unroll 16
                for (int t = 0; t < n; t++) sw = __dadd_rn(sw, sb[2][t]);
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32) grid_ordered_long_kernel(int mode, const LongRun *__restrict__ runs,
                                                               const unsigned int *__restrict__ nruns,
                                                               const double *__restrict__ v_re,
                                                               const double *__restrict__ v_im,
                                                               const double *__restrict__ v_w, double *out_re,
                                                               double *out_im, double *out_w)
{
    __shared__ double sbuf[3][OS_CH];
    const int lane = threadIdx.x;
    for (unsigned int r = blockIdx.x; r < *nruns; r += gridDim.x) {
        const LongRun run = runs[r];
        double sr = 0, si = 0, sw = 0;
        if (lane == 0) {
            sw = out_w[run.cell];
            if (mode == 0) {
                sr = out_re[run.cell];
                si = out_im[run.cell];
            }
        }
        ordered_stream(mode, lane, 0, run.p, run.e, v_re, v_im, v_w, sbuf, sr, si, sw);
        if (lane == 0) {
            out_w[run.cell] = sw;
            if (mode == 0) {
                out_re[run.cell] = sr;
                out_im[run.cell] = si;
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) grid_ordered_sum_kernel(int mode, int64_t ncell,
                                                               const uint32_t *__restrict__ keys, int64_t ncontrib,
                                                               const double *__restrict__ v_re,
                                                               const double *__restrict__ v_im,
                                                               const double *__restrict__ v_w, double *out_re,
                                                               double *out_im, double *out_w, LongRun *runs,
                                                               unsigned int *nruns)
{
    __shared__ double sbuf[4][3][OS_CH];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t cell = (int64_t)blockIdx.x * 128 + threadIdx.x;      // (l*G+m)*nch + c
    int64_t p = 0, e = 0;
    if (cell < ncell) {
        p = lower_bound_u32(keys, ncontrib, (uint32_t)cell);
        if (p < ncontrib && keys[p] == (uint32_t)cell) {
            e = p + 1;
            int64_t step = 1;                      // end of the run: gallop then bisect
            while (e < ncontrib && keys[e] == (uint32_t)cell) {
                e += step;
                step <<= 1;
            }
            if (e > ncontrib) e = ncontrib;
            int64_t lo = p, hi = e;                // keys[lo]==cell ; keys[hi]!=cell or hi==ncontrib
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (keys[mid] == (uint32_t)cell) lo = mid;
                else hi = mid;
            }
            e = lo + 1;
        } else {
            p = e = 0;
        }
    }
    bool has = e > p;
    if (has && runs && e - p > OS_XLONG) {            // deferred to grid_ordered_long_kernel
        runs[atomicAdd(nruns, 1u)] = LongRun{(long long)cell, (long long)p, (long long)e};
        has = false;
    }
    double sr = 0, si = 0, sw = 0;
    if (has) {
        sw = out_w[cell];
        if (mode == 0) {
            sr = out_re[cell];
            si = out_im[cell];
        }
    }
    const bool is_long = has && (e - p) > OS_LONG;
    if (has && !is_long) {
        for (; p < e; p++) {
            if (mode == 0) {
                sr = __dadd_rn(sr, v_re[p]);
                si = __dadd_rn(si, v_im[p]);
            }
            sw = __dadd_rn(sw, v_w[p]);
        }
    }
    unsigned long_mask = __ballot_sync(0xffffffffu, is_long);
    while (long_mask) {
        const int owner = __ffs(long_mask) - 1;
        long_mask &= long_mask - 1;
        const int64_t rp = __shfl_sync(0xffffffffu, p, owner), re_ = __shfl_sync(0xffffffffu, e, owner);
        ordered_stream(mode, lane, owner, rp, re_, v_re, v_im, v_w, sbuf[wid], sr, si, sw);
    }
    if (has) {
        out_w[cell] = sw;
        if (mode == 0) {
            out_re[cell] = sr;
            out_im[cell] = si;
        }
    }
}

// ---- fast (atomic) scatter -------------------------------------------------------------------
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
    double w;
    if (f2) w = P.w[idx] / __dadd_rn(1.0, __dmul_rn(f2[n], b));
    else w = P.w[idx] / b;
    P.w[idx] = w;
}

// Column sums with stride, stage 1: part[b][c] = sum over this block's rows of f(a[q*ncol + c]);
// square != 0 sums a^2.  grid = (row blocks, columns).  Stage 2 adds the row blocks in order.
constexpr int SUM_BLOCKS = 256;
__global__ void __launch_bounds__(256) strided_sum_kernel(const double *__restrict__ a, int64_t nrow, int ncol,
                                                          int square, double *__restrict__ part)
{
    __shared__ double sh[8];
    const int c = blockIdx.y;
    double s = 0.0;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < nrow; q += (int64_t)gridDim.x * 256) {
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
        part[(size_t)blockIdx.x * ncol + c] = t;
    }
}
__global__ void __launch_bounds__(256) strided_sum_final_kernel(const double *__restrict__ part, int nb, int ncol,
                                                                double *__restrict__ out)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= ncol) return;
    double t = 0.0;
    for (int b = 0; b < nb; b++) t += part[(size_t)b * ncol + c];
    out[c] = t;
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

// ---- average(): nearest-cell binning with weighted mean positions (libinterferometry.pyx:151-311) ----
// Everything per element runs here: the weight clamp (:179-181), the uvdist != 0 and on-grid filters
// (:183-189, :250-260), numpy.round bin indices (:221-248) with numpy's operation order and float64 -> uint32
// cast, the ordered (k, n) accumulation (:262-277), the normalisation (:279-284), the channel-weighted mean
// positions (:286-294) and the compaction of the non-empty cells (:296-311).
struct AvgParams {
    const double *u, *v, *uvdist, *log_uvdist, *freq, *re, *im, *w;
    int64_t nuv;          // input rows
    int nf;               // input channels
    int mfs;              // rows are (k, n) pairs scaled to the mean frequency (freqcorrect, :161-170), one channel each
    double mfs_inv_freq;
    int64_t nrow;         // rows after mfs: nuv * nf or nuv
    int nfe;              // channels per row after mfs: 1 or nf
    int G, Gj, nch, spectral;
    int radial;           // 0: (u, v) grid, 1: linear radial bins, 2: log radial bins
    double binsize, half, log_min, dtemp;
    uint32_t *row_cell;   // [nrow] j * G + i, or KEY_DEAD
    double *row_u, *row_v;   // [nrow] (mfs only; otherwise u, v are read directly)
};

__global__ void __launch_bounds__(256) avg_prep_kernel(AvgParams P, unsigned long long *n_dropped)
{
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= P.nrow) return;
    double ur, vr, dist;
    if (P.mfs) {
        const double scale = __dmul_rn(P.freq[r % P.nf], P.mfs_inv_freq);       // data.freq * inv_freq  :598
        ur = __dmul_rn(P.u[r / P.nf], scale);
        vr = __dmul_rn(P.v[r / P.nf], scale);
        dist = __dsqrt_rn(__dadd_rn(__dmul_rn(ur, ur), __dmul_rn(vr, vr)));     // Visibilities(): sqrt(u**2 + v**2)  :27
        P.row_u[r] = ur;
        P.row_v[r] = vr;
    } else {
        ur = P.u[r];
        vr = P.v[r];
        dist = P.uvdist[r];
    }
    uint32_t cell = KEY_DEAD;
    if (dist != 0.0) {                                                          // good_data = uvdist != 0.0  :183
        uint32_t i, j = 0u;
        if (P.radial == 0) {
            i = np_f64_to_u32(rint(__dadd_rn(__ddiv_rn(ur, P.binsize), P.half)));       // :243-248
            j = np_f64_to_u32(rint(__dadd_rn(__ddiv_rn(vr, P.binsize), P.half)));
        } else if (P.radial == 1) {
            i = np_f64_to_u32(rint(__ddiv_rn(dist, P.binsize)));                // :225
        } else {
            // (log10(uvdist) - log10(logmin)) / dtemp - 0.5, log10(uvdist) from the host's libm  :221
            i = np_f64_to_u32(rint(__dsub_rn(__ddiv_rn(__dsub_rn(P.log_uvdist[r], P.log_min), P.dtemp), 0.5)));
        }
        if (i < (uint32_t)P.G && j < (uint32_t)P.Gj) cell = j * (uint32_t)P.G + i;
        else atomicAdd(n_dropped, 1ull);
    }
    P.row_cell[r] = cell;
}

__global__ void __launch_bounds__(256) avg_emit_keys_kernel(AvgParams P, int64_t first, int64_t count, uint32_t *keys,
                                                            uint32_t *ids)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= count) return;
    const int64_t idx = first + t;
    const uint32_t cell = P.row_cell[idx / P.nfe];
    keys[t] = cell == KEY_DEAD ? KEY_DEAD : cell * (uint32_t)P.nch + (P.spectral ? (uint32_t)(idx % P.nfe) : 0u);
    ids[t] = (uint32_t)t;
}

// addends in sorted order: which = 0: (real*w, imag*w, w) ; which = 1: (u*w, v*w, w), w clamped as :179-181
__global__ void __launch_bounds__(256) avg_values_kernel(AvgParams P, int which, int64_t first,
                                                         const uint32_t *__restrict__ ids, int64_t ncontrib,
                                                         double *__restrict__ va, double *__restrict__ vb,
                                                         double *__restrict__ vw)
{
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= ncontrib) return;
    const int64_t idx = first + ids[p];               // element of the [nrow, nfe] view == element of [nuv, nf]
    const double re = P.re[idx], im = P.im[idx];
    double ww = P.w[idx];
    ww = ww < 0 ? 0.0 : ww;
    if (re == 0 && im == 0) ww = 0.0;
    if (which == 0) {
        va[p] = __dmul_rn(re, ww);
        vb[p] = __dmul_rn(im, ww);
    } else {
        const int64_t r = idx / P.nfe;
        va[p] = __dmul_rn(P.mfs ? P.row_u[r] : P.u[r], ww);
        vb[p] = __dmul_rn(P.mfs ? P.row_v[r] : P.v[r], ww);
    }
    vw[p] = ww;
}

// numpy's pairwise sum of a contiguous run (what ndarray.sum(axis=-1) does per row): < 8 sequential, <= 128 in
// eight interleaved partial sums, above that split in halves (multiples of 8)
__device__ double np_pairwise_sum_dev(const double *a, int n)
{
    if (n < 8) {
        double r = 0.;
        for (int i = 0; i < n; i++) r = __dadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int q = 0; q < 8; q++) r[q] = a[q];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int q = 0; q < 8; q++) r[q] = __dadd_rn(r[q], a[i + q]);
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __dadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise_sum_dev(a, n2), np_pairwise_sum_dev(a + n2, n - n2));
}

// :279-294 per position (j, i): normalise the channels that received weight, flag the position when any did,
// and (2-D grid) form the channel-weighted mean position (new_u * new_weights).sum(axis=2) / new_weights.sum(axis=2).
// m_u / m_v are overwritten with the products new_u * new_weights (scratch for the pairwise sums).
__global__ void __launch_bounds__(128) avg_finish_kernel(int64_t npos, int nch, int radial, double *m_re, double *m_im,
                                                         double *m_w, double *m_u, double *m_v, double *pos_u,
                                                         double *pos_v, uint32_t *flag)
{
    const int64_t q = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (q >= npos) return;
    bool any = false;
    for (int c = 0; c < nch; c++) {
        const int64_t e = q * nch + c;
        const double w = m_w[e];
        if (w != 0.0) {
            any = true;
            m_re[e] = __ddiv_rn(m_re[e], w);
            m_im[e] = __ddiv_rn(m_im[e], w);
            if (!radial) {
                m_u[e] = __dmul_rn(__ddiv_rn(m_u[e], w), w);
                m_v[e] = __dmul_rn(__ddiv_rn(m_v[e], w), w);
            }
        } else if (!radial) {
            m_u[e] = __dmul_rn(m_u[e], w);
            m_v[e] = __dmul_rn(m_v[e], w);
        }
    }
    flag[q] = any ? 1u : 0u;
    if (any && !radial) {
        const double ws = np_pairwise_sum_dev(m_w + q * nch, nch);
        pos_u[q] = __ddiv_rn(np_pairwise_sum_dev(m_u + q * nch, nch), ws);
        pos_v[q] = __ddiv_rn(np_pairwise_sum_dev(m_v + q * nch, nch), ws);
    }
}

// :296-311: the flagged positions, in row-major (j, i) order; `rank` is the exclusive scan of the flags
__global__ void __launch_bounds__(256) avg_compact_kernel(int64_t npos, int nch, int radial, int G,
                                                          const uint32_t *__restrict__ flag_rank,
                                                          const uint32_t *__restrict__ flag,
                                                          const double *__restrict__ m_re, const double *__restrict__ m_im,
                                                          const double *__restrict__ m_w, const double *__restrict__ pos_u,
                                                          const double *__restrict__ pos_v,
                                                          const double *__restrict__ centres, double *o_u, double *o_v,
                                                          double *o_re, double *o_im, double *o_w)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= npos * nch) return;
    const int64_t q = t / nch;
    const int c = (int)(t % nch);
    if (!flag[q]) return;
    const int64_t p = flag_rank[q];
    o_re[p * nch + c] = m_re[t];
    o_im[p * nch + c] = m_im[t];
    o_w[p * nch + c] = m_w[t];
    if (c == 0) {
        o_u[p] = radial ? centres[q % G] : pos_u[q];
        o_v[p] = radial ? 0.0 : pos_v[q];
    }
}

// ---- center(): data * conj(point model)  (center.py:5-25, model.py:102-104) -------------------------
__global__ void __launch_bounds__(256) center_kernel(const double *__restrict__ u, const double *__restrict__ v,
                                                     const double *__restrict__ freq, const double *__restrict__ re,
                                                     const double *__restrict__ im, int64_t nuv, int nf,
                                                     double mean_freq, double x0, double y0, double *__restrict__ ore,
                                                     double *__restrict__ oim)
{
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= nuv * nf) return;
    const int64_t k = idx / nf;
    const double f = freq[idx % nf];
    // model(data.u*data.freq[i]/data.freq.mean(), ...) -> point_model: flux*exp(-2*3.14159*(0+1j*(u*x0+v*y0)))
    const double us = __ddiv_rn(__dmul_rn(u[k], f), mean_freq), vs = __ddiv_rn(__dmul_rn(v[k], f), mean_freq);
    const double a = __dadd_rn(__dmul_rn(us, x0), __dmul_rn(vs, y0));
    const double b = __dmul_rn(-2 * 3.14159, a);
    double s, c;
    sincos(b, &s, &c);
    // centered = (re + i im) * conj(c + i s) = (re + i im) * (c - i s)
    const double dr = re[idx], di = im[idx], ms = -s;
    ore[idx] = __dsub_rn(__dmul_rn(dr, c), __dmul_rn(di, ms));
    oim[idx] = __dadd_rn(__dmul_rn(dr, ms), __dmul_rn(di, c));
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

// weights_phase != 0: only the first half of the re-weighting (:429-461) - the box sums of the clamped weights
// (WITHOUT the initial ones) into binned_out [G*G*nch] (device) and the per-channel weight sums into
// sumw_out [nf] (host) - for the multi-GPU paths, which reduce both over the ranks before the second half.
static int grid_impl(const double *u, const double *v, const double *freq, const double *real, const double *imag,
                     const double *weights, int64_t nuv, int nf, int in_kind, int gridsize, double binsize,
                     const double *uu, const double *vv, int convolution, int weighting, double robust, int npixels,
                     int mode, int imaging, int deterministic, double *out_real, double *out_imag, double *out_weights,
                     uint32_t *out_i, uint32_t *out_j, double *out_wmod, int out_kind, int64_t *n_outside,
                     int weights_phase, double *binned_out, double *sumw_out)
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
            PDSB_CHECK(copy_h2d(p, src, n * sizeof(double)));
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
    const size_t work_bytes = (size_t)nvis * (sizeof(double) + 2 * sizeof(uint32_t) + 1) + 6 * 64 +
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
    // device outputs: accumulate straight into the caller's maps (no copy of 3 x G^2 x nch doubles at the end)
    const bool direct_out = out_kind == PDSB_DEVICE && out_real && out_imag && out_weights;
    if (direct_out) {
        o_re = out_real;
        o_im = out_imag;
        o_w = out_weights;
    }

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

    P.row_lo = 0;
    P.row_hi = (uint32_t)G;
    const bool ext_weights = !weights_phase && c.grid_ext_binned != nullptr;
    if (c.grid_row_hi > 0) {           // pdsb_set_grid_band: this process owns a band of output rows
        PDSB_REQUIRE(deterministic && (weights_phase || imaging == 2),
                     "a row band needs the ordered mode and raw sums (imaging = 2)");
        PDSB_REQUIRE(weights_phase || weighting == PDSB_WT_NATURAL || ext_weights,
                     "re-weighting inside a row band needs the reduced weight map (pdsb_set_grid_reweight)");
        PDSB_REQUIRE(c.grid_row_lo >= 0 && c.grid_row_lo < c.grid_row_hi && c.grid_row_hi <= G, "row band");
        P.row_lo = (uint32_t)c.grid_row_lo;
        P.row_hi = (uint32_t)c.grid_row_hi;
    }

    PDSB_CUDA(cudaMemsetAsync(d_nout, 0, sizeof(unsigned long long), c.stream));
    if (direct_out) {
        PDSB_CUDA(cudaMemsetAsync(o_re, 0, (size_t)ncell * sizeof(double), c.stream));
        PDSB_CUDA(cudaMemsetAsync(o_im, 0, (size_t)ncell * sizeof(double), c.stream));
        PDSB_CUDA(cudaMemsetAsync(o_w, 0, (size_t)ncell * sizeof(double), c.stream));
    } else {
        PDSB_CUDA(cudaMemsetAsync(o_re, 0, (size_t)ncell * 3 * sizeof(double), c.stream));
    }
    // The prep kernel (working weights, index maps, good mask) feeds the ordered mode, the re-weighting and the map
    // outputs; the fast mode with natural weights forms everything it needs on the fly from the inputs.
    const bool prepared = deterministic || weights_phase || weighting != PDSB_WT_NATURAL || out_i || out_j || out_wmod;
    int counted_outside = 0;
    if (nvis > 0 && prepared) {
        LaunchScope ls("grid_prep");
        grid_prep_kernel<<<ceil_div(nvis, 256), 256, 0, c.stream>>>(P, d_nout);
        PDSB_CUDA(cudaGetLastError());
    }

    auto hist_bytes = [&](int64_t n) -> size_t {
        const size_t nh = (size_t)256 * ceil_div(n, RS_TILE);
        return (nh + nh / SCAN_SEG + 16) * sizeof(uint32_t) + 1024;
    };
    // out[c] = sum_q f(a[q*ncol + c]) in a fixed order (two stages)
    auto sum_columns = [&](const double *a, int64_t nrow, int ncol, int square, double *out) -> int {
        const int nb = (int)std::min<int64_t>(SUM_BLOCKS, std::max<int64_t>(1, (nrow + 255) / 256));
        PDSB_CHECK(c.red.ensure((size_t)nb * ncol * sizeof(double)));
        LaunchScope ls("grid_sum");
        strided_sum_kernel<<<dim3(nb, ncol), 256, 0, c.stream>>>(a, nrow, ncol, square, c.red.as<double>());
        strided_sum_final_kernel<<<ceil_div(ncol, 256), 256, 0, c.stream>>>(c.red.as<double>(), nb, ncol, out);
        PDSB_CUDA(cudaGetLastError());
        return PDSB_OK;
    };

    // one engine for both scatters (mode 0 main, mode 1 binned weights)
    auto scatter = [&](int smode, uint32_t lo, uint32_t hi, double *t_re, double *t_im, double *t_w) -> int {
        if (nvis == 0) return PDSB_OK;
        const int fp = (int)((lo + hi + 1) * (lo + hi + 1));
        if (!deterministic)
            return grid_fast_scatter(P, smode, lo, hi, prepared ? w_work : nullptr, t_re, t_im, t_w,
                                     (!prepared && !counted_outside++) ? d_nout : nullptr);
        // deterministic: batches of at most 2^27 contributions, processed in (k,n) order
        const int compact = (smode == 0 && convolution == PDSB_CONV_PILLBOX) ? 1 : 0;
        const int fp_emit = compact ? 1 : fp;
        const int64_t max_contrib = (int64_t)1 << 27;
        int64_t per_batch = std::max<int64_t>(1, max_contrib / fp_emit);
        const int nbits = bits_for((uint64_t)ncell);          // dead keys (all ones) sort last
        const int sort_bits = ((nbits + 7) / 8) * 8 > 32 ? 32 : ((nbits + 7) / 8) * 8;
        for (int64_t first = 0; first < nvis; first += per_batch) {
            const int64_t count = std::min(per_batch, nvis - first);
            const int64_t ncontrib = count * fp_emit;
            PDSB_CHECK(c.stage_d.ensure((size_t)ncontrib * 4 * sizeof(uint32_t) + hist_bytes(ncontrib)));
            PDSB_CHECK(c.stage_e.ensure((size_t)ncontrib * 3 * sizeof(double)));
            uint32_t *k0 = c.stage_d.as<uint32_t>(), *v0 = k0 + ncontrib, *k1 = v0 + ncontrib, *v1 = k1 + ncontrib;
            uint32_t *hist = v1 + ncontrib;
            double *val_re = c.stage_e.as<double>(), *val_im = val_re + ncontrib, *val_w = val_im + ncontrib;
            {
                LaunchScope ls("grid_emit_keys");
                grid_emit_keys_kernel<<<ceil_div(ncontrib, 256), 256, 0, c.stream>>>(P, smode, lo, hi, first, count, fp,
                                                                                    compact, k0, v0);
                PDSB_CUDA(cudaGetLastError());
            }
            uint32_t *ko, *vo;
            SortBufs sb{k0, v0, k1, v1, hist};
            // live keys are < 2^nbits <= 2^sort_bits; dead keys (all ones) compare >= any live key on
            // the low sort_bits, so they end up behind every live run.
            PDSB_CHECK(radix_sort(sb, ncontrib, sort_bits, &ko, &vo));
            {
                LaunchScope ls("grid_values");
                grid_values_kernel<<<ceil_div(ncontrib, 256), 256, 0, c.stream>>>(P, smode, first, fp_emit, ko, vo,
                                                                                 ncontrib, val_re, val_im, val_w);
                PDSB_CUDA(cudaGetLastError());
            }
            // very long runs (> OS_XLONG addends) are listed, then summed one warp per run
            const size_t max_runs = (size_t)(ncontrib / OS_XLONG) + 2;
            PDSB_CHECK(c.folded.ensure(max_runs * sizeof(LongRun) + 64));      // (the DFT's scratch is free here)
            unsigned int *nruns = c.folded.as<unsigned int>();
            LongRun *runs = reinterpret_cast<LongRun *>(c.folded.as<unsigned char>() + 64);
            PDSB_CUDA(cudaMemsetAsync(nruns, 0, sizeof(unsigned int), c.stream));
            LaunchScope ls("grid_ordered_sum");
            grid_ordered_sum_kernel<<<ceil_div(ncell, 128), 128, 0, c.stream>>>(smode, ncell, ko, ncontrib, val_re,
                                                                               val_im, val_w, t_re, t_im, t_w, runs, nruns);
            grid_ordered_long_kernel<<<(unsigned)std::min<size_t>(max_runs, 16384), 32, 0, c.stream>>>(
                smode, runs, nruns, val_re, val_im, val_w, t_re, t_im, t_w);
            PDSB_CUDA(cudaGetLastError());
        }
        return PDSB_OK;
    };

    if (weights_phase) {
        PDSB_REQUIRE(weighting != PDSB_WT_NATURAL && binned_out && sumw_out, "weights phase arguments");
        const uint32_t npix = weighting == PDSB_WT_SUPERUNIFORM ? 3u : (uint32_t)npixels;
        PDSB_CUDA(cudaMemsetAsync(binned, 0, (size_t)ncell * sizeof(double), c.stream));
        if (weights_phase == 2) {
            // the ones of :430 in the rows this process owns: the ordered sums then start from 1.0 exactly as on
            // one GPU, and the all-reduce over the ranks adds exact zeros
            const int64_t b0 = (int64_t)P.row_lo * G * nch, b1 = (int64_t)P.row_hi * G * nch;
            LaunchScope ls("grid_fill");
            fill_kernel<<<ceil_div(b1 - b0, 256), 256, 0, c.stream>>>(binned + b0, b1 - b0, 1.0);
            PDSB_CUDA(cudaGetLastError());
        }
        PDSB_CHECK(scatter(1, npix, npix, nullptr, nullptr, binned));
        PDSB_CHECK(sum_columns(w_work, nuv, nf, 0, small + nf));
        PDSB_CUDA(cudaMemcpyAsync(binned_out, binned, (size_t)ncell * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(sumw_out, small + nf, (size_t)nf * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        unsigned long long hn = 0;
        PDSB_CUDA(cudaMemcpyAsync(&hn, d_nout, sizeof(hn), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
        if (n_outside) *n_outside = (int64_t)hn;
        return PDSB_OK;
    }

    // ---- optional re-weighting (:429-485) ----
    if (weighting != PDSB_WT_NATURAL) {
        uint32_t npix = weighting == PDSB_WT_SUPERUNIFORM ? 3u : (uint32_t)npixels;
        if (ext_weights) {
            // multi-GPU: the box sums were made per rank (weights phase), reduced, and handed back with the ones added
            PDSB_REQUIRE(c.grid_ext_ncell == ncell && (int)c.grid_ext_sumw.size() == nf, "pdsb_set_grid_reweight sizes");
            PDSB_CUDA(cudaMemcpyAsync(binned, c.grid_ext_binned, (size_t)ncell * sizeof(double), cudaMemcpyDeviceToDevice,
                                      c.stream));
        } else {
            {
                LaunchScope ls("grid_fill");
                fill_kernel<<<ceil_div(ncell, 256), 256, 0, c.stream>>>(binned, ncell, 1.0);     // numpy.ones :430
                PDSB_CUDA(cudaGetLastError());
            }
            PDSB_CHECK(scatter(1, npix, npix, nullptr, nullptr, binned));
        }
        const double *f2 = nullptr;
        if (weighting == PDSB_WT_ROBUST) {
            double *sumb2 = small, *sumw = small + nf, *f2w = small + 2 * nf;
            PDSB_CHECK(sum_columns(binned, (int64_t)G * G, nch, 1, sumb2));
            if (ext_weights)
                PDSB_CUDA(cudaMemcpyAsync(sumw, c.grid_ext_sumw.data(), (size_t)nf * sizeof(double), cudaMemcpyHostToDevice,
                                          c.stream));
            else
                PDSB_CHECK(sum_columns(w_work, nuv, nf, 0, sumw));
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

    // ---- normalisation (:525-533); imaging == 2 leaves the raw sums (multi-GPU: reduce first) ----
    if (imaging != 2) {
        double *wsum = small + 3 * nf;
        if (imaging) {
            PDSB_REQUIRE(nch <= nf, "channels");
            // per-channel sum of the weight map (binned is free now)
            if (nch != 1) wsum = binned;
            PDSB_CHECK(sum_columns(o_w, (int64_t)G * G, nch, 0, wsum));
        }
        LaunchScope ls("grid_normalise");
        grid_normalise_kernel<<<ceil_div(ncell, 256), 256, 0, c.stream>>>(o_re, o_im, o_w, ncell, nch, imaging ? 1 : 0,
                                                                         wsum);
        PDSB_CUDA(cudaGetLastError());
    }

    // ---- outputs ----
    auto give = [&](void *dst, const void *src, size_t bytes) -> int {
        if (out_kind == PDSB_DEVICE) {
            PDSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c.stream));
            return PDSB_OK;
        }
        return copy_d2h(dst, src, bytes);
    };
    if (!direct_out) {
        if (out_real) PDSB_CHECK(give(out_real, o_re, (size_t)ncell * sizeof(double)));
        if (out_imag) PDSB_CHECK(give(out_imag, o_im, (size_t)ncell * sizeof(double)));
        if (out_weights) PDSB_CHECK(give(out_weights, o_w, (size_t)ncell * sizeof(double)));
    }
    if (nvis > 0) {
        if (out_i) PDSB_CHECK(give(out_i, gi, (size_t)nvis * sizeof(uint32_t)));
        if (out_j) PDSB_CHECK(give(out_j, gj, (size_t)nvis * sizeof(uint32_t)));
        if (out_wmod) PDSB_CHECK(give(out_wmod, w_work, (size_t)nvis * sizeof(double)));
    }
    unsigned long long hn = 0;
    if (n_outside || out_kind == PDSB_HOST) {
        PDSB_CUDA(cudaMemcpyAsync(&hn, d_nout, sizeof(hn), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
        if (n_outside) *n_outside = (int64_t)hn;
    }
    return PDSB_OK;
}

int pdsb_grid(const double *u, const double *v, const double *freq, const double *real, const double *imag,
              const double *weights, int64_t nuv, int nf, int in_kind, int gridsize, double binsize,
              const double *uu, const double *vv, int convolution, int weighting, double robust, int npixels,
              int mode, int imaging, int deterministic, double *out_real, double *out_imag, double *out_weights,
              uint32_t *out_i, uint32_t *out_j, double *out_wmod, int out_kind, int64_t *n_outside)
{
    return grid_impl(u, v, freq, real, imag, weights, nuv, nf, in_kind, gridsize, binsize, uu, vv, convolution, weighting,
                     robust, npixels, mode, imaging, deterministic, out_real, out_imag, out_weights, out_i, out_j, out_wmod,
                     out_kind, n_outside, 0, nullptr, nullptr);
}

int pdsb_grid_weights_map(const double *u, const double *v, const double *freq, const double *real, const double *imag,
                          const double *weights, int64_t nuv, int nf, int in_kind, int gridsize, double binsize,
                          const double *uu, const double *vv, int weighting, int npixels, int mode, int deterministic,
                          int with_ones, double *binned_dev, double *sumw_host, int64_t *n_outside)
{
    return grid_impl(u, v, freq, real, imag, weights, nuv, nf, in_kind, gridsize, binsize, uu, vv, PDSB_CONV_PILLBOX,
                     weighting, 0.0, npixels, mode, 2, deterministic, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                     PDSB_DEVICE, n_outside, with_ones ? 2 : 1, binned_dev, sumw_host);
}

int pdsb_set_grid_reweight(const double *binned_dev, int64_t ncell, const double *sumw_host, int nf)
{
    Context &c = ctx();
    if (!binned_dev) {
        c.grid_ext_binned = nullptr;
        c.grid_ext_ncell = 0;
        c.grid_ext_sumw.clear();
        return PDSB_OK;
    }
    PDSB_REQUIRE(ncell > 0 && nf > 0 && sumw_host, "pdsb_set_grid_reweight arguments");
    c.grid_ext_binned = binned_dev;
    c.grid_ext_ncell = ncell;
    c.grid_ext_sumw.assign(sumw_host, sumw_host + nf);
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

int pdsb_average(const double *u, const double *v, const double *uvdist, const double *log_uvdist, const double *freq,
                 const double *real, const double *imag, const double *weights, int64_t nuv, int nf, int mfs,
                 double mfs_freq, int gridsize, double binsize, int radial, double log_min, double dtemp,
                 const double *centres, int spectral, double *out_u, double *out_v, double *out_real,
                 double *out_imag, double *out_weights, int64_t *n_out, int64_t *n_dropped)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(nuv >= 0 && nf > 0 && gridsize > 0, "sizes");
    PDSB_REQUIRE(radial >= 0 && radial <= 2, "radial");
    PDSB_REQUIRE(out_u && out_v && out_real && out_imag && out_weights && n_out, "outputs");
    PDSB_REQUIRE(radial == 2 || binsize > 0, "binsize");
    PDSB_REQUIRE(radial != 2 || (log_uvdist && dtemp != 0 && !mfs), "log bins need log10(uvdist), dtemp, no mfs");
    PDSB_REQUIRE(radial == 0 || centres, "radial bins need their centres");
    PDSB_REQUIRE(!mfs || mfs_freq > 0, "mfs frequency");
    const int G = gridsize, Gj = radial ? 1 : gridsize;
    const int nfe = mfs ? 1 : nf;
    const int64_t nrow = mfs ? nuv * nf : nuv;
    const int nch = spectral ? nfe : 1;
    const int64_t npos = (int64_t)G * Gj, ncell = npos * nch;
    PDSB_REQUIRE(ncell < (int64_t)KEY_DEAD, "grid too large");
    const int64_t nvis = nuv * nf;
    PDSB_REQUIRE(nvis < (int64_t)1 << 32, "nuv*nf must fit 32 bits");
    *n_out = 0;
    if (n_dropped) *n_dropped = 0;

    // maps: re, im, w, u, v, scratch-w, pos_u, pos_v (per position), flags + rank
    const size_t map_bytes = (size_t)ncell * 6 * sizeof(double) + (size_t)npos * 2 * sizeof(double) +
                             (size_t)npos * 2 * sizeof(uint32_t) + (size_t)(npos / SCAN_SEG + 16) * sizeof(uint32_t) + 1024;
    PDSB_CHECK(c.stage_c.ensure(map_bytes));
    double *m_re = c.stage_c.as<double>(), *m_im = m_re + ncell, *m_w = m_im + ncell, *m_u = m_w + ncell,
           *m_v = m_u + ncell, *m_w2 = m_v + ncell, *pos_u = m_w2 + ncell, *pos_v = pos_u + npos;
    uint32_t *flag = reinterpret_cast<uint32_t *>(pos_v + npos), *rank = flag + npos, *seg = rank + npos;
    PDSB_CUDA(cudaMemsetAsync(m_re, 0, (size_t)ncell * 6 * sizeof(double), c.stream));

    // inputs (host) -> stage_a; per-row work arrays -> stage_b
    AvgParams P;
    {
        const size_t nd = (size_t)2 * nuv + (radial == 2 ? nuv : 0) + (mfs ? 0 : nuv) + nf + (size_t)3 * nvis + (radial ? G : 0);
        PDSB_CHECK(c.stage_a.ensure(nd * sizeof(double) + 64));
        double *p = c.stage_a.as<double>();
        auto put = [&](const double *src, size_t n, const double **dst) -> int {
            *dst = p;
            if (n == 0 || !src) return PDSB_OK;
            PDSB_CHECK(copy_h2d(p, src, n * sizeof(double)));
            p += n;
            return PDSB_OK;
        };
        if (nvis > 0) PDSB_REQUIRE(u && v && freq && real && imag && weights && (mfs || uvdist), "input arrays");
        PDSB_CHECK(put(u, nuv, &P.u));
        PDSB_CHECK(put(v, nuv, &P.v));
        PDSB_CHECK(put(mfs ? nullptr : uvdist, nuv, &P.uvdist));
        PDSB_CHECK(put(radial == 2 ? log_uvdist : nullptr, nuv, &P.log_uvdist));
        PDSB_CHECK(put(freq, nf, &P.freq));
        PDSB_CHECK(put(real, nvis, &P.re));
        PDSB_CHECK(put(imag, nvis, &P.im));
        PDSB_CHECK(put(weights, nvis, &P.w));
        const double *dc = nullptr;
        PDSB_CHECK(put(radial ? centres : nullptr, G, &dc));
        centres = dc;
    }
    PDSB_CHECK(c.stage_b.ensure((size_t)nrow * (sizeof(uint32_t) + 2 * sizeof(double)) + 256));
    P.row_u = c.stage_b.as<double>();
    P.row_v = P.row_u + nrow;
    P.row_cell = reinterpret_cast<uint32_t *>(P.row_v + nrow);
    unsigned long long *d_drop = reinterpret_cast<unsigned long long *>(seg + (npos / SCAN_SEG + 8));
    d_drop = reinterpret_cast<unsigned long long *>(((uintptr_t)d_drop + 7) & ~(uintptr_t)7);
    PDSB_CUDA(cudaMemsetAsync(d_drop, 0, sizeof(unsigned long long), c.stream));
    P.nuv = nuv; P.nf = nf; P.mfs = mfs; P.mfs_inv_freq = mfs ? 1. / mfs_freq : 0.0; P.nrow = nrow; P.nfe = nfe;
    P.G = G; P.Gj = Gj; P.nch = nch; P.spectral = spectral; P.radial = radial;
    P.binsize = binsize; P.half = (G % 2 == 0) ? G / 2. : (G - 1) / 2.; P.log_min = log_min; P.dtemp = dtemp;

    if (nvis > 0) {
        {
            LaunchScope ls("avg_prep");
            avg_prep_kernel<<<ceil_div(nrow, 256), 256, 0, c.stream>>>(P, d_drop);
            PDSB_CUDA(cudaGetLastError());
        }
        const int nbits = bits_for((uint64_t)ncell);
        const int sort_bits = ((nbits + 7) / 8) * 8 > 32 ? 32 : ((nbits + 7) / 8) * 8;
        const int64_t per_batch = (int64_t)1 << 27;
        for (int64_t first = 0; first < nvis; first += per_batch) {
            const int64_t n = std::min(per_batch, nvis - first);
            const size_t nh = (size_t)256 * ceil_div(n, RS_TILE);
            PDSB_CHECK(c.stage_d.ensure((size_t)n * 4 * sizeof(uint32_t) + (nh + nh / SCAN_SEG + 16) * sizeof(uint32_t) + 1024));
            PDSB_CHECK(c.stage_e.ensure((size_t)n * 3 * sizeof(double)));
            uint32_t *k0 = c.stage_d.as<uint32_t>(), *v0 = k0 + n, *k1 = v0 + n, *v1 = k1 + n, *hist = v1 + n;
            double *va = c.stage_e.as<double>(), *vb = va + n, *vw = vb + n;
            {
                LaunchScope ls("avg_emit_keys");
                avg_emit_keys_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(P, first, n, k0, v0);
                PDSB_CUDA(cudaGetLastError());
            }
            uint32_t *ko, *vo;
            SortBufs sb{k0, v0, k1, v1, hist};
            PDSB_CHECK(radix_sort(sb, n, sort_bits, &ko, &vo));
            for (int which = 0; which < (radial ? 1 : 2); which++) {
                {
                    LaunchScope ls("avg_values");
                    avg_values_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(P, which, first, vo, n, va, vb, vw);
                    PDSB_CUDA(cudaGetLastError());
                }
                LaunchScope ls("grid_ordered_sum");
                grid_ordered_sum_kernel<<<ceil_div(ncell, 128), 128, 0, c.stream>>>(
                    0, ncell, ko, n, va, vb, vw, which ? m_u : m_re, which ? m_v : m_im, which ? m_w2 : m_w, nullptr,
                    nullptr);
                PDSB_CUDA(cudaGetLastError());
            }
        }
    }
    {
        LaunchScope ls("avg_finish");
        avg_finish_kernel<<<ceil_div(npos, 128), 128, 0, c.stream>>>(npos, nch, radial ? 1 : 0, m_re, m_im, m_w, m_u, m_v,
                                                                     pos_u, pos_v, flag);
        PDSB_CUDA(cudaMemcpyAsync(rank, flag, (size_t)npos * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
        const int nseg = ceil_div(npos, SCAN_SEG);
        rs_scan_seg_kernel<<<nseg, 1024, 0, c.stream>>>(rank, npos, seg);
        rs_scan_totals_kernel<<<1, 1024, 0, c.stream>>>(seg, nseg);
        rs_scan_add_kernel<<<nseg, 1024, 0, c.stream>>>(rank, npos, seg);
        PDSB_CUDA(cudaGetLastError());
    }
    uint32_t last[2] = {0, 0};
    unsigned long long hdrop = 0;
    PDSB_CUDA(cudaMemcpyAsync(&last[0], rank + npos - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaMemcpyAsync(&last[1], flag + npos - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaMemcpyAsync(&hdrop, d_drop, sizeof(hdrop), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    const int64_t ngood = (int64_t)last[0] + last[1];
    *n_out = ngood;
    if (n_dropped) *n_dropped = (int64_t)hdrop;
    if (ngood == 0) return PDSB_OK;
    // compact into the sort scratch (free now), then to the caller's host arrays
    PDSB_CHECK(c.stage_e.ensure((size_t)ngood * (2 + 3 * nch) * sizeof(double)));
    double *o_u = c.stage_e.as<double>(), *o_v = o_u + ngood, *o_re = o_v + ngood, *o_im = o_re + ngood * nch,
           *o_w = o_im + ngood * nch;
    {
        LaunchScope ls("avg_compact");
        avg_compact_kernel<<<ceil_div(ncell, 256), 256, 0, c.stream>>>(npos, nch, radial ? 1 : 0, G, rank, flag, m_re, m_im, m_w,
                                                                      pos_u, pos_v, centres, o_u, o_v, o_re, o_im, o_w);
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CUDA(cudaMemcpyAsync(out_u, o_u, (size_t)ngood * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaMemcpyAsync(out_v, o_v, (size_t)ngood * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaMemcpyAsync(out_real, o_re, (size_t)ngood * nch * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaMemcpyAsync(out_imag, o_im, (size_t)ngood * nch * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaMemcpyAsync(out_weights, o_w, (size_t)ngood * nch * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    return PDSB_OK;
}

int pdsb_center(const double *u, const double *v, const double *freq, const double *real, const double *imag,
                int64_t nuv, int nf, double mean_freq, double x0_rad, double y0_rad, int kind, double *out_real,
                double *out_imag)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(nuv >= 0 && nf > 0 && mean_freq > 0, "sizes/mean_freq");
    if (nuv == 0) return PDSB_OK;
    PDSB_REQUIRE(u && v && freq && real && imag && out_real && out_imag, "arrays");
    const int64_t n = nuv * nf;
    const double *du = u, *dv = v, *df = freq, *dre = real, *dim = imag;
    double *ore = out_real, *oim = out_imag;
    if (kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_a.ensure((size_t)(2 * nuv + nf + 2 * n) * sizeof(double) + 64));
        double *p = c.stage_a.as<double>();
        auto put = [&](const double *src, size_t cnt, const double **dst) -> int {
            PDSB_CHECK(copy_h2d(p, src, cnt * sizeof(double)));
            *dst = p;
            p += cnt;
            return PDSB_OK;
        };
        PDSB_CHECK(put(u, nuv, &du));
        PDSB_CHECK(put(v, nuv, &dv));
        PDSB_CHECK(put(freq, nf, &df));
        PDSB_CHECK(put(real, n, &dre));
        PDSB_CHECK(put(imag, n, &dim));
        PDSB_CHECK(c.stage_b.ensure((size_t)2 * n * sizeof(double)));
        ore = c.stage_b.as<double>();
        oim = ore + n;
    }
    {
        LaunchScope ls("center");
        center_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(du, dv, df, dre, dim, nuv, nf, mean_freq, x0_rad, y0_rad, ore, oim);
        PDSB_CUDA(cudaGetLastError());
    }
    if (kind == PDSB_HOST) {
        PDSB_CUDA(cudaMemcpyAsync(out_real, ore, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(out_imag, oim, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}

// Multi-GPU bit-exact gridding: restrict the main scatter of later pdsb_grid calls
// to output rows [row_lo, row_hi); (0, 0) lifts the restriction.
int pdsb_set_grid_band(int row_lo, int row_hi)
{
    PDSB_REQUIRE((row_lo == 0 && row_hi == 0) || (row_lo >= 0 && row_hi > row_lo), "row band");
    ctx().grid_row_lo = row_lo;
    ctx().grid_row_hi = row_hi;
    return PDSB_OK;
}

// Normalisation step of grid() on DEVICE maps [G*G, nch] (after a multi-GPU reduction of raw sums).
int pdsb_grid_normalise(double *real, double *imag, double *weights, int gridsize, int nch, int imaging)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(real && imag && weights && gridsize > 0 && nch > 0, "arguments");
    const int64_t ncell = (int64_t)gridsize * gridsize * nch;
    PDSB_CHECK(c.stage_b.ensure((size_t)(nch + SUM_BLOCKS * nch + 8) * sizeof(double)));
    double *wsum = c.stage_b.as<double>(), *part = wsum + nch;
    if (imaging) {
        const int64_t nrow = (int64_t)gridsize * gridsize;
        const int nb = (int)std::min<int64_t>(SUM_BLOCKS, std::max<int64_t>(1, (nrow + 255) / 256));
        LaunchScope ls("grid_sum");
        strided_sum_kernel<<<dim3(nb, nch), 256, 0, c.stream>>>(weights, nrow, nch, 0, part);
        strided_sum_final_kernel<<<ceil_div(nch, 256), 256, 0, c.stream>>>(part, nb, nch, wsum);
        PDSB_CUDA(cudaGetLastError());
    }
    LaunchScope ls("grid_normalise");
    grid_normalise_kernel<<<ceil_div(ncell, 256), 256, 0, c.stream>>>(real, imag, weights, ncell, nch, imaging ? 1 : 0, wsum);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

}  // extern "C"
