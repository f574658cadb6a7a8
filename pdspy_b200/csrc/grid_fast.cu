// Fast (non-bit-exact) mode of grid(): the scatter loop of pdspy/interferometry/libinterferometry.pyx:489-521
// as  counting sort by 8x8-cell uv tile  ->  contiguous 48-byte records  ->  register accumulation per tile
// ->  shared-memory region  ->  one fp64 atomic per region cell and map.
//
// Pipeline (every kernel streams its inputs once; no gathers, no multi-pass radix sort):
//   gf_hist_kernel     u, v -> home tile of every visibility, histogram over the tiles (warp-aggregated atomics)
//   gf_scan_*          exclusive scan of (count, work items) per tile, packed in one 64-bit word
//   gf_items_kernel    work list: every tile's run cut into items of <= GF_ITEM visibilities
//   gf_scatter_kernel  second pass over the inputs: 48-byte record [pos_u, w][w re, w im][pos_v, (gi | gj << 32)]
//                      written at the visibility's slot of its tile run (warp-aggregated cursor atomics; the L2 merges
//                      the 64 K write streams, so DRAM sees full lines)
//   gf_tile_kernel     persistent warps, one work item each:
//        * the item's records are bucketed by home ROW inside the tile (8 buckets, ballot counting sort);
//        * factor phase, one lane per visibility: the 2 x WIDTH one-dimensional kernel factors (the exp*sinc kernel
//          is separable) in Horner form, staged in shared memory;
//        * accumulate phase, one half-warp per visibility: lane = region column, registers = the WIDTH rows of the
//          visibility's footprint x 3 maps.  Because the bucket fixes the rows, all WIDTH x 3 accumulators of a
//          lane are live for every visibility (the round-1 kernel kept all 13 region rows in registers and issued
//          39 DFMAs per lane and visibility for 18 useful ones);
//        * when the bucket changes the accumulators are added to the warp's private region in shared memory (plain
//          read-modify-write, no atomics: the two half-warps take turns), and at the end of the item the region goes
//          to the map with one fp64 atomic per cell and map.
// Summation order differs from the reference's (k, n) loop, so results agree to rounding (1e-15 relative), not bit
// for bit; the ordered mode of grid.cu is the bit-exact one.
#include "grid.cuh"
#include <algorithm>

namespace pdsb {

constexpr int GF_ITEM = 256;        // visibilities per work item (one warp)
constexpr int GF_WARPS = 4;         // warps per CTA, each on its own item
constexpr int GF_SCAN_THREADS = 1024;
constexpr int GF_SCAN_PER = 4;
constexpr int GF_SCAN_SEG = GF_SCAN_THREADS * GF_SCAN_PER;
constexpr uint32_t GF_DEAD = 0xffffffffu;

struct GfItem {
    uint32_t begin, end, key, pad;
};

struct GfVis {
    double pu, pv;          // u f / mean_f, v f / mean_f as the reference forms them (:509-511)
    uint32_t gi, gj;
    bool good;
};

__device__ __forceinline__ GfVis gf_vis(const GridParams &P, int64_t idx)
{
    GfVis x;
    const int64_t k = idx / P.nf;
    const double f = P.freq[idx % P.nf];
    x.pu = __dmul_rn(__dmul_rn(P.u[k], f), P.inv_freq);
    x.pv = __dmul_rn(__dmul_rn(P.v[k], f), P.inv_freq);
    x.gi = np_f64_to_u32(__dadd_rn(__ddiv_rn(x.pu, P.binsize), P.half));        // :388-403
    x.gj = np_f64_to_u32(__dadd_rn(__ddiv_rn(x.pv, P.binsize), P.half));
    x.good = x.gi < (uint32_t)P.G && x.gj < (uint32_t)P.G;                        // :421-423
    return x;
}

__device__ __forceinline__ uint32_t gf_key(const GridParams &P, const GfVis &x, int64_t idx, uint32_t tg)
{
    const uint32_t tile = (x.gj >> 3) * tg + (x.gi >> 3);
    return (P.spectral ? (uint32_t)(idx % P.nf) : 0u) * tg * tg + tile;
}

// ---- pass 1: histogram over the tiles ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) gf_hist_kernel(GridParams P, uint32_t tg, uint32_t *__restrict__ hist,
                                                      unsigned long long *n_outside)
{
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t key = GF_DEAD;
    bool outside = false;
    if (idx < P.nuv * P.nf) {
        const GfVis x = gf_vis(P, idx);
        if (x.good) key = gf_key(P, x, idx, tg);
        else outside = true;
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (key != GF_DEAD && (int)(__ffs(peers) - 1) == lane) atomicAdd(hist + key, (uint32_t)__popc(peers));
    if (n_outside) {
        const uint32_t m = __ballot_sync(0xffffffffu, outside);
        if (lane == 0 && m) atomicAdd(n_outside, (unsigned long long)__popc(m));
    }
}

// ---- exclusive scan of packed (count | items << 40) ---------------------------------------------------------
__device__ __forceinline__ uint64_t gf_pack(uint32_t cnt)
{
    return (uint64_t)cnt | ((uint64_t)((cnt + GF_ITEM - 1) / GF_ITEM) << 40);
}

__device__ __forceinline__ uint64_t gf_block_excl_scan(uint64_t x, uint64_t *warp_sums, uint64_t *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        warp_sums[lane] = s;
    }
    __syncthreads();
    const uint64_t base = wid ? warp_sums[wid - 1] : 0ull;
    *total = warp_sums[31];
    __syncthreads();
    return base + inc - x;
}

// excl[e] = exclusive scan within the segment; seg_tot[segment] = segment total.  n entries (the caller scans
// nkeys + 1 entries so that the last one holds the totals).
__global__ void __launch_bounds__(GF_SCAN_THREADS) gf_scan_seg_kernel(const uint32_t *__restrict__ hist, uint32_t nkeys,
                                                                      uint64_t *__restrict__ excl,
                                                                      uint64_t *__restrict__ seg_tot)
{
    __shared__ uint64_t warp_sums[32];
    const uint32_t e0 = blockIdx.x * GF_SCAN_SEG + threadIdx.x * GF_SCAN_PER;
    uint64_t v[GF_SCAN_PER], tsum = 0;
#pragma unroll
    for (int q = 0; q < GF_SCAN_PER; q++) {
        v[q] = (e0 + q < nkeys) ? gf_pack(hist[e0 + q]) : 0ull;
        tsum += v[q];
    }
    uint64_t total;
    uint64_t ex = gf_block_excl_scan(tsum, warp_sums, &total);
#pragma unroll
    for (int q = 0; q < GF_SCAN_PER; q++) {
        if (e0 + q <= nkeys) excl[e0 + q] = ex;
        ex += v[q];
    }
    if (threadIdx.x == 0) seg_tot[blockIdx.x] = total;
}

// single block: exclusive scan of the segment totals in place
__global__ void __launch_bounds__(GF_SCAN_THREADS) gf_scan_totals_kernel(uint64_t *__restrict__ t, int nseg)
{
    __shared__ uint64_t warp_sums[32];
    uint64_t carry = 0;
    for (int base = 0; base < nseg; base += GF_SCAN_THREADS) {
        const int e = base + threadIdx.x;
        const uint64_t x = e < nseg ? t[e] : 0ull;
        uint64_t total;
        const uint64_t ex = gf_block_excl_scan(x, warp_sums, &total);
        if (e < nseg) t[e] = carry + ex;
        carry += total;
    }
}

// global exclusive prefix of entry e: low 40 bits = visibilities before tile e, high bits = work items before it
__device__ __forceinline__ uint64_t gf_prefix(const uint64_t *__restrict__ excl, const uint64_t *__restrict__ seg_off,
                                              uint32_t e)
{
    return excl[e] + seg_off[e / GF_SCAN_SEG];
}
constexpr uint64_t GF_LOW40 = (1ull << 40) - 1ull;

// ---- work list ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gf_items_kernel(const uint64_t *__restrict__ excl,
                                                       const uint64_t *__restrict__ seg_off, uint32_t nkeys,
                                                       GfItem *__restrict__ items)
{
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    const uint32_t total = (uint32_t)(gf_prefix(excl, seg_off, nkeys) >> 40);
    if (t >= total) return;
    // the tile whose item range holds t: largest key with items_before(key) <= t  (empty tiles repeat the value of
    // their successor, so the search lands on the non-empty one)
    uint32_t lo = 0, hi = nkeys;                       // items_before(lo) <= t < items_before(hi)
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((uint32_t)(gf_prefix(excl, seg_off, mid) >> 40) <= t) lo = mid;
        else hi = mid;
    }
    const uint64_t p0 = gf_prefix(excl, seg_off, lo), p1 = gf_prefix(excl, seg_off, lo + 1);
    const uint32_t j = t - (uint32_t)(p0 >> 40);
    const uint32_t start = (uint32_t)(p0 & GF_LOW40), stop = (uint32_t)(p1 & GF_LOW40);
    GfItem it;
    it.begin = start + j * GF_ITEM;
    it.end = it.begin + GF_ITEM < stop ? it.begin + GF_ITEM : stop;
    it.key = lo;
    it.pad = 0;
    items[t] = it;
}

// ---- pass 2: records to their tile runs ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) gf_scatter_kernel(GridParams P, uint32_t tg, const double *__restrict__ w_src,
                                                         const uint64_t *__restrict__ excl,
                                                         const uint64_t *__restrict__ seg_off,
                                                         uint32_t *__restrict__ fill, double2 *__restrict__ rec)
{
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t key = GF_DEAD;
    GfVis x;
    if (idx < P.nuv * P.nf) {
        x = gf_vis(P, idx);
        if (x.good) key = gf_key(P, x, idx, tg);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (key != GF_DEAD && leader == lane) base = atomicAdd(fill + key, (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (key == GF_DEAD) return;
    const uint32_t pos = (uint32_t)(gf_prefix(excl, seg_off, key) & GF_LOW40) + base + __popc(peers & ((1u << lane) - 1u));
    double w;
    const double re = P.re[idx], im = P.im[idx];
    if (w_src) w = w_src[idx];
    else {
        w = P.w_in[idx];                                    // :351-353
        w = w < 0 ? 0.0 : w;
        if (re == 0 && im == 0) w = 0.0;
    }
    double2 *r = rec + 3 * (size_t)pos;
    r[0] = make_double2(x.pu, w);
    r[1] = make_double2(re * w, im * w);
    r[2] = make_double2(x.pv, __longlong_as_double((long long)((unsigned long long)x.gi | ((unsigned long long)x.gj << 32))));
}

// ---- convolution kernel factors, Horner form ----------------------------------------------------------------
// exp_sinc is separable: exp_sinc(u, v) = g(u) g(v) / norm, g(x) = sinc(x/1.55) exp(-(x/2.52)^2) inside |x| < 3,
// with the reference's degree-16 Taylor polynomial for sinc and degree-5 polynomial for exp (:547-574).  Same
// polynomials as k_exp_sinc (grid.cuh), evaluated by Horner's rule: 1e-16-level differences from the reference's
// term-by-term form, below this mode's own summation-order noise.
__device__ __forceinline__ double gf_exp_sinc_1d(double x)
{
    if (fabs(x) >= 3.0) return 0.;
    const double xp = x * (1. / 1.55) * 3.14159265358979323846;
    const double y = xp * xp;
    double s = 1. / 355687428096000.;
    s = fma(s, y, -1. / 1307674368000.);
    s = fma(s, y, 1. / 6227020800.);
    s = fma(s, y, -1. / 39916800.);
    s = fma(s, y, 1. / 362880.);
    s = fma(s, y, -1. / 5040.);
    s = fma(s, y, 1. / 120.);
    s = fma(s, y, -1. / 6.);
    s = fma(s, y, 1.);
    const double a = x * (1. / 2.52);
    const double z = -(a * a);
    double e = 1. / 120.;
    e = fma(e, z, 1. / 24.);
    e = fma(e, z, 1. / 6.);
    e = fma(e, z, 0.5);
    e = fma(e, z, 1.);
    e = fma(e, z, 1.);
    return s * e;
}

// ---- tile kernel ------------------------------------------------------------------------------------------------
// WIDTH = footprint width lo + hi + 1 (pillbox 3, exp*sinc 6, box sums 3 / 5 / 7); MODE 0 = main sums (three maps),
// MODE 1 = box sums of the weights (one map, all factors 1).
template <int WIDTH, int MODE>
__global__ void __launch_bounds__(GF_WARPS * 32) gf_tile_kernel(GridParams P, int lo, uint32_t tg,
                                                                const double2 *__restrict__ rec,
                                                                const GfItem *__restrict__ items,
                                                                const uint64_t *__restrict__ excl,
                                                                const uint64_t *__restrict__ seg_off, uint32_t nkeys,
                                                                uint32_t *__restrict__ work_counter, double *out_re,
                                                                double *out_im, double *out_w)
{
    constexpr int SIDE = 8 + WIDTH - 1;                 // region side (<= 14)
    constexpr int NMAP = MODE == 0 ? 3 : 1;
    constexpr int FUS = 33;                             // padded lane stride of the FU columns (bank-conflict free)
    __shared__ double s_reg[GF_WARPS][NMAP][SIDE][16];  // the warp's private region: [map][row][column]
    __shared__ double s_fu[GF_WARPS][WIDTH][FUS];       // column factors [slot][staged visibility]
    __shared__ __align__(16) double s_b[GF_WARPS][32][12];   // per staged visibility: FV[0..6], w, w re, w im, meta
    __shared__ uint16_t s_ord[GF_WARPS][GF_ITEM];       // the item's visibilities in bucket (home row) order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, c = lane & 15;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t nitems = (uint32_t)(gf_prefix(excl, seg_off, nkeys) >> 40);
    double(*reg)[SIDE][16] = s_reg[warp];
    double(*fu)[FUS] = s_fu[warp];
    double(*sb)[12] = s_b[warp];
    uint16_t *ord = s_ord[warp];

    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(work_counter, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    while (t < nitems) {
        const GfItem item = items[t];
        // next item's ticket: the atomic's latency hides behind this item's work
        uint32_t t_next = 0;
        if (lane == 0) t_next = atomicAdd(work_counter, 1u);
        const uint32_t chan = item.key / (tg * tg), tile = item.key % (tg * tg);
        const int tl = (int)(tile / tg), tm = (int)(tile % tg);
        const int n = (int)(item.end - item.begin);
        const double2 *irec = rec + 3 * (size_t)item.begin;

        // zero the region
        {
            double *z = &reg[0][0][0];
            for (int i = lane; i < NMAP * SIDE * 16; i += 32) z[i] = 0.0;
        }
        // bucket = home row inside the tile: counting sort of the item's visibilities with ballots
        int bk[GF_ITEM / 32];                          // (chunks past the item's end are skipped: uniform branches)
        int cnt[8];
#pragma unroll
        for (int b = 0; b < 8; b++) cnt[b] = 0;
#pragma unroll
        for (int q = 0; q < GF_ITEM / 32; q++) {
            const int i = q * 32 + lane;
            bk[q] = 8;
            if (q * 32 >= n) continue;
            if (i < n) {
                const unsigned long long ij = (unsigned long long)__double_as_longlong(irec[3 * i + 2].y);
                bk[q] = (int)(uint32_t)(ij >> 32) - tl * 8;
            }
#pragma unroll
            for (int b = 0; b < 8; b++) cnt[b] += __popc(__ballot_sync(0xffffffffu, bk[q] == b));
        }
        {
            int run = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int x = cnt[b];
                cnt[b] = run;
                run += x;
            }
        }
#pragma unroll
        for (int q = 0; q < GF_ITEM / 32; q++) {
            if (q * 32 >= n) continue;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const uint32_t m = __ballot_sync(0xffffffffu, bk[q] == b);
                if (bk[q] == b) ord[cnt[b] + __popc(m & lt)] = (uint16_t)(q * 32 + lane);
                cnt[b] += __popc(m);
            }
        }
        __syncwarp();

        double acc[NMAP][WIDTH];
#pragma unroll
        for (int m = 0; m < NMAP; m++)
#pragma unroll
            for (int r = 0; r < WIDTH; r++) acc[m][r] = 0.0;
        int cb = -1;                                     // bucket the accumulators of this half-warp belong to
        auto flush = [&]() {
#pragma unroll
            for (int m = 0; m < NMAP; m++)
#pragma unroll
                for (int r = 0; r < WIDTH; r++) {
                    reg[m][cb + r][c] += acc[m][r];
                    acc[m][r] = 0.0;
                }
        };

        for (int p0 = 0; p0 < n; p0 += 32) {
            const int nst = n - p0 < 32 ? n - p0 : 32;
            // ---- factor phase: lane = visibility ----
            if (lane < nst) {
                const double2 *r = irec + 3 * (size_t)ord[p0 + lane];
                const double2 c0 = r[0], c1 = r[1], c2 = r[2];
                const unsigned long long ij = (unsigned long long)__double_as_longlong(c2.y);
                const int gi = (int)(uint32_t)ij, gj = (int)(uint32_t)(ij >> 32);
#pragma unroll
                for (int o = 0; o < WIDTH; o++) {
                    const int cu = gi - lo + o, cv = gj - lo + o;
                    double gu = 0.0, gv = 0.0;
                    if (MODE == 0) {
                        if (cu >= 0 && cu < P.G) {
                            const double d = (c0.x - P.uu[cu]) * P.inv_binsize;
                            // 1/norm goes with the u factor
                            gu = P.conv ? gf_exp_sinc_1d(d) * (1. / 2.350016262343186) : (fabs(d) >= 0.5 ? 0.0 : 1.0);
                        }
                        if (cv >= 0 && cv < P.G) {
                            const double d = (c2.x - P.vv[cv]) * P.inv_binsize;
                            gv = P.conv ? gf_exp_sinc_1d(d) : (fabs(d) >= 0.5 ? 0.0 : 1.0);
                        }
                    } else {
                        gu = (cu >= 0 && cu < P.G) ? 1.0 : 0.0;
                        gv = (cv >= 0 && cv < P.G) ? 1.0 : 0.0;
                    }
                    fu[o][lane] = gu;
                    sb[lane][o] = gv;
                }
                sb[lane][7] = c0.y;
                sb[lane][8] = c1.x;
                sb[lane][9] = c1.y;
                sb[lane][10] = __longlong_as_double((long long)((gi - tm * 8) | ((gj - tl * 8) << 8)));
            }
            __syncwarp();
            // ---- accumulate phase: half-warp = visibility, lane = region column ----
            for (int t0 = 0; t0 < nst; t0 += 2) {
                const int v = t0 + half;
                const bool valid = v < nst;
                const double *vb = sb[valid ? v : 0];
                const int meta = (int)__double_as_longlong(vb[10]);
                const int di = meta & 0xff, b = meta >> 8;
                const bool need = valid && cb >= 0 && b != cb;
                if (__any_sync(0xffffffffu, need)) {
                    if (need && half == 0) flush();
                    __syncwarp();
                    if (need && half == 1) flush();
                    __syncwarp();
                }
                if (valid) {
                    cb = b;
                    const int o = c - di;
                    const double f = (o >= 0 && o < WIDTH) ? fu[o][v] : 0.0;
                    const double xw = f * vb[7];
                    if (MODE == 0) {
                        const double xr = f * vb[8], xi = f * vb[9];
#pragma unroll
                        for (int r = 0; r < WIDTH; r++) {
                            const double fv = vb[r];
                            acc[0][r] = fma(fv, xw, acc[0][r]);
                            acc[1][r] = fma(fv, xr, acc[1][r]);
                            acc[2][r] = fma(fv, xi, acc[2][r]);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < WIDTH; r++) acc[0][r] = fma(vb[r], xw, acc[0][r]);
                    }
                }
            }
            __syncwarp();
        }
        // the two half-warps' last accumulators, then the region to the map
        if (cb >= 0 && half == 0) flush();
        __syncwarp();
        if (cb >= 0 && half == 1) flush();
        __syncwarp();
        {
            const int m0 = tm * 8 - lo + c;
            for (int rr0 = 0; rr0 < SIDE; rr0 += 2) {
                const int rr = rr0 + half;
                const int l = tl * 8 - lo + rr;
                if (rr >= SIDE || c >= SIDE || l < 0 || m0 < 0 || l >= P.G || m0 >= P.G) continue;
                const int64_t cell = ((int64_t)l * P.G + m0) * P.nch + chan;
                const double vw = reg[0][rr][c];
                if (MODE == 0) {
                    const double vr = reg[1][rr][c], vi = reg[2][rr][c];
                    if (vw != 0.0) atomicAdd(out_w + cell, vw);
                    if (vr != 0.0) atomicAdd(out_re + cell, vr);
                    if (vi != 0.0) atomicAdd(out_im + cell, vi);
                } else if (vw != 0.0)
                    atomicAdd(out_w + cell, vw);
            }
        }
        __syncwarp();
        t = __shfl_sync(0xffffffffu, t_next, 0);
    }
}

// wide footprints (box sums with npixels > 3): plain atomics, one thread per (visibility, footprint cell)
__global__ void __launch_bounds__(256) gf_wide_atomic_kernel(GridParams P, const double *__restrict__ w_src, uint32_t lo,
                                                             uint32_t hi, int fp, double *out_w)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= P.nuv * P.nf * fp) return;
    const int64_t idx = t / fp;
    const int f = (int)(t % fp);
    const GfVis x = gf_vis(P, idx);
    if (!x.good) return;
    const int side = (int)(lo + hi + 1);
    const long long l = (long long)x.gj - lo + f / side, m = (long long)x.gi - lo + f % side;
    if (l < 0 || m < 0 || l >= P.G || m >= P.G) return;
    const int64_t cell = ((int64_t)l * P.G + m) * P.nch + (P.spectral ? idx % P.nf : 0);
    atomicAdd(out_w + cell, w_src[idx]);
}

template <int WIDTH, int MODE>
static void launch_tile(const GridParams &P, int lo, uint32_t tg, const double2 *rec, const GfItem *items,
                        const uint64_t *excl, const uint64_t *seg_off, uint32_t nkeys, uint32_t *counter, double *t_re,
                        double *t_im, double *t_w)
{
    Context &c = ctx();
    static int per_sm = 0;
    if (!per_sm) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gf_tile_kernel<WIDTH, MODE>, GF_WARPS * 32, 0);
        if (per_sm < 1) per_sm = 1;
    }
    gf_tile_kernel<WIDTH, MODE><<<c.sm_count * per_sm, GF_WARPS * 32, 0, c.stream>>>(P, lo, tg, rec, items, excl, seg_off,
                                                                                    nkeys, counter, t_re, t_im, t_w);
}

int grid_fast_scatter(const GridParams &P, int smode, uint32_t lo, uint32_t hi, const double *w_src, double *t_re,
                      double *t_im, double *t_w, unsigned long long *n_outside)
{
    Context &c = ctx();
    const int64_t nvis = P.nuv * P.nf;
    if (nvis == 0) return PDSB_OK;
    const int width = (int)(lo + hi + 1);
    if (smode == 1 && width > 7) {
        PDSB_REQUIRE(w_src != nullptr, "box sums need the prepared weights");
        const int fp = width * width;
        LaunchScope ls("grid_scatter_atomic");
        gf_wide_atomic_kernel<<<ceil_div(nvis * fp, 256), 256, 0, c.stream>>>(P, w_src, lo, hi, fp, t_w);
        PDSB_CUDA(cudaGetLastError());
        return PDSB_OK;
    }
    PDSB_REQUIRE((smode == 0 && (width == 3 || width == 6)) || (smode == 1 && (width == 3 || width == 5 || width == 7)),
                 "footprint width of the fast gridding mode");
    const uint32_t tg = ((uint32_t)P.G + 7u) >> 3;
    const uint64_t nkeys64 = (uint64_t)tg * tg * (uint64_t)P.nch;
    PDSB_REQUIRE(nkeys64 < (1ull << 31), "too many uv tiles");
    const uint32_t nkeys = (uint32_t)nkeys64;
    const int nseg = ceil_div((int64_t)nkeys + 1, GF_SCAN_SEG);
    const uint64_t max_items = std::min<uint64_t>(nkeys64, (uint64_t)nvis) + (uint64_t)nvis / GF_ITEM + 2;
    PDSB_REQUIRE(max_items < (1ull << 31), "too many gridding work items");

    // scratch: stage_d = hist | fill | counter | excl | seg_off | items ; stage_e = records
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t o_hist = 0, o_fill = o_hist + up((size_t)nkeys * 4), o_cnt = o_fill + up((size_t)nkeys * 4),
                 o_excl = o_cnt + 256, o_seg = o_excl + up(((size_t)nseg * GF_SCAN_SEG + 1) * 8),
                 o_items = o_seg + up((size_t)nseg * 8), total = o_items + up((size_t)max_items * sizeof(GfItem));
    PDSB_CHECK(c.stage_d.ensure(total));
    PDSB_CHECK(c.stage_e.ensure((size_t)nvis * 3 * sizeof(double2)));
    unsigned char *base = c.stage_d.as<unsigned char>();
    uint32_t *hist = (uint32_t *)(base + o_hist), *fill = (uint32_t *)(base + o_fill), *counter = (uint32_t *)(base + o_cnt);
    uint64_t *excl = (uint64_t *)(base + o_excl), *seg_off = (uint64_t *)(base + o_seg);
    GfItem *items = (GfItem *)(base + o_items);
    double2 *rec = c.stage_e.as<double2>();
    PDSB_CUDA(cudaMemsetAsync(base, 0, o_excl, c.stream));              // hist, fill, counter
    {
        LaunchScope ls("grid_tile_hist");
        gf_hist_kernel<<<ceil_div(nvis, 256), 256, 0, c.stream>>>(P, tg, hist, n_outside);
        PDSB_CUDA(cudaGetLastError());
    }
    {
        LaunchScope ls("grid_tile_scan");
        gf_scan_seg_kernel<<<nseg, GF_SCAN_THREADS, 0, c.stream>>>(hist, nkeys, excl, seg_off);
        gf_scan_totals_kernel<<<1, GF_SCAN_THREADS, 0, c.stream>>>(seg_off, nseg);
        gf_items_kernel<<<ceil_div((int64_t)max_items, 256), 256, 0, c.stream>>>(excl, seg_off, nkeys, items);
        PDSB_CUDA(cudaGetLastError());
    }
    {
        LaunchScope ls("grid_tile_records");
        gf_scatter_kernel<<<ceil_div(nvis, 256), 256, 0, c.stream>>>(P, tg, w_src, excl, seg_off, fill, rec);
        PDSB_CUDA(cudaGetLastError());
    }
    {
        LaunchScope ls("grid_tile_accum");
#define PDSB_GF(W, M)                                                                                              \
    launch_tile<W, M>(P, (int)lo, tg, rec, items, excl, seg_off, nkeys, counter, t_re, t_im, t_w)
        if (smode == 0 && width == 3) PDSB_GF(3, 0);
        else if (smode == 0) PDSB_GF(6, 0);
        else if (width == 3) PDSB_GF(3, 1);
        else if (width == 5) PDSB_GF(5, 1);
        else PDSB_GF(7, 1);
#undef PDSB_GF
        PDSB_CUDA(cudaGetLastError());
    }
    return PDSB_OK;
}

}  // namespace pdsb
