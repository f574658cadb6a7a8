// Fast (non-bit-exact) mode of grid(): the scatter loop of pdspy/interferometry/libinterferometry.pyx:489-521
// as  counting sort by 8x8-cell uv tile  ->  contiguous 48-byte records  ->  register accumulation per tile
// ->  shared-memory region  ->  one fp64 atomic per region cell and map.
//
// Pipeline (every kernel streams its inputs once; no gathers, no multi-pass radix sort):
//   gf_hist_kernel     u, v -> sort key of every visibility = (home tile, home row inside the tile); histogram over
//                      the keys (shared-memory window for the hot central tiles, else warp-aggregated atomics).  The
//                      atomics' return values give every visibility its rank inside its key's run (rank[idx]): the
//                      only atomic pass of the pipeline.
//   gf_scan_*          exclusive scan of (count per key, work items per tile), packed in one 64-bit word
//   gf_items_kernel    work list: every tile's run cut into items of <= GF_ITEM visibilities
//   gf_scatter_kernel  second pass over the inputs, pure streaming: 48-byte record [pos_u, w][w re, w im][pos_v, (gi | gj << 32)]
//                      written at prefix(key) + rank (the L2 merges the write streams, so DRAM sees full lines).
//                      A tile's run is therefore already in home-row order.
//   gf_tile_kernel     persistent warps, one work item each; the item's records are streamed in order (coalesced,
//                      next chunk prefetched into registers while the current one is worked on):
//        * the 8 row buckets of the item are read off the scan (no sorting in the kernel);
//        * factor phase: the 2 x WIDTH one-dimensional kernel factors (the exp*sinc kernel is separable), staged in
//          shared memory;
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

#ifndef GF_ITEM_N
#define GF_ITEM_N 512
#endif
constexpr int GF_ITEM = GF_ITEM_N;  // visibilities per work item (one warp)
constexpr int GF_WARPS = 4;         // warps per CTA, each on its own item
#ifndef GF_MINB
#define GF_MINB 4
#endif
constexpr int GF_SCAN_THREADS = 1024;
constexpr int GF_SCAN_PER = 8;          // = the 8 row keys of one tile: a scan thread owns a whole tile
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
    // nuv * nf < 2^32 (checked by pdsb_grid): 32-bit division, an order of magnitude cheaper than the 64-bit routine
    const uint32_t k = P.nf == 1 ? (uint32_t)idx : (uint32_t)idx / (uint32_t)P.nf;
    const double f = P.freq[P.nf == 1 ? 0u : (uint32_t)idx - k * (uint32_t)P.nf];
    x.pu = __dmul_rn(__dmul_rn(P.u[k], f), P.inv_freq);
    x.pv = __dmul_rn(__dmul_rn(P.v[k], f), P.inv_freq);
    x.gi = np_f64_to_u32(__dadd_rn(__ddiv_rn(x.pu, P.binsize), P.half));        // :388-403
    x.gj = np_f64_to_u32(__dadd_rn(__ddiv_rn(x.pv, P.binsize), P.half));
    x.good = x.gi < (uint32_t)P.G && x.gj < (uint32_t)P.G;                        // :421-423
    return x;
}

// sort key = (channel, home tile, home row inside the tile): 8 keys per tile, so that a tile's run arrives in the
// row-bucket order the tile kernel accumulates in
__device__ __forceinline__ uint32_t gf_key(const GridParams &P, const GfVis &x, int64_t idx, uint32_t tg)
{
    const uint32_t tile = (x.gj >> 3) * tg + (x.gi >> 3);
    return (((P.spectral ? (uint32_t)idx % (uint32_t)P.nf : 0u) * tg * tg + tile) << 3) | (x.gj & 7u);
}

// Interferometric uv coverage peaks at the short baselines: the few tiles around the grid centre receive a large
// share of all visibilities (26 % of BASELINE configs[3] land in the four central tiles), and the L2 serialises
// atomics per address.  Both passes therefore count the central GF_WIN x GF_WIN tiles in a shared-memory window per
// block first (one global atomic per block and window tile); the rest of the plane goes to global memory directly.
// The window is used when the key has no channel part (nch == 1); otherwise equal keys are aggregated per warp.
constexpr int GF_WIN = 16;
constexpr int GF_WIN_KEYS = GF_WIN * GF_WIN * 8;
constexpr int GF_PASS_ITEMS = 4;                   // visibilities per thread in the two passes
constexpr int GF_PASS_TILE = 256 * GF_PASS_ITEMS;

__device__ __forceinline__ int gf_window_index(uint32_t gi, uint32_t gj, uint32_t tg)
{
    const int org = (int)(tg / 2) - GF_WIN / 2 > 0 ? (int)(tg / 2) - GF_WIN / 2 : 0;
    const int wr = (int)(gj >> 3) - org, wc = (int)(gi >> 3) - org;
    return (wr >= 0 && wr < GF_WIN && wc >= 0 && wc < GF_WIN) ? ((wr * GF_WIN + wc) << 3) | (int)(gj & 7u) : -1;
}
__device__ __forceinline__ uint32_t gf_window_key(int widx, uint32_t tg)
{
    const int org = (int)(tg / 2) - GF_WIN / 2 > 0 ? (int)(tg / 2) - GF_WIN / 2 : 0;
    const int wt = widx >> 3;
    return (((uint32_t)(org + wt / GF_WIN) * tg + (uint32_t)(org + wt % GF_WIN)) << 3) | (uint32_t)(widx & 7);
}

// ---- pass 1: histogram over the keys, and every visibility's rank inside its key's run ------------------------
// The atomics that build the histogram return the old count: that IS the visibility's slot among the visibilities of
// its key (window keys: the block's base from the global atomic + the slot inside the block from the shared one).
// Stored as rank[idx] (0xffffffff = off the grid), it lets the second pass run without any atomic or barrier.
__global__ void __launch_bounds__(256) gf_hist_kernel(GridParams P, uint32_t tg, uint32_t *__restrict__ hist,
                                                      uint32_t *__restrict__ rank, unsigned long long *n_outside)
{
    __shared__ uint32_t sh[GF_WIN_KEYS];
    const bool use_win = P.nch == 1;
    const int lane = threadIdx.x & 31;
    if (use_win) {
        for (int i = threadIdx.x; i < GF_WIN_KEYS; i += 256) sh[i] = 0;
        __syncthreads();
    }
    uint32_t n_out = 0;
    uint32_t slot[GF_PASS_ITEMS];
    int widx[GF_PASS_ITEMS];
#pragma unroll
    for (int it = 0; it < GF_PASS_ITEMS; it++) {
        const int64_t idx = (int64_t)blockIdx.x * GF_PASS_TILE + it * 256 + threadIdx.x;
        uint32_t key = GF_DEAD;
        widx[it] = -1;
        slot[it] = GF_DEAD;
        if (idx < P.nuv * P.nf) {
            const GfVis x = gf_vis(P, idx);
            if (x.good) {
                key = gf_key(P, x, idx, tg);
                if (use_win) widx[it] = gf_window_index(x.gi, x.gj, tg);
            } else
                n_out++;
        }
        if (use_win) {
            if (widx[it] >= 0) slot[it] = atomicAdd(&sh[widx[it]], 1u);
            else if (key != GF_DEAD) slot[it] = atomicAdd(hist + key, 1u);
        } else {
            const uint32_t peers = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (key != GF_DEAD && leader == lane) base = atomicAdd(hist + key, (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (key != GF_DEAD) slot[it] = base + __popc(peers & ((1u << lane) - 1u));
        }
    }
    if (n_outside) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_out += __shfl_xor_sync(0xffffffffu, n_out, o);
        if (lane == 0 && n_out) atomicAdd(n_outside, (unsigned long long)n_out);
    }
    if (use_win) {
        // the block's base slot in every window key: all atomics go out before the first result is needed (issue
        // is in order: a store of the result right behind its atomic would serialise the round trips)
        __syncthreads();
        constexpr int NQ = GF_WIN_KEYS / 256;
        uint32_t cnt[NQ], base[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) cnt[q] = sh[q * 256 + threadIdx.x];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            base[q] = 0;
            if (cnt[q]) base[q] = atomicAdd(hist + gf_window_key(q * 256 + threadIdx.x, tg), cnt[q]);
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) sh[q * 256 + threadIdx.x] = base[q];
        __syncthreads();
    }
#pragma unroll
    for (int it = 0; it < GF_PASS_ITEMS; it++) {
        const int64_t idx = (int64_t)blockIdx.x * GF_PASS_TILE + it * 256 + threadIdx.x;
        if (idx < P.nuv * P.nf) rank[idx] = slot[it] + (widx[it] >= 0 ? sh[widx[it]] : 0u);
    }
}

// ---- exclusive scan of packed (count | items << 40) ---------------------------------------------------------
// counts are per key (8 row keys per tile); work items are per TILE (its 8 runs together, cut into pieces of
// GF_ITEM) and are attached to the tile's last key, so that prefix(8 T) >> 40 = items before tile T.
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
        v[q] = (e0 + q < nkeys) ? (uint64_t)hist[e0 + q] : 0ull;
        tsum += v[q];
    }
    v[GF_SCAN_PER - 1] += ((tsum + GF_ITEM - 1) / GF_ITEM) << 40;
    tsum += ((tsum + GF_ITEM - 1) / GF_ITEM) << 40;
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
    // the tile whose item range holds t: largest tile T with items_before(T) <= t  (empty tiles repeat the value of
    // their successor, so the search lands on the non-empty one)
    uint32_t lo = 0, hi = nkeys >> 3;                  // items_before(lo) <= t < items_before(hi)
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((uint32_t)(gf_prefix(excl, seg_off, mid << 3) >> 40) <= t) lo = mid;
        else hi = mid;
    }
    const uint64_t p0 = gf_prefix(excl, seg_off, lo << 3), p1 = gf_prefix(excl, seg_off, (lo + 1) << 3);
    const uint32_t j = t - (uint32_t)(p0 >> 40);
    const uint32_t start = (uint32_t)(p0 & GF_LOW40), stop = (uint32_t)(p1 & GF_LOW40);
    GfItem it;
    it.begin = start + j * GF_ITEM;
    it.end = it.begin + GF_ITEM < stop ? it.begin + GF_ITEM : stop;
    it.key = lo;                                       // tile id (channel * tg^2 + tile)
    it.pad = 0;
    items[t] = it;
}

// ---- pass 2: records to their key runs (streaming: no atomics, no barriers) -----------------------------------
__global__ void __launch_bounds__(256) gf_scatter_kernel(GridParams P, uint32_t tg, const double *__restrict__ w_src,
                                                         const uint64_t *__restrict__ excl,
                                                         const uint64_t *__restrict__ seg_off,
                                                         const uint32_t *__restrict__ rank, double2 *__restrict__ rec)
{
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= P.nuv * P.nf) return;
    const uint32_t rk = rank[idx];
    const double re = P.re[idx], im = P.im[idx];
    double w = w_src ? w_src[idx] : P.w_in[idx];
    if (rk == GF_DEAD) return;
    const GfVis x = gf_vis(P, idx);
    const uint32_t key = gf_key(P, x, idx, tg);
    const uint32_t pos = (uint32_t)(gf_prefix(excl, seg_off, key) & GF_LOW40) + rk;
    if (!w_src) {                                           // :351-353
        w = w < 0 ? 0.0 : w;
        if (re == 0 && im == 0) w = 0.0;
    }
    double2 *r = rec + 3 * (size_t)pos;
    r[0] = make_double2(x.pu, w);
    r[1] = make_double2(re * w, im * w);
    r[2] = make_double2(x.pv, __longlong_as_double((long long)((unsigned long long)x.gi |
                                                               ((unsigned long long)x.gj << 32))));
}

// ---- convolution kernel factors ------------------------------------------------------------------------------
// exp_sinc is separable: exp_sinc(u, v) = g(u) g(v) / norm, g(x) = sinc(x/1.55) exp(-(x/2.52)^2) inside |x| < 3,
// with the reference's degree-16 Taylor polynomial for sinc and degree-5 polynomial for exp (:547-574).  The powers
// and the sum are formed in the reference's order (like k_sinc / k_exp of grid.cuh), NOT by Horner's rule: where the
// sinc polynomial crosses zero (x = 1.55) the factor is pure rounding residue, and a weighted mean over a cell fed by
// such factors moves by 4e-11 when the residue changes - the same operation order keeps the residues correlated
// with the reference's (measured: 1e-12 vs 4e-11 with Horner).  Coefficients live in the constant bank: as literals
// every DFMA drags two UMOVs along to build its 64-bit operand.
__constant__ double gf_poly[16] = {1. / 6., 1. / 120., 1. / 5040., 1. / 362880., 1. / 39916800., 1. / 6227020800.,
                                   1. / 1307674368000., 1. / 355687428096000.,
                                   0.5, 1. / 6., 1. / 24., 1. / 120.,
                                   1. / 1.55, 3.14159265358979323846, 1. / 2.52, 0.};
__device__ __forceinline__ double gf_exp_sinc_1d(double x)
{
    if (fabs(x) >= 3.0) return 0.;
    const double xp = x * gf_poly[12] * gf_poly[13];
    const double x2 = xp * xp, x4 = x2 * x2, x6 = x4 * x2, x8 = x4 * x4, x10 = x8 * x2, x12 = x8 * x4, x14 = x8 * x6,
                 x16 = x8 * x8;
    const double s = 1. - x2 * gf_poly[0] + x4 * gf_poly[1] - x6 * gf_poly[2] + x8 * gf_poly[3] - x10 * gf_poly[4] +
                     x12 * gf_poly[5] - x14 * gf_poly[6] + x16 * gf_poly[7];
    const double a = x * gf_poly[14];
    const double z = -1 * (a * a);
    const double z2 = z * z, z3 = z2 * z, z4 = z2 * z2, z5 = z4 * z;
    const double e = 1 + z + z2 * gf_poly[8] + z3 * gf_poly[9] + z4 * gf_poly[10] + z5 * gf_poly[11];
    return s * e;
}

// ---- tile kernel ------------------------------------------------------------------------------------------------
// WIDTH = footprint width lo + hi + 1 (pillbox 3, exp*sinc 6, box sums 1 / 3 / 5 / 7); MODE 0 = main sums (three maps),
// MODE 1 = box sums of the weights (one map, all factors 1).
//
// One warp per work item (<= GF_ITEM consecutive records of one tile, in home-row order).  Per chunk of 32 records:
//   A  lane = visibility: the 48-byte record (prefetched into registers one chunk ahead) into shared memory;
//   B  lane = (visibility, axis): the WIDTH kernel factors of that axis, WIDTH independent polynomial evaluations per lane;
//   C  half-warp = visibility, lane = region column: WIDTH x NMAP DFMAs per lane into registers, all of them live
//      because the bucket (home row) fixes the rows.  When the bucket changes the two half-warps add their registers
//      to the warp's private region in shared memory (plain read-modify-write, one half after the other).
template <int WIDTH, int MODE>
__global__ void __launch_bounds__(GF_WARPS * 32, GF_MINB) gf_tile_kernel(GridParams P, int lo, uint32_t tg,
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
    constexpr int SBS = 18;                             // doubles per staged visibility (16-byte aligned, spreads banks)
    constexpr int RC = SIDE + (SIDE & 1);               // region columns held (even: rows stay 16-byte aligned)
    __shared__ double s_reg[GF_WARPS][NMAP][SIDE][RC];  // the warp's private region: [map][row][column]
    __shared__ double s_fu[GF_WARPS][WIDTH][FUS];       // column factors: [slot][visibility]
    // per staged visibility: [0..6] row factors, [8] w, [9] w re, [10] w im, [11] meta, [12] pos_u, [13] pos_v
    __shared__ __align__(16) double s_b[GF_WARPS][32][SBS];
    __shared__ double s_cen[GF_WARPS][2][16];           // cell centres of the region's columns / rows
    __shared__ int s_cnt[GF_WARPS][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, c = lane & 15;
    const uint32_t nitems = (uint32_t)(gf_prefix(excl, seg_off, nkeys) >> 40);
    double(*reg)[SIDE][RC] = s_reg[warp];
    double(*fu)[FUS] = s_fu[warp];
    double(*sb)[SBS] = s_b[warp];
    int *bcnt = s_cnt[warp];
    const double *fu_lane = &fu[0][0] + c * FUS;

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
        // first chunk of records: lane = visibility, 32 consecutive 48-byte records = one contiguous run
        double2 c0 = make_double2(0., 0.), c1 = c0, c2 = c0;
        if (lane < n) {
            c0 = irec[3 * lane];
            c1 = irec[3 * lane + 1];
            c2 = irec[3 * lane + 2];
        }
        // zero the region; bucket starts (bcnt[b] = first record of home row b inside the item, bcnt[8] = n) read
        // off the scan; centres of the region's cells
        {
            double *z = &reg[0][0][0];
            for (int i = lane; i < NMAP * SIDE * RC; i += 32) z[i] = 0.0;
            if (lane < 9) {
                const long long s = (long long)(gf_prefix(excl, seg_off, (item.key << 3) + lane) & GF_LOW40) -
                                    (long long)item.begin;
                bcnt[lane] = s < 0 ? 0 : s > n ? n : (int)s;
            }
            const int cell = (half ? tl : tm) * 8 - lo + c;
            s_cen[warp][half][c] = (cell >= 0 && cell < P.G) ? (half ? P.vv[cell] : P.uu[cell]) : 0.0;
        }
        __syncwarp();

        double acc[NMAP][WIDTH];
#pragma unroll
        for (int m = 0; m < NMAP; m++)
#pragma unroll
            for (int r = 0; r < WIDTH; r++) acc[m][r] = 0.0;
        int cb = -1;                                     // bucket the accumulators belong to (warp-uniform)
        auto flush = [&]() {
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                if (half == hh && c < RC) {
#pragma unroll
                    for (int m = 0; m < NMAP; m++)
#pragma unroll
                        for (int r = 0; r < WIDTH; r++) {
                            reg[m][cb + r][c] += acc[m][r];
                            acc[m][r] = 0.0;
                        }
                }
                __syncwarp();
            }
        };

        for (int p0 = 0; p0 < n; p0 += 32) {
            const int nst = n - p0 < 32 ? n - p0 : 32;
            // ---- A: lane = visibility ----
            if (lane < nst) {
                const unsigned long long ij = (unsigned long long)__double_as_longlong(c2.y);
                const int di = (int)(uint32_t)ij - tm * 8, dj = (int)(uint32_t)(ij >> 32) - tl * 8;
                // meta: low word = offset of this visibility's column factor for region column 0 (lane c adds c * FUS),
                // high word = home column | home row << 4 (inside the tile)
                const long long meta = (long long)(uint32_t)(lane - FUS * di) | ((long long)(di | (dj << 4)) << 32);
                double2 *d = reinterpret_cast<double2 *>(sb[lane]);
                d[4] = make_double2(c0.y, c1.x);
                d[5] = make_double2(c1.y, __longlong_as_double(meta));
                d[6] = make_double2(c0.x, c2.x);
            }
            if (p0 + 32 + lane < n) {                    // next chunk: in flight during B and C
                const double2 *r = irec + 3 * (size_t)(p0 + 32 + lane);
                c0 = r[0];
                c1 = r[1];
                c2 = r[2];
            }
            __syncwarp();
            // ---- B: lane = (visibility, axis) ----
            for (int v = lane >> 1; v < nst; v += 16) {
                const int axis = lane & 1;
                const int dd = (int)((unsigned long long)__double_as_longlong(sb[v][11]) >> 32);
                const int home = axis ? (dd >> 4) : (dd & 15);
                const int cell0 = (axis ? tl : tm) * 8 - lo + home;
                const double pos = sb[v][12 + axis];
                const double *cen = s_cen[warp][axis] + home;
                double g[WIDTH];
#pragma unroll
                for (int o = 0; o < WIDTH; o++) {
                    g[o] = 0.0;
                    if (cell0 + o >= 0 && cell0 + o < P.G) {
                        if (MODE == 0) {
                            const double d = (pos - cen[o]) * P.inv_binsize;
                            g[o] = P.conv ? gf_exp_sinc_1d(d) : (fabs(d) >= 0.5 ? 0.0 : 1.0);
                        } else
                            g[o] = 1.0;
                    }
                }
                if (axis) {
#pragma unroll
                    for (int o = 0; o < WIDTH; o++) sb[v][o] = g[o];
                } else {
                    const double norm = (MODE == 0 && P.conv) ? (1. / 2.350016262343186) : 1.0;   // 1/norm with the u factor
#pragma unroll
                    for (int o = 0; o < WIDTH; o++) fu[o][v] = g[o] * norm;
                }
            }
            __syncwarp();
            // ---- C: half-warp = visibility, lane = region column; bucket by bucket ----
            for (int b = 0; b < 8; b++) {
                const int s0 = (bcnt[b] > p0 ? bcnt[b] : p0) - p0;
                const int e0 = (bcnt[b + 1] < p0 + nst ? bcnt[b + 1] : p0 + nst) - p0;
                if (s0 >= e0) continue;
                if (b != cb) {
                    if (cb >= 0) flush();
                    cb = b;
                }
                for (int v = s0 + half; v < e0; v += 2) {
                    const double *vb = sb[v];
                    const double2 wx = *reinterpret_cast<const double2 *>(vb + 8);       // w, w re
                    const double2 ym = *reinterpret_cast<const double2 *>(vb + 10);      // w im, meta
                    const long long meta = __double_as_longlong(ym.y);
                    const int off = (int)meta, di = (int)(meta >> 32) & 15;
                    double fv[8];
#pragma unroll
                    for (int r = 0; r < (WIDTH + 1) / 2; r++)
                        *reinterpret_cast<double2 *>(&fv[2 * r]) = *reinterpret_cast<const double2 *>(vb + 2 * r);
                    const double f = (unsigned)(c - di) < (unsigned)WIDTH ? fu_lane[off] : 0.0;
#pragma unroll
                    for (int m = 0; m < NMAP; m++) {
                        const double x = f * (m == 0 ? wx.x : m == 1 ? wx.y : ym.x);
#pragma unroll
                        for (int r = 0; r < WIDTH; r++) acc[m][r] = fma(fv[r], x, acc[m][r]);
                    }
                }
            }
            __syncwarp();
        }
        if (cb >= 0) flush();
        // the region to the map
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
    PDSB_REQUIRE((smode == 0 && (width == 3 || width == 6)) || (smode == 1 && (width == 1 || width == 3 || width == 5 || width == 7)),
                 "footprint width of the fast gridding mode");
    const uint32_t tg = ((uint32_t)P.G + 7u) >> 3;
    const uint64_t nkeys64 = (uint64_t)tg * tg * (uint64_t)P.nch * 8ull;          // 8 row keys per tile
    PDSB_REQUIRE(nkeys64 < (1ull << 31), "too many uv tiles");
    const uint32_t nkeys = (uint32_t)nkeys64;
    const int nseg = ceil_div((int64_t)nkeys + 1, GF_SCAN_SEG);
    const uint64_t max_items = std::min<uint64_t>(nkeys64 >> 3, (uint64_t)nvis) + (uint64_t)nvis / GF_ITEM + 2;
    PDSB_REQUIRE(max_items < (1ull << 31), "too many gridding work items");

    // scratch: stage_d = hist | rank | counter | excl | seg_off | items ; stage_e = records
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t o_hist = 0, o_cnt = o_hist + up((size_t)nkeys * 4), o_rank = o_cnt + 256,
                 o_excl = o_rank + up((size_t)nvis * 4), o_seg = o_excl + up(((size_t)nseg * GF_SCAN_SEG + 1) * 8),
                 o_items = o_seg + up((size_t)nseg * 8), total = o_items + up((size_t)max_items * sizeof(GfItem));
    PDSB_CHECK(c.stage_d.ensure(total));
    PDSB_CHECK(c.stage_e.ensure((size_t)nvis * 3 * sizeof(double2)));
    unsigned char *base = c.stage_d.as<unsigned char>();
    uint32_t *hist = (uint32_t *)(base + o_hist), *rank = (uint32_t *)(base + o_rank), *counter = (uint32_t *)(base + o_cnt);
    uint64_t *excl = (uint64_t *)(base + o_excl), *seg_off = (uint64_t *)(base + o_seg);
    GfItem *items = (GfItem *)(base + o_items);
    double2 *rec = c.stage_e.as<double2>();
    PDSB_CUDA(cudaMemsetAsync(base, 0, o_rank, c.stream));              // hist, counter
    {
        LaunchScope ls("grid_tile_hist");
        gf_hist_kernel<<<ceil_div(nvis, GF_PASS_TILE), 256, 0, c.stream>>>(P, tg, hist, rank, n_outside);
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
        gf_scatter_kernel<<<ceil_div(nvis, 256), 256, 0, c.stream>>>(P, tg, w_src, excl, seg_off, rank, rec);
        PDSB_CUDA(cudaGetLastError());
    }
    {
        LaunchScope ls("grid_tile_accum");
#define PDSB_GF(W, M)                                                                                              \
    launch_tile<W, M>(P, (int)lo, tg, rec, items, excl, seg_off, nkeys, counter, t_re, t_im, t_w)
        if (smode == 0 && width == 3) PDSB_GF(3, 0);
        else if (smode == 0) PDSB_GF(6, 0);
        else if (width == 1) PDSB_GF(1, 1);
        else if (width == 3) PDSB_GF(3, 1);
        else if (width == 5) PDSB_GF(5, 1);
        else PDSB_GF(7, 1);
#undef PDSB_GF
        PDSB_CUDA(cudaGetLastError());
    }
    return PDSB_OK;
}

}  // namespace pdsb
