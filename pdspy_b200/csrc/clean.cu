// Hogbom CLEAN of a dirty image cube: the loop of pdspy/interferometry/clean.py:51-106 on the device.
//
// The reference works on host numpy arrays: per iteration a masked maximum, a scipy fftconvolve of a
// one-pixel image with the dirty beam (= the beam shifted to the peak, scaled), and two full-array
// medians (astropy.stats.mad_std) for the stopping rule.  Here everything stays in HBM; the host only
// reads back three scalars per iteration to take the reference's decisions in the reference's order.
//
//   masked maximum / plain maximum   two-stage (value, first index) reduction
//   mad_std(x) = 1.482602218505602 * median(|x - median(x)|)
//                                    exact order statistics by MSD radix select on order-preserving
//                                    64-bit keys (8 passes of 8 bits: shared-memory histograms, one pick
//                                    thread), numpy.median's mean of the two middle elements for even n
//   beam subtraction                 closed form of fftconvolve(delta, beam, mode="same"):
//                                    out[y, x] = val * beam[y + ny - 1 - y0, x + nx - 1 - x0]
//   restore (clean.py:110-113)       sum over the (<= maxiter) non-zero model components of the shifted
//                                    clean beam, plus the residuals
// Layout: images [ny, nx, nf] (channel fastest, the reference's image[:, :, :, 0]), beams [2ny, 2nx, nf].
#include "common.cuh"
#include <algorithm>
#include <cmath>

namespace pdsb {

// ---- order-preserving keys ----
__device__ __forceinline__ unsigned long long dkey(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

struct SelectState {
    unsigned long long prefix;     // key bits fixed so far (right aligned)
    unsigned long long rank;       // rank still to descend within the prefix class
    unsigned long long next_min;   // smallest key greater than the selected one
    unsigned int need_next;        // the following order statistic is a different value
    unsigned int hist[256];
};

// mode 0: key of x[i]; mode 1: key of |x[i] - *centre|
__global__ void __launch_bounds__(256) select_hist_kernel(const double *__restrict__ x, int64_t n, int mode,
                                                          const double *__restrict__ centre, int pass,
                                                          SelectState *__restrict__ st)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    const int shift = 56 - 8 * pass;
    const double c = mode ? *centre : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = dkey(mode ? fabs(x[i] - c) : x[i]);
        if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&h[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void select_pick_kernel(SelectState *st, int pass)
{
    if (threadIdx.x != 0) return;
    unsigned long long r = st->rank;
    int b = 0;
    for (; b < 255; b++) {
        const unsigned long long hb = st->hist[b];
        if (r < hb) break;
        r -= hb;
    }
    if (pass == 7) st->need_next = (r + 1 >= (unsigned long long)st->hist[b]) ? 1u : 0u;
    st->prefix = (st->prefix << 8) | (unsigned long long)b;
    st->rank = r;
    for (int i = 0; i < 256; i++) st->hist[i] = 0;
}

__global__ void __launch_bounds__(256) select_next_kernel(const double *__restrict__ x, int64_t n, int mode,
                                                          const double *__restrict__ centre, SelectState *__restrict__ st)
{
    const unsigned long long sel = st->prefix;
    const double c = mode ? *centre : 0.0;
    unsigned long long best = ~0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = dkey(mode ? fabs(x[i] - c) : x[i]);
        if (k > sel && k < best) best = k;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&st->next_min, best);
}

// out = median of the selected statistic(s): the lower middle element, averaged with the next one when n is even
__global__ void select_finish_kernel(const SelectState *st, int even, double *out)
{
    const double a = dunkey(st->prefix);
    if (!even) {
        *out = a;
        return;
    }
    const double b = st->need_next ? dunkey(st->next_min) : a;
    *out = (a + b) / 2.0;
}

__global__ void select_init_kernel(SelectState *st, unsigned long long rank)
{
    if (threadIdx.x == 0) {
        st->prefix = 0;
        st->rank = rank;
        st->next_min = ~0ull;
        st->need_next = 0;
    }
    st->hist[threadIdx.x] = 0;
}

// median of x (mode 0) or of |x - *centre| (mode 1) into *out (device), numpy.median semantics (no NaNs)
static int device_median(const double *x, int64_t n, int mode, const double *centre, SelectState *st, double *out)
{
    Context &c = ctx();
    const int even = (n % 2 == 0);
    const unsigned long long rank = even ? (unsigned long long)(n / 2 - 1) : (unsigned long long)(n / 2);
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c.sm_count * 8);
    LaunchScope ls("clean_median");
    select_init_kernel<<<1, 256, 0, c.stream>>>(st, rank);
    for (int pass = 0; pass < 8; pass++) {
        select_hist_kernel<<<blocks, 256, 0, c.stream>>>(x, n, mode, centre, pass, st);
        select_pick_kernel<<<1, 32, 0, c.stream>>>(st, pass);
    }
    if (even) select_next_kernel<<<blocks, 256, 0, c.stream>>>(x, n, mode, centre, st);
    select_finish_kernel<<<1, 1, 0, c.stream>>>(st, even, out);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

// ---- maxima ----
struct MaxItem {
    double v;
    long long i;
};
__device__ __forceinline__ MaxItem max_item(MaxItem a, MaxItem b)      // larger value, ties -> smaller index
{
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
// masked != null: the maximum of dirty*mask exactly as numpy forms it (0 where the mask is 0)
__global__ void __launch_bounds__(256) max_partial_kernel(const double *__restrict__ x, const double *__restrict__ mask,
                                                          int64_t n, MaxItem *__restrict__ part)
{
    MaxItem m{0.0, -1};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = mask ? x[i] * mask[i] : x[i];
        m = max_item(m, MaxItem{v, (long long)i});
    }
    __shared__ MaxItem sm[256];
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] = max_item(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) max_final_kernel(const MaxItem *__restrict__ part, int nb, MaxItem *out)
{
    MaxItem m{0.0, -1};
    for (int i = threadIdx.x; i < nb; i += 256) m = max_item(m, part[i]);
    __shared__ MaxItem sm[256];
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] = max_item(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}
static int device_max(const double *x, const double *mask, int64_t n, MaxItem *part, MaxItem *out)
{
    Context &c = ctx();
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 1024);
    LaunchScope ls("clean_max");
    max_partial_kernel<<<blocks, 256, 0, c.stream>>>(x, mask, n, part);
    max_final_kernel<<<1, 256, 0, c.stream>>>(part, blocks, out);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

// ---- loop body kernels ----
__global__ void __launch_bounds__(256) clean_init_kernel(const double *__restrict__ dirty, int64_t n,
                                                         unsigned char *__restrict__ wherezero, double *__restrict__ model,
                                                         double *__restrict__ mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    wherezero[i] = dirty[i] == 0.0;
    model[i] = 0.0;
    mask[i] = 0.0;
}
// mask = mask or (dirty > threshold)      (clean.py:60-61, 72-75)
__global__ void __launch_bounds__(256) clean_mask_kernel(const double *__restrict__ dirty, int64_t n, double threshold,
                                                         double *__restrict__ mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (dirty[i] > threshold) mask[i] = 1.0;
}
// model[p] += dirty[p]*gain; dirty[:, :, ch] -= fftconvolve(delta, beam[:, :, ch], "same"); dirty[wherezero] = 0
// (clean.py:83-94).  One thread per pixel of channel ch; the peak value is read before anything is written.
__global__ void __launch_bounds__(256) clean_subtract_kernel(double *__restrict__ dirty, const double *__restrict__ beam,
                                                             const unsigned char *__restrict__ wherezero,
                                                             double *__restrict__ model, const MaxItem *__restrict__ peak,
                                                             double gain, int ny, int nx, int nf, double *__restrict__ sub)
{
    const long long p = peak->i;
    const int ch = (int)(p % nf);
    const int x0 = (int)((p / nf) % nx), y0 = (int)(p / ((long long)nf * nx));
    const double val = *sub;                       // dirty[p]*gain, latched by clean_latch_kernel
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)ny * nx) return;
    const int y = (int)(idx / nx), x = (int)(idx % nx);
    const int64_t o = ((int64_t)y * nx + x) * nf + ch;
    const int by = y + ny - 1 - y0, bx = x + nx - 1 - x0;        // always inside the [2ny, 2nx] beam
    const double d = dirty[o] - val * beam[((int64_t)by * (2 * nx) + bx) * nf + ch];
    dirty[o] = wherezero[o] ? 0.0 : d;
    if (o == p) model[o] += val;
}
__global__ void clean_latch_kernel(const double *__restrict__ dirty, const MaxItem *__restrict__ peak, double gain,
                                   double *__restrict__ sub)
{
    *sub = dirty[peak->i] * gain;
}

// ---- restore ----
__global__ void __launch_bounds__(256) clean_compact_kernel(const double *__restrict__ model, int64_t n,
                                                            long long *__restrict__ comp_idx, unsigned int *__restrict__ ncomp,
                                                            unsigned int cap)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || model[i] == 0.0) return;
    const unsigned int k = atomicAdd(ncomp, 1u);
    if (k < cap) comp_idx[k] = i;
}
// clean_image = fftconvolve(model, clean_beam, "same") + residuals, components taken in index order
__global__ void __launch_bounds__(256) clean_restore_kernel(const double *__restrict__ model, const double *__restrict__ cbeam,
                                                            const double *__restrict__ resid,
                                                            const long long *__restrict__ comp_idx, unsigned int ncomp,
                                                            int ny, int nx, int nf, double *__restrict__ out)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (int64_t)ny * nx * nf) return;
    const int ch = (int)(o % nf);
    const int x = (int)((o / nf) % nx), y = (int)(o / ((int64_t)nf * nx));
    double s = 0.0;
    for (unsigned int k = 0; k < ncomp; k++) {
        const long long p = comp_idx[k];
        if ((int)(p % nf) != ch) continue;
        const int x0 = (int)((p / nf) % nx), y0 = (int)(p / ((long long)nf * nx));
        s += model[p] * cbeam[((int64_t)(y + ny - 1 - y0) * (2 * nx) + (x + nx - 1 - x0)) * nf + ch];
    }
    out[o] = s + resid[o];
}

struct CleanScalars {          // device scalars read back by the host loop
    MaxItem masked, plain;
    double med, mad, sub;
};

}  // namespace pdsb

using namespace pdsb;

static const double kMadScale = 1.482602218505602;      // astropy.stats.mad_std

int pdsb_mad_std(const double *x, int64_t n, int kind, double *out)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(n > 0 && x && out, "array");
    const void *dx = nullptr;
    PDSB_CHECK(to_device(x, kind, (size_t)n * sizeof(double), c.stage_a, &dx));
    PDSB_CHECK(c.small_dev.ensure(sizeof(SelectState) + sizeof(CleanScalars) + 64));
    SelectState *st = c.small_dev.as<SelectState>();
    CleanScalars *sc = reinterpret_cast<CleanScalars *>(st + 1);
    PDSB_CHECK(device_median(static_cast<const double *>(dx), n, 0, nullptr, st, &sc->med));
    PDSB_CHECK(device_median(static_cast<const double *>(dx), n, 1, &sc->med, st, &sc->mad));
    double mad = 0;
    PDSB_CUDA(cudaMemcpyAsync(&mad, &sc->mad, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    *out = mad * kMadScale;
    return PDSB_OK;
}

int pdsb_clean_loop(double *dirty, const double *dirty_beam, int ny, int nx, int nf, double beam_resid_max, double gain,
                    int maxiter, double nsigma, int kind, double *model, double *mask, int *niter, double *threshold_out)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(ny > 0 && nx > 0 && nf > 0 && maxiter >= 0, "sizes");
    PDSB_REQUIRE(dirty && dirty_beam && model && mask, "arrays");
    const int64_t n = (int64_t)ny * nx * nf, nb = 4 * n;
    double *d_dirty = dirty, *d_model = model, *d_mask = mask;
    const double *d_beam = dirty_beam;
    if (kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_b.ensure((size_t)(3 * n + nb) * sizeof(double)));
        d_dirty = c.stage_b.as<double>();
        d_model = d_dirty + n;
        d_mask = d_model + n;
        double *b = d_mask + n;
        PDSB_CUDA(cudaMemcpyAsync(d_dirty, dirty, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(b, dirty_beam, (size_t)nb * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        d_beam = b;
    }
    PDSB_CHECK(c.stage_c.ensure((size_t)n + 1024 * sizeof(MaxItem) + 64));
    unsigned char *wherezero = c.stage_c.as<unsigned char>();
    MaxItem *part = reinterpret_cast<MaxItem *>(wherezero + ((n + 63) / 64) * 64);
    PDSB_CHECK(c.small_dev.ensure(sizeof(SelectState) + sizeof(CleanScalars) + 64));
    SelectState *st = c.small_dev.as<SelectState>();
    CleanScalars *sc = reinterpret_cast<CleanScalars *>(st + 1);
    CleanScalars h;

    auto fetch = [&]() -> int {
        PDSB_CUDA(cudaMemcpyAsync(&h, sc, sizeof(CleanScalars), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
        return PDSB_OK;
    };
    auto mad_std = [&]() -> int {             // leaves sc->mad
        PDSB_CHECK(device_median(d_dirty, n, 0, nullptr, st, &sc->med));
        return device_median(d_dirty, n, 1, &sc->med, st, &sc->mad);
    };
    // threshold = max((dirty_beam - clean_beam).max() * dirty.max(), 5*mad_std(dirty)); mask |= dirty > threshold
    double threshold = 0;
    auto update_mask = [&]() -> int {
        PDSB_CHECK(device_max(d_dirty, nullptr, n, part, &sc->plain));
        PDSB_CHECK(mad_std());
        PDSB_CHECK(fetch());
        threshold = std::fmax(beam_resid_max * h.plain.v, 5.0 * (h.mad * kMadScale));
        LaunchScope ls("clean_mask");
        clean_mask_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(d_dirty, n, threshold, d_mask);
        PDSB_CUDA(cudaGetLastError());
        return PDSB_OK;
    };

    {
        LaunchScope ls("clean_init");
        clean_init_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(d_dirty, n, wherezero, d_model, d_mask);
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CHECK(update_mask());                                                   // clean.py:57-61
    int it = 0;
    bool stop = false;
    while (it < maxiter && !stop) {
        PDSB_CHECK(device_max(d_dirty, d_mask, n, part, &sc->masked));
        PDSB_CHECK(fetch());
        if (h.masked.v < threshold) {                                            // clean.py:70-76
            PDSB_CHECK(update_mask());
            PDSB_CHECK(device_max(d_dirty, d_mask, n, part, &sc->masked));
            PDSB_CHECK(fetch());
        }
        if (!(h.masked.v > 0.0)) break;       // nothing positive under the mask: the reference's where() would
                                              // select every unmasked pixel here; stop instead
        {
            LaunchScope ls("clean_subtract");
            clean_latch_kernel<<<1, 1, 0, c.stream>>>(d_dirty, &sc->masked, gain, &sc->sub);
            clean_subtract_kernel<<<ceil_div((int64_t)ny * nx, 256), 256, 0, c.stream>>>(d_dirty, d_beam, wherezero, d_model,
                                                                                      &sc->masked, gain, ny, nx, nf, &sc->sub);
            PDSB_CUDA(cudaGetLastError());
        }
        PDSB_CHECK(device_max(d_dirty, d_mask, n, part, &sc->masked));           // clean.py:98
        PDSB_CHECK(mad_std());
        PDSB_CHECK(fetch());
        stop = h.masked.v < nsigma * (h.mad * kMadScale);
        it++;
    }
    if (kind == PDSB_HOST) {
        PDSB_CUDA(cudaMemcpyAsync(dirty, d_dirty, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(model, d_model, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(mask, d_mask, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    }
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    if (niter) *niter = it;
    if (threshold_out) *threshold_out = threshold;
    return PDSB_OK;
}

int pdsb_clean_restore(const double *model, const double *clean_beam, const double *residuals, int ny, int nx, int nf,
                       int kind, double *clean_image)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(ny > 0 && nx > 0 && nf > 0, "sizes");
    PDSB_REQUIRE(model && clean_beam && residuals && clean_image, "arrays");
    const int64_t n = (int64_t)ny * nx * nf, nb = 4 * n;
    const double *d_model = model, *d_cb = clean_beam, *d_res = residuals;
    double *d_out = clean_image;
    if (kind == PDSB_HOST) {
        PDSB_CHECK(c.stage_b.ensure((size_t)(3 * n + nb) * sizeof(double)));
        double *p = c.stage_b.as<double>();
        PDSB_CUDA(cudaMemcpyAsync(p, model, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(p + n, residuals, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        PDSB_CUDA(cudaMemcpyAsync(p + 3 * n, clean_beam, (size_t)nb * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        d_model = p;
        d_res = p + n;
        d_out = p + 2 * n;
        d_cb = p + 3 * n;
    }
    // non-zero model components, in index order (the summation order of the restore)
    PDSB_CHECK(c.stage_c.ensure((size_t)n * sizeof(long long) + 64));
    long long *comp = c.stage_c.as<long long>();
    PDSB_CHECK(c.small_dev.ensure(64));
    unsigned int *ncomp_d = c.small_dev.as<unsigned int>();
    PDSB_CUDA(cudaMemsetAsync(ncomp_d, 0, sizeof(unsigned int), c.stream));
    unsigned int ncomp = 0;
    {
        LaunchScope ls("clean_compact");
        clean_compact_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(d_model, n, comp, ncomp_d, (unsigned int)n);
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CUDA(cudaMemcpyAsync(&ncomp, ncomp_d, sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    if (ncomp > 1) {          // atomics give an arbitrary order: sort the (few) indices on the host
        std::vector<long long> hc(ncomp);
        PDSB_CUDA(cudaMemcpy(hc.data(), comp, (size_t)ncomp * sizeof(long long), cudaMemcpyDeviceToHost));
        std::sort(hc.begin(), hc.end());
        PDSB_CUDA(cudaMemcpyAsync(comp, hc.data(), (size_t)ncomp * sizeof(long long), cudaMemcpyHostToDevice, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    {
        LaunchScope ls("clean_restore");
        clean_restore_kernel<<<ceil_div(n, 256), 256, 0, c.stream>>>(d_model, d_cb, d_res, comp, ncomp, ny, nx, nf, d_out);
        PDSB_CUDA(cudaGetLastError());
    }
    if (kind == PDSB_HOST)
        PDSB_CUDA(cudaMemcpyAsync(clean_image, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    return PDSB_OK;
}
