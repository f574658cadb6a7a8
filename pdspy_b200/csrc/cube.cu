// Channel post-processing of a model cube before it is Fourier-sampled
// (pdspy/modeling/run_flared_model.py:308-366): RADMC-3D renders `subsample` sub-channels per output
// channel and, when the data were averaged on the sky, `averaging` times more; the reference then
//   (1) means each block of `subsample` sub-channels                       (:316-320 / :345-349)
//   (2) optionally Hanning-smooths along the channel axis with numpy.hanning(5)/sum = [0, 1/4, 1/2, 1/4, 0],
//       zero-padded at the ends (scipy.signal.fftconvolve(..., mode="same"))  (:324-329 / :353-358)
//   (3) means each block of `averaging` channels                            (:333-339 / :362-366)
// Layout [pixel][channel] (the reference's [ny, nx, nf, 1] regular cube and [npts, nf] unstructured image
// are both that), fp64.  HBM-bound: nf_in reads + nf_out writes of 8 B per pixel; one thread per output
// element, channel fastest, so a warp reads one contiguous run of sub-channels.
#include "common.cuh"

namespace pdsb {

// out[p][c] = in[p][c0 + c]: the compact cube of a channel shard
__global__ void __launch_bounds__(256) channel_slice_kernel(const double *__restrict__ in, double *__restrict__ out,
                                                            int64_t npix, int nf_in, int c0, int nf_out)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= npix * nf_out) return;
    const int64_t p = t / nf_out;
    const int ch = (int)(t - p * nf_out);
    out[t] = in[p * nf_in + c0 + ch];
}

// in_scale (device, [nf_in], may be null): per-INPUT-channel factor applied before step (1) - the reference's
// extinction, image[:,:,i,:] *= extinction[i] (:286-299), which precedes the sub-sample mean and the smoothing
// and does not commute with them.
__global__ void __launch_bounds__(256) channel_post_kernel(const double *__restrict__ in, double *__restrict__ out,
                                                           int64_t npix, int nf_in, int subsample, int hanning,
                                                           int averaging, int nf_mid, int nf_out,
                                                           const double *__restrict__ in_scale)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * nf_out) return;
    const int k = (int)(idx % nf_out);
    const double *src = in + (idx / nf_out) * nf_in;
    auto sub_mean = [&](int m) -> double {                  // step (1): sums in index order, then one division
        double s = 0.0;
        if (in_scale)
            for (int j = 0; j < subsample; j++) s += __dmul_rn(src[m * subsample + j], in_scale[m * subsample + j]);
        else
            for (int j = 0; j < subsample; j++) s += src[m * subsample + j];
        return s / (double)subsample;
    };
    double acc = 0.0;
    for (int a = 0; a < averaging; a++) {
        const int m = k * averaging + a;
        double r;
        if (hanning) {                                      // step (2): taps at m-1, m, m+1; the outer two taps are 0
            r = 0.5 * sub_mean(m);
            if (m > 0) r += 0.25 * sub_mean(m - 1);
            if (m + 1 < nf_mid) r += 0.25 * sub_mean(m + 1);
        } else
            r = sub_mean(m);
        acc += r;
    }
    out[idx] = acc / (double)averaging;                     // step (3)
}

// Piecewise-linear regridding of an unstructured image (scattered points, e.g. RADMC-3D's circular
// images, pdspy/modeling/Model.py:536-558) onto the regular pixel grid the transform reads: every pixel
// carries the three vertices of the Delaunay triangle that contains it and its barycentric weights
// (computed once per geometry on the host); per likelihood call only this gather runs:
//     out[p, f] = scale * sum_k bary[p, k] * values[tri[p, k], f]        (0 outside the hull, tri < 0)
// i.e. a sparse matrix with three non-zeros per row applied to all channels; HBM-bound.
__global__ void __launch_bounds__(256) regrid_linear_kernel(const double *__restrict__ values, const int *__restrict__ tri,
                                                            const double *__restrict__ bary, int64_t npix, int nf,
                                                            double scale, double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * nf) return;
    const int64_t p = idx / nf;
    const int f = (int)(idx % nf);
    const int t0 = tri[3 * p];
    double v = 0.0;
    if (t0 >= 0) {
        v = bary[3 * p] * values[(int64_t)t0 * nf + f] + bary[3 * p + 1] * values[(int64_t)tri[3 * p + 1] * nf + f] +
            bary[3 * p + 2] * values[(int64_t)tri[3 * p + 2] * nf + f];
        v *= scale;
    }
    out[idx] = v;
}

}  // namespace pdsb

using namespace pdsb;

int pdsb_regrid_linear(const double *values, int64_t npts, const int *tri, const double *bary, int64_t npix, int nf,
                       double scale, int kind, double *out)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(npts > 0 && npix >= 0 && nf > 0, "sizes");
    if (npix == 0) return PDSB_OK;
    PDSB_REQUIRE(values && tri && bary && out, "arrays");
    const double *dv = values, *db = bary;
    const int *dt = tri;
    double *dout = out;
    if (kind == PDSB_HOST) {
        const size_t bv = (size_t)npts * nf * sizeof(double), bb = (size_t)npix * 3 * sizeof(double),
                     bt = (size_t)npix * 3 * sizeof(int);
        PDSB_CHECK(c.stage_a.ensure(bv + bb + bt + 256));
        unsigned char *p = c.stage_a.as<unsigned char>();
        PDSB_CHECK(copy_h2d(p, values, bv));
        dv = reinterpret_cast<const double *>(p);
        p += (bv + 63) / 64 * 64;
        PDSB_CHECK(copy_h2d(p, bary, bb));
        db = reinterpret_cast<const double *>(p);
        p += (bb + 63) / 64 * 64;
        PDSB_CHECK(copy_h2d(p, tri, bt));
        dt = reinterpret_cast<const int *>(p);
        PDSB_CHECK(c.stage_b.ensure((size_t)npix * nf * sizeof(double)));
        dout = c.stage_b.as<double>();
    }
    {
        LaunchScope ls("regrid_linear");
        regrid_linear_kernel<<<ceil_div(npix * nf, 256), 256, 0, c.stream>>>(dv, dt, db, npix, nf, scale, dout);
        PDSB_CUDA(cudaGetLastError());
    }
    if (kind == PDSB_HOST) {
        PDSB_CHECK(copy_d2h(out, dout, (size_t)npix * nf * sizeof(double)));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}

int pdsb_channel_slice(const double *image, int kind, int64_t npix, int nf_in, int c0, int nf_out, double *out_dev)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(npix >= 0 && nf_in > 0 && nf_out > 0 && c0 >= 0 && c0 + nf_out <= nf_in, "channel window");
    if (npix == 0) return PDSB_OK;
    PDSB_REQUIRE(image && out_dev, "arrays");
    const void *p = nullptr;
    PDSB_CHECK(to_device(image, kind, (size_t)npix * nf_in * sizeof(double), c.img64, &p));
    LaunchScope ls("channel_slice");
    channel_slice_kernel<<<ceil_div(npix * nf_out, 256), 256, 0, c.stream>>>(static_cast<const double *>(p), out_dev, npix,
                                                                           nf_in, c0, nf_out);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

int pdsb_channel_postprocess(const double *image, int64_t npix, int nf_in, int subsample, int hanning, int averaging,
                             int kind, double *out)
{
    return pdsb_channel_postprocess_scaled(image, npix, nf_in, subsample, hanning, averaging, nullptr, kind, out);
}

int pdsb_channel_postprocess_scaled(const double *image, int64_t npix, int nf_in, int subsample, int hanning, int averaging,
                                    const double *in_scale, int kind, double *out)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_REQUIRE(npix >= 0 && nf_in > 0 && subsample >= 1 && averaging >= 1, "sizes");
    PDSB_REQUIRE(nf_in % (subsample * averaging) == 0, "nf_in must be nf_out * averaging * subsample");
    if (npix == 0) return PDSB_OK;
    PDSB_REQUIRE(image && out, "arrays");
    const int nf_mid = nf_in / subsample, nf_out = nf_mid / averaging;
    const double *din = image;
    double *dout = out;
    if (kind == PDSB_HOST) {
        const void *p = nullptr;
        PDSB_CHECK(to_device(image, PDSB_HOST, (size_t)npix * nf_in * sizeof(double), c.stage_a, &p));
        din = static_cast<const double *>(p);
        PDSB_CHECK(c.stage_b.ensure((size_t)npix * nf_out * sizeof(double)));
        dout = c.stage_b.as<double>();
    }
    const double *dscale = nullptr;
    if (in_scale) {                                           // always a host array: nf_in doubles
        PDSB_CHECK(c.small_dev.ensure((size_t)nf_in * sizeof(double)));
        PDSB_CHECK(copy_h2d(c.small_dev.ptr, in_scale, (size_t)nf_in * sizeof(double)));
        dscale = c.small_dev.as<double>();
    }
    {
        LaunchScope ls("channel_post");
        channel_post_kernel<<<ceil_div(npix * nf_out, 256), 256, 0, c.stream>>>(din, dout, npix, nf_in, subsample,
                                                                               hanning ? 1 : 0, averaging, nf_mid, nf_out,
                                                                               dscale);
        PDSB_CUDA(cudaGetLastError());
    }
    if (kind == PDSB_HOST) {
        PDSB_CHECK(copy_d2h(out, dout, (size_t)npix * nf_out * sizeof(double)));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}
