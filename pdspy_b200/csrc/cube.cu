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

__global__ void __launch_bounds__(256) channel_post_kernel(const double *__restrict__ in, double *__restrict__ out,
                                                           int64_t npix, int nf_in, int subsample, int hanning,
                                                           int averaging, int nf_mid, int nf_out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * nf_out) return;
    const int k = (int)(idx % nf_out);
    const double *src = in + (idx / nf_out) * nf_in;
    auto sub_mean = [&](int m) -> double {                  // step (1): sums in index order, then one division
        double s = 0.0;
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

}  // namespace pdsb

using namespace pdsb;

int pdsb_channel_postprocess(const double *image, int64_t npix, int nf_in, int subsample, int hanning, int averaging,
                             int kind, double *out)
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
    {
        LaunchScope ls("channel_post");
        channel_post_kernel<<<ceil_div(npix * nf_out, 256), 256, 0, c.stream>>>(din, dout, npix, nf_in, subsample,
                                                                               hanning ? 1 : 0, averaging, nf_mid, nf_out);
        PDSB_CUDA(cudaGetLastError());
    }
    if (kind == PDSB_HOST) {
        PDSB_CUDA(cudaMemcpyAsync(out, dout, (size_t)npix * nf_out * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        PDSB_CUDA(cudaStreamSynchronize(c.stream));
    }
    return PDSB_OK;
}
