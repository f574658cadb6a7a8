/* pdsb.h - C-ABI of the B200-native model-to-visibility path for pdspy.
 *
 * Plain C: pointers, sizes, opaque handles.  No torch / numpy types cross this
 * boundary.  Every entry point returns 0 on success and a non-zero code on failure
 * (PDSB_ERR_*); pdsb_last_error() returns a human-readable message for the last
 * failure on the calling thread.  Nothing throws.  There is NO CPU fallback: without
 * a CUDA device every compute entry point fails with PDSB_ERR_CUDA.
 *
 * Reference interfaces replaced (paths under the reference tree, psheehan/pdspy v2.0.8):
 *   pdsb_sample_image      galario.double.sampleImage as called by
 *                          pdspy/interferometry/interpolate_model.py:22-30 (per-channel loop,
 *                          row flip, conjugation, dRA/dDec phase)
 *   pdsb_loglike*          interpolate_model (above) followed by the visibility term of
 *                          pdspy/utils/emcee.py:31-43 == pdspy/utils/dynesty.py:47-59
 *   pdsb_chi2              the same likelihood term on caller-supplied model arrays
 *   pdsb_chisq             chisq()/chisq_calc  pdspy/interferometry/libinterferometry.pyx:610-633
 *   pdsb_grid              grid()              pdspy/interferometry/libinterferometry.pyx:313-541
 *   pdsb_freqcorrect       freqcorrect()       pdspy/interferometry/libinterferometry.pyx:587-608
 *
 * Threading: one calling thread per process (the reference is single-threaded Python,
 * one MPI rank per likelihood); N processes may share a GPU.
 *
 * Memory kinds: every array argument is either a host pointer (PDSB_HOST; pageable or
 * pinned) or a device pointer on the initialised device (PDSB_DEVICE); the `kind`
 * argument next to a group of pointers says which.  Host results are valid when the
 * call returns; device results are ordered on the library's stream (pdsb_get_stream).
 */
#ifndef PDSB_H
#define PDSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDSB_VERSION 1

enum { PDSB_HOST = 0, PDSB_DEVICE = 1 };

enum {
    PDSB_OK = 0,
    PDSB_ERR_ARG = 1,        /* bad argument (null pointer, size, enum)           */
    PDSB_ERR_CUDA = 2,       /* CUDA runtime error, or no device                  */
    PDSB_ERR_STATE = 3,      /* call out of order (no init, no data uploaded...)  */
    PDSB_ERR_NOMEM = 4
};

/* convolution kernels of grid(): libinterferometry.pyx:405-417 */
enum { PDSB_CONV_PILLBOX = 0, PDSB_CONV_EXPSINC = 1 };
/* weighting schemes of grid(): libinterferometry.pyx:429 */
enum { PDSB_WT_NATURAL = 0, PDSB_WT_UNIFORM = 1, PDSB_WT_SUPERUNIFORM = 2, PDSB_WT_ROBUST = 3 };
/* mode of grid(): libinterferometry.pyx:363-367 */
enum { PDSB_MODE_CONTINUUM = 0, PDSB_MODE_SPECTRALLINE = 1 };

typedef struct pdsb_dataset pdsb_dataset;

/* ---- lifecycle -------------------------------------------------------------- */
int pdsb_version(void);
int pdsb_device_count(int *count);
/* Select the device, create the library stream, read SM count / clocks.  Idempotent. */
int pdsb_init(int device);
int pdsb_shutdown(void);
const char *pdsb_last_error(void);
/* cudaStream_t the library launches on, as an integer. */
int pdsb_get_stream(uint64_t *stream);
/* Launch on a caller-owned stream instead (e.g. torch's current stream; 0 is the legacy default
 * stream).  pdsb_reset_stream goes back to the library's own stream. */
int pdsb_set_stream(uint64_t stream);
int pdsb_reset_stream(void);
int pdsb_synchronize(void);
int pdsb_device_info(int *sm_count, int *sm_clock_khz, int64_t *mem_bytes, int *cc_major, int *cc_minor);

/* ---- raw buffers (so callers need no other CUDA binding) ---------------------- */
int pdsb_device_alloc(void **ptr, int64_t bytes);
int pdsb_device_free(void *ptr);
int pdsb_host_alloc_pinned(void **ptr, int64_t bytes);
int pdsb_host_free_pinned(void *ptr);
/* kind_dst/kind_src: PDSB_HOST or PDSB_DEVICE; asynchronous on the library stream when the
 * host side is pinned, then pdsb_synchronize() before touching the host buffer. */
int pdsb_memcpy(void *dst, int kind_dst, const void *src, int kind_src, int64_t bytes);
int pdsb_memset(void *dev_ptr, int value, int64_t bytes);

/* ---- timing on the library stream (CUDA events) ------------------------------ */
int pdsb_timer_start(void);
int pdsb_timer_stop(double *elapsed_ms);      /* synchronises on the stop event */
/* Per-kernel device time: when enabled every kernel launch is bracketed by events.
 * pdsb_profile_get sums them (synchronising) for kernels whose name starts with `prefix`. */
int pdsb_profile_enable(int on);
int pdsb_profile_reset(void);
int pdsb_profile_get(const char *prefix, double *total_ms, int64_t *launches);
/* Number of kernels this library has launched since init (all kinds). */
int pdsb_launch_count(int64_t *count);

/* ---- dataset: the static part of one observation ------------------------------ */
/* u, v: [nuv] fp64, wavelengths at the mean frequency (what interpolate_model receives).
 * If the list is Hermitian-doubled (second half == minus first half, exactly, as the
 * reference's readers produce: readuvfits.py:68-73) the transform is evaluated for the
 * first half only and conjugated for the second; detected, never assumed. */
int pdsb_dataset_create(const double *u, const double *v, int64_t nuv, int kind, pdsb_dataset **out);
/* Observed visibilities for the likelihood: real, imag, weights [nuv, nf] fp64 row-major
 * (libinterferometry.pyx:16-22).  Precomputes sum(log(w/2pi)) over w>0 (emcee.py:32,37). */
int pdsb_dataset_set_data(pdsb_dataset *ds, const double *real, const double *imag,
                          const double *weights, int nf, int kind);
int pdsb_dataset_info(const pdsb_dataset *ds, int64_t *nuv, int64_t *nuv_unique, int *nf, int *hermitian);
int pdsb_dataset_destroy(pdsb_dataset *ds);

/* ---- image cube -> model visibilities ------------------------------------------ */
/* image: [ny, nx, nf] fp64, channel fastest: the reference's image[ny,nx,nf,1] buffer
 * (imaging/libimaging.pyx:11).  dxy, dRA, dDec in RADIANS (the caller applies
 * `*arcsec`, interpolate_model.py:20,24).  out_real/out_imag: [nuv, nf] fp64, in the
 * reference's final convention (after its imag -> -imag, interpolate_model.py:27):
 *   V_i(u,v) = sum_{j,c} image[j,c,i] exp(+2 pi i dxy (u (c - nx/2) + v (ny/2 - 1 - j)))
 *                                     exp(-2 pi i (u dRA + v dDec))
 * evaluated as an exact (direct) transform: fp32 products with fp64 phase seeding and
 * fp64 accumulation across row chunks. */
int pdsb_sample_image(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind,
                      double dxy, double dRA, double dDec,
                      double *out_real, double *out_imag, int out_kind);

/* Fused transform + likelihood: the model visibilities never leave the device.
 *   chi2[nf] (nullable): sum_k w (|d - m|^2) per channel
 *   lnlike:  -0.5*sum((d.re-m.re)^2 w) - L - 0.5*sum((d.im-m.im)^2 w) - L,
 *            L = sum(log(w[w>0]/2pi))      (emcee.py:31-43, verbatim incl. the doubled L)
 * Outputs are host doubles.  nf must equal the dataset's nf. */
int pdsb_loglike(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind,
                 double dxy, double dRA, double dDec, double *chi2, double *lnlike);
/* The same two calls with the model modifiers the reference applies in host passes around
 * interpolate_model folded into the device epilogue:
 *   chan_scale[nf] (host, nullable): V_i *= chan_scale[i] - the flux-calibration factor
 *       (run_disk_model.py:319, run_flared_model.py:303) times the per-channel extinction exp(-tau_i)
 *       (run_flared_model.py:286-299); scaling the image and scaling its transform are the same thing;
 *   ff_flux, ff_x0, ff_y0 (radians): free-free point source, added to the REAL part only, every channel:
 *       real += ff_flux*cos(2*3.14159*(u*ff_x0 + v*ff_y0))   (run_disk_model.py:329-334, model.py:102-104). */
int pdsb_sample_image_ex(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind,
                         double dxy, double dRA, double dDec, const double *chan_scale, double ff_flux,
                         double ff_x0, double ff_y0, double *out_real, double *out_imag, int out_kind);
int pdsb_loglike_ex(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind,
                    double dxy, double dRA, double dDec, const double *chan_scale, double ff_flux,
                    double ff_x0, double ff_y0, double *chi2, double *lnlike);
/* The reference's OWN algorithm for the same call, for numerical continuity with galario-based runs:
 * galario.double.sampleImage as interpolate_model.py:23-27 uses it - FFT of the (row-flipped, shifted)
 * image, bilinear interpolation at (|u|/du, n/2 + v/du), conjugate for u < 0, phase shift, imag -> -imag -
 * restated from galario's published algorithm (parity unpinned).  O(n^2 log n + nuv)
 * instead of the direct transform's O(n^2 nuv), at the price of galario's interpolation error (1e-3..4e-2
 * of max|V| off the FFT grid).  Square images, side a power of two <= 4096.
 *   pdsb_loglike_fft: out[0..3] = sum (d.re-m.re)^2 w, sum (d.im-m.im)^2 w, L, lnlike (as pdsb_chi2). */
int pdsb_sample_image_fft(pdsb_dataset *ds, const double *image, int n, int nf, int image_kind, double dxy,
                          double dRA, double dDec, double *out_real, double *out_imag, int out_kind);
int pdsb_loglike_fft(pdsb_dataset *ds, const double *image, int n, int nf, int image_kind, double dxy,
                     double dRA, double dDec, double *out);
/* interpolate_model(code="nufft") - an extension: the EXACT transform (what pdsb_sample_image sums directly) by a
 * type-2 non-uniform FFT: image divided by the transform of an 8-point "exponential of semicircle" kernel, zero-padded
 * to 2n, FFT of every channel, 8 x 8 kernel-weighted sum per (visibility, channel).  O(n^2 log n + 64 nuv) per channel
 * instead of O(n^2 nuv); 4e-8 of max|V| from the exact transform (galario's 2 x 2 bilinear scheme: 1e-3..4e-2).
 * Image [ny, nx, nf] with even sides <= 2048 (any even size, rectangular included: the image is embedded in the
 * power-of-two grid about its centre pixel (ny/2, nx/2)); other arguments and outputs as the two entry points above. */
int pdsb_sample_image_nufft(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                            double dRA, double dDec, double *out_real, double *out_imag, int out_kind);
int pdsb_loglike_nufft(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind, double dxy,
                       double dRA, double dDec, double *out);
/* Asynchronous variant for multi-GPU runs: chi2 per channel is left in DEVICE memory
 * (chi2_dev[nf]) on the library stream, ready for an NCCL all-reduce over uv shards; no host
 * synchronisation.  pdsb_dataset_logsum returns this shard's L. */
int pdsb_loglike_device(pdsb_dataset *ds, const double *image, int ny, int nx, int nf, int image_kind,
                        double dxy, double dRA, double dDec, double *chi2_dev);
int pdsb_dataset_logsum(const pdsb_dataset *ds, double *logsum);
/* nwalkers cubes [W, ny, nx, nf] against one dataset; dRA/dDec per walker [W] (host). */
int pdsb_loglike_batch(pdsb_dataset *ds, const double *images, int nwalkers, int ny, int nx, int nf,
                       int image_kind, double dxy, const double *dRA, const double *dDec,
                       double *lnlike);

/* interpolate_model(code="trift") (interpolate_model.py:49-55; the third-party `trift` package's "extended" mode:
 * the exact transform of the Delaunay piecewise-linear interpolant of a scattered-point image, PARITY UNPINNED like
 * sampleImage).  tri_records (HOST): ntri records of pdsb_triangle_record_bytes() bytes each =
 *   { double x[3], y[3] (vertex coordinates, radians, in the frame V = Int I exp(+2 pi i (u x - v y)));
 *     double B[9] (value at vertex a = sum_k B[3a + k] * values[idx[k]]: identity, or the barycentric rows of a
 *     sub-divided triangle); double area2 (twice the area); int idx[3]; int pad }.
 * The caller (pdspy_b200/interferometry/trift.py) triangulates once per geometry and sub-divides so that the
 * vertex phases of a triangle stay within ~4 rad of their mean for the longest baseline.  values [npts, nf] Jy/sr. */
int pdsb_sample_triangles(pdsb_dataset *ds, const void *tri_records, int ntri, const double *values, int64_t npts, int nf,
                          int values_kind, double dRA, double dDec, double *out_real, double *out_imag, int out_kind);
int pdsb_triangle_record_bytes(void);

/* ---- likelihood on caller-supplied model arrays --------------------------------- */
/* All arrays [n] fp64 (n = nuv*nf).  out[0]=sum (d.re-m.re)^2 w, out[1]=sum (d.im-m.im)^2 w,
 * out[2]=sum log(w[w>0]/2pi), out[3]=the emcee.py:31-43 value.  out is a host double[4]. */
int pdsb_chi2(const double *d_real, const double *d_imag, const double *weights,
              const double *m_real, const double *m_imag, int64_t n, int kind, double *out);
/* The same sums with the data side taken from the dataset handle's device-resident real / imag / weights
 * (pdsb_dataset_set_data): only the model arrays [nuv*nf] are read from the caller (`kind`).  This is what
 * utils.emcee.lnlike calls when the model Visibilities still live on the device (interpolate_model's lazy
 * result): no host round trip between interpolate_model.py:57 and emcee.py:31-43. */
int pdsb_chi2_dataset(pdsb_dataset *ds, const double *m_real, const double *m_imag, int kind, double *out);
/* chisq(): channel 0 of [nuv, nf] arrays, result rounded through a C float as the
 * reference's `cdef float chisq_calc` does (libinterferometry.pyx:616). */
int pdsb_chisq(const double *d_real, const double *d_imag, const double *weights,
               const double *m_real, const double *m_imag, int64_t nuv, int nf, int kind, float *out);

/* ---- convolutional gridding ------------------------------------------------------- */
/* Inputs as grid() has them after channel/mfs selection: u, v [nuv]; freq [nf];
 * real, imag, weights [nuv, nf] (weights NOT yet clamped: the clamp of :351 and the
 * zeroing of :353 happen inside).  uu, vv: [G] cell centres (numpy.linspace of :370-381,
 * computed by the caller because linspace's rounding is part of the reference result).
 * Outputs, host or device per out_kind (any may be NULL):
 *   out_real, out_imag, out_weights  [G*G, nch] fp64, row = v index, col = u index (:535-541)
 *   out_i, out_j                     [nuv, nf] uint32 index maps (:388-403)
 *   out_wmod                         [nuv, nf] fp64 weights after clamp + re-weighting
 *   n_outside                        count of (k,n) outside the grid (the WARNING of :424)
 * imaging: 0 / 1 as the reference's flag; 2 = leave the raw sums (see pdsb_grid_normalise).
 * deterministic != 0: every cell is accumulated in the reference's (k, n) order, so
 * pillbox maps are bit-exact against the reference; 0: shared-memory tiles + atomics. */
int pdsb_grid(const double *u, const double *v, const double *freq,
              const double *real, const double *imag, const double *weights,
              int64_t nuv, int nf, int in_kind,
              int gridsize, double binsize, const double *uu, const double *vv,
              int convolution, int weighting, double robust, int npixels, int mode, int imaging,
              int deterministic,
              double *out_real, double *out_imag, double *out_weights,
              uint32_t *out_i, uint32_t *out_j, double *out_wmod, int out_kind,
              int64_t *n_outside);

/* Multi-GPU gridding with bit-exact results: each process owns a band of
 * output rows.  After pdsb_set_grid_band(row_lo, row_hi) the main scatter of pdsb_grid keeps only
 * contributions to rows [row_lo, row_hi) (ordered mode, natural weighting, imaging = 2 raw sums only);
 * the bands are then summed (adding exact zeros) and normalised with pdsb_grid_normalise.  (0, 0) lifts
 * the restriction. */
int pdsb_set_grid_band(int row_lo, int row_hi);
/* Re-weighting (uniform / superuniform / robust, :429-485) when the data or the output are split over processes.
 * pdsb_grid_weights_map: the first half - box sums of the clamped weights over +-npixels cells (superuniform: 3)
 * around every home cell into binned_dev [gridsize^2 * nch] (device), and the per-channel sums of the clamped
 * weights into sumw_host [nf]; honours pdsb_set_grid_band.  with_ones = 0: raw sums (data split over the ranks:
 * the caller all-reduces map and sums and adds the ones of :430); with_ones = 1: the rows of the band start at 1.0
 * and the sums are added in (k, n) order on top, exactly as on one GPU (output rows split over the ranks: the
 * all-reduce adds exact zeros).  The finished map goes back through pdsb_set_grid_reweight (NULL resets); while
 * set, pdsb_grid with a re-weighting scheme skips its own box sums and uses the map (and, robust, the sums). */
int pdsb_grid_weights_map(const double *u, const double *v, const double *freq, const double *real, const double *imag,
                          const double *weights, int64_t nuv, int nf, int in_kind, int gridsize, double binsize,
                          const double *uu, const double *vv, int weighting, int npixels, int mode, int deterministic,
                          int with_ones, double *binned_dev, double *sumw_host, int64_t *n_outside);
int pdsb_set_grid_reweight(const double *binned_dev, int64_t ncell, const double *sumw_host, int nf);
/* Multi-GPU gridding (throughput mode): every rank grids its share of the
 * visibilities with imaging = 2 (raw sums, no normalisation) into device maps, the maps are summed
 * over ranks (NCCL all-reduce), then this applies the :525-533 normalisation on the device maps. */
int pdsb_grid_normalise(double *real, double *imag, double *weights, int gridsize, int nch, int imaging);

/* freqcorrect(): u' = (u[:,None]*freq/fbar).ravel() etc.; arrays on host or device. */
int pdsb_freqcorrect(const double *u, const double *v, const double *freq, int64_t nuv, int nf,
                     double new_freq, int kind, double *out_u, double *out_v);

/* ---- callers either side of the path --------------------------------------- */
/* average() (libinterferometry.pyx:151-311), everything per element on the device: the weight clamp (:179-181),
 * the uvdist != 0 and on-grid filters (:183-189, :250-260), the numpy.round bin indices (:221-248, numpy's operation
 * order and float64 -> uint32 cast), the accumulation in the reference's (k, n) order (:262-277), the normalisation,
 * the channel-weighted mean positions and the compaction of the non-empty cells in (j, i) order (:279-311).
 * Every output is bit-identical to the reference.
 *   HOST inputs: u, v, uvdist [nuv] (the object's own uvdist; ignored with mfs), freq [nf], real / imag / weights
 *   [nuv, nf].  mfs != 0: rows become the (k, n) pairs at u freq[n] / mfs_freq with one channel each (freqcorrect,
 *   :161-170).  radial: 0 = (u, v) grid of gridsize^2 cells of `binsize`; 1 = gridsize linear radial bins of
 *   `binsize`; 2 = log radial bins - the caller passes log_uvdist = log10(uvdist) [nuv], log_min = log10(logmin) and
 *   dtemp (:207-221) so that the one transcendental is the host libm's, as in the reference (not with mfs).
 *   centres [gridsize] (HOST): the radial bin centres returned as u (:208-212); ignored for radial = 0.
 *   HOST outputs with room for gridsize^2 (radial: gridsize) positions: out_u, out_v [n_out], out_real / out_imag /
 *   out_weights [n_out, nch] with nch = nf (1 with mfs) when spectral, else 1.  n_dropped: rows with uvdist != 0
 *   that fall off the grid (the reference prints its WARNING when this is non-zero, :256-257). */
int pdsb_average(const double *u, const double *v, const double *uvdist, const double *log_uvdist, const double *freq,
                 const double *real, const double *imag, const double *weights, int64_t nuv, int nf, int mfs,
                 double mfs_freq, int gridsize, double binsize, int radial, double log_min, double dtemp,
                 const double *centres, int spectral, double *out_u, double *out_v, double *out_real,
                 double *out_imag, double *out_weights, int64_t *n_out, int64_t *n_dropped);
/* center(): data * conj(point model at (x0, y0)), pdspy/interferometry/center.py:5-25 with
 * point_model of model.py:102-104 (incl. its literal 3.14159).  x0, y0 in radians. */
int pdsb_center(const double *u, const double *v, const double *freq, const double *real, const double *imag,
                int64_t nuv, int nf, double mean_freq, double x0_rad, double y0_rad, int kind,
                double *out_real, double *out_imag);

/* Channel post-processing of a model cube before interpolate_model
 * (pdspy/modeling/run_flared_model.py:308-366): mean over blocks of `subsample` sub-channels, optional
 * Hanning smoothing along the channel axis (numpy.hanning(5)/sum, zero-padded, mode="same"), mean over
 * blocks of `averaging` channels.  image [npix, nf_in] (the [ny,nx,nf,1] cube or an unstructured
 * [npts,nf] image), out [npix, nf_in/subsample/averaging]; both host or both device (`kind`). */
int pdsb_channel_postprocess(const double *image, int64_t npix, int nf_in, int subsample, int hanning,
                             int averaging, int kind, double *out);
/* The same with a per-INPUT-channel factor in_scale[nf_in] (HOST array, or NULL) applied before the sub-sample
 * mean: the reference's extinction, image[:,:,i,:] *= extinction[i] (run_flared_model.py:286-299), which comes
 * before the channel post-processing and does not commute with it. */
int pdsb_channel_postprocess_scaled(const double *image, int64_t npix, int nf_in, int subsample, int hanning,
                                    int averaging, const double *in_scale, int kind, double *out);

/* Channels [c0, c0 + nf_out) of a cube image [npix, nf_in] (HOST or DEVICE, `kind`) as a compact DEVICE array
 * out_dev [npix, nf_out]: the cube a rank evaluates when the likelihood is partitioned over the box by frequency
 * channels (BASELINE.json north_star; pdspy_b200.dist.ShardedLikelihood(channels=...)).  The reference's per-channel
 * loop (pdspy/interferometry/interpolate_model.py:22-30) is what makes channels independent units. */
int pdsb_channel_slice(const double *image, int kind, int64_t npix, int nf_in, int c0, int nf_out, double *out_dev);

/* Piecewise-linear regridding of an unstructured image (interpolate_model code="galario-unstructured",
 * pdspy/interferometry/interpolate_model.py:32-47; the scattered images of Model.py:536-558): values
 * [npts, nf]; per output pixel the three vertices of its Delaunay triangle tri [npix, 3] (tri[p,0] < 0:
 * outside the hull -> 0) and barycentric weights bary [npix, 3]; out [npix, nf] = scale * weighted sum.
 * All host or all device (`kind`). */
int pdsb_regrid_linear(const double *values, int64_t npts, const int *tri, const double *bary, int64_t npix,
                       int nf, double scale, int kind, double *out);

/* invert(): the per-channel image synthesis of pdspy/interferometry/invert.py:63-84 from gridded
 * visibilities (grid(..., imaging=True)): g_real, g_imag [imsize*imsize, nch]; conv [imsize, imsize] =
 * conv_func(u, v, binsize, binsize) of :94-121 evaluated by the caller on the grid; image_out
 * [imsize, imsize, nch] = (fftshift(ifft2(ifftshift(real+i imag))).real*imsize^2 /
 * fftshift(ifft2(ifftshift(conv))).real)[:, ::-1].  imsize must be a power of two (<= 4096). */
int pdsb_invert_image(const double *g_real, const double *g_imag, const double *conv, int imsize, int nch,
                      int kind, double *image_out);

/* clean(): the Hogbom loop of pdspy/interferometry/clean.py:51-106 and the restore step :108-113.
 * Images are [ny, nx, nf] (the reference's image[:, :, :, 0]), beams [2ny, 2nx, nf]; all host or all
 * device (`kind`).
 *   pdsb_mad_std: astropy.stats.mad_std = 1.482602218505602 * median(|x - median(x)|), exact order
 *       statistics (numpy.median's mean of the two middle elements for even n); no NaN handling.
 *   pdsb_clean_loop: `dirty` in: dirty image, out: residuals.  beam_resid_max = (dirty_beam -
 *       clean_beam).max() (:57).  Writes model and mask (0/1), the iteration count and the last mask
 *       threshold.  Exactly equal maxima are resolved to the first in C order; the loop ends when
 *       nothing positive is left under the mask (the reference would select every unmasked pixel).
 *   pdsb_clean_restore: clean_image = fftconvolve(model, clean_beam, mode="same") + residuals as the
 *       direct sum over the non-zero model components in index order. */
int pdsb_mad_std(const double *x, int64_t n, int kind, double *out);
int pdsb_clean_loop(double *dirty, const double *dirty_beam, int ny, int nx, int nf, double beam_resid_max,
                    double gain, int maxiter, double nsigma, int kind, double *model, double *mask, int *niter,
                    double *threshold_out);
int pdsb_clean_restore(const double *model, const double *clean_beam, const double *residuals, int ny, int nx,
                       int nf, int kind, double *clean_image);

/* 64-bit content hash of a HOST array (multi-threaded, memory-bound; needs no device).  The Python handle
 * cache compares it before trusting a device copy of caller-owned arrays (the reference mutates arrays in
 * place, invert.py:15-47). */
int pdsb_hash64(const void *host_ptr, int64_t bytes, uint64_t *out);

/* ---- tuning / measurement ----------------------------------------------------------- */
/* DFT kernel: 0 = auto (the FP32-pipe kernel the north star asks for); 1..2 = its two tilings;
 * 200 = the tcgen05/TMEM tensor-core kernel (lattice-split fp16 operands; 201 = the same with round 1's MMA issue
 * order, kept to time the difference); 300 = all-fp64 reference kernel (1e-13, ~8x slower).  All meet the same
 * 1e-5 parity bound; anything else is PDSB_ERR_ARG. */
int pdsb_set_dft_variant(int variant);
int pdsb_set_dft_split(int nsplit);          /* 0 = auto */
/* Register-resident FP32 FMA peak on all SMs (the roofline denominator bench.py reports beside the nominal one):
 * variant 0 = FFMA, 1 = FFMA2 (f32x2).  Returns achieved TFLOP/s (2 flop per FMA lane). */
int pdsb_bench_fma(int variant, int iters, double *tflops, double *ms);
/* One accumulator round of the tcgen05 DFT kernel's MMA sequence (M = N = 128, K = 64 as 12 MMAs of K = 16 on
 * fp16 hi/lo operands) with the accumulator pre-loaded with c0, on HOST operands a_*[128][64], b_*[128][64]
 * (fp16 bit patterns, row-major, k fastest); out[128][128] fp32 (host).  order 0 = the kernel's issue order,
 * 1 = cross products first.  Measures how the tensor core rounds its accumulator. */
int pdsb_tc5_accum_probe(const uint16_t *a_hi, const uint16_t *a_lo, const uint16_t *b_hi, const uint16_t *b_lo,
                         float c0, int order, float *out);

#ifdef __cplusplus
}
#endif
#endif /* PDSB_H */
