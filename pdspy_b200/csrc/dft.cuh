// Direct (exact) Fourier sampling of an image cube at (u,v) points: device code.
//
// Replaces galario.double.sampleImage as called from
// pdspy/interferometry/interpolate_model.py:22-30 (reference tree).  Written from the
// transform's definition, not from galario's FFT+interpolation algorithm:
//
//   V_i(u,v) = sum_{j,c} I[j,c,i] exp(+2 pi i (u (x_c - dRA) + v (y_j - dDec)))
//   x_c = dxy (c - nx/2),  y_j = dxy (ny/2 - 1 - j)
//
// Design (see DESIGN.md "DFT kernel"):
//  * The phase separates, exp(i(a_c + b_j)) = exp(i a_c) exp(i b_j): a thread owns UVT uv
//    points and keeps the COLUMN factors of one tile of columns in registers (seeded and
//    advanced in fp64, rounded to fp32), while the ROW factor is advanced down the rows by
//    an fp32 complex rotation, re-seeded from an fp64-reduced phase at every chunk of RC
//    row pairs.  No sincos per (pixel, uv) pair.
//  * The image is real, so columns c and nx-1-c (and rows j, ny-1-j) see conjugate
//    factors about the image centre.  The fold kernel stores, per 2x2 mirror quad, the four
//    parity combinations SS, SD, DS, DD (computed in fp64, stored fp32); the inner loop then
//    needs ONE fp32 FMA per pixel per uv point instead of two:
//        Re += cos(b) * sum_t SS cos(a_t) - sin(b) * sum_t DD sin(a_t)
//        Im += cos(b) * sum_t DS sin(a_t) + sin(b) * sum_t SD cos(a_t)
//  * Folded tiles are laid out so that everything one CTA streams is one contiguous run of
//    8/16 KB chunks; chunks are brought into shared memory by 1-D bulk TMA
//    (cp.async.bulk + mbarrier) through a 3-stage ring, and read back with warp-broadcast
//    LDS.128.
//  * Inner products are packed fp32x2 FMAs (FFMA2, new on sm_100) pairing adjacent columns,
//    which halves the issue slots per FMA lane; a scalar-FFMA variant is kept for comparison.
//  * fp32 partial sums live for one chunk (RC row pairs x TCP column pairs) and are then
//    added into fp64 accumulators.
#pragma once
#include "common.cuh"

namespace pdsb {

constexpr int DFT_RC = 32;          // row pairs per shared-memory chunk
constexpr int DFT_NSTAGE = 3;       // TMA ring depth
constexpr int DFT_THREADS = 128;

struct DftGeom {
    int ny, nx, nf;
    int npx, npy;        // column pairs, row pairs (ceil(n/2))
    int hx2, hy2;        // 1 when the pair offsets are half-integers (even size), else 0
    int tcp;             // column pairs per tile
    int ntile, nchunk;   // tiles along columns, chunks along rows (padded)
    double xcen, ycen;   // fold centre in pixel units times dxy: x_c - xcen = +-(t + hx) dxy
};

inline DftGeom make_geom(int ny, int nx, int nf, int tcp, double dxy)
{
    DftGeom g;
    g.ny = ny; g.nx = nx; g.nf = nf;
    g.npx = (nx + 1) / 2;
    g.npy = (ny + 1) / 2;
    g.hx2 = (nx % 2 == 0);
    g.hy2 = (ny % 2 == 0);
    g.tcp = tcp;
    g.ntile = (g.npx + tcp - 1) / tcp;
    g.nchunk = (g.npy + DFT_RC - 1) / DFT_RC;
    g.xcen = dxy * ((nx - 1) * 0.5 - (double)(nx / 2));
    g.ycen = dxy * ((double)(ny / 2) - 1.0 - (ny - 1) * 0.5);
    return g;
}

inline size_t folded_floats(const DftGeom &g)
{
    return (size_t)g.nf * g.ntile * g.nchunk * DFT_RC * 4 * g.tcp;
}

struct DftParams {
    const float *F;          // folded image
    const double *u, *v;     // unique uv points
    int64_t nuvh;
    double dxy;
    int ntile, nchunk, nsplit, nf;
    double hx, hy;           // 0.5 or 0
    double2 *part;           // [nsplit][nf][nuvh]
};

int launch_fold(const double *img_dev, float *F, const DftGeom &g);
int launch_dft(const DftParams &p, int variant, int *tcp_of_variant);
int dft_variant_tcp(int variant);
int dft_variant_count();
int dft_pick_variant(int nx);
int dft_auto_split(int variant, int64_t nuvh, int nf, int ntile);

// tcgen05 / TMEM variant (dft_tc5.cu), selected with pdsb_set_dft_variant(200)
constexpr int DFT_VARIANT_TC5 = 200;
size_t tc5_operand_bytes(int ny, int nx, int nf);
size_t tc5_ws_bytes(int ny, int nx, int nf);
const double *tc5_plane_unscale(void *ws, int ny, int nx, int nf);
int launch_fold_tc5(const double *img_dev, unsigned char *B, void *ws, int ny, int nx, int nf);
int tc5_auto_split(int64_t nuvh, int nf, int nx);
int launch_dft_tc5(DftParams p, const unsigned char *B, void *ws, int ny, int nx);

// NUFFT variant (vis.cu: type-2 non-uniform FFT, the exact transform at FFT cost), selected with pdsb_set_dft_variant(400)
constexpr int DFT_VARIANT_NUFFT = 400;

// fp64 reference variant (dft_f64.cu), selected with pdsb_set_dft_variant(300)
constexpr int DFT_VARIANT_F64 = 300;
int launch_dft_f64(const double *img_dev, int ny, int nx, int nf, const double *u, const double *v, int64_t nuvh, double dxy,
                   double2 *part);

// trift.cu: exact transform of a triangulated scattered-point image (interpolate_model code="trift")
int launch_trift(const void *tris_dev, int ntri, const double *values_dev, int nf, const double *u, const double *v,
                 int64_t nuvh, double2 *part);
size_t trift_record_bytes();

// fft.cu: inverse-sign 2-D transform of every channel of a cube, both-axes fftshift on input and output,
// channel-fastest output (used by the galario-algorithm path of vis.cu)
int fft2_planes(const double *cube_dev, int n, int nf, int flip, double2 *T, double2 *Y);
int rfft2_planes(const double *cube_dev, int n, int nf, int flip, double2 *T, double2 *Yh);
int rfft2_planes_padded(const double *cube_dev, int nsy, int nsx, int n, int nf, int flip, const double *corr_y,
                        const double *corr_x, double2 *T, double2 *Yh);

}  // namespace pdsb
