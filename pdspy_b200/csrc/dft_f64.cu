// High-precision variant of the direct Fourier sampling (pdsb_set_dft_variant(300)): the same separable
// sum in fp64 throughout, no fold, no fp32 anywhere - about 1e-13 of max|V| from the CPU oracle.  It is the
// on-device reference the full-size parity tests hold the FP32 and tensor-core kernels to on EVERY uv point
// (the CPU oracle manages a few thousand points of C2 / C3 in the time a test may take), and a mode for
// callers who want fp64 visibilities; ~7x slower than the default.
//
//   partial[plane][k] = sum_j exp(2 pi i fv ((ny-1)/2 - j)) * sum_c I[j, c, plane] exp(2 pi i fu (c - (nx-1)/2)),
//   fu = u_k dxy, fv = v_k dxy   (phases about the image centre; the epilogue of vis.cu adds the centre and
//   the dRA/dDec shift, as for the other variants)
//
// CTA = 128 uv points x one plane.  Column tiles of 16: the thread keeps its 16 column factors in registers
// (fp64-reduced seed -> sincospi, fp64 rotation); image tiles [32 rows x 16 columns] are staged in shared
// memory and read by warp broadcast; the row factor is an fp64 rotation re-seeded at every tile.
#include "dft.cuh"

namespace pdsb {

constexpr int F64_TC = 16, F64_RC = 32, F64_THREADS = 128;

__global__ void __launch_bounds__(F64_THREADS) dft_f64_kernel(const double *__restrict__ img, int ny, int nx, int nf,
                                                              const double *__restrict__ u, const double *__restrict__ v,
                                                              int64_t nuvh, double dxy, double2 *__restrict__ part)
{
    __shared__ __align__(16) double tile[F64_RC][F64_TC];
    const int plane = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * F64_THREADS + threadIdx.x;
    const bool valid = k < nuvh;
    const double fu = valid ? u[k] * dxy : 0.0, fv = valid ? v[k] * dxy : 0.0;
    double dcr, dci, drr, dri;                      // column step exp(2 pi i fu), row step exp(-2 pi i fv)
    sincospi(2.0 * (fu - rint(fu)), &dci, &dcr);
    sincospi(-2.0 * (fv - rint(fv)), &dri, &drr);
    const double xc = 0.5 * (double)(nx - 1), yc = 0.5 * (double)(ny - 1);
    double vr = 0.0, vi = 0.0;
    for (int c0 = 0; c0 < nx; c0 += F64_TC) {
        double cr[F64_TC], ci[F64_TC];
        {
            double a = fu * ((double)c0 - xc);
            sincospi(2.0 * (a - rint(a)), &ci[0], &cr[0]);
#pragma unroll
            for (int c = 1; c < F64_TC; c++) {
                cr[c] = cr[c - 1] * dcr - ci[c - 1] * dci;
                ci[c] = cr[c - 1] * dci + ci[c - 1] * dcr;
            }
        }
        for (int j0 = 0; j0 < ny; j0 += F64_RC) {
            __syncthreads();
            for (int e = threadIdx.x; e < F64_RC * F64_TC; e += F64_THREADS) {
                const int r = e / F64_TC, c = e % F64_TC;
                const int j = j0 + r, col = c0 + c;
                tile[r][c] = (j < ny && col < nx) ? img[((int64_t)j * nx + col) * nf + plane] : 0.0;
            }
            __syncthreads();
            double er, ei;
            {
                double b = fv * (yc - (double)j0);
                sincospi(2.0 * (b - rint(b)), &ei, &er);
            }
#pragma unroll 4
            for (int r = 0; r < F64_RC; r++) {
                double tr = 0.0, ti = 0.0;
#pragma unroll
                for (int c = 0; c < F64_TC; c++) {
                    const double x = tile[r][c];
                    tr = fma(x, cr[c], tr);
                    ti = fma(x, ci[c], ti);
                }
                vr += er * tr - ei * ti;
                vi += er * ti + ei * tr;
                const double nr = er * drr - ei * dri;
                ei = er * dri + ei * drr;
                er = nr;
            }
        }
    }
    if (valid) part[(size_t)plane * nuvh + k] = make_double2(vr, vi);
}

int launch_dft_f64(const double *img_dev, int ny, int nx, int nf, const double *u, const double *v, int64_t nuvh, double dxy,
                   double2 *part)
{
    Context &c = ctx();
    if (nuvh <= 0) return PDSB_OK;
    LaunchScope ls("dft_f64");
    dft_f64_kernel<<<dim3((unsigned)ceil_div(nuvh, F64_THREADS), (unsigned)nf), F64_THREADS, 0, c.stream>>>(img_dev, ny, nx, nf, u, v,
                                                                                                          nuvh, dxy, part);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

}  // namespace pdsb
