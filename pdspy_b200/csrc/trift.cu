// interpolate_model(code="trift") (pdspy/interferometry/interpolate_model.py:49-55): the EXACT Fourier transform of
// the piecewise-linear (Delaunay) interpolant of a scattered-point image - what the third-party `trift` package
// computes in its "extended" mode (absent here: PARITY UNPINNED, conventions anchored on the reference's other two
// codes for the same image; see oracle/trift.py).
//
//   V_f(u, v) = exp(-2 pi i (u dRA + v dDec)) * sum_T  Int_T I_f(r) exp(i q.r) dA,     q = 2 pi (u, -v)
//   Int_T lambda_a exp(i q.r) dA = 2 A_T exp(i pbar) * sum_k i^k h_k(y_a, y_a, y_b, y_c) / (k + 3)!
// with p_k = q.r_k the vertex phases, pbar their mean, y_k = p_k - pbar, and h_k the complete homogeneous symmetric
// polynomials (the Taylor expansion of the divided difference exp[i p_a, i p_a, i p_b, i p_c], Hermite-Genocchi):
// no division by phase differences, so nothing special happens when q is perpendicular to an edge or zero.  The
// host subdivides triangles until |y| stays below ~4 for the longest baseline (trift.py), so ~30 terms reach
// 1e-15; all recurrences are REAL (the i^k only routes a term to the real or the imaginary sum).
//
// One thread per (unique) uv point, looping over all triangles (records staged in shared memory by the block);
// blockIdx.y takes a chunk of TR_CH channels whose sums sit in registers.  fp64 throughout: this is the optional
// `--ftcode trift` path of the reference, O(nuv ntri), not the headline kernel.  The partial sums land in the same
// [plane][uv] buffer as the pixel kernels' and go through the same epilogues (centre phase, Hermitian halves,
// [nuv, nf] layout / fused chi^2).
#include "dft.cuh"

namespace pdsb {

constexpr int TR_THREADS = 128;
constexpr int TR_CH = 16;            // channels per thread
constexpr int TR_BATCH = 64;         // triangles staged per round
constexpr int TR_KMAX = 40;

struct TriRec {                      // 19 doubles + 3 ints
    double x[3], y[3];               // vertex coordinates (radians)
    double B[9];                     // B[a*3 + k]: value at vertex a = sum_k B[a][k] * image[idx[k]]  (sub-divided triangles)
    double area2;                    // 2 A
    int idx[3];
    int pad;
};

__constant__ double c_inv_fact[TR_KMAX + 4];     // 1 / (k + 3)!

__global__ void __launch_bounds__(TR_THREADS) trift_kernel(const TriRec *__restrict__ tris, int ntri,
                                                           const double *__restrict__ values, int nf,
                                                           const double *__restrict__ u, const double *__restrict__ v,
                                                           int64_t nuvh, double2 *__restrict__ part)
{
    __shared__ TriRec s_tri[TR_BATCH];
    const int64_t k = (int64_t)blockIdx.x * TR_THREADS + threadIdx.x;
    const int ch0 = blockIdx.y * TR_CH;
    const int nch = nf - ch0 < TR_CH ? nf - ch0 : TR_CH;
    const bool valid = k < nuvh;
    const double qx = valid ? 6.283185307179586476925286766559 * u[k] : 0.0;
    const double qy = valid ? -6.283185307179586476925286766559 * v[k] : 0.0;
    double ar[TR_CH], ai[TR_CH];
#pragma unroll
    for (int c = 0; c < TR_CH; c++) ar[c] = ai[c] = 0.0;

    for (int t0 = 0; t0 < ntri; t0 += TR_BATCH) {
        const int nb = ntri - t0 < TR_BATCH ? ntri - t0 : TR_BATCH;
        __syncthreads();
        {
            const int words = nb * (int)(sizeof(TriRec) / 8);
            const double *src = reinterpret_cast<const double *>(tris + t0);
            double *dst = reinterpret_cast<double *>(s_tri);
            for (int i = threadIdx.x; i < words; i += TR_THREADS) dst[i] = src[i];
        }
        __syncthreads();
        for (int t = 0; t < nb; t++) {
            const TriRec &T = s_tri[t];
            const double p0 = qx * T.x[0] + qy * T.y[0], p1 = qx * T.x[1] + qy * T.y[1], p2 = qx * T.x[2] + qy * T.y[2];
            const double pm = (p0 + p1 + p2) * (1.0 / 3.0);
            const double y0 = p0 - pm, y1 = p1 - pm, y2 = p2 - pm;
            const double ymax = fmax(fabs(y0), fmax(fabs(y1), fabs(y2)));
            int K = (int)(3.5 * ymax) + 16;                       // truncation below 1e-15 of the sum for ymax <= 6
            K = K > TR_KMAX ? TR_KMAX : K;
            // h_k over growing node sets, all real:  e1 = y0^k, e2 = h_k(y0, y1), s = h_k(y0, y1, y2),
            // g_a = h_k(y_a, y0, y1, y2) = s_k + y_a g_a(k-1)
            double e1 = 1.0, e2 = 1.0, s = 1.0, g0 = 1.0, g1 = 1.0, g2 = 1.0;
            double r0 = 0, r1 = 0, r2 = 0, i0 = 0, i1 = 0, i2 = 0;          // real / imaginary sums per vertex
            for (int kk = 0; kk < K; kk += 4) {
                // k = kk: +real, kk+1: +imag, kk+2: -real, kk+3: -imag
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double f = ((j & 2) ? -1.0 : 1.0) * c_inv_fact[kk + j];
                    if (j & 1) {
                        i0 = fma(f, g0, i0);
                        i1 = fma(f, g1, i1);
                        i2 = fma(f, g2, i2);
                    } else {
                        r0 = fma(f, g0, r0);
                        r1 = fma(f, g1, r1);
                        r2 = fma(f, g2, r2);
                    }
                    e1 *= y0;
                    e2 = fma(y1, e2, e1);
                    s = fma(y2, s, e2);
                    g0 = fma(y0, g0, s);
                    g1 = fma(y1, g1, s);
                    g2 = fma(y2, g2, s);
                }
            }
            double sm, cm;
            sincos(pm, &sm, &cm);
            cm *= T.area2;
            sm *= T.area2;
            // W_a = 2A e^{i pm} (r_a + i i_a); weights on the ORIGINAL points: Wo_k = sum_a B[a][k] W_a
            const double wr0 = cm * r0 - sm * i0, wi0 = cm * i0 + sm * r0;
            const double wr1 = cm * r1 - sm * i1, wi1 = cm * i1 + sm * r1;
            const double wr2 = cm * r2 - sm * i2, wi2 = cm * i2 + sm * r2;
#pragma unroll
            for (int kv = 0; kv < 3; kv++) {
                const double b0 = T.B[kv], b1 = T.B[3 + kv], b2 = T.B[6 + kv];
                const double wr = b0 * wr0 + b1 * wr1 + b2 * wr2, wi = b0 * wi0 + b1 * wi1 + b2 * wi2;
                const double *val = values + (size_t)T.idx[kv] * nf + ch0;
#pragma unroll
                for (int c = 0; c < TR_CH; c++) {
                    if (c < nch) {
                        const double f = __ldg(val + c);
                        ar[c] = fma(f, wr, ar[c]);
                        ai[c] = fma(f, wi, ai[c]);
                    }
                }
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int c = 0; c < TR_CH; c++)
            if (c < nch) part[(size_t)(ch0 + c) * (size_t)nuvh + k] = make_double2(ar[c], ai[c]);
    }
}

// partial sums of the triangle transform for every unique uv point and channel: part[nf][nuvh]
int launch_trift(const void *tris_dev, int ntri, const double *values_dev, int nf, const double *u, const double *v,
                 int64_t nuvh, double2 *part)
{
    Context &c = ctx();
    static bool table = false;
    if (!table) {
        double h[TR_KMAX + 4];
        double f = 6.0;                                  // 3!
        for (int k = 0; k < TR_KMAX + 4; k++) {
            h[k] = 1.0 / f;
            f *= (double)(k + 4);
        }
        PDSB_CUDA(cudaMemcpyToSymbol(c_inv_fact, h, sizeof(h)));
        table = true;
    }
    if (nuvh <= 0) return PDSB_OK;
    dim3 grid((unsigned)ceil_div(nuvh, TR_THREADS), (unsigned)ceil_div(nf, TR_CH));
    LaunchScope ls("trift");
    trift_kernel<<<grid, TR_THREADS, 0, c.stream>>>(reinterpret_cast<const TriRec *>(tris_dev), ntri, values_dev, nf, u, v,
                                                    nuvh, part);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

size_t trift_record_bytes() { return sizeof(TriRec); }

}  // namespace pdsb
