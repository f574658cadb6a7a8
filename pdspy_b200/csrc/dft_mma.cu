// EXPERIMENTAL opt-in variant of the direct Fourier sampling (pdsb_set_dft_variant(100)):
// the same separable, mirror-folded contraction as dft.cu, but the column contraction
//     C[uv, (row, comp)] = sum_t trig[uv, t] * X[t, (row, comp)]
// runs on the warp-level tensor-core path (mma.sync m16n8k16, fp16 inputs, fp32 accumulate) with both
// operands split into fp16 hi + lo parts (3 MMAs per product: hi*hi + hi*lo + lo*hi; the dropped lo*lo
// term is 2^-22 relative).  The folded image is scaled per plane by a power of two so that its fp16
// split keeps ~22 bits; numerics emulated in numpy before writing this: 1.5e-7 of max|V|.
//
// NOT the default: BASELINE.json's north star asks for an FP32-pipe kernel with FP32-pipe evidence
// (dft.cu).  This file exists to measure what the tensor cores give on the identical data flow
// (TMA ring, fp64 phase seeds, fp32 row rotation, fp64 accumulation, same partial-sum layout).
#include "dft.cuh"
#include <cuda_fp16.h>
#include <algorithm>

namespace pdsb {

constexpr int MMA_KT = 32;                         // column pairs per K tile
constexpr int MMA_ROW_BYTES = 80;                  // 32 halfs + 8 halfs of padding (conflict-free ldmatrix)
constexpr int MMA_ROWS = DFT_RC * 4;               // B rows per chunk: 32 row pairs x 4 components
constexpr int MMA_PART_BYTES = MMA_ROWS * MMA_ROW_BYTES;     // one of hi / lo
constexpr int MMA_CHUNK_BYTES = 2 * MMA_PART_BYTES;          // 20480

// ---- PTX helpers (same TMA / mbarrier protocol as dft.cu) ----
__device__ __forceinline__ uint32_t m_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void m_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool m_mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(m_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void m_tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     m_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(m_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_half(float x, __half &hi, __half &lo)
{
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_half2(__half a, __half b)
{
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// ---- per-plane |pixel| maximum -> power-of-two scale ----
__global__ void __launch_bounds__(256) plane_absmax_kernel(const double *__restrict__ img, int64_t npix, int nf,
                                                           int64_t nstrips, unsigned long long *__restrict__ planemax)
{
    // thread = (pixel strip, channel), channel fastest: coalesced over channels.  With few channels
    // hundreds of thousands of threads would hit the same few addresses, so the block's values are first
    // combined per channel in shared memory: one atomic per (block, channel).
    __shared__ double sm[256];
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double m = 0.0;
    if (idx < nstrips * nf) {
        const int p = (int)(idx % nf);
        for (int64_t pix = idx / nf; pix < npix; pix += nstrips) {
            const double a = fabs(img[pix * nf + p]);
            m = a > m ? a : m;            // NaNs are ignored (comparison false)
        }
        if (nf > 256) {
            if (m > 0.0) atomicMax(planemax + p, (unsigned long long)__double_as_longlong(m));
            return;
        }
    }
    if (nf > 256) return;
    sm[threadIdx.x] = m;
    __syncthreads();
    if ((int)threadIdx.x < nf) {
        // entries j of this block with (blockIdx.x*256 + j) % nf == my channel
        const int first = (int)(((int64_t)blockIdx.x * 256) % nf);
        const int p = ((int)threadIdx.x + first) % nf;        // channel of entry j = threadIdx.x
        double best = 0.0;
        for (int j = threadIdx.x; j < 256; j += nf) best = sm[j] > best ? sm[j] : best;
        if (best > 0.0) atomicMax(planemax + p, (unsigned long long)__double_as_longlong(best));
    }
}

__global__ void plane_scale_kernel(const unsigned long long *__restrict__ planemax, int nf, double *__restrict__ scale,
                                   double *__restrict__ unscale)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nf) return;
    const double m = __longlong_as_double((long long)planemax[p]);
    double s = 1.0;
    // folded values are bounded by 4*max|pixel|; put that bound at 2^14 (fp16 max is 65504)
    if (m > 0.0 && isfinite(m)) s = exp2(floor(log2(4096.0 / m)));
    if (!(s > 0.0) || !isfinite(s)) s = 1.0;
    scale[p] = s;
    unscale[p] = 1.0 / s;
}

// ---- fold into fp16 hi/lo operand tiles in ldmatrix order ----
// Bh[plane][ktile][chunk][hi|lo][row n][40 halfs]; n = ng*16 + type*8 + s_l*2 + c with the chunk's row
// pair s = ng*4 + s_l, component = type*2 + c (SS, SD | DS, DD): an n8 tile is 4 row pairs x the two
// components that share a trig factor (cos for SS/SD, sin for DS/DD).
__global__ void __launch_bounds__(256) fold_half_kernel(const double *__restrict__ img, unsigned char *__restrict__ B,
                                                        const double *__restrict__ scale, int ny, int nx, int nf,
                                                        int npx, int npy, int nkt, int nchunk)
{
    const int64_t tw = (int64_t)nkt * MMA_KT;
    const int64_t sw = (int64_t)nchunk * DFT_RC;
    const int64_t total = (int64_t)nf * tw * sw;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int p = (int)(idx % nf);
    const int64_t r = idx / nf;
    const int t = (int)(r % tw);
    const int s = (int)(r / tw);
    double pp = 0, mp = 0, pm = 0, mm = 0;
    if (t < npx && s < npy) {
        int c_hi, c_lo, j_lo, j_hi;
        if (nx % 2 == 0) { c_hi = nx / 2 + t; c_lo = nx / 2 - 1 - t; }
        else { c_hi = (nx - 1) / 2 + t; c_lo = (nx - 1) / 2 - t; }
        if (ny % 2 == 0) { j_lo = ny / 2 - 1 - s; j_hi = ny / 2 + s; }
        else { j_lo = (ny - 1) / 2 - s; j_hi = (ny - 1) / 2 + s; }
        const bool selfc = (c_hi == c_lo), selfr = (j_lo == j_hi);
        pp = img[((int64_t)j_lo * nx + c_hi) * nf + p];
        if (!selfc) mp = img[((int64_t)j_lo * nx + c_lo) * nf + p];
        if (!selfr) pm = img[((int64_t)j_hi * nx + c_hi) * nf + p];
        if (!selfc && !selfr) mm = img[((int64_t)j_hi * nx + c_lo) * nf + p];
    }
    const double Sp = pp + mp, Dp = pp - mp, Sm = pm + mm, Dm = pm - mm;
    const double comp[4] = {Sp + Sm, Sp - Sm, Dp + Dm, Dp - Dm};
    const double sc = scale[p];
    const int kt = t / MMA_KT, tl = t % MMA_KT;
    const int chunk = s / DFT_RC, sl_chunk = s % DFT_RC;
    const int ng = sl_chunk / 4, s_l = sl_chunk % 4;
    unsigned char *base = B + (((int64_t)p * nkt + kt) * nchunk + chunk) * (int64_t)MMA_CHUNK_BYTES;
#pragma unroll
    for (int cidx = 0; cidx < 4; cidx++) {
        const int type = cidx >> 1, c = cidx & 1;
        const int n = ng * 16 + type * 8 + s_l * 2 + c;
        const double x = comp[cidx] * sc;
        const __half hi = __double2half(x);
        const __half lo = __double2half(x - (double)__half2float(hi));
        *reinterpret_cast<__half *>(base + (size_t)n * MMA_ROW_BYTES + tl * 2) = hi;
        *reinterpret_cast<__half *>(base + MMA_PART_BYTES + (size_t)n * MMA_ROW_BYTES + tl * 2) = lo;
    }
}

// ---- the kernel: 128 threads = 4 warps; a warp owns MT m16 tiles = 16*MT uv points ----
// One CTA: 4 warps x MT m16 tiles of uv points, a range of K tiles (column tiles) and a GROUP of `pg`
// planes (channels share uv points, hence the trig fragments: they are generated once per K tile and
// reused for every plane of the group).  Loop order: K tile > plane > chunk of 32 row pairs.
template <int MT, int MINB, int UNR>
__global__ void __launch_bounds__(DFT_THREADS, MINB) dft_mma_kernel(const DftParams P,
                                                                    const unsigned char *__restrict__ Bg, int nkt, int pg)
{
    constexpr int NQ = 2 * MT;                    // uv points this thread contributes to (rows g and g+8 of each tile)
    constexpr int UVB = 4 * 16 * MT;              // uv points per block
    constexpr int RESEED = 4;                     // chunks between fp64-seeded row phases (32 rotations of D^4)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *fst = reinterpret_cast<double *>(smem_raw + (size_t)DFT_NSTAGE * MMA_CHUNK_BYTES);
    __shared__ __align__(8) uint64_t full_bar[DFT_NSTAGE];
#define MFU(q) fst[(0 * NQ + (q)) * DFT_THREADS + tid]
#define MFV(q) fst[(1 * NQ + (q)) * DFT_THREADS + tid]
#define MVR(q) fst[(2 * NQ + (q)) * DFT_THREADS + tid]
#define MVI(q) fst[(3 * NQ + (q)) * DFT_THREADS + tid]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, j = lane & 3;
    const int plane0 = blockIdx.z * pg, sp = blockIdx.y;
    const int npl = (P.nf - plane0) < pg ? (P.nf - plane0) : pg;
    const int kt0 = (int)(((int64_t)sp * nkt) / P.nsplit), kt1 = (int)(((int64_t)(sp + 1) * nkt) / P.nsplit);
    const int nkl = kt1 - kt0;
    const int per_k = npl * P.nchunk;
    const int nit = nkl * per_k;
    auto chunk_src = [&](int it) -> const unsigned char * {
        const int kl = it / per_k, rem = it - kl * per_k;
        const int pl = rem / P.nchunk, ch = rem - pl * P.nchunk;
        return Bg + (((size_t)(plane0 + pl) * nkt + (kt0 + kl)) * (size_t)P.nchunk + ch) * MMA_CHUNK_BYTES;
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DFT_NSTAGE; s++) m_mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DFT_NSTAGE; s++)
            if (s < nit) {
                m_mbar_expect_tx(&full_bar[s], MMA_CHUNK_BYTES);
                m_tma_load_1d(smem_raw + (size_t)s * MMA_CHUNK_BYTES, chunk_src(s), MMA_CHUNK_BYTES, &full_bar[s]);
            }
    }

    // this thread's uv points: q = 2*m + h  ->  row (h ? g+8 : g) of m16 tile m of this warp
    float D4r[NQ], D4i[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int64_t k = (int64_t)blockIdx.x * UVB + warp * (16 * MT) + (q >> 1) * 16 + ((q & 1) ? g + 8 : g);
        const bool valid = k < P.nuvh;
        const double fuq = valid ? P.u[k] * P.dxy : 0.0, fvq = valid ? P.v[k] * P.dxy : 0.0;
        MFU(q) = fuq;
        MFV(q) = fvq;
        const double a4 = 4.0 * fvq;             // this thread's rows advance by 4 per n-group
        double s, c;
        sincospi(2.0 * (a4 - rint(a4)), &s, &c);
        D4r[q] = (float)c;
        D4i[q] = (float)s;
    }

    // ldmatrix row address of this lane inside an n8 tile: matrix = lane/8 (k segment), row = lane%8
    const uint32_t ld_off = (uint32_t)((lane & 7) * MMA_ROW_BYTES + (lane >> 3) * 16);

    int it = 0;
    for (int kl = 0; kl < nkl; kl++) {
        // ---- A fragments of this K tile: trig[uv, t] split into fp16 hi + lo.  Per uv point one seed
        //      at column t0 = 32 kt + 2j and an fp64 rotation through the 8 columns this lane needs
        //      (offsets 0,1,8,9,16,17,24,25: steps +1,+7,+1,+7,...).
        uint32_t Ach[MT][2][4], Acl[MT][2][4], Ash[MT][2][4], Asl[MT][2][4];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const double fuq = MFU(q);
            const int m = q >> 1, h = q & 1;
            double a0 = fuq * ((double)((kt0 + kl) * MMA_KT + 2 * j) + P.hx);
            double a1 = fuq, a7 = 7.0 * fuq;
            a0 -= rint(a0);
            a1 -= rint(a1);
            a7 -= rint(a7);
            float sf, cf;
            sincospif((float)(2.0 * a0), &sf, &cf);
            double cr = cf, ci = sf;
            sincospif((float)(2.0 * a1), &sf, &cf);
            const double r1c = cf, r1s = sf;
            sincospif((float)(2.0 * a7), &sf, &cf);
            const double r7c = cf, r7s = sf;
#pragma unroll
            for (int pi = 0; pi < 4; pi++) {
                __half ch[2], cl[2], sh[2], sl[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    split_half((float)cr, ch[e], cl[e]);
                    split_half((float)ci, sh[e], sl[e]);
                    const double rc = e ? r7c : r1c, rs = e ? r7s : r1s;     // +1 then +7
                    const double nr = cr * rc - ci * rs;
                    ci = cr * rs + ci * rc;
                    cr = nr;
                }
                const int ks = pi >> 1, reg = (pi & 1) * 2 + h;
                Ach[m][ks][reg] = pack_half2(ch[0], ch[1]);
                Acl[m][ks][reg] = pack_half2(cl[0], cl[1]);
                Ash[m][ks][reg] = pack_half2(sh[0], sh[1]);
                Asl[m][ks][reg] = pack_half2(sl[0], sl[1]);
            }
        }

        for (int pl = 0; pl < npl; pl++) {
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                MVR(q) = 0.0;
                MVI(q) = 0.0;
            }
            float Er[NQ], Ei[NQ];
            for (int ch = 0; ch < P.nchunk; ch++, it++) {
                const int st = it % DFT_NSTAGE;
                const uint32_t parity = (uint32_t)((it / DFT_NSTAGE) & 1);
                float vre[NQ], vim[NQ];
                if (ch % RESEED == 0) {
#pragma unroll
                    for (int q = 0; q < NQ; q++) {
                        double b0 = MFV(q) * ((double)(ch * DFT_RC + j) + P.hy);      // this thread's rows: j mod 4
                        b0 -= rint(b0);
                        sincospif((float)(2.0 * b0), &Ei[q], &Er[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    vre[q] = 0.f;
                    vim[q] = 0.f;
                }
                while (!m_mbar_try_wait(&full_bar[st], parity)) {
                }
                const uint32_t sbase = m_smem_u32(smem_raw + (size_t)st * MMA_CHUNK_BYTES);

#pragma unroll UNR
                for (int ng = 0; ng < DFT_RC / 4; ng++) {
                    float C[2][MT][4];
#pragma unroll
                    for (int ty = 0; ty < 2; ty++)
#pragma unroll
                        for (int m = 0; m < MT; m++)
#pragma unroll
                            for (int e = 0; e < 4; e++) C[ty][m][e] = 0.f;
#pragma unroll
                    for (int ty = 0; ty < 2; ty++) {
                        const uint32_t rowaddr = sbase + (uint32_t)((ng * 16 + ty * 8) * MMA_ROW_BYTES) + ld_off;
                        uint32_t bh[4], bl[4];
                        ldmatrix_x4(rowaddr, bh[0], bh[1], bh[2], bh[3]);
                        ldmatrix_x4(rowaddr + MMA_PART_BYTES, bl[0], bl[1], bl[2], bl[3]);
#pragma unroll
                        for (int m = 0; m < MT; m++)
#pragma unroll
                            for (int ks = 0; ks < 2; ks++) {
                                const uint32_t(&ahi)[4] = ty ? Ash[m][ks] : Ach[m][ks];
                                const uint32_t(&alo)[4] = ty ? Asl[m][ks] : Acl[m][ks];
                                mma16816(C[ty][m], ahi, bh[2 * ks], bh[2 * ks + 1]);
                                mma16816(C[ty][m], ahi, bl[2 * ks], bl[2 * ks + 1]);
                                mma16816(C[ty][m], alo, bh[2 * ks], bh[2 * ks + 1]);
                            }
                    }
                    // row-phase epilogue for this thread's row (s = ch*32 + ng*4 + j) of each of its uv points
#pragma unroll
                    for (int q = 0; q < NQ; q++) {
                        const int m = q >> 1, h = q & 1;
                        const float SS = C[0][m][2 * h], SD = C[0][m][2 * h + 1], DS = C[1][m][2 * h],
                                    DD = C[1][m][2 * h + 1];
                        vre[q] = fmaf(Er[q], SS, vre[q]);
                        vre[q] = fmaf(-Ei[q], DD, vre[q]);
                        vim[q] = fmaf(Er[q], DS, vim[q]);
                        vim[q] = fmaf(Ei[q], SD, vim[q]);
                        const float nr = Er[q] * D4r[q] - Ei[q] * D4i[q];
                        Ei[q] = Er[q] * D4i[q] + Ei[q] * D4r[q];
                        Er[q] = nr;
                    }
                }
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    MVR(q) += (double)vre[q];
                    MVI(q) += (double)vim[q];
                }
                __syncthreads();
                if (tid == 0 && it + DFT_NSTAGE < nit) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    m_mbar_expect_tx(&full_bar[st], MMA_CHUNK_BYTES);
                    m_tma_load_1d(smem_raw + (size_t)st * MMA_CHUNK_BYTES, chunk_src(it + DFT_NSTAGE), MMA_CHUNK_BYTES,
                                  &full_bar[st]);
                }
            }
            // this (K tile, plane) pass is complete: the four lanes of a quad hold the sums over rows = j mod 4
            // of the same uv points; reduce and add into the partial-sum array (this thread owns the entry)
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                double vr = MVR(q), vi = MVI(q);
                vr += __shfl_xor_sync(0xffffffffu, vr, 1);
                vi += __shfl_xor_sync(0xffffffffu, vi, 1);
                vr += __shfl_xor_sync(0xffffffffu, vr, 2);
                vi += __shfl_xor_sync(0xffffffffu, vi, 2);
                const int64_t k = (int64_t)blockIdx.x * UVB + warp * (16 * MT) + (q >> 1) * 16 + ((q & 1) ? g + 8 : g);
                if (j == 0 && k < P.nuvh) {
                    double2 *dst = P.part + ((size_t)sp * P.nf + (plane0 + pl)) * (size_t)P.nuvh + k;
                    if (kl == 0) *dst = make_double2(vr, vi);
                    else {
                        const double2 o = *dst;
                        *dst = make_double2(o.x + vr, o.y + vi);
                    }
                }
            }
        }
    }
#undef MFU
#undef MFV
#undef MVR
#undef MVI
}

// ---- host side ----
size_t mma_operand_bytes(int ny, int nx, int nf)
{
    const int npx = (nx + 1) / 2, npy = (ny + 1) / 2;
    const int nkt = (npx + MMA_KT - 1) / MMA_KT, nchunk = (npy + DFT_RC - 1) / DFT_RC;
    return (size_t)nf * nkt * nchunk * MMA_CHUNK_BYTES;
}

// per-plane power-of-two scale factors: scale_ws = [pmax (as u64) | scale | unscale], nf doubles each
int launch_plane_scale(const double *img_dev, double *scale_ws, int ny, int nx, int nf)
{
    Context &c = ctx();
    unsigned long long *pmax = reinterpret_cast<unsigned long long *>(scale_ws);
    PDSB_CUDA(cudaMemsetAsync(pmax, 0, (size_t)nf * sizeof(unsigned long long), c.stream));
    LaunchScope ls("plane_absmax");
    const int64_t npix = (int64_t)ny * nx;
    int64_t nstrips = ((int64_t)c.sm_count * 2048 + nf - 1) / nf;
    if (nstrips > npix) nstrips = npix;
    plane_absmax_kernel<<<ceil_div(nstrips * nf, 256), 256, 0, c.stream>>>(img_dev, npix, nf, nstrips, pmax);
    plane_scale_kernel<<<ceil_div(nf, 128), 128, 0, c.stream>>>(pmax, nf, scale_ws + nf, scale_ws + 2 * nf);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

// image (device fp64 [ny,nx,nf]) -> fp16 operand tiles + per-plane unscale factors (device double[nf])
int launch_fold_half(const double *img_dev, unsigned char *B, double *scale_ws /* [3*nf] */, int ny, int nx, int nf)
{
    Context &c = ctx();
    const int npx = (nx + 1) / 2, npy = (ny + 1) / 2;
    const int nkt = (npx + MMA_KT - 1) / MMA_KT, nchunk = (npy + DFT_RC - 1) / DFT_RC;
    PDSB_CHECK(launch_plane_scale(img_dev, scale_ws, ny, nx, nf));
    const int64_t total = (int64_t)nf * nkt * MMA_KT * nchunk * DFT_RC;
    LaunchScope ls("fold_half");
    fold_half_kernel<<<ceil_div(total, 256), 256, 0, c.stream>>>(img_dev, B, scale_ws + nf, ny, nx, nf, npx, npy, nkt, nchunk);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

static int mma_mt(int variant) { return variant == 101 || variant == 103 ? 4 : 2; }
static int mma_pg(int nf) { return nf < 8 ? nf : 8; }      // planes per CTA (share the trig fragments)

int mma_auto_split(int64_t nuvh, int nf, int nx)
{
    Context &c = ctx();
    const int UVB = 4 * 16 * mma_mt(c.dft_variant);
    const int nkt = ((nx + 1) / 2 + MMA_KT - 1) / MMA_KT;
    if (c.dft_split > 0) return c.dft_split < nkt ? c.dft_split : nkt;
    const int64_t uvtiles = (nuvh + UVB - 1) / UVB;
    const int pgroups = (nf + mma_pg(nf) - 1) / mma_pg(nf);
    const int64_t capacity = (int64_t)c.sm_count * 3, base = std::max<int64_t>(1, uvtiles * pgroups);
    const int64_t ns = (20 * capacity + base - 1) / base;
    return (int)(ns < 1 ? 1 : (ns > nkt ? nkt : ns));
}

template <int MT, int MINB, int UNR>
static int launch_mma_variant(DftParams p, const unsigned char *B, int nkt, const char *name)
{
    Context &c = ctx();
    constexpr int UVB = 4 * 16 * MT;
    constexpr size_t smem = (size_t)DFT_NSTAGE * MMA_CHUNK_BYTES + (size_t)4 * (2 * MT) * DFT_THREADS * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        PDSB_CUDA(cudaFuncSetAttribute(dft_mma_kernel<MT, MINB, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        attr_set = true;
    }
    const int64_t uvtiles = (p.nuvh + UVB - 1) / UVB;
    if (uvtiles <= 0) return PDSB_OK;
    const int pg = mma_pg(p.nf);
    dim3 grid((unsigned)uvtiles, (unsigned)p.nsplit, (unsigned)((p.nf + pg - 1) / pg));
    LaunchScope ls(name);
    dft_mma_kernel<MT, MINB, UNR><<<grid, DFT_THREADS, smem, c.stream>>>(p, B, nkt, pg);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

int launch_dft_mma(DftParams p, const unsigned char *B, int ny, int nx)
{
    Context &c = ctx();
    const int npx = (nx + 1) / 2, npy = (ny + 1) / 2;
    const int nkt = (npx + MMA_KT - 1) / MMA_KT;
    p.nchunk = (npy + DFT_RC - 1) / DFT_RC;
    PDSB_REQUIRE(p.nsplit >= 1 && p.nsplit <= nkt, "mma split");
    PDSB_REQUIRE(p.nf <= 65535 * 8, "grid z dimension");
    switch (c.dft_variant) {
        case 100: return launch_mma_variant<2, 3, 1>(p, B, nkt, "dft_mma_mt2");
        case 101: return launch_mma_variant<4, 2, 1>(p, B, nkt, "dft_mma_mt4");
        case 102: return launch_mma_variant<2, 3, 2>(p, B, nkt, "dft_mma_mt2_unr2");
        case 103: return launch_mma_variant<4, 2, 2>(p, B, nkt, "dft_mma_mt4_unr2");
        case 104: return launch_mma_variant<2, 2, 2>(p, B, nkt, "dft_mma_mt2_unr2_occ2");
    }
    set_error("unknown tensor-core DFT variant %d", c.dft_variant);
    return PDSB_ERR_ARG;
}

}  // namespace pdsb
