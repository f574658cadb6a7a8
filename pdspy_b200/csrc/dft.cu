// Direct Fourier sampling kernels (fold + DFT).  See dft.cuh for the design.
#include "dft.cuh"

namespace pdsb {

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk TMA, packed fp32x2 arithmetic.
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ float sum2(u64 a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    return lo + hi;
}

// ---------------------------------------------------------------------------------
// Fold: fp64 image [ny,nx,nf] (channel fastest, the reference's [ny,nx,nf,1] buffer) ->
// fp32 parity planes F[p][tile][s][comp][TCP], s over padded row pairs, comp = SS,SD,DS,DD.
//   "+" column of pair t: x offset +(t+hx) dxy  (c_hi);  "-" column: c_lo
//   "+" row of pair s:    y offset +(s+hy) dxy  (j_lo, because y_j decreases with j)
// A self-paired middle column/row (odd size) contributes once.
__global__ void __launch_bounds__(256) fold_image_kernel(const double *__restrict__ img, float *__restrict__ F,
                                                         int ny, int nx, int nf, int npx, int npy, int tcp,
                                                         int ntile, int nchunk)
{
    const int64_t tw = (int64_t)ntile * tcp;           // padded column pairs
    const int64_t sw = (int64_t)nchunk * DFT_RC;       // padded row pairs
    const int64_t total = (int64_t)nf * tw * sw;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int p = (int)(idx % nf);
    int64_t r = idx / nf;
    const int t = (int)(r % tw);
    const int s = (int)(r / tw);

    double pp = 0, mp = 0, pm = 0, mm = 0;   // I(+x,+y), I(-x,+y), I(+x,-y), I(-x,-y)
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
    const int tile = t / tcp, tl = t % tcp;
    float *o = F + ((((int64_t)p * ntile + tile) * sw + s) * 4) * tcp + tl;
    o[0] = (float)(Sp + Sm);
    o[tcp] = (float)(Sp - Sm);
    o[2 * tcp] = (float)(Dp + Dm);
    o[3 * tcp] = (float)(Dp - Dm);
}

int launch_fold(const double *img_dev, float *F, const DftGeom &g)
{
    int64_t total = (int64_t)g.nf * g.ntile * g.tcp * g.nchunk * DFT_RC;
    LaunchScope ls("fold_image");
    fold_image_kernel<<<ceil_div(total, 256), 256, 0, ctx().stream>>>(img_dev, F, g.ny, g.nx, g.nf, g.npx, g.npy,
                                                                      g.tcp, g.ntile, g.nchunk);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

// ---------------------------------------------------------------------------------
// DFT kernel.  grid = (uv tiles, column-tile splits, planes); 128 threads; UVT uv/thread.
// SMS: keep the per-thread fp64 state (phase rates fu, fv and the fp64 accumulators) in shared
// memory instead of registers; it is touched once per chunk, and the registers it frees let the
// wide variants hold their column tables without spilling.
template <int UVT, int TCP, bool F2, int MINB, bool SMS>
__global__ void __launch_bounds__(DFT_THREADS, MINB) dft_kernel(const DftParams P)
{
    constexpr int CHUNK_FLOATS = DFT_RC * 4 * TCP;
    constexpr uint32_t CHUNK_BYTES = CHUNK_FLOATS * sizeof(float);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stages = reinterpret_cast<float *>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[DFT_NSTAGE];

    const int tid = threadIdx.x;
    const int plane = blockIdx.z, sp = blockIdx.y;
    const int tile0 = (int)(((int64_t)sp * P.ntile) / P.nsplit);
    const int tile1 = (int)(((int64_t)(sp + 1) * P.ntile) / P.nsplit);
    const int ntl = tile1 - tile0;
    const int nit = ntl * P.nchunk;
    // everything this CTA reads is one contiguous run of nit chunks
    const float *gbase = P.F + ((size_t)plane * P.ntile + tile0) * (size_t)P.nchunk * CHUNK_FLOATS;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DFT_NSTAGE; s++) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DFT_NSTAGE; s++)
            if (s < nit) {
                mbar_expect_tx(&full_bar[s], CHUNK_BYTES);
                tma_load_1d(stages + s * CHUNK_FLOATS, gbase + (size_t)s * CHUNK_FLOATS, CHUNK_BYTES, &full_bar[s]);
            }
    }

    // per-thread uv points: phase advance per pixel in turns (fp64), row rotation (fp32)
    double fu_r[SMS ? 1 : UVT], fv_r[SMS ? 1 : UVT], Vr_r[SMS ? 1 : UVT], Vi_r[SMS ? 1 : UVT];
    double *fst = reinterpret_cast<double *>(smem_raw + (size_t)DFT_NSTAGE * CHUNK_BYTES);
#define FU(q) (SMS ? fst[(0 * UVT + (q)) * DFT_THREADS + tid] : fu_r[SMS ? 0 : (q)])
#define FV(q) (SMS ? fst[(1 * UVT + (q)) * DFT_THREADS + tid] : fv_r[SMS ? 0 : (q)])
#define VR(q) (SMS ? fst[(2 * UVT + (q)) * DFT_THREADS + tid] : Vr_r[SMS ? 0 : (q)])
#define VI(q) (SMS ? fst[(3 * UVT + (q)) * DFT_THREADS + tid] : Vi_r[SMS ? 0 : (q)])
    float Dr[UVT], Di[UVT];
#pragma unroll
    for (int q = 0; q < UVT; q++) {
        const int64_t k = (int64_t)blockIdx.x * (DFT_THREADS * UVT) + q * DFT_THREADS + tid;
        const bool valid = k < P.nuvh;
        const double fuq = valid ? P.u[k] * P.dxy : 0.0;
        const double fvq = valid ? P.v[k] * P.dxy : 0.0;
        FU(q) = fuq;
        FV(q) = fvq;
        double s, c;
        sincospi(2.0 * (fvq - rint(fvq)), &s, &c);
        Dr[q] = (float)c;
        Di[q] = (float)s;
        VR(q) = 0.0;
        VI(q) = 0.0;
    }

    int it = 0;
    for (int tl = 0; tl < ntl; tl++) {
        // ---- column factors of this tile: fp64 seed + fp64 rotation, rounded to fp32 ----
        float tc[UVT][TCP], ts[UVT][TCP];
#pragma unroll
        for (int q = 0; q < UVT; q++) {
            const double fuq = FU(q);
            double a0 = fuq * ((double)((tile0 + tl) * TCP) + P.hx);
            double cr, ci, rc, rs;
            sincospi(2.0 * (a0 - rint(a0)), &ci, &cr);
            sincospi(2.0 * (fuq - rint(fuq)), &rs, &rc);
#pragma unroll
            for (int t = 0; t < TCP; t++) {
                tc[q][t] = (float)cr;
                ts[q][t] = (float)ci;
                const double nr = cr * rc - ci * rs;
                ci = cr * rs + ci * rc;
                cr = nr;
            }
        }
        u64 tc2[UVT][TCP / 2], ts2[UVT][TCP / 2];
        if (F2) {
#pragma unroll
            for (int q = 0; q < UVT; q++)
#pragma unroll
                for (int t = 0; t < TCP / 2; t++) {
                    tc2[q][t] = pack2(tc[q][2 * t], tc[q][2 * t + 1]);
                    ts2[q][t] = pack2(ts[q][2 * t], ts[q][2 * t + 1]);
                }
        }

        for (int ch = 0; ch < P.nchunk; ch++, it++) {
            const int st = it % DFT_NSTAGE;
            const uint32_t parity = (uint32_t)((it / DFT_NSTAGE) & 1);
            // ---- row factor seed for the first row pair of the chunk (fp64 range reduction) ----
            float Er[UVT], Ei[UVT];
#pragma unroll
            for (int q = 0; q < UVT; q++) {
                double b0 = FV(q) * ((double)(ch * DFT_RC) + P.hy);
                b0 -= rint(b0);
                sincospif((float)(2.0 * b0), &Ei[q], &Er[q]);
            }
            mbar_wait(&full_bar[st], parity);
            const float *sm = stages + st * CHUNK_FLOATS;

            if (F2) {
                u64 vre[UVT], vim[UVT];
#pragma unroll
                for (int q = 0; q < UVT; q++) { vre[q] = 0ull; vim[q] = 0ull; }
#pragma unroll 2
                for (int r = 0; r < DFT_RC; r++) {
                    const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(sm + r * (4 * TCP));
                    u64 a1[UVT], a2[UVT], b1[UVT], b2[UVT];
#pragma unroll
                    for (int g = 0; g < TCP / 4; g++) {
                        const ulonglong2 ss = row[g], sd = row[TCP / 4 + g], ds = row[2 * (TCP / 4) + g],
                                         dd = row[3 * (TCP / 4) + g];
#pragma unroll
                        for (int q = 0; q < UVT; q++) {
                            if (g == 0) {
                                a1[q] = mul2(ss.x, tc2[q][0]);
                                a2[q] = mul2(sd.x, tc2[q][0]);
                                b1[q] = mul2(ds.x, ts2[q][0]);
                                b2[q] = mul2(dd.x, ts2[q][0]);
                            } else {
                                a1[q] = fma2(ss.x, tc2[q][2 * g], a1[q]);
                                a2[q] = fma2(sd.x, tc2[q][2 * g], a2[q]);
                                b1[q] = fma2(ds.x, ts2[q][2 * g], b1[q]);
                                b2[q] = fma2(dd.x, ts2[q][2 * g], b2[q]);
                            }
                            a1[q] = fma2(ss.y, tc2[q][2 * g + 1], a1[q]);
                            a2[q] = fma2(sd.y, tc2[q][2 * g + 1], a2[q]);
                            b1[q] = fma2(ds.y, ts2[q][2 * g + 1], b1[q]);
                            b2[q] = fma2(dd.y, ts2[q][2 * g + 1], b2[q]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < UVT; q++) {
                        const u64 er2 = pack2(Er[q], Er[q]), ei2 = pack2(Ei[q], Ei[q]),
                                  nei2 = pack2(-Ei[q], -Ei[q]);
                        vre[q] = fma2(er2, a1[q], vre[q]);
                        vre[q] = fma2(nei2, b2[q], vre[q]);
                        vim[q] = fma2(er2, b1[q], vim[q]);
                        vim[q] = fma2(ei2, a2[q], vim[q]);
                        const float nr = Er[q] * Dr[q] - Ei[q] * Di[q];
                        Ei[q] = Er[q] * Di[q] + Ei[q] * Dr[q];
                        Er[q] = nr;
                    }
                }
#pragma unroll
                for (int q = 0; q < UVT; q++) {
                    VR(q) += (double)sum2(vre[q]);
                    VI(q) += (double)sum2(vim[q]);
                }
            } else {
                float vre[UVT], vim[UVT];
#pragma unroll
                for (int q = 0; q < UVT; q++) { vre[q] = 0.f; vim[q] = 0.f; }
#pragma unroll 2
                for (int r = 0; r < DFT_RC; r++) {
                    const float4 *row = reinterpret_cast<const float4 *>(sm + r * (4 * TCP));
                    float a1[UVT], a2[UVT], b1[UVT], b2[UVT];
#pragma unroll
                    for (int g = 0; g < TCP / 4; g++) {
                        const float4 ss = row[g], sd = row[TCP / 4 + g], ds = row[2 * (TCP / 4) + g],
                                     dd = row[3 * (TCP / 4) + g];
#pragma unroll
                        for (int q = 0; q < UVT; q++) {
                            if (g == 0) {
                                a1[q] = ss.x * tc[q][0];
                                a2[q] = sd.x * tc[q][0];
                                b1[q] = ds.x * ts[q][0];
                                b2[q] = dd.x * ts[q][0];
                            } else {
                                a1[q] = fmaf(ss.x, tc[q][4 * g], a1[q]);
                                a2[q] = fmaf(sd.x, tc[q][4 * g], a2[q]);
                                b1[q] = fmaf(ds.x, ts[q][4 * g], b1[q]);
                                b2[q] = fmaf(dd.x, ts[q][4 * g], b2[q]);
                            }
                            a1[q] = fmaf(ss.y, tc[q][4 * g + 1], a1[q]);
                            a2[q] = fmaf(sd.y, tc[q][4 * g + 1], a2[q]);
                            b1[q] = fmaf(ds.y, ts[q][4 * g + 1], b1[q]);
                            b2[q] = fmaf(dd.y, ts[q][4 * g + 1], b2[q]);
                            a1[q] = fmaf(ss.z, tc[q][4 * g + 2], a1[q]);
                            a2[q] = fmaf(sd.z, tc[q][4 * g + 2], a2[q]);
                            b1[q] = fmaf(ds.z, ts[q][4 * g + 2], b1[q]);
                            b2[q] = fmaf(dd.z, ts[q][4 * g + 2], b2[q]);
                            a1[q] = fmaf(ss.w, tc[q][4 * g + 3], a1[q]);
                            a2[q] = fmaf(sd.w, tc[q][4 * g + 3], a2[q]);
                            b1[q] = fmaf(ds.w, ts[q][4 * g + 3], b1[q]);
                            b2[q] = fmaf(dd.w, ts[q][4 * g + 3], b2[q]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < UVT; q++) {
                        vre[q] = fmaf(Er[q], a1[q], vre[q]);
                        vre[q] = fmaf(-Ei[q], b2[q], vre[q]);
                        vim[q] = fmaf(Er[q], b1[q], vim[q]);
                        vim[q] = fmaf(Ei[q], a2[q], vim[q]);
                        const float nr = Er[q] * Dr[q] - Ei[q] * Di[q];
                        Ei[q] = Er[q] * Di[q] + Ei[q] * Dr[q];
                        Er[q] = nr;
                    }
                }
#pragma unroll
                for (int q = 0; q < UVT; q++) {
                    VR(q) += (double)vre[q];
                    VI(q) += (double)vim[q];
                }
            }

            // every thread is done reading this stage: refill it with chunk it+NSTAGE
            __syncthreads();
            if (tid == 0 && it + DFT_NSTAGE < nit) {
                fence_proxy_async();
                mbar_expect_tx(&full_bar[st], CHUNK_BYTES);
                tma_load_1d(stages + st * CHUNK_FLOATS, gbase + (size_t)(it + DFT_NSTAGE) * CHUNK_FLOATS,
                            CHUNK_BYTES, &full_bar[st]);
            }
        }
    }

#pragma unroll
    for (int q = 0; q < UVT; q++) {
        const int64_t k = (int64_t)blockIdx.x * (DFT_THREADS * UVT) + q * DFT_THREADS + tid;
        if (k < P.nuvh) P.part[((size_t)sp * P.nf + plane) * (size_t)P.nuvh + k] = make_double2(VR(q), VI(q));
    }
#undef FU
#undef FV
#undef VR
#undef VI
}

// ---------------------------------------------------------------------------------
struct VariantInfo {
    const char *name;
    int uvt, tcp, f2, minb, sms;
};
// Round 1 measured 22 tilings (2-6 uv points per thread, 16-32 column pairs, column- or uv-paired FFMA2, scalar
// FFMA, 1-4 accumulator chains; profiles/r01_dft_ncu.md): all within 74-77 % FMA-pipe activity, the widest that
// fits the register file was fastest.  Kept: that one, and the narrow tiling for images of <= 32 columns.
static const VariantInfo kVariants[] = {
    {"dft_f2_uv3_tc32_sms", 3, 32, 1, 2, 1},   // 1  default
    {"dft_f2_uv2_tc16", 2, 16, 1, 4, 0},       // 2  narrow images (column pairs <= 16)
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

int dft_variant_count() { return kNumVariants; }
int dft_pick_variant(int nx)
{
    const int v = ctx().dft_variant;
    if (v >= 1 && v <= kNumVariants) return v;
    return (nx + 1) / 2 <= 16 ? 2 : 1;
}
int dft_variant_tcp(int variant) { return kVariants[variant - 1].tcp; }

int dft_auto_split(int variant, int64_t nuvh, int nf, int ntile)
{
    if (ctx().dft_split > 0) return ctx().dft_split < ntile ? ctx().dft_split : ntile;
    const VariantInfo &vi = kVariants[variant - 1];
    int64_t uvtiles = (nuvh + (int64_t)DFT_THREADS * vi.uvt - 1) / ((int64_t)DFT_THREADS * vi.uvt);
    int64_t capacity = (int64_t)ctx().sm_count * vi.minb;
    int64_t base = uvtiles * nf;
    int64_t want = 20 * capacity;        // >= ~20 waves keeps the tail under ~5 %
    int64_t ns = (want + base - 1) / base;
    if (ns < 1) ns = 1;
    if (ns > ntile) ns = ntile;
    return (int)ns;
}

template <int UVT, int TCP, bool F2, int MINB, bool SMS>
static int launch_variant(const DftParams &p, const char *name)
{
    constexpr size_t smem = (size_t)DFT_NSTAGE * DFT_RC * 4 * TCP * sizeof(float) +
                            (SMS ? (size_t)4 * UVT * DFT_THREADS * sizeof(double) : 0);
    static bool attr_set = false;
    if (!attr_set) {
        PDSB_CUDA(cudaFuncSetAttribute(dft_kernel<UVT, TCP, F2, MINB, SMS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int64_t uvtiles = (p.nuvh + (int64_t)DFT_THREADS * UVT - 1) / ((int64_t)DFT_THREADS * UVT);
    if (uvtiles <= 0) return PDSB_OK;
    PDSB_REQUIRE(p.nsplit <= 65535 && p.nf <= 65535, "grid y/z dimensions");
    dim3 grid((unsigned)uvtiles, (unsigned)p.nsplit, (unsigned)p.nf);
    LaunchScope ls(name);
    dft_kernel<UVT, TCP, F2, MINB, SMS><<<grid, DFT_THREADS, smem, ctx().stream>>>(p);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

int launch_dft(const DftParams &p, int variant, int *tcp_of_variant)
{
    PDSB_REQUIRE(variant >= 1 && variant <= kNumVariants, "dft variant");
    if (tcp_of_variant) *tcp_of_variant = kVariants[variant - 1].tcp;
    const char *name = kVariants[variant - 1].name;
    switch (variant) {
        case 1: return launch_variant<3, 32, true, 2, true>(p, name);
        case 2: return launch_variant<2, 16, true, 4, false>(p, name);
    }
    return PDSB_ERR_ARG;
}

}  // namespace pdsb
