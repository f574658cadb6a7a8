// Direct Fourier sampling on the 5th-generation tensor cores (opt-in: pdsb_set_dft_variant(200),
// set_dft_kernel("tcgen05"), bench.py --dft tcgen05): tcgen05.mma with TMEM accumulators, operands staged by bulk TMA.
//
// Same algorithm as dft.cu: separable phases, mirror-folded real image, fp64-seeded phases, fp32 row-phase
// rotation in the epilogue, fp64 partial sums in the same [split][plane][uv] layout.  Both operands are
// split into two fp16 limbs (3 MMAs per product: hi.lo, lo.hi, hi.hi).
//
// Accumulator rounding.  tcgen05.mma aligns the 16 products of an instruction and its fp32 accumulator to
// the largest exponent and truncates toward zero (measured with pdsb_tc5_accum_probe: about 0.4 ulp per
// instruction from the accumulator plus 0.04 ulp per product).  With fp16-rounded limbs and 12
// instructions per accumulator round, sums of like-signed products (low spatial frequencies of a positive
// image) came out 3.4e-7 low - a bias, not noise.  The split used here removes it:
//   * the hi limbs live on a fixed-point LATTICE: trig factors on multiples of 2^-9, image values of one
//     (plane, K tile, chunk) on integers <= 2047 after a power-of-two scale chosen so that the largest row
//     sum of |hi| over the K tile stays below 2^15.  Every hi.hi product is a multiple of 2^-9 and every
//     partial sum of them is below 2^15: exactly representable in the accumulator, nothing to truncate;
//   * the cross products (2^-9 of the result) are issued FIRST, into the empty accumulator, where their
//     rounding is 2^-9 smaller; the hi.hi products follow.
// What remains is one truncation of the cross sum's low bits at the ulp of the result: measured -0.50 ulp
// = -4.3e-8 on like-signed sums (was -4.0 ulp), i.e. chi^2 per channel within 1e-7 of the fp64 kernel.
// The dropped lo.lo products are zero-mean and 2^-19 of (trig quantum x image quantum).
//
//   CTA = 128 uv points (UMMA M = 128), 10 warps with fixed roles:
//     warps 0-7  epilogue: thread (w & 3) * 32 + lane owns uv row r (TMEM lane r); warps 0-3 take row pairs
//                0..31 of every chunk, warps 4-7 row pairs 32..63.  They also write row r of the A operand
//                (cos/sin x hi/lo of the K tile's 64 columns, fp16) straight into TMEM with tcgen05.st,
//                double buffered over K tiles: the MMAs then read only B from shared memory, which is
//                what keeps the shared-memory data pipe (128 B/clk, shared with the TMA writes) off the
//                critical path.
//     warp 8     one thread issues the bulk TMA copies of the B units (6-stage ring of 16 KB)
//     warp 9     one elected thread issues the tcgen05.mma's and commits them to the mbarriers
//   B units ([hi|lo][N = 128 rows = 2 comps x 64 row pairs][K = 32], one trig type, one K sub-tile) are
//   written by the fold kernel directly in the canonical UMMA layout; one 16 KB 1-D bulk TMA copy feeds
//   6 MMAs (2 k-steps x 3 split products) of shape 128 x 128 x 16.
//   TMEM: 512 columns = [cos-type accumulator 128 | sin-type accumulator 128 | A buffer 0 128 | A buffer 1
//   128].  The two accumulators are the two pipeline buffers: while the epilogue warps read the cos-type
//   tile (SS, SD) the tensor core fills the sin-type tile (DS, -DD) and vice versa.  Everything is handed
//   over through mbarriers (stage full/empty, TMEM full/empty, A full); no CTA-wide barrier in steady state.
//   Epilogue arithmetic is packed fp32x2 (FFMA2) over adjacent row pairs, four independent phase chains.
//
// NOT the default: BASELINE.json's north star prescribes the FP32 pipe for the headline kernel.
#include "dft.cuh"
#include <cuda_fp16.h>
#include <algorithm>

namespace pdsb {

constexpr int T5_KSUB = 32;                     // column pairs per B sub-chunk (two UMMA k-steps of 16)
constexpr int T5_NSUB = 2;                      // sub-chunks accumulated in TMEM before the epilogue reads them
constexpr int T5_KT = T5_KSUB * T5_NSUB;        // column pairs per K tile (one set of A tiles)
constexpr int T5_RC = 64;                       // row pairs per chunk -> N = 128 per trig type
constexpr int T5_N = 2 * T5_RC;                 // UMMA N
constexpr int T5_M = 128;                       // UMMA M = uv points per CTA
constexpr int T5_KCHUNK_BYTES = (T5_N / 8) * 128;        // 2048: one 8-wide K slab of a [128 x K] tile
constexpr int T5_TILE_BYTES = (T5_KSUB / 8) * T5_KCHUNK_BYTES; // 8192: [128 x 32] fp16
constexpr int T5_UNIT_BYTES = 2 * T5_TILE_BYTES;         // B unit: [hi|lo] of one type, one K sub-tile (16 KB)
constexpr int T5_NSTAGE = 6;
constexpr int T5_ACOL = 2 * T5_N;                        // first TMEM column of the A buffers
constexpr int T5_APART = T5_KT / 2;                      // TMEM columns of one A part (two fp16 per column)
constexpr int T5_ABUF = 4 * T5_APART;                    // A buffer: [cos hi | cos lo | sin hi | sin lo]
static_assert(T5_ACOL + 2 * T5_ABUF <= 512, "TMEM budget");
constexpr uint32_t T5_IDESC = (1u << 4) | ((uint32_t)(T5_N >> 3) << 17) | ((uint32_t)(T5_M >> 4) << 24);
//                            D = f32      N                              M        (A, B = f16, K-major)

// canonical K-major, no-swizzle layout of a [128 rows x 32 k] fp16 tile: 8x8 core matrices of 128
// contiguous bytes; row groups 128 B apart (SBO), 8-wide K slabs 2048 B apart (LBO)
__host__ __device__ __forceinline__ int t5_off(int r, int k)
{
    return (k >> 3) * T5_KCHUNK_BYTES + (r >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2;
}

__device__ __forceinline__ uint32_t t5_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t t5_desc(uint32_t saddr)
{
    const uint32_t lo = ((saddr >> 4) & 0x3fffu) | (((uint32_t)(T5_KCHUNK_BYTES >> 4) & 0x3fffu) << 16);   // start, LBO
    const uint32_t hi = ((uint32_t)(128 >> 4) & 0x3fffu) | (1u << 14);                                     // SBO, version 1
    return ((uint64_t)hi << 32) | lo;                                                                      // layout: none
}
// low word of the descriptor of the tile at `saddr`; tiles at saddr + off have low word + (off >> 4)
__device__ __forceinline__ uint32_t t5_desc_lo(uint32_t saddr)
{
    return ((saddr >> 4) & 0x3fffu) | (((uint32_t)(T5_KCHUNK_BYTES >> 4) & 0x3fffu) << 16);
}
constexpr uint32_t T5_DESC_HI = ((uint32_t)(128 >> 4) & 0x3fffu) | (1u << 14);      // SBO, version 1, no swizzle
__device__ __forceinline__ bool t5_elect()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void t5_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t5_smem(bar)), "r"(count));
}
__device__ __forceinline__ void t5_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t5_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool t5_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(t5_smem(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void t5_wait(uint64_t *bar, uint32_t parity)
{
    while (!t5_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void t5_tma(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     t5_smem(dst)),
                 "l"(src), "r"(bytes), "r"(t5_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void t5_mma(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(adesc_lo), "r"(bdesc_lo), "r"(T5_IDESC), "r"(accumulate), "r"(T5_DESC_HI)
        : "memory");
}
// A operand from TMEM (lane = row, two fp16 per column along K), B from shared memory
__device__ __forceinline__ void t5_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t accumulate,
                                          uint32_t bdesc_hi = T5_DESC_HI)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(bdesc_lo), "r"(T5_IDESC), "r"(accumulate), "r"(bdesc_hi)
        : "memory");
}
__device__ __forceinline__ void t5_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ uint32_t t5_h2(__half a, __half b)
{
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ void t5_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t5_smem(bar)) : "memory");
}
__device__ __forceinline__ void t5_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void t5_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void t5_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void t5_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t5_smem(bar)) : "memory");
}
typedef unsigned long long t5_u64;
__device__ __forceinline__ t5_u64 t5_fma2(t5_u64 a, t5_u64 b, t5_u64 c)
{
    t5_u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ t5_u64 t5_mul2(t5_u64 a, t5_u64 b)
{
    t5_u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ t5_u64 t5_pack(float lo, float hi)
{
    t5_u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ t5_u64 t5_packu(uint32_t lo, uint32_t hi)
{
    t5_u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__device__ __forceinline__ t5_u64 t5_add2(t5_u64 a, t5_u64 b)
{
    t5_u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float t5_sum2(t5_u64 a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    return lo + hi;
}
// Phases are kept as 64-bit fixed-point fractions of a turn (2^64 = one turn): H = frac(f / 2) so that the
// phase of a pixel at half-integer offset n/2 is n * H (mod 2^64), exactly, in integer arithmetic.
// Nothing in the steady state touches fp64: a dense DFMA stream runs at 33.8 TFLOP/s on this GPU
// (pdsb_bench_fma variant 13), but the handful of fp64 instructions per accumulator round an earlier
// version had (two DADD + four DMUL/DFMA per thread) drew 35 % of all stall samples as
// math_pipe_throttle -- sparse use of the fp64 pipe is far more expensive than its peak rate suggests.
__device__ __forceinline__ unsigned long long t5_fix(double f)
{
    double h = 0.5 * f;
    h -= rint(h);                                                    // [-0.5, 0.5]
    return (unsigned long long)__double2ll_rn(h * 9223372036854775808.0) << 1;
}
__device__ __forceinline__ void t5_trig(unsigned long long H, int n, float &c, float &s)
{
    const unsigned long long ph = (unsigned long long)(long long)n * H;
    const float t = (float)(int)(unsigned)(ph >> 32) * 4.656612873077393e-10f;      // turns * 2 in [-1, 1)
    sincospif(t, &s, &c);
}
// (hi, lo) += x, error-free (Knuth two-sum)
__device__ __forceinline__ void t5_dfadd(float &hi, float &lo, float x)
{
    const float s = hi + x, bb = s - hi;
    lo += (hi - (s - bb)) + (x - bb);
    hi = s;
}
// trig factor -> lattice hi limb (multiple of 2^-9, exact in fp16) + fp16 remainder
__device__ __forceinline__ void t5_split(float x, __half &hi, __half &lo)
{
    const float h = rintf(x * 512.0f) * 0.001953125f;
    hi = __float2half_rn(h);
    lo = __float2half_rn(x - h);
}

// ---- fold into the canonical UMMA B layout ----
// B[plane][ktile][chunk][type][sub][hi|lo][t5_off(r, k)], r = c*64 + s_l (component c of the type: SS,SD |
// DS,-DD; row pair s_l of the chunk), k = column pair within the sub-chunk.
// Quad of mirror pixels -> the four parity components (SS, SD, DS, -DD) of column pair t, row pair s.
__device__ __forceinline__ void t5_quad(const double *__restrict__ img, int ny, int nx, int nf, int npx, int npy, int p, int t,
                                        int s, double (&comp)[4])
{
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
    comp[0] = Sp + Sm;
    comp[1] = Sp - Sm;
    comp[2] = Dp + Dm;
    comp[3] = -(Dp - Dm);
}

// Statistics of every accumulator round's B tiles (plane p, K tile kt, chunk): the largest |component| and the
// largest row sum of |component| over the K tile's 64 column pairs; stats[((p nkt + kt) nchunk + chunk) 2 + {0,1}]
// as the bit patterns of non-negative doubles (atomicMax on them orders like the values).  Thread = (plane,
// row pair, K tile), plane fastest: coalesced over the channel-fastest cube.
__global__ void __launch_bounds__(256) tc5_tile_stats_kernel(const double *__restrict__ img, unsigned long long *__restrict__ stats,
                                                             int ny, int nx, int nf, int npx, int npy, int nkt, int nchunk)
{
    const int64_t sw = (int64_t)nchunk * T5_RC;
    const int64_t total = (int64_t)nf * sw * nkt;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int p = (int)(idx % nf);
    const int64_t rr = idx / nf;
    const int s = (int)(rr % sw), kt = (int)(rr / sw);
    double l1[4] = {0, 0, 0, 0}, mx = 0.0;
    for (int k = 0; k < T5_KT; k++) {
        double comp[4];
        t5_quad(img, ny, nx, nf, npx, npy, p, kt * T5_KT + k, s, comp);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double a = fabs(comp[c]);
            if (a <= 1.79e308) {           // NaN / inf pixels do not set the scale (they poison the result anyway)
                l1[c] += a;
                mx = a > mx ? a : mx;
            }
        }
    }
    const double l1m = fmax(fmax(l1[0], l1[1]), fmax(l1[2], l1[3]));
    unsigned long long *o = stats + (((int64_t)p * nkt + kt) * nchunk + s / T5_RC) * 2;
    if (mx > 0.0) {
        atomicMax(o, (unsigned long long)__double_as_longlong(mx));
        atomicMax(o + 1, (unsigned long long)__double_as_longlong(l1m));
    }
}

constexpr double T5_HI_MAX = 2047.0;             // largest integer hi limb (exact in fp16)
constexpr double T5_L1_MAX = 32768.0 - 1024.0;   // largest row sum of |hi| over a K tile: hi.hi sums stay below 2^15
// Thread = plane.  Exponent e of every round's lattice quantum 2^e; the plane's unscale factor is the largest
// quantum, and qrel[round] = quantum / that (a float power of two <= 1; clamped at 2^-100, where the tile no
// longer matters against the plane).  inv_q[round] = 1 / quantum for the fold.
__global__ void tc5_tile_scale_kernel(const unsigned long long *__restrict__ stats, int nf, int nround, float *__restrict__ qrel,
                                      double *__restrict__ inv_q, double *__restrict__ plane_unscale)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nf) return;
    int emax = -100000;
    for (int r = 0; r < nround; r++) {
        const double mx = __longlong_as_double((long long)stats[((int64_t)p * nround + r) * 2]);
        const double l1 = __longlong_as_double((long long)stats[((int64_t)p * nround + r) * 2 + 1]);
        if (mx > 0.0) {
            const int e = (int)ceil(log2(fmax(mx / T5_HI_MAX, l1 / T5_L1_MAX)));
            emax = e > emax ? e : emax;
        }
    }
    if (emax == -100000) emax = 0;                // all-zero plane
    plane_unscale[p] = exp2((double)emax);
    for (int r = 0; r < nround; r++) {
        const double mx = __longlong_as_double((long long)stats[((int64_t)p * nround + r) * 2]);
        const double l1 = __longlong_as_double((long long)stats[((int64_t)p * nround + r) * 2 + 1]);
        int e = emax;
        if (mx > 0.0) e = (int)ceil(log2(fmax(mx / T5_HI_MAX, l1 / T5_L1_MAX)));
        if (e < emax - 100) e = emax - 100;
        qrel[(int64_t)p * nround + r] = exp2f((float)(e - emax));
        inv_q[(int64_t)p * nround + r] = exp2((double)(-e));
    }
}

__global__ void __launch_bounds__(256) fold_tc5_kernel(const double *__restrict__ img, unsigned char *__restrict__ B,
                                                       const double *__restrict__ inv_q, int ny, int nx, int nf,
                                                       int npx, int npy, int nkt, int nchunk)
{
    const int64_t tw = (int64_t)nkt * T5_KT, sw = (int64_t)nchunk * T5_RC;
    const int64_t total = (int64_t)nf * tw * sw;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int p = (int)(idx % nf);
    const int64_t rr = idx / nf;
    const int t = (int)(rr % tw), s = (int)(rr / tw);
    double comp[4];                                                      // SS, SD, DS, -DD
    t5_quad(img, ny, nx, nf, npx, npy, p, t, s, comp);
    const int kt = t / T5_KT, sub = (t % T5_KT) / T5_KSUB, k = t % T5_KSUB;
    const int chunk = s / T5_RC, s_l = s % T5_RC;
    const double sc = inv_q[((int64_t)p * nkt + kt) * nchunk + chunk];  // 1 / lattice quantum of this round
    unsigned char *base = B + (((int64_t)p * nkt + kt) * nchunk + chunk) * (int64_t)(2 * T5_NSUB * T5_UNIT_BYTES);
#pragma unroll
    for (int cidx = 0; cidx < 4; cidx++) {
        const int type = cidx >> 1, c = cidx & 1;
        const int off = t5_off(c * T5_RC + s_l, k);
        const double x = comp[cidx] * sc;
        const double xh = rint(x);                                       // integer hi limb, |xh| <= 2047
        const __half hi = __double2half(xh);
        const __half lo = __double2half(x - xh);
        unsigned char *unit = base + (type * T5_NSUB + sub) * T5_UNIT_BYTES;
        *reinterpret_cast<__half *>(unit + off) = hi;
        *reinterpret_cast<__half *>(unit + T5_TILE_BYTES + off) = lo;
    }
}

// ---- the kernel ----
constexpr int T5_EPI_THREADS = 256;             // warps 0-7
constexpr int T5_THREADS = T5_EPI_THREADS + 64; // + TMA warp + MMA warp
// ORDER 1 (the product): per accumulator round all cross products first, then the hi.hi products (see the header).
// ORDER 0: the three products of a k-step together (round 1's order; kept as variant 201 to time the difference -
// it is NOT bias-free: every later instruction truncates an accumulator that already has full magnitude).
// Round-1 timing probes (C3/4, 250k uv): full kernel 7.66 ms, epilogue removed 7.22 ms, TMA removed as well
// 7.23 ms -> the kernel is bound by the tcgen05.mma rate itself (~82 cycles per 128x128x16 TS-mode MMA).
template <int ORDER>
__global__ void __launch_bounds__(T5_THREADS, 1) dft_tc5_kernel(const DftParams P, const unsigned char *__restrict__ Bg,
                                                                const float *__restrict__ qrel, int nkt, int pg)
{
    extern __shared__ __align__(1024) unsigned char Bs[];       // T5_NSTAGE x 16 KB
    __shared__ __align__(8) uint64_t full_bar[T5_NSTAGE];       // TMA landed           (producer -> MMA)
    __shared__ __align__(8) uint64_t empty_bar[T5_NSTAGE];      // MMAs read the stage  (MMA commit -> producer)
    __shared__ __align__(8) uint64_t tmem_full[2];              // accumulator ready    (MMA commit -> epilogue)
    __shared__ __align__(8) uint64_t tmem_empty[2];             // accumulator read     (epilogue -> MMA)
    __shared__ __align__(8) uint64_t a_full[2];                 // A buffer written     (epilogue -> MMA)
    __shared__ uint32_t tmem_base_s;
    __shared__ float4 vx[T5_M];                  // pass-end exchange between the two halves of a uv point
    __shared__ float2 ctab[T5_KT / 2][T5_M];     // (cos, sin) of j column steps of uv row r: ctab[j][r], 32 KB

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int plane0 = blockIdx.z * pg, sp = blockIdx.y;
    const int npl = (P.nf - plane0) < pg ? (P.nf - plane0) : pg;
    const int kt0 = (int)(((int64_t)sp * nkt) / P.nsplit), kt1 = (int)(((int64_t)(sp + 1) * nkt) / P.nsplit);
    const int nkl = kt1 - kt0;
    const int per_k = npl * P.nchunk;            // accumulator rounds per K tile
    const int nit = nkl * per_k;

    if (tid == 0) {
        for (int s = 0; s < T5_NSTAGE; s++) {
            t5_mbar_init(&full_bar[s], 1);
            t5_mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            t5_mbar_init(&tmem_full[b], 1);
            t5_mbar_init(&tmem_empty[b], T5_EPI_THREADS);
            t5_mbar_init(&a_full[b], T5_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {           // one warp allocates all 512 TMEM columns (this CTA is alone on its SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t5_smem(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 8) {
        // ================= TMA producer =================
        if (lane == 0) {
            // the units of one (K tile, plane) pass are contiguous: [chunk][type][sub]
            int kl = 0, pl = 0, cs = 0;
            const int ncs = P.nchunk * 2 * T5_NSUB;
            const int nunit = nit * 2 * T5_NSUB;
            for (int it = 0; it < nunit; it++) {
                const int st = it % T5_NSTAGE;
                if (it >= T5_NSTAGE) t5_wait(&empty_bar[st], (uint32_t)((it / T5_NSTAGE - 1) & 1));
                const unsigned char *src =
                    Bg + (((size_t)(plane0 + pl) * nkt + (kt0 + kl)) * (size_t)ncs + cs) * T5_UNIT_BYTES;
                t5_expect_tx(&full_bar[st], T5_UNIT_BYTES);
                t5_tma(Bs + (size_t)st * T5_UNIT_BYTES, src, T5_UNIT_BYTES, &full_bar[st]);
                if (++cs == ncs) {
                    cs = 0;
                    if (++pl == npl) {
                        pl = 0;
                        kl++;
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================
        // the whole warp walks the loops and waits (warp-uniform control flow); one elected lane issues
        const bool leader = t5_elect();
        const uint32_t b_lo0 = t5_desc_lo(t5_smem(Bs));
        int it = 0, is = 0, st = 0;                  // accumulator rounds, units, ring stage
        for (int kl = 0; kl < nkl; kl++) {
            t5_wait(&a_full[kl & 1], (uint32_t)((kl >> 1) & 1));
            const uint32_t a_buf = tmem_base + (uint32_t)(T5_ACOL + (kl & 1) * T5_ABUF);
            for (int r = 0; r < per_k; r++, it++) {
#pragma unroll
                for (int ty = 0; ty < 2; ty++) {     // accumulator ty: cos-type, sin-type
                    const uint32_t d = tmem_base + (uint32_t)(ty * T5_N);
                    if (it >= 1) t5_wait(&tmem_empty[ty], (uint32_t)((it - 1) & 1));
                    // A part (ty, hi|lo) at column (ty*2 + part) * APART; 16 k = 8 columns
                    const uint32_t a_hi0 = a_buf + (uint32_t)((ty * 2 + 0) * T5_APART);
                    const uint32_t a_lo0 = a_buf + (uint32_t)((ty * 2 + 1) * T5_APART);
                    if (ORDER == 0) {
#pragma unroll
                        for (int sub = 0; sub < T5_NSUB; sub++, is++) {
                            t5_wait(&full_bar[st], (uint32_t)((is / T5_NSTAGE) & 1));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            if (leader) {
                                const uint32_t b_s = b_lo0 + (uint32_t)st * (uint32_t)(T5_UNIT_BYTES >> 4);
#pragma unroll
                                for (int ks = 0; ks < 2; ks++) {
                                    const uint32_t ahi = a_hi0 + (uint32_t)(sub * (T5_KSUB / 2) + ks * 8);
                                    const uint32_t alo = a_lo0 + (uint32_t)(sub * (T5_KSUB / 2) + ks * 8);
                                    const uint32_t bhi = b_s + (uint32_t)((ks * 2 * T5_KCHUNK_BYTES) >> 4);
                                    const uint32_t blo = b_s + (uint32_t)((T5_TILE_BYTES + ks * 2 * T5_KCHUNK_BYTES) >> 4);
                                    t5_mma_ts(d, ahi, bhi, (sub | ks) ? 1u : 0u);
                                    t5_mma_ts(d, ahi, blo, 1u);
                                    t5_mma_ts(d, alo, bhi, 1u);
                                }
                                t5_commit(&empty_bar[st]);   // stage (and, at the end of a K tile, the A buffer) consumed
                                if (sub == T5_NSUB - 1) t5_commit(&tmem_full[ty]);
                            }
                            __syncwarp();
                            if (++st == T5_NSTAGE) st = 0;
                        }
                    } else {
                        // both units of the round stay until its hi.hi products have been issued
                        static_assert(T5_NSUB == 2 && T5_NSTAGE % T5_NSUB == 0, "a round's units are adjacent ring stages");
                        t5_wait(&full_bar[st], (uint32_t)((is / T5_NSTAGE) & 1));
                        t5_wait(&full_bar[st + 1], (uint32_t)(((is + 1) / T5_NSTAGE) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (leader) {
#pragma unroll
                            for (int sub = 0; sub < T5_NSUB; sub++) {          // cross products into the empty accumulator
                                const uint32_t b_s = b_lo0 + (uint32_t)(st + sub) * (uint32_t)(T5_UNIT_BYTES >> 4);
#pragma unroll
                                for (int ks = 0; ks < 2; ks++) {
                                    const uint32_t ahi = a_hi0 + (uint32_t)(sub * (T5_KSUB / 2) + ks * 8);
                                    const uint32_t alo = a_lo0 + (uint32_t)(sub * (T5_KSUB / 2) + ks * 8);
                                    const uint32_t bhi = b_s + (uint32_t)((ks * 2 * T5_KCHUNK_BYTES) >> 4);
                                    const uint32_t blo = b_s + (uint32_t)((T5_TILE_BYTES + ks * 2 * T5_KCHUNK_BYTES) >> 4);
                                    t5_mma_ts(d, ahi, blo, (sub | ks) ? 1u : 0u);
                                    t5_mma_ts(d, alo, bhi, 1u);
                                }
                            }
#pragma unroll
                            for (int sub = 0; sub < T5_NSUB; sub++) {          // lattice products: exact sums
                                const uint32_t b_s = b_lo0 + (uint32_t)(st + sub) * (uint32_t)(T5_UNIT_BYTES >> 4);
#pragma unroll
                                for (int ks = 0; ks < 2; ks++) {
                                    const uint32_t ahi = a_hi0 + (uint32_t)(sub * (T5_KSUB / 2) + ks * 8);
                                    const uint32_t bhi = b_s + (uint32_t)((ks * 2 * T5_KCHUNK_BYTES) >> 4);
                                    t5_mma_ts(d, ahi, bhi, 1u);
                                }
                                t5_commit(&empty_bar[st + sub]);   // stage (and, at the end of a K tile, the A buffer) consumed
                            }
                            t5_commit(&tmem_full[ty]);
                        }
                        __syncwarp();
                        is += T5_NSUB;
                        st += T5_NSUB;
                        if (st == T5_NSTAGE) st = 0;
                    }
                }
            }
        }
    } else {
        // ================= epilogue warps =================
        const int row = tid & (T5_M - 1), half_id = tid >> 7;      // uv row (TMEM lane), which half of the work
        const int64_t kuv = (int64_t)blockIdx.x * T5_M + row;
        const bool valid = kuv < P.nuvh;
        const unsigned long long Hu = t5_fix(valid ? P.u[kuv] * P.dxy : 0.0), Hv = t5_fix(valid ? P.v[kuv] * P.dxy : 0.0);
        const int hx2 = P.hx != 0.0 ? 1 : 0, hy2 = P.hy != 0.0 ? 1 : 0;    // pixel offsets in half pixels
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        float D1r, D1i;                                  // row-phase step of one row pair
        t5_u64 D2r, D2i, D2n, D8r, D8i, D8n;             // packed (x, x): steps of two and of eight row pairs
        {
            float s, c;
            t5_trig(Hv, 2, D1r, D1i);
            t5_trig(Hv, 4, c, s);
            D2r = t5_pack(c, c);
            D2i = t5_pack(s, s);
            D2n = t5_pack(-s, -s);
            t5_trig(Hv, 16, c, s);
            D8r = t5_pack(c, c);
            D8i = t5_pack(s, s);
            D8n = t5_pack(-s, -s);
        }

        // A operand of K tile kl: row `row` = trig of this uv point at the tile's 64 columns, fp16 hi + lo,
        // written to TMEM lane `row`; this thread does half of the columns, each from its exact phase
        // The column phase of this thread's first column is exact (fixed point -> sincospif); the other 31
        // come from it by the angle-addition formulas with the per-uv step table ctab (one sincospif per K
        // tile and thread instead of 32; 1.2e-7 instead of 6e-8 on a factor that is split to 22 bits anyway).
        {
            constexpr int HALF_K = T5_KT / 2;
#pragma unroll 1
            for (int j = half_id * (HALF_K / 2); j < (half_id + 1) * (HALF_K / 2); j++) {
                float cj, sj;
                t5_trig(Hu, 2 * j, cj, sj);
                ctab[j][row] = make_float2(cj, sj);
            }
            asm volatile("bar.sync 5, %0;" ::"n"(T5_EPI_THREADS) : "memory");        // epilogue warps only
        }
        auto gen_a = [&](int kl) {
            constexpr int HALF_K = T5_KT / 2;
            const uint32_t ab = lane_base + (uint32_t)(T5_ACOL + (kl & 1) * T5_ABUF + half_id * (HALF_K / 2));
            const int n0 = 2 * ((kt0 + kl) * T5_KT + half_id * HALF_K) + hx2;
            float c0, s0;
            t5_trig(Hu, n0, c0, s0);
#pragma unroll 1
            for (int kc = 0; kc < HALF_K / 8; kc++) {
                __half ch[8], cl[8], sh[8], sl[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const float2 d = ctab[kc * 8 + e][row];
                    const float cf = c0 * d.x - s0 * d.y, sf = s0 * d.x + c0 * d.y;
                    t5_split(cf, ch[e], cl[e]);
                    t5_split(sf, sh[e], sl[e]);
                }
                const uint32_t col = ab + (uint32_t)(kc * 4);
                t5_st4(col + 0 * T5_APART, t5_h2(ch[0], ch[1]), t5_h2(ch[2], ch[3]), t5_h2(ch[4], ch[5]), t5_h2(ch[6], ch[7]));
                t5_st4(col + 1 * T5_APART, t5_h2(cl[0], cl[1]), t5_h2(cl[2], cl[3]), t5_h2(cl[4], cl[5]), t5_h2(cl[6], cl[7]));
                t5_st4(col + 2 * T5_APART, t5_h2(sh[0], sh[1]), t5_h2(sh[2], sh[3]), t5_h2(sh[4], sh[5]), t5_h2(sh[6], sh[7]));
                t5_st4(col + 3 * T5_APART, t5_h2(sl[0], sl[1]), t5_h2(sl[2], sl[3]), t5_h2(sl[4], sl[5]), t5_h2(sl[6], sl[7]));
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            t5_arrive(&a_full[kl & 1]);
        };
        if (nkl > 0) gen_a(0);
        if (nkl > 1) gen_a(1);

        float Vrh = 0.f, Vrl = 0.f, Vih = 0.f, Vil = 0.f;    // two-float sums of the current (K tile, plane) pass
        int kl = 0, pl = 0, ch = 0;
        for (int e = 0; e < nit; e++) {
            // packed phase chains c = 0..3 holding rows (2c, 2c+1) of every group of eight row pairs, seeded
            // from the exact phase of this thread's first row pair of the chunk
            t5_u64 Sr[4], Si[4];
            {
                float e0r, e0i;
                t5_trig(Hv, 2 * (ch * T5_RC + half_id * 32) + hy2, e0r, e0i);
                // the round's lattice quantum (relative to the plane's) rides on the row phases
                const float qr = __ldg(qrel + ((size_t)(plane0 + pl) * nkt + (kt0 + kl)) * P.nchunk + ch);
                e0r *= qr;
                e0i *= qr;
                const float e1r = e0r * D1r - e0i * D1i, e1i = e0r * D1i + e0i * D1r;
                Sr[0] = t5_pack(e0r, e1r);
                Si[0] = t5_pack(e0i, e1i);
#pragma unroll
                for (int c = 1; c < 4; c++) {
                    Sr[c] = t5_fma2(Si[c - 1], D2n, t5_mul2(Sr[c - 1], D2r));
                    Si[c] = t5_fma2(Si[c - 1], D2r, t5_mul2(Sr[c - 1], D2i));
                }
            }
#pragma unroll
            for (int ty = 0; ty < 2; ty++) {
                // tile ty holds components (X | Y) = (SS | SD) or (DS | -DD):  p += cos(b) X,  q += sin(b) Y
                t5_u64 Er[4], Ei[4], aP[4] = {0, 0, 0, 0}, aQ[4] = {0, 0, 0, 0};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    Er[c] = Sr[c];
                    Ei[c] = Si[c];
                }
                t5_wait(&tmem_full[ty], (uint32_t)(e & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tb = lane_base + (uint32_t)(ty * T5_N + half_id * 32);
                // the whole tile goes to registers first so that the accumulator is handed back to the tensor
                // core after the TMEM load latency only, not after this thread's arithmetic
                uint32_t xv[32], yv[32];
                t5_ld32(tb, xv);
                t5_ld32(tb + (uint32_t)T5_RC, yv);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                t5_arrive(&tmem_empty[ty]);
#pragma unroll
                for (int g = 0; g < 4; g++) {                                     // eight row pairs each
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        aP[c] = t5_fma2(Er[c], t5_packu(xv[g * 8 + 2 * c], xv[g * 8 + 2 * c + 1]), aP[c]);
                        aQ[c] = t5_fma2(Ei[c], t5_packu(yv[g * 8 + 2 * c], yv[g * 8 + 2 * c + 1]), aQ[c]);
                        if (g < 3) {                                              // advance the chain by eight row pairs
                            const t5_u64 nr = t5_fma2(Ei[c], D8n, t5_mul2(Er[c], D8r));
                            Ei[c] = t5_fma2(Ei[c], D8r, t5_mul2(Er[c], D8i));
                            Er[c] = nr;
                        }
                    }
                }
                const float sp_ = t5_sum2(t5_add2(t5_add2(aP[0], aP[1]), t5_add2(aP[2], aP[3])));
                const float sq_ = t5_sum2(t5_add2(t5_add2(aQ[0], aQ[1]), t5_add2(aQ[2], aQ[3])));
                if (ty == 0) {
                    t5_dfadd(Vrh, Vrl, sp_);
                    t5_dfadd(Vih, Vil, sq_);
                } else {
                    t5_dfadd(Vih, Vil, sp_);
                    t5_dfadd(Vrh, Vrl, sq_);
                }
            }
            if (++ch == P.nchunk) {                  // end of this (K tile, plane) pass: combine the two halves
                ch = 0;
                const int bar_id = 1 + (warp & 3);   // named barrier of the warp pair (w, w + 4)
                if (half_id == 1) vx[row] = make_float4(Vrh, Vrl, Vih, Vil);
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                if (half_id == 0 && valid) {
                    const float4 o2 = vx[row];
                    const double vr = ((double)Vrh + (double)o2.x) + ((double)Vrl + (double)o2.y);
                    const double vi = ((double)Vih + (double)o2.z) + ((double)Vil + (double)o2.w);
                    double2 *dst = P.part + ((size_t)sp * P.nf + (plane0 + pl)) * (size_t)P.nuvh + kuv;
                    if (kl == 0) *dst = make_double2(vr, vi);
                    else {
                        const double2 o = *dst;
                        *dst = make_double2(o.x + vr, o.y + vi);
                    }
                }
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                Vrh = Vrl = Vih = Vil = 0.f;
                if (++pl == npl) {                   // last chunk of K tile kl: its A buffer is free again
                    pl = 0;
                    if (kl + 2 < nkl) gen_a(kl + 2);
                    kl++;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// ---- host side ----
size_t tc5_operand_bytes(int ny, int nx, int nf)
{
    const int npx = (nx + 1) / 2, npy = (ny + 1) / 2;
    const int nkt = (npx + T5_KT - 1) / T5_KT, nchunk = (npy + T5_RC - 1) / T5_RC;
    return (size_t)nf * nkt * nchunk * 2 * T5_NSUB * T5_UNIT_BYTES;
}

static int tc5_pg(int nf) { return nf < 8 ? nf : 8; }

int tc5_auto_split(int64_t nuvh, int nf, int nx)
{
    Context &c = ctx();
    const int nkt = ((nx + 1) / 2 + T5_KT - 1) / T5_KT;
    if (c.dft_split > 0) return c.dft_split < nkt ? c.dft_split : nkt;
    const int64_t uvtiles = (nuvh + T5_M - 1) / T5_M;
    const int pgroups = (nf + tc5_pg(nf) - 1) / tc5_pg(nf);
    const int64_t capacity = (int64_t)c.sm_count, base = std::max<int64_t>(1, uvtiles * pgroups);
    const int64_t ns = (20 * capacity + base - 1) / base;
    return (int)(ns < 1 ? 1 : (ns > nkt ? nkt : ns));
}

// Workspace of one fold: [stats: 2 u64 per round | inv_q: double per round | plane_unscale: double per plane |
// qrel: float per round], rounds = nf * nkt * nchunk.
size_t tc5_ws_bytes(int ny, int nx, int nf)
{
    const int npx = (nx + 1) / 2, npy = (ny + 1) / 2;
    const size_t nround = (size_t)nf * ((npx + T5_KT - 1) / T5_KT) * ((npy + T5_RC - 1) / T5_RC);
    return nround * (2 * sizeof(unsigned long long) + sizeof(double) + sizeof(float)) + (size_t)nf * sizeof(double) + 64;
}
struct Tc5Ws {
    unsigned long long *stats;
    double *inv_q, *plane_unscale;
    float *qrel;
};
static Tc5Ws tc5_ws(void *ws, int ny, int nx, int nf)
{
    const int npx = (nx + 1) / 2, npy = (ny + 1) / 2;
    const size_t nround = (size_t)nf * ((npx + T5_KT - 1) / T5_KT) * ((npy + T5_RC - 1) / T5_RC);
    Tc5Ws w;
    w.stats = reinterpret_cast<unsigned long long *>(ws);
    w.inv_q = reinterpret_cast<double *>(w.stats + 2 * nround);
    w.plane_unscale = w.inv_q + nround;
    w.qrel = reinterpret_cast<float *>(w.plane_unscale + nf);
    return w;
}
const double *tc5_plane_unscale(void *ws, int ny, int nx, int nf) { return tc5_ws(ws, ny, nx, nf).plane_unscale; }

int launch_fold_tc5(const double *img_dev, unsigned char *B, void *ws, int ny, int nx, int nf)
{
    Context &c = ctx();
    const int npx = (nx + 1) / 2, npy = (ny + 1) / 2;
    const int nkt = (npx + T5_KT - 1) / T5_KT, nchunk = (npy + T5_RC - 1) / T5_RC;
    const Tc5Ws w = tc5_ws(ws, ny, nx, nf);
    const int nround = nkt * nchunk;
    PDSB_CUDA(cudaMemsetAsync(w.stats, 0, (size_t)nf * nround * 2 * sizeof(unsigned long long), c.stream));
    {
        LaunchScope ls("tc5_tile_stats");
        const int64_t total = (int64_t)nf * nchunk * T5_RC * nkt;
        tc5_tile_stats_kernel<<<ceil_div(total, 256), 256, 0, c.stream>>>(img_dev, w.stats, ny, nx, nf, npx, npy, nkt, nchunk);
        tc5_tile_scale_kernel<<<ceil_div(nf, 64), 64, 0, c.stream>>>(w.stats, nf, nround, w.qrel, w.inv_q, w.plane_unscale);
        PDSB_CUDA(cudaGetLastError());
    }
    const int64_t total = (int64_t)nf * nkt * T5_KT * nchunk * T5_RC;
    LaunchScope ls("fold_tc5");
    fold_tc5_kernel<<<ceil_div(total, 256), 256, 0, c.stream>>>(img_dev, B, w.inv_q, ny, nx, nf, npx, npy, nkt, nchunk);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

int launch_dft_tc5(DftParams p, const unsigned char *B, void *ws, int ny, int nx)
{
    Context &c = ctx();
    const int npx = (nx + 1) / 2, npy = (ny + 1) / 2;
    const int nkt = (npx + T5_KT - 1) / T5_KT;
    p.nchunk = (npy + T5_RC - 1) / T5_RC;
    PDSB_REQUIRE(p.nsplit >= 1 && p.nsplit <= nkt, "tc5 split");
    constexpr size_t smem_bytes = (size_t)T5_NSTAGE * T5_UNIT_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        PDSB_CUDA(cudaFuncSetAttribute(dft_tc5_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        PDSB_CUDA(cudaFuncSetAttribute(dft_tc5_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        attr_set = true;
    }
    const int64_t uvtiles = (p.nuvh + T5_M - 1) / T5_M;
    if (uvtiles <= 0) return PDSB_OK;
    const int pg = tc5_pg(p.nf);
    const float *qrel = tc5_ws(ws, ny, nx, p.nf).qrel;
    dim3 grid((unsigned)uvtiles, (unsigned)p.nsplit, (unsigned)((p.nf + pg - 1) / pg));
    LaunchScope ls("dft_tc5_tcgen05");
    if (c.dft_variant == DFT_VARIANT_TC5 + 1) dft_tc5_kernel<0><<<grid, T5_THREADS, smem_bytes, c.stream>>>(p, B, qrel, nkt, pg);
    else dft_tc5_kernel<1><<<grid, T5_THREADS, smem_bytes, c.stream>>>(p, B, qrel, nkt, pg);
    PDSB_CUDA(cudaGetLastError());
    return PDSB_OK;
}

// ---- accumulation probe --------------------------------------------------------------------------
// One CTA, one accumulator round of the product kernel's MMA sequence (K = 64: 2 sub-tiles x 2 k-steps x
// [hi.hi, hi.lo, lo.hi]) on caller-supplied fp16 operands, with the accumulator pre-loaded with `c0`.
// It answers how the tensor core rounds its fp32 accumulator (the host compares against the exact
// fp64 sum of the same fp16 products): tests/test_gpu_tc5_accum.py, scripts/gpu_tc5_accum.py.
//   order 0: the product kernel's issue order;  1: all cross products first, the hi.hi products last.
__global__ void __launch_bounds__(128, 1) tc5_probe_kernel(const __half *__restrict__ Ahi, const __half *__restrict__ Alo,
                                                           const __half *__restrict__ Bhi, const __half *__restrict__ Blo,
                                                           float c0, int order, float *__restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char Bs[];       // 2 units of 16 KB
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        t5_mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t5_smem(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    // B row `tid` of both sub-tiles, canonical layout
    for (int k = 0; k < T5_KT; k++) {
        const int sub = k / T5_KSUB, off = t5_off(tid, k % T5_KSUB);
        *reinterpret_cast<__half *>(Bs + sub * T5_UNIT_BYTES + off) = Bhi[tid * T5_KT + k];
        *reinterpret_cast<__half *>(Bs + sub * T5_UNIT_BYTES + T5_TILE_BYTES + off) = Blo[tid * T5_KT + k];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // A row `tid` -> TMEM lane tid: parts [hi | lo], two fp16 per column
    for (int kc = 0; kc < T5_KT / 8; kc++) {
        const __half *h = Ahi + tid * T5_KT + kc * 8, *l = Alo + tid * T5_KT + kc * 8;
        const uint32_t col = lane_base + (uint32_t)(T5_ACOL + kc * 4);
        t5_st4(col, t5_h2(h[0], h[1]), t5_h2(h[2], h[3]), t5_h2(h[4], h[5]), t5_h2(h[6], h[7]));
        t5_st4(col + T5_APART, t5_h2(l[0], l[1]), t5_h2(l[2], l[3]), t5_h2(l[4], l[5]), t5_h2(l[6], l[7]));
    }
    const uint32_t cb = __float_as_uint(c0);
    for (int c4 = 0; c4 < T5_N / 4; c4++) t5_st4(lane_base + (uint32_t)(c4 * 4), cb, cb, cb, cb);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t b_lo0 = t5_desc_lo(t5_smem(Bs));
        const uint32_t a_buf = tmem_base + (uint32_t)T5_ACOL;
        for (int pass = 0; pass < (order == 1 ? 2 : 1); pass++)
            for (int sub = 0; sub < T5_NSUB; sub++)
                for (int ks = 0; ks < 2; ks++) {
                    const uint32_t ahi = a_buf + (uint32_t)(sub * (T5_KSUB / 2) + ks * 8);
                    const uint32_t alo = ahi + (uint32_t)T5_APART;
                    const uint32_t b_s = b_lo0 + (uint32_t)sub * (uint32_t)(T5_UNIT_BYTES >> 4);
                    const uint32_t bhi = b_s + (uint32_t)((ks * 2 * T5_KCHUNK_BYTES) >> 4);
                    const uint32_t blo = b_s + (uint32_t)((T5_TILE_BYTES + ks * 2 * T5_KCHUNK_BYTES) >> 4);
                    if (order == 0) {
                        t5_mma_ts(tmem_base, ahi, bhi, 1u);
                        t5_mma_ts(tmem_base, ahi, blo, 1u);
                        t5_mma_ts(tmem_base, alo, bhi, 1u);
                    } else if (pass == 0) {
                        t5_mma_ts(tmem_base, ahi, blo, 1u);
                        t5_mma_ts(tmem_base, alo, bhi, 1u);
                    } else t5_mma_ts(tmem_base, ahi, bhi, 1u);
                }
        t5_commit(&done_bar);
    }
    t5_wait(&done_bar, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c32 = 0; c32 < T5_N / 32; c32++) {
        uint32_t v[32];
        t5_ld32(lane_base + (uint32_t)(c32 * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 32; e++) out[tid * T5_N + c32 * 32 + e] = __uint_as_float(v[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

int tc5_probe(const uint16_t *a_hi, const uint16_t *a_lo, const uint16_t *b_hi, const uint16_t *b_lo, float c0, int order,
              float *out)
{
    Context &c = ctx();
    const size_t opb = (size_t)T5_M * T5_KT * sizeof(__half), outb = (size_t)T5_M * T5_N * sizeof(float);
    PDSB_CHECK(c.stage_a.ensure(4 * opb + outb));
    unsigned char *d = c.stage_a.as<unsigned char>();
    const uint16_t *src[4] = {a_hi, a_lo, b_hi, b_lo};
    for (int i = 0; i < 4; i++) PDSB_CUDA(cudaMemcpyAsync(d + i * opb, src[i], opb, cudaMemcpyHostToDevice, c.stream));
    constexpr int smem_bytes = 2 * T5_UNIT_BYTES;
    {
        LaunchScope ls("tc5_probe");
        tc5_probe_kernel<<<1, 128, smem_bytes, c.stream>>>((const __half *)d, (const __half *)(d + opb), (const __half *)(d + 2 * opb),
                                                           (const __half *)(d + 3 * opb), c0, order, (float *)(d + 4 * opb));
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CUDA(cudaMemcpyAsync(out, d + 4 * opb, outb, cudaMemcpyDeviceToHost, c.stream));
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    return PDSB_OK;
}

}  // namespace pdsb

extern "C" int pdsb_tc5_accum_probe(const uint16_t *a_hi, const uint16_t *a_lo, const uint16_t *b_hi, const uint16_t *b_lo,
                                    float c0, int order, float *out)
{
    PDSB_CHECK(pdsb::require_init());
    PDSB_REQUIRE(a_hi && a_lo && b_hi && b_lo && out, "arguments");
    return pdsb::tc5_probe(a_hi, a_lo, b_hi, b_lo, c0, order, out);
}
