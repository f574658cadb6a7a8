// Microbenchmark (not part of libpdsb): how fast can FFMA2 (fma.rn.f32x2) stream on sm_100a as a function of where
// its operands come from?  The direct-Fourier kernel's inner loop is  acc[q] = fma2(x, trig[q][t], acc[q])  with x
// shared by the UVT uv points of a thread and trig / acc distinct per instruction; this probe runs that pattern
// with everything in registers (no memory at all) for several UVT, plus the two bounding cases.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/probe_ffma2 scripts/probe_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}

// MODE 0: acc[q] = fma2(x[j], T[q][j], acc[q])   j outer, q inner            (the kernel's order: x reused over q)
// MODE 1: same operands, snake order over (x pair, q) so that consecutive instructions share x or T alternately
// MODE 2: acc[q] = fma2(A, B, acc[q])             two fixed operands         (upper bound: reuse cache always hits)
// MODE 3: acc[q] = fma2(x[j], T[q][j], acc[q])   q outer, j inner            (nothing shared between neighbours)
template <int UVT, int NT, int MODE>
__global__ void __launch_bounds__(128) k(float *out, int iters, float seed)
{
    u64 T[UVT][NT], x[NT], acc[UVT][2];
#pragma unroll
    for (int j = 0; j < NT; j++) {
        x[j] = pack2(seed + 1e-3f * (threadIdx.x + j), seed - 1e-3f * (threadIdx.x * 3 + j));
#pragma unroll
        for (int q = 0; q < UVT; q++) T[q][j] = pack2(1e-3f * (q + j + threadIdx.x), 1e-3f * (q * 7 + j));
    }
#pragma unroll
    for (int q = 0; q < UVT; q++) acc[q][0] = acc[q][1] = 0ull;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < NT; j += 2)
#pragma unroll
                for (int q = 0; q < UVT; q++) {
                    acc[q][0] = fma2(x[j], T[q][j], acc[q][0]);
                    acc[q][1] = fma2(x[j + 1], T[q][j], acc[q][1]);
                }
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < NT; j += 2)
#pragma unroll
                for (int q = 0; q < UVT; q++) {
                    if (q & 1) {
                        acc[q][1] = fma2(x[j + 1], T[q][j], acc[q][1]);
                        acc[q][0] = fma2(x[j], T[q][j], acc[q][0]);
                    } else {
                        acc[q][0] = fma2(x[j], T[q][j], acc[q][0]);
                        acc[q][1] = fma2(x[j + 1], T[q][j], acc[q][1]);
                    }
                }
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < NT; j += 2)
#pragma unroll
                for (int q = 0; q < UVT; q++) {
                    acc[q][0] = fma2(x[0], T[0][0], acc[q][0]);
                    acc[q][1] = fma2(x[0], T[0][0], acc[q][1]);
                }
        } else {
#pragma unroll
            for (int q = 0; q < UVT; q++)
#pragma unroll
                for (int j = 0; j < NT; j += 2) {
                    acc[q][0] = fma2(x[j], T[q][j], acc[q][0]);
                    acc[q][1] = fma2(x[j + 1], T[q][j + 1], acc[q][1]);
                }
        }
    }
    float s = 0;
#pragma unroll
    for (int q = 0; q < UVT; q++) {
        float2 a = *reinterpret_cast<float2 *>(&acc[q][0]), b = *reinterpret_cast<float2 *>(&acc[q][1]);
        s += a.x + a.y + b.x + b.y;
    }
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int UVT, int NT, int MODE>
static void run(float *out, int sm, const char *what)
{
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<UVT, NT, MODE>, 128, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k<UVT, NT, MODE>);
    const int blocks = sm * per_sm, iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<UVT, NT, MODE><<<blocks, 128>>>(out, iters, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double ffma2 = (double)blocks * 128 * iters * UVT * NT;          // lane instructions
    const double tf = ffma2 * 4.0 / (best * 1e-3) / 1e12;
    printf("%-34s uv%d nt%-2d regs %3d CTAs/SM %2d  %.3f ms  %.2f TFLOP/s = %.3f of 74.45\n", what, UVT, NT, fa.numRegs, per_sm,
           best, tf, tf / 74.45);
}

int main()
{
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    cudaMalloc(&out, (size_t)sm * 32 * 128 * sizeof(float));
    run<3, 16, 2>(out, sm, "two fixed operands (reuse)");
    run<2, 16, 0>(out, sm, "x shared over q (kernel order)");
    run<3, 16, 0>(out, sm, "x shared over q (kernel order)");
    run<4, 16, 0>(out, sm, "x shared over q (kernel order)");
    run<6, 16, 0>(out, sm, "x shared over q (kernel order)");
    run<8, 8, 0>(out, sm, "x shared over q (kernel order)");
    run<3, 32, 0>(out, sm, "x shared over q (kernel order)");
    run<3, 16, 1>(out, sm, "snake: x or T shared alternately");
    run<4, 16, 1>(out, sm, "snake: x or T shared alternately");
    run<3, 32, 1>(out, sm, "snake: x or T shared alternately");
    run<3, 16, 3>(out, sm, "nothing shared between neighbours");
    run<3, 32, 3>(out, sm, "nothing shared between neighbours");
    return 0;
}
