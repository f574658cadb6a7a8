// Shared host-side state and helpers of libpdsb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/pdsb.h"

namespace pdsb {

void set_error(const char *fmt, ...);

#define PDSB_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            pdsb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),      \
                            __FILE__, __LINE__);                                          \
            return PDSB_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define PDSB_CHECK(expr)                       \
    do {                                       \
        int rc__ = (expr);                     \
        if (rc__ != PDSB_OK) return rc__;      \
    } while (0)

#define PDSB_REQUIRE(cond, msg)                                          \
    do {                                                                 \
        if (!(cond)) {                                                   \
            pdsb::set_error("bad argument: %s (%s)", msg, #cond);        \
            return PDSB_ERR_ARG;                                         \
        }                                                                \
    } while (0)

// Grow-only device scratch buffer.
struct Scratch {
    void *ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <class T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

struct ProfileEntry {
    const char *name;
    cudaEvent_t start, stop;
};

struct Context {
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    int sm_clock_khz = 0;
    size_t mem_bytes = 0;
    int cc_major = 0, cc_minor = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;        // walker batches: the next cube travels while the current one is evaluated
    cudaEvent_t cube_ready[2] = {nullptr, nullptr}, cube_free[2] = {nullptr, nullptr};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool profiling = false;
    std::vector<ProfileEntry> prof;
    std::vector<cudaEvent_t> event_pool;
    int64_t launches = 0;
    int dft_variant = 0;
    int dft_split = 0;
    int grid_row_lo = 0, grid_row_hi = 0;      // pdsb_set_grid_band: output rows this process accumulates (0, 0 = all)
    const double *grid_ext_binned = nullptr;   // pdsb_set_grid_reweight: binned weight map reduced over the ranks (device)
    int64_t grid_ext_ncell = 0;
    std::vector<double> grid_ext_sumw;         //   and the per-channel weight sums
    // scratch
    Scratch img64, img64_b, folded, partial, red, stage_a, stage_b, stage_c, stage_d, stage_e;
    Scratch small_dev;     // tiny per-call device arrays (channel scale factors)
    Scratch fft_fb;        // invert(), sizes that are not a power of two: transform of Bluestein's chirp, for size fft_fb_n
    int fft_fb_n = 0;
    Scratch fft_tw;        // galario path: twiddle table exp(+2 pi i j / n), j < n/2, for size fft_tw_n
    int fft_tw_n = 0;
    Scratch fft_flags;     // per-plane "holds a non-zero value" flags of the cube being transformed
    Scratch nufft_corr;    // NUFFT path: deapodisation factors 1 / psi_hat(X / N) for image side nufft_corr_n
    int nufft_corr_n = 0, nufft_corr_N = 0;
    Scratch mma_ws;        // tensor-core kernel: per-round lattice quanta and per-plane unscale factors
};

Context &ctx();
int require_init();

// Kernel launch bookkeeping: counts launches, brackets them with events when profiling.
struct LaunchScope {
    const char *name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    explicit LaunchScope(const char *n);
    ~LaunchScope();
};

// Pointer staging: returns a device pointer for `p`; copies from host when kind==HOST.
int to_device(const void *p, int kind, size_t bytes, Scratch &scratch, const void **dev);
// host <-> device copies on the context's stream; large pageable arrays go through a multi-threaded pinned ring
// (runtime.cu).  copy_h2d returns when the source may be reused, copy_d2h when the destination holds the data.
int copy_h2d(void *dst_dev, const void *src_host, size_t bytes);
int copy_h2d_on(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream);      // the same on another stream
int copy_d2h(void *dst_host, const void *src_dev, size_t bytes);

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace pdsb
