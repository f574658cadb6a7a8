// libpdsb runtime: context, stream, buffers, timing, profiling, the FP32 FMA peak measurement.
#include "common.cuh"
#include "dft.cuh"
#include <thread>
#include <algorithm>

namespace pdsb {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

Context &ctx()
{
    static Context c;
    return c;
}

int require_init()
{
    if (!ctx().inited) {
        int rc = pdsb_init(0);
        if (rc != PDSB_OK) return rc;
    }
    return PDSB_OK;
}

int Scratch::ensure(size_t bytes)
{
    if (bytes <= cap) return PDSB_OK;
    if (ptr) {
        // the previous buffer may still be in use by queued work on the stream
        cudaStreamSynchronize(ctx().stream);
        cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        ptr = nullptr;
        return PDSB_ERR_NOMEM;
    }
    cap = want;
    return PDSB_OK;
}

void Scratch::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

static cudaEvent_t get_event()
{
    Context &c = ctx();
    if (!c.event_pool.empty()) {
        cudaEvent_t e = c.event_pool.back();
        c.event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

LaunchScope::LaunchScope(const char *n) : name(n)
{
    Context &c = ctx();
    c.launches++;
    if (c.profiling) {
        e0 = get_event();
        e1 = get_event();
        cudaEventRecord(e0, c.stream);
    }
}

LaunchScope::~LaunchScope()
{
    Context &c = ctx();
    if (e0) {
        cudaEventRecord(e1, c.stream);
        c.prof.push_back({name, e0, e1});
    }
}

// ---------------------------------------------------------------------------------
// Large transfers between PAGEABLE host memory (what numpy hands over) and the device.  cudaMemcpyAsync stages such
// copies through the driver's own pinned buffer with one thread (8-12 GB/s here); several host threads copying
// 4 MB chunks into a ring of pinned buffers of ours, each chunk going on with cudaMemcpyAsync as soon as it is
// staged, keep the PCIe link busy instead.  Pinned / registered memory and small arrays go straight through.
// work(t) for t < nt on nt threads (the caller is thread 0).  A thread that cannot be created (resource limits) must
// not throw across the C ABI: its share runs on the calling thread instead.
template <class F>
static void run_threads(int nt, F &&work)
{
    std::vector<std::thread> th;
    std::vector<int> inline_ids;
    for (int t = 1; t < nt; t++) {
        try {
            th.emplace_back(work, t);
        } catch (...) {
            inline_ids.push_back(t);
        }
    }
    work(0);
    for (int t : inline_ids) work(t);
    for (auto &t : th) t.join();
}

struct StagePool {
    static constexpr size_t CHUNK = (size_t)4 << 20;
    int nthreads = 0;
    std::vector<void *> bufs;          // 2 per thread
    std::vector<cudaEvent_t> evs;
};
static StagePool &stage_pool()
{
    static StagePool sp;
    if (sp.nthreads == 0) {
        // copy threads: the host cores divided among the ranks of this node, between 2 and 8 (measured on the 16-core
        // B200 host: 8 threads move a pageable 134 MB cube as fast as 16 or 24; PDSB_COPY_THREADS overrides)
        unsigned nt = std::thread::hardware_concurrency();
        const char *lw = getenv("LOCAL_WORLD_SIZE"), *ov = getenv("PDSB_COPY_THREADS");
        const unsigned ranks = lw && atoi(lw) > 0 ? (unsigned)atoi(lw) : 1u;
        nt = nt == 0 ? 4 : nt / ranks;
        nt = nt < 2 ? 2 : nt > 8 ? 8 : nt;
        if (ov && atoi(ov) > 0) nt = (unsigned)std::min(atoi(ov), 32);
        for (unsigned i = 0; i < 2 * nt; i++) {
            void *b = nullptr;
            cudaEvent_t e = nullptr;
            if (cudaMallocHost(&b, StagePool::CHUNK) != cudaSuccess ||
                cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                break;
            }
            sp.bufs.push_back(b);
            sp.evs.push_back(e);
        }
        sp.nthreads = (int)(sp.bufs.size() / 2);
        if (sp.nthreads == 0) sp.nthreads = -1;          // no pinned memory to be had: always the direct path
    }
    return sp;
}
static bool pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
static constexpr size_t STAGED_MIN_BYTES = (size_t)16 << 20;

int copy_h2d(void *dst, const void *src, size_t bytes) { return copy_h2d_on(dst, src, bytes, ctx().stream); }

int copy_h2d_on(void *dst, const void *src, size_t bytes, cudaStream_t stream)
{
    Context &c = ctx();
    if (bytes == 0) return PDSB_OK;
    StagePool *sp = bytes >= STAGED_MIN_BYTES && pageable(src) ? &stage_pool() : nullptr;
    if (!sp || sp->nthreads < 1) {
        PDSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return PDSB_OK;
    }
    const size_t nchunk = (bytes + StagePool::CHUNK - 1) / StagePool::CHUNK;
    const int nt = (int)std::min<size_t>((size_t)sp->nthreads, nchunk);
    std::vector<cudaError_t> err((size_t)nt, cudaSuccess);
    auto work = [&](int t) {
        cudaSetDevice(c.device);
        int k = 0;
        for (size_t ch = (size_t)t; ch < nchunk; ch += (size_t)nt, k++) {
            const int b = 2 * t + (k & 1);
            const size_t off = ch * StagePool::CHUNK, len = std::min(StagePool::CHUNK, bytes - off);
            cudaError_t e = k >= 2 ? cudaEventSynchronize(sp->evs[b]) : cudaSuccess;     // buffer free again?
            memcpy(sp->bufs[b], static_cast<const char *>(src) + off, len);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(static_cast<char *>(dst) + off, sp->bufs[b], len, cudaMemcpyHostToDevice, stream);
            if (e == cudaSuccess) e = cudaEventRecord(sp->evs[b], stream);
            if (e != cudaSuccess) err[t] = e;
        }
        for (int q = 0; q < 2 && q < k; q++) {                                          // leave the ring idle
            const cudaError_t e = cudaEventSynchronize(sp->evs[2 * t + q]);
            if (e != cudaSuccess) err[t] = e;
        }
    };
    run_threads(nt, work);
    for (cudaError_t e : err) PDSB_CUDA(e);
    return PDSB_OK;
}

int copy_d2h(void *dst, const void *src, size_t bytes)
{
    Context &c = ctx();
    if (bytes == 0) return PDSB_OK;
    StagePool *sp = bytes >= STAGED_MIN_BYTES && pageable(dst) ? &stage_pool() : nullptr;
    if (!sp || sp->nthreads < 1) {
        PDSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.stream));
        return PDSB_OK;
    }
    const size_t nchunk = (bytes + StagePool::CHUNK - 1) / StagePool::CHUNK;
    const int nt = (int)std::min<size_t>((size_t)sp->nthreads, nchunk);
    std::vector<cudaError_t> err((size_t)nt, cudaSuccess);
    auto work = [&](int t) {
        cudaSetDevice(c.device);
        int k = 0;
        size_t prev_off = 0, prev_len = 0;
        auto drain = [&](int kk) {                       // chunk kk of this thread: wait for it, hand it to the caller
            const int b = 2 * t + (kk & 1);
            const cudaError_t e = cudaEventSynchronize(sp->evs[b]);
            if (e != cudaSuccess) err[t] = e;
            else memcpy(static_cast<char *>(dst) + prev_off, sp->bufs[b], prev_len);
        };
        for (size_t ch = (size_t)t; ch < nchunk; ch += (size_t)nt, k++) {
            const int b = 2 * t + (k & 1);
            const size_t off = ch * StagePool::CHUNK, len = std::min(StagePool::CHUNK, bytes - off);
            cudaError_t e = cudaMemcpyAsync(sp->bufs[b], static_cast<const char *>(src) + off, len, cudaMemcpyDeviceToHost,
                                            c.stream);
            if (e == cudaSuccess) e = cudaEventRecord(sp->evs[b], c.stream);
            if (e != cudaSuccess) err[t] = e;
            if (k >= 1) drain(k - 1);                    // the previous chunk, while this one is on the wire
            prev_off = off;
            prev_len = len;
        }
        if (k >= 1) drain(k - 1);
    };
    run_threads(nt, work);
    for (cudaError_t e : err) PDSB_CUDA(e);
    return PDSB_OK;
}

int to_device(const void *p, int kind, size_t bytes, Scratch &scratch, const void **dev)
{
    if (kind == PDSB_DEVICE) {
        *dev = p;
        return PDSB_OK;
    }
    PDSB_CHECK(scratch.ensure(bytes));
    PDSB_CHECK(copy_h2d(scratch.ptr, p, bytes));
    *dev = scratch.ptr;
    return PDSB_OK;
}

// ---------------------------------------------------------------------------------
// FMA microbenchmark: register-resident dependent chains, enough independent chains per
// thread to cover the pipe latency.  variant 0: scalar FFMA; variant 1: FFMA2 (f32x2).
__device__ __forceinline__ unsigned long long fma2_(unsigned long long a, unsigned long long b,
                                                    unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int VAR>
__global__ void __launch_bounds__(256) fma_bench_kernel(float *out, int iters, float seed)
{
    constexpr int NCH = 16;
    float m = 1.0f + seed * 1e-9f, b = seed * 1e-9f;
    if (VAR == 0) {
        float acc[NCH];
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = threadIdx.x * 1e-3f + i;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < NCH; i++) acc[i] = fmaf(acc[i], m, b);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < NCH; i++) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        unsigned long long acc[NCH / 2];
        float2 mm = make_float2(m, m), bb = make_float2(b, b);
        unsigned long long m2 = *reinterpret_cast<unsigned long long *>(&mm);
        unsigned long long b2 = *reinterpret_cast<unsigned long long *>(&bb);
#pragma unroll
        for (int i = 0; i < NCH / 2; i++) {
            float2 a = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
            acc[i] = *reinterpret_cast<unsigned long long *>(&a);
        }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < NCH / 2; i++) acc[i] = fma2_(acc[i], m2, b2);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < NCH / 2; i++) {
            float2 a = *reinterpret_cast<float2 *>(&acc[i]);
            s += a.x + a.y;
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

}  // namespace pdsb

using namespace pdsb;

extern "C" {

int pdsb_version(void) { return PDSB_VERSION; }

const char *pdsb_last_error(void) { return g_err; }

int pdsb_device_count(int *count)
{
    PDSB_REQUIRE(count, "count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return PDSB_ERR_CUDA;
    }
    *count = n;
    return PDSB_OK;
}

int pdsb_init(int device)
{
    Context &c = ctx();
    if (c.inited && c.device == device) return PDSB_OK;
    if (c.inited) {
        set_error("already initialised on device %d; call pdsb_shutdown first", c.device);
        return PDSB_ERR_STATE;
    }
    int n = 0;
    PDSB_CHECK(pdsb_device_count(&n));
    if (n <= 0) {
        set_error("no CUDA device visible: libpdsb has no CPU fallback");
        return PDSB_ERR_CUDA;
    }
    PDSB_REQUIRE(device >= 0 && device < n, "device index");
    PDSB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    PDSB_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; libpdsb is built for sm_100a only", device, p.major, p.minor);
        return PDSB_ERR_CUDA;
    }
    c.device = device;
    c.sm_count = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    c.sm_clock_khz = khz;
    c.mem_bytes = p.totalGlobalMem;
    c.cc_major = p.major;
    c.cc_minor = p.minor;
    PDSB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    PDSB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        PDSB_CUDA(cudaEventCreateWithFlags(&c.cube_ready[i], cudaEventDisableTiming));
        PDSB_CUDA(cudaEventCreateWithFlags(&c.cube_free[i], cudaEventDisableTiming));
    }
    PDSB_CUDA(cudaEventCreate(&c.t0));
    PDSB_CUDA(cudaEventCreate(&c.t1));
    c.inited = true;
    return PDSB_OK;
}

int pdsb_shutdown(void)
{
    Context &c = ctx();
    if (!c.inited) return PDSB_OK;
    cudaStreamSynchronize(c.stream);
    cudaStreamSynchronize(c.copy_stream);
    for (Scratch *s : {&c.img64, &c.img64_b, &c.folded, &c.partial, &c.red, &c.stage_a, &c.stage_b, &c.stage_c,
                       &c.stage_d, &c.stage_e, &c.small_dev, &c.mma_ws, &c.fft_fb, &c.fft_tw, &c.fft_flags, &c.nufft_corr})
        s->release();
    c.fft_fb_n = 0;
    for (auto &p : c.prof) {
        cudaEventDestroy(p.start);
        cudaEventDestroy(p.stop);
    }
    c.prof.clear();
    for (auto e : c.event_pool) cudaEventDestroy(e);
    c.event_pool.clear();
    cudaEventDestroy(c.t0);
    cudaEventDestroy(c.t1);
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(c.cube_ready[i]);
        cudaEventDestroy(c.cube_free[i]);
    }
    cudaStreamDestroy(c.copy_stream);
    cudaStreamDestroy(c.own_stream);
    c = Context();
    return PDSB_OK;
}

int pdsb_get_stream(uint64_t *stream)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(stream, "stream");
    *stream = (uint64_t)(uintptr_t)ctx().stream;
    return PDSB_OK;
}

int pdsb_set_stream(uint64_t stream)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    c.stream = (cudaStream_t)(uintptr_t)stream;      // 0 is the legacy default stream (torch's default)
    return PDSB_OK;
}

int pdsb_reset_stream(void)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    c.stream = c.own_stream;
    return PDSB_OK;
}

int pdsb_synchronize(void)
{
    PDSB_CHECK(require_init());
    PDSB_CUDA(cudaStreamSynchronize(ctx().stream));
    return PDSB_OK;
}

int pdsb_device_info(int *sm_count, int *sm_clock_khz, int64_t *mem_bytes, int *cc_major, int *cc_minor)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    if (sm_count) *sm_count = c.sm_count;
    if (sm_clock_khz) *sm_clock_khz = c.sm_clock_khz;
    if (mem_bytes) *mem_bytes = (int64_t)c.mem_bytes;
    if (cc_major) *cc_major = c.cc_major;
    if (cc_minor) *cc_minor = c.cc_minor;
    return PDSB_OK;
}

int pdsb_device_alloc(void **ptr, int64_t bytes)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ptr && bytes >= 0, "ptr/bytes");
    *ptr = nullptr;
    if (bytes == 0) return PDSB_OK;
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
        return PDSB_ERR_NOMEM;
    }
    return PDSB_OK;
}

int pdsb_device_free(void *ptr)
{
    if (!ptr) return PDSB_OK;
    PDSB_CHECK(require_init());
    PDSB_CUDA(cudaStreamSynchronize(ctx().stream));
    PDSB_CUDA(cudaFree(ptr));
    return PDSB_OK;
}

int pdsb_host_alloc_pinned(void **ptr, int64_t bytes)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(ptr && bytes > 0, "ptr/bytes");
    cudaError_t e = cudaMallocHost(ptr, (size_t)bytes);
    if (e != cudaSuccess) {
        set_error("cudaMallocHost(%lld): %s", (long long)bytes, cudaGetErrorString(e));
        return PDSB_ERR_NOMEM;
    }
    return PDSB_OK;
}

int pdsb_host_free_pinned(void *ptr)
{
    if (!ptr) return PDSB_OK;
    PDSB_CUDA(cudaFreeHost(ptr));
    return PDSB_OK;
}

int pdsb_memcpy(void *dst, int kind_dst, const void *src, int kind_src, int64_t bytes)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(bytes >= 0, "bytes");
    if (bytes == 0) return PDSB_OK;
    PDSB_REQUIRE(dst && src, "dst/src");
    cudaMemcpyKind k = kind_dst == PDSB_DEVICE
                           ? (kind_src == PDSB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                           : (kind_src == PDSB_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
    if (k == cudaMemcpyHostToDevice) return copy_h2d(dst, src, (size_t)bytes);      // large pageable arrays: staged ring
    if (k == cudaMemcpyDeviceToHost) return copy_d2h(dst, src, (size_t)bytes);
    PDSB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, k, ctx().stream));
    return PDSB_OK;
}

int pdsb_memset(void *dev_ptr, int value, int64_t bytes)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(dev_ptr && bytes >= 0, "ptr/bytes");
    PDSB_CUDA(cudaMemsetAsync(dev_ptr, value, (size_t)bytes, ctx().stream));
    return PDSB_OK;
}

int pdsb_timer_start(void)
{
    PDSB_CHECK(require_init());
    PDSB_CUDA(cudaEventRecord(ctx().t0, ctx().stream));
    return PDSB_OK;
}

int pdsb_timer_stop(double *elapsed_ms)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(elapsed_ms, "elapsed_ms");
    PDSB_CUDA(cudaEventRecord(ctx().t1, ctx().stream));
    PDSB_CUDA(cudaEventSynchronize(ctx().t1));
    float ms = 0;
    PDSB_CUDA(cudaEventElapsedTime(&ms, ctx().t0, ctx().t1));
    *elapsed_ms = ms;
    return PDSB_OK;
}

int pdsb_profile_enable(int on)
{
    PDSB_CHECK(require_init());
    ctx().profiling = on != 0;
    return PDSB_OK;
}

int pdsb_profile_reset(void)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    for (auto &p : c.prof) {
        c.event_pool.push_back(p.start);
        c.event_pool.push_back(p.stop);
    }
    c.prof.clear();
    return PDSB_OK;
}

int pdsb_profile_get(const char *prefix, double *total_ms, int64_t *launches)
{
    PDSB_CHECK(require_init());
    Context &c = ctx();
    PDSB_CUDA(cudaStreamSynchronize(c.stream));
    double tot = 0;
    int64_t n = 0;
    size_t plen = prefix ? strlen(prefix) : 0;
    for (auto &p : c.prof) {
        if (plen && strncmp(p.name, prefix, plen) != 0) continue;
        float ms = 0;
        PDSB_CUDA(cudaEventElapsedTime(&ms, p.start, p.stop));
        tot += ms;
        n++;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return PDSB_OK;
}

int pdsb_launch_count(int64_t *count)
{
    PDSB_REQUIRE(count, "count");
    *count = ctx().launches;
    return PDSB_OK;
}

// Content hash of a host array: what the Python handle cache compares before it trusts a device copy
// (pdspy_b200/device.py).  Memory-bound: 4 interleaved multiply-rotate lanes per thread, chunks of the array
// on up to 32 host threads, chunk hashes combined in order.  Needs no CUDA device.
static uint64_t hash_chunk(const unsigned char *p, size_t n)
{
    const uint64_t K = 0x9E3779B97F4A7C15ull;
    uint64_t h[4] = {0x243F6A8885A308D3ull, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
    const size_t nw = n / 8;
    size_t i = 0;
    for (; i + 4 <= nw; i += 4) {
        uint64_t w[4];
        memcpy(w, p + i * 8, 32);
        for (int l = 0; l < 4; l++) {
            h[l] = (h[l] ^ w[l]) * K;
            h[l] = (h[l] << 29) | (h[l] >> 35);
        }
    }
    uint64_t tail = 0;
    for (; i < nw; i++) {
        uint64_t w;
        memcpy(&w, p + i * 8, 8);
        h[i & 3] = ((h[i & 3] ^ w) * K);
        h[i & 3] = (h[i & 3] << 29) | (h[i & 3] >> 35);
    }
    if (n % 8) memcpy(&tail, p + nw * 8, n % 8);
    uint64_t r = (uint64_t)n * K ^ tail;
    for (int l = 0; l < 4; l++) r = ((r ^ h[l]) * K) ^ (r >> 31);
    return r;
}

int pdsb_hash64(const void *host_ptr, int64_t bytes, uint64_t *out)
{
    PDSB_REQUIRE(out && bytes >= 0 && (bytes == 0 || host_ptr), "hash arguments");
    const unsigned char *p = static_cast<const unsigned char *>(host_ptr);
    const size_t n = (size_t)bytes, chunk = (size_t)4 << 20;
    const size_t nchunk = (n + chunk - 1) / chunk;
    std::vector<uint64_t> hs(nchunk ? nchunk : 1, 0);
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 32) nt = 32;
    if (nt > nchunk) nt = (unsigned)(nchunk ? nchunk : 1);
    auto work = [&](unsigned t) {
        for (size_t c = t; c < nchunk; c += nt) {
            const size_t o = c * chunk;
            hs[c] = hash_chunk(p + o, n - o < chunk ? n - o : chunk);
        }
    };
    run_threads((int)nt, work);
    *out = hash_chunk(reinterpret_cast<const unsigned char *>(hs.data()), nchunk * sizeof(uint64_t)) ^ (uint64_t)n;
    return PDSB_OK;
}

int pdsb_set_dft_variant(int variant)
{
    bool ok = variant == 0 || (variant >= 1 && variant <= dft_variant_count()) ||
              variant == DFT_VARIANT_TC5 || variant == DFT_VARIANT_TC5 + 1 || variant == DFT_VARIANT_F64 ||
              variant == DFT_VARIANT_NUFFT;
    PDSB_REQUIRE(ok, "unknown DFT kernel variant");
    ctx().dft_variant = variant;
    return PDSB_OK;
}

int pdsb_set_dft_split(int nsplit)
{
    PDSB_REQUIRE(nsplit >= 0, "nsplit");
    ctx().dft_split = nsplit;
    return PDSB_OK;
}

int pdsb_bench_fma(int variant, int iters, double *tflops, double *ms_out)
{
    PDSB_CHECK(require_init());
    PDSB_REQUIRE(tflops && iters > 0, "tflops/iters");
    PDSB_REQUIRE(variant == 0 || variant == 1, "variant (0 = FFMA, 1 = FFMA2)");
    Context &c = ctx();
    const int blocks = c.sm_count * 8, threads = 256;
    PDSB_CHECK(c.red.ensure((size_t)blocks * 256 * sizeof(float)));
    const double fmas_per_thread_iter = 8.0 * 16.0;
    for (int rep = 0; rep < 2; rep++) {
        if (rep == 1) PDSB_CUDA(cudaEventRecord(c.t0, c.stream));
        {
            LaunchScope ls("fma_bench");
            float *o = c.red.as<float>();
            if (variant == 0) fma_bench_kernel<0><<<blocks, threads, 0, c.stream>>>(o, iters, 1.0f);
            else fma_bench_kernel<1><<<blocks, threads, 0, c.stream>>>(o, iters, 1.0f);
        }
        PDSB_CUDA(cudaGetLastError());
    }
    PDSB_CUDA(cudaEventRecord(c.t1, c.stream));
    PDSB_CUDA(cudaEventSynchronize(c.t1));
    float ms = 0;
    PDSB_CUDA(cudaEventElapsedTime(&ms, c.t0, c.t1));
    double fmas = (double)blocks * threads * (double)iters * fmas_per_thread_iter;
    *tflops = 2.0 * fmas / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return PDSB_OK;
}

}  // extern "C"
