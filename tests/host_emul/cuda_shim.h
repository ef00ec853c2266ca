// cuda_shim.h -- TEST INFRASTRUCTURE.  A minimal host re-interpretation of the CUDA constructs used by csrc/optim.cu and
// csrc/densify.cu (blocks run one after another; the threads of a block are real host threads, so __syncthreads and the
// warp shuffles keep their meaning).  It lets the CPU test-suite execute the kernels' SOURCE -- index arithmetic, scans,
// row ordering -- where no GPU exists.  It is not a product path and is never shipped or timed: the library proper is
// CUDA only (tests/test_kernel_source_on_host.py builds this into a throw-away .so under the pytest tmp dir).
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static

struct Idx3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Idx3 threadIdx, blockIdx;
static Idx3 gridDim;

struct BlockCtx {
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<uint32_t> xchg;
    unsigned block_threads = 0;
};
static BlockCtx *g_ctx = nullptr;

inline void __syncthreads() { g_ctx->block_bar->arrive_and_wait(); }

template <typename F>
inline uint32_t warp_exchange(uint32_t v, F pick) {
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    g_ctx->xchg[w * 32 + lane] = v;
    g_ctx->warp_bar[w]->arrive_and_wait();
    const int src = pick((int)lane);
    const uint32_t r = (src >= 0 && src < 32) ? g_ctx->xchg[w * 32 + src] : v;
    g_ctx->warp_bar[w]->arrive_and_wait();
    return r;
}
inline uint32_t __shfl_up_sync(unsigned, uint32_t v, int d) { return warp_exchange(v, [d](int l) { return l - d >= 0 ? l - d : l; }); }
inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m) { return warp_exchange(v, [m](int l) { return l ^ m; }); }
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicMin(uint32_t *p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline uint32_t atomicMax(uint32_t *p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline uint32_t __ballot_sync(unsigned, bool pred) {
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    g_ctx->xchg[w * 32 + lane] = pred ? 1u : 0u;
    g_ctx->warp_bar[w]->arrive_and_wait();
    uint32_t m = 0;
    const unsigned lanes = std::min(32u, g_ctx->block_threads - 32 * w);
    for (unsigned l = 0; l < lanes; ++l) m |= g_ctx->xchg[w * 32 + l] << l;
    g_ctx->warp_bar[w]->arrive_and_wait();
    return m;
}
inline uint32_t __match_any_sync(unsigned, uint32_t v) {
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    g_ctx->xchg[w * 32 + lane] = v;
    g_ctx->warp_bar[w]->arrive_and_wait();
    uint32_t m = 0;
    const unsigned lanes = std::min(32u, g_ctx->block_threads - 32 * w);
    for (unsigned l = 0; l < lanes; ++l) m |= (g_ctx->xchg[w * 32 + l] == v ? 1u : 0u) << l;
    g_ctx->warp_bar[w]->arrive_and_wait();
    return m;
}
inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return warp_exchange(v, [src](int) { return src; }); }
inline void __syncwarp() { g_ctx->warp_bar[threadIdx.x / 32]->arrive_and_wait(); }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
using std::isnan;
using std::max;
using std::min;

// ---- what the sources expect from common.cuh (copied definitions: plain C++) ------------------------------------------------
typedef void *cudaStream_t;
typedef int cudaError_t;
constexpr int cudaSuccess = 0;
constexpr int cudaMemcpyDeviceToDevice = 3;
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
enum { GSB_OK = 0, GSB_ERR_ARGUMENT = 1, GSB_ERR_CUDA = 2, GSB_ERR_ALLOC = 3, GSB_ERR_OVERFLOW = 4 };

namespace gsb {
inline void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
struct Carver {
    char *base;
    size_t off;
    explicit Carver(void *p) : base(static_cast<char *>(p)), off(0) {}
    template <typename T>
    T *take(size_t count) {
        off = align_up(off, 256);
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
    size_t bytes() const { return align_up(off, 256); }
};

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
    dim3(int x_) : x((unsigned)x_), y(1), z(1) {}
};

template <typename F>
inline void emul_launch(dim3 grid3, unsigned block, F body) {
    BlockCtx ctx;
    ctx.block_threads = block;
    ctx.block_bar = std::make_unique<std::barrier<>>(block);
    const unsigned warps = (block + 31) / 32;
    for (unsigned w = 0; w < warps; ++w) ctx.warp_bar.emplace_back(std::make_unique<std::barrier<>>(std::min(32u, block - 32 * w)));
    ctx.xchg.assign(warps * 32, 0u);
    g_ctx = &ctx;
    gridDim.x = grid3.x; gridDim.y = grid3.y;
    for (unsigned by = 0; by < grid3.y; ++by)
        for (unsigned b = 0; b < grid3.x; ++b) {
            std::vector<std::thread> th;
            th.reserve(block);
            for (unsigned t = 0; t < block; ++t)
                th.emplace_back([&, b, by, t] { blockIdx.x = b; blockIdx.y = by; threadIdx.x = t; body(); });
            for (auto &x : th) x.join();
        }
    g_ctx = nullptr;
}
}  // namespace gsb

#define GSB_LAUNCH(name, debug, stream, kernel, grid, block, smem, ...) \
    gsb::emul_launch(gsb::dim3(grid), (unsigned)(block), [&] { kernel(__VA_ARGS__); })
#define GSB_CUDA(expr) do { if ((expr) != cudaSuccess) return GSB_ERR_CUDA; } while (0)
