// cuda_shim.h -- TEST INFRASTRUCTURE.  A host re-interpretation of the CUDA constructs used by gaussian-splatting_b200/csrc:
// with -DGSB_HOST_EMUL, csrc/common.cuh includes this file instead of <cuda_runtime.h> and the UNMODIFIED kernel sources
// compile with g++ into a throw-away library (tests/host_emul/build.py).  Blocks run one after another; the threads of a
// block are real host threads, so __syncthreads, the warp shuffles / votes and shared memory keep their meaning; "device"
// memory is host memory and streams / events are no-ops.
//
// Purpose: let the CPU test-suite execute the kernels' SOURCE -- index arithmetic, scans, orderings, the blend recurrences --
// against the oracle where no GPU exists, and let a kernel change be checked functionally before it costs GPU time.
// It is NOT a product path: it is never shipped, never timed, and the product library proper is CUDA only (the Python layer
// refuses CPU tensors).  Approximate-math instructions (ex2.approx, rcp.approx, ...) become their exact libm counterparts.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>

// ---- language ---------------------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __align__(n) alignas(n)
#define __constant__ static

struct GsbEmulIdx { unsigned x = 0, y = 0, z = 0; };
extern thread_local GsbEmulIdx threadIdx, blockIdx;
extern GsbEmulIdx gridDim, blockDim;

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
    dim3(int x_) : x((unsigned)x_), y(1), z(1) {}
    dim3(long x_) : x((unsigned)x_), y(1), z(1) {}
};

// ---- vector types -----------------------------------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

// ---- block context ------------------------------------------------------------------------------------------------------------
struct GsbEmulBlock {
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<uint64_t> xchg;
    unsigned threads = 0;
    std::vector<unsigned char> dyn_smem;
    std::atomic<int> vote{0};
};
extern GsbEmulBlock *g_emul_block;
inline void *gsb_emul_dynamic_smem() { return g_emul_block->dyn_smem.data(); }

inline void __syncthreads() { g_emul_block->block_bar->arrive_and_wait(); }
// block-wide votes: count the predicates between two barriers; the third barrier protects the reset
inline int __syncthreads_count(int pred) {
    GsbEmulBlock *b = g_emul_block;
    if (pred) b->vote.fetch_add(1);
    b->block_bar->arrive_and_wait();
    const int r = b->vote.load();
    b->block_bar->arrive_and_wait();
    b->vote.store(0);
    b->block_bar->arrive_and_wait();
    return r;
}
inline int __syncthreads_or(int pred) { return __syncthreads_count(pred) != 0; }
inline int __syncthreads_and(int pred) { return __syncthreads_count(!pred) == 0; }
inline void __syncwarp(unsigned = 0xffffffffu) { g_emul_block->warp_bar[threadIdx.x / 32]->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

inline unsigned gsb_emul_warp_lanes() { return std::min(32u, g_emul_block->threads - 32 * (threadIdx.x / 32)); }

// every lane publishes 64 bits, then reads the slot `pick(lane)` chooses (its own value when out of range)
template <typename T, typename F>
inline T gsb_emul_exchange(T v, F pick) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    g_emul_block->xchg[w * 32 + lane] = bits;
    g_emul_block->warp_bar[w]->arrive_and_wait();
    const int src = pick((int)lane);
    uint64_t got = (src >= 0 && src < (int)gsb_emul_warp_lanes()) ? g_emul_block->xchg[w * 32 + src] : bits;
    g_emul_block->warp_bar[w]->arrive_and_wait();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return gsb_emul_exchange(v, [src](int) { return src; }); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return gsb_emul_exchange(v, [d](int l) { return l - (int)d >= 0 ? l - (int)d : l; }); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return gsb_emul_exchange(v, [d](int l) { return l + (int)d < 32 ? l + (int)d : l; }); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return gsb_emul_exchange(v, [m](int l) { return l ^ m; }); }

inline uint32_t __ballot_sync(unsigned, int pred) {
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    g_emul_block->xchg[w * 32 + lane] = pred ? 1u : 0u;
    g_emul_block->warp_bar[w]->arrive_and_wait();
    uint32_t m = 0;
    for (unsigned l = 0; l < gsb_emul_warp_lanes(); ++l) m |= (uint32_t)g_emul_block->xchg[w * 32 + l] << l;
    g_emul_block->warp_bar[w]->arrive_and_wait();
    return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned m, int pred) {
    const unsigned lanes = gsb_emul_warp_lanes();
    return __ballot_sync(m, pred) == (lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u));
}
inline uint32_t __activemask() { const unsigned l = gsb_emul_warp_lanes(); return l == 32 ? 0xffffffffu : ((1u << l) - 1u); }
inline uint32_t __match_any_sync(unsigned, uint32_t v) {
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    g_emul_block->xchg[w * 32 + lane] = v;
    g_emul_block->warp_bar[w]->arrive_and_wait();
    uint32_t m = 0;
    for (unsigned l = 0; l < gsb_emul_warp_lanes(); ++l) m |= (g_emul_block->xchg[w * 32 + l] == v ? 1u : 0u) << l;
    g_emul_block->warp_bar[w]->arrive_and_wait();
    return m;
}
template <typename T>
inline T gsb_emul_reduce(T v, bool want_max) {
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    g_emul_block->xchg[w * 32 + lane] = bits;
    g_emul_block->warp_bar[w]->arrive_and_wait();
    T best = v;
    for (unsigned l = 0; l < gsb_emul_warp_lanes(); ++l) {
        T o;
        memcpy(&o, &g_emul_block->xchg[w * 32 + l], sizeof(T));
        best = want_max ? std::max(best, o) : std::min(best, o);
    }
    g_emul_block->warp_bar[w]->arrive_and_wait();
    return best;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return gsb_emul_reduce(v, true); }
inline int __reduce_max_sync(unsigned, int v) { return gsb_emul_reduce(v, true); }
inline unsigned __reduce_min_sync(unsigned, unsigned v) { return gsb_emul_reduce(v, false); }

// ---- atomics --------------------------------------------------------------------------------------------------------------------
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline float atomicAdd(float *p, float v) {
    uint32_t *u = reinterpret_cast<uint32_t *>(p);
    uint32_t old = __atomic_load_n(u, __ATOMIC_SEQ_CST);
    for (;;) {
        float f;
        memcpy(&f, &old, 4);
        f += v;
        uint32_t nu;
        memcpy(&nu, &f, 4);
        if (__atomic_compare_exchange_n(u, &old, nu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { memcpy(&f, &old, 4); return f; }
    }
}
template <typename T>
inline T gsb_emul_atomic_minmax(T *p, T v, bool want_max) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while ((want_max ? v > old : v < old) && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline uint32_t atomicMin(uint32_t *p, uint32_t v) { return gsb_emul_atomic_minmax(p, v, false); }
inline uint32_t atomicMax(uint32_t *p, uint32_t v) { return gsb_emul_atomic_minmax(p, v, true); }
inline int atomicMax(int *p, int v) { return gsb_emul_atomic_minmax(p, v, true); }
inline uint32_t atomicOr(uint32_t *p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

// ---- scalar intrinsics ------------------------------------------------------------------------------------------------------------
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
inline uint32_t __brev(uint32_t v) { uint32_t r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float a) { return expf(a); }
inline float __logf(float a) { return logf(a); }
inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
inline int __float2int_rn(float a) { return (int)nearbyintf(a); }
inline int __float2int_rd(float a) { return (int)floorf(a); }
inline int __float2int_ru(float a) { return (int)ceilf(a); }
inline int __float2int_rz(float a) { return (int)a; }
template <typename T> inline T __ldg(const T *p) { return *p; }
using std::isfinite;
using std::isnan;
// CUDA's min / max accept mixed integer types
template <typename A, typename B> inline typename std::common_type<A, B>::type min(A a, B b) { using C = typename std::common_type<A, B>::type; return (C)a < (C)b ? (C)a : (C)b; }
template <typename A, typename B> inline typename std::common_type<A, B>::type max(A a, B b) { using C = typename std::common_type<A, B>::type; return (C)a > (C)b ? (C)a : (C)b; }

// ---- runtime API ----------------------------------------------------------------------------------------------------------------------
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
enum { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
constexpr unsigned cudaHostAllocDefault = 0, cudaEventDisableTiming = 2;
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = calloc(1, n ? n : 1); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "host emulation"; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = reinterpret_cast<cudaEvent_t>(1); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(1); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 8;
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// ---- launch ---------------------------------------------------------------------------------------------------------------------------
void gsb_emul_run_block(unsigned threads, size_t dyn_smem, const std::function<void()> &body);

template <typename F>
inline void gsb_emul_launch(dim3 grid, dim3 block, size_t smem, F body) {
    gridDim.x = grid.x; gridDim.y = grid.y; gridDim.z = grid.z;
    blockDim.x = block.x; blockDim.y = 1; blockDim.z = 1;
    const std::function<void()> fn = body;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                GsbEmulIdx b;
                b.x = bx; b.y = by; b.z = bz;
                const std::function<void()> with_idx = [&fn, b] { blockIdx = b; fn(); };
                gsb_emul_run_block(block.x, smem, with_idx);
            }
}
