// binning.cu -- tile binning: scan of tiles-touched in depth order (K2), instance emission (K3) and
// per-tile ranges of the tile-sorted instance list (K5).
//
// Replaces the cub::DeviceScan::InclusiveSum call, duplicateWithKeys and identifyTileRanges of the
// reference's cuda_rasterizer/rasterizer_impl.cu (named in BASELINE.json north_star; absent from
// /root/reference).  Difference in structure, same resulting order: gaussians are first sorted by
// depth, instances are emitted in that order with the TILE id as the only key, and a stable sort on
// the tile id (radix_sort.cu) then yields per-tile lists ordered by (depth, gaussian index) -- exactly
// what one stable sort on tile<<32|depth keys gives.
#include "kernels.cuh"

namespace gsb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_IPT;

size_t scan_partials_count(int P) { return (size_t)ceil_div(P > 0 ? P : 1, SCAN_CHUNK) + 1; }

__device__ __forceinline__ uint32_t block_exclusive_scan_256(const uint32_t v, uint32_t *warp_sums, uint32_t &block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    uint32_t prefix = 0, total = 0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; ++k) {
        const uint32_t s = warp_sums[k];
        if (k < w) prefix += s;
        total += s;
    }
    block_total = total;
    __syncthreads();
    return prefix + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(BinArgs a) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    a.order += blockIdx.y * a.sv_gauss; a.tiles += blockIdx.y * a.sv_gauss; a.partials += blockIdx.y * a.sv_partials;
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_IPT;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        const int r = base + k;
        if (r < a.P) s += a.tiles[a.order[r]];
    }
    uint32_t total;
    block_exclusive_scan_256(s, warp_sums, total);
    if (threadIdx.x == 0) a.partials[blockIdx.x] = total;
}

// single block: exclusive scan of the per-chunk totals (64-bit running sum), grand total -> *total
__global__ void __launch_bounds__(SCAN_THREADS)
scan_partials_kernel(uint32_t *partials, const int n, unsigned long long *total_out, const size_t sv_partials) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    partials += blockIdx.y * sv_partials; total_out += blockIdx.y;
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    for (int base = 0; base < n; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n ? partials[i] : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_scan_256(v, warp_sums, total);
        const unsigned long long carry = carry_s;
        if (i < n) partials[i] = (uint32_t)(carry + excl);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(BinArgs a) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    a.order += blockIdx.y * a.sv_gauss; a.tiles += blockIdx.y * a.sv_gauss; a.partials += blockIdx.y * a.sv_partials;
    a.offsets += blockIdx.y * a.sv_gauss;
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        const int r = base + k;
        v[k] = r < a.P ? a.tiles[a.order[r]] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan_256(s, warp_sums, total) + a.partials[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        const int r = base + k;
        if (r < a.P) a.offsets[r] = run;
        run += v[k];
    }
}

int launch_tile_scan(const BinArgs &a, int V, bool debug, cudaStream_t stream) {
    const int nchunks = (int)ceil_div(a.P, SCAN_CHUNK);
    GSB_LAUNCH("scan_reduce", debug, stream, scan_reduce_kernel, dim3(nchunks, V), SCAN_THREADS, 0, a);
    GSB_LAUNCH("scan_partials", debug, stream, scan_partials_kernel, dim3(1, V), SCAN_THREADS, 0, a.partials, nchunks, a.total,
               a.sv_partials);
    GSB_LAUNCH("scan_apply", debug, stream, scan_apply_kernel, dim3(nchunks, V), SCAN_THREADS, 0, a);
    return GSB_OK;
}

__global__ void count_max_kernel(unsigned long long *counts, const int V) {
    unsigned long long m = counts[GSB_MAX_VIEWS];
    for (int v = 0; v < V; ++v) m = counts[v] > m ? counts[v] : m;
    counts[GSB_MAX_VIEWS] = m;
}

int launch_count_max(unsigned long long *counts, int V, bool debug, cudaStream_t stream) {
    GSB_LAUNCH("count_max", debug, stream, count_max_kernel, 1, 1, 0, counts, V);
    return GSB_OK;
}

// K3: one thread per depth rank; writes (tile id, gaussian id) for every tile the gaussian's cull ellipse meets.
__global__ void __launch_bounds__(256)
emit_kernel(BinArgs a, uint32_t *__restrict__ inst_tile, uint32_t *__restrict__ inst_gauss, const uint32_t cap) {
    a.order += blockIdx.y * a.sv_gauss; a.tiles += blockIdx.y * a.sv_gauss; a.rect += blockIdx.y * a.sv_gauss;
    a.offsets += blockIdx.y * a.sv_gauss; a.splat += blockIdx.y * a.sv_splat;
    inst_tile += blockIdx.y * a.sv_inst; inst_gauss += blockIdx.y * a.sv_inst;
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= a.P) return;
    const uint32_t g = a.order[r];
    uint32_t n = a.tiles[g];
    if (n == 0) return;
    const uint32_t start = a.offsets[r];
    // speculative capacity (abi.cu): never write past it; the host re-runs binning when the count did not fit
    if (start >= cap) return;
    n = min(n, cap - start);
    const float4 q0 = a.splat[(size_t)g * SPLAT_F4], q1 = a.splat[(size_t)g * SPLAT_F4 + 1], q2 = a.splat[(size_t)g * SPLAT_F4 + 2];
    const uint2 rc = a.rect[g];
    CullGeom cg;
    cg.cx = q0.x; cg.cy = q0.y; cg.A = q0.z; cg.B = q0.w; cg.C = q1.x; cg.lim = q2.z;
    cg.rx0 = (int)(rc.x & 0xffffu); cg.rx1 = (int)(rc.x >> 16);
    cg.ry0 = (int)(rc.y & 0xffffu); cg.ry1 = (int)(rc.y >> 16);
    uint32_t k = 0;
    if (cg.lim > 1.0e38f) {  // culling disabled: the whole rectangle
        for (int ty = cg.ry0; ty < cg.ry1; ++ty)
            for (int tx = cg.rx0; tx < cg.rx1 && k < n; ++tx, ++k) {
                inst_tile[start + k] = (uint32_t)(ty * a.gx + tx);
                inst_gauss[start + k] = g;
            }
    } else {
        int ty0, ty1;
        cull_rows(cg, ty0, ty1);
        for (int ty = ty0; ty < ty1; ++ty) {
            int tx0, tx1;
            cull_span(cg, ty, tx0, tx1);
            for (int tx = tx0; tx < tx1 && k < n; ++tx, ++k) {
                inst_tile[start + k] = (uint32_t)(ty * a.gx + tx);
                inst_gauss[start + k] = g;
            }
        }
    }
    // never taken when count and emit agree; keeps the slot range well defined if they ever do not
    for (; k < n; ++k) {
        inst_tile[start + k] = (uint32_t)a.num_tiles;
        inst_gauss[start + k] = g;
    }
}

int launch_emit(const BinArgs &a, int V, uint32_t *inst_tile, uint32_t *inst_gauss, int64_t cap, bool debug, cudaStream_t stream) {
    GSB_LAUNCH("emit", debug, stream, emit_kernel, dim3((int)ceil_div(a.P, 256), V), 256, 0, a, inst_tile, inst_gauss, (uint32_t)cap);
    return GSB_OK;
}

// K5: ranges[t] = [first, last+1) of tile t in the tile-sorted instance list (ranges pre-zeroed)
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint32_t *__restrict__ sorted_tiles, int64_t D, const unsigned long long *__restrict__ n_dev,
                   const int num_tiles, uint2 *ranges, const size_t sv_inst) {
    sorted_tiles += blockIdx.y * sv_inst; ranges += (size_t)blockIdx.y * num_tiles;
    if (n_dev) D = min((int64_t)n_dev[blockIdx.y], D);
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= D) return;
    const uint32_t t = sorted_tiles[i];
    if (t >= (uint32_t)num_tiles) return;
    if (i == 0 || sorted_tiles[i - 1] != t) ranges[t].x = (uint32_t)i;
    if (i == D - 1 || sorted_tiles[i + 1] != t) ranges[t].y = (uint32_t)(i + 1);
}

int launch_tile_ranges(const uint32_t *sorted_tiles, int64_t D, const unsigned long long *n_dev, int num_tiles,
                       uint2 *ranges, int V, size_t sv_inst, bool debug, cudaStream_t stream) {
    GSB_CUDA(cudaMemsetAsync(ranges, 0, (size_t)V * num_tiles * sizeof(uint2), stream));
    if (D <= 0) return GSB_OK;
    GSB_LAUNCH("tile_ranges", debug, stream, tile_ranges_kernel, dim3((int)ceil_div(D, 256), V), 256, 0, sorted_tiles, D, n_dev,
               num_tiles, ranges, sv_inst);
    return GSB_OK;
}

}  // namespace gsb
