// knn.cu -- mean squared distance to the three nearest neighbours of every point: the replacement of
// simple_knn._C.distCUDA2, which the reference calls once to initialise the scales (scene/gaussian_model.py:21,159;
// SURVEY.md section 8(f) #4).  The simple-knn submodule is absent from /root/reference, so the definition is restated
// [UNVERIFIED_VS_REFERENCE]: out[i] = (d1 + d2 + d3) / 3 with d1 <= d2 <= d3 the squared distances from point i to its
// three nearest OTHER points (identical coordinates count as distance 0; only index i itself is skipped).  With fewer
// than three other points the mean is taken over those that exist (0 for a single point).
//
// Exact search on a uniform grid sized on the device (no host read-back): bounding box -> cell size such that there are
// about two cells per point -> counting sort of the points by cell -> every point scans Chebyshev shells of cells around
// its own and stops as soon as its third-best distance is inside the radius the visited shells cover.
#include "common.cuh"
#include "kernels.cuh"

#include <float.h>

namespace gsb {

constexpr int KNN_THREADS = 256;

struct KnnGrid {
    float origin[3];
    float h, inv_h;
    int dims[3];
};

struct KnnBuffers {
    uint32_t *bbox;        // 6 order-preserving encodings: min xyz, max xyz
    KnnGrid *grid;
    uint32_t *cell_of;     // [P]
    uint32_t *counts;      // [C + 1] -> exclusive starts after the scan (entry C = P)
    uint32_t *cursor;      // [C]
    float4 *sorted;        // [P] xyz + original index
    uint32_t *partials;
    uint32_t *total;
};

static size_t knn_cells(int64_t P) {
    int64_t c = 2 * P;
    if (c < 64) c = 64;
    if (c > ((int64_t)1 << 27)) c = (int64_t)1 << 27;
    return (size_t)c;
}

static KnnBuffers carve_knn(void *base, int64_t P, size_t *bytes) {
    const size_t C = knn_cells(P);
    Carver c(base);
    KnnBuffers b;
    b.bbox = c.take<uint32_t>(8);
    b.grid = c.take<KnnGrid>(1);
    b.cell_of = c.take<uint32_t>((size_t)P);
    b.counts = c.take<uint32_t>(C + 1);
    b.cursor = c.take<uint32_t>(C);
    b.sorted = c.take<float4>((size_t)P);
    b.partials = c.take<uint32_t>(scan_u32_partials((int64_t)C + 1));
    b.total = c.take<uint32_t>(4);
    if (bytes) *bytes = c.bytes();
    return b;
}

size_t knn_scratch_bytes(int64_t P) {
    size_t b = 0;
    carve_knn(nullptr, P, &b);
    return b;
}

// float <-> unsigned with the same ordering
__device__ __forceinline__ uint32_t order_bits(const float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(const uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_bbox_kernel(const float *__restrict__ pts, const uint32_t P, uint32_t *bbox) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = blockIdx.x * KNN_THREADS + threadIdx.x; i < P; i += gridDim.x * KNN_THREADS) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = pts[3u * (size_t)i + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t a = order_bits(lo[k]), b = order_bits(hi[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bbox + k, a);
            atomicMax(bbox + 3 + k, b);
        }
    }
}

// one thread: the smallest cell size on a geometric ladder whose grid fits into max_cells
__global__ void knn_grid_kernel(const uint32_t *__restrict__ bbox, const uint32_t max_cells, KnnGrid *grid) {
    float ext[3], big = 0.0f;
    for (int k = 0; k < 3; ++k) {
        grid->origin[k] = unorder_bits(bbox[k]);
        ext[k] = unorder_bits(bbox[3 + k]) - grid->origin[k];
        big = fmaxf(big, ext[k]);
    }
    float h = big > 0.0f ? big * 1.0001f : 1.0f;
    int dims[3] = {1, 1, 1};
    for (int it = 0; it < 120 && big > 0.0f; ++it) {
        const float hn = h * 0.7937005f;                         // 2^(-1/3): twice the cells per step
        int d[3];
        double cells = 1.0;
        bool ok = true;
        for (int k = 0; k < 3; ++k) {
            const float q = floorf(ext[k] / hn) + 1.0f;
            if (q > 1024.0f) ok = false;
            d[k] = (int)fminf(q, 1024.0f);
            cells *= (double)d[k];
        }
        if (!ok || cells > (double)max_cells) break;
        h = hn;
        for (int k = 0; k < 3; ++k) dims[k] = d[k];
    }
    grid->h = h;
    grid->inv_h = 1.0f / h;
    for (int k = 0; k < 3; ++k) grid->dims[k] = dims[k];
}

__device__ __forceinline__ void cell_coords(const KnnGrid &g, const float x, const float y, const float z, int c[3]) {
    c[0] = min(g.dims[0] - 1, max(0, (int)((x - g.origin[0]) * g.inv_h)));
    c[1] = min(g.dims[1] - 1, max(0, (int)((y - g.origin[1]) * g.inv_h)));
    c[2] = min(g.dims[2] - 1, max(0, (int)((z - g.origin[2]) * g.inv_h)));
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_count_kernel(const float *__restrict__ pts, const uint32_t P, const KnnGrid *__restrict__ grid, uint32_t *__restrict__ cell_of,
                 uint32_t *counts) {
    const uint32_t i = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (i >= P) return;
    const KnnGrid g = *grid;
    int c[3];
    cell_coords(g, pts[3u * (size_t)i], pts[3u * (size_t)i + 1], pts[3u * (size_t)i + 2], c);
    const uint32_t cell = ((uint32_t)c[2] * g.dims[1] + c[1]) * g.dims[0] + c[0];
    cell_of[i] = cell;
    atomicAdd(counts + cell, 1u);
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_scatter_kernel(const float *__restrict__ pts, const uint32_t P, const uint32_t *__restrict__ cell_of,
                   const uint32_t *__restrict__ starts, uint32_t *cursor, float4 *__restrict__ sorted) {
    const uint32_t i = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (i >= P) return;
    const uint32_t cell = cell_of[i];
    const uint32_t pos = starts[cell] + atomicAdd(cursor + cell, 1u);
    sorted[pos] = make_float4(pts[3u * (size_t)i], pts[3u * (size_t)i + 1], pts[3u * (size_t)i + 2], __uint_as_float(i));
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_query_kernel(const float4 *__restrict__ sorted, const uint32_t P, const KnnGrid *__restrict__ grid,
                 const uint32_t *__restrict__ starts, float *__restrict__ out) {
    const uint32_t t = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (t >= P) return;
    const KnnGrid g = *grid;
    const float4 me = sorted[t];
    int c[3];
    cell_coords(g, me.x, me.y, me.z, c);
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    const int rmax = max(g.dims[0], max(g.dims[1], g.dims[2]));
    for (int r = 0; r <= rmax; ++r) {
        for (int dz = -r; dz <= r; ++dz) {
            const int z = c[2] + dz;
            if (z < 0 || z >= g.dims[2]) continue;
            for (int dy = -r; dy <= r; ++dy) {
                const int y = c[1] + dy;
                if (y < 0 || y >= g.dims[1]) continue;
                const bool face = (dz == -r || dz == r || dy == -r || dy == r);     // on the shell whatever dx is
                const int step = face ? 1 : (r > 0 ? 2 * r : 1);
                for (int dx = -r; dx <= r; dx += step) {
                    const int x = c[0] + dx;
                    if (x < 0 || x >= g.dims[0]) continue;
                    const uint32_t cell = ((uint32_t)z * g.dims[1] + y) * g.dims[0] + x;
                    const uint32_t end = starts[cell + 1];
                    for (uint32_t j = starts[cell]; j < end; ++j) {
                        if (j == t) continue;
                        const float4 q = sorted[j];
                        const float ex = q.x - me.x, ey = q.y - me.y, ez = q.z - me.z;
                        const float d = ex * ex + ey * ey + ez * ez;
                        if (d < b2) {
                            if (d < b1) {
                                b2 = b1;
                                if (d < b0) { b1 = b0; b0 = d; } else b1 = d;
                            } else b2 = d;
                        }
                    }
                }
            }
        }
        const float covered = (float)r * g.h * 0.999f;        // everything closer than this has been visited
        if (b2 <= covered * covered) break;
    }
    float sum = 0.0f;
    int n = 0;
    if (b0 < FLT_MAX) { sum += b0; ++n; }
    if (b1 < FLT_MAX) { sum += b1; ++n; }
    if (b2 < FLT_MAX) { sum += b2; ++n; }
    out[__float_as_uint(me.w)] = n == 3 ? (b0 + b1 + b2) / 3.0f : (n ? sum / (float)n : 0.0f);
}

int launch_knn_mean_dist2(const float *points, int64_t P, float *out, void *scratch, cudaStream_t stream) {
    if (P == 0) return GSB_OK;
    const KnnBuffers b = carve_knn(scratch, P, nullptr);
    const uint32_t C = (uint32_t)knn_cells(P), n = (uint32_t)P;
    const uint32_t blocks = (uint32_t)ceil_div(P, KNN_THREADS);
    cudaError_t err = cudaMemsetAsync(b.bbox, 0xFF, 3 * sizeof(uint32_t), stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(b.bbox + 3, 0, 3 * sizeof(uint32_t), stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(b.counts, 0, ((size_t)C + 1) * sizeof(uint32_t), stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(b.cursor, 0, (size_t)C * sizeof(uint32_t), stream);
    if (err != cudaSuccess) { set_error("knn: memset failed: %s", cudaGetErrorString(err)); return GSB_ERR_CUDA; }
    GSB_LAUNCH("knn_bbox", false, stream, knn_bbox_kernel, blocks < 1184u ? blocks : 1184u, KNN_THREADS, 0, points, n, b.bbox);
    GSB_LAUNCH("knn_grid", false, stream, knn_grid_kernel, 1, 1, 0, b.bbox, C, b.grid);
    GSB_LAUNCH("knn_count", false, stream, knn_count_kernel, blocks, KNN_THREADS, 0, points, n, b.grid, b.cell_of, b.counts);
    const int e = scan_u32_exclusive(b.counts, b.counts, (size_t)C + 1, b.partials, b.total, stream);
    if (e) return e;
    GSB_LAUNCH("knn_scatter", false, stream, knn_scatter_kernel, blocks, KNN_THREADS, 0, points, n, b.cell_of, b.counts, b.cursor, b.sorted);
    GSB_LAUNCH("knn_query", false, stream, knn_query_kernel, blocks, KNN_THREADS, 0, b.sorted, n, b.grid, b.counts, out);
    return GSB_OK;
}

}  // namespace gsb
