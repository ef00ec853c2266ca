// kernels.cuh -- argument blocks and host launchers of the rasterizer kernels.
#pragma once
#include "common.cuh"
#include "geom.cuh"

namespace gsb {

// per-gaussian gradient accumulator written by the backward blend kernel (float, 48 B stride):
//   0..4 raw moments of t = dL/d(power): sum t dx, t dy, t dx^2, t dx dy, t dy^2   5 opacity  6,7,8 rgb  9 inverse depth  10,11 pad
constexpr int DACC_STRIDE = 12;

// ---- view-batch variants: per-view arrays are laid out [V][...] with uniform strides (in elements) ----
constexpr int GSB_MAX_VIEWS = 16;
constexpr int GSB_MAX_PEERS = 16;
struct CamArgsBatch {
    int V;
    CamArgs cam[GSB_MAX_VIEWS];
};
struct PreFwdBatchStrides { size_t splat /* float4 */, per_gauss /* u32 / uint2 arrays */, radii; };
struct PreBwdBatchStrides { size_t splat /* float4 */, dacc /* float */, means2D /* float */; };

struct PreFwdArgs {
    int P;
    const float *means, *shs, *colors, *opac, *scales, *rots, *cov_pre;
    float4 *splat;        // [P*3]
    uint32_t *depth_key;  // [P] float bits of view depth, 0xffffffff when culled
    uint32_t *depth_idx;  // [P] identity, sorted along with the keys
    uint32_t *tiles;      // [P] tiles touched after exact culling
    uint2 *rect;          // [P] x0|x1<<16, y0|y1<<16 reference 3-sigma tile rectangle
    int *radii;           // [P] output
    int cull;             // 1: exact tile culling, 0: reference rectangle (debug / A-B)
};

// Fused reduce-scatter (gsb_backward_batch_peer): gaussians [r * rows_per_rank, (r + 1) * rows_per_rank) are OWNED by rank r;
// every rank adds its gradient rows straight into the owner's buffer.  delta[r] = byte offset that turns an address inside THIS
// rank's gradient buffer into the same location of rank r's buffer as mapped into this process (0 for r == rank).
struct PeerTable {
    int world;              // 0 = no peers (plain local output)
    int rows_per_rank;      // multiple of the per-gaussian kernels' block size
    long long delta[GSB_MAX_PEERS];
};

struct PreBwdArgs {
    int P;
    const float *means, *shs, *opac, *scales, *rots, *cov_pre;
    const float4 *splat;
    const float *dacc;    // [P*DACC_STRIDE]
    int p_begin, p_end;   // gaussian range of this launch (chunked backward: the caller reduces finished chunks meanwhile)
    GsbGrads g;
    PeerTable peer;       // world > 0: outputs are ADDED into the owners' buffers (view-batch kernel, TMA rows only)
};

struct BinArgs {
    int P;
    int num_tiles, gx;
    // view-batch launches (blockIdx.y = view): strides between the views' arrays, in elements; 0 for single view
    size_t sv_gauss;          // order / tiles / rect / offsets
    size_t sv_splat;          // float4
    size_t sv_partials;
    size_t sv_inst;           // instance arrays (capacity)
    const uint32_t *order;    // [P] gaussian ids in depth order
    const uint32_t *tiles;    // [P]
    const uint2 *rect;        // [P]
    const float4 *splat;
    uint32_t *offsets;        // [P] exclusive scan of tiles in depth order
    uint32_t *partials;       // scan scratch
    unsigned long long *total;  // device scalar: D
};

// Blend launches cover V views (blockIdx.y): every per-view array is base + view * stride (strides in elements; 0 and V = 1
// for the single-view call).  tile_order (optional, [V][tiles]): blockIdx.x -> tile id permutation (heavy tiles first).
struct RenderFwdArgs {
    int W, H, gx, gy, V;
    size_t sv_ranges, sv_list, sv_splat /* float4 */, sv_color, sv_depth, sv_image;
    const uint2 *ranges;
    const uint32_t *point_list;
    const uint32_t *tile_order;
    const float4 *splat;
    const float *bg[GSB_MAX_VIEWS];
    float *out_color, *out_invdepth, *final_T;
    uint32_t *n_contrib;
};

struct RenderBwdArgs {
    int W, H, gx, gy, V;
    size_t sv_ranges, sv_list, sv_splat /* float4 */, sv_color, sv_depth, sv_image, sv_dacc;
    const uint2 *ranges;
    const uint32_t *point_list;
    const uint32_t *tile_order;
    const float4 *splat;
    const uint32_t *n_contrib;
    const float *out_color, *out_invdepth;   // forward outputs
    const float *dL_dcolor, *dL_dinvdepth;
    float *dacc;
};

int launch_preprocess_fwd(const CamArgs &ca, const PreFwdArgs &a, bool debug, cudaStream_t stream);
int launch_preprocess_bwd(const CamArgs &ca, const PreBwdArgs &a, bool accumulate, bool debug, cudaStream_t stream);
int launch_preprocess_fwd_batch(const CamArgsBatch &cb, const PreFwdArgs &a, const PreFwdBatchStrides &st, bool debug, cudaStream_t stream);
int launch_preprocess_bwd_batch(const CamArgsBatch &cb, const PreBwdArgs &a, const PreBwdBatchStrides &st, bool accumulate, bool debug,
                                cudaStream_t stream);
int launch_mark_visible(int P, const float *means, const float *view, uint8_t *present, cudaStream_t stream);

size_t scan_partials_count(int P);
// V > 1: blockIdx.y = view, arrays offset by the BinArgs strides; totals / n_dev are arrays of V counts
int launch_tile_scan(const BinArgs &a, int V, bool debug, cudaStream_t stream);
// async forward: counts[GSB_MAX_VIEWS] = max(counts[GSB_MAX_VIEWS], counts[0..V))  (running maximum the host polls later)
int launch_count_max(unsigned long long *counts, int V, bool debug, cudaStream_t stream);
int launch_emit(const BinArgs &a, int V, uint32_t *inst_tile, uint32_t *inst_gauss, int64_t cap, bool debug, cudaStream_t stream);
int launch_tile_ranges(const uint32_t *sorted_tiles, int64_t D, const unsigned long long *n_dev, int num_tiles,
                       uint2 *ranges, int V, size_t sv_inst, bool debug, cudaStream_t stream);

int launch_render_fwd(const RenderFwdArgs &a, bool debug, cudaStream_t stream);
int launch_render_bwd(const RenderBwdArgs &a, bool debug, cudaStream_t stream);
int launch_tile_order(const uint2 *ranges, int V, int num_tiles, uint32_t *order, bool debug, cudaStream_t stream);
int launch_l1_loss_grad(const float *img, const float *gt, int64_t n, float scale, float *grad, float *loss_accum,
                        cudaStream_t stream);
int launch_photometric_loss_grad(const float *img, const float *gt, int C, int H, int W, float lambda_dssim, bool clamp_input,
                                 float *grad, float *loss_accum, float *maps, cudaStream_t stream);

// optimizer step and densification on the flat store (optim.cu, densify.cu)
int launch_adam_step(int64_t P, int sh_coeffs, float *params, const float *grads, float *m, float *v, float *act,
                     const uint8_t *visible, const float step_size[6], float beta1, float beta2, float eps, const float bias2_sqrt[6],
                     uint32_t skip_groups, cudaStream_t stream);
int launch_activate(int64_t P, int sh_coeffs, const float *params, float *act, cudaStream_t stream);
size_t densify_scratch_bytes(int64_t P, int n_children);
// exclusive scan of n uint32 (in == out allowed); partials: scan_u32_partials(n) words; *total receives the sum
size_t scan_u32_partials(int64_t n);
int scan_u32_exclusive(const uint32_t *in, uint32_t *out, size_t n, uint32_t *partials, uint32_t *total, cudaStream_t stream);
// 3-nearest-neighbour mean squared distance (knn.cu)
size_t knn_scratch_bytes(int64_t P);
int launch_knn_mean_dist2(const float *points, int64_t P, float *out, void *scratch, cudaStream_t stream);
int launch_densify_plan(int64_t P, int sh_coeffs, int n_children, const float *params, const float *grad_accum, const float *denom,
                        float grad_threshold, float size_limit, float min_opacity, float world_limit, void *scratch,
                        uint32_t **counters_dev, cudaStream_t stream);
int launch_densify_apply(int64_t P, int sh_coeffs, int n_children, const float *params, const float *m, const float *v, void *scratch,
                         const float *unit_samples, int64_t n_split, int64_t P_new, float *new_params, float *new_m, float *new_v,
                         cudaStream_t stream);

}  // namespace gsb
