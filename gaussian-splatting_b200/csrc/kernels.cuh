// kernels.cuh -- argument blocks and host launchers of the rasterizer kernels.
#pragma once
#include "common.cuh"
#include "geom.cuh"

namespace gsb {

// per-gaussian gradient accumulator written by the backward blend kernel (float, 48 B stride):
//   0,1 mean2D (NDC-scaled)  2,3,4 conic A,B,C  5 opacity  6,7,8 rgb  9 inverse depth  10,11 pad
constexpr int DACC_STRIDE = 12;

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

struct PreBwdArgs {
    int P;
    const float *means, *shs, *opac, *scales, *rots, *cov_pre;
    const float4 *splat;
    const float *dacc;    // [P*DACC_STRIDE]
    int moments;          // dacc[0..4] hold raw moments of d(power) (render_mp.cu) instead of mean/conic gradients
    GsbGrads g;
};

struct BinArgs {
    int P;
    int num_tiles, gx;
    const uint32_t *order;    // [P] gaussian ids in depth order
    const uint32_t *tiles;    // [P]
    const uint2 *rect;        // [P]
    const float4 *splat;
    uint32_t *offsets;        // [P] exclusive scan of tiles in depth order
    uint32_t *partials;       // scan scratch
    unsigned long long *total;  // device scalar: D
};

struct RenderFwdArgs {
    int W, H, gx, gy;
    const uint2 *ranges;
    const uint32_t *point_list;
    const float4 *splat;
    const float *bg;
    float *out_color, *out_invdepth, *final_T;
    uint32_t *n_contrib;
};

struct RenderBwdArgs {
    int W, H, gx, gy;
    const uint2 *ranges;
    const uint32_t *point_list;
    const float4 *splat;
    const float *bg;
    const float *final_T;
    const uint32_t *n_contrib;
    const float *out_color, *out_invdepth;   // forward outputs
    const float *dL_dcolor, *dL_dinvdepth;
    float *dacc;
};

int launch_preprocess_fwd(const CamArgs &ca, const PreFwdArgs &a, bool debug, cudaStream_t stream);
int launch_preprocess_bwd(const CamArgs &ca, const PreBwdArgs &a, bool accumulate, bool debug, cudaStream_t stream);
int launch_mark_visible(int P, const float *means, const float *view, uint8_t *present, cudaStream_t stream);

size_t scan_partials_count(int P);
int launch_tile_scan(const BinArgs &a, bool debug, cudaStream_t stream);
int launch_emit(const BinArgs &a, uint32_t *inst_tile, uint32_t *inst_gauss, int64_t cap, bool debug, cudaStream_t stream);
int launch_tile_ranges(const uint32_t *sorted_tiles, int64_t D, const unsigned long long *n_dev, int num_tiles,
                       uint2 *ranges, bool debug, cudaStream_t stream);

int launch_render_fwd(const RenderFwdArgs &a, int variant, bool debug, cudaStream_t stream);
int launch_render_bwd(const RenderBwdArgs &a, int variant, bool debug, cudaStream_t stream);
int launch_render_fwd_mp(const RenderFwdArgs &a, int qh, bool debug, cudaStream_t stream);
int launch_render_bwd_mp(const RenderBwdArgs &a, int qh, bool debug, cudaStream_t stream);
int launch_l1_loss_grad(const float *img, const float *gt, int64_t n, float scale, float *grad, float *loss_accum,
                        cudaStream_t stream);
int launch_render_fwd_ps(const RenderFwdArgs &a, bool debug, cudaStream_t stream);
int launch_render_bwd_ps(const RenderBwdArgs &a, bool debug, cudaStream_t stream);

}  // namespace gsb
