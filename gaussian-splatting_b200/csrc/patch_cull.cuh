// patch_cull.cuh -- sub-tile culling inside the blend kernels.
//
// A 16x16 tile is cut into eight 8x4-pixel patches (bit 2*band + col).  While a gaussian is staged into shared
// memory its 2 ln(255 o) ellipse is intersected with the patches (same span construction as the tile culling in
// geom.cuh, evaluated per 4-row band), and every warp then walks only the gaussians that can reach its own
// pixels.  Conservative (padded), so the image is unchanged; skipped pairs are pairs the blend loop would have
// rejected one pixel at a time.
#pragma once
#include "common.cuh"

namespace gsb {

// q0 = {x, y, A, B}, C, lim = padded 2 ln(255 o) (record q2.z; > 1e30 when culling is disabled)
__device__ __forceinline__ uint32_t patch_mask(const float cx, const float cy, const float A, const float B, const float C,
                                               const float lim, const float ox, const float oy) {
    if (!(lim < 1.0e30f)) return 0xffu;
    if (lim < 0.0f) return 0u;
    const float detc = A * C - B * B;
    if (!(detc > 0.0f)) return 0xffu;
    const float inv_det = rcp_apx(detc);
    const float ex = sqrt_apx(lim * C * inv_det), ey = sqrt_apx(lim * A * inv_det);
    const float m = fmaf(ex, 0.002f, 0.05f), my = fmaf(ey, 0.002f, 0.05f);
    const float dyR = -B * rcp_apx(C) * ex;
    const float Alim = A * lim, invA = rcp_apx(A);
    const float xa = ox - cx;   // tile origin relative to the mean
    uint32_t mask = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float y0 = oy + 4.0f * r - cy, y1 = y0 + 3.0f;
        const float ylo = fmaxf(y0, -ey - my), yhi = fminf(y1, ey + my);
        if (ylo > yhi) continue;
        const float yr = fminf(fmaxf(dyR, ylo), yhi), yl = fminf(fmaxf(-dyR, ylo), yhi);
        const float xr = (-B * yr + sqrt_apx(fmaxf(Alim - detc * yr * yr, 0.0f))) * invA + m;
        const float xl = (-B * yl - sqrt_apx(fmaxf(Alim - detc * yl * yl, 0.0f))) * invA - m;
        // columns: pixels [xa, xa+7] and [xa+8, xa+15] relative to the mean
        const bool c0 = (xr >= xa) && (xl <= xa + 7.0f);
        const bool c1 = (xr >= xa + 8.0f) && (xl <= xa + 15.0f);
        mask |= (c0 ? 1u : 0u) << (2 * r);
        mask |= (c1 ? 2u : 0u) << (2 * r);
    }
    return mask;
}

// Warp-private compaction: indices k < n of the staged batch whose mask meets `want`, in order, into list[]
// (uint8 indices; the batch holds at most 256 gaussians).  Returns the count.  All 32 lanes must call.
__device__ __forceinline__ int compact_hits(const uint8_t *smask, const int n, const uint32_t want, uint8_t *list) {
    const int lane = threadIdx.x & 31;
    int cnt = 0;
    for (int c = 0; c < n; c += 32) {
        const int k = c + lane;
        const bool hit = (k < n) && ((smask[k] & want) != 0u);
        const uint32_t b = __ballot_sync(0xffffffffu, hit);
        if (hit) list[cnt + __popc(b & ((1u << lane) - 1u))] = (uint8_t)k;
        cnt += __popc(b);
    }
    __syncwarp();
    return cnt;
}

}  // namespace gsb
