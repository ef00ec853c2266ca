// loss.cu -- fused L1 photometric loss + its gradient w.r.t. the UNCLAMPED rendered image.
//
// The step right after the hot path in the reference loop (train.py:120-126: Ll1 = l1_loss(image, gt_image), with
// render() clamping the image to [0,1] first, gaussian_renderer/__init__.py:119) costs eight elementwise torch
// kernels forward + backward.  One pass here: loss += scale * sum |clamp(img) - gt|,
// grad = scale * sign(clamp(img) - gt) * [0 <= img <= 1]  (torch.clamp's and torch.abs's gradients).
#include "common.cuh"

namespace gsb {

__global__ void __launch_bounds__(256)
l1_loss_grad_kernel(const float4 *__restrict__ img, const float4 *__restrict__ gt, const int64_t n4, const float scale,
                    float4 *__restrict__ grad, float *loss_accum) {
    __shared__ float warp_part[8];
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
        const float4 a = img[i], b = __ldg(gt + i);
        float4 g;
        const float va[4] = {a.x, a.y, a.z, a.w}, vb[4] = {b.x, b.y, b.z, b.w};
        float vg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float c = fminf(fmaxf(va[k], 0.0f), 1.0f);
            const float d = c - vb[k];
            acc += fabsf(d);
            const float s = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
            vg[k] = (va[k] >= 0.0f && va[k] <= 1.0f) ? s : 0.0f;
        }
        g.x = vg[0]; g.y = vg[1]; g.z = vg[2]; g.w = vg[3];
        grad[i] = g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += warp_part[k];
        atomicAdd(loss_accum, t * scale);
    }
}

int launch_l1_loss_grad(const float *img, const float *gt, int64_t n, float scale, float *grad, float *loss_accum,
                        cudaStream_t stream) {
    if (n <= 0) return GSB_OK;
    if (n % 4 != 0) { set_error("l1_loss_grad: element count must be a multiple of 4"); return GSB_ERR_ARGUMENT; }
    const int64_t n4 = n / 4;
    const int grid = (int)(ceil_div(n4, 256) < 148 * 8 ? ceil_div(n4, 256) : 148 * 8);
    GSB_LAUNCH("l1_loss_grad", false, stream, l1_loss_grad_kernel, grid, 256, 0, reinterpret_cast<const float4 *>(img),
               reinterpret_cast<const float4 *>(gt), n4, scale, reinterpret_cast<float4 *>(grad), loss_accum);
    return GSB_OK;
}

}  // namespace gsb
