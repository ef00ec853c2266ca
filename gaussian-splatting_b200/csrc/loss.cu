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

// ------------------------------------------------------------------------------------------------------------------
// Photometric loss of the reference training step, fused with its gradient:
//     loss = (1 - lambda) * mean|x - y| + lambda * (1 - SSIM(x, y)),   x = clamp(image, 0, 1)
// (/root/reference/train.py:120-126, utils/loss_utils.py:40-86: 11x11 gaussian window, sigma 1.5, zero padding,
// C1 = 0.01^2, C2 = 0.03^2; restated and pinned in oracle/torch_oracle.py photometric_loss).  SURVEY.md section 8(f) #2.
//
// Pass A (ssim_maps_kernel): per 16x16 tile and channel, the five windowed moments by a separable 11-tap convolution in
// shared memory, the SSIM value m and the three partials the backward needs:
//     M1 = dm/dmu1 (total, with sigma's dependence on mu1), M2 = dm/dE[x^2], M3 = dm/dE[xy].
// Pass B (ssim_grad_kernel): dSSIM/dx(q) = conv(M1)(q) + 2 x(q) conv(M2)(q) + y(q) conv(M3)(q) (the window is symmetric),
// combined with the L1 term and clamp's gradient mask into dL/d(image).
// ------------------------------------------------------------------------------------------------------------------
__device__ constexpr float SSIM_W[11] = {1.0283801239e-03f, 7.5987582095e-03f, 3.6000773311e-02f, 1.0936068743e-01f,
                                         2.1300552785e-01f, 2.6601171494e-01f, 2.1300552785e-01f, 1.0936068743e-01f,
                                         3.6000773311e-02f, 7.5987582095e-03f, 1.0283801239e-03f};
constexpr int SSIM_T = 16, SSIM_R = 5, SSIM_S = SSIM_T + 2 * SSIM_R;   // tile, radius, tile + halo = 26
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

__global__ void __launch_bounds__(SSIM_T * SSIM_T)
ssim_maps_kernel(const float *__restrict__ img, const float *__restrict__ gt, const int H, const int W, float *__restrict__ M1,
                 float *__restrict__ M2, float *__restrict__ M3, float *loss_accum, const float w_l1, const float w_ssim) {
    __shared__ float sx[SSIM_S][SSIM_S + 1], sy[SSIM_S][SSIM_S + 1];
    __shared__ float h[5][SSIM_S][SSIM_T + 1];
    __shared__ float red[2][8];
    const int tx = threadIdx.x % SSIM_T, ty = threadIdx.x / SSIM_T;
    const int x0 = blockIdx.x * SSIM_T, y0 = blockIdx.y * SSIM_T, c = blockIdx.z;
    const size_t plane = (size_t)c * H * W;
    for (int i = threadIdx.x; i < SSIM_S * SSIM_S; i += SSIM_T * SSIM_T) {
        const int ly = i / SSIM_S, lx = i % SSIM_S;
        const int gx = x0 + lx - SSIM_R, gy = y0 + ly - SSIM_R;
        float a = 0.f, b = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            a = fminf(fmaxf(img[plane + (size_t)gy * W + gx], 0.0f), 1.0f);
            b = __ldg(gt + plane + (size_t)gy * W + gx);
        }
        sx[ly][lx] = a; sy[ly][lx] = b;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SSIM_S * SSIM_T; i += SSIM_T * SSIM_T) {   // horizontal pass
        const int ly = i / SSIM_T, lx = i % SSIM_T;
        float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float a = sx[ly][lx + k], b = sy[ly][lx + k], w = SSIM_W[k];
            m1 += w * a; m2 += w * b; xx += w * a * a; yy += w * b * b; xy += w * a * b;
        }
        h[0][ly][lx] = m1; h[1][ly][lx] = m2; h[2][ly][lx] = xx; h[3][ly][lx] = yy; h[4][ly][lx] = xy;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, ex2 = 0.f, ey2 = 0.f, exy = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {   // vertical pass
        const float w = SSIM_W[k];
        mu1 += w * h[0][ty + k][tx]; mu2 += w * h[1][ty + k][tx]; ex2 += w * h[2][ty + k][tx];
        ey2 += w * h[3][ty + k][tx]; exy += w * h[4][ty + k][tx];
    }
    const int gx = x0 + tx, gy = y0 + ty;
    float m = 0.f, l1 = 0.f;
    if (gx < W && gy < H) {
        const float s1 = ex2 - mu1 * mu1, s2 = ey2 - mu2 * mu2, s12 = exy - mu1 * mu2;
        const float A = mu1 * mu1 + mu2 * mu2 + SSIM_C1, B = s1 + s2 + SSIM_C2;
        const float Cc = 2.f * mu1 * mu2 + SSIM_C1, D = 2.f * s12 + SSIM_C2;
        const float iAB = 1.0f / (A * B);
        m = Cc * D * iAB;
        const float dm_ds1 = -m / B;              // = -C D / (A B^2)
        const float dm_ds12 = 2.f * Cc * iAB;
        const float dm_dmu1 = 2.f * mu2 * D * iAB - 2.f * mu1 * m / A + dm_ds1 * (-2.f * mu1) + dm_ds12 * (-mu2);
        const size_t pid = plane + (size_t)gy * W + gx;
        M1[pid] = dm_dmu1; M2[pid] = dm_ds1; M3[pid] = dm_ds12;
        l1 = fabsf(sx[ty + SSIM_R][tx + SSIM_R] - sy[ty + SSIM_R][tx + SSIM_R]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m += __shfl_xor_sync(0xffffffffu, m, o); l1 += __shfl_xor_sync(0xffffffffu, l1, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = m; red[1][threadIdx.x >> 5] = l1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sm = 0.f, sl = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { sm += red[0][k]; sl += red[1][k]; }
        // loss_accum[0] total, [1] mean |x - y|, [2] mean SSIM
        atomicAdd(loss_accum + 0, w_l1 * sl - w_ssim * sm);
        atomicAdd(loss_accum + 1, sl * (w_l1 > 0.f ? 1.0f : 0.0f));
        atomicAdd(loss_accum + 2, sm);
    }
}

__global__ void __launch_bounds__(SSIM_T * SSIM_T)
ssim_grad_kernel(const float *__restrict__ img, const float *__restrict__ gt, const int H, const int W,
                 const float *__restrict__ M1, const float *__restrict__ M2, const float *__restrict__ M3,
                 float *__restrict__ grad, const float w_l1, const float w_ssim) {
    __shared__ float s[3][SSIM_S][SSIM_S + 1];
    __shared__ float h[3][SSIM_S][SSIM_T + 1];
    const int tx = threadIdx.x % SSIM_T, ty = threadIdx.x / SSIM_T;
    const int x0 = blockIdx.x * SSIM_T, y0 = blockIdx.y * SSIM_T, c = blockIdx.z;
    const size_t plane = (size_t)c * H * W;
    for (int i = threadIdx.x; i < SSIM_S * SSIM_S; i += SSIM_T * SSIM_T) {
        const int ly = i / SSIM_S, lx = i % SSIM_S;
        const int gx = x0 + lx - SSIM_R, gy = y0 + ly - SSIM_R;
        float a = 0.f, b = 0.f, d = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t pid = plane + (size_t)gy * W + gx;
            a = __ldg(M1 + pid); b = __ldg(M2 + pid); d = __ldg(M3 + pid);
        }
        s[0][ly][lx] = a; s[1][ly][lx] = b; s[2][ly][lx] = d;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SSIM_S * SSIM_T; i += SSIM_T * SSIM_T) {
        const int ly = i / SSIM_T, lx = i % SSIM_T;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = SSIM_W[k];
            a += w * s[0][ly][lx + k]; b += w * s[1][ly][lx + k]; d += w * s[2][ly][lx + k];
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = d;
    }
    __syncthreads();
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) return;
    float g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const float w = SSIM_W[k];
        g1 += w * h[0][ty + k][tx]; g2 += w * h[1][ty + k][tx]; g3 += w * h[2][ty + k][tx];
    }
    const size_t pid = plane + (size_t)gy * W + gx;
    const float raw = img[pid], y = __ldg(gt + pid);
    const float x = fminf(fmaxf(raw, 0.0f), 1.0f);
    const float dssim = g1 + 2.f * x * g2 + y * g3;
    const float d = x - y;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    const float gx_ = w_l1 * sgn - w_ssim * dssim;
    grad[pid] = (raw >= 0.0f && raw <= 1.0f) ? gx_ : 0.0f;
}

int launch_photometric_loss_grad(const float *img, const float *gt, int C, int H, int W, float lambda_dssim, float *grad,
                                 float *loss_accum, float *maps, cudaStream_t stream) {
    if (C <= 0 || H <= 0 || W <= 0) return GSB_OK;
    const size_t n = (size_t)C * H * W;
    const float w_l1 = (1.0f - lambda_dssim) / (float)n, w_ssim = lambda_dssim / (float)n;
    float *M1 = maps, *M2 = maps + n, *M3 = maps + 2 * n;
    const dim3 grid((W + SSIM_T - 1) / SSIM_T, (H + SSIM_T - 1) / SSIM_T, C);
    GSB_LAUNCH("ssim_maps", false, stream, ssim_maps_kernel, grid, SSIM_T * SSIM_T, 0, img, gt, H, W, M1, M2, M3, loss_accum, w_l1,
               w_ssim);
    GSB_LAUNCH("ssim_grad", false, stream, ssim_grad_kernel, grid, SSIM_T * SSIM_T, 0, img, gt, H, W, M1, M2, M3, grad, w_l1, w_ssim);
    return GSB_OK;
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
