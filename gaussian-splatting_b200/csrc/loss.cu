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
// 32x32-pixel tile per 256-thread block, halo 5: 42x42 inputs.  Both separable passes use register sliding windows: a
// thread produces FOUR adjacent outputs from 14 loaded taps, so shared-memory loads per output drop from 11 to 3.5
// (the first version -- one output per thread per pass -- ran at 40 % of its instruction-count floor, LSU bound).
constexpr int SSIM_TX = 32, SSIM_TY = 32, SSIM_R = 5, SSIM_SX = SSIM_TX + 2 * SSIM_R, SSIM_SY = SSIM_TY + 2 * SSIM_R, SSIM_NT = 256;
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

__global__ void __launch_bounds__(SSIM_NT)
ssim_maps_kernel(const float *__restrict__ img, const float *__restrict__ gt, const int H, const int W, float *__restrict__ M1,
                 float *__restrict__ M2, float *__restrict__ M3, float *loss_accum, const float w_l1, const float w_ssim,
                 const float lo, const float hi) {
    __shared__ float sx[SSIM_SY][SSIM_SX + 1], sy[SSIM_SY][SSIM_SX + 1];
    __shared__ float h[5][SSIM_SY][SSIM_TX + 1];
    __shared__ float red[2][8];
    const int x0 = blockIdx.x * SSIM_TX, y0 = blockIdx.y * SSIM_TY, c = blockIdx.z;
    const size_t plane = (size_t)c * H * W;
    for (int i = threadIdx.x; i < SSIM_SY * SSIM_SX; i += SSIM_NT) {
        const int ly = i / SSIM_SX, lx = i % SSIM_SX;
        const int gx = x0 + lx - SSIM_R, gy = y0 + ly - SSIM_R;
        float a = 0.f, b = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            a = fminf(fmaxf(img[plane + (size_t)gy * W + gx], lo), hi);     // [0, 1], or the whole float range (no clamp)
            b = __ldg(gt + plane + (size_t)gy * W + gx);
        }
        sx[ly][lx] = a; sy[ly][lx] = b;
    }
    __syncthreads();
    // horizontal pass: task = (row, 4-column segment)
    for (int task = threadIdx.x; task < SSIM_SY * (SSIM_TX / 4); task += SSIM_NT) {
        const int ly = task / (SSIM_TX / 4), lx0 = (task % (SSIM_TX / 4)) * 4;
        float a[14], b[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) { a[j] = sx[ly][lx0 + j]; b[j] = sy[ly][lx0 + j]; }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float w = SSIM_W[k], wa = w * a[o + k], wb = w * b[o + k];
                m1 += wa; m2 += wb; xx += wa * a[o + k]; yy += wb * b[o + k]; xy += wa * b[o + k];
            }
            h[0][ly][lx0 + o] = m1; h[1][ly][lx0 + o] = m2; h[2][ly][lx0 + o] = xx; h[3][ly][lx0 + o] = yy; h[4][ly][lx0 + o] = xy;
        }
    }
    __syncthreads();
    // vertical pass: thread = (column, 4-row segment)
    const int tx = threadIdx.x % SSIM_TX, ty0 = (threadIdx.x / SSIM_TX) * 4;
    float q[5][4];
#pragma unroll
    for (int qq = 0; qq < 5; ++qq) {
        float v[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) v[j] = h[qq][ty0 + j][tx];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) acc += SSIM_W[k] * v[o + k];
            q[qq][o] = acc;
        }
    }
    float msum = 0.f, l1sum = 0.f;
    const int gx = x0 + tx;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const int gy = y0 + ty0 + o;
        if (gx < W && gy < H) {
            const float mu1 = q[0][o], mu2 = q[1][o];
            const float s1 = q[2][o] - mu1 * mu1, s2 = q[3][o] - mu2 * mu2, s12 = q[4][o] - mu1 * mu2;
            const float A = mu1 * mu1 + mu2 * mu2 + SSIM_C1, B = s1 + s2 + SSIM_C2;
            const float Cc = 2.f * mu1 * mu2 + SSIM_C1, D = 2.f * s12 + SSIM_C2;
            const float iAB = 1.0f / (A * B);
            const float m = Cc * D * iAB;
            const float dm_ds1 = -m / B;              // = -C D / (A B^2)
            const float dm_ds12 = 2.f * Cc * iAB;
            const float dm_dmu1 = 2.f * mu2 * D * iAB - 2.f * mu1 * m / A + dm_ds1 * (-2.f * mu1) + dm_ds12 * (-mu2);
            const size_t pid = plane + (size_t)gy * W + gx;
            M1[pid] = dm_dmu1; M2[pid] = dm_ds1; M3[pid] = dm_ds12;
            msum += m;
            l1sum += fabsf(sx[ty0 + o + SSIM_R][tx + SSIM_R] - sy[ty0 + o + SSIM_R][tx + SSIM_R]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { msum += __shfl_xor_sync(0xffffffffu, msum, o); l1sum += __shfl_xor_sync(0xffffffffu, l1sum, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = msum; red[1][threadIdx.x >> 5] = l1sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sm = 0.f, sl = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { sm += red[0][k]; sl += red[1][k]; }
        // loss_accum[0] total, [1] sum |x - y|, [2] sum of the SSIM map
        atomicAdd(loss_accum + 0, w_l1 * sl - w_ssim * sm);
        atomicAdd(loss_accum + 1, sl);
        atomicAdd(loss_accum + 2, sm);
    }
}

__global__ void __launch_bounds__(SSIM_NT)
ssim_grad_kernel(const float *__restrict__ img, const float *__restrict__ gt, const int H, const int W,
                 const float *__restrict__ M1, const float *__restrict__ M2, const float *__restrict__ M3,
                 float *__restrict__ grad, const float w_l1, const float w_ssim, const float lo, const float hi) {
    __shared__ float s[3][SSIM_SY][SSIM_SX + 1];
    __shared__ float h[3][SSIM_SY][SSIM_TX + 1];
    const int x0 = blockIdx.x * SSIM_TX, y0 = blockIdx.y * SSIM_TY, c = blockIdx.z;
    const size_t plane = (size_t)c * H * W;
    for (int i = threadIdx.x; i < SSIM_SY * SSIM_SX; i += SSIM_NT) {
        const int ly = i / SSIM_SX, lx = i % SSIM_SX;
        const int gx = x0 + lx - SSIM_R, gy = y0 + ly - SSIM_R;
        float a = 0.f, b = 0.f, d = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t pid = plane + (size_t)gy * W + gx;
            a = __ldg(M1 + pid); b = __ldg(M2 + pid); d = __ldg(M3 + pid);
        }
        s[0][ly][lx] = a; s[1][ly][lx] = b; s[2][ly][lx] = d;
    }
    __syncthreads();
    for (int task = threadIdx.x; task < SSIM_SY * (SSIM_TX / 4); task += SSIM_NT) {
        const int ly = task / (SSIM_TX / 4), lx0 = (task % (SSIM_TX / 4)) * 4;
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {
            float v[14];
#pragma unroll
            for (int j = 0; j < 14; ++j) v[j] = s[qq][ly][lx0 + j];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) acc += SSIM_W[k] * v[o + k];
                h[qq][ly][lx0 + o] = acc;
            }
        }
    }
    __syncthreads();
    const int tx = threadIdx.x % SSIM_TX, ty0 = (threadIdx.x / SSIM_TX) * 4;
    float g[3][4];
#pragma unroll
    for (int qq = 0; qq < 3; ++qq) {
        float v[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) v[j] = h[qq][ty0 + j][tx];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) acc += SSIM_W[k] * v[o + k];
            g[qq][o] = acc;
        }
    }
    const int gx = x0 + tx;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const int gy = y0 + ty0 + o;
        if (gx >= W || gy >= H) continue;
        const size_t pid = plane + (size_t)gy * W + gx;
        const float raw = img[pid], y = __ldg(gt + pid);
        const float x = fminf(fmaxf(raw, lo), hi);
        const float dssim = g[0][o] + 2.f * x * g[1][o] + y * g[2][o];
        const float d = x - y;
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        const float gv = w_l1 * sgn - w_ssim * dssim;
        grad[pid] = (raw >= lo && raw <= hi) ? gv : 0.0f;
    }
}

int launch_photometric_loss_grad(const float *img, const float *gt, int C, int H, int W, float lambda_dssim, bool clamp_input,
                                 float *grad, float *loss_accum, float *maps, cudaStream_t stream) {
    if (C <= 0 || H <= 0 || W <= 0) return GSB_OK;
    const size_t n = (size_t)C * H * W;
    const float w_l1 = (1.0f - lambda_dssim) / (float)n, w_ssim = lambda_dssim / (float)n;
    float *M1 = maps, *M2 = maps + n, *M3 = maps + 2 * n;
    const float lo = clamp_input ? 0.0f : -3.402823466e+38f, hi = clamp_input ? 1.0f : 3.402823466e+38f;
    const dim3 grid((W + SSIM_TX - 1) / SSIM_TX, (H + SSIM_TY - 1) / SSIM_TY, C);
    GSB_LAUNCH("ssim_maps", false, stream, ssim_maps_kernel, grid, SSIM_NT, 0, img, gt, H, W, M1, M2, M3, loss_accum, w_l1, w_ssim, lo, hi);
    GSB_LAUNCH("ssim_grad", false, stream, ssim_grad_kernel, grid, SSIM_NT, 0, img, gt, H, W, M1, M2, M3, grad, w_l1, w_ssim, lo, hi);
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
