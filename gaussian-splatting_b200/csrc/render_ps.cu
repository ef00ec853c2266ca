// render_ps.cu -- "patch-slot" blend kernels (EXPERIMENT, measured slower, NOT the default; fwd/bwd variant 5).
// Kept selectable through gsb_set_option so the A/B in DESIGN.md section 4 can be reproduced (tests/analysis/sweep.py):
// forward 0.49 ms vs 0.33, backward 0.77 ms vs 0.73 at the time -- the per-slot warp-uniform branches serialise the
// four slots and remove the instruction-level parallelism the 2x2 kernels get from evaluating them together.
//
// A 16x16 tile is one CTA of two warps.  Warp w owns rows [8w, 8w+8) = four 8x4 patches; lane l owns the SAME
// in-patch position (l & 7, l >> 3) in each of the four patches, i.e. four pixels 8 columns / 4 rows apart
// ("slots").  This keeps what profiles/r1_render_*.md showed to matter on this issue-bound path:
//   * per-gaussian overhead (shared-memory reads, loop, and in the backward pass the cross-lane reduction of
//     the ten gradient terms) is paid once per FOUR pixels per lane;
//   * culling stays at 8x4-patch granularity: the 8-bit patch mask computed while a gaussian is staged
//     (patch_cull.cuh) tells each warp which of its slots the gaussian can reach at all, so a slot is skipped
//     with a warp-uniform branch before any per-pixel arithmetic;
//   * dx / dy terms are shared between slots (columns differ by 8, rows by 4).
// Arithmetic (log2-domain exponent, front-to-back backward with (T, F) state, moment accumulation) is that of
// render_mp.cu; see its header for the derivation.
#include "blend_common.cuh"

namespace gsb {

constexpr int PS_R = 128;   // gaussians staged per round
constexpr int PS_NT = 64;

__global__ void __launch_bounds__(PS_NT)
render_fwd_ps_kernel(const RenderFwdArgs a) {
    __shared__ float4 s0[PS_R], s1[PS_R];
    __shared__ float2 s2[PS_R];
    __shared__ uint8_t smask[PS_R];
    __shared__ uint8_t slist[2][PS_R];
    const int tile = blockIdx.x;
    const int ox = (tile % a.gx) * TILE, oy = (tile / a.gx) * TILE;
    const int t = threadIdx.x, w = t >> 5, l = t & 31;
    const int px0 = ox + (l & 7), py0 = oy + 8 * w + (l >> 3);
    const float fx0 = (float)px0, fy0 = (float)py0;
    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);

    float T[4], C0[4], C1[4], C2[4], Dp[4];
    uint32_t last[4];
    bool done[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T[i] = 1.0f; C0[i] = C1[i] = C2[i] = Dp[i] = 0.f; last[i] = 0u;
        done[i] = !((px0 + 8 * (i & 1)) < a.W && (py0 + 4 * (i >> 1)) < a.H);
    }

    for (int base = 0; base < todo; base += PS_R) {
        const bool all_done = done[0] && done[1] && done[2] && done[3];
        if (__syncthreads_and(all_done)) break;
        const int n = min(PS_R, todo - base);
        for (int k = t; k < n; k += PS_NT) {
            const uint32_t g = a.point_list[range.x + base + k];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            const float4 q2 = __ldg(rec + 2);
            smask[k] = (uint8_t)patch_mask(q0.x, q0.y, q0.z, q0.w, q1.x, q2.z, (float)ox, (float)oy);
            stage_scale(q0, q1);
            s0[k] = q0; s1[k] = q1; s2[k] = make_float2(q2.x, q2.y);
        }
        __syncthreads();
        const int cnt = compact_hits(smask, n, 0xfu << (4 * w), slist[w]);
        for (int kk = 0; kk < cnt; ++kk) {
            const int j = slist[w][kk];
            const uint32_t m4 = ((uint32_t)smask[j] >> (4 * w)) & 0xfu;
            const float4 q0 = s0[j];
            const float4 q1 = s1[j];
            float dx[2], Axx[2], Bx[2], dy[2], Cyy[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                dx[c] = q0.x - (fx0 + 8.0f * c);
                Axx[c] = __fmul_rn(__fmul_rn(q0.z, dx[c]), dx[c]);
                Bx[c] = __fmul_rn(q0.w, dx[c]);
                dy[c] = q0.y - (fy0 + 4.0f * c);
                Cyy[c] = __fmul_rn(__fmul_rn(q1.x, dy[c]), dy[c]);
            }
            float al[4];
            bool valid[4];
            bool any = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                valid[i] = false;
                al[i] = 0.f;
                if (m4 & (1u << i)) {   // warp-uniform
                    const float p = power2_at(Axx[i & 1], Cyy[i >> 1], Bx[i & 1], dy[i >> 1]);
                    al[i] = fminf(ALPHA_MAX, __fmul_rn(q1.y, ex2_approx(p)));
                    valid[i] = (p <= 0.0f) && (al[i] >= ALPHA_MIN) && !done[i];
                }
                any = any || valid[i];
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            const float2 q2 = s2[j];
            const uint32_t pos = (uint32_t)(base + j + 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!(m4 & (1u << i))) continue;   // warp-uniform
                const float test_T = __fmul_rn(T[i], __fsub_rn(1.0f, al[i]));
                const bool stop = valid[i] && (test_T < T_STOP);
                done[i] = done[i] || stop;
                const bool use = valid[i] && !stop;
                const float wgt = use ? __fmul_rn(al[i], T[i]) : 0.0f;
                C0[i] = __fmaf_rn(q1.z, wgt, C0[i]);
                C1[i] = __fmaf_rn(q1.w, wgt, C1[i]);
                C2[i] = __fmaf_rn(q2.x, wgt, C2[i]);
                Dp[i] = __fmaf_rn(q2.y, wgt, Dp[i]);
                T[i] = use ? test_T : T[i];
                last[i] = use ? pos : last[i];
            }
        }
    }
    const float bg0 = __ldg(a.bg), bg1 = __ldg(a.bg + 1), bg2 = __ldg(a.bg + 2);
    const size_t HW = (size_t)a.W * a.H;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int px = px0 + 8 * (i & 1), py = py0 + 4 * (i >> 1);
        if (px >= a.W || py >= a.H) continue;
        const size_t pid = (size_t)py * a.W + px;
        a.final_T[pid] = T[i];
        a.n_contrib[pid] = last[i];
        a.out_color[pid] = C0[i] + T[i] * bg0;
        a.out_color[HW + pid] = C1[i] + T[i] * bg1;
        a.out_color[2 * HW + pid] = C2[i] + T[i] * bg2;
        a.out_invdepth[pid] = Dp[i];
    }
}

// dacc layout (raw moments, PreBwdArgs::moments = 1): see render_mp.cu
template <bool DEPTH>
__global__ void __launch_bounds__(PS_NT)
render_bwd_ps_kernel(const RenderBwdArgs a) {
    __shared__ float4 s0[PS_R], s1[PS_R];
    __shared__ float2 s2[PS_R];
    __shared__ uint32_t sid[PS_R];
    __shared__ uint8_t smask[PS_R];
    __shared__ uint8_t slist[2][PS_R];
    __shared__ uint32_t s_max;
    const int tile = blockIdx.x;
    const int ox = (tile % a.gx) * TILE, oy = (tile / a.gx) * TILE;
    const int t = threadIdx.x, w = t >> 5, l = t & 31;
    const int px0 = ox + (l & 7), py0 = oy + 8 * w + (l >> 3);
    const float fx0 = (float)px0, fy0 = (float)py0;
    const uint2 range = a.ranges[tile];
    const size_t HW = (size_t)a.W * a.H;

    float T[4], F[4], S[4], dL0[4], dL1[4], dL2[4], dLd[4];
    uint32_t last[4];
    uint32_t my_max = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int px = px0 + 8 * (i & 1), py = py0 + 4 * (i >> 1);
        T[i] = 1.0f; F[i] = 0.f; S[i] = 0.f; dL0[i] = dL1[i] = dL2[i] = dLd[i] = 0.f; last[i] = 0u;
        if (px < a.W && py < a.H) {
            const size_t pid = (size_t)py * a.W + px;
            last[i] = a.n_contrib[pid];
            dL0[i] = a.dL_dcolor[pid]; dL1[i] = a.dL_dcolor[HW + pid]; dL2[i] = a.dL_dcolor[2 * HW + pid];
            S[i] = dL0[i] * a.out_color[pid] + dL1[i] * a.out_color[HW + pid] + dL2[i] * a.out_color[2 * HW + pid];
            if (DEPTH) {
                dLd[i] = a.dL_dinvdepth[pid];
                S[i] += dLd[i] * a.out_invdepth[pid];
            }
        }
        my_max = max(my_max, last[i]);
    }
    if (t == 0) s_max = 0;
    __syncthreads();
    my_max = __reduce_max_sync(0xffffffffu, my_max);   // the deepest list position this warp's pixels blended
    if (l == 0) atomicMax(&s_max, my_max);
    __syncthreads();
    const int todo = (int)s_max;
    const int my_todo = (int)my_max;

    for (int base = 0; base < todo; base += PS_R) {
        __syncthreads();
        const int n = min(PS_R, todo - base);
        for (int k = t; k < n; k += PS_NT) {
            const uint32_t g = a.point_list[range.x + base + k];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            const float4 q2 = __ldg(rec + 2);
            smask[k] = (uint8_t)patch_mask(q0.x, q0.y, q0.z, q0.w, q1.x, q2.z, (float)ox, (float)oy);
            stage_scale(q0, q1);
            s0[k] = q0; s1[k] = q1; s2[k] = make_float2(q2.x, q2.y); sid[k] = g;
        }
        __syncthreads();
        const int nw = min(n, my_todo - base);   // this warp needs nothing behind its own deepest pixel
        const int cnt = compact_hits(smask, nw, 0xfu << (4 * w), slist[w]);
        for (int kk = 0; kk < cnt; ++kk) {
            const int j = slist[w][kk];
            const uint32_t m4 = ((uint32_t)smask[j] >> (4 * w)) & 0xfu;
            const float4 q0 = s0[j];
            const float4 q1 = s1[j];
            const uint32_t pos = (uint32_t)(base + j + 1);
            float dx[2], Axx[2], Bx[2], dy[2], Cyy[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                dx[c] = q0.x - (fx0 + 8.0f * c);
                Axx[c] = __fmul_rn(__fmul_rn(q0.z, dx[c]), dx[c]);
                Bx[c] = __fmul_rn(q0.w, dx[c]);
                dy[c] = q0.y - (fy0 + 4.0f * c);
                Cyy[c] = __fmul_rn(__fmul_rn(q1.x, dy[c]), dy[c]);
            }
            float al[4], G[4];
            bool valid[4];
            bool any = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                valid[i] = false;
                al[i] = 0.f; G[i] = 0.f;
                if (m4 & (1u << i)) {   // warp-uniform
                    const float p = power2_at(Axx[i & 1], Cyy[i >> 1], Bx[i & 1], dy[i >> 1]);
                    G[i] = ex2_approx(p);
                    al[i] = fminf(ALPHA_MAX, __fmul_rn(q1.y, G[i]));
                    valid[i] = (p <= 0.0f) && (al[i] >= ALPHA_MIN) && (pos <= last[i]);
                }
                any = any || valid[i];
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            const float2 q2 = s2[j];
            float m_x = 0.f, m_y = 0.f, m_xx = 0.f, m_xy = 0.f, m_yy = 0.f, g_o = 0.f, g_r = 0.f, g_g = 0.f, g_b = 0.f, g_d = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!__any_sync(0xffffffffu, valid[i])) continue;   // nobody in this patch: nothing changes
                const float ai = valid[i] ? al[i] : 0.0f;
                const float wgt = __fmul_rn(ai, T[i]);
                float g = dL0[i] * q1.z + dL1[i] * q1.w + dL2[i] * q2.x;
                if (DEPTH) g += dLd[i] * q2.y;
                F[i] = __fmaf_rn(wgt, g, F[i]);
                g_r = __fmaf_rn(wgt, dL0[i], g_r); g_g = __fmaf_rn(wgt, dL1[i], g_g); g_b = __fmaf_rn(wgt, dL2[i], g_b);
                if (DEPTH) g_d = __fmaf_rn(wgt, dLd[i], g_d);
                const float om = __fsub_rn(1.0f, ai);
                float dLda = T[i] * g - (S[i] - F[i]) * rcp_approx(om);
                T[i] = __fmul_rn(T[i], om);
                dLda = valid[i] ? dLda : 0.0f;
                g_o = __fmaf_rn(G[i], dLda, g_o);
                const float tt = q1.y * G[i] * dLda;
                const float u = tt * dx[i & 1], v = tt * dy[i >> 1];
                m_x += u; m_y += v;
                m_xx = __fmaf_rn(u, dx[i & 1], m_xx);
                m_xy = __fmaf_rn(u, dy[i >> 1], m_xy);
                m_yy = __fmaf_rn(v, dy[i >> 1], m_yy);
            }
            const float ra = reduce8_transposed(m_x, m_y, m_xx, m_xy, m_yy, g_o, g_r, g_g);
            float *d = a.dacc + (size_t)sid[j] * DACC_STRIDE;
            if ((l & 3) == 0) atomicAdd(d + (l >> 2), ra);
            if (DEPTH) {
                const float rb = reduce2_transposed(g_b, g_d);
                if ((l & 15) == 1) atomicAdd(d + 8 + (l >> 4), rb);
            } else {
                float rb = g_b;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rb += __shfl_xor_sync(0xffffffffu, rb, o);
                if (l == 1) atomicAdd(d + 8, rb);
            }
        }
    }
}

int launch_render_fwd_ps(const RenderFwdArgs &a, bool debug, cudaStream_t stream) {
    const int tiles = a.gx * a.gy;
    if (tiles <= 0) return GSB_OK;
    GSB_LAUNCH("render_fwd", debug, stream, render_fwd_ps_kernel, tiles, PS_NT, 0, a);
    return GSB_OK;
}

int launch_render_bwd_ps(const RenderBwdArgs &a, bool debug, cudaStream_t stream) {
    const int tiles = a.gx * a.gy;
    if (tiles <= 0) return GSB_OK;
    if (a.dL_dinvdepth) {
        GSB_LAUNCH("render_bwd", debug, stream, render_bwd_ps_kernel<true>, tiles, PS_NT, 0, a);
    } else {
        GSB_LAUNCH("render_bwd", debug, stream, render_bwd_ps_kernel<false>, tiles, PS_NT, 0, a);
    }
    return GSB_OK;
}

}  // namespace gsb
