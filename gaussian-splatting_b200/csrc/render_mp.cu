// render_mp.cu -- multi-pixel-per-thread blend kernels.  The DEFAULT backward kernel (K7) lives here:
// render_bwd_mp2x_kernel (2x2 pixels per thread, packed f32x2 arithmetic, sub-tile culling; bwd variant 10).
// render_bwd_mp_kernel is its scalar predecessor; render_fwd_mp_kernel (multi-pixel forward) measured slower than
// the one-pixel forward of render.cu and is kept for A/B only.
//
// ncu on the one-pixel-per-thread kernels (profiles/r1_render_v0.md) shows both are bound by the SM
// issue rate (issue active 81-87 %, DRAM 1-2 %): the cost is instructions per (pixel, gaussian) pair,
// not bytes.  These kernels cut that count:
//   * a thread owns a 2 x QH pixel block, so the shared-memory reads, loop overhead and -- in the
//     backward pass -- the cross-lane reduction of the ten per-gaussian gradient terms are paid once per
//     2*QH pixels instead of once per pixel;
//   * the exponent is evaluated in the log2 domain (conic pre-scaled while staging): one MUFU.EX2;
//   * the backward pass walks FRONT to back with two scalars of state per pixel (T and the running
//     dL-weighted front colour F) instead of the ten the back-to-front recursion carries:
//         dL/dalpha_k = T_k g_k - (S - F_k) / (1 - alpha_k),   g_k = dLdC . c_k + dLdD / z_k,
//         S = dLdC . C_out + dLdD D_out,   F_k = sum_{j<=k} alpha_j T_j g_j
//     (algebraically the recursion of oracle/gs_oracle.c; background enters through C_out);
//   * the geometric gradient terms are accumulated as raw moments of d(power) (sum t dx, t dy, t dx^2,
//     t dx dy, t dy^2) and turned into mean / conic gradients once per gaussian in preprocess_bwd.
#include "blend_common.cuh"

namespace gsb {

template <int QH>
__global__ void __launch_bounds__(256 / (2 * QH))
render_fwd_mp_kernel(const RenderFwdArgs a) {
    constexpr int NT = 256 / (2 * QH);
    constexpr int NPX = 2 * QH;
    __shared__ float4 s0[MP_R], s1[MP_R];
    __shared__ float2 s2[MP_R];
    const int tile = blockIdx.x;
    const int ox = (tile % a.gx) * TILE, oy = (tile / a.gx) * TILE;
    const int t = threadIdx.x;
    const int px0 = ox + 2 * (t & 7), py0 = oy + QH * (t >> 3);
    const float fx0 = (float)px0, fy0 = (float)py0;
    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);

    float T[NPX], C0[NPX], C1[NPX], C2[NPX], Dp[NPX];
    uint32_t last[NPX];
    bool done[NPX];
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
        T[i] = 1.0f; C0[i] = C1[i] = C2[i] = Dp[i] = 0.f; last[i] = 0u;
        done[i] = !((px0 + (i & 1)) < a.W && (py0 + (i >> 1)) < a.H);
    }

    for (int base = 0; base < todo; base += MP_R) {
        bool all_done = true;
#pragma unroll
        for (int i = 0; i < NPX; ++i) all_done = all_done && done[i];
        if (__syncthreads_and(all_done)) break;
        const int n = min(MP_R, todo - base);
        for (int k = t; k < n; k += NT) {
            const uint32_t g = a.point_list[range.x + base + k];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            const float4 q2 = __ldg(rec + 2);
            stage_scale(q0, q1);
            s0[k] = q0; s1[k] = q1; s2[k] = make_float2(q2.x, q2.y);
        }
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            const float4 q0 = s0[j];
            const float4 q1 = s1[j];
            float dx[2], Axx[2], Bx[2], dy[QH], Cyy[QH];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                dx[c] = q0.x - (fx0 + (float)c);
                Axx[c] = __fmul_rn(__fmul_rn(q0.z, dx[c]), dx[c]);
                Bx[c] = __fmul_rn(q0.w, dx[c]);
            }
#pragma unroll
            for (int r = 0; r < QH; ++r) {
                dy[r] = q0.y - (fy0 + (float)r);
                Cyy[r] = __fmul_rn(__fmul_rn(q1.x, dy[r]), dy[r]);
            }
            float al[NPX];
            bool valid[NPX];
            bool any = false;
#pragma unroll
            for (int i = 0; i < NPX; ++i) {
                const float p = power2_at(Axx[i & 1], Cyy[i >> 1], Bx[i & 1], dy[i >> 1]);
                al[i] = fminf(ALPHA_MAX, __fmul_rn(q1.y, ex2_approx(p)));
                valid[i] = (p <= 0.0f) && (al[i] >= ALPHA_MIN) && !done[i];
                any = any || valid[i];
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            const float2 q2 = s2[j];
            const uint32_t pos = (uint32_t)(base + j + 1);
#pragma unroll
            for (int i = 0; i < NPX; ++i) {
                const float test_T = __fmul_rn(T[i], __fsub_rn(1.0f, al[i]));
                const bool stop = valid[i] && (test_T < T_STOP);
                done[i] = done[i] || stop;
                const bool use = valid[i] && !stop;
                const float w = use ? __fmul_rn(al[i], T[i]) : 0.0f;
                C0[i] = __fmaf_rn(q1.z, w, C0[i]);
                C1[i] = __fmaf_rn(q1.w, w, C1[i]);
                C2[i] = __fmaf_rn(q2.x, w, C2[i]);
                Dp[i] = __fmaf_rn(q2.y, w, Dp[i]);
                T[i] = use ? test_T : T[i];
                last[i] = use ? pos : last[i];
            }
        }
    }
    const float bg0 = __ldg(a.bg), bg1 = __ldg(a.bg + 1), bg2 = __ldg(a.bg + 2);
    const size_t HW = (size_t)a.W * a.H;
#pragma unroll
    for (int r = 0; r < QH; ++r) {
        const int py = py0 + r;
        if (py >= a.H) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int px = px0 + c;
            if (px >= a.W) continue;
            const int i = 2 * r + c;
            const size_t pid = (size_t)py * a.W + px;
            a.final_T[pid] = T[i];
            a.n_contrib[pid] = last[i];
            a.out_color[pid] = C0[i] + T[i] * bg0;
            a.out_color[HW + pid] = C1[i] + T[i] * bg1;
            a.out_color[2 * HW + pid] = C2[i] + T[i] * bg2;
            a.out_invdepth[pid] = Dp[i];
        }
    }
}

// dacc layout written here (DACC_MOMENTS): 0 sum t dx, 1 sum t dy, 2 sum t dx^2, 3 sum t dx dy, 4 sum t dy^2,
// 5 sum G dL/dalpha, 6..8 sum w dL/dC, 9 sum w dL/dD, with t = opacity G dL/dalpha = dL/d(power).
template <int QH, bool CULL, int MINB>
__global__ void __launch_bounds__(256 / (2 * QH), MINB)
render_bwd_mp_kernel(const RenderBwdArgs a) {
    constexpr int NT = 256 / (2 * QH);
    constexpr int NPX = 2 * QH;
    constexpr int NW = NT / 32;
    __shared__ float4 s0[MP_R], s1[MP_R];
    __shared__ float2 s2[MP_R];
    __shared__ uint32_t sid[MP_R];
    __shared__ uint32_t s_max;
    __shared__ uint8_t smask[CULL ? MP_R : 1];
    __shared__ uint8_t slist[CULL ? NW : 1][CULL ? MP_R : 1];
    // rows covered by this warp: [16/NW * w, ...) = 4-row bands 4/NW*w .. ; mask bits 2*band + col
    const uint32_t want = CULL ? (((1u << (8 / NW)) - 1u) << ((8 / NW) * (threadIdx.x >> 5))) : 0u;
    const int tile = blockIdx.x;
    const int ox = (tile % a.gx) * TILE, oy = (tile / a.gx) * TILE;
    const int t = threadIdx.x;
    const int px0 = ox + 2 * (t & 7), py0 = oy + QH * (t >> 3);
    const float fx0 = (float)px0, fy0 = (float)py0;
    const uint2 range = a.ranges[tile];
    const size_t HW = (size_t)a.W * a.H;

    float T[NPX], F[NPX], S[NPX], dL0[NPX], dL1[NPX], dL2[NPX], dLd[NPX];
    uint32_t last[NPX];
    uint32_t my_max = 0;
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
        const int px = px0 + (i & 1), py = py0 + (i >> 1);
        T[i] = 1.0f; F[i] = 0.f; S[i] = 0.f; dL0[i] = dL1[i] = dL2[i] = dLd[i] = 0.f; last[i] = 0u;
        if (px < a.W && py < a.H) {
            const size_t pid = (size_t)py * a.W + px;
            last[i] = a.n_contrib[pid];
            dL0[i] = a.dL_dcolor[pid]; dL1[i] = a.dL_dcolor[HW + pid]; dL2[i] = a.dL_dcolor[2 * HW + pid];
            dLd[i] = a.dL_dinvdepth ? a.dL_dinvdepth[pid] : 0.f;
            S[i] = dL0[i] * a.out_color[pid] + dL1[i] * a.out_color[HW + pid] + dL2[i] * a.out_color[2 * HW + pid] +
                   dLd[i] * a.out_invdepth[pid];
        }
        my_max = max(my_max, last[i]);
    }
    if (t == 0) s_max = 0;
    __syncthreads();
    my_max = __reduce_max_sync(0xffffffffu, my_max);   // deepest list position this warp's pixels blended
    if ((t & 31) == 0) atomicMax(&s_max, my_max);
    __syncthreads();
    const int todo = (int)s_max;  // the deepest list position any pixel of the tile blended
    const int my_todo = (int)my_max;

    for (int base = 0; base < todo; base += MP_R) {
        __syncthreads();
        const int n = min(MP_R, todo - base);
        for (int k = t; k < n; k += NT) {
            const uint32_t g = a.point_list[range.x + base + k];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            const float4 q2 = __ldg(rec + 2);
            if (CULL) smask[k] = (uint8_t)patch_mask(q0.x, q0.y, q0.z, q0.w, q1.x, q2.z, (float)ox, (float)oy);
            stage_scale(q0, q1);
            s0[k] = q0; s1[k] = q1; s2[k] = make_float2(q2.x, q2.y); sid[k] = g;
        }
        __syncthreads();
        const int nw = max(0, min(n, my_todo - base));   // nothing behind this warp's own deepest pixel matters to it
        const int cnt = CULL ? compact_hits(smask, nw, want, slist[CULL ? (t >> 5) : 0]) : nw;
        for (int kk = 0; kk < cnt; ++kk) {
            const int j = CULL ? (int)slist[CULL ? (t >> 5) : 0][kk] : kk;
            const float4 q0 = s0[j];
            const float4 q1 = s1[j];
            const uint32_t pos = (uint32_t)(base + j + 1);
            float dx[2], Axx[2], Bx[2], dy[QH], Cyy[QH];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                dx[c] = q0.x - (fx0 + (float)c);
                Axx[c] = __fmul_rn(__fmul_rn(q0.z, dx[c]), dx[c]);
                Bx[c] = __fmul_rn(q0.w, dx[c]);
            }
#pragma unroll
            for (int r = 0; r < QH; ++r) {
                dy[r] = q0.y - (fy0 + (float)r);
                Cyy[r] = __fmul_rn(__fmul_rn(q1.x, dy[r]), dy[r]);
            }
            float al[NPX], G[NPX];
            bool valid[NPX];
            bool any = false;
#pragma unroll
            for (int i = 0; i < NPX; ++i) {
                const float p = power2_at(Axx[i & 1], Cyy[i >> 1], Bx[i & 1], dy[i >> 1]);
                G[i] = ex2_approx(p);
                al[i] = fminf(ALPHA_MAX, __fmul_rn(q1.y, G[i]));
                valid[i] = (p <= 0.0f) && (al[i] >= ALPHA_MIN) && (pos <= last[i]);
                any = any || valid[i];
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            const float2 q2 = s2[j];
            float m_x = 0.f, m_y = 0.f, m_xx = 0.f, m_xy = 0.f, m_yy = 0.f, g_o = 0.f, g_r = 0.f, g_g = 0.f, g_b = 0.f, g_d = 0.f;
#pragma unroll
            for (int i = 0; i < NPX; ++i) {
                const float ai = valid[i] ? al[i] : 0.0f;
                const float w = __fmul_rn(ai, T[i]);
                const float g = dL0[i] * q1.z + dL1[i] * q1.w + dL2[i] * q2.x + dLd[i] * q2.y;
                F[i] = __fmaf_rn(w, g, F[i]);
                g_r = __fmaf_rn(w, dL0[i], g_r); g_g = __fmaf_rn(w, dL1[i], g_g);
                g_b = __fmaf_rn(w, dL2[i], g_b); g_d = __fmaf_rn(w, dLd[i], g_d);
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
            const int lane = t & 31;
            const float ra = reduce8_transposed(m_x, m_y, m_xx, m_xy, m_yy, g_o, g_r, g_g);
            const float rb = reduce2_transposed(g_b, g_d);
            float *d = a.dacc + (size_t)sid[j] * DACC_STRIDE;
            if ((lane & 3) == 0) atomicAdd(d + (lane >> 2), ra);
            if ((lane & 15) == 1) atomicAdd(d + 8 + (lane >> 4), rb);
        }
    }
}

// Packed variant of render_bwd_mp_kernel<2, CULL>: the two pixels of a row of the thread's 2x2 block share every
// per-gaussian operand, so their FP32 mul / add / fma run as f32x2 instructions (FFMA2): ~20 % fewer issue slots.
// Arithmetic per lane is IEEE round-to-nearest exactly as in the scalar kernel.
// SMEM_REDUCE: the cross-lane sum of the ten (packed) gradient terms goes through shared memory instead of the
// shuffle network: every lane stores its ten f32x2 partials, 30 lanes then each add one third of one term's column
// (11 independent loads + packed adds), two shuffles combine the thirds and lanes 0..9 issue the atomics.  Fewer
// instructions (~38 vs ~62) and a much shorter dependency chain than five shuffle levels.
template <bool CULL, bool DEPTH, int MINB, bool PREFETCH, bool SMEM_REDUCE = false>
__global__ void __launch_bounds__(64, MINB)
render_bwd_mp2x_kernel(const RenderBwdArgs a) {
    constexpr int NT = 64;
    __shared__ f32x2 sred[SMEM_REDUCE ? 2 : 1][SMEM_REDUCE ? 10 * 33 : 1];
    __shared__ float4 s0[MP_R], s1[MP_R];
    __shared__ float2 s2[MP_R];
    __shared__ uint32_t sid[MP_R];
    __shared__ uint32_t s_max;
    __shared__ uint8_t smask[CULL ? MP_R : 1];
    __shared__ uint8_t slist[CULL ? 2 : 1][CULL ? MP_R : 1];
    const uint32_t want = CULL ? (0xfu << (4 * (threadIdx.x >> 5))) : 0u;
    const int tile = blockIdx.x;
    const int ox = (tile % a.gx) * TILE, oy = (tile / a.gx) * TILE;
    const int t = threadIdx.x;
    const int px0 = ox + 2 * (t & 7), py0 = oy + 2 * (t >> 3);
    const float fx0 = (float)px0, fy0 = (float)py0;
    const uint2 range = a.ranges[tile];
    const size_t HW = (size_t)a.W * a.H;

    // per-pixel state, packed by row: element c of row r is pixel (px0 + c, py0 + r)
    f32x2 T[2], F[2], S[2], dL0[2], dL1[2], dL2[2], dLd[2];
    uint32_t last[4];
    uint32_t my_max = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float vS[2] = {0.f, 0.f}, v0[2] = {0.f, 0.f}, v1[2] = {0.f, 0.f}, v2[2] = {0.f, 0.f}, vd[2] = {0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int px = px0 + c, py = py0 + r;
            last[2 * r + c] = 0u;
            if (px < a.W && py < a.H) {
                const size_t pid = (size_t)py * a.W + px;
                last[2 * r + c] = a.n_contrib[pid];
                v0[c] = a.dL_dcolor[pid]; v1[c] = a.dL_dcolor[HW + pid]; v2[c] = a.dL_dcolor[2 * HW + pid];
                vS[c] = v0[c] * a.out_color[pid] + v1[c] * a.out_color[HW + pid] + v2[c] * a.out_color[2 * HW + pid];
                if (DEPTH) { vd[c] = a.dL_dinvdepth[pid]; vS[c] += vd[c] * a.out_invdepth[pid]; }
            }
            my_max = max(my_max, last[2 * r + c]);
        }
        T[r] = pk1(1.0f); F[r] = pk1(0.0f); S[r] = pk(vS[0], vS[1]);
        dL0[r] = pk(v0[0], v0[1]); dL1[r] = pk(v1[0], v1[1]); dL2[r] = pk(v2[0], v2[1]); dLd[r] = pk(vd[0], vd[1]);
    }
    if (t == 0) s_max = 0;
    __syncthreads();
    my_max = __reduce_max_sync(0xffffffffu, my_max);
    if ((t & 31) == 0) atomicMax(&s_max, my_max);
    __syncthreads();
    const int todo = (int)s_max;
    const int my_todo = (int)my_max;
    const f32x2 one2 = pk1(1.0f);

    for (int base = 0; base < todo; base += MP_R) {
        __syncthreads();
        const int n = min(MP_R, todo - base);
        for (int k = t; k < n; k += NT) {
            const uint32_t g = a.point_list[range.x + base + k];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            const float4 q2 = __ldg(rec + 2);
            if (CULL) smask[k] = (uint8_t)patch_mask(q0.x, q0.y, q0.z, q0.w, q1.x, q2.z, (float)ox, (float)oy);
            stage_scale(q0, q1);
            s0[k] = q0; s1[k] = q1; s2[k] = make_float2(q2.x, q2.y); sid[k] = g;
        }
        __syncthreads();
        const int nw = max(0, min(n, my_todo - base));
        const int cnt = CULL ? compact_hits(smask, nw, want, slist[CULL ? (t >> 5) : 0]) : nw;
        int jn = 0;
        float4 q0n = make_float4(0.f, 0.f, 0.f, 0.f), q1n = q0n;
        if (PREFETCH && cnt > 0) { jn = CULL ? (int)slist[CULL ? (t >> 5) : 0][0] : 0; q0n = s0[jn]; q1n = s1[jn]; }
        for (int kk = 0; kk < cnt; ++kk) {
            int j;
            float4 q0, q1;
            if (PREFETCH) {   // next record read from shared memory while this one is processed
                j = jn; q0 = q0n; q1 = q1n;
                if (kk + 1 < cnt) { jn = CULL ? (int)slist[CULL ? (t >> 5) : 0][kk + 1] : kk + 1; q0n = s0[jn]; q1n = s1[jn]; }
            } else {
                j = CULL ? (int)slist[CULL ? (t >> 5) : 0][kk] : kk; q0 = s0[j]; q1 = s1[j];
            }
            const uint32_t pos = (uint32_t)(base + j + 1);
            // column-packed terms (shared by both rows): dx, A' dx^2, B' dx
            const f32x2 dx2 = pk(q0.x - fx0, q0.x - (fx0 + 1.0f));   // same rounding as the one-pixel kernels
            const f32x2 Axx2 = mul2(mul2(pk1(q0.z), dx2), dx2);
            const f32x2 Bx2 = mul2(pk1(q0.w), dx2);
            float dy[2], Cyy[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                dy[r] = q0.y - (fy0 + (float)r);
                Cyy[r] = __fmul_rn(__fmul_rn(q1.x, dy[r]), dy[r]);
            }
            float G[4], al[4];
            bool valid[4];
            bool any = false;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                // power2_at(Axx, Cyy, Bx, dy) = fma(Bx, dy, Axx + Cyy), both columns at once
                const f32x2 p2 = fma2(Bx2, pk1(dy[r]), add2(Axx2, pk1(Cyy[r])));
                float p[2];
                unpk(p2, p[0], p[1]);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int i = 2 * r + c;
                    G[i] = ex2_approx(p[c]);
                    al[i] = fminf(ALPHA_MAX, __fmul_rn(q1.y, G[i]));
                    valid[i] = (p[c] <= 0.0f) && (al[i] >= ALPHA_MIN) && (pos <= last[i]);
                    any = any || valid[i];
                }
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            const float2 q2 = s2[j];
            f32x2 m_x = pk1(0.f), m_y = m_x, m_xx = m_x, m_xy = m_x, m_yy = m_x, g_o = m_x, g_r = m_x, g_g = m_x, g_b = m_x, g_d = m_x;
            const f32x2 cr = pk1(q1.z), cg = pk1(q1.w), cb = pk1(q2.x), cd = pk1(q2.y), op2 = pk1(q1.y);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i0 = 2 * r, i1 = 2 * r + 1;
                // invalid lanes: alpha = 0 and G = 0 make every contribution below exactly zero
                const f32x2 ai = pk(valid[i0] ? al[i0] : 0.0f, valid[i1] ? al[i1] : 0.0f);
                const f32x2 Gm = pk(valid[i0] ? G[i0] : 0.0f, valid[i1] ? G[i1] : 0.0f);
                const f32x2 w = mul2(ai, T[r]);
                f32x2 g = fma2(dL2[r], cb, fma2(dL1[r], cg, mul2(dL0[r], cr)));
                if (DEPTH) g = fma2(dLd[r], cd, g);
                F[r] = fma2(w, g, F[r]);
                g_r = fma2(w, dL0[r], g_r); g_g = fma2(w, dL1[r], g_g); g_b = fma2(w, dL2[r], g_b);
                if (DEPTH) g_d = fma2(w, dLd[r], g_d);
                const f32x2 om = sub2(one2, ai);
                float om0, om1;
                unpk(om, om0, om1);
                const f32x2 rcp = pk(rcp_approx(om0), rcp_approx(om1));
                // dL/dalpha = T g - (S - F) / (1 - alpha); multiplied by G (0 for invalid lanes) wherever it is used
                const f32x2 dLda = sub2(mul2(T[r], g), mul2(sub2(S[r], F[r]), rcp));
                T[r] = mul2(T[r], om);
                const f32x2 Gd = mul2(Gm, dLda);
                g_o = add2(g_o, Gd);
                const f32x2 tt = mul2(op2, Gd);
                const f32x2 dy2 = pk1(dy[r]);
                const f32x2 u = mul2(tt, dx2), v = mul2(tt, dy2);
                m_x = add2(m_x, u); m_y = add2(m_y, v);
                m_xx = fma2(u, dx2, m_xx); m_xy = fma2(u, dy2, m_xy); m_yy = fma2(v, dy2, m_yy);
            }
            const int lane = t & 31;
            float *d = a.dacc + (size_t)sid[j] * DACC_STRIDE;
            if (SMEM_REDUCE) {
                f32x2 *sr = sred[SMEM_REDUCE ? (t >> 5) : 0];
                sr[0 * 33 + lane] = m_x; sr[1 * 33 + lane] = m_y; sr[2 * 33 + lane] = m_xx; sr[3 * 33 + lane] = m_xy;
                sr[4 * 33 + lane] = m_yy; sr[5 * 33 + lane] = g_o; sr[6 * 33 + lane] = g_r; sr[7 * 33 + lane] = g_g;
                sr[8 * 33 + lane] = g_b; sr[9 * 33 + lane] = g_d;
                __syncwarp();
                // lane = 10 * third + term (lanes 30, 31 idle): sums entries [11*third, 11*third + 11) of the term's column
                const int term = lane % 10, third = lane / 10;
                f32x2 acc2 = pk1(0.f);
                if (lane < 30) {
                    const f32x2 *col = sr + term * 33 + third * 11;
#pragma unroll
                    for (int q = 0; q < 11; ++q)
                        if (third * 11 + q < 32) acc2 = add2(acc2, col[q]);
                }
                float tot = hsum(acc2);
                tot += __shfl_down_sync(0xffffffffu, tot, 10) + __shfl_down_sync(0xffffffffu, tot, 20);
                if (lane < (DEPTH ? 10 : 9)) atomicAdd(d + lane, tot);
                __syncwarp();
                continue;
            }
            // both shuffle networks are issued before either atomic so that their (independent) chains overlap
            const float ra = reduce8_transposed(hsum(m_x), hsum(m_y), hsum(m_xx), hsum(m_xy), hsum(m_yy), hsum(g_o), hsum(g_r), hsum(g_g));
            float rb;
            if (DEPTH) {
                rb = reduce2_transposed(hsum(g_b), hsum(g_d));
            } else {
                rb = hsum(g_b);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rb += __shfl_xor_sync(0xffffffffu, rb, o);
            }
            if ((lane & 3) == 0) atomicAdd(d + (lane >> 2), ra);
            if (DEPTH) {
                if ((lane & 15) == 1) atomicAdd(d + 8 + (lane >> 4), rb);
            } else {
                if (lane == 1) atomicAdd(d + 8, rb);
            }
        }
    }
}

int launch_render_fwd_mp(const RenderFwdArgs &a, int qh, bool debug, cudaStream_t stream) {
    const int tiles = a.gx * a.gy;
    if (tiles <= 0) return GSB_OK;
    if (qh == 4) {
        GSB_LAUNCH("render_fwd", debug, stream, render_fwd_mp_kernel<4>, tiles, 32, 0, a);
    } else {
        GSB_LAUNCH("render_fwd", debug, stream, render_fwd_mp_kernel<2>, tiles, 64, 0, a);
    }
    return GSB_OK;
}

int launch_render_bwd_mp(const RenderBwdArgs &a, int qh, bool debug, cudaStream_t stream) {
    const int tiles = a.gx * a.gy;
    if (tiles <= 0) return GSB_OK;
    if (qh == 4) {
        GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp_kernel<4, false, 1>), tiles, 32, 0, a);
    } else if (qh == 2) {
        GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp_kernel<2, false, 1>), tiles, 64, 0, a);
    } else if (qh == -20) {   // packed f32x2 arithmetic (FFMA2), sub-tile culling
        if (a.dL_dinvdepth) {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, true, 1, false>), tiles, 64, 0, a);
        } else {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, false, 1, false>), tiles, 64, 0, a);
        }
    } else if (qh == -21 || qh == -22) {   // same, register budget for 12 / 16 CTAs per SM
        if (a.dL_dinvdepth) {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, true, 12, false>), tiles, 64, 0, a);
        } else if (qh == -21) {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, false, 12, false>), tiles, 64, 0, a);
        } else {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, false, 16, false>), tiles, 64, 0, a);
        }
    } else if (qh == -23) {   // 16 CTAs/SM + shared-memory prefetch of the next record
        if (a.dL_dinvdepth) {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, true, 12, true>), tiles, 64, 0, a);
        } else {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, false, 16, true>), tiles, 64, 0, a);
        }
    } else if (qh == -24) {   // packed arithmetic, 16 CTAs/SM, shared-memory reduction
        if (a.dL_dinvdepth) {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, true, 12, false, true>), tiles, 64, 0, a);
        } else {
            GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp2x_kernel<true, false, 16, false, true>), tiles, 64, 0, a);
        }
    } else if (qh == -12) {   // as -2, register budget for 12 CTAs / SM
        GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp_kernel<2, true, 12>), tiles, 64, 0, a);
    } else if (qh == -16) {   // as -2, register budget for 16 CTAs / SM
        GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp_kernel<2, true, 16>), tiles, 64, 0, a);
    } else {   // qh == -2: 2x2 pixels per thread + sub-tile culling
        GSB_LAUNCH("render_bwd", debug, stream, (render_bwd_mp_kernel<2, true, 1>), tiles, 64, 0, a);
    }
    return GSB_OK;
}

}  // namespace gsb
