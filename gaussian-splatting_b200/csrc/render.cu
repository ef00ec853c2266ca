// render.cu -- one-pixel-per-thread blend kernels.  The DEFAULT forward kernel (K6) lives here:
// render_fwd_pc_kernel (8x4 pixel patch per warp, sub-tile culling, log2-domain exponent; fwd variant 6).
// render_fwd_kernel / render_bwd_kernel are the round's first versions (back-to-front recursion with butterfly
// reductions), kept as the A/B baseline of DESIGN.md section 4; render_fwd_pc2_kernel is a rejected experiment.
//
// Replaces FORWARD::renderCUDA / BACKWARD::renderCUDA of the reference's
// cuda_rasterizer/{forward,backward}.cu (named in BASELINE.json north_star; absent from
// /root/reference).  Blend rules restated in oracle/gs_oracle.c.
//
// Data movement: a tile's sorted gaussian ids are contiguous in point_list; the 48-byte splat
// records they point to are gathered from the L2-resident record table (P * 48 B) into shared
// memory 256 at a time.  One 16x16 tile per CTA; a warp owns an 8x4 pixel patch so that a gaussian
// which misses the patch is skipped by the whole warp.
#include "blend_common.cuh"

namespace gsb {

constexpr int RB = 256;  // gaussians staged per round

__device__ __forceinline__ void pixel_of_thread(const int tile_x, const int tile_y, int &px, int &py) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    px = tile_x * TILE + (w & 1) * 8 + (l & 7);
    py = tile_y * TILE + (w >> 1) * 4 + (l >> 3);
}

__global__ void __launch_bounds__(256)
render_fwd_kernel(const RenderFwdArgs a) {
    __shared__ float4 s0[RB], s1[RB], s2[RB];
    const int tile = blockIdx.x;
    const int tile_x = tile % a.gx, tile_y = tile / a.gx;
    int px, py;
    pixel_of_thread(tile_x, tile_y, px, py);
    const bool inside = px < a.W && py < a.H;
    const float fx = (float)px, fy = (float)py;
    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);
    const int rounds = (todo + RB - 1) / RB;

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;
    uint32_t contributor = 0, last = 0;

    for (int rd = 0; rd < rounds; ++rd) {
        if (__syncthreads_count(done) == 256) break;
        const int idx = rd * RB + threadIdx.x;
        if (idx < todo) {
            const uint32_t g = a.point_list[range.x + idx];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            s0[threadIdx.x] = __ldg(rec);
            s1[threadIdx.x] = __ldg(rec + 1);
            s2[threadIdx.x] = __ldg(rec + 2);
        }
        __syncthreads();
        const int n = min(RB, todo - rd * RB);
        for (int j = 0; !done && j < n; ++j) {
            ++contributor;
            const float4 q0 = s0[j];
            const float4 q1 = s1[j];
            const float dx = q0.x - fx, dy = q0.y - fy;
            const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(ALPHA_MAX, q1.y * __expf(power));
            if (alpha < ALPHA_MIN) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < T_STOP) { done = true; continue; }
            const float4 q2 = s2[j];
            const float w = alpha * T;
            C0 += q1.z * w; C1 += q1.w * w; C2 += q2.x * w;
            Dp += q2.y * w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t pid = (size_t)py * a.W + px, HW = (size_t)a.W * a.H;
        a.final_T[pid] = T;
        a.n_contrib[pid] = last;
        a.out_color[pid] = C0 + T * __ldg(a.bg);
        a.out_color[HW + pid] = C1 + T * __ldg(a.bg + 1);
        a.out_color[2 * HW + pid] = C2 + T * __ldg(a.bg + 2);
        a.out_invdepth[pid] = Dp;
    }
}

// forward with sub-tile culling: each warp (8x4 patch) walks only the staged gaussians whose cull ellipse
// meets its patch (patch_cull.cuh).  Blend arithmetic identical to render_fwd_kernel.
template <int MINB, bool PREFETCH>
__global__ void __launch_bounds__(256, MINB)
render_fwd_pc_kernel(const RenderFwdArgs a) {
    __shared__ float4 s0[RB], s1[RB];
    __shared__ float2 s2[RB];
    __shared__ uint8_t smask[RB];
    __shared__ uint8_t slist[8][RB];
    const int tile = blockIdx.x;
    const int tile_x = tile % a.gx, tile_y = tile / a.gx;
    int px, py;
    pixel_of_thread(tile_x, tile_y, px, py);
    const bool inside = px < a.W && py < a.H;
    const float fx = (float)px, fy = (float)py;
    const float ox = (float)(tile_x * TILE), oy = (float)(tile_y * TILE);
    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);
    const int rounds = (todo + RB - 1) / RB;
    const int w = threadIdx.x >> 5;

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;
    uint32_t last = 0;

    for (int rd = 0; rd < rounds; ++rd) {
        if (__syncthreads_count(done) == 256) break;
        const int idx = rd * RB + threadIdx.x;
        if (idx < todo) {
            const uint32_t g = a.point_list[range.x + idx];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            const float4 q2 = __ldg(rec + 2);
            smask[threadIdx.x] = (uint8_t)patch_mask(q0.x, q0.y, q0.z, q0.w, q1.x, q2.z, ox, oy);
            stage_scale(q0, q1);   // conic pre-scaled for the log2-domain exponent (blend_common.cuh)
            s0[threadIdx.x] = q0; s1[threadIdx.x] = q1; s2[threadIdx.x] = make_float2(q2.x, q2.y);
        }
        __syncthreads();
        const int n = min(RB, todo - rd * RB);
        const int cnt = compact_hits(smask, n, 1u << w, slist[w]);
        // PREFETCH: the next list entry's record is read from shared memory while the current one is blended
        int jn = 0;
        float4 q0n = make_float4(0.f, 0.f, 0.f, 0.f), q1n = q0n;
        if (PREFETCH && cnt > 0) { jn = slist[w][0]; q0n = s0[jn]; q1n = s1[jn]; }
        for (int k = 0; !done && k < cnt; ++k) {
            int j;
            float4 q0, q1;
            if (PREFETCH) {
                j = jn; q0 = q0n; q1 = q1n;
                if (k + 1 < cnt) { jn = slist[w][k + 1]; q0n = s0[jn]; q1n = s1[jn]; }
            } else {
                j = slist[w][k]; q0 = s0[j]; q1 = s1[j];
            }
            const float dx = q0.x - fx, dy = q0.y - fy;
            // same operation order as render_mp.cu / render_ps.cu, so forward and backward agree bit for bit on alpha
            const float power = power2_at(__fmul_rn(__fmul_rn(q0.z, dx), dx), __fmul_rn(__fmul_rn(q1.x, dy), dy), __fmul_rn(q0.w, dx), dy);
            if (power > 0.0f) continue;
            const float alpha = fminf(ALPHA_MAX, __fmul_rn(q1.y, ex2_approx(power)));
            if (alpha < ALPHA_MIN) continue;
            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
            if (test_T < T_STOP) { done = true; continue; }
            const float2 q2 = s2[j];
            const float wgt = __fmul_rn(alpha, T);
            C0 = __fmaf_rn(q1.z, wgt, C0); C1 = __fmaf_rn(q1.w, wgt, C1); C2 = __fmaf_rn(q2.x, wgt, C2);
            Dp = __fmaf_rn(q2.y, wgt, Dp);
            T = test_T;
            last = (uint32_t)(rd * RB + j + 1);
        }
    }
    if (inside) {
        const size_t pid = (size_t)py * a.W + px, HW = (size_t)a.W * a.H;
        a.final_T[pid] = T;
        a.n_contrib[pid] = last;
        a.out_color[pid] = C0 + T * __ldg(a.bg);
        a.out_color[HW + pid] = C1 + T * __ldg(a.bg + 1);
        a.out_color[2 * HW + pid] = C2 + T * __ldg(a.bg + 2);
        a.out_invdepth[pid] = Dp;
    }
}

// Forward, two pixels per lane: a warp owns one 16x4 band of the tile (both 8x4 patches of that band); lane l
// blends pixels (l & 7, l >> 3) of the left and of the right patch, i.e. two pixels 8 columns apart that share dy and
// every per-gaussian operand, so their FP32 arithmetic runs as packed f32x2 instructions (FFMA2).  Culling stays at
// patch granularity through the staged 8-bit mask: a warp walks the gaussians that reach either of its two patches.
__global__ void __launch_bounds__(128)
render_fwd_pc2_kernel(const RenderFwdArgs a) {
    constexpr int NT = 128;
    __shared__ float4 s0[RB], s1[RB];
    __shared__ float2 s2[RB];
    __shared__ uint8_t smask[RB];
    __shared__ uint8_t slist[4][RB];
    const int tile = blockIdx.x;
    const int tile_x = tile % a.gx, tile_y = tile / a.gx;
    const int t = threadIdx.x, w = t >> 5, l = t & 31;
    const int px0 = tile_x * TILE + (l & 7), py = tile_y * TILE + 4 * w + (l >> 3);
    const float fx0 = (float)px0, fy = (float)py;
    const float ox = (float)(tile_x * TILE), oy = (float)(tile_y * TILE);
    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);
    const int rounds = (todo + RB - 1) / RB;

    bool done0 = !(px0 < a.W && py < a.H), done1 = !(px0 + 8 < a.W && py < a.H);
    f32x2 T = pk1(1.0f), C0 = pk1(0.f), C1 = C0, C2 = C0, Dp = C0;
    uint32_t last0 = 0, last1 = 0;
    const f32x2 one2 = pk1(1.0f);

    for (int rd = 0; rd < rounds; ++rd) {
        if (__syncthreads_and(done0 && done1)) break;
        const int n = min(RB, todo - rd * RB);
        for (int k = t; k < n; k += NT) {
            const uint32_t g = a.point_list[range.x + rd * RB + k];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            float4 q0 = __ldg(rec), q1 = __ldg(rec + 1);
            const float4 q2 = __ldg(rec + 2);
            smask[k] = (uint8_t)patch_mask(q0.x, q0.y, q0.z, q0.w, q1.x, q2.z, ox, oy);
            stage_scale(q0, q1);
            s0[k] = q0; s1[k] = q1; s2[k] = make_float2(q2.x, q2.y);
        }
        __syncthreads();
        const int cnt = compact_hits(smask, n, 3u << (2 * w), slist[w]);
        for (int k = 0; k < cnt; ++k) {
            // warp-uniform exit: the vote below needs all 32 lanes, so a lane may not leave on its own
            if (__all_sync(0xffffffffu, done0 && done1)) break;
            const int j = slist[w][k];
            const float4 q0 = s0[j];
            const float4 q1 = s1[j];
            const f32x2 dx2 = pk(q0.x - fx0, q0.x - (fx0 + 8.0f));
            const float dy = q0.y - fy;
            const float Cyy = __fmul_rn(__fmul_rn(q1.x, dy), dy);
            const f32x2 Axx2 = mul2(mul2(pk1(q0.z), dx2), dx2);
            const f32x2 p2 = fma2(mul2(pk1(q0.w), dx2), pk1(dy), add2(Axx2, pk1(Cyy)));   // power2_at, both columns
            float p0, p1;
            unpk(p2, p0, p1);
            const float al0 = fminf(ALPHA_MAX, __fmul_rn(q1.y, ex2_approx(p0)));
            const float al1 = fminf(ALPHA_MAX, __fmul_rn(q1.y, ex2_approx(p1)));
            const bool v0 = (p0 <= 0.0f) && (al0 >= ALPHA_MIN) && !done0;
            const bool v1 = (p1 <= 0.0f) && (al1 >= ALPHA_MIN) && !done1;
            if (!__any_sync(0xffffffffu, v0 || v1)) continue;
            const f32x2 al2 = pk(al0, al1);
            const f32x2 tT = mul2(T, sub2(one2, al2));
            float t0, t1;
            unpk(tT, t0, t1);
            const bool stop0 = v0 && (t0 < T_STOP), stop1 = v1 && (t1 < T_STOP);
            done0 = done0 || stop0; done1 = done1 || stop1;
            const bool u0 = v0 && !stop0, u1 = v1 && !stop1;
            const f32x2 wgt = mul2(pk(u0 ? al0 : 0.0f, u1 ? al1 : 0.0f), T);
            const float2 q2 = s2[j];
            C0 = fma2(pk1(q1.z), wgt, C0); C1 = fma2(pk1(q1.w), wgt, C1); C2 = fma2(pk1(q2.x), wgt, C2);
            Dp = fma2(pk1(q2.y), wgt, Dp);
            float T0, T1;
            unpk(T, T0, T1);
            T = pk(u0 ? t0 : T0, u1 ? t1 : T1);
            const uint32_t pos = (uint32_t)(rd * RB + j + 1);
            last0 = u0 ? pos : last0; last1 = u1 ? pos : last1;
        }
    }
    const float bg0 = __ldg(a.bg), bg1 = __ldg(a.bg + 1), bg2 = __ldg(a.bg + 2);
    const size_t HW = (size_t)a.W * a.H;
    float Tl, Th, c0l, c0h, c1l, c1h, c2l, c2h, dl, dh;
    unpk(T, Tl, Th); unpk(C0, c0l, c0h); unpk(C1, c1l, c1h); unpk(C2, c2l, c2h); unpk(Dp, dl, dh);
    if (py < a.H) {
        if (px0 < a.W) {
            const size_t pid = (size_t)py * a.W + px0;
            a.final_T[pid] = Tl; a.n_contrib[pid] = last0;
            a.out_color[pid] = c0l + Tl * bg0; a.out_color[HW + pid] = c1l + Tl * bg1; a.out_color[2 * HW + pid] = c2l + Tl * bg2;
            a.out_invdepth[pid] = dl;
        }
        if (px0 + 8 < a.W) {
            const size_t pid = (size_t)py * a.W + px0 + 8;
            a.final_T[pid] = Th; a.n_contrib[pid] = last1;
            a.out_color[pid] = c0h + Th * bg0; a.out_color[HW + pid] = c1h + Th * bg1; a.out_color[2 * HW + pid] = c2h + Th * bg2;
            a.out_invdepth[pid] = dh;
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 8 values per lane -> lane L ends with the warp total of value (L >> 2): 9 shuffles instead of 40.
// Each xor step halves the number of values a lane carries: the lane keeps the half selected by its own
// bit and adds the partner's copy of that half.
__device__ __forceinline__ float warp_reduce8_transposed(const float v0, const float v1, const float v2, const float v3,
                                                         const float v4, const float v5, const float v6, const float v7) {
    const int lane = threadIdx.x & 31;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    const float k0 = (b4 ? v4 : v0) + __shfl_xor_sync(0xffffffffu, b4 ? v0 : v4, 16);
    const float k1 = (b4 ? v5 : v1) + __shfl_xor_sync(0xffffffffu, b4 ? v1 : v5, 16);
    const float k2 = (b4 ? v6 : v2) + __shfl_xor_sync(0xffffffffu, b4 ? v2 : v6, 16);
    const float k3 = (b4 ? v7 : v3) + __shfl_xor_sync(0xffffffffu, b4 ? v3 : v7, 16);
    const float m0 = (b3 ? k2 : k0) + __shfl_xor_sync(0xffffffffu, b3 ? k0 : k2, 8);
    const float m1 = (b3 ? k3 : k1) + __shfl_xor_sync(0xffffffffu, b3 ? k1 : k3, 8);
    float r = (b2 ? m1 : m0) + __shfl_xor_sync(0xffffffffu, b2 ? m0 : m1, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;  // value index = 4*b4 + 2*b3 + b2 = lane >> 2
}

// 2 values per lane -> lanes 0..15 end with the total of v0, lanes 16..31 with the total of v1: 5 shuffles
__device__ __forceinline__ float warp_reduce2_transposed(const float v0, const float v1) {
    const bool b4 = threadIdx.x & 16;
    float r = (b4 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, b4 ? v0 : v1, 16);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// per-pixel threads; the 10 per-gaussian gradient terms are reduced over the warp and added to the
// per-gaussian accumulator.  VARIANT 0: butterfly all-reduce, lane 0 issues 10 atomics.
// VARIANT 1: transposed reduction (14 shuffles), 8 + 2 lanes issue one atomic each.
template <int VARIANT>
__global__ void __launch_bounds__(256)
render_bwd_kernel(const RenderBwdArgs a) {
    __shared__ float4 s0[RB], s1[RB], s2[RB];
    __shared__ uint32_t sid[RB];
    __shared__ uint32_t s_max;
    const int tile = blockIdx.x;
    const int tile_x = tile % a.gx, tile_y = tile / a.gx;
    int px, py;
    pixel_of_thread(tile_x, tile_y, px, py);
    const bool inside = px < a.W && py < a.H;
    const float fx = (float)px, fy = (float)py;
    const uint2 range = a.ranges[tile];
    const size_t pid = (size_t)py * a.W + px, HW = (size_t)a.W * a.H;

    const float T_final = inside ? a.final_T[pid] : 0.f;
    const uint32_t last = inside ? a.n_contrib[pid] : 0u;
    float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLd = 0.f;
    if (inside) {
        dLp0 = a.dL_dcolor[pid]; dLp1 = a.dL_dcolor[HW + pid]; dLp2 = a.dL_dcolor[2 * HW + pid];
        if (a.dL_dinvdepth) dLd = a.dL_dinvdepth[pid];
    }
    const float bg_dot = __ldg(a.bg) * dLp0 + __ldg(a.bg + 1) * dLp1 + __ldg(a.bg + 2) * dLp2;
    const float ddx = 0.5f * a.W, ddy = 0.5f * a.H;

    // the deepest position any pixel of the tile reached
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, last);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, wmax);
    __syncthreads();
    const int todo = (int)s_max;  // positions todo .. 1 (1-based) are visited back to front
    const int rounds = (todo + RB - 1) / RB;

    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f;
    float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f, last_alpha = 0.f;

    for (int rd = 0; rd < rounds; ++rd) {
        __syncthreads();
        const int k = rd * RB + threadIdx.x;  // k-th from the back
        if (k < todo) {
            const uint32_t g = a.point_list[range.x + (todo - 1 - k)];
            const float4 *rec = a.splat + (size_t)g * SPLAT_F4;
            sid[threadIdx.x] = g;
            s0[threadIdx.x] = __ldg(rec);
            s1[threadIdx.x] = __ldg(rec + 1);
            s2[threadIdx.x] = __ldg(rec + 2);
        }
        __syncthreads();
        const int n = min(RB, todo - rd * RB);
        for (int j = 0; j < n; ++j) {
            const uint32_t pos = (uint32_t)(todo - (rd * RB + j));  // 1-based position in the tile list
            const float4 q0 = s0[j];
            const float4 q1 = s1[j];
            const float dx = q0.x - fx, dy = q0.y - fy;
            const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
            const float G = __expf(power);
            const float alpha = fminf(ALPHA_MAX, q1.y * G);
            const bool valid = (pos <= last) && (power <= 0.0f) && (alpha >= ALPHA_MIN);
            if (!__any_sync(0xffffffffu, valid)) continue;
            float g_mx = 0.f, g_my = 0.f, g_A = 0.f, g_B = 0.f, g_C = 0.f, g_o = 0.f, g_r = 0.f, g_g = 0.f, g_b = 0.f, g_d = 0.f;
            if (valid) {
                const float4 q2 = s2[j];
                T = T / (1.0f - alpha);
                const float w = alpha * T;
                acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = q1.z;
                acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = q1.w;
                acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = q2.x;
                accd = last_alpha * ld + (1.f - last_alpha) * accd; ld = q2.y;
                float dL_dalpha = (q1.z - acc0) * dLp0 + (q1.w - acc1) * dLp1 + (q2.x - acc2) * dLp2 + (q2.y - accd) * dLd;
                g_r = w * dLp0; g_g = w * dLp1; g_b = w * dLp2; g_d = w * dLd;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dpow = q1.y * G * dL_dalpha;
                g_mx = dL_dpow * (-q0.z * dx - q0.w * dy) * ddx;
                g_my = dL_dpow * (-q1.x * dy - q0.w * dx) * ddy;
                g_A = dL_dpow * (-0.5f * dx * dx);
                g_B = dL_dpow * (-dx * dy);
                g_C = dL_dpow * (-0.5f * dy * dy);
                g_o = G * dL_dalpha;
            }
            float *d = a.dacc + (size_t)sid[j] * DACC_STRIDE;
            if (VARIANT == 0) {
                g_mx = warp_sum(g_mx); g_my = warp_sum(g_my); g_A = warp_sum(g_A); g_B = warp_sum(g_B); g_C = warp_sum(g_C);
                g_o = warp_sum(g_o); g_r = warp_sum(g_r); g_g = warp_sum(g_g); g_b = warp_sum(g_b); g_d = warp_sum(g_d);
                if ((threadIdx.x & 31) == 0) {
                    atomicAdd(d + 0, g_mx); atomicAdd(d + 1, g_my); atomicAdd(d + 2, g_A); atomicAdd(d + 3, g_B);
                    atomicAdd(d + 4, g_C); atomicAdd(d + 5, g_o); atomicAdd(d + 6, g_r); atomicAdd(d + 7, g_g);
                    atomicAdd(d + 8, g_b); atomicAdd(d + 9, g_d);
                }
            } else {
                const int lane = threadIdx.x & 31;
                const float ra = warp_reduce8_transposed(g_mx, g_my, g_A, g_B, g_C, g_o, g_r, g_g);
                const float rb = warp_reduce2_transposed(g_b, g_d);
                if ((lane & 3) == 0) atomicAdd(d + (lane >> 2), ra);
                if ((lane & 15) == 1) atomicAdd(d + 8 + (lane >> 4), rb);
            }
        }
    }
}

int launch_render_fwd(const RenderFwdArgs &a, int variant, bool debug, cudaStream_t stream) {
    const int tiles = a.gx * a.gy;
    if (tiles <= 0) return GSB_OK;
    if (variant == 4) {
        GSB_LAUNCH("render_fwd", debug, stream, (render_fwd_pc_kernel<1, false>), tiles, 256, 0, a);
    } else if (variant == 7) {   // 6 CTAs/SM + shared-memory prefetch of the next record
        GSB_LAUNCH("render_fwd", debug, stream, (render_fwd_pc_kernel<6, true>), tiles, 256, 0, a);
    } else if (variant == 9) {   // two pixels per lane, packed f32x2 arithmetic
        GSB_LAUNCH("render_fwd", debug, stream, render_fwd_pc2_kernel, tiles, 128, 0, a);
    } else if (variant == 6) {   // register budget for 6 CTAs (48 warps) per SM
        GSB_LAUNCH("render_fwd", debug, stream, (render_fwd_pc_kernel<6, false>), tiles, 256, 0, a);
    } else if (variant == 8) {   // 8 CTAs (64 warps) per SM
        GSB_LAUNCH("render_fwd", debug, stream, (render_fwd_pc_kernel<8, false>), tiles, 256, 0, a);
    } else {
        GSB_LAUNCH("render_fwd", debug, stream, render_fwd_kernel, tiles, 256, 0, a);
    }
    return GSB_OK;
}

int launch_render_bwd(const RenderBwdArgs &a, int variant, bool debug, cudaStream_t stream) {
    const int tiles = a.gx * a.gy;
    if (tiles <= 0) return GSB_OK;
    if (variant == 0) {
        GSB_LAUNCH("render_bwd", debug, stream, render_bwd_kernel<0>, tiles, 256, 0, a);
    } else {
        GSB_LAUNCH("render_bwd", debug, stream, render_bwd_kernel<1>, tiles, 256, 0, a);
    }
    return GSB_OK;
}

}  // namespace gsb
