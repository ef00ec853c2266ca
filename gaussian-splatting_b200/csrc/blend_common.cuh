// blend_common.cuh -- helpers of the blend kernels (render.cu): staged-record scaling, the pinned exponent, packed f32x2 arithmetic,
// the transposed warp reductions of the backward pass.
#pragma once
#include "kernels.cuh"
#include "patch_cull.cuh"

namespace gsb {

constexpr int MP_R = 128;           // gaussians staged per round (256 measured 2 % slower)
constexpr float LOG2E = 1.4426950408889634f;

#ifdef GSB_HOST_EMUL     // tests/host_emul: the approximate instructions become their libm counterparts
inline float ex2_approx(const float x) { return exp2f(x); }
inline float rcp_approx(const float x) { return 1.0f / x; }
#else
__device__ __forceinline__ float ex2_approx(const float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 1/x, one MUFU.RCP (<= 1 ulp); x = 1 - alpha is in [0.01, 1], no special cases to handle
__device__ __forceinline__ float rcp_approx(const float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

// staged record: q0 = {x, y, A', C'}, q1 = {B', opacity, r, g} with A' = -0.5 log2e A, B' = -log2e B, C' = -0.5 log2e C.
// (x, y) and (A', C') are 64-bit register pairs after the 128-bit shared-memory load, i.e. ready-made f32x2 operands:
// (dx, dy) = (x, y) - pixel is ONE packed add and (A' dx dx, C' dy dy) two packed multiplies.
__device__ __forceinline__ void stage_scale(float4 &q0, float4 &q1) {
    const float a = __fmul_rn(q0.z, -0.5f * LOG2E), b = __fmul_rn(q0.w, -LOG2E), c = __fmul_rn(q1.x, -0.5f * LOG2E);
    q0.z = a; q0.w = c; q1.x = b;
}

// log2-domain exponent at offset (dx, dy); pinned operation order: forward and backward must agree bit for bit
__device__ __forceinline__ float power2_at(const float Axx, const float Cyy, const float Bx, const float dy) {
    return __fmaf_rn(Bx, dy, __fadd_rn(Axx, Cyy));
}

// lane L ends with the warp total of value (L >> 2): 9 shuffles
__device__ __forceinline__ float reduce8_transposed(const float v0, const float v1, const float v2, const float v3,
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
    return r;
}

// lanes 0..15 end with the total of v0, lanes 16..31 with the total of v1: 5 shuffles
__device__ __forceinline__ float reduce2_transposed(const float v0, const float v1) {
    const bool b4 = threadIdx.x & 16;
    float r = (b4 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, b4 ? v0 : v1, 16);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// ---- packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for two lanes' worth of FP32) ----
// tools/micro/ffma2_bench.cu: FFMA2 runs at half the FFMA issue rate (same FLOP/s), so it frees issue slots for
// the ALU / MUFU / shuffle work around it -- what the issue-bound blend kernels need.
typedef unsigned long long f32x2;
#ifdef GSB_HOST_EMUL
inline f32x2 pk(const float lo, const float hi) { f32x2 r; float t[2] = {lo, hi}; memcpy(&r, t, 8); return r; }
inline void unpk(const f32x2 v, float &lo, float &hi) { float t[2]; memcpy(t, &v, 8); lo = t[0]; hi = t[1]; }
#define GSB_F32X2_OP(name, expr_lo, expr_hi) \
    { float al, ah, bl, bh, cl = 0.f, ch = 0.f; unpk(a, al, ah); unpk(b, bl, bh); (void)cl; (void)ch; return pk(expr_lo, expr_hi); }
inline f32x2 fma2(const f32x2 a, const f32x2 b, const f32x2 c) { float al, ah, bl, bh, cl, ch; unpk(a, al, ah); unpk(b, bl, bh); unpk(c, cl, ch); return pk(fmaf(al, bl, cl), fmaf(ah, bh, ch)); }
inline f32x2 mul2(const f32x2 a, const f32x2 b) GSB_F32X2_OP(mul2, al * bl, ah * bh)
inline f32x2 add2(const f32x2 a, const f32x2 b) GSB_F32X2_OP(add2, al + bl, ah + bh)
inline f32x2 sub2(const f32x2 a, const f32x2 b) GSB_F32X2_OP(sub2, al - bl, ah - bh)
#else
__device__ __forceinline__ f32x2 pk(const float lo, const float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(const f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(const f32x2 a, const f32x2 b, const f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(const f32x2 a, const f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(const f32x2 a, const f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(const f32x2 a, const f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
#endif
__device__ __forceinline__ f32x2 pk1(const float v) { return pk(v, v); }
__device__ __forceinline__ float lo_of(const f32x2 v) { float a, b; unpk(v, a, b); return a; }
__device__ __forceinline__ float hi_of(const f32x2 v) { float a, b; unpk(v, a, b); return b; }
__device__ __forceinline__ float hsum(const f32x2 v) { float a, b; unpk(v, a, b); return a + b; }

}  // namespace gsb
