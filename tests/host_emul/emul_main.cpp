// emul_main.cpp -- TEST INFRASTRUCTURE: compiles the UNMODIFIED kernel sources of csrc/optim.cu and csrc/densify.cu (their
// two #include lines removed by the test, which writes optim_body.inc / densify_body.inc next to this file's copy) against
// cuda_shim.h and exports thin C entry points for ctypes.
#include "cuda_shim.h"

namespace gsb {
int launch_adam_step(int64_t P, int sh_coeffs, float *params, const float *grads, float *m, float *v, float *act,
                     const uint8_t *visible, const float step_size[6], float beta1, float beta2, float eps, float bias2_sqrt,
                     cudaStream_t stream);
}
namespace gsb {
size_t scan_u32_partials(int64_t n);
int scan_u32_exclusive(const uint32_t *in, uint32_t *out, size_t n, uint32_t *partials, uint32_t *total, cudaStream_t stream);
}
namespace gsb {
size_t sort_scratch_bytes(int64_t n, int V = 1);
int sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_alt, uint32_t *vals_alt, int64_t n, const unsigned long long *n_dev,
               int begin_bit, int end_bit, void *scratch, bool debug, cudaStream_t stream, int V = 1, size_t sv = 0);
}
#include "radix_sort_body.inc"
#include "optim_body.inc"
#include "densify_body.inc"
#include "knn_body.inc"

extern "C" {
int emul_adam_step(int64_t P, int M, float *params, const float *grads, float *m, float *v, float *act, const uint8_t *visible,
                   const float *step_size, float b1, float b2, float eps, float bc2s) {
    return gsb::launch_adam_step(P, M, params, grads, m, v, act, visible, step_size, b1, b2, eps, bc2s, nullptr);
}
int emul_activate(int64_t P, int M, const float *params, float *act) { return gsb::launch_activate(P, M, params, act, nullptr); }
size_t emul_densify_scratch_bytes(int64_t P, int N) { return gsb::densify_scratch_bytes(P, N); }
int emul_densify_plan(int64_t P, int M, int N, const float *params, const float *accum, const float *denom, float thr, float size_limit,
                      float min_opacity, float world_limit, void *scratch, int64_t *counts) {
    uint32_t *dev = nullptr;
    const int e = gsb::launch_densify_plan(P, M, N, params, accum, denom, thr, size_limit, min_opacity, world_limit, scratch, &dev, nullptr);
    if (e) return e;
    for (int k = 0; k < 4; ++k) counts[k] = dev[k];
    return dev[4] == dev[1] ? 0 : 2;
}
size_t emul_knn_scratch_bytes(int64_t P) { return gsb::knn_scratch_bytes(P); }
int emul_knn_mean_dist2(const float *points, int64_t P, float *out, void *scratch) {
    return gsb::launch_knn_mean_dist2(points, P, out, scratch, nullptr);
}
size_t emul_sort_scratch_bytes(int64_t n, int V) { return gsb::sort_scratch_bytes(n, V); }
// V independent sorts (view batch) of n pairs each, view v at element offset v * sv; n_dev (optional) = per-view counts
int emul_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_alt, uint32_t *vals_alt, int64_t n, const unsigned long long *n_dev,
                    int begin_bit, int end_bit, void *scratch, int V, size_t sv, int variant, int force_small, int big_ipt) {
    gsb::g_sort_variant = variant;
    gsb::g_sort_force_small = force_small;
    gsb::g_sort_big_ipt = big_ipt;
    return gsb::sort_pairs(keys, vals, keys_alt, vals_alt, n, n_dev, begin_bit, end_bit, scratch, false, nullptr, V, sv);
}
int emul_exclusive_scan(const uint32_t *in, uint32_t *out, size_t n, uint32_t *partials, uint32_t *total) {
    return gsb::exclusive_scan_u32(in, out, n, partials, total, nullptr);
}
size_t emul_scan_partials(size_t n) { return gsb::partial_count((int64_t)n); }
int emul_densify_apply(int64_t P, int M, int N, const float *params, const float *m, const float *v, void *scratch, const float *unit,
                       int64_t n_split, int64_t P_new, float *np, float *nm, float *nv) {
    return gsb::launch_densify_apply(P, M, N, params, m, v, scratch, unit, n_split, P_new, np, nm, nv, nullptr);
}
}
