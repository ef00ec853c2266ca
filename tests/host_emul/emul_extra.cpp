// emul_extra.cpp -- TEST INFRASTRUCTURE (see cuda_shim.h): C entry points for internals of the host build that the public
// C ABI does not expose (the view-batch / block-size arguments of the sort, the bare exclusive scan).
#include "kernels.cuh"

extern "C" {
int emul_exclusive_scan(const uint32_t *in, uint32_t *out, size_t n, uint32_t *partials, uint32_t *total) {
    return gsb::scan_u32_exclusive(in, out, n, partials, total, nullptr);
}
size_t emul_scan_partials(size_t n) { return gsb::scan_u32_partials((int64_t)n); }
size_t emul_sort_scratch_bytes(int64_t n, int V) { return gsb::sort_scratch_bytes(n, V); }
// V independent sorts (view batch) of n pairs each, view v at element offset v * sv; n_dev (optional) = per-view counts
int emul_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_alt, uint32_t *vals_alt, int64_t n, const unsigned long long *n_dev,
                    int begin_bit, int end_bit, void *scratch, int V, size_t sv, int force_small) {
    gsb::g_sort_force_small = force_small;
    const int rc = gsb::sort_pairs(keys, vals, keys_alt, vals_alt, n, n_dev, begin_bit, end_bit, scratch, false, nullptr, V, sv);
    gsb::g_sort_force_small = 0;
    return rc;
}
}
