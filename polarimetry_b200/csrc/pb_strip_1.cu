// pb_strip_1.cu -- instantiates the strip-march kernel (pb_strip2.cuh): 1PF, the reference's standard 18 angles (PP:1926), 3x3 window.
#include "pb_strip2.cuh"

cudaError_t pb_strip_nh1(const PixArgs &a, const Basis1 &bs, int bw, int n_stacks, unsigned char *redo, cudaStream_t st) {
    const bool same = a.sub == a.dark;
    // FAST: whole-image thresholding, float32 maps + mask for every stack
    const bool fast = same && a.simple && a.all_lean && !(a.flags & PB_FLAG_PROJECT_ONLY);
    if (bw != 3) return cudaErrorInvalidConfiguration;
    if (fast) return strip2::launch<3, 18, true, true>(a, bs, n_stacks, redo, st);
    return same ? strip2::launch<3, 18, true, false>(a, bs, n_stacks, redo, st) : strip2::launch<3, 18, false, false>(a, bs, n_stacks, redo, st);
}
