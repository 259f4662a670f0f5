// pb_march_u8_1.cu -- instantiates the fused row-march kernel (pb_march_impl.cuh) for uint8_t stacks, 1 harmonic(s).
#include "pb_march_impl.cuh"

cudaError_t pb_march_u8_nh1(const PixArgs &a, const Basis1 &bs, const pb_params *p, int n_stacks, cudaStream_t st) {
    return launch_nh<uint8_t, 1>(a, bs, p, n_stacks, st);
}
