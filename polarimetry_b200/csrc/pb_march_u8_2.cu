// pb_march_u8_2.cu -- instantiates the fused row-march kernel (pb_march_impl.cuh) for uint8_t stacks, 2 harmonic(s).
#include "pb_march_impl.cuh"

cudaError_t pb_march_u8_nh2(const PixArgs &a, const Basis2 &bs, const pb_params *p, int n_stacks, cudaStream_t st) {
    return launch_nh<uint8_t, 2>(a, bs, p, n_stacks, st);
}
