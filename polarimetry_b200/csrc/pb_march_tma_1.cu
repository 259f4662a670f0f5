// pb_march_tma_1.cu -- instantiates the TMA row-march kernel (pb_march_tma.cuh) for 1 harmonic (1PF).
#include "pb_march_tma.cuh"

cudaError_t pb_march_tma_nh1(const PixArgs &a, const Basis1 &bs, int bw, int n_stacks, cudaStream_t st) {
    return tma::launch_nh<1>(a, bs, bw, n_stacks, st);
}
