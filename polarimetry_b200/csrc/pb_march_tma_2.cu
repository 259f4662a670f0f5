// pb_march_tma_2.cu -- instantiates the TMA row-march kernel (pb_march_tma.cuh) for 2 harmonics (SRS / SHG / 2PF).
#include "pb_march_tma.cuh"

cudaError_t pb_march_tma_nh2(const PixArgs &a, const Basis2 &bs, int bw, int n_stacks, cudaStream_t st) {
    return tma::launch_nh<2>(a, bs, bw, n_stacks, st);
}
