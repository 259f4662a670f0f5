// pb_march_tma_1.cu -- instantiates the TMA row-march kernel (pb_march_tma.cuh) for 1 harmonic (1PF).
#include "pb_march_tma.cuh"

cudaError_t pb_march_tma_nh1(const PixArgs &a, const Basis1 &bs, int bw, int n_stacks, cudaStream_t st) {
    return tma::launch_nh<1>(a, bs, bw, n_stacks, st);
}

// second pass after k_strip (pb_strip.cuh): the 16-row blocks flagged in `redo` only, and in them only the pixels whose fit / chi2
// decision needs the exact loop of PP:2990-3000
cudaError_t pb_march_tma_redo_nh1_bw3_n18(const PixArgs &a, const Basis1 &bs, int n_stacks, const unsigned char *redo, cudaStream_t st) {
    return tma::launch<1, 3, 16, 18>(a, bs, n_stacks, st, redo);
}
