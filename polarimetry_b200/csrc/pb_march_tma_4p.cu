// pb_march_tma_4p.cu -- instantiates the TMA row-march kernel (pb_march_tma.cuh) for 4POLAR 2D / 3D with spatial binning: the four
// channel planes are box-filtered by the march (PP:2942-2946) and each pixel goes through the polarization-matrix inversion of
// PP:3001-3014 / 3041-3051 (pb_pixel4.cuh).
#include "pb_march_tma.cuh"

cudaError_t pb_march_tma_4polar(const PixArgs &a, int bw, int n_stacks, cudaStream_t st) {
    static const Basis1 none = {};
    return tma::launch_nh<0>(a, none, bw, n_stacks, st);
}
