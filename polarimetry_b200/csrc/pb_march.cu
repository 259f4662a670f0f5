// pb_march.cu -- fused single-pass analysis WITH spatial binning (placeholder until the row-march kernel lands).
#include "pb_dev.cuh"
bool pb_march_supported(const pb_params *) { return false; }
cudaError_t pb_launch_march(const PixArgs &, const Basis2 &, const pb_params *, int, cudaStream_t, long *) { return cudaErrorNotSupported; }
