// i2f64.cu -- int32 -> fp64 on B200 (sm_100a): I2F.F64.S32 against the magic-number conversion (hi word 0x43300000 + DADD), alone and
// mixed with DFMA work. Prints warp-instructions per clock per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o i2f64 i2f64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void k(double *out, int iters, int seed, long long *cyc) {
    int z[ILP];
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { z[i] = threadIdx.x * 7 + i + seed; acc[i] = 0.0; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            double d;
            if (MODE == 0) d = __int2double_rn(z[i]);                                                     // I2F.F64.S32
            else d = __hiloint2double(0x43300000, z[i]) - 4503599627370496.0;                            // MOV + DADD
            if (MODE == 2) d = __int2double_rn(z[i]);
            acc[i] = (MODE == 2) ? __fma_rn(d, 1.0000001, acc[i]) : acc[i] + d;                           // consumer keeps the conversion alive
            z[i] = (z[i] * 3 + it) & 0xfffff;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int ILP>
void run(const char *name, int warps_per_sm) {
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int threads = 32 * warps_per_sm, blocks = p.multiProcessorCount;
    if (threads > 1024) { blocks *= threads / 1024; threads = 1024; }
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<MODE, ILP><<<blocks, threads>>>(out, 16, 1, cyc);
    k<MODE, ILP><<<blocks, threads>>>(out, iters, 1, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s ILP=%d warps/SM=%2d : %.3f conversions (warp-inst)/clk/SM, %.2f clk per loop body per warp\n", name, ILP, warps_per_sm,
           (double)iters * ILP * warps_per_sm / (double)c, (double)c / iters);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 4>("I2F.F64.S32 + DADD", 8); run<0, 4>("I2F.F64.S32 + DADD", 32);
    run<1, 4>("magic MOV+DADD + DADD", 8); run<1, 4>("magic MOV+DADD + DADD", 32);
    run<2, 4>("I2F.F64.S32 + DFMA", 32);
    return 0;
}
