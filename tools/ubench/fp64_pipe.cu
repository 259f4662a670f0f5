// fp64 pipe micro-benchmark for B200 (sm_100a): throughput / latency of DFMA, DADD, DMUL chains at varying ILP and occupancy,
// and of the same chains interleaved with integer work. Prints warp-instructions per clock per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe fp64_pipe.cu && ./fp64_pipe
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP, int MIX>
__global__ void k(double *out, int iters, double a, double b, long long *cyc) {
    double x[ILP];
    int z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = a + i + threadIdx.x; z[i] = threadIdx.x + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = __fma_rn(x[i], a, b);
            if (OP == 1) x[i] = __dadd_rn(x[i], b);
            if (OP == 2) x[i] = __dmul_rn(x[i], a);
            if (MIX) {
#pragma unroll
                for (int m = 0; m < MIX; ++m) z[i] = max(z[i] - (int)threadIdx.x, m) + (z[i] >> 3);
            }
        }
    }
    long long t1 = clock64();
    double s = 0; int zz = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { s += x[i]; zz += z[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + zz;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int ILP, int MIX>
void run(const char *name, int warps_per_sm) {
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int threads = 32 * warps_per_sm;
    int blocks = p.multiProcessorCount;
    if (threads > 1024) { blocks *= threads / 1024; threads = 1024; }
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<OP, ILP, MIX><<<blocks, threads>>>(out, 16, 1.0000001, 1e-9, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP, ILP, MIX><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double winst = (double)iters * ILP * warps_per_sm;      // fp64 warp-instructions per SM
    printf("%-6s ILP=%d mix=%d warps/SM=%2d : %.3f fp64 warp-inst/clk/SM (%.1f cyc per dependent op), %.3f ms, %.0f MHz eff\n", name, ILP, MIX, warps_per_sm,
           winst / (double)c, (double)c / iters, ms, (double)c / (ms * 1e3));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1, 0>("DFMA", 1); run<0, 1, 0>("DFMA", 4); run<0, 2, 0>("DFMA", 4); run<0, 4, 0>("DFMA", 4); run<0, 8, 0>("DFMA", 4);
    run<0, 4, 0>("DFMA", 8); run<0, 4, 0>("DFMA", 16); run<0, 4, 0>("DFMA", 32); run<0, 8, 0>("DFMA", 32);
    run<1, 1, 0>("DADD", 1); run<1, 4, 0>("DADD", 16); run<1, 4, 0>("DADD", 32);
    run<2, 1, 0>("DMUL", 1); run<2, 4, 0>("DMUL", 16); run<2, 4, 0>("DMUL", 32);
    run<0, 4, 1>("DFMA", 16); run<0, 4, 2>("DFMA", 16); run<0, 4, 4>("DFMA", 16); run<0, 4, 2>("DFMA", 32); run<0, 4, 4>("DFMA", 32);
    return 0;
}
