// tma_box.cu -- which 4-D uint16 box geometries does cp.async.bulk.tensor accept?  nvcc -arch=sm_100a -o tma_box tma_box.cu ; ./tma_box
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, int bytes, uint16_t *out, int n) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(smem), bb = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bb), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(sb),
                     "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(x), "r"(y), "r"(0), "r"(0), "r"(bb)
                     : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bb), "r"(0)
        : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(smem)[i];
}

int main() {
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int W = 88, H = 70, N = 18, S = 2;
    std::vector<uint16_t> h((size_t)S * N * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i % 65521);
    uint16_t *d, *o;
    cudaMalloc(&d, h.size() * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 65536 * 2);
    struct Cfg { int bx, by, bz; CUtensorMapSwizzle sw; int x, y; const char *name; };
    Cfg cfgs[] = {
        {16, 18, 18, CU_TENSOR_MAP_SWIZZLE_32B, 0, 0, "16x18x18 sw32 x=0"},
        {16, 18, 18, CU_TENSOR_MAP_SWIZZLE_NONE, 0, 0, "16x18x18 none x=0"},
        {8, 34, 18, CU_TENSOR_MAP_SWIZZLE_NONE, 0, 0, "8x34x18 none x=0"},
        {8, 34, 18, CU_TENSOR_MAP_SWIZZLE_NONE, 8, 0, "8x34x18 none x=8"},
        {8, 34, 18, CU_TENSOR_MAP_SWIZZLE_NONE, 1, 0, "8x34x18 none x=1"},
        {8, 34, 18, CU_TENSOR_MAP_SWIZZLE_NONE, -7, -1, "8x34x18 none x=-7 y=-1"},
        {8, 34, 18, CU_TENSOR_MAP_SWIZZLE_NONE, -8, -1, "8x34x18 none x=-8 y=-1"},
        {8, 18, 18, CU_TENSOR_MAP_SWIZZLE_NONE, 0, 0, "8x18x18 none x=0"},
        {8, 34, 9, CU_TENSOR_MAP_SWIZZLE_NONE, 0, 0, "8x34x9 none x=0"},
        {16, 34, 18, CU_TENSOR_MAP_SWIZZLE_32B, 1, -1, "16x34x18 sw32 x=1 y=-1"},
    };
    for (auto &c : cfgs) {
        CUtensorMap map;
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)S};
        const cuuint64_t strides[3] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2, (cuuint64_t)W * H * N * 2};
        const cuuint32_t box[4] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, (cuuint32_t)c.bz, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int n = c.bx * c.by * c.bz;
        if (r != CUDA_SUCCESS) { printf("%-28s encode failed %d\n", c.name, (int)r); continue; }
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        k<<<1, 32, 64 * 1024>>>(map, c.x, c.y, n * 2, o, n);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-28s kernel: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        std::vector<uint16_t> got(n);
        cudaMemcpy(got.data(), o, n * 2, cudaMemcpyDeviceToHost);
        // check element (plane 1, row 2, col 3 of the box) when the swizzle is none
        int bad = 0;
        if (c.sw == CU_TENSOR_MAP_SWIZZLE_NONE)
            for (int p = 0; p < c.bz; ++p) for (int r2 = 0; r2 < c.by; ++r2) for (int cc = 0; cc < c.bx; ++cc) {
                const int gx = c.x + cc, gy = c.y + r2;
                const uint16_t want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : h[((size_t)p * H + gy) * W + gx];
                bad += got[(p * c.by + r2) * c.bx + cc] != want;
            }
        printf("%-28s ok, mismatches %d\n", c.name, bad);
    }
    return 0;
}
