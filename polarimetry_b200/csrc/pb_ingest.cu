// pb_ingest.cu -- "next" row N3: the numeric part of Polarimetry.define_stack (PyPOLAR.py:1923-1942), and the
// layout conversion for angle-minor [H,W,N] stacks.
//
// A TIFF stack that is not of an integer type is rescaled to uint16 before anything else sees it (PP:1929-1930):
//     stack_vals = (65535 * (a - np.amin(a)) / np.ptp(a)).astype(np.uint16)
// evaluated by NumPy in the array's own precision (float32 TIFFs stay float32: the Python int 65535 is a weak scalar),
// one separately rounded operation at a time: d = a - amin ; m = 65535 * d ; q = m / ptp ; truncation to uint16.
// Two kernels: a min / max reduction over the whole stack (order-independent, so any reduction tree is exact) and the
// element-wise rescale. HBM-bound: the stack is read twice, the uint16 stack written once.
#include "pb_dev.cuh"

namespace {

// order-preserving map float <-> unsigned so that min / max can use integer atomics
__device__ __forceinline__ unsigned ord32(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float unord32(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__device__ __forceinline__ unsigned long long ord64(double f) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(f);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unord64(unsigned long long u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

// mm[0] = ordered min, mm[1] = ordered max (64-bit slots for both types), mm[2] = number of NaNs seen
template <typename T>
__global__ void __launch_bounds__(256) k_minmax(const T *__restrict__ a, long long n, unsigned long long *__restrict__ mm) {
    __shared__ unsigned long long smin[8], smax[8];
    __shared__ unsigned snan[8];
    unsigned long long lo = ~0ull, hi = 0ull;
    unsigned nn = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const T v = a[p];
        if (v != v) { ++nn; continue; }
        unsigned long long o;
        if constexpr (sizeof(T) == 4) o = ord32(v); else o = ord64(v);
        lo = o < lo ? o : lo;
        hi = o > hi ? o : hi;
    }
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long l2 = __shfl_down_sync(0xffffffffu, lo, s), h2 = __shfl_down_sync(0xffffffffu, hi, s);
        lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
        nn += __shfl_down_sync(0xffffffffu, nn, s);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { smin[w] = lo; smax[w] = hi; snan[w] = nn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { lo = smin[k] < lo ? smin[k] : lo; hi = smax[k] > hi ? smax[k] : hi; nn += snan[k]; }
        atomicMin(mm + 0, lo);
        atomicMax(mm + 1, hi);
        if (nn) atomicAdd(mm + 2, (unsigned long long)nn);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_rescale(const T *__restrict__ a, long long n, const unsigned long long *__restrict__ mm,
                                                 uint16_t *__restrict__ out) {
    T amin, amax;
    if constexpr (sizeof(T) == 4) { amin = unord32((unsigned)mm[0]); amax = unord32((unsigned)mm[1]); }
    else { amin = unord64(mm[0]); amax = unord64(mm[1]); }
    const T ptp = amax - amin;                                   // np.ptp: max - min in the array's precision
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const T d = a[p] - amin;
        const T m = (T)65535 * d;
        const T q = m / ptp;                                     // IEEE division (-prec-div default), no contraction (-fmad=false)
        // astype(np.uint16): C truncation toward zero; q is in [0, 65535.0x] here
        out[p] = (uint16_t)(q >= (T)65535 ? 65535u : (q > (T)0 ? (unsigned)q : 0u));
    }
}

// Angle-minor stacks [H,W,N] (pixel-major: the N angles of a pixel are contiguous -- the layout of interleaved acquisitions and
// the "north-star" variant of SURVEY 8b) -> the reference's angle-major [N,H,W]. One CTA = PPB consecutive pixels: their PPB*N
// elements are one contiguous span, read with fully coalesced loads into shared memory, then written plane by plane (PPB
// consecutive elements per plane). Reads and writes the stack once.
constexpr int PPB = 64;
template <typename T>
__global__ void __launch_bounds__(256) k_hwn_to_nhw(const T *__restrict__ in, T *__restrict__ out, long long P, int N) {
    extern __shared__ __align__(16) unsigned char tsm[];
    T *tile = reinterpret_cast<T *>(tsm);
    const long long p0 = (long long)blockIdx.x * PPB;
    const int np = (int)((P - p0) < PPB ? (P - p0) : PPB);
    const int cnt = np * N;
    const T *src = in + p0 * N;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) tile[k] = src[k];
    __syncthreads();
    for (int k = threadIdx.x; k < N * PPB; k += blockDim.x) {
        const int n = k / PPB, q = k - n * PPB;
        if (q < np) out[(long long)n * P + p0 + q] = tile[q * N + n];
    }
}

}  // namespace

cudaError_t pb_launch_hwn_to_nhw(const void *in, void *out, int elem_bytes, long long P, int N, cudaStream_t st) {
    const unsigned blocks = (unsigned)((P + PPB - 1) / PPB);
    const size_t smem = (size_t)PPB * N * elem_bytes;
    switch (elem_bytes) {
    case 1: k_hwn_to_nhw<uint8_t><<<blocks, 256, smem, st>>>((const uint8_t *)in, (uint8_t *)out, P, N); break;
    case 2: k_hwn_to_nhw<uint16_t><<<blocks, 256, smem, st>>>((const uint16_t *)in, (uint16_t *)out, P, N); break;
    case 4: k_hwn_to_nhw<float><<<blocks, 256, smem, st>>>((const float *)in, (float *)out, P, N); break;
    case 8: {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_hwn_to_nhw<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k_hwn_to_nhw<double><<<blocks, 256, smem, st>>>((const double *)in, (double *)out, P, N);
        break;
    }
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// mm_dev: 3 x unsigned long long workspace. Enqueues the min/max pass and the rescale; *mm_dev keeps (min, max, #NaN) for the caller.
cudaError_t pb_launch_rescale(const void *a, int is_f64, long long n, unsigned long long *mm_dev, uint16_t *out, cudaStream_t st) {
    const unsigned long long init[3] = {~0ull, 0ull, 0ull};
    cudaError_t e = cudaMemcpyAsync(mm_dev, init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    long long want = (n + 255) / 256;
    const int blocks = (int)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16);
    if (is_f64) {
        k_minmax<double><<<blocks, 256, 0, st>>>((const double *)a, n, mm_dev);
        k_rescale<double><<<blocks, 256, 0, st>>>((const double *)a, n, mm_dev, out);
    } else {
        k_minmax<float><<<blocks, 256, 0, st>>>((const float *)a, n, mm_dev);
        k_rescale<float><<<blocks, 256, 0, st>>>((const float *)a, n, mm_dev, out);
    }
    return cudaGetLastError();
}
