// pb_compact.cu -- "next" row N4: the masked-value compaction behind the result writers.
//
// save_mat (PyPOLAR.py:2829-2842) stores, per variable, var.values[mask & isfinite(var.values)].astype(float16);
// save_csv (PP:2794-2827) stores xmap[filter], ymap[filter] and var.values[filter] with filter = mask & isfinite(vars[0]).
// Both are boolean-mask gathers in row-major order, i.e. a STABLE stream compaction of maps the GPU already holds:
//   k_compact_count   : kept elements per tile of 2048 pixels
//   k_compact_scan    : exclusive scan of the tile counts (one CTA; 32768 tiles for an 8192x8192 frame)
//   k_compact_scatter : in-tile ballot scan, writes the float16 values (one correctly rounded conversion from the fp64 value,
//                       as numpy's astype does), the full-precision values and the linear pixel indices (x = p % W, y = p / W).
// HBM-bound: the map (+ filter map + selection) is read twice, the kept values written once.
#include <cuda_fp16.h>
#include "pb_dev.cuh"

namespace {

constexpr int CT = 256;            // threads per CTA
constexpr int CI = 8;              // consecutive-in-stride items per thread (item k of thread t = tile*2048 + k*256 + t)
constexpr int TILE = CT * CI;

template <typename T>
__device__ __forceinline__ bool keep(const T *__restrict__ vals, const T *__restrict__ filt, const uint8_t *__restrict__ sel,
                                     const uint32_t *__restrict__ bits, uint32_t group, long long p) {
    if (sel && !sel[p]) return false;
    if (bits && !(bits[p] & group)) return false;
    return pb_finite((double)(filt ? filt[p] : vals[p]));
}

template <typename T>
__global__ void __launch_bounds__(CT) k_compact_count(const T *__restrict__ vals, const T *__restrict__ filt, const uint8_t *__restrict__ sel,
                                                      const uint32_t *__restrict__ bits, uint32_t group, long long n,
                                                      unsigned *__restrict__ tile_count) {
    __shared__ unsigned wsum[CT / 32];
    const long long base = (long long)blockIdx.x * TILE;
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < CI; ++k) {
        const long long p = base + k * CT + threadIdx.x;
        c += (p < n && keep(vals, filt, sel, bits, group, p)) ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned s = 0;
        for (int w = 0; w < CT / 32; ++w) s += wsum[w];
        tile_count[blockIdx.x] = s;
    }
}

// exclusive scan of the tile counts, in place (64-bit offsets: a frame can keep more than 2^32 / few elements only in theory, but
// the running total is what pb_compact returns); total -> offs[ntiles]
__global__ void __launch_bounds__(1024) k_compact_scan(const unsigned *__restrict__ tile_count, unsigned long long *__restrict__ offs, int ntiles) {
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < ntiles; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const unsigned long long v = i < ntiles ? tile_count[i] : 0ull;
        unsigned long long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wtot[w] = x;
        __syncthreads();
        if (w == 0) {
            unsigned long long t = wtot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            wtot[lane] = t;                                     // inclusive totals of the warps
        }
        __syncthreads();
        const unsigned long long before = carry + (w ? wtot[w - 1] : 0ull) + (x - v);
        if (i < ntiles) offs[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offs[ntiles] = carry;
}

template <typename T>
__global__ void __launch_bounds__(CT) k_compact_scatter(const T *__restrict__ vals, const T *__restrict__ filt, const uint8_t *__restrict__ sel,
                                                        const uint32_t *__restrict__ bits, uint32_t group, long long n,
                                                        const unsigned long long *__restrict__ offs, __half *__restrict__ out_f16,
                                                        T *__restrict__ out_val, uint32_t *__restrict__ out_index) {
    __shared__ unsigned wcnt[CI][CT / 32];
    __shared__ unsigned wbase[CI][CT / 32];
    const long long base = (long long)blockIdx.x * TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    bool kp[CI];
    unsigned ball[CI];
#pragma unroll
    for (int k = 0; k < CI; ++k) {
        const long long p = base + k * CT + threadIdx.x;
        kp[k] = p < n && keep(vals, filt, sel, bits, group, p);
        ball[k] = __ballot_sync(0xffffffffu, kp[k]);
        if (lane == 0) wcnt[k][w] = __popc(ball[k]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                     // 64 entries, row-major order = (k, warp)
        unsigned s = 0;
        for (int k = 0; k < CI; ++k)
            for (int q = 0; q < CT / 32; ++q) { wbase[k][q] = s; s += wcnt[k][q]; }
    }
    __syncthreads();
    const unsigned long long o0 = offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < CI; ++k) {
        if (kp[k]) {
            const long long p = base + k * CT + threadIdx.x;
            const unsigned long long e = o0 + wbase[k][w] + __popc(ball[k] & ((1u << lane) - 1u));
            const T v = vals[p];
            if (out_f16) out_f16[e] = __double2half((double)v);   // cvt.rn.f16.f64: one rounding, like numpy's astype(float16)
            if (out_val) out_val[e] = v;
            if (out_index) out_index[e] = (uint32_t)p;
        }
    }
}

}  // namespace

size_t pb_compact_ws_bytes(long long n) {
    const long long ntiles = (n + TILE - 1) / TILE;
    return (size_t)ntiles * 4 + 16 + (size_t)(ntiles + 1) * 8;
}

// ws: [ntiles] unsigned counts (padded to 16 B), then [ntiles + 1] unsigned long long offsets; the total lands in offs[ntiles]
cudaError_t pb_launch_compact(const void *vals, const void *filt, int is_f64, const uint8_t *sel, const uint32_t *bits, uint32_t group,
                              long long n, void *ws, void *out_f16, void *out_val, uint32_t *out_index,
                              const unsigned long long **total_dev, cudaStream_t st) {
    const int ntiles = (int)((n + TILE - 1) / TILE);
    unsigned *cnt = reinterpret_cast<unsigned *>(ws);
    unsigned long long *offs = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(ws) + (((size_t)ntiles * 4 + 15) & ~(size_t)15));
    if (is_f64) k_compact_count<double><<<ntiles, CT, 0, st>>>((const double *)vals, (const double *)filt, sel, bits, group, n, cnt);
    else k_compact_count<float><<<ntiles, CT, 0, st>>>((const float *)vals, (const float *)filt, sel, bits, group, n, cnt);
    k_compact_scan<<<1, 1024, 0, st>>>(cnt, offs, ntiles);
    if (is_f64) k_compact_scatter<double><<<ntiles, CT, 0, st>>>((const double *)vals, (const double *)filt, sel, bits, group, n, offs,
                                                               (__half *)out_f16, (double *)out_val, out_index);
    else k_compact_scatter<float><<<ntiles, CT, 0, st>>>((const float *)vals, (const float *)filt, sel, bits, group, n, offs, (__half *)out_f16,
                                                        (float *)out_val, out_index);
    *total_dev = offs + ntiles;
    return cudaGetLastError();
}
