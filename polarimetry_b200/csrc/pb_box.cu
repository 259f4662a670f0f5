// pb_box.cu -- the unfused pre-processing kernels: dark subtraction / clamp / sqrt / channel swap,
// total-intensity sum, and the exact two-pass running-sum box filter.
//
// Replaces analyze_stack's pre-processing (PyPOLAR.py:2953-2957), Polarimetry.binning (PP:2942-2946) and
// Stack.get_intensity (pypolar_classes.py:260-264). The box filter replays scipy.ndimage.uniform_filter1d
// (mode='reflect', origin 0) op for op: per line tmp = sum(ext[0..n-1]) left to right, out[0] = tmp/n, then
// tmp += ext[k+n-1] - ext[k-1]; out[k] = tmp/n -- axis 1 (H) first, then axis 2 (W). The recurrence is
// sequential along each line by construction (its rounding history is part of the reference's result), so
// the parallelism is over lines: one thread per column for the H pass, one thread per row for the W pass,
// the latter staged through shared memory so that global traffic stays coalesced.
#include "pb_dev.cuh"
#include <cstdlib>

void pb_ddrecip(int n, double *hi, double *lo);   // pb_api.cu

namespace {

template <typename T>
__global__ void k_field(const T *__restrict__ v, double *__restrict__ out, long long P, int N, double sub, int do_sqrt, int swap12) {
    const long long total = P * N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long i = e / P, p = e - i * P;
        long long src = (swap12 && (i == 1 || i == 2)) ? 3 - i : i;
        double x = pb_max0((double)__ldg(v + src * P + p) - sub);
        out[e] = do_sqrt ? sqrt(x) : x;
    }
}

template <typename T>
__global__ void k_isum(const T *__restrict__ v, double *__restrict__ out, long long P, int N, double dark) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        double s = pb_max0((double)__ldg(v + p) - dark);
        for (int i = 1; i < N; ++i) s = s + pb_max0((double)__ldg(v + (long long)i * P + p) - dark);
        out[p] = s;
    }
}

// tmp / n, correctly rounded without a division: fma(x, rhi, x*rlo) with (rhi, rlo) the double-double 1/n (tests/test_shortcuts.py:
// x/n is never within 2^-60 of a rounding boundary, the approximation error is 2^-105); values outside the range where neither
// product can underflow or overflow (never produced by image data) take the IEEE division
__device__ __forceinline__ double div_small(double x, double rhi, double rlo, double dn) {
    const double ax = fabs(x);
    return ((ax > 1e-280 && ax < 1e280) || x == 0.0) ? fma(x, rhi, x * rlo) : x / dn;
}

// H pass: thread = (plane, x); walks y, four rows at a time so that their eight loads are in flight together. Coalesced across x.
// Out of place.
__global__ void k_box_y(const double *__restrict__ in, double *__restrict__ out, int planes, int H, int W, int n, double rhi, double rlo) {
    const long long cols = (long long)planes * W;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const long long pl = c / W;
    const int x = (int)(c - pl * W);
    const double *src = in + pl * (long long)H * W + x;
    double *dst = out + pl * (long long)H * W + x;
    const int h = n / 2;
    const double dn = (double)n;
    double tmp = 0.0;
    for (int k = 0; k < n; ++k) tmp = tmp + src[(long long)pb_reflect(k - h, H) * W];
    dst[0] = div_small(tmp, rhi, rlo, dn);
    int k = 1;
    for (; k + 3 < H; k += 4) {
        double nw[4], od[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            nw[j] = src[(long long)pb_reflect(k + j + n - 1 - h, H) * W];
            od[j] = src[(long long)pb_reflect(k + j - 1 - h, H) * W];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tmp = tmp + (nw[j] - od[j]);
            dst[(long long)(k + j) * W] = div_small(tmp, rhi, rlo, dn);
        }
    }
    for (; k < H; ++k) {
        double nw = src[(long long)pb_reflect(k + n - 1 - h, H) * W];
        double od = src[(long long)pb_reflect(k - 1 - h, H) * W];
        tmp = tmp + (nw - od);
        dst[(long long)k * W] = div_small(tmp, rhi, rlo, dn);
    }
}

// W pass: thread = one row (plane*H + y); the extended line ext[k] = line[reflect(k - h)] is staged through a
// 64-entry per-row ring in shared memory, 32 entries (one coalesced 256-byte segment per row) at a time.
constexpr int BX_ROWS = 64;   // rows per block = threads per block
constexpr int BX_RING = 64;   // ring entries per row (n <= 32)
constexpr int BX_LD = BX_RING + 1;

__global__ void __launch_bounds__(BX_ROWS) k_box_x(const double *__restrict__ in, double *__restrict__ out, long long rows, int W, int n, double rhi, double rlo) {
    extern __shared__ __align__(16) unsigned char bx_smem[];
    double *ring = reinterpret_cast<double *>(bx_smem);   // [BX_ROWS][BX_LD]
    double *obuf = ring + BX_ROWS * BX_LD;                 // [BX_ROWS][33]
    const long long row0 = (long long)blockIdx.x * BX_ROWS;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int h = n / 2;
    const int E = W + n - 1;   // extended length
    const bool live = row0 + t < rows;
    const double dn = (double)n;
    double tmp = 0.0;
    const int ntiles = (E + 31) / 32;
    int jbeg = 0;              // first output index not yet produced (identical for every row)
    for (int m = 0; m < ntiles; ++m) {
        // stage ext[32m .. 32m+31] of every row: warp w loads rows w, w+2, ... (one coalesced segment per row)
        for (int r = warp; r < BX_ROWS; r += BX_ROWS / 32) {
            long long rr = row0 + r;
            int k = 32 * m + lane;
            if (rr < rows && k < E) ring[r * BX_LD + (k & (BX_RING - 1))] = __ldg(in + rr * W + pb_reflect(k - h, W));
        }
        __syncthreads();
        // outputs whose entering element ext[j+n-1] is now staged: j <= kmax - n + 1
        const int kmax = min(32 * m + 31, E - 1);
        const int jend = max(jbeg, min(W, kmax - n + 2));
        if (live) {
            double *ob = obuf + t * 33;
            const double *rg = ring + t * BX_LD;
            for (int j = jbeg; j < jend; ++j) {
                if (j == 0) {
                    for (int k = 0; k < n; ++k) tmp = tmp + rg[k & (BX_RING - 1)];
                } else {
                    tmp = tmp + (rg[(j + n - 1) & (BX_RING - 1)] - rg[(j - 1) & (BX_RING - 1)]);
                }
                ob[j - jbeg] = div_small(tmp, rhi, rlo, dn);
            }
        }
        __syncthreads();
        const int cnt = jend - jbeg;   // <= 32
        for (int r = warp; r < BX_ROWS; r += BX_ROWS / 32) {
            long long rr = row0 + r;
            if (rr < rows && lane < cnt) out[rr * W + jbeg + lane] = obuf[r * 33 + lane];
        }
        jbeg = jend;
        // the next tile overwrites ring slots of tile m-1 and obuf: all reads above are done after this barrier
        __syncthreads();
    }
}

// W pass, one thread per row reading and writing its row directly (no shared-memory staging, no block barriers): consecutive steps of a
// thread touch consecutive doubles, so every 32-byte sector it fetches or writes is used in full (4 steps); the entering and the
// leaving element of the window are loaded four steps at a time. Same operations in the same order as k_box_x.
__global__ void k_box_x_rows(const double *__restrict__ in, double *__restrict__ out, long long rows, int W, int n, double rhi, double rlo) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double *src = in + r * W;
    double *dst = out + r * W;
    const int h = n / 2;
    const double dn = (double)n;
    double tmp = 0.0;
    for (int k = 0; k < n; ++k) tmp = tmp + src[pb_reflect(k - h, W)];
    dst[0] = div_small(tmp, rhi, rlo, dn);
    // output j: entering element ext[j+n-1] = src[reflect(j+n-1-h)], leaving element ext[j-1] = src[reflect(j-1-h)]
    auto step = [&](int j) {
        const double nw = src[pb_reflect(j + n - 1 - h, W)], od = src[pb_reflect(j - 1 - h, W)];
        tmp = tmp + (nw - od);
        dst[j] = div_small(tmp, rhi, rlo, dn);
    };
    int j = 1;
    const int j_in = min(h + 1, W);                 // first output whose two elements lie inside the row without reflection
    for (; j < j_in; ++j) step(j);
    const int j_last = W - n + h;                   // last such output
    for (; j + 3 <= j_last; j += 4) {
        double nw[4], od[4];
        const double *pn = src + (j + n - 1 - h), *po = src + (j - 1 - h);
#pragma unroll
        for (int q = 0; q < 4; ++q) { nw[q] = pn[q]; od[q] = po[q]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            tmp = tmp + (nw[q] - od[q]);
            dst[j + q] = div_small(tmp, rhi, rlo, dn);
        }
    }
    for (; j < W; ++j) step(j);
}

}  // namespace

cudaError_t pb_launch_field(const void *v, int dtype, double *out, long long P, int N, double sub, int do_sqrt, int swap12, cudaStream_t st) {
    const int threads = 256;
    long long total = P * N;
    int blocks = (int)((total + threads - 1) / threads < 148LL * 16 ? (total + threads - 1) / threads : 148LL * 16);
    if (blocks < 1) blocks = 1;
    switch (dtype) {
    case PB_U8: k_field<uint8_t><<<blocks, threads, 0, st>>>((const uint8_t *)v, out, P, N, sub, do_sqrt, swap12); break;
    case PB_U16: k_field<uint16_t><<<blocks, threads, 0, st>>>((const uint16_t *)v, out, P, N, sub, do_sqrt, swap12); break;
    case PB_F32: k_field<float><<<blocks, threads, 0, st>>>((const float *)v, out, P, N, sub, do_sqrt, swap12); break;
    case PB_F64: k_field<double><<<blocks, threads, 0, st>>>((const double *)v, out, P, N, sub, do_sqrt, swap12); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t pb_launch_isum(const void *v, int dtype, double *out, long long P, int N, double dark, cudaStream_t st) {
    const int threads = 256;
    int blocks = (int)((P + threads - 1) / threads < 148LL * 16 ? (P + threads - 1) / threads : 148LL * 16);
    if (blocks < 1) blocks = 1;
    switch (dtype) {
    case PB_U8: k_isum<uint8_t><<<blocks, threads, 0, st>>>((const uint8_t *)v, out, P, N, dark); break;
    case PB_U16: k_isum<uint16_t><<<blocks, threads, 0, st>>>((const uint16_t *)v, out, P, N, dark); break;
    case PB_F32: k_isum<float><<<blocks, threads, 0, st>>>((const float *)v, out, P, N, dark); break;
    case PB_F64: k_isum<double><<<blocks, threads, 0, st>>>((const double *)v, out, P, N, dark); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// in -> out along H (n > 1)
cudaError_t pb_launch_box_y(const double *in, double *out, int planes, int H, int W, int n, cudaStream_t st) {
    const int threads = 128;
    long long cols = (long long)planes * W;
    double rhi, rlo;
    pb_ddrecip(n, &rhi, &rlo);
    k_box_y<<<(unsigned)((cols + threads - 1) / threads), threads, 0, st>>>(in, out, planes, H, W, n, rhi, rlo);
    return cudaGetLastError();
}

// in -> out along W (1 < n <= 32)
cudaError_t pb_launch_box_x(const double *in, double *out, int planes, int H, int W, int n, cudaStream_t st) {
    long long rows = (long long)planes * H;
    static const bool staged = [] { const char *e = getenv("PB_BOX_X_STAGED"); return e && atoi(e) != 0; }();   // A/B: the shared-memory staged kernel
    if (!staged) {
        double rhi, rlo;
        pb_ddrecip(n, &rhi, &rlo);
        const int threads = rows >= 148LL * 256 ? 128 : 32;        // few rows (one plane): spread them over the SMs
        k_box_x_rows<<<(unsigned)((rows + threads - 1) / threads), threads, 0, st>>>(in, out, rows, W, n, rhi, rlo);
        return cudaGetLastError();
    }
    const size_t smem = sizeof(double) * BX_ROWS * (BX_LD + 33);
    cudaError_t e = cudaFuncSetAttribute(k_box_x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    double rhi, rlo;
    pb_ddrecip(n, &rhi, &rlo);
    k_box_x<<<(unsigned)((rows + BX_ROWS - 1) / BX_ROWS), BX_ROWS, smem, st>>>(in, out, rows, W, n, rhi, rlo);
    return cudaGetLastError();
}
