// pb_reduce.cu -- standalone histogram, ROI statistics and fit/chi2 map kernels.
//
// pb_histo   : numeric part of Variable.histo (pypolar_classes.py:340-369) on a finished map.
// pb_stats   : the reductions of return_vecexcel (PyPOLAR.py:2744-2779): N, (circular) mean, std, mean delta.
// pb_fit_maps: field_fit / chi2 for the individual-fit viewer (PP:2990-2999, PP:3073-3077), fp64, reference order.
#include "pb_dev.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_histo(const T *__restrict__ vals, const uint8_t *__restrict__ sel, long long n, int is_rho,
                                               double rot_fig, int fold90, const double *__restrict__ edges, int nbins, double inv_w,
                                               unsigned long long *__restrict__ counts) {
    __shared__ unsigned cnt[PB_MAX_BINS];
    __shared__ double e[PB_MAX_BINS + 1];
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) cnt[k] = 0u;
    for (int k = threadIdx.x; k <= nbins; k += blockDim.x) e[k] = edges[k];
    __syncthreads();
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        if (sel && !sel[p]) continue;
        double d = (double)vals[p];
        if (!pb_finite(d)) continue;
        if (is_rho) d = pb_pymod(2.0 * (d + rot_fig), 360.0) / 2.0;
        if (fold90 && d >= 90.0) d = 180.0 - d;
        int b = pb_bin(d, e, nbins, inv_w);
        if (b >= 0) atomicAdd(&cnt[b], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nbins; k += blockDim.x)
        if (cnt[k]) atomicAdd(counts + k, (unsigned long long)cnt[k]);
}

__device__ __forceinline__ double block_sum(double v, double *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    return r;
}

constexpr double D2R = 3.141592653589793238462643383279502884 / 180.0;

template <typename T>
__device__ __forceinline__ bool stat_value(const T *vals, const T *rho, const uint8_t *sel, long long p, int circular, double rot_fig,
                                           int rewrap, int fold90, double &d) {
    if (sel && !sel[p]) return false;
    if (rho && !pb_finite((double)rho[p])) return false;
    d = (double)vals[p];
    if (circular) {
        if (!pb_finite(d)) return false;
        if (rewrap) d = pb_pymod(2.0 * (d + rot_fig), 360.0) / 2.0;
        if (fold90 && d >= 90.0) d = 180.0 - d;
    }
    return true;
}

// pass 1: acc[0] += count, acc[1] += sum d (or sum cos 2d), acc[2] += sum sin 2d
template <typename T>
__global__ void __launch_bounds__(256) k_stats1(const T *vals, const T *rho, const uint8_t *sel, long long n, int circular, double rot_fig,
                                                int rewrap, int fold90, double *acc) {
    __shared__ double sh[8];
    double c = 0.0, s1 = 0.0, s2 = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        double d;
        if (!stat_value(vals, rho, sel, p, circular, rot_fig, rewrap, fold90, d)) continue;
        c += 1.0;
        if (circular) { double sn, cs; sincos(2.0 * d * D2R, &sn, &cs); s1 += cs; s2 += sn; }
        else s1 += d;
    }
    c = block_sum(c, sh); s1 = block_sum(s1, sh); s2 = block_sum(s2, sh);
    if (threadIdx.x == 0) { atomicAdd(acc + 0, c); atomicAdd(acc + 1, s1); atomicAdd(acc + 2, s2); }
}

// pass 2: acc[3] += sum delta, acc[4] += sum delta^2, delta = wrapto180(2(d-mean))/2 (circular) or d - mean
template <typename T>
__global__ void __launch_bounds__(256) k_stats2(const T *vals, const T *rho, const uint8_t *sel, long long n, int circular, double rot_fig,
                                                int rewrap, int fold90, double mean, double *acc) {
    __shared__ double sh[8];
    double s1 = 0.0, s2 = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        double d;
        if (!stat_value(vals, rho, sel, p, circular, rot_fig, rewrap, fold90, d)) continue;
        double dl = circular ? remainder(2.0 * (d - mean), 360.0) / 2.0 : d - mean;
        s1 += dl; s2 += dl * dl;
    }
    s1 = block_sum(s1, sh); s2 = block_sum(s2, sh);
    if (threadIdx.x == 0) { atomicAdd(acc + 3, s1); atomicAdd(acc + 4, s2); }
}


// ---------------------------------------------------------------------------------------------------
// Batch statistics (return_vecexcel, PP:2739-2779, over every map of a batch): one launch per pass for all stacks,
// groups and variables, accumulators in device memory, means finalised on the device -> no host round trip between the passes,
// and the accumulators are plain sums, so a sharded batch needs one small all-reduce per pass.
struct StatsBatchArgs {
    const pb_outputs *outs;      // device array [n_stacks]
    const uint32_t *roi_bits;    // [P] or NULL
    long long P;
    int n_vars;                  // primary variables (Rho first); statistic slot n_vars is intmap
    int f64, per_stack, n_groups;
    double rot_fig;
    uint32_t group_bits[PB_MAX_ROIS];
};
constexpr int SB_SLOTS = PB_MAX_VARS + 1;

template <typename T, int PASS>
__global__ void __launch_bounds__(256) k_stats_batch(StatsBatchArgs a, double *__restrict__ acc1, double *__restrict__ acc2) {
    __shared__ double sh[8];
    const int g = blockIdx.z;
    const pb_outputs o = a.outs[blockIdx.y];
    if (a.per_stack) {                          // one accumulator set per stack instead of one for the whole batch
        acc1 += (size_t)blockIdx.y * a.n_groups * SB_SLOTS * 4;
        acc2 += (size_t)blockIdx.y * a.n_groups * SB_SLOTS * 2;
    }
    const uint32_t gb = a.group_bits[g];
    const T *maps[SB_SLOTS];
#pragma unroll
    for (int v = 0; v < SB_SLOTS; ++v) maps[v] = nullptr;
#pragma unroll
    for (int v = 0; v < PB_MAX_VARS; ++v)
        if (v < a.n_vars) maps[v] = reinterpret_cast<const T *>(o.var[v]);
    const T *im = reinterpret_cast<const T *>(o.intmap);
    double mean[SB_SLOTS];
    if (PASS == 2) {
        // the means of pass 1 (already summed over the ranks by the caller, if any): circularmean (PC:55-56) for Rho
#pragma unroll
        for (int v = 0; v < SB_SLOTS; ++v) {
            const double *q = acc1 + ((size_t)g * SB_SLOTS + v) * 4;
            const double n = q[0];
            if (v == 0) {
                double ang = atan2(q[2] / n, q[1] / n) * (180.0 / 3.141592653589793238462643383279502884);
                ang = fmod(ang, 360.0);
                if (ang < 0) ang += 360.0;
                mean[v] = ang / 2.0;
            } else {
                mean[v] = q[1] / n;
            }
        }
    }
    double cnt = 0.0, s1[SB_SLOTS], s2[SB_SLOTS];
#pragma unroll
    for (int v = 0; v < SB_SLOTS; ++v) s1[v] = s2[v] = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += (long long)gridDim.x * blockDim.x) {
        const double rho = (double)maps[0][p];
        if (!pb_finite(rho)) continue;                                            // keep = roi & isfinite(Rho), PP:2742-2743
        if (a.roi_bits && !(__ldg(a.roi_bits + p) & gb)) continue;
        cnt += 1.0;
        const double d = pb_pymod(2.0 * (rho + a.rot_fig), 360.0) / 2.0;         // PP:2744
        if (PASS == 1) {
            double sn, cs;
            sincos(2.0 * d * D2R, &sn, &cs);
            s1[0] += cs; s2[0] += sn;
        } else {
            const double dl = remainder(2.0 * (d - mean[0]), 360.0) / 2.0;       // wrapto180(2(d - mean))/2, PC:73-74
            s1[0] += dl; s2[0] += dl * dl;
        }
#pragma unroll
        for (int v = 1; v < SB_SLOTS; ++v) {
            const T *m = v < PB_MAX_VARS ? maps[v] : nullptr;
            if (v == a.n_vars) m = im;
            if (v > a.n_vars || !m) continue;
            const double x = (double)m[p];
            if (PASS == 1) s1[v] += x;
            else { const double dl = x - mean[v]; s1[v] += dl; s2[v] += dl * dl; }
        }
    }
    cnt = block_sum(cnt, sh);
#pragma unroll
    for (int v = 0; v < SB_SLOTS; ++v) {
        if (v > a.n_vars) continue;
        const double r1 = block_sum(s1[v], sh), r2 = block_sum(s2[v], sh);
        if (threadIdx.x == 0) {
            if (PASS == 1) {
                double *q = acc1 + ((size_t)g * SB_SLOTS + v) * 4;
                atomicAdd(q + 0, cnt); atomicAdd(q + 1, r1); atomicAdd(q + 2, r2);
            } else {
                double *q = acc2 + ((size_t)g * SB_SLOTS + v) * 2;
                atomicAdd(q + 0, r1); atomicAdd(q + 1, r2);
            }
        }
    }
}

template <int NH>
__global__ void __launch_bounds__(256) k_fit_maps(const double *__restrict__ f, const uint8_t *__restrict__ mask, long long P, int N,
                                                  typename BasisSel<NH>::type bs, double *__restrict__ fit, double *__restrict__ chi2o) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const bool m = mask ? mask[p] != 0 : true;
    double S0 = 0.0, Sc = 0.0, Ss = 0.0, Sc4 = 0.0, Ss4 = 0.0;
    for (int i = 0; i < N; ++i) {
        double x = f[(long long)i * P + p];
        S0 = S0 + x; Sc = Sc + x * bs.c2[i]; Ss = Ss + x * bs.s2[i];
        if constexpr (NH == 2) { Sc4 = Sc4 + x * bs.c4[i]; Ss4 = Ss4 + x * bs.s4[i]; }
    }
    const double dN = (double)N, k2 = 2.0 / dN;
    double a0 = S0 / dN;
    if (a0 == 0.0) a0 = pb_nan();
    const double a2r = k2 * Sc, a2i = k2 * Ss, a4r = k2 * Sc4, a4i = k2 * Ss4;
    double acc = 0.0;
    for (int i = 0; i < N; ++i) {
        double ft = a0 + fma(a2r, bs.c2[i], a2i * bs.s2[i]);
        if constexpr (NH == 2) ft = ft + fma(a4r, bs.c4[i], a4i * bs.s4[i]);
        double d = f[(long long)i * P + p] - ft;
        acc = acc + (d * d) / ft;
        if (fit) fit[(long long)i * P + p] = m ? ft : pb_nan();
    }
    if (chi2o) chi2o[p] = m ? acc / dN : pb_nan();
}

// Stack.compute_dark (PC:266-278), step 1: per size x size cell of the FIRST angle plane, the sum and the count of the
// non-zero samples (numpy.ma.masked_equal(.., 0).mean is sum / count, both exact integers). One warp per cell.
template <typename T>
__global__ void __launch_bounds__(256) k_dark_cells(const T *__restrict__ plane0, int W, int ny, int nx, int cell,
                                                    unsigned long long *__restrict__ sums, unsigned *__restrict__ counts) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= ny * nx) return;
    const int ci = warp / nx, cj = warp - ci * nx;
    unsigned long long s = 0ull;
    unsigned c = 0u;
    for (int e = lane; e < cell * cell; e += 32) {
        const int dy = e / cell, dx = e - dy * cell;
        const unsigned v = (unsigned)plane0[(long long)(ci * cell + dy) * W + cj * cell + dx];
        s += v; c += v != 0u;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); c += __shfl_down_sync(0xffffffffu, c, o); }
    if (lane == 0) { sums[warp] = s; counts[warp] = c; }
}

// step 2: the same sum / count over ALL angle planes of the chosen cell (PC:276-278)
template <typename T>
__global__ void __launch_bounds__(256) k_dark_cell_all(const T *__restrict__ stack, long long P, int W, int N, int y0, int x0, int cell,
                                                       unsigned long long *__restrict__ out /* [0]=sum, [1]=count */) {
    unsigned long long s = 0ull, c = 0ull;
    const int per = cell * cell, total = per * N;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int i = e / per, r = e - i * per, dy = r / cell, dx = r - dy * cell;
        const unsigned v = (unsigned)stack[(long long)i * P + (long long)(y0 + dy) * W + x0 + dx];
        s += v; c += v != 0u;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); c += __shfl_down_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, s); atomicAdd(out + 1, c); }
}

}  // namespace

cudaError_t pb_launch_dark_cells(const void *plane0, int dtype, int W, int ny, int nx, int cell, unsigned long long *sums, unsigned *counts,
                                 cudaStream_t st) {
    const int warps = ny * nx, blocks = (warps * 32 + 255) / 256;
    if (dtype == PB_U8) k_dark_cells<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t *)plane0, W, ny, nx, cell, sums, counts);
    else if (dtype == PB_U16) k_dark_cells<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t *)plane0, W, ny, nx, cell, sums, counts);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t pb_launch_dark_cell_all(const void *stack, int dtype, long long P, int W, int N, int y0, int x0, int cell,
                                    unsigned long long *out, cudaStream_t st) {
    const int blocks = (cell * cell * N + 255) / 256 < 64 ? (cell * cell * N + 255) / 256 : 64;
    if (dtype == PB_U8) k_dark_cell_all<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t *)stack, P, W, N, y0, x0, cell, out);
    else if (dtype == PB_U16) k_dark_cell_all<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t *)stack, P, W, N, y0, x0, cell, out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t pb_launch_histo(const void *vals, int is_f64, const uint8_t *sel, long long n, int is_rho, double rot_fig, int fold90,
                            const double *edges_dev, int nbins, double inv_w, int64_t *counts, cudaStream_t st) {
    int blocks = (int)((n + 255) / 256 < 148LL * 8 ? (n + 255) / 256 : 148LL * 8);
    if (blocks < 1) blocks = 1;
    if (is_f64) k_histo<double><<<blocks, 256, 0, st>>>((const double *)vals, sel, n, is_rho, rot_fig, fold90, edges_dev, nbins, inv_w, (unsigned long long *)counts);
    else k_histo<float><<<blocks, 256, 0, st>>>((const float *)vals, sel, n, is_rho, rot_fig, fold90, edges_dev, nbins, inv_w, (unsigned long long *)counts);
    return cudaGetLastError();
}

cudaError_t pb_launch_stats(int pass, const void *vals, const void *rho, int is_f64, const uint8_t *sel, long long n, int circular,
                            double rot_fig, int rewrap, int fold90, double mean, double *acc, cudaStream_t st) {
    int blocks = (int)((n + 255) / 256 < 148LL * 8 ? (n + 255) / 256 : 148LL * 8);
    if (blocks < 1) blocks = 1;
    if (is_f64) {
        if (pass == 1) k_stats1<double><<<blocks, 256, 0, st>>>((const double *)vals, (const double *)rho, sel, n, circular, rot_fig, rewrap, fold90, acc);
        else k_stats2<double><<<blocks, 256, 0, st>>>((const double *)vals, (const double *)rho, sel, n, circular, rot_fig, rewrap, fold90, mean, acc);
    } else {
        if (pass == 1) k_stats1<float><<<blocks, 256, 0, st>>>((const float *)vals, (const float *)rho, sel, n, circular, rot_fig, rewrap, fold90, acc);
        else k_stats2<float><<<blocks, 256, 0, st>>>((const float *)vals, (const float *)rho, sel, n, circular, rot_fig, rewrap, fold90, mean, acc);
    }
    return cudaGetLastError();
}

cudaError_t pb_launch_fit_maps(const double *f, const uint8_t *mask, long long P, int N, int nharm, const Basis2 &bs, double *fit,
                               double *chi2, cudaStream_t st) {
    unsigned blocks = (unsigned)((P + 255) / 256);
    if (nharm == 1) {
        Basis1 b1;
        memcpy(b1.c2, bs.c2, sizeof(b1.c2)); memcpy(b1.s2, bs.s2, sizeof(b1.s2)); memset(b1.kc2, 0, sizeof(b1.kc2)); memset(b1.ks2, 0, sizeof(b1.ks2));
        k_fit_maps<1><<<blocks, 256, 0, st>>>(f, mask, P, N, b1, fit, chi2);
    } else {
        k_fit_maps<2><<<blocks, 256, 0, st>>>(f, mask, P, N, bs, fit, chi2);
    }
    return cudaGetLastError();
}

cudaError_t pb_launch_stats_batch(int pass, const pb_outputs *outs_dev, int n_stacks, const uint32_t *roi_bits, long long P, int n_vars, int f64,
                                  double rot_fig, const uint32_t *group_bits, int n_groups, int per_stack, double *acc1, double *acc2, cudaStream_t st) {
    StatsBatchArgs a;
    a.outs = outs_dev; a.roi_bits = roi_bits; a.P = P; a.n_vars = n_vars; a.f64 = f64; a.rot_fig = rot_fig; a.per_stack = per_stack; a.n_groups = n_groups;
    for (int g = 0; g < PB_MAX_ROIS; ++g) a.group_bits[g] = g < n_groups ? group_bits[g] : 0u;
    int bx = (int)((P + 255) / 256 < 148LL * 4 ? (P + 255) / 256 : 148LL * 4);
    if (bx < 1) bx = 1;
    dim3 grid(bx, n_stacks, n_groups);
    if (f64) {
        if (pass == 1) k_stats_batch<double, 1><<<grid, 256, 0, st>>>(a, acc1, acc2);
        else k_stats_batch<double, 2><<<grid, 256, 0, st>>>(a, acc1, acc2);
    } else {
        if (pass == 1) k_stats_batch<float, 1><<<grid, 256, 0, st>>>(a, acc1, acc2);
        else k_stats_batch<float, 2><<<grid, 256, 0, st>>>(a, acc1, acc2);
    }
    return cudaGetLastError();
}
