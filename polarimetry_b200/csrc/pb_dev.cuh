// pb_dev.cuh -- device-side helpers shared by the kernels of libpolarb200.
//
// Arithmetic contract: every fp64 operation that feeds a mask, a LUT coordinate or a histogram bin is a
// separately rounded IEEE operation in the reference's evaluation order (SURVEY.md Appendix B,
// oracle/contract.c). The translation units are compiled with -fmad=false so that nvcc never contracts
// a*b+c; the fused operations of the contract are spelled fma()/__fma_rn() explicitly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/polar_b200.h"

#define PB_MAX_ANGLES 180

// harmonic basis (host computed, PP:2983-2986 / PP:2993) and k* = -(2^52 * c), the exact constants of the
// one-instruction rounded product RN(f c) = fma(2^52 + f, c, k) used on the integer fast path
struct Basis1 { double c2[PB_MAX_ANGLES], s2[PB_MAX_ANGLES], kc2[PB_MAX_ANGLES], ks2[PB_MAX_ANGLES]; };
struct Basis2 { double c2[PB_MAX_ANGLES], s2[PB_MAX_ANGLES], c4[PB_MAX_ANGLES], s4[PB_MAX_ANGLES],
                       kc2[PB_MAX_ANGLES], ks2[PB_MAX_ANGLES], kc4[PB_MAX_ANGLES], ks4[PB_MAX_ANGLES]; };
#ifndef PB_PIXEL_MIN_BLOCKS
#define PB_PIXEL_MIN_BLOCKS 8
#endif
template <int NH> struct BasisSel { using type = Basis1; };
template <> struct BasisSel<2> { using type = Basis2; };

// Per-launch constants of the per-pixel kernels (passed by value: lives in the constant bank).
struct PixArgs {
    const void *stack;          // [n_stacks][N][P] raw stack, or double field
    long long stack_stride;     // elements between consecutive stacks
    const double *intensity_in; // [P] precomputed threshold image (unfused path) or NULL (computed inline)
    const uint8_t *mask_in;     // [P] incoming mask (compute_fields seam) or NULL
    const uint32_t *roi_bits;   // [P] or NULL
    const double *edge;         // [P] or NULL
    const pb_outputs *outs;     // device array [n_stacks]
    const double2 *lut;         // [n*n] (rho, psi) interleaved
    const double *grid;         // [n]
    const double2 *lut_hy;      // [n-1] (h = g[k+1]-g[k], RN(1/h)) per cell
    const double *hist_edges;   // [n_vars][nbins+1]
    int lut_n;
    int epi_smem_off;           // per-pixel kernel: byte offset of the stage-B queue / result area in dynamic shared memory (set by the launcher)
    int N, H, W;
    long long P;
    int method;
    int do_sqrt, from_field;    // from_field: input is the already pre-processed double field
    int n_vars, nbins, n_groups, n_rois;
    unsigned flags;
    unsigned hist_is_rho;       // bit v set: variable v is re-wrapped before binning
    double sub, dark;           // sub = dark + removed_intensity
    double rot_stick, rot_fig, ref_angle, chi2thr;
    double invN_hi, invN_lo, k2, dN;   // double-double 1/N, 2.0/N, (double)N (host computed)
    float thr_lo_n, thr_hi_n;          // chi2 threshold * N with -/+ 1e-5 margins (screening)
    double grid_lo, grid_hi, half_nm1; // LUT grid end points and (n-1)/2
    int simple;                        // no ROI bits, no incoming mask, one histogram group: whole-image thresholding
    int out_aligned;                   // every output map of every stack is 16-byte aligned (mask: 4-byte): vector stores allowed
    int all_lean;                      // every stack asks for the throughput-mode output set (store_is_lean) -- host evaluated
    double inv_bin_width[PB_MAX_VARS];
    double hist_lo[PB_MAX_VARS], hist_hi[PB_MAX_VARS];   // first / last histogram edge of every variable
    double roi_ilow[PB_MAX_ROIS];
    uint32_t group_bits[PB_MAX_ROIS];
    double invk[16];
};

__device__ __forceinline__ bool pb_finite(double x) { return fabs(x) <= 1.7976931348623157e308; }
__device__ __forceinline__ double pb_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// numpy.maximum(x, 0): NaN propagates
__device__ __forceinline__ double pb_max0(double x) { return (x > 0.0) ? x : ((x != x) ? x : 0.0); }

// np.mod(a, b) for b > 0: python-sign modulo; a tiny negative a gives exactly b
static __device__ __noinline__ double pb_pymod_general(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) { if (r < 0.0) r = r + b; }
    else r = 0.0;
    return r;
}
__device__ __forceinline__ double pb_pymod(double a, double b) {
    if (a > 0.0 && a < b) return a;               // fmod(a,b) == a: the common case, kept inline
    return pb_pymod_general(a, b);
}

// vector loads of PX consecutive elements (callers guarantee alignment when PX > 1)
template <typename T, int PX> struct VecLoad;
template <typename T> struct VecLoad<T, 1> {
    __device__ static __forceinline__ void ld(const T *p, double *o) { o[0] = (double)__ldg(p); }
};
template <> struct VecLoad<uint16_t, 4> {
    __device__ static __forceinline__ void ld(const uint16_t *p, double *o) {
        uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
        o[0] = (double)(r.x & 0xffffu); o[1] = (double)(r.x >> 16); o[2] = (double)(r.y & 0xffffu); o[3] = (double)(r.y >> 16);
    }
};
template <> struct VecLoad<uint8_t, 4> {
    __device__ static __forceinline__ void ld(const uint8_t *p, double *o) {
        unsigned r = __ldg(reinterpret_cast<const unsigned *>(p));
        o[0] = (double)(r & 0xffu); o[1] = (double)((r >> 8) & 0xffu); o[2] = (double)((r >> 16) & 0xffu); o[3] = (double)(r >> 24);
    }
};
template <> struct VecLoad<float, 4> {
    __device__ static __forceinline__ void ld(const float *p, double *o) {
        float4 r = __ldg(reinterpret_cast<const float4 *>(p));
        o[0] = (double)r.x; o[1] = (double)r.y; o[2] = (double)r.z; o[3] = (double)r.w;
    }
};
template <> struct VecLoad<double, 4> {
    __device__ static __forceinline__ void ld(const double *p, double *o) {
        double2 a = __ldg(reinterpret_cast<const double2 *>(p));
        double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
};

// scipy.ndimage 'reflect' (half-sample symmetric) index
__device__ __forceinline__ int pb_reflect(int k, int L) {
    while (k < 0 || k >= L) {
        if (k < 0) k = -k - 1;
        if (k >= L) k = 2 * L - 1 - k;
    }
    return k;
}

// histogram bin of d for edges e[0..nbins] (np.histogram rule: [e_k, e_{k+1}), last bin closed); -1 = dropped
__device__ __forceinline__ int pb_bin(double d, const double *e, int nbins, double inv_w) {
    if (!(d >= e[0]) || !(d <= e[nbins])) return -1;
    int k = (int)((d - e[0]) * inv_w);
    k = max(0, min(k, nbins - 1));
    while (k > 0 && d < e[k]) --k;
    while (k < nbins - 1 && d >= e[k + 1]) ++k;
    return k;
}
