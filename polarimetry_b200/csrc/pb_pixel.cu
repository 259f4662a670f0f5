// pb_pixel.cu -- per-pixel analysis kernels (no spatial coupling): harmonic projection, fit/chi2
// decisions, disk-cone LUT inversion, 4-harmonic outputs, 4POLAR inversion, threshold/ROI mask and
// per-CTA shared-memory histograms.
//
// Used (a) as the FUSED single-pass path when no binning is requested (raw uint8/uint16/float stack in,
// stack read once, maps + histograms out) and (b) as the last stage of the unfused path (double field in).
// Replaces compute_fields, PyPOLAR.py:2981-3078, plus the mask logic of PP:2891/PP:2967 and the
// histogram of pypolar_classes.py:340-369. Arithmetic order follows oracle/contract.c.
#include "pb_dev.cuh"

namespace {

struct HistSmem {
    // [n_groups][n_vars][nbins] counters followed by the edges; carved from dynamic shared memory
    unsigned *cnt;
    double *edges;
};

__device__ __forceinline__ HistSmem hist_setup(const PixArgs &a, unsigned char *smem) {
    HistSmem h;
    int ne = a.n_vars * (a.nbins + 1);
    h.edges = reinterpret_cast<double *>(smem);
    h.cnt = reinterpret_cast<unsigned *>(smem + sizeof(double) * ne);
    int nc = a.n_groups * a.n_vars * a.nbins;
    for (int k = threadIdx.x; k < ne; k += blockDim.x) h.edges[k] = a.hist_edges[k];
    for (int k = threadIdx.x; k < nc; k += blockDim.x) h.cnt[k] = 0u;
    __syncthreads();
    return h;
}

__device__ __forceinline__ void hist_flush(const PixArgs &a, const HistSmem &h, int64_t *dst) {
    __syncthreads();
    int nc = a.n_groups * a.n_vars * a.nbins;
    for (int k = threadIdx.x; k < nc; k += blockDim.x) {
        unsigned c = h.cnt[k];
        if (c) atomicAdd(reinterpret_cast<unsigned long long *>(dst) + k, (unsigned long long)c);
    }
}

__device__ __forceinline__ void hist_add(const PixArgs &a, const HistSmem &h, int var, double val, uint32_t bits) {
    if (!pb_finite(val)) return;
    double d = val;
    if ((a.hist_is_rho >> var) & 1u) d = pb_pymod(2.0 * (d + a.rot_fig), 360.0) / 2.0;
    int b = pb_bin(d, h.edges + var * (a.nbins + 1), a.nbins, a.inv_bin_width[var]);
    if (b < 0) return;
    for (int g = 0; g < a.n_groups; ++g)
        if (bits & a.group_bits[g]) atomicAdd(&h.cnt[(g * a.n_vars + var) * a.nbins + b], 1u);
}

template <typename OUT> __device__ __forceinline__ void store_map(void *base, long long p, double v) {
    if (base) reinterpret_cast<OUT *>(base)[p] = (OUT)v;
}
__device__ __forceinline__ void store_any(void *base, long long p, double v, bool f64) {
    if (f64) store_map<double>(base, p, v); else store_map<float>(base, p, v);
}

// threshold / ROI test, PP:2891 + PP:2967: any_r( bit_r & I >= ILow_r )
__device__ __forceinline__ bool roi_pass(const PixArgs &a, uint32_t bits, double I) {
    if (a.n_rois == 0) return (bits != 0u) && (I >= a.roi_ilow[0]);
    bool ok = false;
    for (int r = 0; r < a.n_rois; ++r) ok |= ((bits >> r) & 1u) && (I >= a.roi_ilow[r]);
    return ok;
}

// disk-cone LUT cell: g[i] <= x < g[i+1], last cell closed (scipy RegularGridInterpolator rule)
__device__ __forceinline__ int lut_cell(const double *g, int n, double x) {
    int k = (int)floor((x + 1.0) * (0.5 * (double)(n - 1)));
    k = max(0, min(k, n - 2));
    while (k > 0 && x < __ldg(g + k)) --k;
    while (k < n - 2 && x >= __ldg(g + k + 1)) ++k;
    return k;
}

__device__ __forceinline__ void lut_lookup(const PixArgs &a, double x, double y, double &rho, double &psi) {
    rho = pb_nan(); psi = pb_nan();
    const double *g = a.grid;
    const int n = a.lut_n;
    double g0 = __ldg(g), g1 = __ldg(g + n - 1);
    if (!(x >= g0 && x <= g1 && y >= g0 && y <= g1)) return;   // also rejects NaN
    int i0 = lut_cell(g, n, x), i1 = lut_cell(g, n, y);
    double ga = __ldg(g + i0), gb = __ldg(g + i0 + 1), gc = __ldg(g + i1), gd = __ldg(g + i1 + 1);
    double t0 = (x - ga) / (gb - ga), t1 = (y - gc) / (gd - gc);
    double u0 = 1.0 - t0, u1 = 1.0 - t1;
    double w00 = (1.0 * u0) * u1, w01 = (1.0 * u0) * t1, w10 = (1.0 * t0) * u1, w11 = (1.0 * t0) * t1;
    const double2 *T = a.lut;
    double2 v00 = __ldg(T + (long long)i0 * n + i1), v01 = __ldg(T + (long long)i0 * n + i1 + 1);
    double2 v10 = __ldg(T + (long long)(i0 + 1) * n + i1), v11 = __ldg(T + (long long)(i0 + 1) * n + i1 + 1);
    double r = 0.0, q = 0.0;
    r = r + v00.x * w00; r = r + v01.x * w01; r = r + v10.x * w10; r = r + v11.x * w11;
    q = q + v00.y * w00; q = q + v01.y * w01; q = q + v10.y * w10; q = q + v11.y * w11;
    rho = r; psi = q;
}

constexpr double RAD2DEG = 180.0 / 3.141592653589793238462643383279502884;
constexpr double DEG2RAD = 3.141592653589793238462643383279502884 / 180.0;

// ---------------------------------------------------------------------------------------------------
// Harmonic modalities (1PF: NH=1, CARS/SRS/SHG/2PF: NH=2). One thread = PX consecutive pixels.
template <typename T, int NH, int PX>
__global__ void __launch_bounds__(256) k_pixel_harm(PixArgs a, typename BasisSel<NH>::type bs) {
    extern __shared__ __align__(16) unsigned char smem[];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0;
    HistSmem hs;
    if (do_hist) hs = hist_setup(a, smem);
    const int stack = blockIdx.y;
    const pb_outputs out = a.outs[stack];
    const long long P = a.P;
    const long long p0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * PX;
    const bool out64 = a.flags & PB_FLAG_OUT_F64;
    if (p0 < P) {
        const T *base = reinterpret_cast<const T *>(a.stack) + (long long)stack * a.stack_stride + p0;
        const int N = a.N;
        const bool pre = !a.from_field;
        const bool sep_int = pre && !a.intensity_in && (a.do_sqrt || a.sub != a.dark);
        double S0[PX], Sc[PX], Ss[PX], Sc4[PX], Ss4[PX], Sff[PX], Isum[PX];
#pragma unroll
        for (int k = 0; k < PX; ++k) { S0[k] = Sc[k] = Ss[k] = Sc4[k] = Ss4[k] = Sff[k] = Isum[k] = 0.0; }
#pragma unroll 6
        for (int i = 0; i < N; ++i) {
            double v[PX];
            VecLoad<T, PX>::ld(base + (long long)i * P, v);
            const double c2 = bs.c2[i], s2 = bs.s2[i];
#pragma unroll
            for (int k = 0; k < PX; ++k) {
                double f = v[k];
                if (pre) {
                    if (sep_int) Isum[k] = Isum[k] + pb_max0(v[k] - a.dark);
                    f = pb_max0(f - a.sub);
                    if (a.do_sqrt) f = sqrt(f);
                }
                S0[k] = S0[k] + f;
                Sc[k] = Sc[k] + f * c2;
                Ss[k] = Ss[k] + f * s2;
                if constexpr (NH == 2) {
                    Sc4[k] = Sc4[k] + f * bs.c4[i];
                    Ss4[k] = Ss4[k] + f * bs.s4[i];
                }
                Sff[k] = fma(f, f, Sff[k]);   // screening only (not parity relevant)
            }
        }
        const double dN = (double)N, k2 = 2.0 / dN;
        bool need_exact[PX], harm_ok[PX];
        double a0[PX], a2r[PX], a2i[PX], a4r[PX], a4i[PX];
        bool any_exact = false;
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            a0[k] = S0[k] / dN;
            if (a0[k] == 0.0) a0[k] = pb_nan();
            a2r[k] = k2 * Sc[k]; a2i[k] = k2 * Ss[k];
            a4r[k] = k2 * Sc4[k]; a4i[k] = k2 * Ss4[k];
            // Certified screening of (all fit_i > 0) & (chi2 <= thr) & (chi2 > 0) from the orthogonality identity
            //   sum_i (f_i - fit_i)^2 = sum f^2 - N (a0^2 + |a2|^2/2 + |a4|^2/2),  a0 - |a2| - |a4| <= fit_i <= a0 + |a2| + |a4|.
            // Undecided pixels (and N too small for the identity) replay the reference's fp64 loop below.
            need_exact[k] = true; harm_ok[k] = false;
            if (!(a.flags & PB_FLAG_EXACT_CHI2) && N >= (NH == 2 ? 5 : 3) && pb_finite(a0[k])) {
                double q2 = a2r[k] * a2r[k] + a2i[k] * a2i[k], q4 = a4r[k] * a4r[k] + a4i[k] * a4i[k];
                double amp = sqrt(q2) + (NH == 2 ? sqrt(q4) : 0.0);
                double model = dN * (a0[k] * a0[k] + 0.5 * (q2 + (NH == 2 ? q4 : 0.0)));
                double sse = Sff[k] - model;
                double err = 1e-11 * (Sff[k] + model);
                double fmin = a0[k] - amp * (1.0 + 1e-9), fmax = a0[k] + amp * (1.0 + 1e-9);
                if (fmin > 1e-9 * a0[k] && pb_finite(fmax)) {
                    double hi = (sse + err) / (dN * fmin) * (1.0 + 1e-9);
                    double lo = (sse - err) / (dN * fmax) * (1.0 - 1e-9);
                    if (hi <= a.chi2thr && lo > 0.0) { need_exact[k] = false; harm_ok[k] = true; }
                    else if (lo > a.chi2thr) { need_exact[k] = false; harm_ok[k] = false; }
                }
            }
            any_exact |= need_exact[k];
        }
        if (any_exact) {
            bool valid[PX]; double acc[PX];
#pragma unroll
            for (int k = 0; k < PX; ++k) { valid[k] = true; acc[k] = 0.0; }
            for (int i = 0; i < N; ++i) {
                double v[PX];
                VecLoad<T, PX>::ld(base + (long long)i * P, v);
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    double f = v[k];
                    if (pre) { f = pb_max0(f - a.sub); if (a.do_sqrt) f = sqrt(f); }
                    // numpy's fused SIMD complex multiply: Re(a2 conj(e2)) = fma(a2r, c, -(a2i * (-s)))
                    double ft = a0[k] + fma(a2r[k], bs.c2[i], a2i[k] * bs.s2[i]);
                    if constexpr (NH == 2) ft = ft + fma(a4r[k], bs.c4[i], a4i[k] * bs.s4[i]);
                    valid[k] = valid[k] && (ft > 0.0) && pb_finite(ft);
                    double d = f - ft;
                    acc[k] = acc[k] + (d * d) / ft;
                }
            }
#pragma unroll
            for (int k = 0; k < PX; ++k)
                if (need_exact[k]) {
                    double chi2 = acc[k] / dN;
                    harm_ok[k] = valid[k] && (chi2 <= a.chi2thr) && (chi2 > 0.0);
                }
        }
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            const long long p = p0 + k;
            if (p >= P) break;
            const uint32_t bits = a.roi_bits ? __ldg(a.roi_bits + p) : 1u;
            double I = a.intensity_in ? __ldg(a.intensity_in + p) : (sep_int ? Isum[k] : S0[k]);
            bool m = a.mask_in ? (__ldg(a.mask_in + p) != 0) : roi_pass(a, bits, I);
            m = m && harm_ok[k];
            const double inv = 1.0 / a0[k];
            double A2 = a2r[k] * inv, B2 = a2i[k] * inv;
            if (!(pb_finite(A2) && pb_finite(B2))) { A2 = pb_nan(); B2 = pb_nan(); }
            double v0 = pb_nan(), v1 = pb_nan(), v2 = pb_nan(), v3 = pb_nan();
            if (NH == 1) {
                if (m) {
                    lut_lookup(a, A2, B2, v0, v1);
                    if (pb_finite(v0)) v0 = pb_pymod(v0 + a.rot_stick, 180.0);
                    m = pb_finite(v0) && pb_finite(v1);
                    // the reference keeps a finite rho next to a NaN psi (PP:3021-3024); both tables share their NaN rim
                }
            } else {
                double A4 = a4r[k] * inv, B4 = a4i[k] * inv;
                if (!(pb_finite(A4) && pb_finite(B4))) { A4 = pb_nan(); B4 = pb_nan(); }
                if (m) {
                    double r = (atan2(B2, A2) * RAD2DEG) / 2.0;
                    r = pb_pymod(r + a.rot_stick, 180.0);
                    double ab2 = hypot(A2, B2), ab4 = hypot(A4, B4);
                    v0 = r;
                    v1 = 1.5 * ab2;
                    v2 = (6.0 * ab4) * cos(4.0 * (0.25 * atan2(B4, A4) - r * DEG2RAD));
                    if (a.method == PB_SHG) v3 = (-0.5 * (ab4 - ab2)) / (ab4 + ab2) - 0.65;
                }
            }
            store_any(out.var[0], p, v0, out64);
            store_any(out.var[1], p, v1, out64);
            if constexpr (NH == 2) {
                store_any(out.var[2], p, v2, out64);
                if (a.method == PB_SHG) store_any(out.var[3], p, v3, out64);
            }
            store_any(out.intmap, p, m ? a0[k] : pb_nan(), out64);
            if (out.intensity && !a.intensity_in) reinterpret_cast<double *>(out.intensity)[p] = I;
            if (out.mask) out.mask[p] = (uint8_t)m;
            if (out.rho_angle) store_any(out.rho_angle, p, (v0 == v0) ? pb_pymod(v0 - a.ref_angle, 180.0) : pb_nan(), out64);
            if (out.rho_contour && a.edge) {
                double e = __ldg(a.edge + p);
                store_any(out.rho_contour, p, (pb_finite(v0) && pb_finite(e)) ? pb_pymod(v0 - e, 180.0) : pb_nan(), out64);
            }
            if (do_hist && out.hist) {
                hist_add(a, hs, 0, v0, bits);
                hist_add(a, hs, 1, v1, bits);
                if constexpr (NH == 2) {
                    hist_add(a, hs, 2, v2, bits);
                    if (a.method == PB_SHG) hist_add(a, hs, 3, v3, bits);
                }
            }
        }
    }
    if (do_hist && out.hist) hist_flush(a, hs, out.hist);
}

// ---------------------------------------------------------------------------------------------------
// 4POLAR 2D / 3D (PP:3001-3014, PP:3041-3051). One thread = one pixel; 4 planes.
template <typename T>
__global__ void __launch_bounds__(256) k_pixel_4polar(PixArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0;
    HistSmem hs;
    if (do_hist) hs = hist_setup(a, smem);
    const int stack = blockIdx.y;
    const pb_outputs out = a.outs[stack];
    const long long P = a.P;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool out64 = a.flags & PB_FLAG_OUT_F64;
    const bool three_d = a.method == PB_4POLAR_3D;
    if (p < P) {
        const T *base = reinterpret_cast<const T *>(a.stack) + (long long)stack * a.stack_stride + p;
        double f[4], Isum = 0.0;
        const bool pre = !a.from_field;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double v = (double)__ldg(base + (long long)i * P);
            if (pre) {
                Isum = Isum + pb_max0(v - a.dark);
                v = pb_max0(v - a.sub);
            }
            // planes 1 <-> 2 swapped (PP:2957) when reading the raw stack
            f[pre ? ((i == 1 || i == 2) ? 3 - i : i) : i] = v;
        }
        double m4[4] = {0.0, 0.0, 0.0, 0.0};
        const int rows = three_d ? 4 : 3;
        for (int r = 0; r < rows; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = acc + a.invk[r * 4 + j] * f[j];
            m4[r] = acc;
        }
        double s = three_d ? ((m4[0] + m4[1]) + m4[2]) : ((m4[0] + m4[1]) + 0.0);
        if (s == 0.0) s = pb_nan();
        double pxy = (m4[0] - m4[1]) / s;
        double puv = (2.0 * m4[three_d ? 3 : 2]) / s;
        double pzz = 0.0, lam;
        if (three_d) {
            pzz = m4[2] / s;
            lam = (1.0 - (pzz + sqrt(puv * puv + pxy * pxy))) / 2.0;
        } else {
            double p2 = sqrt(puv * puv + pxy * pxy);
            lam = (1.0 - p2) / (3.0 - p2);
        }
        double a0 = ((((f[0] + f[1]) + f[2]) + f[3]) / 4.0) / 4.0;
        if (a0 == 0.0) a0 = pb_nan();
        const uint32_t bits = a.roi_bits ? __ldg(a.roi_bits + p) : 1u;
        // raw-stack intensity sums planes in acquisition order; addition order matters only for non-integers
        double I = a.intensity_in ? __ldg(a.intensity_in + p) : Isum;
        bool m = a.mask_in ? (__ldg(a.mask_in + p) != 0) : roi_pass(a, bits, I);
        m = m && (lam < 1.0 / 3.0) && (lam > 0.0) && (three_d ? (pzz > lam) : true);
        double v0 = pb_nan(), v1 = pb_nan(), v2 = pb_nan();
        if (m) {
            v0 = pb_pymod(0.5 * (atan2(puv, pxy) * RAD2DEG) + a.rot_stick, 180.0);
            v1 = 2.0 * (acos((-1.0 + sqrt(9.0 - 24.0 * lam)) / 2.0) * RAD2DEG);
            if (three_d) v2 = acos(sqrt((pzz - lam) / (1.0 - 3.0 * lam))) * RAD2DEG;
        }
        store_any(out.var[0], p, v0, out64);
        store_any(out.var[1], p, v1, out64);
        if (three_d) store_any(out.var[2], p, v2, out64);
        store_any(out.intmap, p, m ? a0 : pb_nan(), out64);
        if (out.intensity && !a.intensity_in) reinterpret_cast<double *>(out.intensity)[p] = I;
        if (out.mask) out.mask[p] = (uint8_t)m;
        if (out.rho_angle) store_any(out.rho_angle, p, (v0 == v0) ? pb_pymod(v0 - a.ref_angle, 180.0) : pb_nan(), out64);
        if (out.rho_contour && a.edge) {
            double e = __ldg(a.edge + p);
            store_any(out.rho_contour, p, (pb_finite(v0) && pb_finite(e)) ? pb_pymod(v0 - e, 180.0) : pb_nan(), out64);
        }
        if (do_hist && out.hist) {
            hist_add(a, hs, 0, v0, bits);
            hist_add(a, hs, 1, v1, bits);
            if (three_d) hist_add(a, hs, 2, v2, bits);
        }
    }
    if (do_hist && out.hist) hist_flush(a, hs, out.hist);
}

template <typename T, int NH>
cudaError_t launch_harm_t(const PixArgs &a, const typename BasisSel<NH>::type &bs, int n_stacks, size_t smem, cudaStream_t st) {
    const int threads = 256;
    // PX = 4 needs every plane (and every stack) to start on a 4-element boundary
    bool vec = (a.P % 4 == 0) && (a.stack_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.stack) % (4 * sizeof(T))) == 0);
    if (vec) {
        long long nthr = a.P / 4;
        dim3 grid((unsigned)((nthr + threads - 1) / threads), n_stacks);
        k_pixel_harm<T, NH, 4><<<grid, threads, smem, st>>>(a, bs);
    } else {
        dim3 grid((unsigned)((a.P + threads - 1) / threads), n_stacks);
        k_pixel_harm<T, NH, 1><<<grid, threads, smem, st>>>(a, bs);
    }
    return cudaGetLastError();
}

}  // namespace

size_t pb_hist_smem_bytes(int n_vars, int nbins, int n_groups) {
    return sizeof(double) * n_vars * (nbins + 1) + sizeof(unsigned) * n_groups * n_vars * nbins + 16;
}

cudaError_t pb_launch_pixel_harm(const PixArgs &a, const Basis2 &bs, int dtype, int n_stacks, cudaStream_t st) {
    size_t smem = pb_hist_smem_bytes(a.n_vars, a.nbins, a.n_groups);
    const bool one = a.method == PB_1PF;
    Basis1 b1;
    if (one) { memcpy(b1.c2, bs.c2, sizeof(b1.c2)); memcpy(b1.s2, bs.s2, sizeof(b1.s2)); }
#define PB_DISPATCH(T)                                                                     \
    return one ? launch_harm_t<T, 1>(a, b1, n_stacks, smem, st) : launch_harm_t<T, 2>(a, bs, n_stacks, smem, st)
    switch (dtype) {
    case PB_U8: PB_DISPATCH(uint8_t);
    case PB_U16: PB_DISPATCH(uint16_t);
    case PB_F32: PB_DISPATCH(float);
    case PB_F64: PB_DISPATCH(double);
    }
#undef PB_DISPATCH
    return cudaErrorInvalidValue;
}

cudaError_t pb_launch_pixel_4polar(const PixArgs &a, int dtype, int n_stacks, cudaStream_t st) {
    size_t smem = pb_hist_smem_bytes(a.n_vars, a.nbins, a.n_groups);
    const int threads = 256;
    dim3 grid((unsigned)((a.P + threads - 1) / threads), n_stacks);
    switch (dtype) {
    case PB_U8: k_pixel_4polar<uint8_t><<<grid, threads, smem, st>>>(a); break;
    case PB_U16: k_pixel_4polar<uint16_t><<<grid, threads, smem, st>>>(a); break;
    case PB_F32: k_pixel_4polar<float><<<grid, threads, smem, st>>>(a); break;
    case PB_F64: k_pixel_4polar<double><<<grid, threads, smem, st>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
