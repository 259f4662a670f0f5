// pb_pixel.cuh -- per-pixel device code shared by the fused per-pixel kernel (pb_pixel.cu) and the fused
// row-march kernel (pb_march.cu): shared-memory histograms, threshold/ROI test, disk-cone LUT inversion,
// certified fit/chi2 screening and the per-pixel epilogue of the harmonic modalities.
// Arithmetic order of everything that reaches a mask, a LUT coordinate or a bin follows oracle/contract.c.
#pragma once
#include "pb_dev.cuh"

struct HistSmem {
    unsigned *cnt;     // [n_groups][n_vars][nbins]
    double *edges;     // [n_vars][nbins+1]
};

__device__ __forceinline__ HistSmem hist_setup(const PixArgs &a, unsigned char *smem) {
    HistSmem h;
    const int ne = a.n_vars * (a.nbins + 1);
    h.edges = reinterpret_cast<double *>(smem);
    h.cnt = reinterpret_cast<unsigned *>(smem + sizeof(double) * ne);
    const int nc = a.n_groups * a.n_vars * a.nbins;
    for (int k = threadIdx.x; k < ne; k += blockDim.x) h.edges[k] = a.hist_edges[k];
    for (int k = threadIdx.x; k < nc; k += blockDim.x) h.cnt[k] = 0u;
    __syncthreads();
    return h;
}

__device__ __forceinline__ void hist_flush(const PixArgs &a, const HistSmem &h, int64_t *dst) {
    __syncthreads();
    const int nc = a.n_groups * a.n_vars * a.nbins;
    for (int k = threadIdx.x; k < nc; k += blockDim.x) {
        const unsigned c = h.cnt[k];
        if (c) atomicAdd(reinterpret_cast<unsigned long long *>(dst) + k, (unsigned long long)c);
    }
}

// np.histogram rule on explicit edges: [e_k, e_{k+1}), last bin closed, outside (and NaN / inf) dropped.
// The first and last edge come from the constant bank; the arithmetic bin guess is verified against the two edges around it
// (they are uniform -- np.linspace, checked by pb_set_hist -- so the guess is off by at most one bin) without a branch.
// SIMPLE = 1: the caller has established a.simple (whole-image thresholding, one histogram group)
template <int SIMPLE = 0>
__device__ __forceinline__ void hist_add(const PixArgs &a, const HistSmem &h, int var, double val, uint32_t bits) {
    double d = val;
    if ((a.hist_is_rho >> var) & 1u) {
        // Variable.histo re-wrap d = mod(2(d + rot), 360)/2 (PC:341-342); for rot == 0 and 0 <= d <= 180 it is d, except 180 -> 0
        if (a.rot_fig == 0.0 && d >= 0.0 && d <= 180.0) d = (d == 180.0) ? 0.0 : d;
        else d = pb_pymod(2.0 * (d + a.rot_fig), 360.0) / 2.0;
    }
    const int nb = a.nbins;
    const double e0 = a.hist_lo[var];
    const bool inr = (d >= e0) && (d <= a.hist_hi[var]);          // false for NaN
    int k = (int)((d - e0) * a.inv_bin_width[var]);               // NaN -> 0; saturates
    k = max(0, min(k, nb - 1));
    const double *e = h.edges + var * (nb + 1) + k;
    const double ea = e[0], eb = e[1];
    const bool dn = d < ea;
    const bool up = !dn && (k < nb - 1) && (d >= eb);
    k = k + (up ? 1 : 0) - (dn ? 1 : 0);
    if (!inr) return;
    unsigned *c = h.cnt + var * nb + k;
    if (SIMPLE || a.simple) {
        if (a.flags & PB_FLAG_HIST_MATCH) {
            // lanes of this warp that update the same counter elect one of them, which adds their number
            const unsigned peers = __match_any_sync(__activemask(), var * nb + k);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(c, (unsigned)__popc(peers));
        } else {
            atomicAdd(c, 1u);
        }
    } else if (a.n_groups == 1) {
        if (bits & a.group_bits[0]) atomicAdd(c, 1u);
    } else {
#pragma unroll 1
        for (int g = 0; g < a.n_groups; ++g)
            if (bits & a.group_bits[g]) atomicAdd(c + g * a.n_vars * nb, 1u);
    }
}

// threshold / ROI test, PP:2891 + PP:2967: any_r( bit_r & I >= ILow_r )
__device__ __forceinline__ bool roi_pass(const PixArgs &a, uint32_t bits, double I) {
    if (a.n_rois == 0) return (bits != 0u) && (I >= a.roi_ilow[0]);
    bool ok = false;
#pragma unroll 1
    for (int r = 0; r < a.n_rois; ++r) ok |= ((bits >> r) & 1u) && (I >= a.roi_ilow[r]);   // rolled: keeps the kernels inside the instruction cache
    return ok;
}

// disk-cone LUT cell and normalised distance along one axis: g[i] <= x < g[i+1], last cell closed
// (scipy RegularGridInterpolator rule); t = (x - g[i]) / (g[i+1] - g[i]). Caller guarantees g[0] <= x <= g[n-1].
// The arithmetic guess floor((x+1)(n-1)/2) is exact up to one cell for a linspace grid; it is then corrected against
// the real grid values.
__device__ __forceinline__ int lut_axis(const double *__restrict__ g, const double2 *__restrict__ hy, int n, double half_nm1, double x, double &t) {
    int k = __double2int_rd(fma(x, half_nm1, half_nm1));
    k = max(0, min(k, n - 2));
    double ga = __ldg(g + k), gb = __ldg(g + k + 1);
    // the grid is uniform (np.linspace, PC:463; checked by pb_set_diskcone): the guess is off by at most one cell
    if (x < ga) { --k; ga = __ldg(g + k); }
    else if (x >= gb && k < n - 2) { ++k; ga = gb; }
    // t = (x - ga) / (gb - ga), correctly rounded from the tabulated spacing h = gb - ga and y = RN(1/h)
    // (Markstein: q0 = RN(d y), r = d - q0 h exactly, t = RN(q0 + r y); tests/test_shortcuts.py)
    const double2 c = __ldg(hy + k);
    const double d = x - ga, q0 = d * c.y;
    t = fma(fma(-q0, c.x, d), c.y, q0);
    return k;
}

// bilinear interpolation of the interleaved (rho, psi) table in the reference's evaluation order
// (SURVEY 8a row a6): v = 0 + V00 w00 + V01 w01 + V10 w10 + V11 w11, w = ((1)(1-t0))(1-t1) ... ; NaN corners propagate.
// The cell is first GUESSED arithmetically (exact up to one cell for the linspace grid) and everything that depends on it -- grid
// values, the (h, 1/h) entries and the four table corners -- is loaded at once, so the lookup costs one memory latency instead of
// three dependent ones; the guess is then verified against the real grid values and the rare miss takes the searched path.
__device__ __forceinline__ void lut_lookup(const PixArgs &a, double x, double y, double &rho, double &psi) {
    rho = pb_nan(); psi = pb_nan();
    if (!(x >= a.grid_lo && x <= a.grid_hi && y >= a.grid_lo && y <= a.grid_hi)) return;   // out of range or NaN -> fill value NaN
    const int n = a.lut_n;
    int i0 = max(0, min(__double2int_rd(fma(x, a.half_nm1, a.half_nm1)), n - 2));
    int i1 = max(0, min(__double2int_rd(fma(y, a.half_nm1, a.half_nm1)), n - 2));
    const double ga0 = __ldg(a.grid + i0), gb0 = __ldg(a.grid + i0 + 1), ga1 = __ldg(a.grid + i1), gb1 = __ldg(a.grid + i1 + 1);
    const double2 c0 = __ldg(a.lut_hy + i0), c1 = __ldg(a.lut_hy + i1);
    const double2 *T = a.lut + (i0 * n + i1);
    double2 v00 = __ldg(T), v01 = __ldg(T + 1), v10 = __ldg(T + n), v11 = __ldg(T + n + 1);
    double t0, t1;
    if (x >= ga0 && (x < gb0 || i0 == n - 2) && y >= ga1 && (y < gb1 || i1 == n - 2)) {
        // t = (x - ga) / (gb - ga), correctly rounded from the tabulated spacing h = gb - ga and y = RN(1/h)
        // (Markstein: q0 = RN(d y), r = d - q0 h exactly, t = RN(q0 + r y); tests/test_shortcuts.py)
        const double d0 = x - ga0, q0 = d0 * c0.y, d1 = y - ga1, q1 = d1 * c1.y;
        t0 = fma(fma(-q0, c0.x, d0), c0.y, q0);
        t1 = fma(fma(-q1, c1.x, d1), c1.y, q1);
    } else {
        i0 = lut_axis(a.grid, a.lut_hy, n, a.half_nm1, x, t0);
        i1 = lut_axis(a.grid, a.lut_hy, n, a.half_nm1, y, t1);
        T = a.lut + (i0 * n + i1);
        v00 = __ldg(T); v01 = __ldg(T + 1); v10 = __ldg(T + n); v11 = __ldg(T + n + 1);
    }
    const double u0 = 1.0 - t0, u1 = 1.0 - t1;
    const double w00 = u0 * u1, w01 = u0 * t1, w10 = t0 * u1, w11 = t0 * t1;    // (1.0 * u0) == u0 exactly
    // 0.0 + p == p for every p that can matter here (p = -0.0 only for a zero table entry times a zero weight)
    double r = v00.x * w00, q = v00.y * w00;
    r = r + v01.x * w01; q = q + v01.y * w01;
    r = r + v10.x * w10; q = q + v10.y * w10;
    r = r + v11.x * w11; q = q + v11.y * w11;
    rho = r; psi = q;
}

constexpr double PB_RAD2DEG = 180.0 / 3.141592653589793238462643383279502884;
constexpr double PB_DEG2RAD = 3.141592653589793238462643383279502884 / 180.0;

struct PixOut {
    double v0, v1, v2, v3, a0m;
    bool m;
};

// state of one pixel between the two stages of the epilogue
struct HarmPix {
    double a0, a2r, a2i, a4r, a4i;
    uint32_t bits;
    bool m;       // still a candidate (threshold / ROI passed, a0 != 0, screening did not reject)
    bool need;    // the fit/chi2 decision needs the reference's fp64 loop (ExactAcc)
};

// accumulator of the reference's fit/chi2 loop for one pixel (PP:2990-3000)
struct ExactAcc {
    double acc;
    bool valid;
    __device__ __forceinline__ void init() { acc = 0.0; valid = true; }
    // ft = a0 + Re(a2 conj(e2_i)) [+ Re(a4 conj(e4_i))], numpy's fused SIMD complex multiply: fma(a2r, c, -(a2i * (-s)))
    template <int NH>
    __device__ __forceinline__ void step(const HarmPix &h, double f, double c2, double s2, double c4, double s4) {
        double ft = h.a0 + fma(h.a2r, c2, h.a2i * s2);
        if constexpr (NH == 2) ft = ft + fma(h.a4r, c4, h.a4i * s4);
        valid = valid && (ft > 0.0) && pb_finite(ft);
        const double d = f - ft;
        acc = acc + (d * d) / ft;
    }
    __device__ __forceinline__ bool decide(const PixArgs &a) const {
        const double chi2 = acc / a.dN;
        return valid && (chi2 <= a.chi2thr) && (chi2 > 0.0);
    }
};

// threshold / ROI part of the mask (PP:2891 + PP:2967) for one pixel; `bits` receives the pixel's ROI bit set
__device__ __forceinline__ bool harm_threshold(const PixArgs &a, long long p, double I, uint32_t &bits) {
    if (a.simple) {                                            // whole image, one threshold (PP:2875-2876)
        bits = 1u;
        return I >= a.roi_ilow[0];
    }
    bits = a.roi_bits ? __ldg(a.roi_bits + p) : 1u;
    return a.mask_in ? (__ldg(a.mask_in + p) != 0) : roi_pass(a, bits, I);
}

// Stage A (PP:2987-3000 decisions) for a pixel that passed the threshold / ROI test: a0 / a2 / a4, certified screening of
// (all fit_i > 0) & (chi2 <= thr) & (chi2 > 0).
//   S0,Sc,Ss,(Sc4,Ss4): the sequential sums of PP:2987-2994 ; Sff: sum f^2 (screening only)
template <int NH>
__device__ __forceinline__ HarmPix harm_stage_a_core(const PixArgs &a, uint32_t bits, double S0, double Sc, double Ss, double Sc4, double Ss4,
                                                     double Sff) {
    HarmPix h;
    h.a0 = pb_nan(); h.a2r = h.a2i = h.a4r = h.a4i = 0.0;
    h.need = false;
    h.bits = bits;
    h.m = true;
    // S0 / N correctly rounded without a division: fma(S0, rhi, S0*rlo) with (rhi, rlo) the double-double 1/N
    // (tests/test_shortcuts.py; x/N is never within 2^-60 of a rounding boundary, the approximation error is 2^-105)
    const double a0 = fma(S0, a.invN_hi, S0 * a.invN_lo);
    if (a0 == 0.0) { h.m = false; return h; }                  // a0 -> NaN (PP:2988): fit is NaN, pixel invalid
    h.a0 = a0;
    h.a2r = a.k2 * Sc; h.a2i = a.k2 * Ss;
    if constexpr (NH == 2) { h.a4r = a.k2 * Sc4; h.a4i = a.k2 * Ss4; }
    // From the orthogonality identity
    //   sum_i (f_i - fit_i)^2 = sum f^2 - N (a0^2 + |a2|^2/2 + |a4|^2/2),  a0 - |a2| - |a4| <= fit_i <= a0 + |a2| + |a4|
    // evaluated in fp64 up to the cancellation, then in fp32 with explicit margins covering every rounding.
    h.need = true;
    if (!(a.flags & PB_FLAG_EXACT_CHI2)) {
        const double q2 = fma(h.a2r, h.a2r, h.a2i * h.a2i);
        double qq = q2;
        const float q2f = (float)q2;
        float ampf = q2f > 0.0f ? q2f * rsqrtf(q2f) : 0.0f;
        if constexpr (NH == 2) {
            const double q4 = fma(h.a4r, h.a4r, h.a4i * h.a4i);
            qq = q2 + q4;
            const float q4f = (float)q4;
            ampf += q4f > 0.0f ? q4f * rsqrtf(q4f) : 0.0f;
        }
        const double model = a.dN * fma(0.5, qq, a0 * a0);
        const float ssef = (float)(Sff - model), a0f = (float)a0, Sf = (float)Sff;
        const float errf = fmaf(4e-12f, Sf, 1e-6f * fabsf(ssef));
        const float fmn = fmaf(a0f, 1.0f - 2e-6f, -ampf * (1.0f + 1e-5f));
        const float fmx = fmaf(a0f, 1.0f + 2e-6f, ampf * (1.0f + 1e-5f));
        const float lo = ssef - errf, hi = ssef + errf;
        if (fmn > 1e-6f * a0f && fmx < 1e30f && Sf < 1e30f) {
            if (lo > 0.0f && hi * (1.0f + 1e-5f) <= a.thr_lo_n * fmn) h.need = false;                 // certainly passes
            else if (lo * (1.0f - 1e-5f) > a.thr_hi_n * fmx) { h.need = false; h.m = false; }         // certainly fails
        }
    }
    return h;
}

// Stage B (PP:3017-3040 outputs, PP:3052, histograms) for a pixel whose fit/chi2 decision is h.m.
// SIMPLE = 1: a.simple holds and PB_FLAG_PROJECT_ONLY is not set (established by the caller)
template <int NH, int SIMPLE = 0>
__device__ __forceinline__ PixOut harm_stage_b(const PixArgs &a, const HistSmem &hs, bool do_hist, const HarmPix &h) {
    PixOut o;
    o.v0 = o.v1 = o.v2 = o.v3 = o.a0m = pb_nan();
    o.m = false;
    if (!h.m) return o;
    if constexpr (NH == 1 && !SIMPLE) {
        if (a.flags & PB_FLAG_PROJECT_ONLY) {      // the disk-independent part only (calibration sweep): a2, a0 and the mask of PP:3000
            o.v0 = h.a2r; o.v1 = h.a2i; o.a0m = h.a0; o.m = true;
            return o;
        }
    }
    const double inv = 1.0 / h.a0;
    double A2 = h.a2r * inv, B2 = h.a2i * inv;
    double v0, v1;
    bool m = true;
    if constexpr (NH == 1) {
        // a non-finite A2/B2 is NaN'ed by divide_ext (PC:58-62) and then falls out of the LUT range: same result here
        lut_lookup(a, A2, B2, v0, v1);
        if (pb_finite(v0)) v0 = pb_pymod(v0 + a.rot_stick, 180.0);
        // the reference keeps a finite rho next to a NaN psi (PP:3021-3024); the mask needs both finite
        m = pb_finite(v0) && pb_finite(v1);
        o.v0 = v0; o.v1 = v1;
    } else {
        double A4 = h.a4r * inv, B4 = h.a4i * inv;
        if (!(pb_finite(A2) && pb_finite(B2))) { A2 = pb_nan(); B2 = pb_nan(); }
        if (!(pb_finite(A4) && pb_finite(B4))) { A4 = pb_nan(); B4 = pb_nan(); }
        double r = (atan2(B2, A2) * PB_RAD2DEG) / 2.0;
        r = pb_pymod(r + a.rot_stick, 180.0);
        const double ab2 = hypot(A2, B2), ab4 = hypot(A4, B4);
        o.v0 = v0 = r;
        o.v1 = v1 = 1.5 * ab2;
        o.v2 = (6.0 * ab4) * cos(4.0 * (0.25 * atan2(B4, A4) - r * PB_DEG2RAD));
        if (a.method == PB_SHG) o.v3 = (-0.5 * (ab4 - ab2)) / (ab4 + ab2) - 0.65;
    }
    o.m = m;
    if (m) o.a0m = h.a0;
    if (do_hist) {
        // non-finite values fall out of every histogram range (hist_add drops them)
        hist_add<SIMPLE>(a, hs, 0, v0, h.bits);
        hist_add<SIMPLE>(a, hs, 1, v1, h.bits);
        if constexpr (NH == 2) {
            hist_add<SIMPLE>(a, hs, 2, o.v2, h.bits);
            if (a.method == PB_SHG) hist_add<SIMPLE>(a, hs, 3, o.v3, h.bits);
        }
    }
    return o;
}

// scalar stores of one pixel's results (float or double maps)
__device__ __forceinline__ void store1(void *base, long long p, double v, bool f64) {
    if (!base) return;
    if (f64) reinterpret_cast<double *>(base)[p] = v; else reinterpret_cast<float *>(base)[p] = (float)v;
}

// the common output set of the throughput mode -- float32 maps for every variable + intmap + mask, nothing else -- is
// recognised once per kernel (uniform) so that the per-pixel store sequence has no per-map pointer / type tests
template <int NH>
__device__ __forceinline__ bool store_is_lean(const PixArgs &a, const pb_outputs &out) {
    bool ok = !(a.flags & PB_FLAG_OUT_F64) && out.var[0] && out.var[1] && out.intmap && out.mask && !out.intensity &&
              !out.rho_angle && !(out.rho_contour && a.edge);
    if constexpr (NH == 2) ok = ok && out.var[2] && (a.method != PB_SHG || out.var[3]);
    return ok;
}

template <int NH>
__device__ __forceinline__ void store_pixel(const PixArgs &a, const pb_outputs &out, long long p, const PixOut &o, double I, bool f64, bool lean = false) {
    if (lean) {
        reinterpret_cast<float *>(out.var[0])[p] = (float)o.v0;
        reinterpret_cast<float *>(out.var[1])[p] = (float)o.v1;
        if constexpr (NH == 2) {
            reinterpret_cast<float *>(out.var[2])[p] = (float)o.v2;
            if (a.method == PB_SHG) reinterpret_cast<float *>(out.var[3])[p] = (float)o.v3;
        }
        reinterpret_cast<float *>(out.intmap)[p] = (float)o.a0m;
        out.mask[p] = (uint8_t)o.m;
        return;
    }
    store1(out.var[0], p, o.v0, f64);
    store1(out.var[1], p, o.v1, f64);
    if constexpr (NH == 2) {
        store1(out.var[2], p, o.v2, f64);
        if (a.method == PB_SHG) store1(out.var[3], p, o.v3, f64);
    }
    store1(out.intmap, p, o.a0m, f64);
    if (out.mask) out.mask[p] = (uint8_t)o.m;
    if (out.intensity && !a.intensity_in) reinterpret_cast<double *>(out.intensity)[p] = I;
    if (out.rho_angle) store1(out.rho_angle, p, (o.v0 == o.v0) ? pb_pymod(o.v0 - a.ref_angle, 180.0) : pb_nan(), f64);
    if (out.rho_contour && a.edge) {
        const double e = __ldg(a.edge + p);
        store1(out.rho_contour, p, (pb_finite(o.v0) && pb_finite(e)) ? pb_pymod(o.v0 - e, 180.0) : pb_nan(), f64);
    }
}

// an invalid pixel on the lean path: NaN maps and a zero mask, no conversions
template <int NH>
__device__ __forceinline__ void store_invalid_lean(const PixArgs &a, const pb_outputs &out, long long p) {
    const float nanf_ = __int_as_float(0x7fc00000);
    reinterpret_cast<float *>(out.var[0])[p] = nanf_;
    reinterpret_cast<float *>(out.var[1])[p] = nanf_;
    if constexpr (NH == 2) {
        reinterpret_cast<float *>(out.var[2])[p] = nanf_;
        if (a.method == PB_SHG) reinterpret_cast<float *>(out.var[3])[p] = nanf_;
    }
    reinterpret_cast<float *>(out.intmap)[p] = nanf_;
    out.mask[p] = 0;
}
