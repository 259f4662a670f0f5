// pb_pixel4.cuh -- the 4POLAR per-pixel evaluation shared by the per-pixel kernel (pb_pixel.cu) and the fused row march (pb_march_tma.cuh,
// 4POLAR with spatial binning).
#pragma once
#include "pb_pixel.cuh"

// 4POLAR 2D / 3D (PP:3001-3014, PP:3041-3051). One pixel; 4 planes.
//
// Two evaluation paths with identical decisions (mask, histogram bins):
//   exact : the reference's statements one rounded fp64 operation at a time (three divisions, sqrt, atan2, two acos): what the
//           double-precision output mode always runs.
//   fast  : float32-output mode. The mask tests and the angles are rewritten on the un-normalised quantities (no division by s):
//               D = m0 - m1, U = 2 m_last, R = sqrt(U^2 + D^2) (fp64)
//           3D: lambda = (s - m2 - R)/(2s);  lambda > 0 <=> s - m2 - R > 0;  lambda < 1/3 <=> 3(m2 + R) - s > 0;  pzz > lambda <=> 3 m2 + R - s > 0
//           2D: lambda = (s - R)/(3s - R);   0 < lambda < 1/3 <=> 0 < R < s
//               rho = atan2(U, D)/2,  psi = 4 asin(sqrt(6 lambda/(3 + sqrt(9 - 24 lambda)))) [= 2 acos((-1 + sqrt(9 - 24 lambda))/2)],
//               eta = atan2(sqrt(2R), sqrt(3 m2 + R - s)) [= acos(sqrt((pzz - lambda)/(1 - 3 lambda)))]
//           -- forms without cancellation, whose float32 evaluation is within PB_P4_EPS degrees of the fp64 value. A decision is
//           taken from the fast value only when that bound proves it: a mask expression farther than 2^-40 s from zero, a value
//           farther than PB_P4_EPS from every histogram edge, range end and wrap point. Anything else is recomputed by the exact
//           statements (per variable), so both paths give the same mask and the same histograms; the float32 maps differ from the
//           rounded fp64 values by at most PB_P4_EPS (north_star asks 1e-3 degrees).
#define PB_P4_EPS 5e-4

struct P4Exact {                    // the reference's normalised quantities, PP:3003-3012
    double pxy, puv, pzz, lam;
};

static __device__ __noinline__ P4Exact p4_exact_core(double m0, double m1, double m2, double mlast, bool three_d) {
    P4Exact e;
    double s = three_d ? ((m0 + m1) + m2) : ((m0 + m1) + 0.0);
    if (s == 0.0) s = pb_nan();
    e.pxy = (m0 - m1) / s;
    e.puv = (2.0 * mlast) / s;
    e.pzz = 0.0;
    if (three_d) {
        e.pzz = m2 / s;
        e.lam = (1.0 - (e.pzz + sqrt(e.puv * e.puv + e.pxy * e.pxy))) / 2.0;
    } else {
        const double p2 = sqrt(e.puv * e.puv + e.pxy * e.pxy);
        e.lam = (1.0 - p2) / (3.0 - p2);
    }
    return e;
}
static __device__ __noinline__ double p4_exact_rho(const P4Exact &e, double rot) { return pb_pymod(0.5 * (atan2(e.puv, e.pxy) * PB_RAD2DEG) + rot, 180.0); }
static __device__ __noinline__ double p4_exact_psi(const P4Exact &e) { return 2.0 * (acos((-1.0 + sqrt(9.0 - 24.0 * e.lam)) / 2.0) * PB_RAD2DEG); }
static __device__ __noinline__ double p4_exact_eta(const P4Exact &e) { return acos(sqrt((e.pzz - e.lam) / (1.0 - 3.0 * e.lam))) * PB_RAD2DEG; }

// Is the histogram bin (or the decision "dropped") of variable `var` the same for every value within PB_P4_EPS of d?
// d is the value as binned (rho already re-wrapped). Mirrors hist_add's rule: [e_k, e_k+1), last bin closed, outside dropped.
__device__ __forceinline__ bool p4_bin_certain(const PixArgs &a, const HistSmem &h, int var, double d) {
    const double *e = h.edges + var * (a.nbins + 1);
    const double lo = e[0], hi = e[a.nbins];
    if (d < lo - PB_P4_EPS || d > hi + PB_P4_EPS) return true;                       // certainly dropped
    if (fabs(d - lo) <= PB_P4_EPS || fabs(d - hi) <= PB_P4_EPS) return false;
    const double x = (d - lo) * a.inv_bin_width[var];
    const double fr = x - floor(x);
    const double w = PB_P4_EPS * a.inv_bin_width[var];                               // the bound in units of a bin
    return fr > w && (1.0 - fr) > w;
}

// One pixel: f = the four pre-processed (and, where asked, binned) planes in the order of PP:2957 (acquisition planes 1 and 2 swapped),
// I = the pixel's threshold-image value (PC:260-264). Writes every requested map and adds to the histograms.
template <bool FAST>
__device__ __forceinline__ void p4_pixel(const PixArgs &a, const HistSmem &hs, bool do_hist, const pb_outputs &out, long long p, const double *f, double I) {
    const bool out64 = a.flags & PB_FLAG_OUT_F64;
    const bool three_d = a.method == PB_4POLAR_3D;
    double m4[4] = {0.0, 0.0, 0.0, 0.0};
    const int rows = three_d ? 4 : 3;
    for (int r = 0; r < rows; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc = acc + a.invk[r * 4 + j] * f[j];
        m4[r] = acc;
    }
    const double mlast = m4[three_d ? 3 : 2];
    double a0 = ((((f[0] + f[1]) + f[2]) + f[3]) / 4.0) / 4.0;
    if (a0 == 0.0) a0 = pb_nan();
    const uint32_t bits = a.roi_bits ? __ldg(a.roi_bits + p) : 1u;
    bool m = a.mask_in ? (__ldg(a.mask_in + p) != 0) : roi_pass(a, bits, I);
    double v0 = pb_nan(), v1 = pb_nan(), v2 = pb_nan();
    bool done = false;
    if constexpr (FAST) {
        // ---- certified fast path (see the header of this section)
        const double s = three_d ? ((m4[0] + m4[1]) + m4[2]) : (m4[0] + m4[1]);
        const double D = m4[0] - m4[1], U = 2.0 * mlast;
        const double R = sqrt(fma(U, U, D * D));
        const double tol = s * 9.094947017729282e-13;                      // 2^-40 s
        if (s > 0.0 && s < 1e300 && pb_finite(R)) {
            double nl, n2 = 0.0, n3 = 0.0;
            int verdict;                                                   // 1: mask certainly true, 0: certainly false, -1: undecided
            if (three_d) {
                nl = (s - m4[2]) - R;                                      // 2 s lambda
                n2 = 3.0 * (m4[2] + R) - s;                                // 2 s (1 - 3 lambda)
                n3 = fma(3.0, m4[2], R) - s;                               // 2 s (pzz - lambda)
                verdict = (nl > tol && n2 > tol && n3 > tol) ? 1 : ((nl < -tol || n2 < -tol || n3 < -tol) ? 0 : -1);
            } else {
                nl = s - R;                                                // (3 s - R) lambda
                verdict = (nl > tol && R > tol) ? 1 : ((nl < -tol) ? 0 : -1);
            }
            if (verdict == 0 || !m) {
                m = false;
                done = true;
            } else if (verdict == 1) {
                // float32 evaluation of the rewritten angles
                const float Uf = (float)U, Df = (float)D, sf = (float)s, Rf = (float)R;
                float rf = 28.64788975654116f * atan2f(Uf, Df);           // 0.5 * rad2deg
                double r0 = pb_pymod((double)rf + a.rot_stick, 180.0);
                const float lamf = three_d ? (float)nl / (2.0f * sf) : (float)nl / (3.0f * sf - Rf);
                const float wf = sqrtf(9.0f - 24.0f * lamf);
                float psf = 229.1831180523293f * asinf(sqrtf(6.0f * lamf / (3.0f + wf)));   // 4 * rad2deg
                float etf = 0.0f;
                if (three_d) etf = 57.29577951308232f * atan2f(sqrtf(2.0f * Rf), sqrtf((float)n3));
                // every decision taken from these values must be provable: wrap point of rho, histogram bins
                // (eta: near isotropy both pzz - lambda and 1 - 3 lambda cancel and the reference's own quotient is ill-conditioned:
                //  there the exact statements decide)
                const double tol20 = s * 9.5367431640625e-07;              // 2^-20 s
                bool ok0 = r0 > PB_P4_EPS && r0 < 180.0 - PB_P4_EPS, ok1 = true, ok2 = !three_d || (n2 > tol20 && n3 > tol20);
                if (do_hist) {
                    double d0 = r0;
                    if ((a.hist_is_rho & 1u) && a.rot_fig != 0.0) {
                        d0 = pb_pymod(2.0 * (r0 + a.rot_fig), 360.0) / 2.0;
                        ok0 = ok0 && d0 > PB_P4_EPS && d0 < 180.0 - PB_P4_EPS;
                    }
                    ok0 = ok0 && p4_bin_certain(a, hs, 0, d0);
                    ok1 = p4_bin_certain(a, hs, 1, (double)psf);
                    if (three_d) ok2 = ok2 && p4_bin_certain(a, hs, 2, (double)etf);
                }
                v0 = r0; v1 = (double)psf; v2 = three_d ? (double)etf : pb_nan();
                if (!(ok0 && ok1 && ok2)) {
                    const P4Exact e = p4_exact_core(m4[0], m4[1], m4[2], mlast, three_d);
                    if (!ok0) v0 = p4_exact_rho(e, a.rot_stick);
                    if (!ok1) v1 = p4_exact_psi(e);
                    if (!ok2) v2 = p4_exact_eta(e);
                }
                done = true;
            }
        }
    }
    if (!done) {
        // ---- the reference's statements (PP:3003-3012, PP:3042-3051)
        const P4Exact e = p4_exact_core(m4[0], m4[1], m4[2], mlast, three_d);
        m = m && (e.lam < 1.0 / 3.0) && (e.lam > 0.0) && (three_d ? (e.pzz > e.lam) : true);
        if (m) {
            v0 = p4_exact_rho(e, a.rot_stick);
            v1 = p4_exact_psi(e);
            if (three_d) v2 = p4_exact_eta(e);
        }
    }
    if (!m) v0 = v1 = v2 = pb_nan();
    store1(out.var[0], p, v0, out64);
    store1(out.var[1], p, v1, out64);
    if (three_d) store1(out.var[2], p, v2, out64);
    store1(out.intmap, p, m ? a0 : pb_nan(), out64);
    if (out.intensity && !a.intensity_in) reinterpret_cast<double *>(out.intensity)[p] = I;
    if (out.mask) out.mask[p] = (uint8_t)m;
    if (out.rho_angle) store1(out.rho_angle, p, (v0 == v0) ? pb_pymod(v0 - a.ref_angle, 180.0) : pb_nan(), out64);
    if (out.rho_contour && a.edge) {
        double e = __ldg(a.edge + p);
        store1(out.rho_contour, p, (pb_finite(v0) && pb_finite(e)) ? pb_pymod(v0 - e, 180.0) : pb_nan(), out64);
    }
    if (do_hist && m) {
        hist_add(a, hs, 0, v0, bits);
        hist_add(a, hs, 1, v1, bits);
        if (three_d && pb_finite(v2)) hist_add(a, hs, 2, v2, bits);
    }
}
