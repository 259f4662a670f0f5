/* TEST INFRASTRUCTURE ONLY -- op-for-op fp64 arithmetic contract of PyPOLAR's per-pixel analysis path.
 *
 * Plain scalar C restatement of what the reference's NumPy/SciPy statements compute, one separately
 * rounded IEEE-754 double operation at a time, in the reference's evaluation order.
 * It exists to pin, on a CPU, the exact sequence the CUDA kernels replay. It is checked bit-for-bit
 * against oracle/restate.py (which calls the same NumPy/SciPy routines as the reference) and against the
 * golden vectors generated from the unmodified reference (tests/golden). It is never linked into or
 * called by the product.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC  (contraction MUST stay off; the one
 * fused operation of the path is spelled fma()).
 *
 * PP = pypolar/PyPOLAR.py, PC = pypolar/pypolar_classes.py (cchandre/Polarimetry v2.9.4).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { PC_U8 = 0, PC_U16 = 1, PC_F32 = 2, PC_F64 = 3 };
enum { M_1PF = 0, M_CARS = 1, M_SRS = 2, M_SHG = 3, M_2PF = 4, M_4POLAR2D = 5, M_4POLAR3D = 6 };

static double load_val(const void *v, int dtype, long k) {
    switch (dtype) {
    case PC_U8: return (double)((const uint8_t *)v)[k];
    case PC_U16: return (double)((const uint16_t *)v)[k];
    case PC_F32: return (double)((const float *)v)[k];
    default: return ((const double *)v)[k];
    }
}

/* numpy maximum(x, 0): NaN propagates */
static double max0(double x) { return (x != x) ? x : (x > 0.0 ? x : 0.0); }

/* scipy.ndimage 'reflect' (half-sample symmetric) index: d c b a | a b c d | d c b a */
static long reflect_idx(long k, long L) {
    while (k < 0 || k >= L) {
        if (k < 0) k = -k - 1;
        if (k >= L) k = 2 * L - 1 - k;
    }
    return k;
}

/* scipy.ndimage.uniform_filter1d, mode='reflect', origin 0, on one strided line (in place allowed
 * through the extension buffer `ext`, L + n doubles). Running sum: tmp += new - old; out = tmp / n.
 * Restates the 1-D passes behind PP:2945 and PC:264. */
static void box_line(double *line, long stride, long L, int n, double *ext) {
    long h = n / 2, E = L + n - 1;
    for (long k = 0; k < E; ++k) ext[k] = line[reflect_idx(k - h, L) * stride];
    double tmp = 0.0;
    for (int k = 0; k < n; ++k) tmp = tmp + ext[k];
    line[0] = tmp / (double)n;
    for (long k = 1; k < L; ++k) {
        tmp = tmp + (ext[k + n - 1] - ext[k - 1]);
        line[k * stride] = tmp / (double)n;
    }
}

/* uniform_filter(f, size=[0,bh,bw]) on [P,H,W]: axis 1 (H) first, then axis 2 (W); sizes <= 1 skipped. */
void pc_box(double *f, long P, long H, long W, int bh, int bw) {
    long m = (H > W ? H : W) + (bh > bw ? bh : bw) + 2;
    double *ext = (double *)malloc(sizeof(double) * m);
    for (long p = 0; p < P; ++p) {
        double *pl = f + p * H * W;
        if (bh > 1)
            for (long x = 0; x < W; ++x) box_line(pl + x, W, H, bh, ext);
        if (bw > 1)
            for (long y = 0; y < H; ++y) box_line(pl + y * W, 1, W, bw, ext);
    }
    free(ext);
}

/* field = max(values - sub, 0) ; CARS: sqrt ; 4POLAR: planes 1 <-> 2. Restates PP:2953-2957. */
void pc_field(const void *v, int dtype, long N, long H, long W, double sub, int do_sqrt, int swap12, double *out) {
    long P = H * W;
    for (long i = 0; i < N; ++i) {
        long src = i;
        if (swap12 && (i == 1 || i == 2)) src = 3 - i;
        for (long p = 0; p < P; ++p) {
            double x = max0(load_val(v, dtype, src * P + p) - sub);
            out[i * P + p] = do_sqrt ? sqrt(x) : x;
        }
    }
}

/* Total-intensity image: sum_i max(v_i - dark, 0), then 2-D box filter. Restates PC:260-264. */
void pc_intensity(const void *v, int dtype, long N, long H, long W, double dark, int bh, int bw, double *out) {
    long P = H * W;
    for (long p = 0; p < P; ++p) {
        double s = max0(load_val(v, dtype, p) - dark);
        for (long i = 1; i < N; ++i) s = s + max0(load_val(v, dtype, i * P + p) - dark);
        out[p] = s;
    }
    if (!(bh == 1 && bw == 1)) pc_box(out, 1, H, W, bh, bw);
}

static int finite_d(double x) { return x == x && x - x == 0.0; }

/* np.mod(a, b), b > 0 (python-sign modulo). */
static double pymod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) {
        if (r < 0.0) r = r + b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

/* Harmonic block PP:2987-3000. nharm = 1 (1PF) or 2 (CARS/SRS/SHG/2PF).
 * Outputs: a0 (NaN where 0), normalised A2,B2,(A4,B4), chi2, mask (in/out, uint8). fit (optional, may be
 * NULL) receives field_fit [N,H,W]. */
void pc_harmonics(const double *f, long N, long H, long W, int nharm, const double *c2, const double *s2,
                  const double *c4, const double *s4, double chi2thr, uint8_t *mask, double *a0o, double *A2,
                  double *B2, double *A4, double *B4, double *chi2o, double *fit_out) {
    long P = H * W;
    double k2 = 2.0 / (double)N;
    double *fit = (double *)malloc(sizeof(double) * N);
    for (long p = 0; p < P; ++p) {
        double S0 = 0.0, Sc = 0.0, Ss = 0.0, Sc4 = 0.0, Ss4 = 0.0;
        for (long i = 0; i < N; ++i) {
            double x = f[i * P + p];
            S0 = S0 + x;
            Sc = Sc + x * c2[i];
            Ss = Ss + x * s2[i];
            if (nharm == 2) {
                Sc4 = Sc4 + x * c4[i];
                Ss4 = Ss4 + x * s4[i];
            }
        }
        double a0 = S0 / (double)N;
        if (a0 == 0.0) a0 = NAN;
        double a2r = k2 * Sc, a2i = k2 * Ss, a4r = k2 * Sc4, a4i = k2 * Ss4;
        int valid = 1;
        double acc = 0.0;
        for (long i = 0; i < N; ++i) {
            /* numpy's SIMD complex multiply: Re(a2 * conj(e2)) = fma(a2r, c, -(a2i * (-s))) */
            double ft = a0 + fma(a2r, c2[i], -(a2i * (-s2[i])));
            if (nharm == 2) ft = ft + fma(a4r, c4[i], -(a4i * (-s4[i])));
            fit[i] = ft;
            if (!(ft > 0.0 && finite_d(ft))) valid = 0;
            double d = f[i * P + p] - ft;
            acc = acc + (d * d) / ft;
            if (fit_out) fit_out[i * P + p] = ft;
        }
        double chi2 = acc / (double)N;
        double inv = 1.0 / a0;
        double x;
        x = a2r * inv; A2[p] = finite_d(x) ? x : NAN;
        x = a2i * inv; B2[p] = finite_d(x) ? x : NAN;
        /* divide_ext NaNs the whole complex number if either part is non-finite (PC:61) */
        if (!(finite_d(A2[p]) && finite_d(B2[p]))) A2[p] = B2[p] = NAN;
        if (nharm == 2) {
            x = a4r * inv; A4[p] = x;
            x = a4i * inv; B4[p] = x;
            if (!(finite_d(A4[p]) && finite_d(B4[p]))) A4[p] = B4[p] = NAN;
        }
        a0o[p] = a0;
        chi2o[p] = chi2;
        mask[p] = (uint8_t)(mask[p] && valid && (chi2 <= chi2thr) && (chi2 > 0.0));
    }
    free(fit);
}

/* scipy RegularGridInterpolator cell search: g[i] <= x < g[i+1], last cell closed. */
static long find_cell(const double *g, long n, double x) {
    if (x == g[n - 1]) return n - 2;
    long lo = 0, hi = n - 1; /* invariant g[lo] <= x < g[hi] */
    while (hi - lo > 1) {
        long mid = (lo + hi) / 2;
        if (x >= g[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

static double bilinear(const double *T, long n, long i0, long i1, double t0, double t1) {
    double u0 = 1.0 - t0, u1 = 1.0 - t1;
    double w00 = (1.0 * u0) * u1, w01 = (1.0 * u0) * t1, w10 = (1.0 * t0) * u1, w11 = (1.0 * t0) * t1;
    double v = 0.0;
    v = v + T[i0 * n + i1] * w00;
    v = v + T[i0 * n + i1 + 1] * w01;
    v = v + T[(i0 + 1) * n + i1] * w10;
    v = v + T[(i0 + 1) * n + i1 + 1] * w11;
    return v;
}

/* 1PF inversion through the disk-cone table. Restates PP:3017-3026 (+ interpn semantics). */
void pc_lut_1pf(const double *A2, const double *B2, long P, const double *rho_t, const double *psi_t,
                const double *g, long n, double rotation, uint8_t *mask, double *rho, double *psi) {
    for (long p = 0; p < P; ++p) {
        rho[p] = NAN; psi[p] = NAN;
        if (!mask[p]) continue;
        double x = A2[p], y = B2[p], r = NAN, q = NAN;
        if (x == x && y == y && !(x < g[0]) && !(x > g[n - 1]) && !(y < g[0]) && !(y > g[n - 1])) {
            long i0 = find_cell(g, n, x), i1 = find_cell(g, n, y);
            double t0 = (x - g[i0]) / (g[i0 + 1] - g[i0]);
            double t1 = (y - g[i1]) / (g[i1 + 1] - g[i1]);
            r = bilinear(rho_t, n, i0, i1, t0, t1);
            q = bilinear(psi_t, n, i0, i1, t0, t1);
        }
        if (finite_d(r)) r = pymod(r + rotation, 180.0);
        rho[p] = r; psi[p] = q;
        mask[p] = (uint8_t)(finite_d(r) && finite_d(q));
    }
}

static const double RAD2DEG = 180.0 / 3.141592653589793238462643383279502884;
static const double DEG2RAD = 3.141592653589793238462643383279502884 / 180.0;

/* 4-harmonic outputs. Restates PP:3027-3040. sshg may be NULL. */
void pc_out_4harm(const double *A2, const double *B2, const double *A4, const double *B4, long P, double rotation,
                  const uint8_t *mask, double *rho, double *s2, double *s4, double *sshg) {
    for (long p = 0; p < P; ++p) {
        rho[p] = NAN; s2[p] = NAN; s4[p] = NAN;
        if (sshg) sshg[p] = NAN;
        if (!mask[p]) continue;
        double r = (atan2(B2[p], A2[p]) * RAD2DEG) / 2.0;
        r = pymod(r + rotation, 180.0);
        double ab2 = hypot(A2[p], B2[p]), ab4 = hypot(A4[p], B4[p]);
        rho[p] = r;
        s2[p] = 1.5 * ab2;
        s4[p] = (6.0 * ab4) * cos(4.0 * (0.25 * atan2(B4[p], A4[p]) - r * DEG2RAD));
        if (sshg) sshg[p] = (-0.5 * (ab4 - ab2)) / (ab4 + ab2) - 0.65;
    }
}

/* 4POLAR 2D/3D. f has 4 planes already swapped to (0,90,45,135). Restates PP:3001-3014, PP:3041-3051.
 * invK is [rows,4] row-major, rows = 3 (2D) or 4 (3D). eta may be NULL (2D). */
void pc_4polar(const double *f, long P, const double *invK, int three_d, double rotation, uint8_t *mask,
               double *a0o, double *rho, double *psi, double *eta) {
    int rows = three_d ? 4 : 3;
    for (long p = 0; p < P; ++p) {
        double m[4] = {0, 0, 0, 0};
        for (int r = 0; r < rows; ++r) {
            double acc = 0.0;
            for (int j = 0; j < 4; ++j) acc = acc + invK[r * 4 + j] * f[j * P + p];
            m[r] = acc;
        }
        double s = three_d ? ((m[0] + m[1]) + m[2]) : ((m[0] + m[1]) + 0.0);
        if (s == 0.0) s = NAN;
        double pxy = (m[0] - m[1]) / s;
        double puv = (2.0 * m[three_d ? 3 : 2]) / s;
        double pzz = 0.0, lam;
        if (three_d) {
            pzz = m[2] / s;
            lam = (1.0 - (pzz + sqrt(puv * puv + pxy * pxy))) / 2.0;
        } else {
            double p2 = sqrt(puv * puv + pxy * pxy);
            lam = (1.0 - p2) / (3.0 - p2);
        }
        double a0 = ((((f[p] + f[P + p]) + f[2 * P + p]) + f[3 * P + p]) / 4.0) / 4.0;
        if (a0 == 0.0) a0 = NAN;
        int ok = mask[p] && (lam < 1.0 / 3.0) && (lam > 0.0) && (three_d ? (pzz > lam) : 1);
        mask[p] = (uint8_t)ok;
        rho[p] = NAN; psi[p] = NAN;
        if (eta) eta[p] = NAN;
        a0o[p] = a0;
        if (!ok) continue;
        rho[p] = pymod(0.5 * (atan2(puv, pxy) * RAD2DEG) + rotation, 180.0);
        psi[p] = 2.0 * (acos((-1.0 + sqrt(9.0 - 24.0 * lam)) / 2.0) * RAD2DEG);
        if (eta) eta[p] = acos(sqrt((pzz - lam) / (1.0 - 3.0 * lam))) * RAD2DEG;
    }
}

/* intmap = a0 where mask else NaN (PP:3052); Rho_angle (PP:3069-3072); Rho_contour (PP:3059-3068). */
void pc_tail(long P, const uint8_t *mask, double *a0, const double *rho, double reference_angle, double *rho_angle,
             const double *edge, double *rho_contour) {
    for (long p = 0; p < P; ++p) {
        if (!mask[p]) a0[p] = NAN;
        if (rho_angle) rho_angle[p] = (rho[p] == rho[p]) ? pymod(rho[p] - reference_angle, 180.0) : NAN;
        if (rho_contour) rho_contour[p] = (finite_d(rho[p]) && finite_d(edge[p])) ? pymod(rho[p] - edge[p], 180.0) : NAN;
    }
}

/* Variable.histo numerics, PC:340-369 + np.histogram with explicit edges:
 * sel & finite(X); Rho/Rho_angle: d = mod(2(d+rot),360)/2; polar3: d>=90 -> 180-d;
 * bins [e_k, e_{k+1}), last closed; outside dropped. */
void pc_hist(const double *X, const uint8_t *sel, long P, int is_rho, double rotation_fig, int fold90,
             const double *edges, int nbins, int64_t *counts) {
    for (int k = 0; k < nbins; ++k) counts[k] = 0;
    for (long p = 0; p < P; ++p) {
        double d = X[p];
        if ((sel && !sel[p]) || !finite_d(d)) continue;
        if (is_rho) d = pymod(2.0 * (d + rotation_fig), 360.0) / 2.0;
        if (fold90 && d >= 90.0) d = 180.0 - d;
        if (d < edges[0] || d > edges[nbins]) continue;
        int lo = 0, hi = nbins; /* edges[lo] <= d, find largest lo with edges[lo] <= d */
        if (d == edges[nbins]) { counts[nbins - 1]++; continue; }
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (d >= edges[mid]) lo = mid; else hi = mid;
        }
        counts[lo]++;
    }
}

/* ---- checks of the exact-arithmetic shortcuts the CUDA kernels use (tests/test_shortcuts.py) ------------ */
static uint64_t rng_next(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

/* x / n == fma(x, hi, x*lo) with (hi, lo) the double-double reciprocal of n: returns the number of mismatches
 * over `count` random doubles (mixed: integers, thirds, running sums) plus all integers below 2^20. */
long pc_check_div_by_const(int n, long count, uint64_t seed) {
    double d = (double)n, hi = 1.0 / d, lo = fma(-hi, d, 1.0) / d;
    long bad = 0;
    uint64_t s = seed ? seed : 88172645463325252ULL;
    for (long k = 0; k < (1L << 20); ++k) {
        double x = (double)k;
        if (fma(x, hi, x * lo) != x / d) ++bad;
    }
    for (long k = 0; k < count; ++k) {
        uint64_t r = rng_next(&s);
        double x;
        switch (k & 3) {
        case 0: x = (double)(r >> 11) * (1.0 / 9007199254740992.0) * 1048576.0; break;      /* uniform [0, 2^20) */
        case 1: x = (double)(r % 16777216ULL) / 3.0 + (double)((r >> 32) % 65536ULL) / 7.0; break;
        case 2: x = ldexp((double)(r >> 11), (int)(r % 80) - 60); break;                      /* wide exponent range */
        default: x = (double)(r % 4294967296ULL) / (double)((r >> 40) % 20 + 1); break;
        }
        if (fma(x, hi, x * lo) != x / d) ++bad;
    }
    return bad;
}

/* RN(f * c) == fma(2^52 + f, c, -(2^52 * c)) for integers 0 <= f < 2^32: mismatches over all f < 2^16 and random f */
long pc_check_magic_product(double c, long count, uint64_t seed) {
    long bad = 0;
    uint64_t s = seed ? seed : 1234567ULL;
    const double T = 4503599627370496.0, k = -(T * c);
    for (long f = 0; f < 65536 + count; ++f) {
        double fv = f < 65536 ? (double)f : (double)(rng_next(&s) % 4294967296ULL);
        double a = fma(T + fv, c, k), b = fv * c;
        if (a != b && !(a == 0.0 && b == 0.0)) ++bad;
    }
    return bad;
}

/* d / h == q with y = RN(1/h), q0 = RN(d*y), r = fma(-q0, h, d) (exact), q = fma(r, y, q0)  (Markstein's correctly
 * rounded division from a correctly rounded reciprocal). Used for the LUT cell coordinate t = (x - g[i]) / (g[i+1] - g[i])
 * where h takes the n-1 values of the grid spacing. Returns mismatches over `count` numerators 0 <= d <= 1.5 h. */
long pc_check_markstein_div(double h, long count, uint64_t seed) {
    const double y = 1.0 / h;
    long bad = 0;
    uint64_t s = seed ? seed : 99991ULL;
    for (long k = 0; k < count; ++k) {
        uint64_t r = rng_next(&s);
        double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);        /* [0,1) */
        double d;
        switch (k & 3) {
        case 0: d = u * h; break;
        case 1: d = u * h * 1.5; break;
        case 2: d = ldexp(u, -(int)(r % 40)) * h; break;                    /* tiny numerators */
        default: d = (double)(r % 1000003ULL) * (h / 1000003.0); break;
        }
        double q0 = d * y, rr = fma(-q0, h, d), q = fma(rr, y, q0);
        if (q != d / h) ++bad;
    }
    return bad;
}

/* ---- the arithmetic of the rows either side of the path ----------------------------------------------- */
/* define_stack's rescale of a non-integer stack (PyPOLAR.py:1929-1930), one separately rounded operation at a time in the
 * array's precision: d = a - amin ; m = 65535 * d ; q = m / ptp ; truncation to uint16. */
void pc_rescale_f32(const float *a, long n, uint16_t *out) {
    float mn = a[0], mx = a[0];
    for (long k = 1; k < n; ++k) { if (a[k] < mn) mn = a[k]; if (a[k] > mx) mx = a[k]; }
    const volatile float ptp = mx - mn;
    for (long k = 0; k < n; ++k) {
        volatile float d = a[k] - mn;
        volatile float m = 65535.0f * d;
        volatile float q = m / ptp;
        out[k] = (uint16_t)(q >= 65535.0f ? 65535u : (q > 0.0f ? (unsigned)q : 0u));
    }
}
void pc_rescale_f64(const double *a, long n, uint16_t *out) {
    double mn = a[0], mx = a[0];
    for (long k = 1; k < n; ++k) { if (a[k] < mn) mn = a[k]; if (a[k] > mx) mx = a[k]; }
    const volatile double ptp = mx - mn;
    for (long k = 0; k < n; ++k) {
        volatile double d = a[k] - mn;
        volatile double m = 65535.0 * d;
        volatile double q = m / ptp;
        out[k] = (uint16_t)(q >= 65535.0 ? 65535u : (q > 0.0 ? (unsigned)q : 0u));
    }
}

/* IEEE binary64 -> binary16, ONE rounding to nearest even (what numpy's astype(np.float16) does, PP:2839, and what the
 * device's cvt.rn.f16.f64 does): bit pattern of the half. */
uint16_t pc_f64_to_f16(double x) {
    uint64_t u; memcpy(&u, &x, 8);
    const uint16_t sign = (uint16_t)((u >> 48) & 0x8000u);
    const int e = (int)((u >> 52) & 0x7ff);
    const uint64_t man = u & 0xfffffffffffffull;
    if (e == 0x7ff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0u));             /* inf / NaN */
    if (e == 0) return sign;                                                              /* zero and binary64 subnormals -> 0 */
    const int eh = e - 1023 + 15;                                                         /* biased half exponent */
    if (eh >= 31) return (uint16_t)(sign | 0x7c00u);                                      /* overflow -> inf */
    uint64_t sig = man | (1ull << 52);                                                    /* 53-bit significand */
    int shift = 42;                                                                       /* 52 - 10 */
    uint16_t hexp = (uint16_t)eh;
    if (eh <= 0) {                                                                        /* half subnormal (or underflow) */
        shift += 1 - eh;
        hexp = 0;
        if (shift > 63) return sign;
    }
    const uint64_t q = sig >> shift, rem = sig & ((1ull << shift) - 1), half = 1ull << (shift - 1);
    uint64_t r = q + ((rem > half || (rem == half && (q & 1))) ? 1 : 0);
    /* normal: q has the implicit bit at position 10; adding the exponent field handles the carry into the next binade */
    uint32_t h = (eh > 0) ? (uint32_t)(((uint32_t)(hexp - 1) << 10) + (uint32_t)r) : (uint32_t)r;
    if (h >= 0x7c00u) h = 0x7c00u;
    return (uint16_t)(sign | h);
}

/* save_mat's gather (PP:2838-2840): values[sel & isfinite(values)] in row-major order, cast to float16. Returns the count. */
long pc_compact_f16(const double *v, const uint8_t *sel, long n, uint16_t *out) {
    long c = 0;
    for (long k = 0; k < n; ++k)
        if ((!sel || sel[k]) && finite_d(v[k])) out[c++] = pc_f64_to_f16(v[k]);
    return c;
}
