// pb_march.cu -- fused single-pass analysis WITH spatial binning: the stack is read once and the maps,
// mask and histograms come out, with scipy's two-pass running-sum box filter replayed bit for bit.
//
// Replaces, in one kernel: Stack.get_intensity (pypolar_classes.py:260-264), analyze_stack's pre-processing
// (PyPOLAR.py:2953-2957), Polarimetry.binning (PP:2942-2946), compute_fields (PP:2981-3078), the mask logic
// (PP:2891/2967) and the histograms (PC:340-369).
//
// Why a "row march". scipy.ndimage.uniform_filter1d keeps a running sum along each line,
//     tmp += ext[k+n-1] - ext[k-1];  out[k] = tmp / n,
// so along W (the second pass, whose inputs RN(S/bh) are no longer integers) every output carries the
// rounding history of its whole row from column 0: the bits -- hence the LUT cell, hence the histogram bin
//  -- can only be reproduced by walking each (plane, row) line sequentially. The first pass
// (along H) sums integers, which is exact in any order, so it is done as a direct window sum.
//
// Decomposition (one CTA = 32 rows of one stack, marching over W in steps of XT extended columns):
//   loader : raw stack -> shared memory, already dark-subtracted and clamped (uint16), plus the per-position
//            total-intensity partial sums; next step's global loads are in flight during the phases below.
//   phase A: one thread per (plane, row) line ("chain"): integer window sum over bh rows, q = RN(S/bh),
//            tmp += q_new - q_old (register ring of BW), f = RN(tmp/bw) -> shared memory (fp64).
//            Plane N is the total-intensity image the threshold compares against.
//   phase B: one thread per pixel of the step: sequential harmonic sums over the N planes in the reference's
//            order (PP:2987-2994), then the shared per-pixel epilogue (pb_pixel.cuh).
// Divisions by the bin sizes are the two-instruction correctly rounded form fma(x, rhi, x*rlo) (tests/test_shortcuts.py).
#include <type_traits>
#include "pb_pixel.cuh"

void pb_ddrecip(int n, double *hi, double *lo);   // pb_api.cu

namespace {

constexpr int MR = 32;   // rows per CTA (= lanes of a chain warp)
constexpr int PL = 2;    // planes (lines) per chain thread

struct MarchArgs {
    int bh, hh, hw;              // H-window size, its left extent bh/2, W-window left extent bw/2
    int steps;                   // number of XT-wide steps over the extended row (W + bw - 1)
    int tile_rows;               // MR + bh - 1
    int rp;                      // raw tile row pitch in uint16 (2 * odd)
    int groups;                  // loader groups splitting the angle loop
    int npos;                    // tile_rows * XT loader positions
    double bh_hi, bh_lo, bw_hi, bw_lo;   // double-double reciprocals of the bin sizes
    unsigned off_T, off_f;       // byte offsets of the T partials and of fbuf in dynamic shared memory (after the histogram area)
    unsigned off_raw;
};

template <int BW> struct StepWidth { static constexpr int M = (8 + BW - 1) / BW; static constexpr int XT = BW * M; };
template <> struct StepWidth<1> { static constexpr int M = 8; static constexpr int XT = 8; };

template <typename UT, int NH, int BW>
__global__ void __launch_bounds__(512, 2) k_march(PixArgs a, typename BasisSel<NH>::type bs, MarchArgs ma) {
    constexpr int XT = StepWidth<BW>::XT;
    constexpr int FP = (XT % 2) ? XT : XT + 1;          // fbuf row pitch in doubles (odd: conflict-free chain stores)
    constexpr int MAXLD = 24;                           // register budget for the prefetched raw values per thread
    extern __shared__ __align__(16) unsigned char smem[];
    const pb_outputs out = a.outs[blockIdx.y];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    HistSmem hs;
    hs.cnt = nullptr; hs.edges = nullptr;
    if (do_hist) hs = hist_setup(a, smem);
    uint16_t *raw = reinterpret_cast<uint16_t *>(smem + ma.off_raw);      // [N][tile_rows][rp] clamped (v - sub)
    int *Tp = reinterpret_cast<int *>(smem + ma.off_T);                   // [groups][tile_rows][XT] partial sums of clamp(v - dark)
    double *fbuf = reinterpret_cast<double *>(smem + ma.off_f);           // [N+1][MR][FP]

    const int N = a.N, H = a.H, W = a.W;
    const long long P = a.P;
    const UT *stack = reinterpret_cast<const UT *>(a.stack) + (long long)blockIdx.y * a.stack_stride;
    const int y0 = blockIdx.x * MR;
    const int tid = threadIdx.x, lane = tid & 31, plane0 = (tid >> 5) * PL;   // chains (plane0 .. plane0+PL-1, row y0 + lane)
    const int isub = (int)a.sub, idark = (int)a.dark;
    const bool sep = isub != idark;
    const int bh = ma.bh, TR = ma.tile_rows, RP = ma.rp, G = ma.groups, NPOS = ma.npos;
    const int plane_pitch = TR * RP;

    // loader role: position (r', j) of the tile, angles i = g, g + G, ...
    const int lpos = tid % NPOS, lg = tid / NPOS;
    const bool ldr = lg < G;
    const int lr = lpos / XT, lj = lpos - lr * XT;
    const long long lrow_off = (long long)pb_reflect(y0 - ma.hh + lr, H) * W;
    unsigned ldv[MAXLD];

    auto prefetch = [&](int t) {
        if (!ldr) return;
        const int col = pb_reflect(t * XT + lj - ma.hw, W);
        const UT *src = stack + lrow_off + col;
#pragma unroll
        for (int u = 0; u < MAXLD; ++u) {
            const int i = lg + u * G;
            if (i < N) ldv[u] = (unsigned)__ldg(src + (long long)i * P);
        }
    };
    auto commit = [&]() {
        if (!ldr) return;
        int tsum = 0;
        uint16_t *dst = raw + lr * RP + lj;
#pragma unroll
        for (int u = 0; u < MAXLD; ++u) {
            const int i = lg + u * G;
            if (i < N) {
                const int v = (int)ldv[u];
                const int f = max(v - isub, 0);
                dst[i * plane_pitch] = (uint16_t)f;
                tsum += sep ? max(v - idark, 0) : f;
            }
        }
        Tp[lg * NPOS + lpos] = tsum;
    };

    // chain state
    double tmp[PL], ring[PL][BW];
#pragma unroll
    for (int c = 0; c < PL; ++c) {
        tmp[c] = 0.0;
#pragma unroll
        for (int s = 0; s < BW; ++s) ring[c][s] = 0.0;
    }

    prefetch(0);
    for (int t = 0; t < ma.steps; ++t) {
        commit();
        __syncthreads();
        if (t + 1 < ma.steps) prefetch(t + 1);
        // ---- phase A: the (plane, row) lines advance XT extended positions
#pragma unroll
        for (int c = 0; c < PL; ++c) {
            const int plane = plane0 + c;
            if (plane > N) continue;
            const bool is_int_plane = plane == N;
            double *fo = fbuf + (plane * MR + lane) * FP;
            const int e0 = t * XT;
#pragma unroll
            for (int j = 0; j < XT; ++j) {
                int S1 = 0;
                if (!is_int_plane) {
                    const uint16_t *rp = raw + plane * plane_pitch + lane * RP + j;
                    for (int dy = 0; dy < bh; ++dy) S1 += (int)rp[dy * RP];
                } else {
                    for (int g = 0; g < G; ++g) {
                        const int *tp = Tp + g * NPOS + lane * XT + j;
                        for (int dy = 0; dy < bh; ++dy) S1 += tp[dy * XT];
                    }
                }
                // (double)S1 exactly: (2^52 + S1) - 2^52 ; then RN(S1 / bh) unless the H pass is skipped (bh <= 1, PP:2944)
                const double u = __hiloint2double(0x43300000, S1) - 4503599627370496.0;
                const double q = (bh > 1) ? fma(u, ma.bh_hi, u * ma.bh_lo) : u;
                double f;
                if constexpr (BW == 1) {
                    f = q;                                   // W pass skipped (size <= 1)
                } else {
                    const int e = e0 + j;
                    const double old = ring[c][j % BW];
                    ring[c][j % BW] = q;
                    tmp[c] = (e < BW) ? tmp[c] + q : tmp[c] + (q - old);
                    f = fma(tmp[c], ma.bw_hi, tmp[c] * ma.bw_lo);   // RN(tmp / bw)
                }
                fo[j] = f;
            }
        }
        __syncthreads();
        // ---- phase B: the pixels completed by this step
        if (tid < MR * XT) {
            const int r = tid / XT, j = tid - r * XT;
            const int k = t * XT + j - (BW - 1), y = y0 + r;
            if (k >= 0 && k < W && y < H) {
                const double *fp = fbuf + r * FP + j;
                double S0 = 0.0, Sc = 0.0, Ss = 0.0, Sc4 = 0.0, Ss4 = 0.0, Sff = 0.0;
                for (int i = 0; i < N; ++i) {
                    const double f = fp[i * (MR * FP)];
                    S0 = S0 + f;
                    Sc = Sc + f * bs.c2[i];
                    Ss = Ss + f * bs.s2[i];
                    if constexpr (NH == 2) {
                        Sc4 = Sc4 + f * bs.c4[i];
                        Ss4 = Ss4 + f * bs.s4[i];
                    }
                    Sff = fma(f, f, Sff);                    // screening only
                }
                const double I = fp[N * (MR * FP)];
                const long long p = (long long)y * W + k;
                HarmPix hp = harm_stage_a<NH>(a, p, S0, Sc, Ss, Sc4, Ss4, Sff, I);
                if (hp.need) {
                    ExactAcc ex;
                    ex.init();
                    for (int i = 0; i < N; ++i) {
                        double c4 = 0.0, s4 = 0.0;
                        if constexpr (NH == 2) { c4 = bs.c4[i]; s4 = bs.s4[i]; }
                        ex.template step<NH>(hp, fp[i * (MR * FP)], bs.c2[i], bs.s2[i], c4, s4);
                    }
                    hp.m = ex.decide(a);
                }
                const PixOut o = harm_stage_b<NH>(a, hs, do_hist, hp);
                store_pixel<NH>(a, out, p, o, I, a.flags & PB_FLAG_OUT_F64);
            }
        }
        __syncthreads();
    }
    if (do_hist) hist_flush(a, hs, out.hist);
}

size_t hist_bytes(const PixArgs &a) {
    size_t b = sizeof(double) * a.n_vars * (a.nbins + 1) + sizeof(unsigned) * a.n_groups * a.n_vars * a.nbins;
    return (b + 15) & ~(size_t)15;
}

template <typename UT, int NH, int BW>
cudaError_t launch_bw(const PixArgs &a, const typename BasisSel<NH>::type &bs, const pb_params *p, int n_stacks, cudaStream_t st) {
    constexpr int XT = StepWidth<BW>::XT;
    constexpr int FP = (XT % 2) ? XT : XT + 1;
    MarchArgs ma;
    const int bh = p->bin_h > 1 ? p->bin_h : 1;
    ma.bh = bh; ma.hh = bh / 2; ma.hw = BW / 2;
    ma.steps = (a.W + BW - 1 + XT - 1) / XT;
    ma.tile_rows = MR + bh - 1;
    int words = (XT + 1) / 2;
    if (words % 2 == 0) words += 1;
    ma.rp = 2 * words;
    const int threads = 32 * ((a.N + 1 + PL - 1) / PL);
    ma.npos = ma.tile_rows * XT;
    ma.groups = threads / ma.npos;
    if (ma.groups < 1) return cudaErrorInvalidConfiguration;
    if (ma.groups > a.N) ma.groups = a.N;
    if ((a.N + ma.groups - 1) / ma.groups > 24) return cudaErrorInvalidConfiguration;
    pb_ddrecip(bh, &ma.bh_hi, &ma.bh_lo);
    pb_ddrecip(BW, &ma.bw_hi, &ma.bw_lo);
    size_t off = hist_bytes(a);
    ma.off_raw = (unsigned)off; off += ((size_t)a.N * ma.tile_rows * ma.rp * 2 + 15) & ~(size_t)15;
    ma.off_T = (unsigned)off; off += ((size_t)ma.groups * ma.npos * 4 + 15) & ~(size_t)15;
    ma.off_f = (unsigned)off; off += (size_t)(a.N + 1) * MR * FP * 8;
    if (off > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = k_march<UT, NH, BW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off);
    if (e != cudaSuccess) return e;
    dim3 grid((a.H + MR - 1) / MR, n_stacks);
    kern<<<grid, threads, off, st>>>(a, bs, ma);
    return cudaGetLastError();
}

template <typename UT, int NH>
cudaError_t launch_nh(const PixArgs &a, const typename BasisSel<NH>::type &bs, const pb_params *p, int n_stacks, cudaStream_t st) {
    const int bw = p->bin_w > 1 ? p->bin_w : 1;
    switch (bw) {
    case 1: return launch_bw<UT, NH, 1>(a, bs, p, n_stacks, st);
    case 2: return launch_bw<UT, NH, 2>(a, bs, p, n_stacks, st);
    case 3: return launch_bw<UT, NH, 3>(a, bs, p, n_stacks, st);
    case 4: return launch_bw<UT, NH, 4>(a, bs, p, n_stacks, st);
    case 5: return launch_bw<UT, NH, 5>(a, bs, p, n_stacks, st);
    case 6: return launch_bw<UT, NH, 6>(a, bs, p, n_stacks, st);
    case 7: return launch_bw<UT, NH, 7>(a, bs, p, n_stacks, st);
    case 8: return launch_bw<UT, NH, 8>(a, bs, p, n_stacks, st);
    }
    return cudaErrorInvalidConfiguration;
}

}  // namespace

// The march kernel covers: harmonic modalities without sqrt (1PF, SRS, SHG, 2PF), integer stacks, integer-valued
// dark / dark+removed (so that the H pass sums integers), bin_w <= 8, N + 1 <= 32 chain warps. Everything else
// (CARS, 4POLAR, float stacks, fractional dark, wider windows, more angles) takes the unfused kernels.
bool pb_march_supported(const pb_params *p) {
    if (p->method == PB_CARS || p->method >= PB_4POLAR_2D) return false;
    if (p->dtype != PB_U8 && p->dtype != PB_U16) return false;
    const double sub = p->dark + p->removed_intensity;
    if (sub != floor(sub) || p->dark != floor(p->dark) || sub < 0 || sub > 65535.0 || p->dark < 0 || p->dark > 65535.0) return false;
    if (p->bin_w > 8 || p->bin_h > 32) return false;
    if (p->N + 1 > 32 || p->N < 1) return false;
    const int threads = 32 * ((p->N + 1 + PL - 1) / PL);
    const int bh = p->bin_h > 1 ? p->bin_h : 1, bw = p->bin_w > 1 ? p->bin_w : 1;
    const int xt = bw == 1 ? 8 : bw * ((8 + bw - 1) / bw);
    if (threads < (MR + bh - 1) * xt || threads < MR * xt) return false;      // one loader group, one pixel thread per step pixel
    return true;
}

cudaError_t pb_launch_march(const PixArgs &a, const Basis2 &bs, const pb_params *p, int n_stacks, cudaStream_t st, long *launches) {
    cudaError_t e;
    if (p->method == PB_1PF) {
        Basis1 b1;
        memcpy(b1.c2, bs.c2, sizeof(b1.c2)); memcpy(b1.s2, bs.s2, sizeof(b1.s2));
        memcpy(b1.kc2, bs.kc2, sizeof(b1.kc2)); memcpy(b1.ks2, bs.ks2, sizeof(b1.ks2));
        e = p->dtype == PB_U8 ? launch_nh<uint8_t, 1>(a, b1, p, n_stacks, st) : launch_nh<uint16_t, 1>(a, b1, p, n_stacks, st);
    } else {
        e = p->dtype == PB_U8 ? launch_nh<uint8_t, 2>(a, bs, p, n_stacks, st) : launch_nh<uint16_t, 2>(a, bs, p, n_stacks, st);
    }
    if (e == cudaSuccess && launches) ++*launches;
    return e;
}
