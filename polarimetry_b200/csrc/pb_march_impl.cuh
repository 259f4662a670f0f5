// pb_march_impl.cuh -- fused single-pass analysis WITH spatial binning: the stack is read once and the maps,
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
#pragma once
#include <cstdlib>
#include <type_traits>
#include "pb_pixel.cuh"

void pb_ddrecip(int n, double *hi, double *lo);   // pb_api.cu

namespace {

// MR = rows per CTA is a template parameter: 32 (N + 1 <= 32 lines per row), 16 (<= 64) or 8 (<= 128); a chain warp holds
// 32/MR groups of MR rows
constexpr int PL = 2;      // planes (lines) per chain thread
constexpr int QMAX = 2;    // tile positions per loader thread
#ifndef PB_MARCH_LA
#define PB_MARCH_LA 6
#endif
#ifndef PB_MARCH_MAXREG
#define PB_MARCH_MAXREG 64
#endif
constexpr int LA = PB_MARCH_LA;      // angle planes per loader batch
#ifndef PB_MARCH_PIPE
#define PB_MARCH_PIPE 0              // 1: line phase of step t+1 and pixel phase of step t share one barrier interval (double-buffered filtered tile)
#endif

struct MarchArgs {
    int bh, hh, hw;              // H-window size, its left extent bh/2, W-window left extent bw/2
    int steps;                   // number of XT-wide steps over the extended row (W + bw - 1)
    int tile_rows;               // MR + bh - 1
    int cw, lw;                  // compute (chain + pixel) warps, loader warps
    int isub, idark;             // dark + removed_intensity, dark (integers on this path)
    double bh_hi, bh_lo, bw_hi, bw_lo;   // double-double reciprocals of the bin sizes
    unsigned off_raw, off_T, off_f;      // byte offsets in dynamic shared memory (after the histogram area)
    unsigned raw_bytes, T_bytes;         // size of ONE buffer of the double-buffered raw / T tiles
    unsigned f_bytes;                    // size of ONE filtered tile (double-buffered when PB_MARCH_PIPE)
};

#ifndef PB_MARCH_XT_MIN
#define PB_MARCH_XT_MIN 8          // step width: the smallest multiple of the W window >= this
#endif
// Plane pitch of the raw tile in 32-bit words. With fewer than 32 rows per CTA a chain warp holds 32/MR line groups, PL planes
// apart: the pitch is padded so that the groups fall on disjoint banks (rows are RW words apart, RW odd; the groups must be
// MR*RW words apart modulo 32).
__host__ __device__ constexpr int march_plane_words(int rows, int RW, int MR) {
    int w = rows * RW;
    if (MR < 32) { while ((PL * w - MR * RW) % 32 != 0) ++w; }
    return w;
}

template <int BW> struct StepWidth { static constexpr int M = (PB_MARCH_XT_MIN + BW - 1) / BW; static constexpr int XT = BW * M; };
template <> struct StepWidth<1> { static constexpr int M = 8; static constexpr int XT = 8; };

// SQ: the H window equals the W window (bh == BW), known at compile time (the common square binning); else bh is a runtime value
template <typename UT, int NH, int BW, bool SQ, int MR>
__global__ void __maxnreg__(PB_MARCH_MAXREG) k_march(PixArgs a, typename BasisSel<NH>::type bs, MarchArgs ma) {
    constexpr int SUBS = 32 / MR;                       // line groups per chain warp
    constexpr int XT = StepWidth<BW>::XT;
    constexpr int FP = (XT % 2) ? XT : XT + 1;          // fbuf row pitch in doubles (odd: conflict-free chain stores)
    constexpr int RW = ((XT + 1) / 2) | 1;              // raw tile row pitch in 32-bit words (odd: conflict-free chain loads)
    constexpr int RP = 2 * RW;                          // ... in uint16
    extern __shared__ __align__(16) unsigned char smem[];
    const pb_outputs out = a.outs[blockIdx.y];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    HistSmem hs;
    hs.cnt = nullptr; hs.edges = nullptr;
    if (do_hist) hs = hist_setup(a, smem);

    const int N = a.N, H = a.H, W = a.W;
    const int y0 = blockIdx.x * MR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bh = SQ ? BW : ma.bh;
    const int TR = SQ ? MR + BW - 1 : ma.tile_rows;
    const int NPOS = TR * XT;
    constexpr int PPW_SQ = march_plane_words(MR + BW - 1, RW, MR);
    const int plane_pitch = SQ ? 2 * PPW_SQ : 2 * march_plane_words(ma.tile_rows, RW, MR);   // uint16 per plane of the raw tile
    const int steps = ma.steps;

    if (warp >= ma.cw) {
        // =========================== loader warps ===========================
        // raw stack -> shared memory, dark-subtracted and clamped (PP:2953), for step t+1 while the compute warps work on step t;
        // also the per-position total-intensity sums sum_i max(v - dark, 0) (PC:261) for the threshold image.
        const UT *stack = reinterpret_cast<const UT *>(a.stack) + (long long)blockIdx.y * a.stack_stride;
        asm volatile("" : "+l"(stack));
        const int ltid = tid - ma.cw * 32, LT = ma.lw * 32;
        const int isub = ma.isub, idark = ma.idark;
        const bool sep = isub != idark;
        const unsigned uP = (unsigned)a.P;
        unsigned rowoff[QMAX];
        int jj[QMAX], soff[QMAX];
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            const int pos = ltid + q * LT;
            const int r = pos / XT;
            jj[q] = pos - r * XT;
            soff[q] = r * RP + jj[q];
            rowoff[q] = pos < NPOS ? (unsigned)pb_reflect(y0 - ma.hh + r, H) * (unsigned)W : 0u;
            if (pos >= NPOS) jj[q] = -1;
        }
        for (int t = 0; t <= steps + PB_MARCH_PIPE; ++t) {
            // fill buffer t&1 with the data of step t (t == steps: nothing left, just take part in the barriers)
            if (t < steps) {
                uint16_t *rawb = reinterpret_cast<uint16_t *>(smem + ma.off_raw + (t & 1) * ma.raw_bytes);
                int *Tb = reinterpret_cast<int *>(smem + ma.off_T + (t & 1) * ma.T_bytes);
                unsigned off[QMAX];
                int tsum[QMAX];
#pragma unroll
                for (int q = 0; q < QMAX; ++q) {
                    off[q] = rowoff[q] + (unsigned)(jj[q] >= 0 ? pb_reflect(t * XT + jj[q] - ma.hw, W) : 0);
                    tsum[q] = 0;
                }
                // running plane pointers kept opaque so that the 64-bit stack base is not re-derived for every load:
                // per element one address IMAD.WIDE, the load, the fused subtract/clamp, the store and half an add
                const UT *bp = stack;
                uint16_t *dst = rawb;
                int i = 0;
                for (; i + LA <= N; i += LA) {            // LA planes x QMAX positions of loads in flight before any is consumed
                    asm volatile("" : "+l"(bp));
                    unsigned v[LA][QMAX];
                    const UT *b = bp;
#pragma unroll
                    for (int l = 0; l < LA; ++l) {
#pragma unroll
                        for (int q = 0; q < QMAX; ++q) v[l][q] = (unsigned)__ldg(b + off[q]);
                        b += uP;
                    }
#pragma unroll
                    for (int q = 0; q < QMAX; ++q) {
                        if (jj[q] >= 0) {
#pragma unroll
                            for (int l = 0; l < LA; ++l) {
                                const int f = max((int)v[l][q] - isub, 0);
                                dst[soff[q] + l * plane_pitch] = (uint16_t)f;
                                tsum[q] += sep ? max((int)v[l][q] - idark, 0) : f;
                            }
                        }
                    }
                    bp += (size_t)LA * uP;
                    dst += LA * plane_pitch;
                }
                for (; i < N; ++i) {
                    asm volatile("" : "+l"(bp));
#pragma unroll
                    for (int q = 0; q < QMAX; ++q) {
                        if (jj[q] >= 0) {
                            const int v = (int)__ldg(bp + off[q]);
                            const int f = max(v - isub, 0);
                            dst[soff[q]] = (uint16_t)f;
                            tsum[q] += sep ? max(v - idark, 0) : f;
                        }
                    }
                    bp += uP;
                    dst += plane_pitch;
                }
#pragma unroll
                for (int q = 0; q < QMAX; ++q)
                    if (jj[q] >= 0) Tb[ltid + q * LT] = tsum[q];
            }
            __syncthreads();                 // end-of-step barrier: buffer t&1 complete, step t-1 fully consumed
        }
    } else {
        // =========================== compute warps ===========================
        const int crow = lane % MR;                                            // chains (plane0 .. plane0+PL-1, row y0 + crow)
        const int plane0 = (warp * SUBS + lane / MR) * PL;
        // chain state: running sum and the last BW first-pass outputs. The ring starts at 0.0 so that the first BW steps
        // compute tmp + (q - 0.0) = tmp + q, i.e. scipy's left-to-right initial window sum, with the same instruction.
        double tmp[PL], ring[PL][BW];
#pragma unroll
        for (int c = 0; c < PL; ++c) {
            tmp[c] = 0.0;
#pragma unroll
            for (int s = 0; s < BW; ++s) ring[c][s] = 0.0;
        }
        const int cfo_off = (plane0 * MR + crow) * FP;
        const int pr = tid / XT, pj = tid - pr * XT;                           // pixel role
        const bool pix = tid < MR * XT && (y0 + pr) < H;
        const int pfp_off = pr * FP + pj;
        const bool f64 = a.flags & PB_FLAG_OUT_F64;
        const bool lean = store_is_lean<NH>(a, out);

        // ---- phase A: the (plane, row) lines advance XT extended positions: tile of step t -> filtered tile
        // one step of a line: q = RN(S1 / bh) [(double)S1 = (2^52 + S1) - 2^52 exactly; H pass skipped if bh <= 1, PP:2944],
        // tmp += q - ring (W pass, skipped if BW == 1; scipy's running sum, ring starts at 0.0), f = RN(tmp / bw)
        auto phase_a = [&](int t) {
            const uint16_t *rawb = reinterpret_cast<const uint16_t *>(smem + ma.off_raw + (t & 1) * ma.raw_bytes);
            const int *Tb = reinterpret_cast<const int *>(smem + ma.off_T + (t & 1) * ma.T_bytes);
            double *const cfo = reinterpret_cast<double *>(smem + ma.off_f + (PB_MARCH_PIPE ? (t & 1) * ma.f_bytes : 0u)) + cfo_off;
            if (plane0 + PL <= N) {
                // all PL lines of this thread are image planes. Blocks of JB positions, each in three stages -- all the
                // shared-memory loads and the independent first-pass quotients, then the (only) sequential part, the two
                // running sums, then the stores -- so that the independent fp64 chains overlap instead of running one
                // after the other (the compiler cannot move the tile loads across the filtered-tile stores by itself)
                constexpr int JB = (BW == 1) ? 4 : BW;
                const uint16_t *__restrict__ rp = rawb + plane0 * plane_pitch + crow * RP;
#pragma unroll
                for (int j0 = 0; j0 < XT; j0 += JB) {
                    double q[JB][PL];
#pragma unroll
                    for (int jj = 0; jj < JB; ++jj) {
#pragma unroll
                        for (int c = 0; c < PL; ++c) {
                            if (j0 + jj < XT) {
                                int S1 = 0;
                                if constexpr (SQ) {
#pragma unroll
                                    for (int dy = 0; dy < BW; ++dy) S1 += (int)rp[c * plane_pitch + dy * RP + j0 + jj];
                                } else {
                                    for (int dy = 0; dy < bh; ++dy) S1 += (int)rp[c * plane_pitch + dy * RP + j0 + jj];
                                }
                                const double u = __hiloint2double(0x43300000, S1) - 4503599627370496.0;
                                q[jj][c] = (bh > 1) ? fma(u, ma.bh_hi, u * ma.bh_lo) : u;
                            }
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < JB; ++jj) {
#pragma unroll
                        for (int c = 0; c < PL; ++c) {
                            if (j0 + jj < XT) {
                                if constexpr (BW > 1) {
                                    const double d = q[jj][c] - ring[c][(j0 + jj) % BW];
                                    ring[c][(j0 + jj) % BW] = q[jj][c];
                                    tmp[c] = tmp[c] + d;
                                    q[jj][c] = fma(tmp[c], ma.bw_hi, tmp[c] * ma.bw_lo);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < JB; ++jj) {
#pragma unroll
                        for (int c = 0; c < PL; ++c)
                            if (j0 + jj < XT) cfo[c * (MR * FP) + j0 + jj] = q[jj][c];
                    }
                }
            } else if (plane0 <= N) {
                // the last line group: one image plane (odd N) and / or the total-intensity line (PC:260-264: the same two
                // passes on the integer sums over the angles). Same staged form, one line at a time -- all the other warps
                // wait for this one at the end of the phase.
#pragma unroll
                for (int c = 0; c < PL; ++c) {
                    const int plane = plane0 + c;
                    if (plane > N) break;
                    const bool is_t = plane == N;
                    const uint16_t *__restrict__ rp = rawb + plane * plane_pitch + crow * RP;
                    const int *__restrict__ tp = Tb + crow * XT;
                    double *fo = cfo + c * (MR * FP);
                    constexpr int JB = (BW == 1) ? 4 : BW;
#pragma unroll
                    for (int j0 = 0; j0 < XT; j0 += JB) {
                        double q[JB];
#pragma unroll
                        for (int jj = 0; jj < JB; ++jj) {
                            if (j0 + jj < XT) {
                                int S1 = 0;
                                if constexpr (SQ) {
#pragma unroll
                                    for (int dy = 0; dy < BW; ++dy) S1 += is_t ? tp[dy * XT + j0 + jj] : (int)rp[dy * RP + j0 + jj];
                                } else {
                                    for (int dy = 0; dy < bh; ++dy) S1 += is_t ? tp[dy * XT + j0 + jj] : (int)rp[dy * RP + j0 + jj];
                                }
                                const double u = __hiloint2double(0x43300000, S1) - 4503599627370496.0;
                                q[jj] = (bh > 1) ? fma(u, ma.bh_hi, u * ma.bh_lo) : u;
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < JB; ++jj) {
                            if (j0 + jj < XT) {
                                if constexpr (BW > 1) {
                                    const double d = q[jj] - ring[c][(j0 + jj) % BW];
                                    ring[c][(j0 + jj) % BW] = q[jj];
                                    tmp[c] = tmp[c] + d;
                                    q[jj] = fma(tmp[c], ma.bw_hi, tmp[c] * ma.bw_lo);
                                }
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < JB; ++jj)
                            if (j0 + jj < XT) fo[j0 + jj] = q[jj];
                    }
                }
            }
        };
        // ---- phase B: the pixels completed by step t
        auto phase_b = [&](int t) {
            const double *const pfp = reinterpret_cast<const double *>(smem + ma.off_f + (PB_MARCH_PIPE ? (t & 1) * ma.f_bytes : 0u)) + pfp_off;
            const int k = t * XT + pj - (BW - 1);
            if (pix && k >= 0 && k < W) {
                const long long p = (long long)(y0 + pr) * W + k;
                const double I = pfp[N * (MR * FP)];             // plane N: the binned total intensity
                // the threshold / ROI test needs only I: pixels that fail it skip the harmonic sums altogether
                // (their maps are NaN whatever the sums are, PP:3017-3025 / PP:3052)
                bool cand;
                if (a.simple) cand = I >= a.roi_ilow[0];
                else if (a.mask_in) cand = __ldg(a.mask_in + p) != 0;
                else cand = roi_pass(a, a.roi_bits ? __ldg(a.roi_bits + p) : 1u, I);
                PixOut o;
                o.v0 = o.v1 = o.v2 = o.v3 = o.a0m = pb_nan();
                o.m = false;
                if (!cand && lean) {
                    store_invalid_lean<NH>(a, out, p);
                    return;
                }
                if (cand) {
                    double S0 = 0.0, Sc = 0.0, Ss = 0.0, Sc4 = 0.0, Ss4 = 0.0, Sff = 0.0;
                    const double *fp = pfp;
#pragma unroll 6
                    for (int i = 0; i < N; ++i) {
                        const double f = *fp;
                        fp += MR * FP;
                        S0 = S0 + f;
                        Sc = Sc + f * bs.c2[i];
                        Ss = Ss + f * bs.s2[i];
                        if constexpr (NH == 2) {
                            Sc4 = Sc4 + f * bs.c4[i];
                            Ss4 = Ss4 + f * bs.s4[i];
                        }
                        Sff = fma(f, f, Sff);                    // screening only
                    }
                    // the threshold / ROI test is `cand` (above): stage A proper
                    const uint32_t bits = (a.simple || !a.roi_bits) ? 1u : __ldg(a.roi_bits + p);
                    HarmPix hp = harm_stage_a_core<NH>(a, bits, S0, Sc, Ss, Sc4, Ss4, Sff);
                    if (hp.need) {
                        ExactAcc ex;
                        ex.init();
                        for (int i = 0; i < N; ++i) {
                            double c4 = 0.0, s4 = 0.0;
                            if constexpr (NH == 2) { c4 = bs.c4[i]; s4 = bs.s4[i]; }
                            ex.template step<NH>(hp, pfp[i * (MR * FP)], bs.c2[i], bs.s2[i], c4, s4);
                        }
                        hp.m = ex.decide(a);
                    }
                    o = harm_stage_b<NH>(a, hs, do_hist, hp);
                }
                store_pixel<NH>(a, out, p, o, I, f64, lean);
            }
        };

        __syncthreads();                      // tile of step 0 in place
#if PB_MARCH_PIPE
        // software pipeline: between two barriers a compute warp filters the lines of step t+1 (into the other filtered tile) and
        // then finishes the pixels of step t, so that line warps and pixel warps of one CTA overlap; the loader runs two steps ahead
        // (one instance of each phase in the code: the kernel has to stay inside the instruction cache)
        for (int t = -1; t < steps; ++t) {
            if (t + 1 < steps) phase_a(t + 1);
            if (t >= 0) phase_b(t);
            __syncthreads();                  // filtered tile of step t+1 complete, raw tile of step t+2 in place
        }
#else
        for (int t = 0; t < steps; ++t) {
            phase_a(t);
            asm volatile("bar.sync 1, %0;" ::"r"(ma.cw * 32) : "memory");   // compute warps only: filtered tile of step t complete
            phase_b(t);
            __syncthreads();                  // end-of-step barrier (all warps): filtered tile free again, tile of step t+1 in place
        }
#endif
    }
    if (do_hist) hist_flush(a, hs, out.hist);
}

size_t hist_bytes(const PixArgs &a) {
    size_t b = sizeof(double) * a.n_vars * (a.nbins + 1) + sizeof(unsigned) * a.n_groups * a.n_vars * a.nbins;
    return (b + 15) & ~(size_t)15;
}

template <typename UT, int NH, int BW, int MR>
cudaError_t launch_bw(const PixArgs &a, const typename BasisSel<NH>::type &bs, const pb_params *p, int n_stacks, cudaStream_t st) {
    constexpr int XT = StepWidth<BW>::XT;
    constexpr int FP = (XT % 2) ? XT : XT + 1;
    constexpr int RP = 2 * (((XT + 1) / 2) | 1);
    constexpr int SUBS = 32 / MR;
    MarchArgs ma;
    const int bh = p->bin_h > 1 ? p->bin_h : 1;
    ma.bh = bh; ma.hh = bh / 2; ma.hw = BW / 2;
    ma.steps = (a.W + BW - 1 + XT - 1) / XT;
    ma.tile_rows = MR + bh - 1;
    ma.isub = (int)a.sub; ma.idark = (int)a.dark;
    const int npos = ma.tile_rows * XT;
    const int groups = (a.N + 1 + PL - 1) / PL;             // PL lines per chain thread
    ma.cw = (groups + SUBS - 1) / SUBS;
    if (ma.cw * 32 < MR * XT) ma.cw = (MR * XT + 31) / 32;
    ma.lw = (npos + 32 * QMAX - 1) / (32 * QMAX);
    if (ma.lw < 2 && npos > 32 * 3) ma.lw = 2;
    const int threads = 32 * (ma.cw + ma.lw);
    if (threads > 1024 || a.P >= 2147483647LL) return cudaErrorInvalidConfiguration;
    pb_ddrecip(bh, &ma.bh_hi, &ma.bh_lo);
    pb_ddrecip(BW, &ma.bw_hi, &ma.bw_lo);
    size_t off = hist_bytes(a);
    ma.raw_bytes = (unsigned)(((size_t)a.N * march_plane_words(ma.tile_rows, RP / 2, MR) * 4 + 15) & ~(size_t)15);
    ma.T_bytes = (unsigned)(((size_t)npos * 4 + 15) & ~(size_t)15);
    ma.off_raw = (unsigned)off; off += 2 * (size_t)ma.raw_bytes;
    ma.off_T = (unsigned)off; off += 2 * (size_t)ma.T_bytes;
    ma.f_bytes = (unsigned)((size_t)(a.N + 1) * MR * FP * 8);
    ma.off_f = (unsigned)off; off += (size_t)ma.f_bytes * (PB_MARCH_PIPE ? 2 : 1);
    if (off > 227 * 1024) return cudaErrorInvalidConfiguration;
    dim3 grid((a.H + MR - 1) / MR, n_stacks);
    cudaError_t e;
    if (bh == BW && BW > 1) {
        auto kern = k_march<UT, NH, BW, true, MR>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off)) != cudaSuccess) return e;
        kern<<<grid, threads, off, st>>>(a, bs, ma);
    } else {
        if constexpr (MR == 32) {
            auto kern = k_march<UT, NH, BW, false, MR>;
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off)) != cudaSuccess) return e;
            kern<<<grid, threads, off, st>>>(a, bs, ma);
        } else {
            return cudaErrorInvalidConfiguration;       // rectangular windows with more than 31 angles: unfused kernels
        }
    }
    return cudaGetLastError();
}

template <typename UT, int NH, int MR>
cudaError_t launch_mr(const PixArgs &a, const typename BasisSel<NH>::type &bs, const pb_params *p, int n_stacks, cudaStream_t st) {
    const int bw = p->bin_w > 1 ? p->bin_w : 1;
    switch (bw) {
    case 1: if constexpr (MR == 32) return launch_bw<UT, NH, 1, MR>(a, bs, p, n_stacks, st); else return cudaErrorInvalidConfiguration;
    case 2: return launch_bw<UT, NH, 2, MR>(a, bs, p, n_stacks, st);
    case 3: return launch_bw<UT, NH, 3, MR>(a, bs, p, n_stacks, st);
    case 4: return launch_bw<UT, NH, 4, MR>(a, bs, p, n_stacks, st);
    case 5: return launch_bw<UT, NH, 5, MR>(a, bs, p, n_stacks, st);
    case 6: return launch_bw<UT, NH, 6, MR>(a, bs, p, n_stacks, st);
    case 7: return launch_bw<UT, NH, 7, MR>(a, bs, p, n_stacks, st);
    case 8: return launch_bw<UT, NH, 8, MR>(a, bs, p, n_stacks, st);
    }
    return cudaErrorInvalidConfiguration;
}

// rows per CTA from the number of lines per row (N planes + the intensity plane); PB_MARCH_ROWS=8|16|32 is a tuning knob
int march_rows_forced() {
    static const int forced = [] { const char *e = getenv("PB_MARCH_ROWS"); return e ? atoi(e) : 0; }();
    return forced;
}
int march_rows(int N) {
    const int need = N + 1 <= 32 ? 32 : (N + 1 <= 64 ? 16 : 8);
    const int forced = march_rows_forced();
    return (forced == 16 || forced == 8) && forced < need ? forced : need;
}

template <typename UT, int NH>
cudaError_t launch_nh(const PixArgs &a, const typename BasisSel<NH>::type &bs, const pb_params *p, int n_stacks, cudaStream_t st) {
    int mr = march_rows(a.N);
    if constexpr (std::is_same<UT, uint16_t>::value) {
        // uint16 stacks with square windows also have the 16- and 8-row variants. 16 rows per CTA (4 CTAs of 8 warps per SM
        // instead of 2 of 15) measured 2-3 % faster than 32 on full batches; with few stacks (the single-stack drop-in call,
        // the host-buffer pipeline) go down to 8 rows until the grid covers the machine twice (2048 rows = 64 CTAs of 32
        // rows for 148 SMs).
        const bool small_ok = p->bin_h == p->bin_w && p->bin_w >= 2 && march_rows_forced() == 0;
        if (small_ok && mr == 32) mr = 16;
        while (small_ok && mr > 8 && (long long)n_stacks * ((a.H + mr - 1) / mr) < 2 * 148) mr /= 2;
    }
    if (mr == 32) return launch_mr<UT, NH, 32>(a, bs, p, n_stacks, st);
    if constexpr (std::is_same<UT, uint16_t>::value) {      // the 16- and 8-row variants are built for uint16 stacks only
        if (mr == 16) return launch_mr<UT, NH, 16>(a, bs, p, n_stacks, st);
        return launch_mr<UT, NH, 8>(a, bs, p, n_stacks, st);
    }
    return cudaErrorInvalidConfiguration;
}

}  // namespace
