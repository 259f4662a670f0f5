// pb_march_tma.cuh -- fused single-pass analysis WITH spatial binning, Blackwell data path: the raw uint16 tiles are brought
// into shared memory by the TMA (cp.async.bulk.tensor, 4-D boxes over [stack][plane][row][column], 32-byte swizzle) completing
// on mbarriers, two stages deep; no thread ever issues a global load for the stack (edge tiles excepted, see fix_tile).
//
// Replaces, in one kernel: Stack.get_intensity (pypolar_classes.py:260-264), analyze_stack's pre-processing (PyPOLAR.py:2953-2957),
// Polarimetry.binning (PP:2942-2946), compute_fields (PP:2981-3078), the mask logic (PP:2891/2967) and the histograms (PC:340-369).
// The arithmetic is that of pb_march_impl.cuh (scipy's running-sum box filter replayed bit for bit); what is new is
// the data movement and the work decomposition:
//
//   one CTA = MR rows of one stack, marching over W in steps of XC = 16 columns (one 32-byte TMA row);
//   tile u (u = 0 .. U) = image columns [16(u-1), 16u), rows [y0 - bh/2, y0 + MR + bh - 1 - bh/2), all N planes, as stored;
//       out-of-image elements arrive as zeros (TMA fill) and the few that the 'reflect' boundary needs are patched from global
//       memory (fix_tile; first / last tile of a row block, first / last row block of a stack);
//   phase A(u)  one thread per (plane, row) line: H window sum straight from the packed raw tile with 128-bit shared loads
//               (sum of max(v, sub) per halfword; the - bh*sub is folded into the 2^52 conversion constant, PP:2953),
//               q = RN(S/bh), tmp += q - ring, f = RN(tmp/bw) -> fp64 tile (128-bit stores, conflict-free pitch);
//               line N is the total-intensity image (PC:260-264) fed from the T tile;
//   phase B(u)  pixel warps: one thread per pixel, sequential harmonic sums over the planes (PP:2987-2994) + the shared epilogue;
//               T warps (concurrently): sum over the planes of max(v, dark) of tile u+1 -> T tile, after patching that tile.
//   Two block barriers per step; the TMA for tile u+2 is issued right after phase A(u) released its stage.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include "pb_pixel4.cuh"

void pb_ddrecip(int n, double *hi, double *lo);   // pb_api.cu

namespace tma {

constexpr int XC = 16;                 // columns per step: 32 bytes of uint16 = the inner extent of a 32-byte-swizzled TMA box
constexpr int FPB = XC * 8 + 16;       // filtered-tile row pitch in bytes (odd multiple of 16: conflict-free 128-bit chain stores)
constexpr int TPB = XC * 4 + 16;       // T-tile row pitch in bytes
#ifndef PB_TMA_MAXREG
#define PB_TMA_MAXREG 64
#endif

struct TmaArgs {
    int U;                       // last tile index: tiles u = 0 .. U
    int isub, idark;             // dark + removed_intensity, dark (integers on this path)
    int threads, pix_warps, t_warps;
    double b_hi, b_lo;           // double-double reciprocal of the (square) bin size
    double magic, magicT;        // 2^52 + bw*isub, 2^52 + bw*N*idark: conversion constants that also remove the clamp offsets
    unsigned off_bar, off_raw, raw_bytes, off_T, off_f;   // byte offsets in dynamic shared memory (off_raw relative to the 1024-aligned base)
    // redo mode (second pass after k_strip, pb_strip.cuh): one flag per CTA (16-row block of a stack); unflagged CTAs exit at once, the
    // others evaluate ONLY the pixels the certified screening leaves undecided (hp.need) -- k_strip stored all the others
    const unsigned char *redo;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap *map, unsigned bar, int x, int y, int z, int w) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(w), "r"(bar)
                 : "memory");
}

// byte offset of 16-byte chunk `half` of logical tile row L inside a stage written with CU_TENSOR_MAP_SWIZZLE_32B
// (address bit 4 ^= address bit 7; rows are 32 bytes, the stage base is 1024-byte aligned)
__device__ __forceinline__ unsigned swz(int L, int half) { return (unsigned)(L * 32 + ((half ^ ((L >> 2) & 1)) << 4)); }

// NT: number of planes known at compile time (0 = run-time N): the harmonic sums are then fully unrolled and the basis constants
// become immediate constant-bank operands (no loop control, no uniform loads)
template <int NH, int BW, int MR, int NT>
__global__ void __maxnreg__(PB_TMA_MAXREG) k_march_tma(const __grid_constant__ CUtensorMap tmap, PixArgs a, typename BasisSel<NH>::type bs, TmaArgs ta) {
    constexpr int TR = MR + BW - 1;                     // tile rows
    constexpr int HH = BW / 2;                          // window extent before the centre (rows above / columns left)
    constexpr int RX = BW - 1 - HH;                     // ... after the centre; outputs lag the consumed column by RX
    constexpr int FPD = FPB / 8;                        // filtered-tile row pitch in doubles
    constexpr int PLD = MR * FPD;                       // ... plane pitch in doubles
    extern __shared__ __align__(16) unsigned char smem[];
    const bool redo = ta.redo != nullptr;
    if (redo && !ta.redo[(size_t)blockIdx.y * gridDim.x + blockIdx.x]) return;
    const pb_outputs out = a.outs[blockIdx.y];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    const int N = NT ? NT : a.N, H = a.H, W = a.W;
    const int y0 = blockIdx.x * MR;
    const int tid = threadIdx.x, warp = tid >> 5;
    const unsigned sbase = smem_u32(smem);
    const unsigned abase = (sbase + 1023u) & ~1023u;                 // 1024-aligned base of the raw stages
    unsigned char *const raw0 = smem + (abase - sbase) + ta.off_raw;
    const unsigned bar0 = sbase + ta.off_bar;                          // two mbarriers: stage s completes on bar0 + 8 s
    const unsigned stage_bytes = ta.raw_bytes;
    const int U = ta.U;
    const int ytop = y0 - HH;

    auto issue = [&](int u) {       // one elected thread: arm the stage's barrier and start the box copy of tile u
        const unsigned bar = bar0 + 8u * (u & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses of this stage (reads, edge patches) before the async write
        mbar_expect_tx(bar, (unsigned)(N * TR * 32));
        tma_load_4d(abase + ta.off_raw + (u & 1) * stage_bytes, &tmap, bar, 16 * (u - 1), ytop, 0, (int)blockIdx.y);
    };
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    HistSmem hs;
    hs.cnt = nullptr; hs.edges = nullptr;
    if (do_hist) hs = hist_setup(a, smem);          // ends with a block barrier
    else __syncthreads();
    if (tid == 0) {
        issue(0);
        if (U >= 1) issue(1);
    }

    // ---- patch of tile u: the elements the 'reflect' boundary needs that lie outside the image (TMA delivered zeros)
    const uint16_t *const gstack = reinterpret_cast<const uint16_t *>(a.stack) + (long long)blockIdx.y * a.stack_stride;
    const int rtop = min(max(-ytop, 0), TR), rbot = min(max(H - ytop, 0), TR);      // tile rows [rtop, rbot) lie inside the image
    auto tile_is_edge = [&](int u) { return rtop > 0 || rbot < TR || u == 0 || 16 * u > W; };
    auto fix_tile = [&](int u, int ftid, int nfix) {
        unsigned char *raw = raw0 + (u & 1) * stage_bytes;
        const int c0 = 16 * (u - 1);
        // out-of-image columns of this tile that are needed: image columns [-HH, 0) and [W, W + RX)
        const int cl0 = min(max(-HH - c0, 0), 16), cl1 = min(max(0 - c0, 0), 16);
        const int cr0 = min(max(W - c0, 0), 16), cr1 = min(max(W + RX - c0, 0), 16);
        const int nl = cl1 - cl0, ncol = nl + (cr1 - cr0);
        const int nrow = rtop + (TR - rbot), nin = rbot - rtop;
        const int per_plane = nrow * 16 + nin * ncol;
        const int total = N * per_plane;
        for (int idx = ftid; idx < total; idx += nfix) {
            const int p = idx / per_plane;
            int rem = idx - p * per_plane, rr, cc;
            if (rem < nrow * 16) {
                const int ri = rem >> 4;
                cc = rem & 15;
                rr = ri < rtop ? ri : rbot + (ri - rtop);
            } else {
                rem -= nrow * 16;
                const int ri = rem / ncol, ci = rem - ri * ncol;
                rr = rtop + ri;
                cc = ci < nl ? cl0 + ci : cr0 + (ci - nl);
            }
            const int c = c0 + cc;
            if (c < -HH || c >= W + RX) continue;                         // never consumed by a valid output: stays zero
            const int x = pb_reflect(c, W), y = pb_reflect(ytop + rr, H);
            const uint16_t v = __ldg(gstack + (long long)p * a.P + (long long)y * W + x);
            const int L = p * TR + rr;
            *reinterpret_cast<uint16_t *>(raw + swz(L, cc >> 3) + (cc & 7) * 2) = v;
        }
        // these generic-proxy stores are later overwritten by the TMA (async proxy) when the stage is refilled
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    };
    // ---- T tile of tile u: per tile position the sum over the planes of max(v, dark) (PC:261 up to the constant N*dark)
    unsigned char *const Tt = smem + ta.off_T;
    // The swizzle bit of tile row L = p*TR + rr is periodic in the plane p with period 8 (L grows by 8*TR = a multiple of 8 rows),
    // so a thread needs 8 byte offsets for all its planes; the loads then take immediate plane offsets.
    auto t_phase = [&](int u, int item) {
        if (item >= TR * 2) return;
        const unsigned char *raw = raw0 + (u & 1) * stage_bytes;
        const int rr = item >> 1, half = item & 1;
        const unsigned dark2 = (unsigned)ta.idark * 0x10001u;
        unsigned offk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) offk[k] = swz(k * TR + rr, half);
        unsigned acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0u;
        auto add_plane = [&](const uint4 w) {
            const unsigned m0 = __vmaxu2(w.x, dark2), m1 = __vmaxu2(w.y, dark2), m2 = __vmaxu2(w.z, dark2), m3 = __vmaxu2(w.w, dark2);
            // __dp2a_lo(m, b, c) = c + m.lo16 * b.byte0 + m.hi16 * b.byte1: b = 0x0001 adds the low halfword, b = 0x0100 the high one
            acc[0] = __dp2a_lo(m0, 0x0001u, acc[0]); acc[1] = __dp2a_lo(m0, 0x0100u, acc[1]);
            acc[2] = __dp2a_lo(m1, 0x0001u, acc[2]); acc[3] = __dp2a_lo(m1, 0x0100u, acc[3]);
            acc[4] = __dp2a_lo(m2, 0x0001u, acc[4]); acc[5] = __dp2a_lo(m2, 0x0100u, acc[5]);
            acc[6] = __dp2a_lo(m3, 0x0001u, acc[6]); acc[7] = __dp2a_lo(m3, 0x0100u, acc[7]);
        };
        int p = 0;
        for (; p + 8 <= N; p += 8) {
            const unsigned char *base = raw + p * (TR * 32);
#pragma unroll
            for (int k = 0; k < 8; ++k) add_plane(*reinterpret_cast<const uint4 *>(base + offk[k]));
        }
        {
            const unsigned char *base = raw + p * (TR * 32);
#pragma unroll
            for (int k = 0; k < 7; ++k)
                if (p + k < N) add_plane(*reinterpret_cast<const uint4 *>(base + offk[k]));
        }
        uint4 *dst = reinterpret_cast<uint4 *>(Tt + rr * TPB + half * 32);
        dst[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
    };

    // ---- prologue: tile 0 in place, patched, its T tile computed (the same work phase B does for tile u+1 in the steady state)
    mbar_wait(bar0, 0);
    fix_tile(0, tid, blockDim.x);                   // tile 0 always holds the left boundary columns
    __syncthreads();
    t_phase(0, tid);
    __syncthreads();

    // ---- roles
    const int lines = (N + 1) * MR;
    const bool chain = tid < lines;
    const int cpl = tid / MR, crow = tid - cpl * MR;              // chain: plane (N = the intensity line) and row
    double tmp = 0.0, ring[BW];
#pragma unroll
    for (int s = 0; s < BW; ++s) ring[s] = 0.0;
    double *const fbuf = reinterpret_cast<double *>(smem + ta.off_f);
    double *const cfo = fbuf + (cpl * MR + crow) * FPD;
    const bool pixw = warp < ta.pix_warps;
    const int pr = tid / XC, pj = tid - pr * XC;                   // pixel role
    const bool prow_ok = pixw && (y0 + pr) < H;
    const double *const pfp = fbuf + pr * FPD + pj;
    const bool f64 = a.flags & PB_FLAG_OUT_F64;
    const bool lean = NH != 0 && store_is_lean<(NH ? NH : 1)>(a, out);
    const unsigned sub2 = (unsigned)ta.isub * 0x10001u;
    const int ft = tid - ta.pix_warps * 32, nft = ta.t_warps * 32;   // T-warp thread index / count

    // one extended position of a line: first-pass sum S (with its clamp offset still in) -> q = RN(S/bh) -> running sum -> f
    auto advance = [&](unsigned S, double magic, int j) -> double {
        const double u = __hiloint2double(0x43300000, (int)S) - magic;
        const double q = fma(u, ta.b_hi, u * ta.b_lo);
        const double d = q - ring[j % BW];
        ring[j % BW] = q;
        tmp = tmp + d;
        return fma(tmp, ta.b_hi, tmp * ta.b_lo);
    };

    for (int u = 0; u <= U; ++u) {
        const unsigned char *raw = raw0 + (u & 1) * stage_bytes;
        mbar_wait(bar0 + 8u * (u & 1), (unsigned)((u >> 1) & 1));      // every thread observes the completion of tile u (already complete)
        // =========================== phase A(u) ===========================
        if (chain) {
            if (cpl < N) {
                const int L0 = cpl * TR + crow;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint4 w[BW];
#pragma unroll
                    for (int dy = 0; dy < BW; ++dy) w[dy] = *reinterpret_cast<const uint4 *>(raw + swz(L0 + dy, half));
                    double f[8];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        unsigned lo = 0u, hi = 0u;
#pragma unroll
                        for (int dy = 0; dy < BW; ++dy) {
                            const unsigned word = k == 0 ? w[dy].x : (k == 1 ? w[dy].y : (k == 2 ? w[dy].z : w[dy].w));
                            const unsigned m = __vmaxu2(word, sub2);
                            lo = __dp2a_lo(m, 0x0001u, lo);       // lo += m & 0xffff, hi += m >> 16: one instruction each
                            hi = __dp2a_lo(m, 0x0100u, hi);
                        }
                        f[2 * k] = advance(lo, ta.magic, half * 8 + 2 * k);
                        f[2 * k + 1] = advance(hi, ta.magic, half * 8 + 2 * k + 1);
                    }
                    double2 *dst = reinterpret_cast<double2 *>(cfo + half * 8);
#pragma unroll
                    for (int k = 0; k < 4; ++k) dst[k] = make_double2(f[2 * k], f[2 * k + 1]);
                }
            } else {
                // the total-intensity line (PC:260-264): the same two passes on the integer sums over the angles
#pragma unroll
                for (int quad = 0; quad < 4; ++quad) {
                    unsigned S[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int dy = 0; dy < BW; ++dy) {
                        const uint4 t = *reinterpret_cast<const uint4 *>(Tt + (crow + dy) * TPB + quad * 16);
                        S[0] += t.x; S[1] += t.y; S[2] += t.z; S[3] += t.w;
                    }
                    double f[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) f[k] = advance(S[k], ta.magicT, quad * 4 + k);
                    double2 *dst = reinterpret_cast<double2 *>(cfo + quad * 4);
                    dst[0] = make_double2(f[0], f[1]);
                    dst[1] = make_double2(f[2], f[3]);
                }
            }
            // the ring is indexed by (column % BW) inside a step: rotate it so that slot 0 is the oldest entry again
            if constexpr (XC % BW != 0) {
                double r2[BW];
#pragma unroll
                for (int s = 0; s < BW; ++s) r2[s] = ring[(s + XC) % BW];
#pragma unroll
                for (int s = 0; s < BW; ++s) ring[s] = r2[s];
            }
        }
        __syncthreads();          // filtered tile of step u complete; raw stage (u & 1) and the T tile are free
        if (tid == 0 && u + 2 <= U) issue(u + 2);
        // =========================== phase B(u) ===========================
        if (pixw) {
            const int k = 16 * (u - 1) + pj - RX;
            if (prow_ok && k >= 0 && k < W) {
                const long long p = (long long)(y0 + pr) * W + k;
                const double I = pfp[N * PLD];                   // line N: the binned total intensity
                if constexpr (NH == 0) {
                    // 4POLAR (N = 4): the binned planes in the order of PP:2957 (acquisition planes 1 and 2 swapped) -> PP:3001-3014, 3041-3051
                    const double f4[4] = {pfp[0], pfp[2 * PLD], pfp[PLD], pfp[3 * PLD]};
                    p4_pixel<false>(a, hs, do_hist, out, p, f4, I);
                } else {
                bool cand;
                if (a.simple) cand = I >= a.roi_ilow[0];
                else if (a.mask_in) cand = __ldg(a.mask_in + p) != 0;
                else cand = roi_pass(a, a.roi_bits ? __ldg(a.roi_bits + p) : 1u, I);
                if (redo && !cand) {
                    // stored by k_strip
                } else if (!cand && lean) {
                    store_invalid_lean<NH>(a, out, p);
                } else {
                    PixOut o;
                    o.v0 = o.v1 = o.v2 = o.v3 = o.a0m = pb_nan();
                    o.m = false;
                    bool emit = true;
                    if (cand) {
                        double S0 = 0.0, Sc = 0.0, Ss = 0.0, Sc4 = 0.0, Ss4 = 0.0, Sff = 0.0;
                        const double *fp = pfp;
#pragma unroll(NT ? NT : 6)
                        for (int i = 0; i < N; ++i) {
                            const double f = *fp;
                            fp += PLD;
                            S0 = S0 + f;
                            Sc = Sc + f * bs.c2[i];
                            Ss = Ss + f * bs.s2[i];
                            if constexpr (NH == 2) {
                                Sc4 = Sc4 + f * bs.c4[i];
                                Ss4 = Ss4 + f * bs.s4[i];
                            }
                            Sff = fma(f, f, Sff);                    // screening only
                        }
                        const uint32_t bits = (a.simple || !a.roi_bits) ? 1u : __ldg(a.roi_bits + p);
                        HarmPix hp = harm_stage_a_core<NH>(a, bits, S0, Sc, Ss, Sc4, Ss4, Sff);
                        if (hp.need) {
                            ExactAcc ex;
                            ex.init();
                            for (int i = 0; i < N; ++i) {
                                double c4 = 0.0, s4 = 0.0;
                                if constexpr (NH == 2) { c4 = bs.c4[i]; s4 = bs.s4[i]; }
                                ex.template step<NH>(hp, pfp[i * PLD], bs.c2[i], bs.s2[i], c4, s4);
                            }
                            hp.m = ex.decide(a);
                        } else if (redo) {
                            emit = false;
                        }
                        if (emit) o = harm_stage_b<NH>(a, hs, do_hist, hp);
                    }
                    if (emit) store_pixel<NH>(a, out, p, o, I, f64, lean);
                }
                }
            }
        } else if (u + 1 <= U) {
            // T warps: tile u+1 (issued one step ago) -> patched -> its T tile, ready for phase A(u+1) at the end-of-step barrier
            mbar_wait(bar0 + 8u * ((u + 1) & 1), (unsigned)(((u + 1) >> 1) & 1));
            if (tile_is_edge(u + 1)) {
                fix_tile(u + 1, ft, nft);
                asm volatile("bar.sync 1, %0;" ::"r"(nft) : "memory");
            }
            t_phase(u + 1, ft);
        }
        __syncthreads();          // filtered tile free again; T tile and patches of tile u+1 complete
    }
    if (do_hist) hist_flush(a, hs, out.hist);
}

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

inline size_t hist_bytes_tma(const PixArgs &a) {
    size_t b = sizeof(double) * a.n_vars * (a.nbins + 1) + sizeof(unsigned) * a.n_groups * a.n_vars * a.nbins;
    return (b + 15) & ~(size_t)15;
}

// shared-memory plan of one CTA; returns the total dynamic size (0 if it does not fit)
template <int BW, int MR>
size_t plan(const PixArgs &a, TmaArgs &ta) {
    constexpr int TR = MR + BW - 1;
    size_t off = hist_bytes_tma(a);
    ta.off_bar = (unsigned)off; off += 16;
    ta.off_T = (unsigned)off; off += (size_t)TR * TPB;
    off = (off + 15) & ~(size_t)15;
    ta.off_f = (unsigned)off; off += (size_t)(a.N + 1) * MR * FPB;
    // raw stages behind a 1024-byte alignment slack (the dynamic shared memory base is only 16-byte aligned by contract)
    off = (off + 1023) & ~(size_t)1023;
    ta.off_raw = (unsigned)off;
    ta.raw_bytes = (unsigned)(((size_t)a.N * TR * 32 + 1023) & ~(size_t)1023);
    off += 2 * (size_t)ta.raw_bytes + 1024;
    const int lines = (a.N + 1) * MR;
    ta.pix_warps = MR * XC / 32;
    ta.t_warps = (TR * 2 + 31) / 32;
    int warps = (lines + 31) / 32;
    if (warps < ta.pix_warps + ta.t_warps) warps = ta.pix_warps + ta.t_warps;
    ta.t_warps = warps - ta.pix_warps;            // every warp that is not a pixel warp helps with the patches
    ta.threads = warps * 32;
    if (ta.threads > 1024 || off > 227 * 1024) return 0;
    return off;
}

template <int NH, int BW, int MR, int NT>
cudaError_t launch(const PixArgs &a, const typename BasisSel<NH>::type &bs, int n_stacks, cudaStream_t st, const unsigned char *redo = nullptr) {
    constexpr int TR = MR + BW - 1;
    TmaArgs ta;
    ta.redo = redo;
    const size_t smem = plan<BW, MR>(a, ta);
    if (!smem) return cudaErrorInvalidConfiguration;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return cudaErrorNotSupported;
    ta.isub = (int)a.sub; ta.idark = (int)a.dark;
    ta.U = (a.W + (BW - 1 - BW / 2) + 15) / 16;            // last tile: the one holding image column W - 1 + RX
    pb_ddrecip(BW, &ta.b_hi, &ta.b_lo);
    ta.magic = 4503599627370496.0 + (double)(BW * ta.isub);
    ta.magicT = 4503599627370496.0 + (double)BW * (double)a.N * (double)ta.idark;
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N, (cuuint64_t)n_stacks};
    const cuuint64_t strides[3] = {(cuuint64_t)a.W * 2, (cuuint64_t)a.P * 2, (cuuint64_t)a.stack_stride * 2};
    const cuuint32_t box[4] = {16, (cuuint32_t)TR, (cuuint32_t)a.N, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void *>(a.stack), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    auto kern = k_march_tma<NH, BW, MR, NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((a.H + MR - 1) / MR, n_stacks);
    kern<<<grid, ta.threads, smem, st>>>(map, a, bs, ta);
    return cudaGetLastError();
}

// rows per CTA (16 / 8 / 4). The march of a row block is sequential over W, so fewer rows per CTA do not shorten a single-stack
// call (measured, one 2048x2048x18 stack: 16 rows 0.47 ms, 8 rows 0.71 ms, 4 rows 0.70 ms): the grid size plays no role.
// Chosen: the largest MR whose CTA leaves room for three resident CTAs per SM (<= 75 KB of shared memory: N = 18 -> 16 rows,
// SHG's N = 36 -> 8 rows, measured 314 k against 270 k Mpx-angles/s with 16 rows = one CTA per SM); if none does, the largest
// that fits at all (N = 90, the configs[4] shape: 8 rows, 728 threads, 346 k against 325 k with 4 rows).
#ifndef PB_TMA_MAX_LINES
#define PB_TMA_MAX_LINES 768
#endif
template <int BW>
int rows_for(const PixArgs &a) {
    static const int forced = [] { const char *e = getenv("PB_TMA_ROWS"); return e ? atoi(e) : 0; }();
    if (forced == 16 || forced == 8 || forced == 4) return forced;
    TmaArgs t;
    const size_t s16 = plan<BW, 16>(a, t), s8 = plan<BW, 8>(a, t), s4 = plan<BW, 4>(a, t);
    const size_t sz[3] = {s16, s8, s4};
    const int mr[3] = {16, 8, 4};
    for (int k = 0; k < 3; ++k)
        if (sz[k] && sz[k] <= 75 * 1024 && (a.N + 1) * mr[k] <= PB_TMA_MAX_LINES) return mr[k];
    for (int k = 0; k < 3; ++k)
        if (sz[k] && (a.N + 1) * mr[k] <= PB_TMA_MAX_LINES) return mr[k];
    return 4;
}

template <int NH, int BW>
cudaError_t launch_bw(const PixArgs &a, const typename BasisSel<NH>::type &bs, int n_stacks, cudaStream_t st) {
    const int mr = rows_for<BW>(a);
    // the reference's standard 1PF acquisition (18 angles, PP:1926) with the small windows has fully unrolled variants
    if constexpr (NH == 1 && BW <= 5) {
        if (a.N == 18) {
            switch (mr) {
            case 16: return launch<NH, BW, 16, 18>(a, bs, n_stacks, st);
            case 8: return launch<NH, BW, 8, 18>(a, bs, n_stacks, st);
            default: return launch<NH, BW, 4, 18>(a, bs, n_stacks, st);
            }
        }
    }
    switch (mr) {
    case 16: return launch<NH, BW, 16, 0>(a, bs, n_stacks, st);
    case 8: return launch<NH, BW, 8, 0>(a, bs, n_stacks, st);
    default: return launch<NH, BW, 4, 0>(a, bs, n_stacks, st);
    }
}

template <int NH>
cudaError_t launch_nh(const PixArgs &a, const typename BasisSel<NH>::type &bs, int bw, int n_stacks, cudaStream_t st) {
    switch (bw) {
    case 2: return launch_bw<NH, 2>(a, bs, n_stacks, st);
    case 3: return launch_bw<NH, 3>(a, bs, n_stacks, st);
    case 4: return launch_bw<NH, 4>(a, bs, n_stacks, st);
    case 5: return launch_bw<NH, 5>(a, bs, n_stacks, st);
    case 6: return launch_bw<NH, 6>(a, bs, n_stacks, st);
    case 7: return launch_bw<NH, 7>(a, bs, n_stacks, st);
    case 8: return launch_bw<NH, 8>(a, bs, n_stacks, st);
    }
    return cudaErrorInvalidConfiguration;
}

}  // namespace tma
