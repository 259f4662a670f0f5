// pb_strip3.cuh -- fused single-pass analysis WITH spatial binning, "strip march", warp-specialised: the march of pb_strip2.cuh
// (two lanes per image row carry the box-filter state of all planes in registers; a pixel's harmonic sums never pass through an
// fp64 tile) and the per-pixel epilogue run in DIFFERENT warps of one CTA, so that the epilogue's table-lookup latencies (disk cone
// in L2) never stall a march warp and the march warps' registers are not held during the epilogue.
//
// Replaces, in one kernel (as k_march_tma does): Stack.get_intensity (pypolar_classes.py:260-264), analyze_stack's pre-processing
// (PyPOLAR.py:2953-2957), Polarimetry.binning (PP:2942-2946), compute_fields (PP:2981-3078), the mask logic (PP:2891/2967) and the
// histograms (PC:340-369). Arithmetic: operation for operation that of pb_march_tma.cuh / oracle/contract.c.
//
//   CTA = 8 warps = 4 pairs; pair i = march warp i + epilogue warp 4+i working on one strip of 16 image rows (CTA: 64 rows of a stack).
//   Warps 0-3 raise their register budget (setmaxnreg.inc), warps 4-7 lower it (setmaxnreg.dec): 2 CTAs per SM.
//   march warp: as k_strip2 -- lane 2r+t holds planes 9t..9t+8 of row r; lane 1 works one 4-column block behind lane 0 and continues
//       from lane 0's partial sums (shared-memory slots); TMA tiles of 8 columns x 18 rows x 18 planes, two stages, issued by lane 0;
//       at the end of a step lane 1 leaves the block's finished sums S0 / Sc / Ss / Sff and the binned intensities in a double-
//       buffered slot and arrives on the pair's `full` mbarrier;
//   epilogue warp: waits for `full`, copies its two pixels' sums into registers, arrives on `empty` (the march warp may refill the
//       slot) and runs the shared per-pixel epilogue (pb_pixel.cuh): threshold, a0 / a2, certified chi2 screening, disk-cone lookup,
//       histograms, stores;
//   pixels whose fit / chi2 decision the certified screening cannot take are left out and their 16-row block is flagged;
//       k_march_tma then runs in redo mode on the flagged blocks only.
#pragma once
#include "pb_march_tma.cuh"

namespace strip3 {

constexpr int XT = 8;       // columns per tile: 16 bytes of uint16 per (plane, row) line
constexpr int CB = 4;       // columns per step
constexpr int RW = 16;      // rows per pair
constexpr int NPAIR = 4;    // pairs per CTA
constexpr int NV = 10;      // 16-byte vectors per slot row: 8 (S0,Sc / Ss,Sff per column) + 2 (window-sum totals | intensities)
#ifndef PB_STRIP3_MARCH_REGS
#define PB_STRIP3_MARCH_REGS 192
#endif
#ifndef PB_STRIP3_EPI_REGS
#define PB_STRIP3_EPI_REGS 64
#endif
// byte offsets inside a pair's shared-memory region (128-byte aligned)
constexpr unsigned P_BAR = 0;                       // tile barriers (2), full (2), empty (2)
constexpr unsigned P_ZERO = 64;                     // one vector of zeros
constexpr unsigned P_PART = 128;                    // lane 0 -> lane 1 partial sums [NV][RW] vectors
constexpr unsigned P_FIN = P_PART + NV * RW * 16;   // finished sums, two buffers [2][NV][RW]
constexpr unsigned P_RAW = (P_FIN + 2 * NV * RW * 16 + 127) & ~127u;

struct Args {
    int U;                       // last tile (the synthetic right-edge tile)
    int b_last;                  // last block: block b = image columns [4(b-2), 4(b-2)+4), tile b >> 1, half b & 1
    int isub, idark;             // dark + removed_intensity, dark
    int nblk16;                  // 16-row blocks per stack (redo flags)
    unsigned off_basis, off_pair, pair_bytes, stage_bytes;
    double b_hi, b_lo;           // double-double reciprocal of the (square) bin size
    double magic, magicT;        // 2^52 + bw*isub, 2^52 + bw*N*idark
    unsigned char *redo;         // [n_stacks][nblk16]
};

// parity wait with a watchdog: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_wd(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    for (unsigned spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok) {
            __nanosleep(64);                       // a waiting warp must not compete for issue slots with the warps it waits for
            if (spin > (1u << 22)) __trap();
        }
    }
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int BW, int NT, bool SAME>
__device__ __forceinline__ void march_warp(const CUtensorMap *tmap, const PixArgs &a, const unsigned char *cs_basis, const Args &sa, unsigned char *pbase,
                                           unsigned pbase_s, int lane, int y0, int stack) {
    constexpr int HH = BW / 2;
    constexpr int TR = RW + BW - 1;                     // tile rows
    constexpr int NP = NT / 2;                          // planes per lane
    constexpr unsigned PLB = TR * 16;                   // plane pitch of a stage in bytes
    const int H = a.H;
    const int r = lane >> 1, t = lane & 1;
    const int y = y0 + r, ytop = y0 - HH;
    const bool row_ok = y < H;
    unsigned char *const raw0 = pbase + P_RAW;
    const unsigned bar0 = pbase_s + P_BAR;              // stage s completes on bar0 + 8 s
    const unsigned full0 = pbase_s + P_BAR + 16, empty0 = pbase_s + P_BAR + 32;
    const unsigned stage_bytes = sa.stage_bytes;
    const int U = sa.U, b_last = sa.b_last;

    auto issue = [&](int u) {       // lane 0: arm the stage's barrier and start the box copy of tile u
        const unsigned bar = bar0 + 8u * (u & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the warp's generic-proxy accesses of this stage before the async write
        tma::mbar_expect_tx(bar, (unsigned)(NT * TR * 16));
        tma::tma_load_4d(pbase_s + P_RAW + (u & 1) * stage_bytes, tmap, bar, XT * (u - 1), ytop, 0, stack);
    };
    auto wait_tile = [&](int u) { mbar_wait_wd(bar0 + 8u * (u & 1), (unsigned)(((u - 1) >> 1) & 1)); };   // real tile u (1 .. U-1) is load number (u-1) >> 1 of its stage
    // synthetic edge tile: the mirror image of the neighbouring real tile (scipy 'reflect': column -1 = column 0, column W = column W-1),
    // restricted to the one column the 3-wide window reaches; the rest stays 0 (= no contribution after the clamp)
    auto synth = [&](int u_dst, int u_src, bool left) {
        const uint4 *src = reinterpret_cast<const uint4 *>(raw0 + (u_src & 1) * stage_bytes);
        uint4 *dst = reinterpret_cast<uint4 *>(raw0 + (u_dst & 1) * stage_bytes);
        for (int L = lane; L < NT * TR; L += 32) {
            const uint4 w = src[L];
            dst[L] = left ? make_uint4(0u, 0u, 0u, w.x << 16) : make_uint4(w.w >> 16, 0u, 0u, 0u);
        }
        __syncwarp();
    };
    if (lane == 0) issue(1);

    double2 *const part = reinterpret_cast<double2 *>(pbase + P_PART);
    double2 *const fin = reinterpret_cast<double2 *>(pbase + P_FIN);
    // lane 1 continues from its partner's partial sums; lane 0 starts every block from zero (all its vectors are the zero vector)
    const double2 *const acc_src = t ? part + r : reinterpret_cast<const double2 *>(pbase + P_ZERO);
    const int acc_stride = t ? RW : 0;

    // per-thread line offsets of the window rows (reflected rows are rows of the same tile) + this lane's first plane
    unsigned lo[BW];
    {
        const int yc = min(y, H - 1);
#pragma unroll
        for (int dy = 0; dy < BW; ++dy) lo[dy] = (unsigned)(pb_reflect(yc - HH + dy, H) - ytop) * 16u + (unsigned)(t * NP) * PLB;
    }
    const double2 *const basis = reinterpret_cast<const double2 *>(cs_basis) + t * NP;      // (cos, sin)(2 alpha_p) of this lane's planes
    double tmp[NP + 1], ring[NP + 1][BW];        // entry NP: the total-intensity line (lane 1)
#pragma unroll
    for (int p = 0; p <= NP; ++p) {
        tmp[p] = 0.0;
#pragma unroll
        for (int s = 0; s < BW; ++s) ring[p][s] = 0.0;
    }
    const unsigned sub2 = (unsigned)sa.isub * 0x10001u, dark2 = (unsigned)sa.idark * 0x10001u;

    wait_tile(1);
    synth(0, 1, true);

    const int s_last = b_last + 1;
    for (int s = 1; s <= s_last; ++s) {
        if (!(s & 1) && (s >> 1) < U) wait_tile(s >> 1);       // lane 0 enters tile s/2 (tile U is synthetic: built when its stage was released)
        const int bl = s - t;                                   // this lane's block
        const bool act = row_ok && bl >= 1 && bl <= b_last;
        // incoming sums
        double S0[CB], Sc[CB], Ss[CB], Sff[CB];
        unsigned Ts[CB];
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            const double2 v0 = acc_src[(2 * c) * acc_stride], v1 = acc_src[(2 * c + 1) * acc_stride];
            S0[c] = v0.x; Sc[c] = v0.y; Ss[c] = v1.x; Sff[c] = v1.y;
        }
        {
            const uint4 v = *reinterpret_cast<const uint4 *>(acc_src + 8 * acc_stride);
            Ts[0] = v.x; Ts[1] = v.y; Ts[2] = v.z; Ts[3] = v.w;
        }
        __syncwarp();          // lane 1's reads of the partial sums before lane 0 rewrites them
        double Iv[CB] = {0.0, 0.0, 0.0, 0.0};
        if (act) {
            const int u = bl >> 1, half = bl & 1;
            const unsigned char *rb = raw0 + (u & 1) * stage_bytes + half * 8;
            const unsigned char *rl[BW];
#pragma unroll
            for (int dy = 0; dy < BW; ++dy) rl[dy] = rb + lo[dy];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                unsigned sw[CB] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int dy = 0; dy < BW; ++dy) {
                    const uint2 w = *reinterpret_cast<const uint2 *>(rl[dy] + p * PLB);
                    const unsigned m0 = __vmaxu2(w.x, sub2), m1 = __vmaxu2(w.y, sub2);
                    // __dp2a_lo(m, b, c) = c + m.lo16 * b.byte0 + m.hi16 * b.byte1
                    sw[0] = __dp2a_lo(m0, 0x0001u, sw[0]); sw[1] = __dp2a_lo(m0, 0x0100u, sw[1]);
                    sw[2] = __dp2a_lo(m1, 0x0001u, sw[2]); sw[3] = __dp2a_lo(m1, 0x0100u, sw[3]);
                    if constexpr (!SAME) {          // removed_intensity != 0: the threshold image clamps at the dark value alone (PC:261)
                        const unsigned d0 = __vmaxu2(w.x, dark2), d1 = __vmaxu2(w.y, dark2);
                        Ts[0] = __dp2a_lo(d0, 0x0001u, Ts[0]); Ts[1] = __dp2a_lo(d0, 0x0100u, Ts[1]);
                        Ts[2] = __dp2a_lo(d1, 0x0001u, Ts[2]); Ts[3] = __dp2a_lo(d1, 0x0100u, Ts[3]);
                    }
                }
                const double2 cs = basis[p];
                const double c2 = cs.x, s2 = cs.y;
                double q[CB];
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    if constexpr (SAME) Ts[c] += sw[c];
                    const double uu = __hiloint2double(0x43300000, (int)sw[c]) - sa.magic;
                    q[c] = fma(uu, sa.b_hi, uu * sa.b_lo);
                }
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const double old = c < BW ? ring[p][c < BW ? c : 0] : q[c >= BW ? c - BW : 0];
                    const double d = q[c] - old;
                    tmp[p] = tmp[p] + d;
                    const double f = fma(tmp[p], sa.b_hi, tmp[p] * sa.b_lo);
                    S0[c] = S0[c] + f;
                    Sc[c] = Sc[c] + f * c2;
                    Ss[c] = Ss[c] + f * s2;
                    Sff[c] = fma(f, f, Sff[c]);                    // screening only
                }
#pragma unroll
                for (int k = 0; k < BW; ++k) ring[p][k] = q[CB - BW + k];
            }
            if (t) {
                // the total-intensity line (PC:260-264): the same two passes on the integer sums over all the angles
                double q[CB];
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const double uu = __hiloint2double(0x43300000, (int)Ts[c]) - sa.magicT;
                    q[c] = fma(uu, sa.b_hi, uu * sa.b_lo);
                }
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const double old = c < BW ? ring[NP][c < BW ? c : 0] : q[c >= BW ? c - BW : 0];
                    const double d = q[c] - old;
                    tmp[NP] = tmp[NP] + d;
                    Iv[c] = fma(tmp[NP], sa.b_hi, tmp[NP] * sa.b_lo);
                }
#pragma unroll
                for (int k = 0; k < BW; ++k) ring[NP][k] = q[CB - BW + k];
            }
        }
        // outgoing sums: lane 0's partial sums for its partner, lane 1's finished sums + intensities for the epilogue warp
        const unsigned fb = (unsigned)(s & 1);
        if (s >= 3) mbar_wait_wd(empty0 + 8u * fb, (unsigned)(((s - 3) >> 1) & 1));     // the epilogue warp has taken step s - 2 out of this buffer
        double2 *const dst = t ? fin + fb * (NV * RW) + r : part + r;
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            dst[(2 * c) * RW] = make_double2(S0[c], Sc[c]);
            dst[(2 * c + 1) * RW] = make_double2(Ss[c], Sff[c]);
        }
        if (t) {
            dst[8 * RW] = make_double2(Iv[0], Iv[1]);
            dst[9 * RW] = make_double2(Iv[2], Iv[3]);
        } else {
            *reinterpret_cast<uint4 *>(dst + 8 * RW) = make_uint4(Ts[0], Ts[1], Ts[2], Ts[3]);
        }
        mbar_arrive(full0 + 8u * fb);          // release: this lane's slot writes are visible to whoever observes the completed phase
        __syncwarp();                          // lane 0's partial sums before lane 1 reads them
        if (!(s & 1)) {
            // lane 1 has read tile s/2 - 1 for the last time: its stage takes tile s/2 + 1
            const int v = (s >> 1) + 1;
            if (v < U) {
                if (lane == 0) issue(v);
            } else if (v == U) {
                wait_tile(U - 1);
                synth(U, U - 1, false);
            }
        }
    }
}

template <int BW, bool FAST>
__device__ __forceinline__ void epilogue_warp(const PixArgs &a, const Args &sa, const pb_outputs &out, const HistSmem &hs, bool do_hist, unsigned char *pbase,
                                              unsigned pbase_s, int lane, int y0, int stack, int blk16) {
    constexpr int RX = BW - 1 - BW / 2;
    const int H = a.H, W = a.W;
    const int r = lane >> 1, t = lane & 1;
    const int y = y0 + r;
    const bool row_ok = y < H;
    const unsigned full0 = pbase_s + P_BAR + 16, empty0 = pbase_s + P_BAR + 32;
    const double2 *const fin = reinterpret_cast<const double2 *>(pbase + P_FIN) + r + 4 * t * RW;      // this lane's two pixels: columns 2t, 2t+1
    const double2 *const fin_I = reinterpret_cast<const double2 *>(pbase + P_FIN) + r + (8 + t) * RW;
    const bool f64 = a.flags & PB_FLAG_OUT_F64;
    const bool lean = store_is_lean<1>(a, out);
    const long long rowp = (long long)y * W;
    bool need_any = false;
    const int b_last = sa.b_last, s_last = b_last + 1;
    for (int s = 1; s <= s_last; ++s) {
        const unsigned fb = (unsigned)(s & 1);
        mbar_wait_wd(full0 + 8u * fb, (unsigned)(((s - 1) >> 1) & 1));
        const int be = s - 1;                                    // the block lane 1 of the march warp has just finished
        const bool work = row_ok && be >= 2 && be <= b_last;
        double2 e0, e1, e2, e3, eI;
        if (work) {
            const double2 *f = fin + fb * (NV * RW);
            e0 = f[0]; e1 = f[RW]; e2 = f[2 * RW]; e3 = f[3 * RW];
            eI = fin_I[fb * (NV * RW)];
        }
        mbar_arrive(empty0 + 8u * fb);         // release: the loads above are done, the march warp may refill the buffer
        if (!work) continue;
        const int k0 = CB * (be - 2) - RX + 2 * t;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int k = k0 + c;
            if (k < 0 || k >= W) continue;
            const double s0 = c ? e2.x : e0.x, sc = c ? e2.y : e0.y, ss = c ? e3.x : e1.x, sf = c ? e3.y : e1.y;
            const double I = c ? eI.y : eI.x;
            const long long p = rowp + k;
            if constexpr (FAST) {
                // whole-image thresholding, float32 maps + mask (established by the launcher)
                float r0 = __int_as_float(0x7fc00000), r1 = r0, r2 = r0;
                uint8_t mk = 0;
                if (I >= a.roi_ilow[0]) {
                    const HarmPix hp = harm_stage_a_core<1>(a, 1u, s0, sc, ss, 0.0, 0.0, sf);
                    if (hp.need) { need_any = true; continue; }          // decided by k_march_tma's redo pass (exact loop of PP:2990-3000)
                    const PixOut o = harm_stage_b<1, 1>(a, hs, do_hist, hp);
                    r0 = (float)o.v0; r1 = (float)o.v1; r2 = (float)o.a0m;
                    mk = (uint8_t)o.m;
                }
                reinterpret_cast<float *>(out.var[0])[p] = r0;
                reinterpret_cast<float *>(out.var[1])[p] = r1;
                reinterpret_cast<float *>(out.intmap)[p] = r2;
                out.mask[p] = mk;
            } else {
                bool cand;
                if (a.simple) cand = I >= a.roi_ilow[0];
                else if (a.mask_in) cand = __ldg(a.mask_in + p) != 0;
                else cand = roi_pass(a, a.roi_bits ? __ldg(a.roi_bits + p) : 1u, I);
                PixOut o;
                o.v0 = o.v1 = o.v2 = o.v3 = o.a0m = pb_nan();
                o.m = false;
                if (cand) {
                    const uint32_t bits = (a.simple || !a.roi_bits) ? 1u : __ldg(a.roi_bits + p);
                    const HarmPix hp = harm_stage_a_core<1>(a, bits, s0, sc, ss, 0.0, 0.0, sf);
                    if (hp.need) { need_any = true; continue; }          // left to the redo pass, nothing stored
                    o = harm_stage_b<1>(a, hs, do_hist, hp);
                }
                if (!cand && lean) store_invalid_lean<1>(a, out, p);
                else store_pixel<1>(a, out, p, o, I, f64, lean);
            }
        }
    }
    if (sa.redo) {
        const unsigned nb = __ballot_sync(0xffffffffu, need_any);
        if (lane == 0) sa.redo[(size_t)stack * sa.nblk16 + blk16] = nb ? 1 : 0;
    }
}

template <int BW, int NT, bool SAME, bool FAST>
__global__ void __launch_bounds__(32 * 2 * NPAIR, 2) k_strip3(const __grid_constant__ CUtensorMap tmap, PixArgs a, Basis1 bs, Args sa) {
    static_assert(BW == 3 && NT % 2 == 0, "written for the 3-wide window (HH = RX = 1) and an even number of planes");
    extern __shared__ __align__(16) unsigned char smem[];
    const pb_outputs out = a.outs[blockIdx.y];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp & (NPAIR - 1);
    const bool marching = warp < NPAIR;
    const int blk16 = blockIdx.x * NPAIR + pair;
    const int y0 = blk16 * RW;
    const bool strip_ok = y0 < a.H;
    const unsigned sbase = tma::smem_u32(smem);
    const unsigned abase = (sbase + sa.off_pair + 127u) & ~127u;      // 128-aligned base of the pair regions
    unsigned char *const pbase = smem + (abase - sbase) + pair * sa.pair_bytes;
    const unsigned pbase_s = abase + pair * sa.pair_bytes;
    if (threadIdx.x < NPAIR) {
        const unsigned b = abase + threadIdx.x * sa.pair_bytes + P_BAR;
        tma::mbar_init(b, 1); tma::mbar_init(b + 8, 1);               // tile stages: one expect_tx arrival each
        tma::mbar_init(b + 16, 32); tma::mbar_init(b + 24, 32);       // full: every march lane arrives
        tma::mbar_init(b + 32, 32); tma::mbar_init(b + 40, 32);       // empty: every epilogue lane arrives
        *reinterpret_cast<double2 *>(smem + (abase - sbase) + threadIdx.x * sa.pair_bytes + P_ZERO) = make_double2(0.0, 0.0);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // harmonic basis as (cos, sin) pairs in shared memory: a lane's planes depend on its parity, which rules out constant-bank operands
    unsigned char *const cs_basis = smem + sa.off_basis;
    if (threadIdx.x < NT) reinterpret_cast<double2 *>(cs_basis)[threadIdx.x] = make_double2(bs.c2[threadIdx.x], bs.s2[threadIdx.x]);
    HistSmem hs;
    hs.cnt = nullptr; hs.edges = nullptr;
    if (do_hist) hs = hist_setup(a, smem);          // edges + counters in shared memory; ends with a block barrier
    else __syncthreads();
    if (marching) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(PB_STRIP3_MARCH_REGS));
        if (strip_ok) march_warp<BW, NT, SAME>(&tmap, a, cs_basis, sa, pbase, pbase_s, lane, y0, (int)blockIdx.y);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(PB_STRIP3_EPI_REGS));
        if (strip_ok) epilogue_warp<BW, FAST>(a, sa, out, hs, do_hist, pbase, pbase_s, lane, y0, (int)blockIdx.y, blk16);
    }
    if (do_hist) hist_flush(a, hs, out.hist);       // begins with a block barrier
}

// ---------------------------------------------------------------------------------------------------
template <int BW, int NT, bool SAME, bool FAST>
cudaError_t launch(const PixArgs &a, const Basis1 &bs, int n_stacks, unsigned char *redo, cudaStream_t st) {
    constexpr int TR = RW + BW - 1;
    Args sa;
    size_t off = tma::hist_bytes_tma(a);
    sa.off_basis = (unsigned)off; off += (size_t)NT * 16;
    sa.off_pair = (unsigned)off;
    sa.stage_bytes = (unsigned)(((size_t)NT * TR * 16 + 127) & ~(size_t)127);
    sa.pair_bytes = (unsigned)((P_RAW + 2 * (size_t)sa.stage_bytes + 127) & ~(size_t)127);
    off += (size_t)NPAIR * sa.pair_bytes + 128;          // + the alignment slack of the dynamic shared memory base
    if (off > 110 * 1024) return cudaErrorInvalidConfiguration;
    tma::EncodeTiledFn enc = tma::encode_fn();
    if (!enc) return cudaErrorNotSupported;
    sa.U = a.W / XT + 1;                               // tiles 1 .. U-1 hold the image columns, tile U starts with image column W
    sa.b_last = 2 * sa.U;
    sa.isub = (int)a.sub; sa.idark = (int)a.dark;
    sa.nblk16 = (a.H + RW - 1) / RW;
    sa.redo = redo;
    pb_ddrecip(BW, &sa.b_hi, &sa.b_lo);
    sa.magic = 4503599627370496.0 + (double)(BW * sa.isub);
    sa.magicT = 4503599627370496.0 + (double)BW * (double)NT * (double)sa.idark;
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)NT, (cuuint64_t)n_stacks};
    const cuuint64_t strides[3] = {(cuuint64_t)a.W * 2, (cuuint64_t)a.P * 2, (cuuint64_t)a.stack_stride * 2};
    const cuuint32_t box[4] = {XT, (cuuint32_t)TR, (cuuint32_t)NT, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void *>(a.stack), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    auto kern = k_strip3<BW, NT, SAME, FAST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off);
    if (e != cudaSuccess) return e;
    // shared memory for the two CTAs the register file holds; the rest of the 256 KB stays L1 (spilled values, disk-cone table lines)
    static const int carve = [] { const char *c = getenv("PB_STRIP_CARVE"); return c ? atoi(c) : -1; }();
    int pct = carve >= 0 ? carve : (int)((2 * (off + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
    if (pct > 100) pct = 100;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e != cudaSuccess) return e;
    dim3 grid((a.H + RW * NPAIR - 1) / (RW * NPAIR), n_stacks);
    kern<<<grid, 32 * 2 * NPAIR, off, st>>>(map, a, bs, sa);
    return cudaGetLastError();
}

}  // namespace strip3
