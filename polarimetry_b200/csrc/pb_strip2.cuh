// pb_strip2.cuh -- fused single-pass analysis WITH spatial binning, "strip march": a thread carries the box-filter state of its
// planes in registers and a pixel's harmonic sums never pass through an fp64 tile in shared memory; no block barrier, no warp roles.
//
// Replaces, in one kernel (as k_march_tma does): Stack.get_intensity (pypolar_classes.py:260-264), analyze_stack's pre-processing
// (PyPOLAR.py:2953-2957), Polarimetry.binning (PP:2942-2946), compute_fields (PP:2981-3078), the mask logic (PP:2891/2967) and the
// histograms (PC:340-369). The arithmetic is, operation for operation, that of pb_march_tma.cuh / oracle/contract.c (scipy's
// running-sum box filter replayed bit for bit; sequential harmonic sums in plane order, PP:2987-2994).
//
//   CTA = one warp = a strip of 16 image rows of one stack; TWO lanes per row: lane 2r+t holds the running sums and ring entries of
//       planes 9t .. 9t+8 of row r (80 registers of state). The sums over the planes are sequential (PP:2987-2994), so lane t=1 works
//       one 4-column block behind lane t=0: at the end of a step lane 0 leaves its partial sums (planes 0-8) of block b in shared
//       memory, lane 1 picks them up and adds planes 9-17 during the next step -- a two-stage systolic pipeline inside the warp.
//       (The pb_strip.cuh predecessor held all 18 planes in one thread: 176 registers of state left no room for instruction-level
//       parallelism and it ran at a third of k_march_tma's speed.)
//   data: 4-D tensor map [stack][plane][row][column]; tile u = image columns [8(u-1), 8u) x rows [y0-1, y0+17) x 18 planes = one
//       16-byte line per (plane, row), cp.async.bulk.tensor.4d on an mbarrier, two stages per warp (the box start must be 16-byte
//       aligned in global memory: tools/ubench/tma_box.cu); a tile is read during three steps (lane 0: 2u, 2u+1; lane 1: 2u+1, 2u+2),
//       then lane 0 issues tile u+2 into its stage. The two lanes of a row always read different 8-byte halves of their lines, which
//       makes the 64-bit shared loads of a half-warp conflict-free. 'reflect': rows by a per-thread line remap, columns by two
//       synthetic tiles mirrored from the neighbouring real tile in shared memory;
//   a step: per plane the H window sums (packed clamp + IDP.2A), q = RN(S/bh), tmp += q - ring, f = RN(tmp/bw) and straight into
//       the pixel's sums S0 / Sc / Ss / Sff; lane 1 also advances the total-intensity line (PC:260-264) from the summed window sums;
//       then both lanes of the row take two of the finished block's four pixels each through the shared epilogue (pb_pixel.cuh);
//   pixels whose fit / chi2 decision the certified screening cannot take (rare: PP:2990-3000 needs the per-plane values again) are
//       left out and their 16-row block is flagged; k_march_tma then runs in redo mode on the flagged blocks only.
#pragma once
#include "pb_march_tma.cuh"

namespace strip2 {

constexpr int XT = 8;      // columns per tile: 16 bytes of uint16 per (plane, row) line
constexpr int CB = 4;      // columns per step
constexpr int RW = 16;     // rows per warp
#ifndef PB_STRIP2_MAXREG
#define PB_STRIP2_MAXREG 168
#endif
constexpr int NVEC = 11;   // 16-byte vectors per lane slot: 8 (S0,Sc / Ss,Sff per column) + 1 (window-sum totals) + 2 (intensity)

struct Args {
    int U;                       // last tile (the synthetic right-edge tile)
    int b_last;                  // last block: block b = image columns [4(b-2), 4(b-2)+4), tile b >> 1, half b & 1
    int isub, idark;             // dark + removed_intensity, dark
    int nblk16;                  // 16-row blocks per stack (redo flags) = gridDim.x
    unsigned off_bar, off_slot, off_raw, stage_bytes;
    double b_hi, b_lo;           // double-double reciprocal of the (square) bin size
    double magic, magicT;        // 2^52 + bw*isub, 2^52 + bw*N*idark
    unsigned char *redo;         // [n_stacks][nblk16]
};

template <int BW, int NT, bool SAME, bool FAST>
__global__ void __maxnreg__(PB_STRIP2_MAXREG) k_strip2(const __grid_constant__ CUtensorMap tmap, PixArgs a, Basis1 bs, Args sa) {
    constexpr int HH = BW / 2;                          // window extent before the centre (rows above / columns left)
    constexpr int RX = BW - 1 - HH;                     // ... after the centre; outputs lag the consumed column by RX
    constexpr int TR = RW + BW - 1;                     // tile rows
    constexpr int NP = NT / 2;                          // planes per lane
    constexpr unsigned PLB = TR * 16;                   // plane pitch of a stage in bytes
    static_assert(BW == 3 && NT % 2 == 0, "written for the 3-wide window (HH = RX = 1) and an even number of planes");
    extern __shared__ __align__(16) unsigned char smem[];
    const pb_outputs out = a.outs[blockIdx.y];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    const int H = a.H, W = a.W;
    const int lane = threadIdx.x, r = lane >> 1, t = lane & 1;
    const int y0 = blockIdx.x * RW, y = y0 + r, ytop = y0 - HH;
    const bool row_ok = y < H;
    const unsigned sbase = tma::smem_u32(smem);
    const unsigned abase = (sbase + 127u) & ~127u;                    // 128-aligned base of the raw stages
    unsigned char *const raw0 = smem + (abase - sbase) + sa.off_raw;
    const unsigned bar0 = sbase + sa.off_bar;                         // stage s completes on bar0 + 8 s
    const unsigned stage_bytes = sa.stage_bytes;
    const int U = sa.U, b_last = sa.b_last;

    auto issue = [&](int u) {       // lane 0: arm the stage's barrier and start the box copy of tile u
        const unsigned bar = bar0 + 8u * (u & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the warp's generic-proxy accesses of this stage before the async write
        tma::mbar_expect_tx(bar, (unsigned)(NT * TR * 16));
        tma::tma_load_4d(abase + sa.off_raw + (u & 1) * stage_bytes, &tmap, bar, XT * (u - 1), ytop, 0, (int)blockIdx.y);
    };
    auto wait_tile = [&](int u) { tma::mbar_wait(bar0 + 8u * (u & 1), (unsigned)(((u - 1) >> 1) & 1)); };   // real tile u (1 .. U-1) is load number (u-1) >> 1 of its stage
    // synthetic edge tile: the mirror image of the neighbouring real tile (scipy 'reflect': column -1 = column 0, column W = column W-1),
    // restricted to the one column the 3-wide window reaches; the rest stays 0 (= no contribution after the clamp)
    auto synth = [&](int u_dst, int u_src, bool left) {
        const uint4 *src = reinterpret_cast<const uint4 *>(raw0 + (u_src & 1) * stage_bytes);
        uint4 *dst = reinterpret_cast<uint4 *>(raw0 + (u_dst & 1) * stage_bytes);
        for (int L = lane; L < NT * TR; L += 32) {
            const uint4 w = src[L];
            // left: halfword 7 (column -1) <- halfword 0 of the source (column 0); right: halfword 0 (column W) <- halfword 7 (column W-1)
            dst[L] = left ? make_uint4(0u, 0u, 0u, w.x << 16) : make_uint4(w.w >> 16, 0u, 0u, 0u);
        }
        __syncwarp();
    };

    if (lane == 0) {
        tma::mbar_init(bar0, 1);
        tma::mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    HistSmem hs;
    hs.cnt = nullptr; hs.edges = nullptr;
    if (do_hist) hs = hist_setup(a, smem);          // edges + counters in shared memory; ends with a block barrier
    else __syncthreads();
    if (lane == 0) issue(1);

    // hand-over slots: vector j of lane L at slot + (32 j + L) 16 bytes (conflict-free 128-bit accesses); one vector of zeros behind them
    double2 *const slot = reinterpret_cast<double2 *>(smem + sa.off_slot);
    if (lane == 0) slot[NVEC * 32] = make_double2(0.0, 0.0);
    // lane 1 continues from its partner's partial sums; lane 0 starts every block from zero (all its vectors are the zero vector)
    const double2 *const acc_src = t ? slot + (lane - 1) : slot + NVEC * 32;
    const int acc_stride = t ? 32 : 0;
    double2 *const my_slot = slot + lane;
    const double2 *const fin = slot + (lane | 1) + 4 * t * 32;       // epilogue: this lane's two pixels (columns 2t, 2t+1 of the block)
    const double2 *const fin_I = slot + (lane | 1) + (9 + t) * 32;

    // per-thread line offsets of the window rows (reflected rows are rows of the same tile) + this lane's first plane
    unsigned lo[BW];
    {
        const int yc = min(y, H - 1);
#pragma unroll
        for (int dy = 0; dy < BW; ++dy) lo[dy] = (unsigned)(pb_reflect(yc - HH + dy, H) - ytop) * 16u + (unsigned)(t * NP) * PLB;
    }
    const int pt = t * NP;
    double tmp[NP + 1], ring[NP + 1][BW];        // entry NP: the total-intensity line (lane 1)
#pragma unroll
    for (int p = 0; p <= NP; ++p) {
        tmp[p] = 0.0;
#pragma unroll
        for (int s = 0; s < BW; ++s) ring[p][s] = 0.0;
    }
    const bool f64 = a.flags & PB_FLAG_OUT_F64;
    const bool lean = store_is_lean<1>(a, out);
    const unsigned sub2 = (unsigned)sa.isub * 0x10001u, dark2 = (unsigned)sa.idark * 0x10001u;
    bool need_any = false;
    const long long rowp = (long long)y * W;

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
        __syncwarp();          // every read of the slots (these and the previous step's epilogue) before they are rewritten
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
                const double c2 = bs.c2[pt + p], s2 = bs.s2[pt + p];
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
        // outgoing sums: lane 0's partial sums for its partner, lane 1's finished sums + intensities for the epilogue
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            my_slot[(2 * c) * 32] = make_double2(S0[c], Sc[c]);
            my_slot[(2 * c + 1) * 32] = make_double2(Ss[c], Sff[c]);
        }
        *reinterpret_cast<uint4 *>(my_slot + 8 * 32) = make_uint4(Ts[0], Ts[1], Ts[2], Ts[3]);
        my_slot[9 * 32] = make_double2(Iv[0], Iv[1]);
        my_slot[10 * 32] = make_double2(Iv[2], Iv[3]);
        __syncwarp();
        // ---- epilogue of the block lane 1 has just finished (block s - 1): image columns k0 .. k0 + 3, two per lane
        const int be = s - 1;
        if (row_ok && be >= 2 && be <= b_last) {
            const int k0 = CB * (be - 2) - RX + 2 * t;
            const double2 e0 = fin[0], e1 = fin[32], e2 = fin[64], e3 = fin[96], eI = fin_I[0];
            if constexpr (FAST) {
                // whole-image thresholding, float32 maps + mask (established by the launcher). (Measured: running the lane's two pixels
                // side by side on stand-in values for the ones that take no part costs more instructions than the overlap of their
                // latencies buys: 361 k -> 321 k Mpx-angles/s.)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int k = k0 + c;
                    if (k < 0 || k >= W) continue;
                    const long long p = rowp + k;
                    float r0 = __int_as_float(0x7fc00000), r1 = r0, r2 = r0;
                    uint8_t mk = 0;
                    if ((c ? eI.y : eI.x) >= a.roi_ilow[0]) {
                        const HarmPix hp = harm_stage_a_core<1>(a, 1u, c ? e2.x : e0.x, c ? e2.y : e0.y, c ? e3.x : e1.x, 0.0, 0.0, c ? e3.y : e1.y);
                        if (hp.need) { need_any = true; continue; }          // decided by k_march_tma's redo pass (exact loop of PP:2990-3000)
                        const PixOut o = harm_stage_b<1, 1>(a, hs, do_hist, hp);
                        r0 = (float)o.v0; r1 = (float)o.v1; r2 = (float)o.a0m;
                        mk = (uint8_t)o.m;
                    }
                    reinterpret_cast<float *>(out.var[0])[p] = r0;
                    reinterpret_cast<float *>(out.var[1])[p] = r1;
                    reinterpret_cast<float *>(out.intmap)[p] = r2;
                    out.mask[p] = mk;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int k = k0 + c;
                    if (k < 0 || k >= W) continue;
                    const double s0 = c ? e2.x : e0.x, sc = c ? e2.y : e0.y, ss = c ? e3.x : e1.x, sf = c ? e3.y : e1.y;
                    const double I = c ? eI.y : eI.x;
                    const long long p = rowp + k;
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
        if (!(s & 1)) {
            // lane 1 has read tile s/2 - 1 for the last time: its stage takes tile s/2 + 1
            __syncwarp();
            const int v = (s >> 1) + 1;
            if (v < U) {
                if (lane == 0) issue(v);
            } else if (v == U) {
                wait_tile(U - 1);
                synth(U, U - 1, false);
            }
        }
    }
    if (sa.redo) {
        const unsigned nb = __ballot_sync(0xffffffffu, need_any);
        if (lane == 0) sa.redo[(size_t)blockIdx.y * sa.nblk16 + blockIdx.x] = nb ? 1 : 0;
    }
    if (do_hist) hist_flush(a, hs, out.hist);
}

// ---------------------------------------------------------------------------------------------------
template <int BW, int NT, bool SAME, bool FAST>
cudaError_t launch(const PixArgs &a, const Basis1 &bs, int n_stacks, unsigned char *redo, cudaStream_t st) {
    constexpr int TR = RW + BW - 1;
    Args sa;
    size_t off = tma::hist_bytes_tma(a);
    sa.off_bar = (unsigned)off; off += 16;
    sa.off_slot = (unsigned)off; off += (size_t)(NVEC * 32 + 1) * 16;
    off = (off + 127) & ~(size_t)127;
    sa.off_raw = (unsigned)off;
    sa.stage_bytes = (unsigned)(((size_t)NT * TR * 16 + 127) & ~(size_t)127);
    off += 2 * (size_t)sa.stage_bytes + 128;           // + the alignment slack of the dynamic shared memory base
    if (off > 200 * 1024) return cudaErrorInvalidConfiguration;
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
    auto kern = k_strip2<BW, NT, SAME, FAST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    dim3 grid((a.H + RW - 1) / RW, n_stacks);
    kern<<<grid, 32, off, st>>>(map, a, bs, sa);
    return cudaGetLastError();
}

}  // namespace strip2
