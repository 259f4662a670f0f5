// pb_strip.cuh -- fused single-pass analysis WITH spatial binning, "strip march": one warp = a strip of 32 image rows of one stack,
// one THREAD = one image row carrying the box-filter state of ALL N planes in registers (N running sums + N x bw ring entries),
// so that a pixel's harmonic sums never leave the thread: no fp64 tile in shared memory, no block barrier, no role switch.
//
// Replaces, in one kernel (as k_march_tma does): Stack.get_intensity (pypolar_classes.py:260-264), analyze_stack's pre-processing
// (PyPOLAR.py:2953-2957), Polarimetry.binning (PP:2942-2946), compute_fields (PP:2981-3078), the mask logic (PP:2891/2967) and the
// histograms (PC:340-369). The arithmetic is, operation for operation, that of pb_march_tma.cuh / oracle/contract.c (scipy's
// running-sum box filter replayed bit for bit; sequential harmonic sums in plane order, PP:2987-2994).
//
//   CTA = one warp (the hardware block scheduler hands out the strips: 64 per 2048-row stack); 10 CTAs per SM (200 registers);
//   data: 4-D tensor map [stack][plane][row][column]; tile u = image columns [8(u-1), 8u) x rows [y0 - bh/2, y0 + 32 + bh-1-bh/2)
//       x all planes = one 16-byte line per (plane, row), brought in by cp.async.bulk.tensor.4d on an mbarrier, two stages per warp
//       (the box start must be 16-byte aligned in global memory: measured, tools/ubench/tma_box.cu); lane 0 issues tile u+2 as soon
//       as the warp has read tile u. The 'reflect' boundary needs no global load: rows by a per-thread line remap (the reflected
//       rows are rows of the same tile), columns by two synthetic tiles mirrored from the neighbouring real tile in shared memory
//       (left of column 0, right of column W-1);
//   a step = a block of 4 columns (one 64-bit shared load per plane and window row): per plane the H window sums (packed clamp +
//       IDP.2A), q = RN(S/bh), tmp += q - ring, f = RN(tmp/bw), and straight into the pixel's sums S0 / Sc / Ss / Sff;
//       then the shared per-pixel epilogue (pb_pixel.cuh) for the thread's 4 pixels and 128-bit stores of the float maps;
//   pixels whose fit / chi2 decision the certified screening cannot take (rare: PP:2990-3000 needs the per-plane values again) are
//       left out and their 16-row block is flagged; k_march_tma then runs in redo mode on the flagged blocks only.
#pragma once
#include "pb_march_tma.cuh"

namespace strip {

constexpr int XT = 8;      // columns per tile: 16 bytes of uint16 per (plane, row) line
constexpr int CB = 4;      // columns per step
constexpr int ROWS = 32;   // rows per strip = lanes
#ifndef PB_STRIP_MAXREG
#define PB_STRIP_MAXREG 200
#endif

struct StripArgs {
    int U;                       // last tile (the synthetic right-edge tile when the window reaches beyond column W-1)
    int b_last;                  // last block: block b = columns [4(b-2), 4(b-2)+4), tile b >> 1, half b & 1
    int isub;                    // dark + removed_intensity
    int idark;
    int nblk16;                  // 16-row blocks per stack (redo flags)
    unsigned off_bar, off_cry, off_raw, stage_bytes;
    double b_hi, b_lo;           // double-double reciprocal of the (square) bin size
    double magic, magicT;        // 2^52 + bw*isub, 2^52 + bw*N*idark
    unsigned char *redo;         // [n_stacks][nblk16]
};

template <int BW, int NT, bool SAME, bool FAST>
__global__ void __maxnreg__(PB_STRIP_MAXREG) k_strip(const __grid_constant__ CUtensorMap tmap, PixArgs a, Basis1 bs, StripArgs sa) {
    constexpr int HH = BW / 2;                          // window extent before the centre (rows above / columns left)
    constexpr int RX = BW - 1 - HH;                     // ... after the centre; outputs lag the consumed column by RX
    constexpr int TR = ROWS + BW - 1;                   // tile rows
    constexpr unsigned PLB = TR * 16;                   // plane pitch of a stage in bytes
    static_assert(BW == 3, "the synthetic edge tiles and the aligned-store carry are written for the 3-wide window (HH = RX = 1)");
    extern __shared__ __align__(16) unsigned char smem[];
    const pb_outputs out = a.outs[blockIdx.y];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    const int H = a.H, W = a.W;
    const int lane = threadIdx.x;
    const int y0 = blockIdx.x * ROWS, y = y0 + lane, ytop = y0 - HH;
    const bool row_ok = y < H;
    const unsigned sbase = tma::smem_u32(smem);
    const unsigned abase = (sbase + 127u) & ~127u;                    // 128-aligned base of the raw stages
    unsigned char *const raw0 = smem + (abase - sbase) + sa.off_raw;
    const unsigned bar0 = sbase + sa.off_bar;                         // stage s completes on bar0 + 8 s
    const unsigned stage_bytes = sa.stage_bytes;
    const int U = sa.U;

    auto issue = [&](int u) {       // lane 0: arm the stage's barrier and start the box copy of tile u
        const unsigned bar = bar0 + 8u * (u & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the warp's generic-proxy accesses of this stage before the async write
        tma::mbar_expect_tx(bar, (unsigned)(NT * TR * 16));
        tma::tma_load_4d(abase + sa.off_raw + (u & 1) * stage_bytes, &tmap, bar, XT * (u - 1), ytop, 0, (int)blockIdx.y);
    };
    auto wait_tile = [&](int u) { tma::mbar_wait(bar0 + 8u * (u & 1), (unsigned)(((u - 1) >> 1) & 1)); };   // real tile u (1 .. U-1) is load number (u-1) >> 1 of its stage
    // synthetic edge tile: the mirror image of the neighbouring real tile (scipy 'reflect': column -1-i = column i, column W+i =
    // column W-1-i), restricted to the columns the window reaches; the rest stays 0 (= no contribution after the clamp)
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
    if constexpr (FAST) {
        // counters in shared memory, edges read through L1 (976 bytes that buy the tenth resident warp)
        if (do_hist) {
            hs.cnt = reinterpret_cast<unsigned *>(smem);
            hs.edges = const_cast<double *>(a.hist_edges);
            for (int k = lane; k < a.n_groups * a.n_vars * a.nbins; k += 32) hs.cnt[k] = 0u;
        }
        __syncthreads();
    } else {
        if (do_hist) hs = hist_setup(a, smem);          // ends with a block barrier
        else __syncthreads();
    }
    // FAST: per-thread carry slots of the aligned stores: [map][lane] float4 + [lane] mask bytes
    float4 *const cry = reinterpret_cast<float4 *>(smem + sa.off_cry) + lane;
    unsigned *const crm = reinterpret_cast<unsigned *>(smem + sa.off_cry + 3 * 32 * 16) + lane;
    if (lane == 0) issue(1);

    // per-thread line offsets of the window rows: reflected rows are rows of the same tile
    unsigned lo[BW];
    {
        const int yc = min(y, H - 1);
#pragma unroll
        for (int dy = 0; dy < BW; ++dy) lo[dy] = (unsigned)(pb_reflect(yc - HH + dy, H) - ytop) * 16u;
    }
    double tmp[NT + 1], ring[NT + 1][BW];
#pragma unroll
    for (int p = 0; p <= NT; ++p) {
        tmp[p] = 0.0;
#pragma unroll
        for (int s = 0; s < BW; ++s) ring[p][s] = 0.0;
    }
    const bool f64 = a.flags & PB_FLAG_OUT_F64;
    const bool lean = store_is_lean<1>(a, out);
    const unsigned sub2 = (unsigned)sa.isub * 0x10001u, dark2 = (unsigned)sa.idark * 0x10001u;
    const float nanf_ = __int_as_float(0x7fc00000);
    bool need_any = false;
    const long long rowp = (long long)y * W;

    wait_tile(1);
    synth(0, 1, true);

    for (int b = 1; b <= sa.b_last; ++b) {
        const int u = b >> 1, half = b & 1;
        if (half == 0 && u < U) wait_tile(u);          // tile U is synthetic (built when its stage was released)
        if (row_ok) {
            const unsigned char *rb = raw0 + (u & 1) * stage_bytes + half * 8;
            const unsigned char *rl[BW];
#pragma unroll
            for (int dy = 0; dy < BW; ++dy) rl[dy] = rb + lo[dy];
            unsigned Ts[CB];
            double S0[CB], Sc[CB], Ss[CB], Sff[CB];
#pragma unroll
            for (int c = 0; c < CB; ++c) { Ts[c] = 0u; S0[c] = 0.0; Sc[c] = 0.0; Ss[c] = 0.0; Sff[c] = 0.0; }
#pragma unroll
            for (int p = 0; p < NT; ++p) {
                unsigned s[CB] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int dy = 0; dy < BW; ++dy) {
                    const uint2 w = *reinterpret_cast<const uint2 *>(rl[dy] + p * PLB);
                    const unsigned m0 = __vmaxu2(w.x, sub2), m1 = __vmaxu2(w.y, sub2);
                    // __dp2a_lo(m, b, c) = c + m.lo16 * b.byte0 + m.hi16 * b.byte1
                    s[0] = __dp2a_lo(m0, 0x0001u, s[0]); s[1] = __dp2a_lo(m0, 0x0100u, s[1]);
                    s[2] = __dp2a_lo(m1, 0x0001u, s[2]); s[3] = __dp2a_lo(m1, 0x0100u, s[3]);
                    if constexpr (!SAME) {          // removed_intensity != 0: the threshold image clamps at the dark value alone (PC:261)
                        const unsigned d0 = __vmaxu2(w.x, dark2), d1 = __vmaxu2(w.y, dark2);
                        Ts[0] = __dp2a_lo(d0, 0x0001u, Ts[0]); Ts[1] = __dp2a_lo(d0, 0x0100u, Ts[1]);
                        Ts[2] = __dp2a_lo(d1, 0x0001u, Ts[2]); Ts[3] = __dp2a_lo(d1, 0x0100u, Ts[3]);
                    }
                }
                double q[CB];
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    if constexpr (SAME) Ts[c] += s[c];
                    const double uu = __hiloint2double(0x43300000, (int)s[c]) - sa.magic;
                    q[c] = fma(uu, sa.b_hi, uu * sa.b_lo);
                }
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const double old = c < BW ? ring[p][c < BW ? c : 0] : q[c >= BW ? c - BW : 0];
                    const double d = q[c] - old;
                    tmp[p] = tmp[p] + d;
                    const double f = fma(tmp[p], sa.b_hi, tmp[p] * sa.b_lo);
                    S0[c] = S0[c] + f;
                    Sc[c] = Sc[c] + f * bs.c2[p];
                    Ss[c] = Ss[c] + f * bs.s2[p];
                    Sff[c] = fma(f, f, Sff[c]);                    // screening only
                }
#pragma unroll
                for (int s2 = 0; s2 < BW; ++s2) ring[p][s2] = q[CB - BW + s2];
            }
            // the total-intensity line (PC:260-264): the same two passes on the integer sums over the angles
            double Iv[CB];
            {
                double q[CB];
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const double uu = __hiloint2double(0x43300000, (int)Ts[c]) - sa.magicT;
                    q[c] = fma(uu, sa.b_hi, uu * sa.b_lo);
                }
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const double old = c < BW ? ring[NT][c < BW ? c : 0] : q[c >= BW ? c - BW : 0];
                    const double d = q[c] - old;
                    tmp[NT] = tmp[NT] + d;
                    Iv[c] = fma(tmp[NT], sa.b_hi, tmp[NT] * sa.b_lo);
                }
#pragma unroll
                for (int s2 = 0; s2 < BW; ++s2) ring[NT][s2] = q[CB - BW + s2];
            }
            // ---- the block's 4 pixels: image columns k0 .. k0 + 3
            const int k0 = CB * (b - 2) - RX;
            float r0[CB], r1[CB], r2[CB];
            unsigned mk = 0u;
#pragma unroll
            for (int c = 0; c < CB; ++c) { r0[c] = nanf_; r1[c] = nanf_; r2[c] = nanf_; }
            if constexpr (FAST) {
                // whole-image thresholding, float32 maps + mask, 16-byte aligned outputs (established by the launcher)
                const double ilow = a.roi_ilow[0];
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const int k = k0 + c;
                    if (k >= 0 && k < W && Iv[c] >= ilow) {
                        const HarmPix hp = harm_stage_a_core<1>(a, 1u, S0[c], Sc[c], Ss[c], 0.0, 0.0, Sff[c]);
                        if (hp.need) need_any = true;        // decided by k_march_tma's redo pass (exact loop of PP:2990-3000)
                        else {
                            const PixOut o = harm_stage_b<1, 1>(a, hs, do_hist, hp);
                            r0[c] = (float)o.v0; r1[c] = (float)o.v1; r2[c] = (float)o.a0m;
                            mk |= (o.m ? 1u : 0u) << (8 * c);
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < CB; ++c) {
                    const int k = k0 + c;
                    if (k < 0 || k >= W) continue;
                    // the thread's 4 sets of sums live in registers: pick set c without indexing
                    const double s0 = c == 0 ? S0[0] : c == 1 ? S0[1] : c == 2 ? S0[2] : S0[3];
                    const double sc = c == 0 ? Sc[0] : c == 1 ? Sc[1] : c == 2 ? Sc[2] : Sc[3];
                    const double ss = c == 0 ? Ss[0] : c == 1 ? Ss[1] : c == 2 ? Ss[2] : Ss[3];
                    const double sf = c == 0 ? Sff[0] : c == 1 ? Sff[1] : c == 2 ? Sff[2] : Sff[3];
                    const double I = c == 0 ? Iv[0] : c == 1 ? Iv[1] : c == 2 ? Iv[2] : Iv[3];
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
            if constexpr (FAST) {
                // 16-byte aligned group g = the previous block's last three pixels (carried in the thread's shared-memory slots) + this
                // block's first one
                const float4 c0 = cry[0], c1 = cry[32], c2 = cry[64];
                const unsigned cm = crm[0];
                cry[0] = make_float4(r0[1], r0[2], r0[3], 0.f);
                cry[32] = make_float4(r1[1], r1[2], r1[3], 0.f);
                cry[64] = make_float4(r2[1], r2[2], r2[3], 0.f);
                crm[0] = mk >> 8;
                const int g = b - 3;
                if (g >= 0 && CB * g < W) {
                    const long long p = rowp + CB * g;
                    *reinterpret_cast<float4 *>(reinterpret_cast<float *>(out.var[0]) + p) = make_float4(c0.x, c0.y, c0.z, r0[0]);
                    *reinterpret_cast<float4 *>(reinterpret_cast<float *>(out.var[1]) + p) = make_float4(c1.x, c1.y, c1.z, r1[0]);
                    *reinterpret_cast<float4 *>(reinterpret_cast<float *>(out.intmap) + p) = make_float4(c2.x, c2.y, c2.z, r2[0]);
                    *reinterpret_cast<unsigned *>(out.mask + p) = cm | (mk << 24);
                }
            }
        }
        if (half == 1) {
            // the warp has read tile u for the last time: its stage takes tile u + 2
            __syncwarp();
            if (u + 2 < U) {
                if (lane == 0) issue(u + 2);
            } else if (u + 2 == U) {
                wait_tile(U - 1);
                synth(U, U - 1, false);
            }
        }
    }
    if (sa.redo) {
        const unsigned nb = __ballot_sync(0xffffffffu, need_any);
        if (lane == 0) {
            unsigned char *f = sa.redo + (long long)blockIdx.y * sa.nblk16 + 2 * blockIdx.x;
            f[0] = (nb & 0xffffu) ? 1 : 0;
            if (2 * (int)blockIdx.x + 1 < sa.nblk16) f[1] = (nb >> 16) ? 1 : 0;
        }
    }
    if (do_hist) hist_flush(a, hs, out.hist);
}

// ---------------------------------------------------------------------------------------------------
template <int BW, int NT, bool SAME, bool FAST>
cudaError_t launch(const PixArgs &a, const Basis1 &bs, int n_stacks, unsigned char *redo, cudaStream_t st) {
    constexpr int TR = ROWS + BW - 1;
    StripArgs sa;
    size_t off = FAST ? (((size_t)4 * a.n_groups * a.n_vars * a.nbins + 15) & ~(size_t)15) : tma::hist_bytes_tma(a);
    sa.off_bar = (unsigned)off; off += 16;
    sa.off_cry = (unsigned)off;
    if (FAST) off += 3 * 32 * 16 + 32 * 4;
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
    sa.nblk16 = (a.H + 15) / 16;
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
    auto kern = k_strip<BW, NT, SAME, FAST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off);
    if (e != cudaSuccess) return e;
    // shared memory for as many CTAs as the register file holds; what is left stays L1 (the few spilled values live there)
    const int ctas = 65536 / (32 * PB_STRIP_MAXREG);
    static const int carve = [] { const char *c = getenv("PB_STRIP_CARVE"); return c ? atoi(c) : -1; }();
    int pct = carve >= 0 ? carve : (int)((ctas * (off + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e != cudaSuccess) return e;
    dim3 grid((a.H + ROWS - 1) / ROWS, n_stacks);
    kern<<<grid, 32, off, st>>>(map, a, bs, sa);
    return cudaGetLastError();
}

}  // namespace strip
