// pb_pixel.cu -- fused per-pixel analysis kernels (no spatial coupling): harmonic projection, fit/chi2
// decisions, disk-cone LUT inversion, 4-harmonic outputs, 4POLAR inversion, threshold/ROI mask and
// per-CTA shared-memory histograms.
//
// Used (a) as the FUSED single-pass path when no binning is requested (raw uint8/uint16/float stack in,
// stack read once, maps + histograms out) and (b) as the last stage of the unfused path (double field in).
// Replaces compute_fields, PyPOLAR.py:2981-3078, plus the mask logic of PP:2891/PP:2967 and the
// histogram of pypolar_classes.py:340-369. Arithmetic order follows oracle/contract.c.
#include <cstdlib>
#include <type_traits>
#include "pb_pixel4.cuh"

namespace {

// raw packed loads of PX consecutive integer elements (no unpacking: fewer live registers while in flight)
template <typename UT, int PX> struct RawLoad;
template <typename UT> struct RawLoad<UT, 1> {
    static constexpr int WORDS = 1;
    __device__ static __forceinline__ void ld(const UT *p, unsigned *w) { w[0] = (unsigned)__ldg(p); }
    __device__ static __forceinline__ unsigned get(const unsigned *w, int) { return w[0]; }
};
template <> struct RawLoad<uint16_t, 4> {
    static constexpr int WORDS = 2;
    __device__ static __forceinline__ void ld(const uint16_t *p, unsigned *w) {
        const uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
        w[0] = r.x; w[1] = r.y;
    }
    __device__ static __forceinline__ unsigned get(const unsigned *w, int k) { return (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu); }
};
template <> struct RawLoad<uint8_t, 4> {
    static constexpr int WORDS = 1;
    __device__ static __forceinline__ void ld(const uint8_t *p, unsigned *w) { w[0] = __ldg(reinterpret_cast<const unsigned *>(p)); }
    __device__ static __forceinline__ unsigned get(const unsigned *w, int k) { return (w[0] >> (8 * k)) & 0xffu; }
};

#ifndef PB_PIXEL_CH
#define PB_PIXEL_CH 6
#endif
constexpr int CH = PB_PIXEL_CH;   // angle planes in flight per thread (depth of the rotating load queue)

template <typename UT, int PX> struct IntState {
    double Sc[PX], Ss[PX], Sc4[PX], Ss4[PX];
    unsigned S0[PX], Ii[PX];
    unsigned long long Sff[PX];
};

// one plane of the projection on the integer fast path: per pixel-angle one unpack, one fused subtract/clamp, integer S0 and
// sum f^2, and per harmonic component ONE DFMA (the rounded product RN(f c) = fma(2^52+f, c, -(2^52 c))) plus ONE DADD (the
// reference's separately rounded accumulation, PP:2989 / einsum)
template <typename UT, int NH, int PX, bool SEP, typename Basis>
__device__ __forceinline__ void compute_one(IntState<UT, PX> &s, const unsigned *w, const Basis &bs, int i, int isub, int idark) {
    const double c2 = bs.c2[i], s2 = bs.s2[i], kc2 = bs.kc2[i], ks2 = bs.ks2[i];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const int v = (int)RawLoad<UT, PX>::get(w, k);
        const int f = max(v - isub, 0);
        if constexpr (SEP) s.Ii[k] += (unsigned)max(v - idark, 0);
        s.S0[k] += (unsigned)f;
        s.Sff[k] += (unsigned long long)((unsigned)f) * (unsigned)f;
        const double X = __hiloint2double(0x43300000, f);
        s.Sc[k] = s.Sc[k] + fma(X, c2, kc2);
        s.Ss[k] = s.Ss[k] + fma(X, s2, ks2);
        if constexpr (NH == 2) {
            s.Sc4[k] = s.Sc4[k] + fma(X, bs.c4[i], bs.kc4[i]);
            s.Ss4[k] = s.Ss4[k] + fma(X, bs.s4[i], bs.ks4[i]);
        }
    }
}

// vector stores of 4 consecutive pixels
__device__ __forceinline__ void store4(void *base, long long p, double x0, double x1, double x2, double x3, bool f64) {
    if (!base) return;
    if (f64) {
        double2 *d = reinterpret_cast<double2 *>(reinterpret_cast<double *>(base) + p);
        d[0] = make_double2(x0, x1); d[1] = make_double2(x2, x3);
    } else {
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(base) + p) = make_float4((float)x0, (float)x1, (float)x2, (float)x3);
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused per-pixel kernel, harmonic modalities (1PF: NH=1, CARS/SRS/SHG/2PF: NH=2). One thread = PX
// consecutive pixels; angle planes are streamed CH at a time with the next chunk's loads in flight.
//   MODE 1: integer fast path -- uint8/uint16 stack, integer-valued dark and dark+removed, no sqrt.
//   MODE 2: the same with removed_intensity != 0 (the threshold image then needs its own clamp with `dark` alone, PC:261).
//   MODE 0: generic -- fp64 pre-processing (float input, fractional dark, CARS sqrt, or an already binned double field).
template <typename T, int NH, int PX, int MODE>
__global__ void __launch_bounds__(128, NH == 1 ? PB_PIXEL_MIN_BLOCKS : 5) k_pixel_harm(PixArgs a, typename BasisSel<NH>::type bs) {
    extern __shared__ __align__(16) unsigned char smem[];
    const long long P = a.P;
    const long long p0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * PX;
    const bool active = p0 < P;
    const T *base = reinterpret_cast<const T *>(a.stack) + (long long)blockIdx.y * a.stack_stride + p0;
    double Sc[PX], Ss[PX], Sc4[PX], Ss4[PX], S0d[PX], Sffd[PX], Id[PX];
    // the projection starts at once: the output pointers and the histogram area are not needed before the epilogue (their
    // global loads and the set-up barrier would otherwise sit in front of the first stack loads, and the pointers would be
    // spilled across the loop)
    if (active) {
        const int N = a.N;
        if constexpr (MODE >= 1) {
            constexpr bool SEP = MODE == 2;
            using UT = typename std::conditional<std::is_same<T, uint8_t>::value, uint8_t, uint16_t>::type;
            const UT *ptr = reinterpret_cast<const UT *>(base);
            const int isub = (int)a.sub, idark = (int)a.dark;
            IntState<UT, PX> s;
#pragma unroll
            for (int k = 0; k < PX; ++k) { s.Sc[k] = s.Ss[k] = s.Sc4[k] = s.Ss4[k] = 0.0; s.S0[k] = s.Ii[k] = 0u; s.Sff[k] = 0ull; }
            // rotating register queue of CH planes: plane i is consumed from slot i % CH and the slot is at once reloaded with
            // plane i + CH, so CH loads stay in flight with CH (not 2 CH) buffers -- with ping-pong buffers the compiler ran out
            // of registers and spilled the data of loads still in flight, which made every prefetch synchronous
            unsigned buf[CH][RawLoad<UT, PX>::WORDS];
#pragma unroll
            for (int j = 0; j < CH; ++j)
                if (j < N) { RawLoad<UT, PX>::ld(ptr, buf[j]); ptr += P; }
            int i = 0;
            for (; i + CH <= N; i += CH) {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    compute_one<UT, NH, PX, SEP>(s, buf[j], bs, i + j, isub, idark);
                    if (i + j + CH < N) { RawLoad<UT, PX>::ld(ptr, buf[j]); ptr += P; }
                }
            }
#pragma unroll
            for (int j = 0; j < CH; ++j)                               // the N % CH planes left in the queue
                if (i + j < N) compute_one<UT, NH, PX, SEP>(s, buf[j], bs, i + j, isub, idark);
#pragma unroll
            for (int k = 0; k < PX; ++k) {
                Sc[k] = s.Sc[k]; Ss[k] = s.Ss[k]; Sc4[k] = s.Sc4[k]; Ss4[k] = s.Ss4[k];
                S0d[k] = (double)s.S0[k];
                Sffd[k] = (double)s.Sff[k];
                Id[k] = SEP ? (double)s.Ii[k] : S0d[k];
            }
        } else {
            const bool pre = !a.from_field;
            const bool sep_int = pre && !a.intensity_in && (a.do_sqrt || a.sub != a.dark);
#pragma unroll
            for (int k = 0; k < PX; ++k) { S0d[k] = Sc[k] = Ss[k] = Sc4[k] = Ss4[k] = Sffd[k] = Id[k] = 0.0; }
#pragma unroll 3
            for (int i = 0; i < N; ++i) {
                double v[PX];
                VecLoad<T, PX>::ld(base + (long long)i * P, v);
                const double c2 = bs.c2[i], s2 = bs.s2[i];
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    double f = v[k];
                    if (pre) {
                        if (sep_int) Id[k] = Id[k] + pb_max0(v[k] - a.dark);
                        f = pb_max0(f - a.sub);
                        if (a.do_sqrt) f = sqrt(f);
                    }
                    S0d[k] = S0d[k] + f;
                    Sc[k] = Sc[k] + f * c2;
                    Ss[k] = Ss[k] + f * s2;
                    if constexpr (NH == 2) {
                        Sc4[k] = Sc4[k] + f * bs.c4[i];
                        Ss4[k] = Ss4[k] + f * bs.s4[i];
                    }
                    Sffd[k] = fma(f, f, Sffd[k]);   // screening only (not parity relevant)
                }
            }
            if (!sep_int) {
#pragma unroll
                for (int k = 0; k < PX; ++k) Id[k] = S0d[k];
            }
        }
    }
    // ---- Epilogue.
    // Owner thread (its PX pixels): only the threshold / ROI test (PP:2891). Pixels that pass are pushed, with their sums, into a
    // per-CTA queue in shared memory.
    // Dense pass (any thread, one queued pixel at a time): a0 / a2 / a4, the certified fit/chi2 screening and, for the rare
    // undecided pixel, the reference's fp64 loop (PP:2987-3000); normalisation, disk-cone LUT or 4-harmonic outputs, histograms
    // (PP:3017-3053, PC:340-369). On thresholded images only a fraction of the pixels gets here: taking them from the queue
    // keeps every lane of the warps that run it busy, instead of every warp paying for its few valid lanes.
    // The owner threads then write their PX consecutive pixels with vector stores.
    constexpr int NQ = 2 + 2 * NH;                                  // S0, Sc, Ss (, Sc4, Ss4), Sff
    constexpr int NR = 2 * NH + 1;                                  // v0, v1 (, v2, v3), a0m
    const int cta_px = blockDim.x * PX;
    static_assert(NR <= NQ, "the dense pass writes its results over the queue entry it consumed");
    double *qv = reinterpret_cast<double *>(smem + a.epi_smem_off);            // [NQ][cta_px] queue entries, then their results
    unsigned short *qidx = reinterpret_cast<unsigned short *>(qv + NQ * cta_px);   // [cta_px] local pixel of each queue entry
    unsigned char *rm = reinterpret_cast<unsigned char *>(qidx + cta_px);     // [cta_px] final mask of the queued pixels
    int *qcount = reinterpret_cast<int *>(rm + cta_px);
    if (threadIdx.x == 0) *qcount = 0;
    const pb_outputs out = a.outs[blockIdx.y];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    HistSmem hs;
    hs.cnt = nullptr; hs.edges = nullptr;
    if (do_hist) hs = hist_setup(a, smem);          // ends with a barrier
    else __syncthreads();
    const bool f64 = a.flags & PB_FLAG_OUT_F64;
    const int lane = threadIdx.x & 31;
    int qslot[PX];                                                  // queue entry of pixel k of this thread, or -1
    double Iv[PX];
    // queue slots. AGG (4-harmonic kernels, 96 registers): one shared-memory atomic per warp for all its PX x 32 pixels, entries
    // numbered pixel slot by pixel slot; otherwise (1PF, 64 registers: the extra live flags cost spills in the projection loop,
    // measured 843 -> 694 k Mpx-angles/s) one atomic per pixel slot.
    constexpr bool AGG = NH == 2;
    bool pm[PX];
    unsigned ball[PX];
    int e0 = 0;
    if constexpr (AGG) {
        int wcount = 0;
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            Iv[k] = 0.0;
            pm[k] = false;
            if (active) {
                uint32_t bits;
                Iv[k] = a.intensity_in ? __ldg(a.intensity_in + p0 + k) : Id[k];
                pm[k] = harm_threshold(a, p0 + k, Iv[k], bits);
            }
            ball[k] = __ballot_sync(0xffffffffu, pm[k]);
            wcount += __popc(ball[k]);
        }
        if (lane == 0 && wcount) e0 = atomicAdd(qcount, wcount);
        e0 = __shfl_sync(0xffffffffu, e0, 0);
    }
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        if constexpr (!AGG) {
            Iv[k] = 0.0;
            pm[k] = false;
            if (active) {
                uint32_t bits;
                Iv[k] = a.intensity_in ? __ldg(a.intensity_in + p0 + k) : Id[k];
                pm[k] = harm_threshold(a, p0 + k, Iv[k], bits);
            }
            ball[k] = __ballot_sync(0xffffffffu, pm[k]);
            e0 = 0;
            if (lane == 0 && ball[k]) e0 = atomicAdd(qcount, __popc(ball[k]));
            e0 = __shfl_sync(0xffffffffu, e0, 0);
        }
        qslot[k] = -1;
        if (pm[k]) {
            const int e = e0 + __popc(ball[k] & ((1u << lane) - 1u));
            qv[e] = S0d[k]; qv[cta_px + e] = Sc[k]; qv[2 * cta_px + e] = Ss[k];
            if constexpr (NH == 2) { qv[3 * cta_px + e] = Sc4[k]; qv[4 * cta_px + e] = Ss4[k]; }
            qv[(NQ - 1) * cta_px + e] = Sffd[k];
            qidx[e] = (unsigned short)(threadIdx.x * PX + k);
            qslot[k] = e;
        }
        if constexpr (AGG) e0 += __popc(ball[k]);
    }
    __syncthreads();
    {
        const int qn = *qcount;
        const long long pcta = (long long)blockIdx.x * cta_px;
        const T *cta_stack = reinterpret_cast<const T *>(a.stack) + (long long)blockIdx.y * a.stack_stride + pcta;
        for (int e = threadIdx.x; e < qn; e += blockDim.x) {
            const int lp = qidx[e];
            const uint32_t bits = (a.simple || !a.roi_bits) ? 1u : __ldg(a.roi_bits + pcta + lp);
            double c4s = 0.0, s4s = 0.0;
            if constexpr (NH == 2) { c4s = qv[3 * cta_px + e]; s4s = qv[4 * cta_px + e]; }
            HarmPix h = harm_stage_a_core<NH>(a, bits, qv[e], qv[cta_px + e], qv[2 * cta_px + e], c4s, s4s, qv[(NQ - 1) * cta_px + e]);
            if (h.need) {
                // rare path: the reference's fp64 fit/chi2 loop for an undecided pixel (stack re-read, L1/L2 hits)
                ExactAcc ex;
                ex.init();
                const T *pp = cta_stack + lp;
#pragma unroll 1
                for (int i = 0; i < a.N; ++i) {
                    double f = (double)__ldg(pp + (long long)i * P);
                    if (!a.from_field) { f = pb_max0(f - a.sub); if (a.do_sqrt) f = sqrt(f); }
                    double c4 = 0.0, s4 = 0.0;
                    if constexpr (NH == 2) { c4 = bs.c4[i]; s4 = bs.s4[i]; }
                    ex.template step<NH>(h, f, bs.c2[i], bs.s2[i], c4, s4);
                }
                h.m = ex.decide(a);
                h.need = false;
            }
            const PixOut o = harm_stage_b<NH>(a, hs, do_hist, h);
            qv[e] = o.v0; qv[cta_px + e] = o.v1;
            if constexpr (NH == 2) { qv[2 * cta_px + e] = o.v2; qv[3 * cta_px + e] = o.v3; }
            qv[(NR - 1) * cta_px + e] = o.a0m;
            rm[e] = (unsigned char)o.m;
        }
    }
    __syncthreads();
    if (active) {
        PixOut o[PX];
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            o[k].v0 = o[k].v1 = o[k].v2 = o[k].v3 = o[k].a0m = pb_nan();
            o[k].m = false;
            if (qslot[k] >= 0) {
                const int e = qslot[k];
                o[k].v0 = qv[e]; o[k].v1 = qv[cta_px + e];
                if constexpr (NH == 2) { o[k].v2 = qv[2 * cta_px + e]; o[k].v3 = qv[3 * cta_px + e]; }
                o[k].a0m = qv[(NR - 1) * cta_px + e];
                o[k].m = rm[e] != 0;
            }
        }
        if constexpr (PX == 4) {
            store4(out.var[0], p0, o[0].v0, o[1].v0, o[2].v0, o[3].v0, f64);
            store4(out.var[1], p0, o[0].v1, o[1].v1, o[2].v1, o[3].v1, f64);
            if constexpr (NH == 2) {
                store4(out.var[2], p0, o[0].v2, o[1].v2, o[2].v2, o[3].v2, f64);
                if (a.method == PB_SHG) store4(out.var[3], p0, o[0].v3, o[1].v3, o[2].v3, o[3].v3, f64);
            }
            store4(out.intmap, p0, o[0].a0m, o[1].a0m, o[2].a0m, o[3].a0m, f64);
            if (out.mask) *reinterpret_cast<uchar4 *>(out.mask + p0) = make_uchar4(o[0].m, o[1].m, o[2].m, o[3].m);
            if (out.intensity && !a.intensity_in) {
                double2 *d = reinterpret_cast<double2 *>(reinterpret_cast<double *>(out.intensity) + p0);
                d[0] = make_double2(Iv[0], Iv[1]); d[1] = make_double2(Iv[2], Iv[3]);
            }
            if (out.rho_angle || (out.rho_contour && a.edge)) {
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    const long long p = p0 + k;
                    if (out.rho_angle) store1(out.rho_angle, p, (o[k].v0 == o[k].v0) ? pb_pymod(o[k].v0 - a.ref_angle, 180.0) : pb_nan(), f64);
                    if (out.rho_contour && a.edge) {
                        const double e = __ldg(a.edge + p);
                        store1(out.rho_contour, p, (pb_finite(o[k].v0) && pb_finite(e)) ? pb_pymod(o[k].v0 - e, 180.0) : pb_nan(), f64);
                    }
                }
            }
        } else {
            store_pixel<NH>(a, out, p0, o[0], Iv[0], f64);
        }
    }
    if (do_hist) hist_flush(a, hs, out.hist);
}

template <typename T, bool FAST>
__global__ void __launch_bounds__(256) k_pixel_4polar(PixArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int stack = blockIdx.y;
    const pb_outputs out = a.outs[stack];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    HistSmem hs;
    hs.cnt = nullptr; hs.edges = nullptr;
    if (do_hist) hs = hist_setup(a, smem);
    const long long P = a.P;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
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
        // raw-stack intensity sums planes in acquisition order; addition order matters only for non-integers
        const double I = a.intensity_in ? __ldg(a.intensity_in + p) : Isum;
        p4_pixel<FAST>(a, hs, do_hist, out, p, f, I);
    }
    if (do_hist) hist_flush(a, hs, out.hist);
}

// ---------------------------------------------------------------------------------------------------
// Disk-dependent tail of the 1PF analysis on projected maps (calibration sweep): one thread = one pixel.
__global__ void __launch_bounds__(256) k_lut_apply(PixArgs a, const double *__restrict__ a0, const double *__restrict__ a2r,
                                                   const double *__restrict__ a2i, const uint8_t *__restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem[];
    const pb_outputs out = a.outs[blockIdx.y];
    const bool do_hist = !(a.flags & PB_FLAG_NO_HIST) && a.n_vars > 0 && out.hist;
    HistSmem hs;
    hs.cnt = nullptr; hs.edges = nullptr;
    if (do_hist) hs = hist_setup(a, smem);
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.P) {
        const long long q = (long long)blockIdx.y * a.P + p;
        HarmPix h;
        h.a0 = a0[q]; h.a2r = a2r[q]; h.a2i = a2i[q]; h.a4r = h.a4i = 0.0;
        h.bits = a.roi_bits ? __ldg(a.roi_bits + p) : 1u;
        h.m = mask[q] != 0; h.need = false;
        const PixOut o = harm_stage_b<1>(a, hs, do_hist, h);
        store_pixel<1>(a, out, p, o, 0.0, a.flags & PB_FLAG_OUT_F64, store_is_lean<1>(a, out));
    }
    if (do_hist) hist_flush(a, hs, out.hist);
}

template <typename T, int NH>
cudaError_t launch_harm_t(const PixArgs &a_in, const typename BasisSel<NH>::type &bs, int n_stacks, size_t smem, cudaStream_t st) {
    const int threads = 128;
    PixArgs a = a_in;
    // PX = 4 needs every plane (and every stack) to start on a 4-element boundary
    const bool vec = (a.P % 4 == 0) && (a.stack_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.stack) % (4 * sizeof(T))) == 0) && a.out_aligned;
    // epilogue queue (results overwrite the entries) behind the histograms: per pixel of the CTA 2 + 2 NH doubles, a 16-bit
    // index and a mask byte; then the queue counter
    smem = (smem + 15) & ~(size_t)15;
    a.epi_smem_off = (int)smem;
    smem += (size_t)threads * (vec ? 4 : 1) * (8 * (2 + 2 * NH) + 3) + 16;
    // integer fast path: integer stack, integer-valued offsets that fit the element range, no sqrt
    constexpr bool is_int = std::is_same<T, uint8_t>::value || std::is_same<T, uint16_t>::value;
    const bool fast = is_int && !a.from_field && !a.do_sqrt && !a.intensity_in && a.sub == floor(a.sub) && a.dark == floor(a.dark) &&
                      a.sub >= 0 && a.sub <= 65536.0 && a.dark >= 0 && a.dark <= 65536.0;
    const long long nthr = vec ? a.P / 4 : a.P;
    dim3 grid((unsigned)((nthr + threads - 1) / threads), n_stacks);
    auto launch = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {      // many ROI groups x bins: opt in to a large dynamic shared memory area
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, threads, smem, st>>>(a, bs);
        return cudaGetLastError();
    };
    if constexpr (is_int) {
        if (fast && a.sub == a.dark) return vec ? launch(k_pixel_harm<T, NH, 4, 1>) : launch(k_pixel_harm<T, NH, 1, 1>);
        if (fast) return vec ? launch(k_pixel_harm<T, NH, 4, 2>) : launch(k_pixel_harm<T, NH, 1, 2>);
    }
    return vec ? launch(k_pixel_harm<T, NH, 4, 0>) : launch(k_pixel_harm<T, NH, 1, 0>);
}

}  // namespace

size_t pb_hist_smem_bytes(int n_vars, int nbins, int n_groups) {
    return sizeof(double) * n_vars * (nbins + 1) + sizeof(unsigned) * n_groups * n_vars * nbins + 16;
}

cudaError_t pb_launch_pixel_harm(const PixArgs &a, const Basis2 &bs, int dtype, int n_stacks, cudaStream_t st) {
    size_t smem = pb_hist_smem_bytes(a.n_vars, a.nbins, a.n_groups);
    const bool one = a.method == PB_1PF;
    Basis1 b1;
    if (one) { memcpy(b1.c2, bs.c2, sizeof(b1.c2)); memcpy(b1.s2, bs.s2, sizeof(b1.s2)); memcpy(b1.kc2, bs.kc2, sizeof(b1.kc2)); memcpy(b1.ks2, bs.ks2, sizeof(b1.ks2)); }
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

cudaError_t pb_launch_lut_apply(const PixArgs &a, const double *a0, const double *a2r, const double *a2i, const uint8_t *mask, int n_stacks,
                                cudaStream_t st) {
    const size_t smem = pb_hist_smem_bytes(a.n_vars, a.nbins, a.n_groups);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_lut_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)((a.P + 255) / 256), n_stacks);
    k_lut_apply<<<grid, 256, smem, st>>>(a, a0, a2r, a2i, mask);
    return cudaGetLastError();
}

cudaError_t pb_launch_pixel_4polar(const PixArgs &a, int dtype, int n_stacks, cudaStream_t st) {
    size_t smem = pb_hist_smem_bytes(a.n_vars, a.nbins, a.n_groups);
    const int threads = 256;
    dim3 grid((unsigned)((a.P + threads - 1) / threads), n_stacks);
    auto launch = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, threads, smem, st>>>(a);
        return cudaGetLastError();
    };
    // float32 maps may take the certified fast evaluation (same mask, same histograms); double maps: the reference's statements
    // throughout. Measured on B200 (4POLAR 3D 4x2048x2048, 32 stacks): exact 117 k, fast 102 k Mpx-angles/s -- the kernel is
    // issue-bound, the full-precision float32 library functions plus the certification cost as many instructions as the fp64
    // statements they replace, and the larger code no longer fits the instruction cache. Hence opt-in: PB_P4_FAST=1.
    static const bool fast_on = [] { const char *e = getenv("PB_P4_FAST"); return e && atoi(e) != 0; }();
    const bool fast = fast_on && !(a.flags & PB_FLAG_OUT_F64);
    switch (dtype) {
    case PB_U8: return fast ? launch(k_pixel_4polar<uint8_t, true>) : launch(k_pixel_4polar<uint8_t, false>);
    case PB_U16: return fast ? launch(k_pixel_4polar<uint16_t, true>) : launch(k_pixel_4polar<uint16_t, false>);
    case PB_F32: return fast ? launch(k_pixel_4polar<float, true>) : launch(k_pixel_4polar<float, false>);
    case PB_F64: return fast ? launch(k_pixel_4polar<double, true>) : launch(k_pixel_4polar<double, false>);
    }
    return cudaErrorInvalidValue;
}
