// pb_march.cu -- planning and dispatch of the fused row-march kernel (pb_march_impl.cuh). The kernel is instantiated in
// four translation units (pb_march_u16_1.cu, ... one per stack type x number of harmonics) so that they compile in parallel.
#include "pb_march_impl.cuh"

cudaError_t pb_march_u16_nh1(const PixArgs &a, const Basis1 &bs, const pb_params *p, int n_stacks, cudaStream_t st);
cudaError_t pb_march_u16_nh2(const PixArgs &a, const Basis2 &bs, const pb_params *p, int n_stacks, cudaStream_t st);
cudaError_t pb_march_u8_nh1(const PixArgs &a, const Basis1 &bs, const pb_params *p, int n_stacks, cudaStream_t st);
cudaError_t pb_march_u8_nh2(const PixArgs &a, const Basis2 &bs, const pb_params *p, int n_stacks, cudaStream_t st);
cudaError_t pb_march_tma_nh1(const PixArgs &a, const Basis1 &bs, int bw, int n_stacks, cudaStream_t st);   // pb_march_tma_1.cu
cudaError_t pb_march_tma_nh2(const PixArgs &a, const Basis2 &bs, int bw, int n_stacks, cudaStream_t st);   // pb_march_tma_2.cu
cudaError_t pb_march_tma_4polar(const PixArgs &a, int bw, int n_stacks, cudaStream_t st);                     // pb_march_tma_4p.cu
cudaError_t pb_strip_nh1(const PixArgs &a, const Basis1 &bs, int bw, int n_stacks, unsigned char *redo, cudaStream_t st);   // pb_strip_1.cu
cudaError_t pb_march_tma_redo_nh1_bw3_n18(const PixArgs &a, const Basis1 &bs, int n_stacks, const unsigned char *redo, cudaStream_t st);

// The TMA kernel (pb_march_tma.cuh) takes the common case: uint16 stacks, square windows 2..8, rows that are a multiple of 16 bytes
// and a 16-byte aligned stack (the tensor map's stride / address rules). PB_MARCH_TMA=0 keeps the register-staged loader.
static bool tma_eligible(const PixArgs &a, const pb_params *p) {
    static const bool on = [] { const char *e = getenv("PB_MARCH_TMA"); return !e || atoi(e) != 0; }();
    if (!on || p->dtype != PB_U16 || p->bin_h != p->bin_w || p->bin_w < 2 || p->bin_w > 8) return false;
    if (a.W % 8 != 0 || a.W < 16 || a.H < 1 || (reinterpret_cast<uintptr_t>(a.stack) & 15) != 0 || (a.stack_stride % 8) != 0) return false;
    return true;
}

// The strip-march kernel (pb_strip.cuh: one thread per image row, every plane's filter state in registers) takes the reference's
// standard 1PF acquisition -- 18 angles, 3x3 window -- when the TMA rules hold; pixels it cannot decide go to k_march_tma's redo pass.
// Opt-in (PB_FLAG_STRIP_MARCH, or PB_STRIP=1 for every eligible call) while it is slower than k_march_tma; PB_FLAG_EXACT_CHI2 (every
// pixel through the exact loop) keeps k_march_tma.
static bool strip_eligible(const PixArgs &a, const pb_params *p, const unsigned char *redo) {
    static const bool env_on = [] { const char *e = getenv("PB_STRIP"); return e && atoi(e) != 0; }();
    return (env_on || (p->flags & PB_FLAG_STRIP_MARCH)) && redo && p->method == PB_1PF && p->N == 18 && p->bin_w == 3 && p->bin_h == 3 && a.W >= 16 && a.H >= 2 &&
           !(p->flags & PB_FLAG_EXACT_CHI2) && tma_eligible(a, p);
}

// The march kernel covers: harmonic modalities without sqrt (1PF, SRS, SHG, 2PF), integer stacks, integer-valued
// dark / dark+removed (so that the H pass sums integers), bin_w <= 8, N + 1 <= 32 chain warps. Everything else
// (CARS, 4POLAR, float stacks, fractional dark, wider windows, more angles) takes the unfused kernels.
bool pb_march_supported(const pb_params *p) {
    if (p->method == PB_CARS) return false;
    if (p->method >= PB_4POLAR_2D) {
        // 4POLAR with binning: the TMA kernel only (uint16 stack of 4 planes, square window 2..8, integer dark); pb_analyze falls back to
        // the unfused kernels when the stack's address / width do not satisfy the tensor-map rules
        const double sub = p->dark + p->removed_intensity;
        return p->dtype == PB_U16 && p->N == 4 && p->bin_h == p->bin_w && p->bin_w >= 2 && p->bin_w <= 8 && sub == floor(sub) && p->dark == floor(p->dark) &&
               sub >= 0 && sub <= 65535.0 && p->dark >= 0 && p->dark <= 65535.0 && (long long)p->H * p->W < 2147483647LL;
    }
    if (p->dtype != PB_U8 && p->dtype != PB_U16) return false;
    const double sub = p->dark + p->removed_intensity;
    if (sub != floor(sub) || p->dark != floor(p->dark) || sub < 0 || sub > 65535.0 || p->dark < 0 || p->dark > 65535.0) return false;
    if (p->bin_w > 8 || p->bin_h > 32) return false;
    if (p->N + 1 > 128 || p->N < 1) return false;
    if ((long long)p->H * p->W >= 2147483647LL) return false;             // 32-bit in-plane offsets in the loader
    const int bh = p->bin_h > 1 ? p->bin_h : 1, bw = p->bin_w > 1 ? p->bin_w : 1;
    const int mr = march_rows(p->N);
    if (mr != 32 && (p->dtype != PB_U16 || bh != bw || bw < 2)) return false;   // many angles: uint16, square windows
    const int xt = bw == 1 ? 8 : bw * ((PB_MARCH_XT_MIN + bw - 1) / bw);
    const int subs = 32 / mr;
    int cw = (((p->N + 1 + PL - 1) / PL) + subs - 1) / subs;
    if (cw * 32 < mr * xt) cw = (mr * xt + 31) / 32;
    const int npos = (mr + bh - 1) * xt;
    int lw = (npos + 32 * QMAX - 1) / (32 * QMAX);
    if (lw < 2) lw = 2;
    if (32 * (cw + lw) > 1024) return false;
    const int rp = 2 * (((xt + 1) / 2) | 1), fp = (xt % 2) ? xt : xt + 1;
    const size_t smem = 2 * ((size_t)p->N * march_plane_words(mr + bh - 1, rp / 2, mr) * 4 + 16) + 2 * ((size_t)npos * 4 + 16) + (size_t)(p->N + 1) * mr * fp * 8 + 16 * 1024;
    return smem <= 227 * 1024;
}

cudaError_t pb_launch_march(const PixArgs &a, const Basis2 &bs, const pb_params *p, int n_stacks, cudaStream_t st, long *launches,
                            unsigned char *redo) {
    cudaError_t e;
    const bool tma_ok = tma_eligible(a, p);
    if (p->method >= PB_4POLAR_2D) {
        if (!tma_ok) return cudaErrorInvalidConfiguration;
        e = pb_march_tma_4polar(a, p->bin_w, n_stacks, st);
        if (e == cudaSuccess && launches) ++*launches;
        return e;
    }
    if (p->method == PB_1PF) {
        Basis1 b1;
        memcpy(b1.c2, bs.c2, sizeof(b1.c2)); memcpy(b1.s2, bs.s2, sizeof(b1.s2));
        memcpy(b1.kc2, bs.kc2, sizeof(b1.kc2)); memcpy(b1.ks2, bs.ks2, sizeof(b1.ks2));
        if (strip_eligible(a, p, redo)) {
            e = pb_strip_nh1(a, b1, p->bin_w, n_stacks, redo, st);
            if (e == cudaSuccess) {
                if (launches) ++*launches;
                e = pb_march_tma_redo_nh1_bw3_n18(a, b1, n_stacks, redo, st);
                if (e == cudaSuccess && launches) ++*launches;
                return e;
            }
            if (e != cudaErrorInvalidConfiguration && e != cudaErrorNotSupported) return e;
        }
        e = tma_ok ? pb_march_tma_nh1(a, b1, p->bin_w, n_stacks, st) : cudaErrorInvalidConfiguration;
        if (e == cudaErrorInvalidConfiguration || e == cudaErrorNotSupported)      // shape outside the TMA kernel's plan: register-staged loader
            e = p->dtype == PB_U8 ? pb_march_u8_nh1(a, b1, p, n_stacks, st) : pb_march_u16_nh1(a, b1, p, n_stacks, st);
    } else {
        e = tma_ok ? pb_march_tma_nh2(a, bs, p->bin_w, n_stacks, st) : cudaErrorInvalidConfiguration;
        if (e == cudaErrorInvalidConfiguration || e == cudaErrorNotSupported)
            e = p->dtype == PB_U8 ? pb_march_u8_nh2(a, bs, p, n_stacks, st) : pb_march_u16_nh2(a, bs, p, n_stacks, st);
    }
    if (e == cudaSuccess && launches) ++*launches;
    return e;
}
