// pb_api.cu -- the C ABI of libpolarb200.so (see include/polar_b200.h): context, calibration upload,
// path planning and kernel dispatch. Host code only; all device work is in pb_pixel.cu / pb_box.cu /
// pb_march.cu / pb_reduce.cu. No CPU fallback: without a usable CUDA device every compute call fails.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pb_dev.cuh"

// kernel launchers defined in the other translation units
size_t pb_hist_smem_bytes(int n_vars, int nbins, int n_groups);
cudaError_t pb_launch_pixel_harm(const PixArgs &a, const Basis2 &bs, int dtype, int n_stacks, cudaStream_t st);
cudaError_t pb_launch_pixel_4polar(const PixArgs &a, int dtype, int n_stacks, cudaStream_t st);
cudaError_t pb_launch_lut_apply(const PixArgs &a, const double *a0, const double *a2r, const double *a2i, const uint8_t *mask, int n_stacks,
                                cudaStream_t st);
cudaError_t pb_launch_rescale(const void *a, int is_f64, long long n, unsigned long long *mm_dev, uint16_t *out, cudaStream_t st);
cudaError_t pb_launch_hwn_to_nhw(const void *in, void *out, int elem_bytes, long long P, int N, cudaStream_t st);
size_t pb_compact_ws_bytes(long long n);
cudaError_t pb_launch_compact(const void *vals, const void *filt, int is_f64, const uint8_t *sel, const uint32_t *bits, uint32_t group, long long n,
                              void *ws, void *out_f16, void *out_val, uint32_t *out_index, const unsigned long long **total_dev, cudaStream_t st);
cudaError_t pb_launch_field(const void *v, int dtype, double *out, long long P, int N, double sub, int do_sqrt, int swap12, cudaStream_t st);
cudaError_t pb_launch_isum(const void *v, int dtype, double *out, long long P, int N, double dark, cudaStream_t st);
cudaError_t pb_launch_box_y(const double *in, double *out, int planes, int H, int W, int n, cudaStream_t st);
cudaError_t pb_launch_box_x(const double *in, double *out, int planes, int H, int W, int n, cudaStream_t st);
cudaError_t pb_launch_histo(const void *vals, int is_f64, const uint8_t *sel, long long n, int is_rho, double rot_fig, int fold90,
                            const double *edges_dev, int nbins, double inv_w, int64_t *counts, cudaStream_t st);
cudaError_t pb_launch_stats(int pass, const void *vals, const void *rho, int is_f64, const uint8_t *sel, long long n, int circular,
                            double rot_fig, int rewrap, int fold90, double mean, double *acc, cudaStream_t st);
cudaError_t pb_launch_fit_maps(const double *f, const uint8_t *mask, long long P, int N, int nharm, const Basis2 &bs, double *fit,
                               double *chi2, cudaStream_t st);
cudaError_t pb_launch_dark_cells(const void *plane0, int dtype, int W, int ny, int nx, int cell, unsigned long long *sums, unsigned *counts,
                                 cudaStream_t st);
cudaError_t pb_launch_dark_cell_all(const void *stack, int dtype, long long P, int W, int N, int y0, int x0, int cell,
                                    unsigned long long *out, cudaStream_t st);
cudaError_t pb_launch_stats_batch(int pass, const pb_outputs *outs_dev, int n_stacks, const uint32_t *roi_bits, long long P, int n_vars, int f64,
                                  double rot_fig, const uint32_t *group_bits, int n_groups, int per_stack, double *acc1, double *acc2, cudaStream_t st);
bool pb_march_supported(const pb_params *p);
cudaError_t pb_launch_march(const PixArgs &a, const Basis2 &bs, const pb_params *p, int n_stacks, cudaStream_t st, long *launches, unsigned char *redo);

static std::string g_create_error;

// double-double reciprocal of a small positive integer: hi = RN(1/n), lo = RN(1/n - hi) (exact residual via fma)
void pb_ddrecip(int n, double *hi, double *lo) {
    const double d = (double)n;
    const double h = 1.0 / d;
    const double r = std::fma(-h, d, 1.0);   // exact: 1 - h*d
    *hi = h;
    *lo = r / d;
}

struct pb_ctx {
    int device = 0;
    std::string err;
    long launches = 0;
    // calibration: `lut_block` holds n_disks tables back to back (each: table | grid | (h, 1/h)); lut_dev / grid_dev / hy_dev
    // point into the selected one
    char *lut_block = nullptr;
    size_t lut_stride = 0;
    int n_disks = 0, cur_disk = 0;
    double2 *lut_dev = nullptr;
    double *grid_dev = nullptr;
    double2 *hy_dev = nullptr;
    int lut_n = 0;
    double grid_lo = -1.0, grid_hi = 1.0;
    double invk[16] = {0};
    int invk_rows = 0;
    Basis2 basis;
    int basis_n = 0;
    bool basis_orthogonal = false;
    // histograms
    double *edges_dev = nullptr;
    int n_vars = 0, nbins = 0;
    unsigned is_rho_bits = 0;
    double inv_w[PB_MAX_VARS] = {0};
    double hist_lo[PB_MAX_VARS] = {0}, hist_hi[PB_MAX_VARS] = {0};
    // ROIs
    int n_rois = 0, n_groups = 1;
    double roi_ilow[PB_MAX_ROIS] = {0};
    uint32_t group_bits[PB_MAX_ROIS] = {0xffffffffu};
    // workspace (unfused path and host API)
    void *ws[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t ws_bytes[6] = {0, 0, 0, 0, 0, 0};
    pb_outputs *outs_dev = nullptr;   // ring of output descriptors
    int outs_cap = 0, outs_pos = 0;
    double *acc_dev = nullptr;        // stats accumulators
    cudaStream_t hs[2] = {nullptr, nullptr};
    int64_t *hist_pinned = nullptr;   // pinned staging of the per-stack histograms of pb_analyze_host
    size_t hist_pinned_n = 0;
    cudaEvent_t ws_event = nullptr;   // recorded after the last kernel that used the shared workspace
    // strip-march redo flags (one byte per 16-row block per stack): a ring of slots, one per call, so that calls in flight on
    // different streams (pb_analyze_host's two pipelines) never share one
    unsigned char *redo_dev = nullptr;
    size_t redo_slot_bytes = 0;
    int redo_pos = 0;
};
static const int PB_REDO_SLOTS = 8;

// flags buffer of one fused-march call (n_stacks x ceil(H/16) bytes); grows with a device synchronisation
static unsigned char *redo_slot(pb_ctx *ctx, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > ctx->redo_slot_bytes) {
        if (cudaDeviceSynchronize() != cudaSuccess) return nullptr;
        if (ctx->redo_dev) cudaFree(ctx->redo_dev);
        ctx->redo_dev = nullptr; ctx->redo_slot_bytes = 0;
        if (cudaMalloc(&ctx->redo_dev, bytes * PB_REDO_SLOTS) != cudaSuccess) return nullptr;
        ctx->redo_slot_bytes = bytes;
    }
    ctx->redo_pos = (ctx->redo_pos + 1) % PB_REDO_SLOTS;
    return ctx->redo_dev + (size_t)ctx->redo_pos * ctx->redo_slot_bytes;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return PB_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

static int fail(pb_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

static int ws_reserve(pb_ctx *ctx, int slot, size_t bytes) {
    if (ctx->ws_bytes[slot] >= bytes) return PB_OK;
    if (ctx->ws[slot]) { CK(cudaDeviceSynchronize()); CK(cudaFree(ctx->ws[slot])); ctx->ws[slot] = nullptr; ctx->ws_bytes[slot] = 0; }
    CK(cudaMalloc(&ctx->ws[slot], bytes));
    ctx->ws_bytes[slot] = bytes;
    return PB_OK;
}

// The workspace (ws[0..4]) is shared by every call on this context: a call enqueued on another stream must not
// start overwriting it while the previous user is still running. ws_acquire makes `st` wait for the previous user,
// ws_release marks the end of this one.
static int ws_acquire(pb_ctx *ctx, cudaStream_t st) {
    if (!ctx->ws_event) CK(cudaEventCreateWithFlags(&ctx->ws_event, cudaEventDisableTiming));
    else CK(cudaStreamWaitEvent(st, ctx->ws_event, 0));
    return PB_OK;
}
static int ws_release(pb_ctx *ctx, cudaStream_t st) {
    CK(cudaEventRecord(ctx->ws_event, st));
    return PB_OK;
}

static size_t dtype_size(int dt) { return dt == PB_U8 ? 1 : dt == PB_U16 ? 2 : dt == PB_F32 ? 4 : 8; }

static int method_nvars(int m) {
    switch (m) {
    case PB_1PF: case PB_4POLAR_2D: return 2;
    case PB_SHG: return 4;
    case PB_CARS: case PB_SRS: case PB_2PF: case PB_4POLAR_3D: return 3;
    }
    return 0;
}

extern "C" {

int pb_abi_version(void) { return PB_ABI_VERSION; }

int pb_create(pb_ctx **out, int device_ordinal) {
    if (!out) return fail(nullptr, PB_ERR_ARG, "pb_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, PB_ERR_CUDA, std::string("pb_create: no CUDA device (") + cudaGetErrorString(e) + "); libpolarb200 has no CPU path");
    if (device_ordinal < 0 || device_ordinal >= n) return fail(nullptr, PB_ERR_ARG, "pb_create: bad device ordinal");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device_ordinal);
    if (prop.major < 10) return fail(nullptr, PB_ERR_UNSUPPORTED, "pb_create: built for sm_100a (Blackwell B200) only");
    pb_ctx *ctx = new pb_ctx();
    ctx->device = device_ordinal;
    memset(&ctx->basis, 0, sizeof(ctx->basis));
    cudaSetDevice(device_ordinal);
    const int cap = 4096;
    if (cudaMalloc(&ctx->outs_dev, sizeof(pb_outputs) * cap) != cudaSuccess || cudaMalloc(&ctx->acc_dev, sizeof(double) * 8) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, PB_ERR_NOMEM, "pb_create: cudaMalloc failed");
    }
    ctx->outs_cap = cap;
    *out = ctx;
    return PB_OK;
}

int pb_destroy(pb_ctx *ctx) {
    if (!ctx) return PB_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->lut_block); cudaFree(ctx->edges_dev); cudaFree(ctx->outs_dev); cudaFree(ctx->acc_dev); cudaFree(ctx->redo_dev);
    for (int k = 0; k < 6; ++k) cudaFree(ctx->ws[k]);
    for (int k = 0; k < 2; ++k) { if (ctx->hs[k]) cudaStreamDestroy(ctx->hs[k]); }
    if (ctx->hist_pinned) cudaFreeHost(ctx->hist_pinned);
    if (ctx->ws_event) cudaEventDestroy(ctx->ws_event);
    delete ctx;
    return PB_OK;
}

const char *pb_last_error(const pb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
long pb_launch_count(const pb_ctx *ctx) { return ctx ? ctx->launches : 0; }

static void select_disk(pb_ctx *ctx, int d) {
    const size_t b_t = sizeof(double2) * (size_t)ctx->lut_n * ctx->lut_n, b_g = (sizeof(double) * ctx->lut_n + 15) & ~(size_t)15;
    char *blk = ctx->lut_block + (size_t)d * ctx->lut_stride;
    ctx->lut_dev = reinterpret_cast<double2 *>(blk);
    ctx->grid_dev = reinterpret_cast<double *>(blk + b_t);
    ctx->hy_dev = reinterpret_cast<double2 *>(blk + b_t + b_g);
    ctx->cur_disk = d;
}

int pb_set_diskcones(pb_ctx *ctx, int n_disks, const double *rho, const double *psi, int n, const double *grid) {
    if (!ctx || !rho || !psi || !grid || n < 2 || n_disks < 1 || n_disks > 4096) return fail(ctx, PB_ERR_ARG, "pb_set_diskcones: bad argument");
    // validate first: a rejected call leaves the previous tables in place.
    // The cell search corrects an arithmetic guess by at most one cell: the grid must be (numerically) uniform, as the
    // reference's always is (np.linspace(-1, 1, NbMapValues), PC:463)
    {
        const double step = (grid[n - 1] - grid[0]) / (double)(n - 1);
        for (int k = 0; k < n; ++k)
            if (!(fabs(grid[k] - (grid[0] + k * step)) <= 0.25 * fabs(step)) || !(step > 0.0))
                return fail(ctx, PB_ERR_UNSUPPORTED, "pb_set_diskcone: the grid must be uniform (np.linspace), as in Calibration.define_disk");
    }
    std::vector<double2> hy((size_t)n - 1);
    for (int k = 0; k + 1 < n; ++k) {
        hy[k].x = grid[k + 1] - grid[k];       // the reference's (grid[i+1] - grid[i]), one rounding
        hy[k].y = 1.0 / hy[k].x;               // correctly rounded reciprocal
        if (!(hy[k].x > 0.0)) return fail(ctx, PB_ERR_ARG, "pb_set_diskcone: grid must be strictly increasing");
    }
    CK(cudaSetDevice(ctx->device));
    // one allocation for all disks, each: table | grid | (h, 1/h); uploaded before the old set is released
    const size_t nn = (size_t)n * n;
    const size_t b_t = sizeof(double2) * nn, b_g = (sizeof(double) * n + 15) & ~(size_t)15, b_h = (sizeof(double2) * hy.size() + 255) & ~(size_t)255;
    const size_t stride = b_t + b_g + b_h;
    char *blk = nullptr;
    CK(cudaMalloc(&blk, stride * n_disks));
    std::vector<double2> t(nn);
    cudaError_t e_ = cudaSuccess;
    for (int d = 0; d < n_disks && e_ == cudaSuccess; ++d) {
        const double *r = rho + (size_t)d * nn, *q = psi + (size_t)d * nn;
        for (size_t k = 0; k < nn; ++k) { t[k].x = r[k]; t[k].y = q[k]; }
        char *dst = blk + (size_t)d * stride;
        e_ = cudaMemcpy(dst, t.data(), b_t, cudaMemcpyHostToDevice);
        if (e_ == cudaSuccess) e_ = cudaMemcpy(dst + b_t, grid, sizeof(double) * n, cudaMemcpyHostToDevice);
        if (e_ == cudaSuccess) e_ = cudaMemcpy(dst + b_t + b_g, hy.data(), sizeof(double2) * hy.size(), cudaMemcpyHostToDevice);
    }
    if (e_ == cudaSuccess && ctx->lut_block) e_ = cudaDeviceSynchronize();   // kernels already enqueued may still read the old tables
    if (e_ != cudaSuccess) { cudaFree(blk); ctx->err = std::string("pb_set_diskcones: ") + cudaGetErrorString(e_); return PB_ERR_CUDA; }
    if (ctx->lut_block) cudaFree(ctx->lut_block);
    ctx->lut_block = blk;
    ctx->lut_stride = stride;
    ctx->n_disks = n_disks;
    ctx->lut_n = n;
    ctx->grid_lo = grid[0]; ctx->grid_hi = grid[n - 1];
    select_disk(ctx, 0);
    return PB_OK;
}

int pb_set_diskcone(pb_ctx *ctx, const double *rho, const double *psi, int n, const double *grid) {
    return pb_set_diskcones(ctx, 1, rho, psi, n, grid);
}

int pb_select_diskcone(pb_ctx *ctx, int disk) {
    if (!ctx || disk < 0 || disk >= ctx->n_disks) return fail(ctx, PB_ERR_ARG, "pb_select_diskcone: no such resident disk cone");
    select_disk(ctx, disk);      // host-side pointer switch only: kernels already enqueued keep the table they were launched with
    return PB_OK;
}

int pb_set_invk(pb_ctx *ctx, const double *invk, int rows) {
    if (!ctx || !invk || (rows != 3 && rows != 4)) return fail(ctx, PB_ERR_ARG, "pb_set_invk: rows must be 3 (2D) or 4 (3D)");
    memset(ctx->invk, 0, sizeof(ctx->invk));
    memcpy(ctx->invk, invk, sizeof(double) * rows * 4);
    ctx->invk_rows = rows;
    return PB_OK;
}

int pb_set_basis(pb_ctx *ctx, int N, const double *c2, const double *s2, const double *c4, const double *s4) {
    if (!ctx || !c2 || !s2 || N < 1 || N > PB_MAX_ANGLES) return fail(ctx, PB_ERR_ARG, "pb_set_basis: 1 <= N <= 180 required");
    memset(&ctx->basis, 0, sizeof(ctx->basis));
    memcpy(ctx->basis.c2, c2, sizeof(double) * N);
    memcpy(ctx->basis.s2, s2, sizeof(double) * N);
    if (c4 && s4) { memcpy(ctx->basis.c4, c4, sizeof(double) * N); memcpy(ctx->basis.s4, s4, sizeof(double) * N); }
    for (int i = 0; i < N; ++i) {
        const double T = 4503599627370496.0;   // 2^52: the products are exact (power-of-two scaling)
        ctx->basis.kc2[i] = -(T * ctx->basis.c2[i]); ctx->basis.ks2[i] = -(T * ctx->basis.s2[i]);
        ctx->basis.kc4[i] = -(T * ctx->basis.c4[i]); ctx->basis.ks4[i] = -(T * ctx->basis.s4[i]);
    }
    ctx->basis_n = N;
    // the chi2 screening relies on {1, cos2a, sin2a, cos4a, sin4a} being orthogonal with norms N, N/2, ...
    const double *v[5] = {nullptr, ctx->basis.c2, ctx->basis.s2, ctx->basis.c4, ctx->basis.s4};
    const int nb = (c4 && s4) ? 5 : 3;
    double worst = 0.0;
    for (int p = 0; p < nb; ++p)
        for (int q = p; q < nb; ++q) {
            double g = 0.0;
            for (int i = 0; i < N; ++i) g += (p ? v[p][i] : 1.0) * (q ? v[q][i] : 1.0);
            double want = (p == q) ? (p == 0 ? (double)N : 0.5 * N) : 0.0;
            worst = fmax(worst, fabs(g - want));
        }
    ctx->basis_orthogonal = worst <= 1e-9 * N;
    return PB_OK;
}

int pb_set_hist(pb_ctx *ctx, int n_vars, int nbins, const double *edges, const int32_t *is_rho) {
    if (!ctx || n_vars < 0 || n_vars > PB_MAX_VARS || nbins < 1 || nbins > PB_MAX_BINS || (n_vars && (!edges || !is_rho)))
        return fail(ctx, PB_ERR_ARG, "pb_set_hist: bad argument");
    // validate first: a rejected call leaves the previous histogram set-up untouched
    // the bin search corrects an arithmetic guess by at most one bin: edges must be uniform (np.linspace, PC:367)
    for (int v = 0; v < n_vars; ++v) {
        const double *e = edges + (size_t)v * (nbins + 1);
        const double step = (e[nbins] - e[0]) / (double)nbins;
        for (int k = 0; k <= nbins; ++k)
            if (!(step > 0.0) || !(fabs(e[k] - (e[0] + k * step)) <= 0.25 * step))
                return fail(ctx, PB_ERR_UNSUPPORTED, "pb_set_hist: edges must be uniform and increasing (np.linspace), as in Variable.histo");
    }
    CK(cudaSetDevice(ctx->device));
    double *dev = nullptr;
    if (n_vars) {
        CK(cudaMalloc(&dev, sizeof(double) * n_vars * (nbins + 1)));
        cudaError_t e_ = cudaMemcpy(dev, edges, sizeof(double) * n_vars * (nbins + 1), cudaMemcpyHostToDevice);
        if (e_ != cudaSuccess) { cudaFree(dev); ctx->err = std::string("pb_set_hist: ") + cudaGetErrorString(e_); return PB_ERR_CUDA; }
    }
    // swap: kernels already enqueued may still read the old edges
    if (ctx->edges_dev) {
        cudaError_t e_ = cudaDeviceSynchronize();
        if (e_ != cudaSuccess) { cudaFree(dev); ctx->err = std::string("pb_set_hist: ") + cudaGetErrorString(e_); return PB_ERR_CUDA; }
        cudaFree(ctx->edges_dev);
    }
    ctx->edges_dev = dev;
    ctx->n_vars = n_vars; ctx->nbins = nbins; ctx->is_rho_bits = 0;
    for (int v = 0; v < n_vars; ++v) {
        const double *e = edges + (size_t)v * (nbins + 1);
        ctx->inv_w[v] = (double)nbins / (e[nbins] - e[0]);
        ctx->hist_lo[v] = e[0]; ctx->hist_hi[v] = e[nbins];
        if (is_rho[v]) ctx->is_rho_bits |= 1u << v;
    }
    return PB_OK;
}

int pb_set_rois(pb_ctx *ctx, int n_rois, const double *ilow, int n_groups, const uint32_t *group_bits) {
    if (!ctx || n_rois < 0 || n_rois > PB_MAX_ROIS || n_groups < 1 || n_groups > PB_MAX_ROIS || (n_rois && !ilow))
        return fail(ctx, PB_ERR_ARG, "pb_set_rois: bad argument");
    ctx->n_rois = n_rois; ctx->n_groups = n_groups;
    for (int r = 0; r < n_rois; ++r) ctx->roi_ilow[r] = ilow[r];
    for (int g = 0; g < n_groups; ++g) ctx->group_bits[g] = group_bits ? group_bits[g] : 0xffffffffu;
    return PB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
static int check_params(pb_ctx *ctx, const pb_params *p) {
    if (!ctx) return PB_ERR_ARG;
    if (!p) return fail(ctx, PB_ERR_ARG, "params is NULL");
    if (p->method < PB_1PF || p->method > PB_4POLAR_3D) return fail(ctx, PB_ERR_ARG, "unknown method");
    if (p->dtype < PB_U8 || p->dtype > PB_F64) return fail(ctx, PB_ERR_ARG, "unknown dtype");
    if (p->N < 1 || p->H < 1 || p->W < 1) return fail(ctx, PB_ERR_ARG, "bad stack shape");
    if (p->bin_h > 32 || p->bin_w > 32) return fail(ctx, PB_ERR_UNSUPPORTED, "bin sizes above 32 are not supported (reference GUI: 1..20)");
    const bool polar4 = p->method >= PB_4POLAR_2D;
    if (polar4) {
        if (p->N != 4) return fail(ctx, PB_ERR_ARG, "4POLAR needs exactly 4 planes");
        if (ctx->invk_rows != (p->method == PB_4POLAR_3D ? 4 : 3)) return fail(ctx, PB_ERR_STATE, "pb_set_invk not called with the K matrix of this method");
    } else {
        if (ctx->basis_n != p->N) return fail(ctx, PB_ERR_STATE, "pb_set_basis not called for this N");
        if (p->method == PB_1PF && !ctx->lut_dev && !(p->flags & PB_FLAG_PROJECT_ONLY)) return fail(ctx, PB_ERR_STATE, "pb_set_diskcone not called");
        if ((p->flags & PB_FLAG_PROJECT_ONLY) && p->method != PB_1PF) return fail(ctx, PB_ERR_UNSUPPORTED, "PB_FLAG_PROJECT_ONLY: 1PF only");
    }
    return PB_OK;
}

static void fill_args(pb_ctx *ctx, const pb_params *p, PixArgs &a) {
    memset(&a, 0, sizeof(a));
    a.lut = ctx->lut_dev; a.grid = ctx->grid_dev; a.lut_hy = ctx->hy_dev; a.lut_n = ctx->lut_n;
    a.hist_edges = ctx->edges_dev;
    a.N = p->N; a.H = p->H; a.W = p->W; a.P = (long long)p->H * p->W;
    a.method = p->method;
    a.do_sqrt = p->method == PB_CARS;
    const int nv = method_nvars(p->method);
    a.n_vars = (ctx->n_vars >= nv && ctx->edges_dev) ? nv : 0;   // histograms only if edges were given for every primary variable
    a.nbins = ctx->nbins > 0 ? ctx->nbins : 1;
    a.n_groups = ctx->n_groups; a.n_rois = ctx->n_rois;
    a.flags = p->flags;
    if (a.flags & PB_FLAG_PROJECT_ONLY) a.flags |= PB_FLAG_OUT_F64 | PB_FLAG_NO_HIST;
    if (!ctx->basis_orthogonal) a.flags |= PB_FLAG_EXACT_CHI2;
    a.hist_is_rho = ctx->is_rho_bits;
    a.sub = p->dark + p->removed_intensity; a.dark = p->dark;
    a.rot_stick = p->rotation_stick; a.rot_fig = p->rotation_fig; a.ref_angle = p->reference_angle; a.chi2thr = p->chi2_threshold;
    pb_ddrecip(p->N, &a.invN_hi, &a.invN_lo);
    a.k2 = 2.0 / (double)p->N;
    a.dN = (double)p->N;
    a.thr_lo_n = (float)(p->chi2_threshold * p->N * (1.0 - 1e-5));
    a.thr_hi_n = (float)(p->chi2_threshold * p->N * (1.0 + 1e-5));
    // the screening identity needs N >= 3 (second harmonic) / N >= 5 (fourth harmonic) equally spaced angles
    if (p->N < (p->method == PB_1PF ? 3 : 5)) a.flags |= PB_FLAG_EXACT_CHI2;
    for (int v = 0; v < PB_MAX_VARS; ++v) { a.inv_bin_width[v] = ctx->inv_w[v]; a.hist_lo[v] = ctx->hist_lo[v]; a.hist_hi[v] = ctx->hist_hi[v]; }
    for (int r = 0; r < PB_MAX_ROIS; ++r) { a.roi_ilow[r] = ctx->roi_ilow[r]; a.group_bits[r] = ctx->group_bits[r]; }
    if (ctx->n_rois == 0) a.roi_ilow[0] = p->ilow;
    memcpy(a.invk, ctx->invk, sizeof(a.invk));
    a.grid_lo = ctx->grid_lo; a.grid_hi = ctx->grid_hi; a.half_nm1 = 0.5 * (double)(ctx->lut_n - 1);
    a.simple = 0;   // set by the callers once roi_bits / mask_in are known
    a.all_lean = 0;
}

// vector stores of the per-pixel kernels need 16-byte aligned maps (4-byte aligned mask); otherwise the scalar kernels run
static int outs_aligned(const pb_outputs *o, int n) {
    uintptr_t acc = 0, m = 0;
    for (int s = 0; s < n; ++s) {
        for (int v = 0; v < PB_MAX_VARS; ++v) acc |= (uintptr_t)o[s].var[v];
        acc |= (uintptr_t)o[s].intmap | (uintptr_t)o[s].intensity | (uintptr_t)o[s].rho_angle | (uintptr_t)o[s].rho_contour;
        m |= (uintptr_t)o[s].mask;
    }
    return (acc % 16 == 0) && (m % 4 == 0);
}

// every stack asks for the throughput-mode output set: float32 maps of the variables + intmap + mask, nothing else
// (the host-side twin of store_is_lean, pb_pixel.cuh)
static int outs_lean(const pb_params *p, const pb_outputs *o, const double *edge, int n) {
    if (p->flags & PB_FLAG_OUT_F64) return 0;
    const int nv = method_nvars(p->method);
    for (int s = 0; s < n; ++s) {
        for (int v = 0; v < nv; ++v) if (!o[s].var[v]) return 0;
        if (!o[s].intmap || !o[s].mask || o[s].intensity || o[s].rho_angle || (o[s].rho_contour && edge)) return 0;
    }
    return 1;
}

// copy n output descriptors into the device ring; returns device pointer
static int push_outs(pb_ctx *ctx, const pb_outputs *outs_host, int n, cudaStream_t st, const pb_outputs **dev) {
    if (n > ctx->outs_cap) return fail(ctx, PB_ERR_ARG, "too many stacks in one call (max 4096)");
    if (ctx->outs_pos + n > ctx->outs_cap) ctx->outs_pos = 0;
    pb_outputs *d = ctx->outs_dev + ctx->outs_pos;
    ctx->outs_pos += n;
    CK(cudaMemcpyAsync(d, outs_host, sizeof(pb_outputs) * n, cudaMemcpyHostToDevice, st));
    *dev = d;
    return PB_OK;
}

static int run_box(pb_ctx *ctx, double *buf, double *tmp, int planes, int H, int W, int bh, int bw, cudaStream_t st, double **result) {
    double *cur = buf, *other = tmp;
    if (bh > 1) { CK(pb_launch_box_y(cur, other, planes, H, W, bh, st)); ctx->launches++; std::swap(cur, other); }
    if (bw > 1) { CK(pb_launch_box_x(cur, other, planes, H, W, bw, st)); ctx->launches++; std::swap(cur, other); }
    *result = cur;
    return PB_OK;
}

static const char *plan_name(pb_ctx *ctx, const pb_params *p) {
    const bool binned = p->bin_h > 1 || p->bin_w > 1;
    if (p->flags & PB_FLAG_FORCE_UNFUSED) return "unfused";
    if (!binned) return "fused_pixel";
    if (pb_march_supported(p)) return "fused_march";
    return "unfused";
}

extern "C" {

const char *pb_plan(pb_ctx *ctx, const pb_params *p) {
    if (!ctx || !p) return "invalid";
    return plan_name(ctx, p);
}

int pb_get_intensity(pb_ctx *ctx, const void *stack, int dtype, int N, int H, int W, double dark, int bh, int bw, double *out, void *stream) {
    if (!ctx || !stack || !out) return fail(ctx, PB_ERR_ARG, "pb_get_intensity: NULL argument");
    if (bh > 32 || bw > 32) return fail(ctx, PB_ERR_UNSUPPORTED, "bin sizes above 32 are not supported");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = (long long)H * W;
    if (bh <= 1 && bw <= 1) { CK(pb_launch_isum(stack, dtype, out, P, N, dark, st)); ctx->launches++; return PB_OK; }
    int rc;
    if ((rc = ws_reserve(ctx, 2, sizeof(double) * P))) return rc;
    if ((rc = ws_reserve(ctx, 3, sizeof(double) * P))) return rc;
    double *a = (double *)ctx->ws[2], *b = (double *)ctx->ws[3], *res;
    if ((rc = ws_acquire(ctx, st))) return rc;
    CK(pb_launch_isum(stack, dtype, a, P, N, dark, st)); ctx->launches++;
    if ((rc = run_box(ctx, a, b, 1, H, W, bh, bw, st, &res))) return rc;
    CK(cudaMemcpyAsync(out, res, sizeof(double) * P, cudaMemcpyDeviceToDevice, st));
    return ws_release(ctx, st);
}

int pb_binning(pb_ctx *ctx, double *field, int planes, int H, int W, int bh, int bw, void *stream) {
    if (!ctx || !field) return fail(ctx, PB_ERR_ARG, "pb_binning: NULL argument");
    if (bh > 32 || bw > 32) return fail(ctx, PB_ERR_UNSUPPORTED, "bin sizes above 32 are not supported");
    if (bh <= 1 && bw <= 1) return PB_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(double) * (size_t)planes * H * W;
    int rc;
    if ((rc = ws_reserve(ctx, 1, bytes))) return rc;
    double *res;
    if ((rc = ws_acquire(ctx, st))) return rc;
    if ((rc = run_box(ctx, field, (double *)ctx->ws[1], planes, H, W, bh, bw, st, &res))) return rc;
    if (res != field) CK(cudaMemcpyAsync(field, res, bytes, cudaMemcpyDeviceToDevice, st));
    return ws_release(ctx, st);
}

int pb_field(pb_ctx *ctx, const pb_params *p, const void *stack, double *field, void *stream) {
    if (!ctx || !p || !stack || !field) return fail(ctx, PB_ERR_ARG, "pb_field: NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = (long long)p->H * p->W;
    CK(pb_launch_field(stack, p->dtype, field, P, p->N, p->dark + p->removed_intensity, p->method == PB_CARS,
                       p->method >= PB_4POLAR_2D, st));
    ctx->launches++;
    return pb_binning(ctx, field, p->N, p->H, p->W, p->bin_h, p->bin_w, stream);
}

static int launch_pixel(pb_ctx *ctx, const pb_params *p, PixArgs &a, int dtype, int n_stacks, cudaStream_t st) {
    if (p->method >= PB_4POLAR_2D) CK(pb_launch_pixel_4polar(a, dtype, n_stacks, st));
    else CK(pb_launch_pixel_harm(a, ctx->basis, dtype, n_stacks, st));
    ctx->launches++;
    return PB_OK;
}

int pb_compute_fields(pb_ctx *ctx, const pb_params *p, const double *field, uint8_t *mask, const uint32_t *roi_bits, const double *edge,
                      const pb_outputs *out_host, void *stream) {
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!field || !out_host) return fail(ctx, PB_ERR_ARG, "pb_compute_fields: NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    PixArgs a;
    fill_args(ctx, p, a);
    pb_outputs o = *out_host;
    // the incoming mask is read and the final one written to the same buffer unless the caller gave another
    if (!o.mask) o.mask = mask;
    const pb_outputs *od;
    if ((rc = push_outs(ctx, &o, 1, st, &od))) return rc;
    a.stack = field; a.stack_stride = 0; a.from_field = 1; a.do_sqrt = 0;
    a.mask_in = mask; a.roi_bits = roi_bits; a.edge = edge; a.outs = od;
    a.out_aligned = outs_aligned(&o, 1);
    return launch_pixel(ctx, p, a, PB_F64, 1, st);
}

int pb_analyze(pb_ctx *ctx, const pb_params *p, const void *stacks, int n_stacks, const uint32_t *roi_bits, const double *edge,
               const pb_outputs *outs_host, void *stream) {
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!stacks || !outs_host || n_stacks < 1) return fail(ctx, PB_ERR_ARG, "pb_analyze: NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    PixArgs a;
    fill_args(ctx, p, a);
    a.roi_bits = roi_bits; a.edge = edge;
    a.simple = (!roi_bits && ctx->n_rois == 0 && ctx->n_groups == 1) ? 1 : 0;
    a.out_aligned = outs_aligned(outs_host, n_stacks);
    a.all_lean = outs_lean(p, outs_host, edge, n_stacks);
    const long long P = a.P;
    const std::string plan = plan_name(ctx, p);
    const pb_outputs *od;
    if ((rc = push_outs(ctx, outs_host, n_stacks, st, &od))) return rc;
    if (plan == "fused_pixel") {
        a.stack = stacks; a.stack_stride = (long long)p->N * P; a.outs = od;
        return launch_pixel(ctx, p, a, p->dtype, n_stacks, st);
    }
    if (plan == "fused_march") {
        a.stack = stacks; a.stack_stride = (long long)p->N * P; a.outs = od;
        unsigned char *redo = redo_slot(ctx, (size_t)n_stacks * ((p->H + 15) / 16));
        if (!redo) return fail(ctx, PB_ERR_NOMEM, "pb_analyze: cudaMalloc of the redo flags failed");
        const cudaError_t me = pb_launch_march(a, ctx->basis, p, n_stacks, st, &ctx->launches, redo);
        if (me == cudaSuccess) return PB_OK;
        // 4POLAR has no register-staged march: a stack the tensor map cannot describe (address / width) takes the unfused kernels
        if (!(p->method >= PB_4POLAR_2D && (me == cudaErrorInvalidConfiguration || me == cudaErrorNotSupported))) {
            ctx->err = std::string("pb_launch_march: ") + cudaGetErrorString(me);
            return PB_ERR_CUDA;
        }
        a.stack = nullptr;
    }
    // unfused: field planes + the total-intensity image as one more plane -> box(H) -> box(W) (one launch each over the N + 1 planes);
    // per-pixel kernel on the double field
    const size_t fbytes = sizeof(double) * (size_t)(p->N + 1) * P;
    if ((rc = ws_reserve(ctx, 0, fbytes))) return rc;
    if ((rc = ws_reserve(ctx, 1, fbytes))) return rc;
    const size_t esz = dtype_size(p->dtype);
    if ((rc = ws_acquire(ctx, st))) return rc;
    for (int s = 0; s < n_stacks; ++s) {
        const char *src = (const char *)stacks + (size_t)s * p->N * P * esz;
        double *Fres;
        CK(pb_launch_isum(src, p->dtype, (double *)ctx->ws[0] + (size_t)p->N * P, P, p->N, p->dark, st)); ctx->launches++;
        CK(pb_launch_field(src, p->dtype, (double *)ctx->ws[0], P, p->N, a.sub, a.do_sqrt, p->method >= PB_4POLAR_2D, st)); ctx->launches++;
        if ((rc = run_box(ctx, (double *)ctx->ws[0], (double *)ctx->ws[1], p->N + 1, p->H, p->W, p->bin_h, p->bin_w, st, &Fres))) return rc;
        const double *Ires = Fres + (size_t)p->N * P;
        PixArgs b = a;
        b.stack = Fres; b.stack_stride = 0; b.from_field = 1; b.do_sqrt = 0; b.intensity_in = Ires; b.outs = od + s;
        if ((rc = launch_pixel(ctx, p, b, PB_F64, 1, st))) return rc;
        if (outs_host[s].intensity) CK(cudaMemcpyAsync(outs_host[s].intensity, Ires, sizeof(double) * P, cudaMemcpyDeviceToDevice, st));
    }
    return ws_release(ctx, st);
}

int pb_lut_apply(pb_ctx *ctx, const pb_params *p, const double *a0, const double *a2r, const double *a2i, const uint8_t *mask, int n_stacks,
                 const uint32_t *roi_bits, const pb_outputs *outs_host, void *stream) {
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!a0 || !a2r || !a2i || !mask || !outs_host || n_stacks < 1) return fail(ctx, PB_ERR_ARG, "pb_lut_apply: NULL argument");
    if (p->method != PB_1PF) return fail(ctx, PB_ERR_UNSUPPORTED, "pb_lut_apply: the disk-cone inversion exists for 1PF only (PP:3017)");
    if (p->flags & PB_FLAG_PROJECT_ONLY) return fail(ctx, PB_ERR_ARG, "pb_lut_apply: PB_FLAG_PROJECT_ONLY belongs to pb_analyze");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    PixArgs a;
    fill_args(ctx, p, a);
    a.roi_bits = roi_bits;
    a.simple = (!roi_bits && ctx->n_rois == 0 && ctx->n_groups == 1) ? 1 : 0;
    const pb_outputs *od;
    if ((rc = push_outs(ctx, outs_host, n_stacks, st, &od))) return rc;
    a.outs = od;
    CK(pb_launch_lut_apply(a, a0, a2r, a2i, mask, n_stacks, st));
    ctx->launches++;
    return PB_OK;
}

int pb_fit_maps(pb_ctx *ctx, const pb_params *p, const double *field, const uint8_t *mask, double *fit, double *chi2, void *stream) {
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (p->method >= PB_4POLAR_2D) return fail(ctx, PB_ERR_UNSUPPORTED, "4POLAR has no fit (PP:3073)");
    CK(cudaSetDevice(ctx->device));
    CK(pb_launch_fit_maps(field, mask, (long long)p->H * p->W, p->N, p->method == PB_1PF ? 1 : 2, ctx->basis, fit, chi2, (cudaStream_t)stream));
    ctx->launches++;
    return PB_OK;
}

int pb_histo(pb_ctx *ctx, const void *values, int is_f64, const uint8_t *sel, long n, int is_rho, double rot_fig, int fold90,
             const double *edges_host, int nbins, int64_t *counts, void *stream) {
    if (!ctx || !values || !edges_host || !counts || nbins < 1 || nbins > PB_MAX_BINS) return fail(ctx, PB_ERR_ARG, "pb_histo: bad argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = ws_reserve(ctx, 4, sizeof(double) * (PB_MAX_BINS + 1)))) return rc;
    if ((rc = ws_acquire(ctx, st))) return rc;
    CK(cudaMemcpyAsync(ctx->ws[4], edges_host, sizeof(double) * (nbins + 1), cudaMemcpyHostToDevice, st));
    CK(pb_launch_histo(values, is_f64, sel, n, is_rho, rot_fig, fold90, (const double *)ctx->ws[4], nbins,
                       (double)nbins / (edges_host[nbins] - edges_host[0]), counts, st));
    ctx->launches++;
    return ws_release(ctx, st);
}

int pb_stats(pb_ctx *ctx, const void *values, const void *rho, int is_f64, const uint8_t *sel, long n, int circular, double rot_fig,
             int fold90, double *out_host, void *stream) {
    if (!ctx || !values || !out_host) return fail(ctx, PB_ERR_ARG, "pb_stats: bad argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    // circular & 1: circular statistics; circular & 2: apply the Rho re-wrap 2(v+rot)%360/2 first (PP:2744)
    const int circ = circular & 1, rewrap = (circular >> 1) & 1;
    double acc[5] = {0, 0, 0, 0, 0};
    int rc0;
    if ((rc0 = ws_acquire(ctx, st))) return rc0;
    CK(cudaMemsetAsync(ctx->acc_dev, 0, sizeof(double) * 8, st));
    CK(pb_launch_stats(1, values, rho, is_f64, sel, n, circ, rot_fig, rewrap, fold90, 0.0, ctx->acc_dev, st)); ctx->launches++;
    CK(cudaMemcpyAsync(acc, ctx->acc_dev, sizeof(double) * 3, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const double cnt = acc[0];
    double mean;
    if (circ) {
        // circularmean, PC:55-56: mod(angle(mean(exp(2j d))), 360)/2 in degrees
        double ang = atan2(acc[2] / cnt, acc[1] / cnt) * (180.0 / 3.141592653589793238462643383279502884);
        ang = fmod(ang, 360.0);
        if (ang < 0) ang += 360.0;
        mean = ang / 2.0;
    } else {
        mean = acc[1] / cnt;
    }
    CK(pb_launch_stats(2, values, rho, is_f64, sel, n, circ, rot_fig, rewrap, fold90, mean, ctx->acc_dev, st)); ctx->launches++;
    CK(cudaMemcpyAsync(acc + 3, ctx->acc_dev + 3, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const double md = acc[3] / cnt;
    double var = acc[4] / cnt - md * md;
    if (var < 0) var = 0;
    out_host[0] = cnt; out_host[1] = mean; out_host[2] = sqrt(var); out_host[3] = md; out_host[4] = circ ? NAN : acc[1];
    return ws_release(ctx, st);
}

int pb_stats_batch(pb_ctx *ctx, const pb_params *p, int pass, int per_stack, const pb_outputs *outs_host, int n_stacks, const uint32_t *roi_bits,
                   double *acc1_dev, double *acc2_dev, void *stream) {
    if (!ctx || !p || !outs_host || !acc1_dev || !acc2_dev || n_stacks < 1 || (pass != 1 && pass != 2))
        return fail(ctx, PB_ERR_ARG, "pb_stats_batch: bad argument");
    const int nv = method_nvars(p->method);
    if (!nv) return fail(ctx, PB_ERR_ARG, "pb_stats_batch: unknown method");
    for (int s = 0; s < n_stacks; ++s) {
        if (!outs_host[s].intmap) return fail(ctx, PB_ERR_ARG, "pb_stats_batch: every stack needs its variable maps and intmap");
        for (int v = 0; v < nv; ++v)
            if (!outs_host[s].var[v]) return fail(ctx, PB_ERR_ARG, "pb_stats_batch: every stack needs its variable maps and intmap");
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const pb_outputs *od;
    int rc;
    if ((rc = push_outs(ctx, outs_host, n_stacks, st, &od))) return rc;
    CK(pb_launch_stats_batch(pass, od, n_stacks, roi_bits, (long long)p->H * p->W, nv, (p->flags & PB_FLAG_OUT_F64) ? 1 : 0, p->rotation_fig,
                             ctx->group_bits, ctx->n_groups, per_stack, acc1_dev, acc2_dev, st));
    ctx->launches++;
    return PB_OK;
}

int pb_stack_from_hwn(pb_ctx *ctx, const void *stack_hwn, int dtype, int N, int H, int W, void *stack_nhw, void *stream) {
    if (!ctx || !stack_hwn || !stack_nhw || N < 1 || H < 1 || W < 1) return fail(ctx, PB_ERR_ARG, "pb_stack_from_hwn: bad argument");
    if (stack_hwn == stack_nhw) return fail(ctx, PB_ERR_ARG, "pb_stack_from_hwn: in-place conversion is not supported");
    if (N > 128) return fail(ctx, PB_ERR_UNSUPPORTED, "pb_stack_from_hwn: more than 128 angles");
    static const int bytes[4] = {1, 2, 4, 8};
    if (dtype < 0 || dtype > 3) return fail(ctx, PB_ERR_ARG, "pb_stack_from_hwn: bad dtype");
    CK(cudaSetDevice(ctx->device));
    CK(pb_launch_hwn_to_nhw(stack_hwn, stack_nhw, bytes[dtype], (long long)H * W, N, (cudaStream_t)stream));
    ctx->launches++;
    return PB_OK;
}

int pb_rescale_stack(pb_ctx *ctx, const void *values, int dtype, long n, uint16_t *out, double *minmax_host, void *stream) {
    if (!ctx || !values || !out || n < 1) return fail(ctx, PB_ERR_ARG, "pb_rescale_stack: bad argument");
    if (dtype != PB_F32 && dtype != PB_F64) return fail(ctx, PB_ERR_UNSUPPORTED, "pb_rescale_stack: integer stacks are used as they are (PP:1929)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = ws_reserve(ctx, 4, 64))) return rc;
    if ((rc = ws_acquire(ctx, st))) return rc;
    unsigned long long *mm = reinterpret_cast<unsigned long long *>(ctx->ws[4]);
    CK(pb_launch_rescale(values, dtype == PB_F64, n, mm, out, st));
    ctx->launches += 2;
    unsigned long long h[3];
    CK(cudaMemcpyAsync(h, mm, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if ((rc = ws_release(ctx, st))) return rc;
    if (h[2]) return fail(ctx, PB_ERR_UNSUPPORTED, "pb_rescale_stack: the stack holds NaN (np.amin is NaN: the reference's cast is undefined)");
    if (minmax_host) {
        for (int k = 0; k < 2; ++k) {
            if (dtype == PB_F64) {
                const unsigned long long u = (h[k] >> 63) ? (h[k] & 0x7fffffffffffffffull) : ~h[k];
                double d; memcpy(&d, &u, 8); minmax_host[k] = d;
            } else {
                const unsigned v = (unsigned)h[k];
                const unsigned u = (v & 0x80000000u) ? (v & 0x7fffffffu) : ~v;
                float f; memcpy(&f, &u, 4); minmax_host[k] = f;
            }
        }
    }
    return PB_OK;
}

int pb_compact(pb_ctx *ctx, const void *values, const void *filter, int is_f64, const uint8_t *sel, const uint32_t *roi_bits, uint32_t group_bits,
               long n, void *out_f16, void *out_val, uint32_t *out_index, long *count_host, void *stream) {
    if (!ctx || !values || !count_host || n < 0) return fail(ctx, PB_ERR_ARG, "pb_compact: bad argument");
    if (n >= 4294967296L && out_index) return fail(ctx, PB_ERR_UNSUPPORTED, "pb_compact: linear indices are 32-bit");
    *count_host = 0;
    if (n == 0) return PB_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = ws_reserve(ctx, 4, pb_compact_ws_bytes(n)))) return rc;
    if ((rc = ws_acquire(ctx, st))) return rc;
    const unsigned long long *total_dev = nullptr;
    CK(pb_launch_compact(values, filter, is_f64, sel, roi_bits, group_bits, n, ctx->ws[4], out_f16, out_val, out_index, &total_dev, st));
    ctx->launches += 3;
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, total_dev, sizeof(total), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *count_host = (long)total;
    return ws_release(ctx, st);
}

int pb_compute_dark(pb_ctx *ctx, const void *stack, int dtype, int N, int H, int W, int size_cell, double *dark_host, void *stream) {
    if (!ctx || !stack || !dark_host || size_cell < 1) return fail(ctx, PB_ERR_ARG, "pb_compute_dark: bad argument");
    if (dtype != PB_U8 && dtype != PB_U16) return fail(ctx, PB_ERR_UNSUPPORTED, "pb_compute_dark: integer stacks only (uint8 / uint16)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int ny = H / size_cell, nx = W / size_cell;
    *dark_host = 0.0;
    if (N < 1 || ny < 1 || nx < 1) return PB_OK;                        // values.size == 0 -> 0.0 (PC:267-268)
    const size_t ncell = (size_t)ny * nx;
    int rc;
    if ((rc = ws_reserve(ctx, 4, ncell * 12 + 64))) return rc;
    if ((rc = ws_acquire(ctx, st))) return rc;
    unsigned long long *sums = (unsigned long long *)ctx->ws[4];
    unsigned *counts = (unsigned *)(sums + ncell + 2);
    CK(pb_launch_dark_cells(stack, dtype, W, ny, nx, size_cell, sums, counts, st)); ctx->launches++;
    std::vector<unsigned long long> hs(ncell);
    std::vector<unsigned> hc(ncell);
    CK(cudaMemcpyAsync(hs.data(), sums, ncell * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hc.data(), counts, ncell * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // np.argmin of the masked means: cells without a non-zero sample are masked (filled with the maximum) -> first minimum
    size_t best = 0;
    double bestv = INFINITY;
    bool any = false;
    for (size_t k = 0; k < ncell; ++k) {
        if (!hc[k]) continue;
        const double m = (double)hs[k] / (double)hc[k];
        if (!any || m < bestv) { bestv = m; best = k; any = true; }
    }
    if (!any) best = 0;
    const int ci = (int)(best / nx), cj = (int)(best % nx);
    unsigned long long *acc = sums + ncell;
    CK(cudaMemsetAsync(acc, 0, 16, st));
    CK(pb_launch_dark_cell_all(stack, dtype, (long long)H * W, W, N, ci * size_cell, cj * size_cell, size_cell, acc, st)); ctx->launches++;
    unsigned long long r[2];
    CK(cudaMemcpyAsync(r, acc, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *dark_host = r[1] ? (double)r[0] / (double)r[1] : NAN;              // an all-zero cell has a fully masked mean
    return ws_release(ctx, st);
}

int pb_analyze_host(pb_ctx *ctx, const pb_params *p, const void *stacks_host, int n_stacks, const uint32_t *roi_bits_host,
                    const double *edge_host, const pb_outputs *outs_host) {
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!stacks_host || !outs_host || n_stacks < 1) return fail(ctx, PB_ERR_ARG, "pb_analyze_host: NULL argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());   // synchronous API: earlier asynchronous work of this context may still use the workspace
    for (int k = 0; k < 2; ++k) {
        if (!ctx->hs[k]) CK(cudaStreamCreateWithFlags(&ctx->hs[k], cudaStreamNonBlocking));
    }
    const long long P = (long long)p->H * p->W;
    const size_t esz = dtype_size(p->dtype), sbytes = (size_t)p->N * P * esz;
    const bool f64 = p->flags & PB_FLAG_OUT_F64;
    const size_t msz = f64 ? 8 : 4;
    const int nv = method_nvars(p->method);
    const int nhist = ctx->n_groups * nv * (ctx->nbins > 0 ? ctx->nbins : 1);
    // device staging per pipeline slot: stack | maps (vars, intmap, rho_angle, rho_contour) | intensity | mask | hist
    auto al = [](size_t n) { return (n + 255) & ~(size_t)255; };
    const size_t per_slot = al(sbytes) + (size_t)(nv + 3) * al(P * msz) + al(P * 8) + al(P) + al(sizeof(int64_t) * nhist);
    if ((rc = ws_reserve(ctx, 5, 2 * per_slot + (size_t)P * 12 + 64))) return rc;
    char *base = (char *)ctx->ws[5];
    uint32_t *roi_dev = nullptr;
    double *edge_dev = nullptr;
    char *extra = base + 2 * per_slot;
    if (edge_host) { edge_dev = (double *)extra; CK(cudaMemcpyAsync(edge_dev, edge_host, P * 8, cudaMemcpyHostToDevice, ctx->hs[0])); }
    if (roi_bits_host) { roi_dev = (uint32_t *)(extra + P * 8); CK(cudaMemcpyAsync(roi_dev, roi_bits_host, P * 4, cudaMemcpyHostToDevice, ctx->hs[0])); }
    CK(cudaStreamSynchronize(ctx->hs[0]));
    // per-stack histograms come back through PINNED staging: a device-to-host copy into pageable memory would block the
    // host until the stream drains and serialise the whole copy/compute pipeline
    const size_t nh_all = (size_t)n_stacks * nhist;
    if (nh_all > ctx->hist_pinned_n) {
        if (ctx->hist_pinned) cudaFreeHost(ctx->hist_pinned);
        ctx->hist_pinned = nullptr; ctx->hist_pinned_n = 0;
        CK(cudaHostAlloc((void **)&ctx->hist_pinned, sizeof(int64_t) * nh_all, cudaHostAllocDefault));
        ctx->hist_pinned_n = nh_all;
    }
    int64_t *htmp = ctx->hist_pinned;
    // the unfused path shares the context workspace between stacks: keep it on one stream
    const bool serial = std::string(plan_name(ctx, p)) == "unfused";
    for (int s = 0; s < n_stacks; ++s) {
        const int k = serial ? 0 : (s & 1);
        cudaStream_t st = ctx->hs[k];
        char *slot = base + (size_t)k * per_slot;
        size_t off = 0;
        auto take = [&](size_t n) { char *q = slot + off; off += (n + 255) & ~(size_t)255; return q; };
        void *dstack = take(sbytes);
        pb_outputs od;
        memset(&od, 0, sizeof(od));
        const pb_outputs &oh = outs_host[s];
        for (int v = 0; v < nv; ++v) if (oh.var[v]) od.var[v] = take(P * msz);
        if (oh.intmap) od.intmap = take(P * msz);
        if (oh.rho_angle) od.rho_angle = take(P * msz);
        if (oh.rho_contour) od.rho_contour = take(P * msz);
        if (oh.intensity) od.intensity = take(P * 8);
        if (oh.mask) od.mask = (uint8_t *)take(P);
        if (oh.hist) od.hist = (int64_t *)take(sizeof(int64_t) * nhist);
        CK(cudaMemcpyAsync(dstack, (const char *)stacks_host + (size_t)s * sbytes, sbytes, cudaMemcpyHostToDevice, st));
        if (od.hist) CK(cudaMemsetAsync(od.hist, 0, sizeof(int64_t) * nhist, st));
        if ((rc = pb_analyze(ctx, p, dstack, 1, roi_dev, edge_dev, &od, st))) return rc;
        for (int v = 0; v < nv; ++v) if (oh.var[v]) CK(cudaMemcpyAsync(oh.var[v], od.var[v], P * msz, cudaMemcpyDeviceToHost, st));
        if (oh.intmap) CK(cudaMemcpyAsync(oh.intmap, od.intmap, P * msz, cudaMemcpyDeviceToHost, st));
        if (oh.rho_angle) CK(cudaMemcpyAsync(oh.rho_angle, od.rho_angle, P * msz, cudaMemcpyDeviceToHost, st));
        if (oh.rho_contour) CK(cudaMemcpyAsync(oh.rho_contour, od.rho_contour, P * msz, cudaMemcpyDeviceToHost, st));
        if (oh.intensity) CK(cudaMemcpyAsync(oh.intensity, od.intensity, P * 8, cudaMemcpyDeviceToHost, st));
        if (oh.mask) CK(cudaMemcpyAsync(oh.mask, od.mask, P, cudaMemcpyDeviceToHost, st));
        if (oh.hist) CK(cudaMemcpyAsync(htmp + (size_t)s * nhist, od.hist, sizeof(int64_t) * nhist, cudaMemcpyDeviceToHost, st));
        // slot k is reused by stack s+2: that stack's H2D is enqueued on the same stream, hence ordered after these copies
    }
    CK(cudaStreamSynchronize(ctx->hs[0]));
    CK(cudaStreamSynchronize(ctx->hs[1]));
    for (int s = 0; s < n_stacks; ++s)
        if (outs_host[s].hist)
            for (int k = 0; k < nhist; ++k) outs_host[s].hist[k] += htmp[(size_t)s * nhist + k];
    return PB_OK;
}

}  // extern "C"
