/* polar_b200.h -- C ABI of libpolarb200.so: the B200 (sm_100a) implementation of PyPOLAR's per-pixel
 * polarimetric analysis path.
 *
 * The reference (cchandre/Polarimetry v2.9.4) has no FFI: its seam is four Python call sites
 * (SURVEY.md 8b). Each entry point below names the reference statement(s) it replaces
 * (PP = pypolar/PyPOLAR.py, PC = pypolar/pypolar_classes.py). INTEGRATION.md shows the ctypes stubs a
 * maintainer of the reference would add to bind them.
 *
 * Conventions
 *  - plain C types only; every pointer is either a HOST pointer (suffix _host / documented) or a DEVICE
 *    pointer owned by the caller. The library never allocates inside an analysis call except through the
 *    context's workspace (grown by pb_reserve / on first use), never synchronises the device unless
 *    documented, and enqueues all work on the `stream` argument (a cudaStream_t passed as void*).
 *  - return value: 0 on success, negative pb_status on error; message via pb_last_error().
 *  - stacks are angle-major [N,H,W] C-contiguous, exactly as the reference holds them (PP:1926).
 *  - invalid pixels are NaN in every floating-point map and 0 in `mask` (PP:3052, PC:326).
 *  - there is no CPU fallback: every compute entry point fails with PB_ERR_CUDA if no device is usable.
 */
#ifndef POLAR_B200_H
#define POLAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 1
#define PB_MAX_ROIS 32   /* ROI membership travels as one uint32 bit set per pixel */
#define PB_MAX_VARS 4    /* Rho,Psi | Rho,S2,S4,S_SHG | Rho,Psi,Eta (PP:3026, PP:3036-3040, PP:3047-3051) */
#define PB_MAX_BINS 128  /* histogram bins per variable (reference GUI allows 10..90, PP:461-462) */

typedef enum pb_status {
    PB_OK = 0, PB_ERR_ARG = -1, PB_ERR_CUDA = -2, PB_ERR_STATE = -3, PB_ERR_UNSUPPORTED = -4, PB_ERR_NOMEM = -5
} pb_status;

/* modality, PP:332 */
typedef enum pb_method {
    PB_1PF = 0, PB_CARS = 1, PB_SRS = 2, PB_SHG = 3, PB_2PF = 4, PB_4POLAR_2D = 5, PB_4POLAR_3D = 6
} pb_method;

/* element type of the input stack */
typedef enum pb_dtype { PB_U8 = 0, PB_U16 = 1, PB_F32 = 2, PB_F64 = 3 } pb_dtype;

/* flags */
#define PB_FLAG_OUT_F64      1u  /* maps are written as double (bit-comparable with the reference); default float */
#define PB_FLAG_NO_HIST      2u  /* skip in-kernel histograms */
#define PB_FLAG_EXACT_CHI2   4u  /* always evaluate fit/chi2 in fp64 (disables the certified screening) */
#define PB_FLAG_FORCE_UNFUSED 8u /* run the separate kernels (field, box filter, per-pixel) even where a fused one exists */
#define PB_FLAG_PROJECT_ONLY 16u /* 1PF: stop before the disk-cone inversion and write var[0] = Re a2, var[1] = Im a2 (un-normalised,
                                    PP:2989), intmap = a0 (PP:2987) and mask = the mask of PP:3000 -- the disk-independent part of the
                                    analysis, input of pb_lut_apply. Implies double maps and no histograms. */

#define PB_FLAG_HIST_MATCH   32u /* warp-aggregate the shared-memory histogram updates with a match on (variable, bin) before the
                                    atomic: one atomic per distinct bin of a warp instead of one per lane (aligned samples whose
                                    rho falls into two or three bins); the default leaves the aggregation to the atomic unit */
#define PB_FLAG_STRIP_MARCH  64u /* binned 1PF, 18 angles, 3x3 window: take the strip-march kernel (pb_strip2.cuh: two lanes per image
                                    row carry the box-filter state of all planes in registers) + the redo pass of the row-march kernel
                                    instead of the row-march kernel alone. Same results bit for bit; opt-in while it is the slower one. */

typedef struct pb_ctx pb_ctx;

/* The settings analyze_stack pulls from the GUI at PP:2950-2968 (SURVEY.md section 5 "Config / flags"). */
typedef struct pb_params {
    int32_t method;            /* pb_method */
    int32_t dtype;             /* pb_dtype of the stack */
    int32_t N, H, W;           /* angles, height, width */
    int32_t bin_h, bin_w;      /* box filter sizes along H (axis 1) and W (axis 2): PP:2943-2945 */
    uint32_t flags;
    double dark;               /* PP:2951; the threshold image uses dark only (PC:261) */
    double removed_intensity;  /* PP:2953 */
    double rotation_stick;     /* rotation[0], PP:3022 */
    double rotation_fig;       /* rotation[1], PC:342 */
    double reference_angle;    /* rotation[2], PP:3069 */
    double chi2_threshold;     /* PP:2981 default 500 */
    double ilow;               /* threshold when no ROI is given, PP:2876 */
} pb_params;

/* Device output buffers of one analysis (any pointer may be NULL = not wanted).
 * Map element type is float, or double with PB_FLAG_OUT_F64. Any alignment natural for the element type is accepted; the
 * vectorised kernels are used when every map is 16-byte aligned (mask: 4-byte), as torch / cudaMalloc allocations are. */
typedef struct pb_outputs {
    void *var[PB_MAX_VARS];    /* [H,W] each, in the reference's datastack.vars order */
    void *intmap;              /* [H,W] a0 under the final mask (PP:3052-3053) */
    void *intensity;           /* [H,W] total-intensity image the threshold compares against (PC:260-264); always double */
    void *rho_angle;           /* [H,W] (rho - reference_angle) % 180 (PP:3069-3072), written if reference_angle != 0 */
    void *rho_contour;         /* [H,W] (rho - edge) % 180 (PP:3059-3068), written if an edge map is given */
    uint8_t *mask;             /* [H,W] final mask (PP:3000, PP:3025, PP:3042) */
    int64_t *hist;             /* [n_groups, n_vars, nbins] ACCUMULATED (+=) -- zero it before the first stack */
} pb_outputs;

/* ---- life cycle ------------------------------------------------------------------------------- */
int pb_abi_version(void);
/* one context per GPU; owns calibration tables, basis, histogram edges and the workspace */
int pb_create(pb_ctx **out, int device_ordinal);
int pb_destroy(pb_ctx *ctx);
const char *pb_last_error(const pb_ctx *ctx);   /* ctx may be NULL: last creation error */

/* ---- calibration / constants (HOST pointers, copied; synchronous) ------------------------------ */
/* replaces Calibration.define_disk (PC:460-464): RoTest/PsiTest [n,n] row-major, grid = linspace(-1,1,n) */
int pb_set_diskcone(pb_ctx *ctx, const double *rho_table_host, const double *psi_table_host, int n, const double *grid_host);
/* several disk cones resident at once (the 1PF calibration sweep loops over every file of the diskcones folder, PP:850-913):
 * rho / psi tables [n_disks][n,n] back to back on the host, one common grid. Disk 0 becomes the current one; pb_select_diskcone
 * switches the table used by the following pb_analyze / pb_lut_apply calls without any copy, allocation or synchronisation. */
int pb_set_diskcones(pb_ctx *ctx, int n_disks, const double *rho_tables_host, const double *psi_tables_host, int n, const double *grid_host);
int pb_select_diskcone(pb_ctx *ctx, int disk);
/* replaces Calibration.invKmat (PC:450-451): pinv(K) row-major [rows,4], rows = 3 (2D) | 4 (3D) */
int pb_set_invk(pb_ctx *ctx, const double *invk_host, int rows);
/* harmonic basis computed on the host by NumPy exactly as PP:2983-2986 / PP:2993 (cos/sin of 2a and 4a) */
int pb_set_basis(pb_ctx *ctx, int N, const double *cos2_host, const double *sin2_host, const double *cos4_host, const double *sin4_host);
/* histogram edges per variable: np.linspace(vmin, vmax, nbins+1) (PC:367); is_rho marks the 2(v+rot)%360/2 re-wrap (PC:341-342) */
int pb_set_hist(pb_ctx *ctx, int n_vars, int nbins, const double *edges_host /*[n_vars, nbins+1]*/, const int32_t *is_rho_host /*[n_vars]*/);
/* ROI thresholds and histogram groups. roi_ilow[r] = float(roi['ILow']) (PP:2880); a histogram group g
 * selects the pixels whose ROI bit set intersects group_bits[g] (PP:2894-2904). n_rois = 0 => whole image. */
int pb_set_rois(pb_ctx *ctx, int n_rois, const double *roi_ilow_host, int n_groups, const uint32_t *group_bits_host);

/* ---- the fused hot path: replaces Polarimetry.analyze_stack's numerics (PP:2948-2968) ------------ */
/* stacks_dev: n_stacks stacks back to back, each [N,H,W] of params->dtype (DEVICE).
 * roi_bits_dev: [H,W] uint32 bit r set <=> pixel inside ROI r and base mask != 0 (PP:2883-2891); NULL = all pixels, bit 0.
 * edge_dev: [H,W] double contour-angle map or NULL (PP:3059).
 * outs_host: n_stacks pb_outputs structs (HOST array of DEVICE pointers); hist pointers may alias to accumulate a batch.
 * The same ROI / edge maps apply to every stack of the batch. */
int pb_analyze(pb_ctx *ctx, const pb_params *params, const void *stacks_dev, int n_stacks, const uint32_t *roi_bits_dev,
               const double *edge_dev, const pb_outputs *outs_host, void *stream);

/* the same analysis from HOST memory (stack and outputs in page-locked or pageable host buffers): copies
 * in, analyses, copies the wanted outputs back, overlapping copies and kernels over two internal streams
 * when n_stacks > 1. Synchronous. outs_host here holds HOST pointers. */
int pb_analyze_host(pb_ctx *ctx, const pb_params *params, const void *stacks_host, int n_stacks, const uint32_t *roi_bits_host,
                    const double *edge_host, const pb_outputs *outs_host);

/* ---- the individual seams (for a drop-in that keeps the reference's call structure) --------------- */
/* replaces Stack.get_intensity (PC:260-264): out [H,W] double */
int pb_get_intensity(pb_ctx *ctx, const void *stack_dev, int dtype, int N, int H, int W, double dark, int bin_h, int bin_w,
                     double *out_dev, void *stream);
/* replaces analyze_stack's pre-processing + Polarimetry.binning (PP:2953-2957, PP:2942-2946): out [N,H,W] double */
int pb_field(pb_ctx *ctx, const pb_params *params, const void *stack_dev, double *field_dev, void *stream);
/* replaces Polarimetry.binning alone (PP:2942-2946) on a double field, in place */
int pb_binning(pb_ctx *ctx, double *field_dev, int planes, int H, int W, int bin_h, int bin_w, void *stream);
/* replaces compute_fields (PP:2981-3078) on a double, already binned field; mask_dev [H,W] uint8 is read as the
 * incoming mask and overwritten with the final one (the reference mutates it in place, PP:3000) */
int pb_compute_fields(pb_ctx *ctx, const pb_params *params, const double *field_dev, uint8_t *mask_dev, const uint32_t *roi_bits_dev,
                      const double *edge_dev, const pb_outputs *out_host, void *stream);
/* fit / chi2 maps for the individual-fit viewer (PP:3073-3077): fit [N,H,W], chi2 [H,W], NaN outside mask */
int pb_fit_maps(pb_ctx *ctx, const pb_params *params, const double *field_dev, const uint8_t *mask_dev, double *fit_dev,
                double *chi2_dev, void *stream);
/* replaces the numeric part of Variable.histo (PC:340-369): counts[nbins] += histogram of values under sel */
int pb_histo(pb_ctx *ctx, const void *values_dev, int is_f64, const uint8_t *sel_dev, long n, int is_rho, double rotation_fig,
             int fold90, const double *edges_host, int nbins, int64_t *counts_dev, void *stream);
/* replaces return_vecexcel's reductions (PP:2744-2779) for one variable under sel & finite(rho):
 * out_host[0..4] = N, mean, std, mean(delta) (circular if is_rho), sum. Synchronises `stream`. */
int pb_stats(pb_ctx *ctx, const void *values_dev, const void *rho_dev, int is_f64, const uint8_t *sel_dev, long n, int circular,
             double rotation_fig, int fold90, double *out_host, void *stream);

/* Batch form of the same reductions (ROI statistics of a batch sharded over GPUs; semantic model merge_histos /
 * return_vecexcel on the concatenated values, PP:2318-2384, PP:2739-2779). Statistic slots v = 0 .. n_vars: the method's primary
 * variables (Rho first: circular, re-wrapped with params->rotation_fig) and, last, intmap; for every histogram group g
 * (pb_set_rois) over the pixels of ALL n_stacks maps with the group's ROI bit set and a finite Rho:
 *   pass 1: acc1[g][v][0..2] += N, sum d (Rho: sum cos 2d), (Rho: sum sin 2d)         acc1: double [n_groups][PB_MAX_VARS+1][4]
 *   pass 2: acc2[g][v][0..1] += sum delta, sum delta^2                                  acc2: double [n_groups][PB_MAX_VARS+1][2]
 *           delta = d - mean (Rho: wrapto180(2(d - mean))/2), the means FINALISED ON THE DEVICE from acc1.
 * The accumulators are plain sums: a batch sharded over several GPUs all-reduces acc1 between the passes and acc2 after
 * (two messages of a few hundred bytes), then N, mean, std = sqrt(sum delta^2 / N - (sum delta / N)^2) follow on the host.
 * per_stack != 0: one accumulator set per stack (acc1 [n_stacks][n_groups][..], acc2 likewise) instead of one for the batch
 * (the calibration sweep wants one statistics row per stack, PP:883-898).
 * outs_host: the DEVICE map pointers of each stack as written by pb_analyze (var[], intmap). Zero the accumulators first.
 * Asynchronous on `stream`; never synchronises. */
int pb_stats_batch(pb_ctx *ctx, const pb_params *params, int pass, int per_stack, const pb_outputs *outs_host, int n_stacks,
                   const uint32_t *roi_bits_dev, double *acc1_dev, double *acc2_dev, void *stream);

/* "next" row N2: replaces Stack.compute_dark (PC:266-278): the mean over all angles of the non-zero samples of
 * the size_cell x size_cell cell whose first-angle non-zero mean is smallest. Integer stacks. Synchronises `stream`. */
int pb_compute_dark(pb_ctx *ctx, const void *stack_dev, int dtype, int N, int H, int W, int size_cell, double *dark_host, void *stream);

/* "next" row N1: the disk-dependent tail of the 1PF analysis (PP:2991, PP:3017-3026, PP:3052, PC:340-369) on the
 * maps written by pb_analyze with PB_FLAG_PROJECT_ONLY: normalisation a2/a0, disk-cone inversion with the CURRENT disk
 * (pb_set_diskcone), rho rotation / wrap, final mask, intmap, histograms. A calibration sweep over D disks (PP:850-913) projects
 * every stack once and calls this D times. a0/a2r/a2i: [n_stacks][H,W] double, mask: [n_stacks][H,W] uint8 (DEVICE). */
int pb_lut_apply(pb_ctx *ctx, const pb_params *params, const double *a0_dev, const double *a2r_dev, const double *a2i_dev,
                 const uint8_t *mask_dev, int n_stacks, const uint32_t *roi_bits_dev, const pb_outputs *outs_host, void *stream);

/* Layout conversion at the boundary: an angle-minor stack [H,W,N] (the N samples of a pixel contiguous) to the
 * reference's angle-major [N,H,W] (PP:1926), which every other entry point takes. Out of place, same dtype, N <= 128. */
int pb_stack_from_hwn(pb_ctx *ctx, const void *stack_hwn_dev, int dtype, int N, int H, int W, void *stack_nhw_dev, void *stream);

/* "next" row N3: the numeric part of Polarimetry.define_stack (PP:1923-1942) for non-integer TIFF stacks:
 * out = (65535 * (a - amin(a)) / ptp(a)).astype(uint16) in the array's own precision (PP:1929-1930), min / max over the WHOLE
 * stack. values_dev: n elements of dtype PB_F32 / PB_F64 (DEVICE), out_dev: n uint16. minmax_host (optional): amin, amax.
 * A stack holding NaN is refused (the reference's cast is undefined there). Synchronises `stream`. */
int pb_rescale_stack(pb_ctx *ctx, const void *values_dev, int dtype, long n, uint16_t *out_dev, double *minmax_host, void *stream);

/* "next" row N4: the boolean-mask gathers of the result writers, on maps the device already holds.
 *   save_mat (PP:2829-2842): var.values[mask & isfinite(var.values)].astype(float16)       -> filter_dev = NULL, out_f16
 *   save_csv (PP:2794-2827): xmap/ymap/var.values[mask & isfinite(vars[0].values)]          -> filter_dev = vars[0], out_val, out_index
 * kept = (sel == NULL || sel[p]) && (roi_bits == NULL || roi_bits[p] & group_bits) && isfinite(filter ? filter[p] : values[p]);
 * kept elements are written in row-major order (stable compaction): out_f16 = IEEE half of the value (one rounding from fp64,
 * numpy's astype), out_val = the value itself (same type as values), out_index = linear pixel index p (x = p % W, y = p / W are
 * the writer's X / Y columns). Any output may be NULL; each must hold up to n elements (or the count of a previous call with the
 * same filter). *count_host = number kept. Synchronises `stream`. */
int pb_compact(pb_ctx *ctx, const void *values_dev, const void *filter_dev, int is_f64, const uint8_t *sel_dev, const uint32_t *roi_bits_dev,
               uint32_t group_bits, long n, void *out_f16_dev, void *out_val_dev, uint32_t *out_index_dev, long *count_host, void *stream);

/* ---- introspection ------------------------------------------------------------------------------ */
/* number of kernel launches this context has enqueued so far (bench.py's gpu_launches) */
long pb_launch_count(const pb_ctx *ctx);
/* name of the code path pb_analyze would take for these params ("fused_pixel", "fused_march", "unfused") */
const char *pb_plan(pb_ctx *ctx, const pb_params *params);

#ifdef __cplusplus
}
#endif
#endif /* POLAR_B200_H */
