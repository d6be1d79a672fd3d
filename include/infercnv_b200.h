/*
 * infercnv_b200.h - C ABI of libinfercnv_b200.so: the B200 (sm_100a) drop-in for the
 * smoothing + HMM hot path of broadinstitute/infercnv.
 *
 * The reference is pure R (DESCRIPTION: "NeedsCompilation: no"; NAMESPACE has no useDynLib), so it
 * has no native interface to bind to.  The boundary is the set of internal R functions that
 * infercnv::run() calls on this path; each entry point below replaces the body of one of them
 * (file:line of the reference given per function) and is what a `.Call()` shim binds
 * (see INTEGRATION.md and infercnv_b200/r/).
 *
 * Conventions
 *   layout   R column-major, X[g + G*c]: g = gene (row, chromosome-ordered), c = cell (column).
 *            A cell's gene vector is contiguous.  Sizes are int64 (R long vectors).
 *   dtype    float64 across the ABI (R `double`); states are int32 (1..m, -1 = not assigned), or one byte each
 *            (1..m, 255 = not assigned) from the *_u8 variants - a quarter of the bytes over PCIe.
 *   indices  0-based everywhere (the R shim subtracts 1).  Index lists are CSR style:
 *            group k owns grp_idx[grp_off[k] .. grp_off[k+1]).
 *   chromosomes  K contiguous row ranges [chr_start[k], chr_start[k]+chr_len[k]) - the reference
 *            pre-sorts rows by chr,start,stop (R/inferCNV.R:407-413).
 *   ownership  caller owns every buffer it passes; inputs are never modified; outputs may alias
 *            inputs only where stated.  The library owns all device memory, streams and a pinned
 *            staging ring (two slabs in, two out per GPU): caller memory that is not page-locked -
 *            an R matrix, a NumPy array - is moved through that ring by the library's copy threads,
 *            page-locked caller memory is used directly.  Everything is freed in icnv_shutdown().
 *   errors   every entry returns 0 or a negative icnv_status and records a per-thread message
 *            readable through icnv_last_error().  Nothing throws across the ABI.  There is no CPU
 *            fallback: without a usable CUDA device every compute entry returns ICNV_E_NO_DEVICE.
 *   threading  calls are synchronous (result resident in the caller's buffer at return) and
 *            serialised internally; call from one host thread at a time per process.
 *
 * Two families:
 *   icnv_*      host pointers in / host pointers out (what R binds).  H2D / D2H inside.
 *   icnv_dev_*  device pointers on the library's current device + a cudaStream_t (as void*);
 *               asynchronous on that stream.  Used by the multi-GPU driver and the benchmark.
 */
#ifndef INFERCNV_B200_H
#define INFERCNV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ICNV_API __attribute__((visibility("default")))
#else
#define ICNV_API
#endif

typedef enum icnv_status {
    ICNV_OK = 0,
    ICNV_E_NO_DEVICE = -1,   /* no CUDA device / driver, or icnv_init failed */
    ICNV_E_CUDA = -2,        /* a CUDA runtime call or kernel failed */
    ICNV_E_BAD_ARG = -3,     /* NULL pointer, negative size, inconsistent ranges, even window, ... */
    ICNV_E_UNSUPPORTED = -4, /* outside the kernels' envelope (e.g. G too large for shared memory) */
    ICNV_E_NONFINITE = -5,   /* NA/NaN/Inf in the input where the reference would stop() */
    ICNV_E_UNDERFLOW = -6,   /* "Problems With Underflow" (R/inferCNV_HMM.R:1165-1166) */
    ICNV_E_NOMEM = -7
} icnv_status;

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Select `device` (cudaSetDevice), create the library's stream and scratch pools.  Idempotent for
 * the same device; re-initialises when called with another one.  device < 0 -> keep the current
 * set-up, or device 0 when there is none.  Tuning switches (ICNV_* environment variables, DESIGN.md
 * section 8) are read here, once, and the ones that are set are reported on stderr. */
ICNV_API int icnv_init(int device);
/* Multi-GPU for a single host process - the form the R shim uses (infercnv::run() is one R process;
 * its num_threads, R/inferCNV_ops.R:388, never reaches this path): initialise n_devices GPUs
 * (device_ids == NULL: devices 0 .. n-1; n_devices <= 0: every device present).  Afterwards the
 * host-pointer entry points that stream cells (icnv_smooth_block_f64, icnv_smooth_hmm_*_f64, the
 * per-cell icnv_viterbi_*_f64) cut the cells into one contiguous range per GPU, driven by one host
 * thread each; every GPU reduces ALL reference cells itself in list order (they are <= ~10 % of the
 * matrix and cross PCIe once per GPU), so no inter-GPU exchange is needed and the result is bit-identical
 * for any device count ("broadcast the ref-cell columns").  All other entry points use
 * the first device.  icnv_dev_* calls address the device of the calling thread's context (the first). */
ICNV_API int icnv_init_devices(int n_devices, const int *device_ids);
ICNV_API int icnv_devices_in_use(void);        /* 0 before initialisation */
/* Host threads that move the caller's (pageable) matrix into / out of the pinned staging ring; they are
 * split over the GPUs in use.  0 = default: half the hardware threads, at most 16. */
ICNV_API int icnv_set_host_threads(int n);
ICNV_API void icnv_shutdown(void);
ICNV_API int icnv_device_count(void);          /* >= 0, or a negative icnv_status */
ICNV_API const char *icnv_last_error(void);    /* never NULL */
ICNV_API const char *icnv_version(void);
/* Number of kernel launches issued by this library since icnv_init (for bench `gpu_launches`). */
ICNV_API int64_t icnv_launch_count(void);

/* HMM arithmetic.  mode 0: reference-order IEEE arithmetic for every sequence.  mode 1 (default): certified FP64 pass -
 * table emission + structured recursion, every arg-max margin checked, and each sequence whose smallest margin is below
 * 1e-7 recomputed in mode-0 arithmetic.  mode 2 (option, measured slower on the benchmark data): a single-precision pass
 * first, certified by the margins along the returned path only; what it cannot certify goes through the mode-1 pass, what
 * that cannot certify through mode 0.  The state calls are the same in all modes.  Also settable with the environment
 * variable ICNV_HMM_MODE (read at icnv_init). */
ICNV_API int icnv_set_hmm_mode(int mode);
/* Number of sequences the last Viterbi call recomputed in reference-order arithmetic (syncs). */
ICNV_API int64_t icnv_hmm_rerun_count(void);
/* Number of sequences the single-precision pass (mode 2) of the last Viterbi call could not certify and handed to the FP64
 * pass (syncs); 0 in modes 0 and 1. */
ICNV_API int64_t icnv_hmm_second_pass_count(void);

/* ---- host-pointer entry points (what the R shim binds) ---------------------------------------- */

/* .get_normal_gene_mean_bounds, R/inferCNV_ops.R:1708-1735.
 * means[g + G*k] = mean(X[g, group k]);  inv_log != 0: log2(mean(2^x - 1) + 1). */
ICNV_API int icnv_ref_means_f64(const double *X, int64_t G, int64_t C, const int32_t *grp_off,
                                const int32_t *grp_idx, int n_grp, int inv_log, double *means);

/* .subtract_expr, R/inferCNV_ops.R:1742-1786 (body of subtract_ref_expr_from_obs, :1678-1702).
 * use_bounds: y = x-max_k(means) if x>max; x-min_k(means) if x<min; else 0.  Otherwise x-mean_k(means).
 * Y may alias X. */
ICNV_API int icnv_subtract_ref_f64(const double *X, double *Y, int64_t G, int64_t C, const double *means,
                                   int n_grp, int use_bounds);

/* smooth_by_chromosome -> .smooth_window -> .smooth_helper, R/inferCNV_ops.R:2406-2532, 2640-2661.
 * Truncated, renormalised triangular ("pyramidinal") moving average per cell per chromosome.
 * window < 2: copy.  Even window: ICNV_E_BAD_ARG (the reference's behaviour is accidental there).
 * Chromosomes with < 2 genes are left untouched (:2417).  NA input: ICNV_E_NONFINITE. */
ICNV_API int icnv_smooth_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                             const int32_t *chr_len, int K, int window);

/* .center_columns, R/inferCNV_ops.R:2094-2109 (center_cell_expr_across_chromosome, :2074-2088).
 * use_median != 0: subtract the per-cell median over all G genes; else the per-cell mean. */
ICNV_API int icnv_center_f64(const double *X, double *Y, int64_t G, int64_t C, int use_median);

/* The two steps either side of the path, so the reference's bundled example runs
 * count.data -> expr.data on the GPU end to end:
 * normalize_counts_by_seq_depth (R/inferCNV_ops.R:3064-3111; normalize_factor < 0 or NaN = median of colSums) and
 * clear_noise_via_ref_mean_sd (R/inferCNV_ops.R:2302-2346, noise_logistic = FALSE). */
ICNV_API int icnv_normalize_counts_by_seq_depth_f64(const double *X, double *Y, int64_t G, int64_t C,
                                                    double normalize_factor);
ICNV_API int icnv_clear_noise_via_ref_mean_sd_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *idx,
                                                  int64_t n_idx, double sd_amplifier);

/* remove_outliers_norm / .remove_outliers_norm, R/inferCNV_ops.R:1969-2056 (run() step 16): clamp to
 * [lower_bound, upper_bound].  A NaN bound (R's NA) selects out_method "average_bound" (.get_average_bounds,
 * ops.R:2734-2742): lower = mean over cells of each cell's smallest value, upper = mean of each cell's largest.
 * bounds_out: optional double[2] receiving the bounds used. */
ICNV_API int icnv_remove_outliers_norm_f64(const double *X, double *Y, int64_t G, int64_t C, double lower_bound,
                                           double upper_bound, double *bounds_out);

/* clear_noise / .clear_noise, R/inferCNV_ops.R:2232-2275 (run() step 22 with a numeric noise_filter): centre = mean
 * over all values of the listed cells (n_idx == 0: all data); values strictly inside centre +- threshold become
 * centre, or - noise_logistic != 0 - are pulled towards it by depress_log_signal_midpt_val (slope 20,
 * R/inferCNV_heatmap.R:2783-2810).  threshold == 0: unchanged copy. */
ICNV_API int icnv_clear_noise_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *idx, int64_t n_idx,
                                  double threshold, int noise_logistic);

/* clear_noise_via_ref_mean_sd(noise_logistic = TRUE), R/inferCNV_ops.R:2325-2329. */
ICNV_API int icnv_clear_noise_via_ref_mean_sd_logistic_f64(const double *X, double *Y, int64_t G, int64_t C,
                                                           const int32_t *idx, int64_t n_idx, double sd_amplifier);

/* Element-wise steps as stand-alone calls (inside icnv_smooth_block_f64 they are fused into the loads
 * and stores): log2xplus1 (R/inferCNV_ops.R:2756-2769), invert_log2 (:2814-2826),
 * apply_max_threshold_bounds (:2970-2983).  n = G*C elements, Y may alias X. */
ICNV_API int icnv_log2xplus1_f64(const double *X, double *Y, int64_t n);
ICNV_API int icnv_invert_log2_f64(const double *X, double *Y, int64_t n);
ICNV_API int icnv_apply_max_threshold_bounds_f64(const double *X, double *Y, int64_t n, double threshold);

/* Fused run() steps 4, 8, 9, 10, 11, 12, 14 (R/inferCNV_ops.R:614, 771, 817, 865, 911, 952, 1031):
 * [log2(x+1)] -> subtract ref (bounds) -> clamp +-threshold -> smooth(window) -> centre by
 * median -> subtract ref again -> 2^x.  X is the depth-normalised matrix (after step 3).
 * n_grp == 0 is BAD_ARG (callers pass the proxy group of all observation cells when there are no
 * references, as the reference does at ops.R:1686-1689).  threshold <= 0 disables the clamp. */
ICNV_API int icnv_smooth_block_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                                   const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx,
                                   int n_grp, int apply_log, double threshold, int window, int use_bounds);

/* Fused icnv_smooth_block_f64 + per-cell icnv_viterbi_f64 in ONE pass over the matrix: run() steps
 * 4..14 followed by step 17 for analysis_mode = "cells" (R/inferCNV_ops.R:614-1031, :1235-1309); with
 * the default prune_outliers = FALSE nothing between those steps touches expr.data.  Y receives the
 * smoothed matrix (what run() keeps as the preliminary object), states the HMM calls on it. */
ICNV_API int icnv_smooth_hmm_f64(const double *X, double *Y, int32_t *states, int64_t G, int64_t C,
                                 const int32_t *chr_start, const int32_t *chr_len, int K, const int32_t *grp_off,
                                 const int32_t *grp_idx, int n_grp, int apply_log, double threshold, int window,
                                 int use_bounds, int m, const double *Pi, const double *delta, const double *mean,
                                 const double *sd);
/* same, states as one byte each (1..m): the host-pointer calls are PCIe-bound, and int32 states are a third of
 * the bytes this call returns.  This is the variant the R shim binds (it widens to R doubles either way). */
ICNV_API int icnv_smooth_hmm_u8_f64(const double *X, double *Y, uint8_t *states, int64_t G, int64_t C,
                                    const int32_t *chr_start, const int32_t *chr_len, int K, const int32_t *grp_off,
                                    const int32_t *grp_idx, int n_grp, int apply_log, double threshold, int window,
                                    int use_bounds, int m, const double *Pi, const double *delta, const double *mean,
                                    const double *sd);

/* Viterbi.dthmm.adj, R/inferCNV_HMM.R:1101-1176, batched over the drivers
 * predict_CNV_via_HMM_on_indiv_cells (HMM.R:284-324), ..._on_tumor_subclusters (:345-408),
 * ..._on_whole_tumor_samples (:509-567) and the i3 twins (R/inferCNV_i3HMM.R:180-389).
 * m = 6 (i6) or 3 (i3).  Pi: m x m column-major, Pi[j + m*k] = P(j -> k); delta[m]; mean[m].
 * n_grp == 0: one sequence per (cell, chromosome), sd[m].
 * n_grp  > 0: one sequence per (group, chromosome) on rowMeans(X[chr, group]); sd[m*n_grp]
 *             (per-group sds as .get_state_emission_params, HMM.R:586-614, computes them);
 *             the trace is written to every cell of the group, other cells get -1.
 * states: int32 G x C.  margins: optional (may be NULL) K x (C or n_grp) doubles receiving, per
 * sequence, the smallest winner/runner-up gap over the traceback decisions (diagnostic).
 * Sequences with < 2 genes get state 3 (HMM.R:1104-1107). */
ICNV_API int icnv_viterbi_f64(const double *X, int64_t G, int64_t C, const int32_t *chr_start,
                              const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx,
                              int n_grp, int m, const double *Pi, const double *delta, const double *mean,
                              const double *sd, int32_t *states, double *margins);
/* same, states as uint8 G x C (1..m, 255 where the int32 variant writes -1) */
ICNV_API int icnv_viterbi_u8_f64(const double *X, int64_t G, int64_t C, const int32_t *chr_start,
                                 const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx,
                                 int n_grp, int m, const double *Pi, const double *delta, const double *mean,
                                 const double *sd, uint8_t *states, double *margins);

/* assign_HMM_states_to_proxy_expr_vals, R/inferCNV_HMM.R:1191-1206 (m = 6: states 1..6 -> 0, 0.5, 1, 1.5, 2, 3) and
 * i3HMM_assign_HMM_states_to_proxy_expr_vals, R/inferCNV_i3HMM.R:405-417 (m = 3: 1..3 -> 0.5, 1, 1.5) on the numeric
 * state matrix (n entries); any other value (e.g. -1) passes through.  Y may alias X. */
ICNV_API int icnv_assign_hmm_states_to_proxy_expr_vals_f64(const double *X, double *Y, int64_t n, int m);

/* predict_CNV_via_HMM_on_tumor_subclusters_per_chr, R/inferCNV_HMM.R:412-471: every chromosome k has its own list of
 * subclusters - groups chr_grp_off[k] .. chr_grp_off[k+1]) of the CSR lists (grp_off, grp_idx); one Viterbi sequence
 * per (chromosome, subcluster) on rowMeans(X[chr, cells]) with that group's sds (sd: m per group, the median is used,
 * HMM.R:1122); the trace is written to the subcluster's cells on that chromosome's genes, cells in no subcluster of a
 * chromosome get 255 (R's -1, HMM.R:436) there.  A cell in two subclusters of one chromosome: ICNV_E_BAD_ARG. */
ICNV_API int icnv_viterbi_per_chr_u8_f64(const double *X, int64_t G, int64_t C, const int32_t *chr_start,
                                         const int32_t *chr_len, int K, const int32_t *chr_grp_off,
                                         const int32_t *grp_off, const int32_t *grp_idx, int m, const double *Pi,
                                         const double *delta, const double *mean, const double *sd, uint8_t *states);

/* apply_median_filtering / .median_filter, R/noise_reduction.R:43-113.  Blocks = chromosome x one
 * index list (a subcluster for observations, a whole group for references), cells in list order.
 * Window radius is (window_size+1)/2 as in the reference (noise_reduction.R:102-106).
 * Cells in no list are copied.  window_size even or < 3: ICNV_E_BAD_ARG. */
ICNV_API int icnv_median_filter_f64(const double *X, double *Y, int64_t G, int64_t C, const int32_t *chr_start,
                                    const int32_t *chr_len, int K, const int32_t *grp_off, const int32_t *grp_idx,
                                    int n_grp, int window_size);

/* .i3HMM_get_sd_trend_by_num_cells_fit, R/inferCNV_i3HMM.R:17-30: mean and sd (n-1) over all
 * values of the listed cells. */
ICNV_API int icnv_mean_sd_f64(const double *X, int64_t G, int64_t C, const int32_t *idx, int64_t n_idx, double *mu,
                              double *sigma);

/* parallelDist(t(expr.data[, cells])) with its default method "euclidean" - the distance matrix behind every hclust() of
 * the reference: R/inferCNV_tumor_subclusters.R:191,411,472,582,609, R/inferCNV_ops.R:1930,3242,
 * R/inferCNV_heatmap.R:719,755,1062,1079.  cells: n 0-based columns (NULL: all C columns in
 * order, n is then ignored).  out: R's "dist" vector, n (n - 1) / 2 doubles, the strict lower triangle by columns -
 * out[n a - a (a + 1) / 2 + (b - a - 1)] for list positions a < b.  Difference form, genes in order, one fused
 * multiply-add per term (no cancellation for near-identical cells).  n < 2: nothing is written. */
ICNV_API int icnv_pairwise_dist_f64(const double *X, int64_t G, int64_t C, const int32_t *cells, int64_t n, double *out);
/* The same distances from the matrix AS parallelDist() RECEIVES IT - observations x variables, column-major, i.e. the
 * t(expr.data[, cells]) the reference builds at every call site (n cells x G genes, cell a / gene g at x[a + n g]): the R
 * closure passes its argument through without transposing it back. */
ICNV_API int icnv_pairwise_dist_rows_f64(const double *x, int64_t n, int64_t G, double *out);

/* Combine per-cell (sum, sd) pairs - e.g. all-gathered from several GPUs - into mu / sigma over all values
 * of those cells, in list order (what icnv_mean_sd_f64 does internally). */
ICNV_API void icnv_combine_cell_stats(const double *sums, const double *sds, int64_t n, int64_t G, double *mu,
                                      double *sigma);

/* ---- gene filters and counts ingest: the steps in front of the path (run() steps 2-3) ------------------------
 * Per-gene statistics of a dense G x C matrix for the two gene filters of run() step 2:
 *   sums[g]  = sum over all cells; means[g] = rowMeans as R rounds it (long-double quotient -> double), what
 *              .below_min_mean_expr_cutoff compares with the cutoff (R/inferCNV_ops.R:2149-2158);
 *   n_pos[g] = number of cells with x > 0 (NaN counts as not expressed), what require_above_min_cells_ref compares
 *              with min_cells_per_gene (ops.R:2177-2209).
 * Any of the three outputs may be NULL. */
ICNV_API int icnv_gene_stats_f64(const double *X, int64_t G, int64_t C, double *sums, int32_t *n_pos, double *means);

/* scale_infercnv_expr, R/inferCNV_ops.R:3174-3186 (run() step 5, scale_data = TRUE): t(scale(t(expr))) - every gene
 * (row) centred at its mean over the cells and divided by its sd (n - 1).  A constant gene gives NaN, as in R. */
ICNV_API int icnv_scale_infercnv_expr_f64(const double *X, double *Y, int64_t G, int64_t C);

/* remove_genes, R/inferCNV.R:445-457: Y (n_keep x C) = X[keep, ]; keep = increasing 0-based row indices. */
ICNV_API int icnv_remove_genes_f64(const double *X, int64_t G, int64_t C, const int32_t *keep, int64_t n_keep,
                                   double *Y);

/* Rows in any order, repeats allowed: Y (n x C) = X[idx, ] - the reordering of the expression matrix to the genomic
 * position table in .order_reduce (R/inferCNV.R:352-428; CreateInfercnvObject's ingest). */
ICNV_API int icnv_gather_genes_f64(const double *X, int64_t G, int64_t C, const int32_t *idx, int64_t n, double *Y);

/* The same statistics for a compressed-sparse-column counts matrix (a dgCMatrix's @p, @i, @x; the reference accepts
 * one as raw_counts_matrix, R/inferCNV.R:158-160): p[C+1] column pointers, i[nnz] 0-based row indices, x[nnz]. */
ICNV_API int icnv_csc_gene_stats_f64(const int32_t *p, const int32_t *i, const double *x, int64_t G, int64_t C,
                                     double *sums, int32_t *n_pos, double *means);

/* Gene removal + normalize_counts_by_seq_depth (ops.R:3064-3111) straight from the compressed columns: Y, dense
 * G_out x C (G_out = n_keep, or G when keep == NULL), = (x / colSums over the kept genes) * normalize_factor; a
 * negative or NaN factor means "median of the column sums" (the reference's NA default).  The matrix crosses PCIe
 * as 12 bytes per stored count instead of 8 bytes per cell-gene.  col_sums: optional C doubles. */
ICNV_API int icnv_csc_normalize_f64(const int32_t *p, const int32_t *i, const double *x, int64_t G, int64_t C,
                                    const int32_t *keep, int64_t n_keep, double normalize_factor, double *Y,
                                    double *col_sums);

/* ---- CNV region calling on the HMM state matrix (the step after the Viterbi kernels) ---------------------
 * States use the one-byte wire format of the *_u8 Viterbi variants: 0..6 = state, 255 = unassigned (R's -1).
 *
 * .get_state_consensus, R/inferCNV_HMM.R:977-988: consensus[g + G*k] = the most frequent state of gene g over
 * the cells of group k; ties go to the smallest state, 255 (-1) ordering before every state, as
 * table() / order(decreasing=TRUE)[1] do.  A byte outside {0..6, 255}: ICNV_E_BAD_ARG. */
ICNV_API int icnv_state_consensus_u8(const uint8_t *states, int64_t G, int64_t C, const int32_t *grp_off,
                                     const int32_t *grp_idx, int n_grp, uint8_t *consensus);

/* The step that ends predict_CNV_via_HMM_on_tumor_subclusters_per_chr (HMM.R:470-483: get_predicted_CNV_regions by
 * "subcluster", then every region's state written to all cells of its group): out[g, c] = consensus state of c's
 * group at gene g, for the genes region calling covers (chromosomes with >= 2 genes); other genes and cells in no
 * group keep their state.  out may alias states. */
ICNV_API int icnv_apply_state_consensus_u8(const uint8_t *states, int64_t G, int64_t C, const int32_t *chr_start,
                                           const int32_t *chr_len, int K, const int32_t *grp_off,
                                           const int32_t *grp_idx, int n_grp, uint8_t *out);

/* .define_cnv_gene_regions + .get_cnv_gene_region_bounds, R/inferCNV_HMM.R:1006-1087, for n_seq state sequences
 * (the columns of seqs, G x n_seq): within every chromosome of >= 2 genes (shorter ones are skipped, :1012-1014)
 * each run of equal states is one region.  Regions are produced in (sequence, chromosome, position) order - the
 * order get_predicted_CNV_regions numbers them in (:735-760), so region i (0-based) of a call is "region_<i+1>".
 * gene_start / gene_stop: gene_order$start / $stop as doubles (exact for integer positions below 2^53).
 * The call leaves the records in the library and reports their number; icnv_cnv_regions_fetch copies them out. */
ICNV_API int icnv_cnv_regions_u8(const uint8_t *seqs, int64_t G, int64_t n_seq, const int32_t *chr_start,
                                 const int32_t *chr_len, int K, const double *gene_start, const double *gene_stop,
                                 int64_t *n_regions);

/* get_predicted_CNV_regions, R/inferCNV_HMM.R:706-764, in one upload of the state matrix: consensus of every
 * cell group (reference groups + observation groups for by = "consensus", the subclusters for by = "subcluster",
 * one single-cell group per cell for by = "cell"; the caller orders the groups as the reference does) followed
 * by region calling on the n_grp consensus sequences.  consensus: optional G x n_grp output (may be NULL). */
ICNV_API int icnv_predicted_cnv_regions_u8(const uint8_t *states, int64_t G, int64_t C, const int32_t *chr_start,
                                           const int32_t *chr_len, int K, const double *gene_start,
                                           const double *gene_stop, const int32_t *grp_off, const int32_t *grp_idx,
                                           int n_grp, uint8_t *consensus, int64_t *n_regions);

/* Records of the last icnv_cnv_regions_u8 / icnv_predicted_cnv_regions_u8 call; n_regions must be the number that
 * call reported.  Per region: seq (0-based sequence / group), chr (0-based chromosome), first_gene .. last_gene
 * (0-based, inclusive), state (0..6, -1 = unassigned), start = min(gene_start), end = max(gene_stop) over its
 * genes (:1078-1079).  Any output pointer may be NULL. */
ICNV_API int icnv_cnv_regions_fetch(int64_t n_regions, int32_t *seq, int32_t *chr, int32_t *first_gene,
                                    int32_t *last_gene, int32_t *state, double *start, double *end);

/* ---- device-pointer entry points ------------------------------------------------------------- */
/* All pointers are device pointers on the icnv_init() device unless marked host.  `stream` is a
 * cudaStream_t; NULL = the library's own (non-blocking) stream - pass cudaStreamLegacy ((void*)1) to
 * run on the legacy default stream.  Asynchronous: the caller synchronises. */

/* Partial sums for group means, fixed summation order so results do not depend on the number of
 * GPUs: cells[0..n_cells) (device, column indices into X) are cut into chunks of `chunk` list
 * entries; partial[g + G*q] = sum over chunk q, in list order, of f(X[g, cell]) with
 * f = identity, or log2(x+1) when apply_log.  n_chunks = ceil(n_cells/chunk). */
ICNV_API int icnv_dev_group_partial_sums_f64(const double *X, int64_t G, int64_t ldx, const int32_t *cells,
                                             int64_t n_cells, int chunk, int apply_log, double *partial,
                                             void *stream);
/* Reference bounds (min / max / mean over the groups of the per-gene group means) from chunk sums laid out
 * [world][tot_rows][G] - what an all-gather of every rank's icnv_dev_group_partial_sums_f64 rows gives; group k owns rows
 * row_off[k] .. row_off[k+1] of each rank's block, counts[k] is its global size.  Replaces .get_normal_gene_mean_bounds'
 * per-group mean() (R/inferCNV_ops.R:1708-1735) + the min / max of .subtract_expr (:1742-1786) in one launch. */
ICNV_API int icnv_dev_bounds_from_partials_f64(const double *partials, int64_t G, int world, int64_t tot_rows, int n_grp,
                                               const int32_t *row_off, const int64_t *counts, double *lo, double *hi,
                                               double *mid, void *stream);
/* The group means themselves from the same layout of chunk sums: means[g + G*k] (rowMeans of a group spread over ranks -
 * the group modes of the HMM, R/inferCNV_HMM.R:383 - in the fixed chunk order, so the rank count does not matter). */
ICNV_API int icnv_dev_means_from_partials_f64(const double *partials, int64_t G, int world, int64_t tot_rows, int n_grp,
                                              const int32_t *row_off, const int64_t *counts, double *means, void *stream);
/* Group-mode broadcast on the device: out[g + G*c] = group_states[g + G*grp_of[c]], 255 where grp_of[c] < 0
 * (R/inferCNV_HMM.R:368, 399: every cell of a group gets the group's trace, cells in no group stay unassigned). */
ICNV_API int icnv_dev_scatter_group_states_u8(const uint8_t *group_states, int64_t G, int64_t C, const int32_t *grp_of,
                                              uint8_t *out, void *stream);
/* means[g] = (sum_q partial[g + G*q], q ascending) / count */
ICNV_API int icnv_dev_combine_partials_f64(const double *partial, int64_t G, int64_t n_chunks, int64_t count,
                                           double *means, void *stream);
/* lo/hi/mid over the n_grp columns of means (each length G). */
ICNV_API int icnv_dev_bounds_from_means_f64(const double *means, int64_t G, int n_grp, double *lo, double *hi,
                                            double *mid, void *stream);

/* The fused per-cell kernel.  For every listed column (cols == NULL: columns 0..n_cols-1 of X):
 *   v = X[:, col]; [v = log2(v+1)]; [v = dead-band subtract (lo1, hi1) or v - mid1];
 *   [clamp +-threshold]; [smooth(window) per chromosome]; [v -= median(v) | mean(v)];
 *   [dead-band subtract (lo2, hi2) or v - mid2]; [v = 2^v]; Y[:, i] = v   (output column i, ld = ldy)
 * Stage selection: lo1/hi1 NULL -> skip; mid1 used when lo1 == NULL && mid1 != NULL; threshold <= 0
 * -> skip; window < 2 -> skip; center: 0 none, 1 median, 2 mean; lo2/hi2/mid2 likewise; apply_exp2.
 * err_flag (device int*, may be NULL) is set to 1 when a non-finite value is seen. */
ICNV_API int icnv_dev_cell_pipeline_f64(const double *X, int64_t G, int64_t ldx, const int32_t *cols, int64_t n_cols,
                                        double *Y, int64_t ldy, const int32_t *chr_start /*host*/,
                                        const int32_t *chr_len /*host*/, int K, int apply_log, const double *lo1,
                                        const double *hi1, const double *mid1, double threshold, int window,
                                        int center, const double *lo2, const double *hi2, const double *mid2,
                                        int apply_exp2, int *err_flag, void *stream);

/* Whole smooth block on device-resident data (single GPU): X, Y are G x C (ld = G). */
ICNV_API int icnv_dev_smooth_block_f64(const double *X, double *Y, int64_t G, int64_t C,
                                       const int32_t *chr_start /*host*/, const int32_t *chr_len /*host*/, int K,
                                       const int32_t *grp_off /*host*/, const int32_t *grp_idx /*host*/, int n_grp,
                                       int apply_log, double threshold, int window, int use_bounds, void *stream);

/* Per-sequence Viterbi on device-resident X (G x C, ld = G), one sequence per (column, chromosome).
 * states_u8: uint8 G x C (1..m).  margins: device K x C doubles or NULL.  Model arrays are host.
 * sd has m entries (sd_per_col == 0) or m per column (sd_per_col != 0, device-order = column). */
ICNV_API int icnv_dev_viterbi_f64(const double *X, int64_t G, int64_t C, const int32_t *chr_start /*host*/,
                                  const int32_t *chr_len /*host*/, int K, int m, const double *Pi /*host*/,
                                  const double *delta /*host*/, const double *mean /*host*/,
                                  const double *sd /*host*/, int sd_per_col, uint8_t *states_u8, double *margins,
                                  int *err_flag, void *stream);

/* Per-cell sum and sd (n-1) of the listed columns (cells == NULL: columns 0..n_cells-1); sums / sds may be NULL. */
ICNV_API int icnv_dev_column_stats_f64(const double *X, int64_t G, const int32_t *cells, int64_t n_cells, double *sums,
                                       double *sds, void *stream);

/* Median filter on device-resident data; index lists are host arrays. */
ICNV_API int icnv_dev_median_filter_f64(const double *X, double *Y, int64_t G, int64_t C,
                                        const int32_t *chr_start /*host*/, const int32_t *chr_len /*host*/, int K,
                                        const int32_t *grp_off /*host*/, const int32_t *grp_idx /*host*/, int n_grp,
                                        int window_size, void *stream);

/* icnv_pairwise_dist_f64 on device pointers: columns X + ldx * (d_cells ? d_cells[a] : a), a < n; out: n (n - 1) / 2
 * doubles (device).  Asynchronous on `stream`. */
ICNV_API int icnv_dev_pairwise_dist_f64(const double *X, int64_t G, int64_t ldx, const int32_t *d_cells, int64_t n,
                                        double *out, void *stream);
ICNV_API int icnv_dev_pairwise_dist_rows_f64(const double *x, int64_t n, int64_t ldx, int64_t G, double *out, void *stream);
/* Region calling on device-resident states (uint8, column stride lds >= G).
 * icnv_dev_state_counts_u8: counts[(k*G + g)*8 + slot] = number of cells of group k (d_cells[h_grp_off[k] ..
 * h_grp_off[k+1]), device) whose state at gene g falls in `slot` (0: unassigned, v+1: state v); integer counts, so
 * ranks holding different cells of a group can add theirs (all-reduce) before icnv_dev_consensus_from_counts.
 * err_flag |= 4 on a byte outside {0..6, 255}.  Synchronises the stream. */
ICNV_API int icnv_dev_state_counts_u8(const uint8_t *S, int64_t G, int64_t lds, const int32_t *d_cells,
                                      const int32_t *h_grp_off /*host*/, int n_grp, uint32_t *d_counts, int *err_flag,
                                      void *stream);
ICNV_API int icnv_dev_consensus_from_counts(const uint32_t *d_counts, int64_t G, int n_grp, uint8_t *d_cons,
                                            void *stream);
ICNV_API int icnv_dev_state_consensus_u8(const uint8_t *S, int64_t G, int64_t lds, const int32_t *d_cells,
                                         const int32_t *h_grp_off /*host*/, int n_grp, uint8_t *d_cons, int *err_flag,
                                         void *stream);
/* Regions of the n_seq sequences d_seqs[:, d_cols ? d_cols[s] : s]; synchronises; records stay in the library. */
ICNV_API int icnv_dev_cnv_regions_u8(const uint8_t *d_seqs, int64_t G, int64_t lds, int64_t n_seq,
                                     const int32_t *d_cols, const int32_t *chr_start /*host*/,
                                     const int32_t *chr_len /*host*/, int K, const double *gene_start /*host*/,
                                     const double *gene_stop /*host*/, int64_t *n_regions /*host*/, void *stream);
/* Device pointers of the last call's records (valid until the next region call); any output may be NULL. */
ICNV_API int icnv_dev_cnv_regions_records(int64_t *n, const int32_t **seq, const int32_t **chr,
                                          const int32_t **first_gene, const int32_t **last_gene,
                                          const int32_t **state, const double **start, const double **end);

/* Gene filters / ingest / outlier clamp on device-resident data (see the host entry points of the same names):
 * per-gene sum and number of values > 0 over the C columns (column stride ldx); row gather Y[i + n_keep*c] =
 * X[keep[i] + ldx*c]; per-cell min / max (NaN skipped); clamp to [lower, upper]. */
ICNV_API int icnv_dev_gene_stats_f64(const double *X, int64_t G, int64_t ldx, int64_t C, double *sums, int32_t *n_pos,
                                     void *stream);
ICNV_API int icnv_dev_gather_rows_f64(const double *X, int64_t ldx, const int32_t *keep, int64_t n_keep, double *Y,
                                      int64_t C, void *stream);
ICNV_API int icnv_dev_column_minmax_f64(const double *X, int64_t G, int64_t C, double *mins, double *maxs, void *stream);
ICNV_API int icnv_dev_clamp_bounds_f64(const double *X, double *Y, int64_t n, double lower, double upper, void *stream);

/* Deterministic synthetic workload of SURVEY section 8(d) written straight into HBM: counter-based
 * generator keyed by (seed, global cell, gene), so any sharding of the cells yields identical data.
 * Fills X (G x n_cells, ld = G) for global cells [cell0, cell0 + n_cells). */
ICNV_API int icnv_dev_synth_f64(double *X, int64_t G, int64_t cell0, int64_t n_cells, int64_t C_total,
                                const int32_t *chr_start /*host*/, const int32_t *chr_len /*host*/, int K,
                                uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* INFERCNV_B200_H */
