/*
 * infercnvb200_shim.c - the `.Call()` glue between R and libinfercnv_b200.so.
 *
 * Build on a machine with R:   R CMD SHLIB infercnvb200_shim.c -L<dir> -linfercnv_b200 -I<repo>/include
 * (tests/test_r_shim_compiles.py compile-checks this file against a minimal mock of Rinternals.h,
 *  because R is not installable in the build image.)
 *
 * Every entry: extract plain vectors from the SEXPs, allocate the result with allocMatrix, make ONE
 * call into the C ABI, and turn a non-zero status into an R error AFTER all locals are dead
 * (Rf_error longjmps).  Index vectors arrive 1-based from R and are shifted here.
 */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include <stdlib.h>

#include "infercnv_b200.h"

/* list of integer vectors (1-based) -> CSR (0-based); caller frees *off and *idx */
static int list_to_csr(SEXP groups, int32_t **off, int32_t **idx) {
    int n = Rf_length(groups);
    int64_t total = 0;
    for (int k = 0; k < n; ++k) total += Rf_length(VECTOR_ELT(groups, k));
    *off = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
    if (!*off || !*idx) return -1;
    int32_t pos = 0;
    (*off)[0] = 0;
    for (int k = 0; k < n; ++k) {
        SEXP v = VECTOR_ELT(groups, k);
        const int *p = INTEGER(v);
        for (int i = 0; i < Rf_length(v); ++i) (*idx)[pos++] = p[i] - 1;
        (*off)[k + 1] = pos;
    }
    return n;
}

/* as.integer(gene_order$chr) (rows pre-sorted by chr) -> contiguous ranges */
static int chr_to_ranges(SEXP chr_codes, int32_t **start, int32_t **len) {
    int G = Rf_length(chr_codes);
    const int *c = INTEGER(chr_codes);
    int K = 0;
    for (int g = 0; g < G; ++g)
        if (g == 0 || c[g] != c[g - 1]) ++K;
    *start = (int32_t *)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));
    *len = (int32_t *)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));
    if (!*start || !*len) return -1;
    int k = -1;
    for (int g = 0; g < G; ++g) {
        if (g == 0 || c[g] != c[g - 1]) {
            ++k;
            (*start)[k] = g;
            (*len)[k] = 0;
        }
        (*len)[k] += 1;
    }
    return K;
}

static void fail_if(int rc) {
    if (rc != 0) Rf_error("infercnv_b200: %s (status %d)", icnv_last_error(), rc);
}

/* subtract_ref_expr_from_obs: .get_normal_gene_mean_bounds + .subtract_expr (ops.R:1678-1786) */
SEXP icnvR_subtract_ref(SEXP expr, SEXP ref_groups, SEXP inv_log, SEXP use_bounds) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int32_t *off = NULL, *idx = NULL;
    int n_grp = list_to_csr(ref_groups, &off, &idx);
    double *means = (double *)malloc(sizeof(double) * (size_t)(G * (n_grp > 0 ? n_grp : 1)));
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = (n_grp < 0 || !means) ? ICNV_E_NOMEM
                                   : icnv_ref_means_f64(REAL(expr), G, C, off, idx, n_grp, Rf_asLogical(inv_log), means);
    if (rc == 0) rc = icnv_subtract_ref_f64(REAL(expr), REAL(ans), G, C, means, n_grp, Rf_asLogical(use_bounds));
    free(off); free(idx); free(means);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* smooth_by_chromosome (ops.R:2406-2434) */
SEXP icnvR_smooth(SEXP expr, SEXP chr_codes, SEXP window_length) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int32_t *cs = NULL, *cl = NULL;
    int K = chr_to_ranges(chr_codes, &cs, &cl);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = K < 0 ? ICNV_E_NOMEM : icnv_smooth_f64(REAL(expr), REAL(ans), G, C, cs, cl, K, Rf_asInteger(window_length));
    free(cs); free(cl);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* center_cell_expr_across_chromosome (ops.R:2074-2109) */
SEXP icnvR_center(SEXP expr, SEXP use_median) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = icnv_center_f64(REAL(expr), REAL(ans), G, C, Rf_asLogical(use_median));
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* fused run() steps 4, 8-12, 14 */
SEXP icnvR_smooth_block(SEXP expr, SEXP chr_codes, SEXP ref_groups, SEXP apply_log, SEXP threshold, SEXP window_length,
                        SEXP use_bounds) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int32_t *cs = NULL, *cl = NULL, *off = NULL, *idx = NULL;
    int K = chr_to_ranges(chr_codes, &cs, &cl);
    int n_grp = list_to_csr(ref_groups, &off, &idx);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = (K < 0 || n_grp < 0) ? ICNV_E_NOMEM
                                  : icnv_smooth_block_f64(REAL(expr), REAL(ans), G, C, cs, cl, K, off, idx, n_grp,
                                                          Rf_asLogical(apply_log), Rf_asReal(threshold),
                                                          Rf_asInteger(window_length), Rf_asLogical(use_bounds));
    free(cs); free(cl); free(off); free(idx);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* fused run() steps 4, 8-12, 14 AND step 17 for analysis_mode = "cells" (per-cell i6 / i3 HMM on the block's output):
 * ONE upload of the matrix, the smoothed matrix and one byte per state back.  Returns list(expr, states). */
SEXP icnvR_smooth_hmm(SEXP expr, SEXP chr_codes, SEXP ref_groups, SEXP apply_log, SEXP threshold, SEXP window_length,
                      SEXP use_bounds, SEXP Pi, SEXP delta, SEXP mean, SEXP sd) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int m = Rf_length(delta);
    int32_t *cs = NULL, *cl = NULL, *off = NULL, *idx = NULL;
    int K = chr_to_ranges(chr_codes, &cs, &cl);
    int n_grp = list_to_csr(ref_groups, &off, &idx);
    uint8_t *st = (uint8_t *)malloc((size_t)(G * C));
    SEXP ans = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP y = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    SEXP states = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = (K < 0 || n_grp < 0 || !st) ? ICNV_E_NOMEM
                                         : icnv_smooth_hmm_u8_f64(REAL(expr), REAL(y), st, G, C, cs, cl, K, off, idx, n_grp,
                                                                  Rf_asLogical(apply_log), Rf_asReal(threshold),
                                                                  Rf_asInteger(window_length), Rf_asLogical(use_bounds), m, REAL(Pi),
                                                                  REAL(delta), REAL(mean), REAL(sd));
    if (rc == 0) {
        double *out = REAL(states);   /* the reference keeps states as doubles (HMM.R:320) */
        for (int64_t i = 0; i < G * C; ++i) out[i] = (st[i] == 255) ? -1.0 : (double)st[i];
    }
    SET_VECTOR_ELT(ans, 0, y);
    SET_VECTOR_ELT(ans, 1, states);
    free(cs); free(cl); free(off); free(idx); free(st);
    UNPROTECT(3);
    fail_if(rc);
    return ans;
}

/* icnv_init_devices: single-process multi-GPU (infercnv::run() is one R process).  ids = NULL: every GPU present. */
SEXP icnvR_init_devices(SEXP ids) {
    int rc = Rf_isNull(ids) ? icnv_init_devices(0, NULL) : icnv_init_devices(Rf_length(ids), INTEGER(ids));
    fail_if(rc);
    return Rf_ScalarInteger(icnv_devices_in_use());
}

/* predict_CNV_via_HMM_on_* / i3HMM_predict_* (HMM.R:284-567, i3HMM.R:180-389): groups = NULL -> per cell */
SEXP icnvR_viterbi(SEXP expr, SEXP chr_codes, SEXP groups, SEXP Pi, SEXP delta, SEXP mean, SEXP sd) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int m = Rf_length(delta);
    int32_t *cs = NULL, *cl = NULL, *off = NULL, *idx = NULL;
    int K = chr_to_ranges(chr_codes, &cs, &cl);
    int n_grp = Rf_isNull(groups) ? 0 : list_to_csr(groups, &off, &idx);
    /* one byte per state over PCIe (255 = cell in no group); the reference keeps states as doubles (HMM.R:320) */
    uint8_t *st = (uint8_t *)malloc((size_t)(G * C));
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = (K < 0 || n_grp < 0 || !st) ? ICNV_E_NOMEM
                                         : icnv_viterbi_u8_f64(REAL(expr), G, C, cs, cl, K, off, idx, n_grp, m, REAL(Pi),
                                                               REAL(delta), REAL(mean), REAL(sd), st, NULL);
    if (rc == 0) {
        double *out = REAL(ans);
        for (int64_t i = 0; i < G * C; ++i) out[i] = (st[i] == 255) ? -1.0 : (double)st[i];
    }
    free(cs); free(cl); free(off); free(idx); free(st);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* apply_median_filtering (noise_reduction.R:43-89) */
SEXP icnvR_median_filter(SEXP expr, SEXP chr_codes, SEXP index_lists, SEXP window_size) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int32_t *cs = NULL, *cl = NULL, *off = NULL, *idx = NULL;
    int K = chr_to_ranges(chr_codes, &cs, &cl);
    int n_grp = list_to_csr(index_lists, &off, &idx);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = (K < 0 || n_grp < 0) ? ICNV_E_NOMEM
                                  : icnv_median_filter_f64(REAL(expr), REAL(ans), G, C, cs, cl, K, off, idx, n_grp,
                                                           Rf_asInteger(window_size));
    free(cs); free(cl); free(off); free(idx);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* .i3HMM_get_sd_trend_by_num_cells_fit: mu, sigma over the listed cells (i3HMM.R:17-30) */
SEXP icnvR_mean_sd(SEXP expr, SEXP cells) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int n = Rf_length(cells);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    double mu = 0, sg = 0;
    int rc = ICNV_E_NOMEM;
    if (idx) {
        for (int i = 0; i < n; ++i) idx[i] = INTEGER(cells)[i] - 1;
        rc = icnv_mean_sd_f64(REAL(expr), G, C, idx, n, &mu, &sg);
    }
    free(idx);
    fail_if(rc);
    SEXP ans = PROTECT(Rf_allocVector(REALSXP, 2));
    REAL(ans)[0] = mu;
    REAL(ans)[1] = sg;
    UNPROTECT(1);
    return ans;
}

/* parallelDist(x) (method "euclidean") for x = observations x variables, the t(expr.data[, cells]) every hclust() call site
 * of the reference builds (R/inferCNV_tumor_subclusters.R:191, R/inferCNV_ops.R:1930, ...): the bare "dist" vector, the R
 * closure attaches Size / Labels / class */
SEXP icnvR_pairwise_dist(SEXP x) {
    SEXP dim = Rf_getAttrib(x, R_DimSymbol);
    int64_t n = INTEGER(dim)[0], G = INTEGER(dim)[1];
    SEXP ans = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)(n * (n - 1) / 2)));
    int rc = icnv_pairwise_dist_rows_f64(REAL(x), n, G, REAL(ans));
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* normalize_counts_by_seq_depth (ops.R:3064-3111); normalize_factor NA -> median of colSums */
SEXP icnvR_normalize(SEXP expr, SEXP normalize_factor) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    double nf = Rf_asReal(normalize_factor);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = icnv_normalize_counts_by_seq_depth_f64(REAL(expr), REAL(ans), G, C, (nf == nf) ? nf : -1.0);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* clear_noise_via_ref_mean_sd (ops.R:2302-2346), noise_logistic = FALSE */
SEXP icnvR_clear_noise(SEXP expr, SEXP cells, SEXP sd_amplifier) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int n = Rf_length(cells);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = ICNV_E_NOMEM;
    if (idx) {
        for (int i = 0; i < n; ++i) idx[i] = INTEGER(cells)[i] - 1;
        rc = icnv_clear_noise_via_ref_mean_sd_f64(REAL(expr), REAL(ans), G, C, idx, n, Rf_asReal(sd_amplifier));
    }
    free(idx);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* get_predicted_CNV_regions (HMM.R:706-764): consensus state per cell group, run-length regions per chromosome and
 * their bounds in one call.  `states` is the numeric state matrix of the HMM object (-1 = unassigned); it crosses
 * PCIe as one byte per entry.  Returns list(seq, chr, first_gene, last_gene, state, start, end), one entry per
 * region in the reference's numbering order, indices 1-based. */
SEXP icnvR_cnv_regions(SEXP states, SEXP chr_codes, SEXP gene_start, SEXP gene_stop, SEXP groups) {
    SEXP dim = Rf_getAttrib(states, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int32_t *cs = NULL, *cl = NULL, *off = NULL, *idx = NULL;
    int K = chr_to_ranges(chr_codes, &cs, &cl);
    int n_grp = list_to_csr(groups, &off, &idx);
    uint8_t *st = (uint8_t *)malloc((size_t)(G * C));
    int64_t n = 0;
    int not_states = 0;
    int rc = (K < 0 || n_grp < 0 || !st) ? ICNV_E_NOMEM : 0;
    if (rc == 0) {
        const double *x = REAL(states);
        for (int64_t i = 0; i < G * C && !not_states; ++i) {
            if (x[i] == -1.0) st[i] = 255;
            else if (x[i] >= 0.0 && x[i] <= 6.0 && x[i] == (double)(int)x[i]) st[i] = (uint8_t)x[i];
            else not_states = 1;            /* not a state matrix: the R wrapper falls back to the reference code */
        }
    }
    if (rc == 0 && !not_states)
        rc = icnv_predicted_cnv_regions_u8(st, G, C, cs, cl, K, REAL(gene_start), REAL(gene_stop), off, idx, n_grp, NULL, &n);
    free(cs); free(cl); free(off); free(idx); free(st);
    if (not_states) Rf_error("infercnv_b200: expr.data does not hold HMM states (-1, 0..6)");
    fail_if(rc);
    SEXP ans = PROTECT(Rf_allocVector(VECSXP, 7));
    SEXP v[7];
    for (int k = 0; k < 7; ++k) {
        v[k] = Rf_allocVector(k < 4 ? INTSXP : REALSXP, (long)n);
        SET_VECTOR_ELT(ans, k, v[k]);
    }
    int32_t *state = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    rc = !state ? ICNV_E_NOMEM
                : icnv_cnv_regions_fetch(n, INTEGER(v[0]), INTEGER(v[1]), INTEGER(v[2]), INTEGER(v[3]), state, REAL(v[5]),
                                         REAL(v[6]));
    if (rc == 0) {
        for (int64_t i = 0; i < n; ++i) {
            for (int k = 0; k < 4; ++k) INTEGER(v[k])[i] += 1;      /* R is 1-based */
            REAL(v[4])[i] = (double)state[i];                      /* states stay numeric, as in @expr.data */
        }
    }
    free(state);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* normalize_counts_by_seq_depth on a dgCMatrix (R/inferCNV.R:158-160 accepts one): @p, @i, @x (0-based, as the
 * Matrix package stores them) -> dense depth-normalised G x C matrix; 12 bytes per stored count over PCIe. */
SEXP icnvR_csc_normalize(SEXP p, SEXP i, SEXP x, SEXP dims, SEXP normalize_factor) {
    int64_t G = INTEGER(dims)[0], C = INTEGER(dims)[1];
    double nf = Rf_asReal(normalize_factor);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = (Rf_length(p) != C + 1) ? ICNV_E_BAD_ARG
                                     : icnv_csc_normalize_f64(INTEGER(p), INTEGER(i), REAL(x), G, C, NULL, 0,
                                                              (nf == nf) ? nf : -1.0, REAL(ans), NULL);
    UNPROTECT(1);
    if (rc == ICNV_E_BAD_ARG && Rf_length(p) != C + 1) Rf_error("infercnv_b200: @p does not match the matrix dimensions");
    fail_if(rc);
    return ans;
}

/* remove_outliers_norm (ops.R:1969-2056): NA bounds select "average_bound" */
SEXP icnvR_remove_outliers(SEXP expr, SEXP lower_bound, SEXP upper_bound) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = icnv_remove_outliers_norm_f64(REAL(expr), REAL(ans), G, C, Rf_asReal(lower_bound), Rf_asReal(upper_bound), NULL);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* clear_noise (ops.R:2232-2263): cells = reference cells (1-based) or NULL for "all data" */
SEXP icnvR_clear_noise_threshold(SEXP expr, SEXP cells, SEXP threshold, SEXP noise_logistic) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int n = Rf_isNull(cells) ? 0 : Rf_length(cells);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = ICNV_E_NOMEM;
    if (idx) {
        for (int i = 0; i < n; ++i) idx[i] = INTEGER(cells)[i] - 1;
        rc = icnv_clear_noise_f64(REAL(expr), REAL(ans), G, C, idx, n, Rf_asReal(threshold), Rf_asLogical(noise_logistic));
    }
    free(idx);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* predict_CNV_via_HMM_on_tumor_subclusters_per_chr (HMM.R:412-487): groups = the subclusters of all chromosomes, one
 * after the other (list of 1-based integer vectors), chr_grp_off = where each chromosome's subclusters start (0-based
 * offsets into that list, K + 1 entries), sds = m values per subcluster.  The per-chromosome traces and the final
 * per-subcluster consensus (consensus_groups: the tumour subclusters) are both done in the library. */
SEXP icnvR_viterbi_per_chr(SEXP expr, SEXP chr_codes, SEXP groups, SEXP chr_grp_off, SEXP Pi, SEXP delta, SEXP mean, SEXP sds,
                           SEXP consensus_groups) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    int m = Rf_length(delta);
    int32_t *cs = NULL, *cl = NULL, *off = NULL, *idx = NULL, *coff = NULL, *cidx = NULL;
    int K = chr_to_ranges(chr_codes, &cs, &cl);
    int n_grp = list_to_csr(groups, &off, &idx);
    int n_cons = list_to_csr(consensus_groups, &coff, &cidx);
    uint8_t *st = (uint8_t *)malloc((size_t)(G * C));
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = (K < 0 || n_grp < 0 || n_cons < 0 || !st || Rf_length(chr_grp_off) != K + 1) ? ICNV_E_NOMEM : 0;
    if (rc == 0)
        rc = icnv_viterbi_per_chr_u8_f64(REAL(expr), G, C, cs, cl, K, INTEGER(chr_grp_off), off, idx, m, REAL(Pi), REAL(delta),
                                         REAL(mean), REAL(sds), st);
    if (rc == 0 && n_cons > 0) rc = icnv_apply_state_consensus_u8(st, G, C, cs, cl, K, coff, cidx, n_cons, st);
    if (rc == 0) {
        double *out = REAL(ans);
        for (int64_t i = 0; i < G * C; ++i) out[i] = st[i] == 255 ? -1.0 : (double)st[i];
    }
    free(cs); free(cl); free(off); free(idx); free(coff); free(cidx); free(st);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

/* scale_infercnv_expr (ops.R:3174-3186) */
SEXP icnvR_scale(SEXP expr) {
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    int64_t G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
    int rc = icnv_scale_infercnv_expr_f64(REAL(expr), REAL(ans), G, C);
    UNPROTECT(1);
    fail_if(rc);
    return ans;
}

SEXP icnvR_available(void) { return Rf_ScalarLogical(icnv_device_count() > 0 && icnv_init(-1) == 0); }

static const R_CallMethodDef call_methods[] = {
    {"icnvR_subtract_ref", (DL_FUNC)&icnvR_subtract_ref, 4}, {"icnvR_smooth", (DL_FUNC)&icnvR_smooth, 3},
    {"icnvR_center", (DL_FUNC)&icnvR_center, 2},             {"icnvR_smooth_block", (DL_FUNC)&icnvR_smooth_block, 7},
    {"icnvR_viterbi", (DL_FUNC)&icnvR_viterbi, 7},           {"icnvR_median_filter", (DL_FUNC)&icnvR_median_filter, 4},
    {"icnvR_mean_sd", (DL_FUNC)&icnvR_mean_sd, 2},           {"icnvR_available", (DL_FUNC)&icnvR_available, 0},
    {"icnvR_normalize", (DL_FUNC)&icnvR_normalize, 2},       {"icnvR_clear_noise", (DL_FUNC)&icnvR_clear_noise, 3},
    {"icnvR_cnv_regions", (DL_FUNC)&icnvR_cnv_regions, 5},   {"icnvR_csc_normalize", (DL_FUNC)&icnvR_csc_normalize, 5},
    {"icnvR_remove_outliers", (DL_FUNC)&icnvR_remove_outliers, 3},
    {"icnvR_viterbi_per_chr", (DL_FUNC)&icnvR_viterbi_per_chr, 9}, {"icnvR_scale", (DL_FUNC)&icnvR_scale, 1},
    {"icnvR_clear_noise_threshold", (DL_FUNC)&icnvR_clear_noise_threshold, 4},
    {"icnvR_smooth_hmm", (DL_FUNC)&icnvR_smooth_hmm, 11},   {"icnvR_init_devices", (DL_FUNC)&icnvR_init_devices, 1},
    {"icnvR_pairwise_dist", (DL_FUNC)&icnvR_pairwise_dist, 1},
    {NULL, NULL, 0}};

void R_init_infercnvb200_shim(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

void R_unload_infercnvb200_shim(DllInfo *dll) {
    (void)dll;
    icnv_shutdown();
}
