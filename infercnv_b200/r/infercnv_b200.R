## infercnv_b200.R - drop-in GPU replacements for inferCNV's smoothing + HMM hot path.
##
## Usage (on a box with R, the infercnv package and a B200):
##   R CMD SHLIB infercnvb200_shim.c -I<repo>/include -L<repo>/infercnv_b200 -linfercnv_b200
##   dyn.load("<repo>/infercnv_b200/libinfercnv_b200.so"); dyn.load("infercnvb200_shim.so")
##   source("infercnv_b200.R"); infercnvb200_install()
##   infercnv::run(...)            # unchanged call; steps 8, 10, 11, 12, 17 now run on the GPU
##
## Each replacement has the SAME name and signature as the internal infercnv function it replaces
## (reference file:line in the comments) and keeps a handle to the original, which it calls whenever
## the input is outside the validated envelope (see .icnv_ok) or the library reports an error.
## The S4 object, run() and every other step are untouched.  GPU use can be switched off with
## options(infercnv.b200 = FALSE).

.icnv_env <- new.env()

## STATUS: experimental.  The shim is compile-checked against a mock of Rinternals.h and the same C entry points are
## exercised through ctypes by the repository's GPU tests, but this R layer has not been executed under R yet (no R in the
## build image); see INTEGRATION.md for the check list of a first run.

.icnv_enabled <- function() {
    if (!isTRUE(getOption("infercnv.b200", TRUE))) return(FALSE)
    ok <- .icnv_try(.Call("icnvR_available"), error = function(e) {
        if (is.null(.icnv_env$warned_unavailable)) {     # once per session: a broken GPU install must not be invisible
            .icnv_env$warned_unavailable <- TRUE
            futile.logger::flog.warn(sprintf("infercnv_b200: GPU path unavailable (%s); every step runs the R implementation", conditionMessage(e)))
        }
        FALSE
    })
    if (!isTRUE(ok) && is.null(.icnv_env$warned_unavailable)) {
        .icnv_env$warned_unavailable <- TRUE
        futile.logger::flog.warn("infercnv_b200: no usable CUDA device; every step runs the R implementation")
    }
    isTRUE(ok)
}

## one .Call into the library; an error is LOGGED (never swallowed silently) and turned into NULL = "use the R original"
.icnv_try <- function(expr, what) {
    tryCatch(expr, error = function(e) {
        futile.logger::flog.warn(sprintf("infercnv_b200: %s failed on the GPU (%s); falling back to the R implementation (hours at scale)",
                                         what, conditionMessage(e)))
        NULL
    })
}

## Single-process multi-GPU: infercnv::run() is one R process and its num_threads never reaches this path
## (R/inferCNV_ops.R:388); after this call the library shards the cells of the streaming entry points over the devices.
infercnvb200_init_devices <- function(device_ids = NULL) {
    .Call("icnvR_init_devices", if (is.null(device_ids)) NULL else as.integer(device_ids))
}

## envelope: base dense double matrix without NA/NaN/Inf (the R smoother strips NAs per cell,
## ops.R:2487-2489; expr.data may also be a dgCMatrix, R/inferCNV.R:40)
.icnv_ok <- function(m) {
    is.matrix(m) && is.double(m) && !anyNA(m) && all(is.finite(range(m)))
}

.icnv_chr_codes <- function(infercnv_obj) {
    chr <- infercnv_obj@gene_order[["chr"]]
    codes <- as.integer(factor(chr, levels = unique(chr)))   # rows are pre-sorted by chr (R/inferCNV.R:407-413)
    if (is.unsorted(codes)) return(NULL)
    codes
}

.icnv_ref_groups <- function(infercnv_obj) {
    if (length(infercnv_obj@reference_grouped_cell_indices) > 0) {
        lapply(infercnv_obj@reference_grouped_cell_indices, as.integer)
    } else {                                               # ops.R:1686-1689: proxy group of all observations
        list(proxyNormal = as.integer(unlist(infercnv_obj@observation_grouped_cell_indices)))
    }
}

.icnv_keep_names <- function(new, old) { dimnames(new) <- dimnames(old); new }

## subtract_ref_expr_from_obs, R/inferCNV_ops.R:1678
b200_subtract_ref_expr_from_obs <- function(infercnv_obj, inv_log=FALSE, use_bounds=TRUE) {
    orig <- .icnv_env$orig$subtract_ref_expr_from_obs
    m <- infercnv_obj@expr.data
    if (!.icnv_enabled() || !.icnv_ok(m)) return(orig(infercnv_obj, inv_log=inv_log, use_bounds=use_bounds))
    futile.logger::flog.info(sprintf("::subtract_ref_expr_from_obs:Start inv_log=%s, use_bounds=%s (B200)", inv_log, use_bounds))
    res <- tryCatch(.Call("icnvR_subtract_ref", m, .icnv_ref_groups(infercnv_obj), inv_log, use_bounds), "available")
    if (is.null(res)) return(orig(infercnv_obj, inv_log=inv_log, use_bounds=use_bounds))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    if (!is.null(infercnv_obj@.hspike)) {
        futile.logger::flog.info("-mirroring for hspike")
        infercnv_obj@.hspike <- b200_subtract_ref_expr_from_obs(infercnv_obj@.hspike, inv_log=inv_log, use_bounds=use_bounds)
    }
    infercnv_obj
}

## smooth_by_chromosome, R/inferCNV_ops.R:2406
b200_smooth_by_chromosome <- function(infercnv_obj, window_length, smooth_ends=TRUE) {
    orig <- .icnv_env$orig$smooth_by_chromosome
    m <- infercnv_obj@expr.data
    codes <- .icnv_chr_codes(infercnv_obj)
    ## even or < 2 windows: the reference's behaviour is accidental / identity - leave it to R
    if (!.icnv_enabled() || !.icnv_ok(m) || is.null(codes) || window_length < 2 || window_length %% 2 == 0)
        return(orig(infercnv_obj, window_length, smooth_ends))
    res <- .icnv_try(.Call("icnvR_smooth", m, codes, as.integer(window_length)), "smooth")
    if (is.null(res)) return(orig(infercnv_obj, window_length, smooth_ends))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    if (!is.null(infercnv_obj@.hspike)) {
        futile.logger::flog.info("-mirroring for hspike")
        infercnv_obj@.hspike <- b200_smooth_by_chromosome(infercnv_obj@.hspike, window_length, smooth_ends)
    }
    infercnv_obj
}

## center_cell_expr_across_chromosome, R/inferCNV_ops.R:2074
b200_center_cell_expr_across_chromosome <- function(infercnv_obj, method="mean") {
    orig <- .icnv_env$orig$center_cell_expr_across_chromosome
    m <- infercnv_obj@expr.data
    if (!.icnv_enabled() || !.icnv_ok(m)) return(orig(infercnv_obj, method))
    futile.logger::flog.info("::center_smooth across chromosomes per cell (B200)")
    res <- .icnv_try(.Call("icnvR_center", m, identical(method, "median")), "center")
    if (is.null(res)) return(orig(infercnv_obj, method))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    if (!is.null(infercnv_obj@.hspike)) {
        futile.logger::flog.info("-mirroring for hspike")
        infercnv_obj@.hspike <- b200_center_cell_expr_across_chromosome(infercnv_obj@.hspike, method)
    }
    infercnv_obj
}

## Viterbi for a set of (group or cell) sequences; used by the six drivers below
.icnv_hmm <- function(infercnv_obj, HMM_info, groups = NULL, sds = NULL) {
    m <- infercnv_obj@expr.data
    codes <- .icnv_chr_codes(infercnv_obj)
    if (!.icnv_enabled() || !.icnv_ok(m) || is.null(codes)) return(NULL)
    sd <- if (is.null(sds)) HMM_info[["state_emission_params"]]$sd else sds
    res <- .icnv_try(.Call("icnvR_viterbi", m, codes, groups, HMM_info[["state_transitions"]], HMM_info[["delta"]],
                          HMM_info[["state_emission_params"]]$mean, as.double(sd)), "viterbi")
    if (is.null(res)) return(NULL)
    .icnv_keep_names(res, m)
}

## predict_CNV_via_HMM_on_indiv_cells, R/inferCNV_HMM.R:284
b200_predict_CNV_via_HMM_on_indiv_cells <- function(infercnv_obj, cnv_mean_sd=infercnv:::get_spike_dists(infercnv_obj@.hspike), t=1e-6) {
    res <- .icnv_hmm(infercnv_obj, infercnv:::.get_HMM(cnv_mean_sd, t))
    if (is.null(res)) return(.icnv_env$orig$predict_CNV_via_HMM_on_indiv_cells(infercnv_obj, cnv_mean_sd, t))
    infercnv_obj@expr.data <- res
    infercnv_obj
}

## predict_CNV_via_HMM_on_tumor_subclusters, R/inferCNV_HMM.R:345
b200_predict_CNV_via_HMM_on_tumor_subclusters <- function(infercnv_obj,
        cnv_mean_sd=infercnv:::get_spike_dists(infercnv_obj@.hspike),
        cnv_level_to_mean_sd_fit=infercnv:::get_hspike_cnv_mean_sd_trend_by_num_cells_fit(infercnv_obj@.hspike), t=1e-6) {
    orig <- .icnv_env$orig$predict_CNV_via_HMM_on_tumor_subclusters
    if (is.null(infercnv_obj@tumor_subclusters)) return(orig(infercnv_obj, cnv_mean_sd, cnv_level_to_mean_sd_fit, t))
    groups <- lapply(unlist(infercnv_obj@tumor_subclusters[["subclusters"]], recursive=FALSE), as.integer)
    sds <- unlist(lapply(groups, function(g)                       # .get_state_emission_params, HMM.R:586-614
        infercnv:::.get_state_emission_params(length(g), cnv_mean_sd, cnv_level_to_mean_sd_fit)$sd))
    res <- .icnv_hmm(infercnv_obj, infercnv:::.get_HMM(cnv_mean_sd, t), groups, sds)
    if (is.null(res)) return(orig(infercnv_obj, cnv_mean_sd, cnv_level_to_mean_sd_fit, t))
    infercnv_obj@expr.data <- res
    infercnv_obj
}

## predict_CNV_via_HMM_on_whole_tumor_samples, R/inferCNV_HMM.R:509
b200_predict_CNV_via_HMM_on_whole_tumor_samples <- function(infercnv_obj, cluster_by_groups,
        cnv_mean_sd=infercnv:::get_spike_dists(infercnv_obj@.hspike),
        cnv_level_to_mean_sd_fit=infercnv:::get_hspike_cnv_mean_sd_trend_by_num_cells_fit(infercnv_obj@.hspike), t=1e-6) {
    obs <- infercnv_obj@observation_grouped_cell_indices
    ## cluster_by_groups = FALSE: the reference writes c(all_observations = unlist(obs), <list of reference groups>)
    ## (R/inferCNV_HMM.R:531, R/inferCNV_i3HMM.R:356); c() of an integer vector and a list makes every observation cell a
    ## list element of its own, i.e. each observation cell is a one-cell "sample" (rowMeans = the cell, sd for num_cells = 1)
    groups <- c(if (isTRUE(cluster_by_groups)) obs else as.list(unlist(obs, use.names=FALSE)),
                infercnv_obj@reference_grouped_cell_indices)
    groups <- lapply(groups, as.integer)
    sds <- unlist(lapply(groups, function(g)
        infercnv:::.get_state_emission_params(length(g), cnv_mean_sd, cnv_level_to_mean_sd_fit)$sd))
    res <- .icnv_hmm(infercnv_obj, infercnv:::.get_HMM(cnv_mean_sd, t), groups, sds)
    if (is.null(res))
        return(.icnv_env$orig$predict_CNV_via_HMM_on_whole_tumor_samples(infercnv_obj, cluster_by_groups, cnv_mean_sd,
                                                                         cnv_level_to_mean_sd_fit, t))
    infercnv_obj@expr.data <- res
    infercnv_obj
}

## i3HMM_predict_CNV_via_HMM_on_indiv_cells, R/inferCNV_i3HMM.R:180 (sd_trend stays R's; group twins below)
b200_i3HMM_predict_CNV_via_HMM_on_indiv_cells <- function(infercnv_obj, i3_p_val=0.05,
        sd_trend=infercnv:::.i3HMM_get_sd_trend_by_num_cells_fit(infercnv_obj, i3_p_val), t=1e-6, use_KS=TRUE) {
    res <- .icnv_hmm(infercnv_obj, infercnv:::.i3HMM_get_HMM(sd_trend, t=t, i3_p_val=i3_p_val, use_KS=use_KS))
    if (is.null(res)) return(.icnv_env$orig$i3HMM_predict_CNV_via_HMM_on_indiv_cells(infercnv_obj, i3_p_val, sd_trend, t, use_KS))
    infercnv_obj@expr.data <- res
    infercnv_obj
}

## i3HMM_predict_CNV_via_HMM_on_tumor_subclusters, R/inferCNV_i3HMM.R:249: one trace per subcluster on rowMeans, the
## same three-state parameters for every group (no group-size dependent sd in the i3 model)
b200_i3HMM_predict_CNV_via_HMM_on_tumor_subclusters <- function(infercnv_obj, i3_p_val=0.05,
        sd_trend=infercnv:::.i3HMM_get_sd_trend_by_num_cells_fit(infercnv_obj, i3_p_val), t=1e-6, use_KS=TRUE) {
    orig <- .icnv_env$orig$i3HMM_predict_CNV_via_HMM_on_tumor_subclusters
    if (is.null(infercnv_obj@tumor_subclusters)) return(orig(infercnv_obj, i3_p_val, sd_trend, t, use_KS))
    groups <- lapply(unlist(infercnv_obj@tumor_subclusters[["subclusters"]], recursive=FALSE), as.integer)
    HMM_info <- infercnv:::.i3HMM_get_HMM(sd_trend, t=t, i3_p_val=i3_p_val, use_KS=use_KS)
    res <- .icnv_hmm(infercnv_obj, HMM_info, groups, rep(HMM_info[["state_emission_params"]]$sd, length(groups)))
    if (is.null(res)) return(orig(infercnv_obj, i3_p_val, sd_trend, t, use_KS))
    infercnv_obj@expr.data <- res
    infercnv_obj
}

## i3HMM_predict_CNV_via_HMM_on_whole_tumor_samples, R/inferCNV_i3HMM.R:332
b200_i3HMM_predict_CNV_via_HMM_on_whole_tumor_samples <- function(infercnv_obj, cluster_by_groups, i3_p_val=0.05,
        sd_trend=infercnv:::.i3HMM_get_sd_trend_by_num_cells_fit(infercnv_obj, i3_p_val), t=1e-6, use_KS=TRUE) {
    obs <- infercnv_obj@observation_grouped_cell_indices
    ## cluster_by_groups = FALSE: the reference writes c(all_observations = unlist(obs), <list of reference groups>)
    ## (R/inferCNV_HMM.R:531, R/inferCNV_i3HMM.R:356); c() of an integer vector and a list makes every observation cell a
    ## list element of its own, i.e. each observation cell is a one-cell "sample" (rowMeans = the cell, sd for num_cells = 1)
    groups <- c(if (isTRUE(cluster_by_groups)) obs else as.list(unlist(obs, use.names=FALSE)),
                infercnv_obj@reference_grouped_cell_indices)
    groups <- lapply(groups, as.integer)
    HMM_info <- infercnv:::.i3HMM_get_HMM(sd_trend, t=t, i3_p_val=i3_p_val, use_KS=use_KS)
    res <- .icnv_hmm(infercnv_obj, HMM_info, groups, rep(HMM_info[["state_emission_params"]]$sd, length(groups)))
    if (is.null(res))
        return(.icnv_env$orig$i3HMM_predict_CNV_via_HMM_on_whole_tumor_samples(infercnv_obj, cluster_by_groups, i3_p_val,
                                                                               sd_trend, t, use_KS))
    infercnv_obj@expr.data <- res
    infercnv_obj
}

## apply_median_filtering, R/noise_reduction.R:43 (exported)
b200_apply_median_filtering <- function(infercnv_obj, window_size=7, on_observations=TRUE, on_references=TRUE) {
    orig <- .icnv_env$orig$apply_median_filtering
    m <- infercnv_obj@expr.data
    codes <- .icnv_chr_codes(infercnv_obj)
    if (!.icnv_enabled() || !.icnv_ok(m) || is.null(codes) || window_size %% 2 != 1 || window_size < 3)
        return(orig(infercnv_obj, window_size, on_observations, on_references))
    lists <- list()
    if (on_observations) for (tt in names(infercnv_obj@observation_grouped_cell_indices))
        lists <- c(lists, lapply(infercnv_obj@tumor_subclusters[["subclusters"]][[tt]], as.integer))
    if (on_references) lists <- c(lists, lapply(infercnv_obj@reference_grouped_cell_indices, as.integer))
    res <- .icnv_try(.Call("icnvR_median_filter", m, codes, lists, as.integer(window_size)), "median_filter")
    if (is.null(res)) return(orig(infercnv_obj, window_size, on_observations, on_references))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    infercnv_obj
}

## normalize_counts_by_seq_depth, R/inferCNV_ops.R:3064
b200_normalize_counts_by_seq_depth <- function(infercnv_obj, normalize_factor=NA) {
    m <- infercnv_obj@expr.data
    if (.icnv_enabled() && methods::is(m, "dgCMatrix") && !anyNA(m@x)) {
        ## sparse counts (R/inferCNV.R:158-160): the compressed columns cross PCIe, the dense normalised matrix comes back
        res <- .icnv_try(.Call("icnvR_csc_normalize", m@p, m@i, as.double(m@x), dim(m), as.double(normalize_factor)), "csc_normalize")
        if (!is.null(res)) {
            futile.logger::flog.info("normalizing counts matrix by depth (B200, sparse input)")
            infercnv_obj@expr.data <- .icnv_keep_names(res, m)
            return(infercnv_obj)
        }
    }
    if (!.icnv_enabled() || !.icnv_ok(m)) return(.icnv_env$orig$normalize_counts_by_seq_depth(infercnv_obj, normalize_factor))
    res <- .icnv_try(.Call("icnvR_normalize", m, as.double(normalize_factor)), "normalize")
    if (is.null(res)) return(.icnv_env$orig$normalize_counts_by_seq_depth(infercnv_obj, normalize_factor))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    infercnv_obj
}

## clear_noise_via_ref_mean_sd, R/inferCNV_ops.R:2302 (noise_logistic=TRUE stays in R)
b200_clear_noise_via_ref_mean_sd <- function(infercnv_obj, sd_amplifier=1.5, noise_logistic=FALSE) {
    orig <- .icnv_env$orig$clear_noise_via_ref_mean_sd
    m <- infercnv_obj@expr.data
    if (!.icnv_enabled() || !.icnv_ok(m) || isTRUE(noise_logistic)) return(orig(infercnv_obj, sd_amplifier, noise_logistic))
    cells <- if (length(infercnv_obj@reference_grouped_cell_indices) > 0)
        unlist(infercnv_obj@reference_grouped_cell_indices) else unlist(infercnv_obj@observation_grouped_cell_indices)
    res <- .icnv_try(.Call("icnvR_clear_noise", m, as.integer(cells), as.double(sd_amplifier)), "clear_noise")
    if (is.null(res)) return(orig(infercnv_obj, sd_amplifier, noise_logistic))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    infercnv_obj
}

## predict_CNV_via_HMM_on_tumor_subclusters_per_chr, R/inferCNV_HMM.R:412 (Leiden + per_chr_hmm_subclusters)
b200_predict_CNV_via_HMM_on_tumor_subclusters_per_chr <- function(infercnv_obj, subclusters_per_chr,
        cnv_mean_sd=infercnv:::get_spike_dists(infercnv_obj@.hspike),
        cnv_level_to_mean_sd_fit=infercnv:::get_hspike_cnv_mean_sd_trend_by_num_cells_fit(infercnv_obj@.hspike), t=1e-6) {
    orig <- .icnv_env$orig$predict_CNV_via_HMM_on_tumor_subclusters_per_chr
    m <- infercnv_obj@expr.data
    codes <- .icnv_chr_codes(infercnv_obj)
    if (is.null(subclusters_per_chr) || !.icnv_enabled() || !.icnv_ok(m) || is.null(codes))
        return(orig(infercnv_obj, subclusters_per_chr, cnv_mean_sd, cnv_level_to_mean_sd_fit, t))
    futile.logger::flog.info("predict_CNV_via_HMM_on_tumor_subclusters_per_chr (B200)")
    chrs <- unique(infercnv_obj@gene_order$chr)
    per <- lapply(chrs, function(chr) lapply(subclusters_per_chr[[chr]], as.integer))
    flat <- unlist(per, recursive=FALSE)
    sds <- unlist(lapply(flat, function(g)
        infercnv:::.get_state_emission_params(length(g), cnv_mean_sd, cnv_level_to_mean_sd_fit)$sd))
    HMM_info <- infercnv:::.get_HMM(cnv_mean_sd, t)
    tumor_subclusters <- lapply(unlist(infercnv_obj@tumor_subclusters[["subclusters"]], recursive=FALSE), as.integer)
    res <- .icnv_try(.Call("icnvR_viterbi_per_chr", m, codes, flat, as.integer(cumsum(c(0L, lengths(per)))),
                          HMM_info[["state_transitions"]], HMM_info[["delta"]], HMM_info[["state_emission_params"]]$mean,
                          as.double(sds), tumor_subclusters), "viterbi_per_chr")
    if (is.null(res)) return(orig(infercnv_obj, subclusters_per_chr, cnv_mean_sd, cnv_level_to_mean_sd_fit, t))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    infercnv_obj
}

## scale_infercnv_expr, R/inferCNV_ops.R:3174 (run() step 5 when scale_data = TRUE)
b200_scale_infercnv_expr <- function(infercnv_obj) {
    orig <- .icnv_env$orig$scale_infercnv_expr
    m <- infercnv_obj@expr.data
    if (!.icnv_enabled() || !.icnv_ok(m)) return(orig(infercnv_obj))
    futile.logger::flog.info("-scaling expr data (B200)")
    res <- .icnv_try(.Call("icnvR_scale", m), "scale")
    if (is.null(res)) return(orig(infercnv_obj))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    if (!is.null(infercnv_obj@.hspike)) {
        futile.logger::flog.info("-mirroring for hspike")
        infercnv_obj@.hspike <- b200_scale_infercnv_expr(infercnv_obj@.hspike)
    }
    infercnv_obj
}

## remove_outliers_norm, R/inferCNV_ops.R:1969 (run() step 16)
b200_remove_outliers_norm <- function(infercnv_obj, out_method="average_bound", lower_bound=NA, upper_bound=NA) {
    orig <- .icnv_env$orig$remove_outliers_norm
    m <- infercnv_obj@expr.data
    hard <- !is.na(lower_bound) & !is.na(upper_bound)
    ## anything but hard bounds / "average_bound" ends in the reference's stop(991) / stop(992): leave it to R
    if (!.icnv_enabled() || !.icnv_ok(m) || !(hard || identical(out_method, "average_bound")))
        return(orig(infercnv_obj, out_method, lower_bound, upper_bound))
    futile.logger::flog.info(paste("::remove_outlier_norm:Start (B200)", "out_method:", out_method, "lower_bound:", lower_bound,
                                   "upper_bound:", upper_bound))
    res <- .icnv_try(.Call("icnvR_remove_outliers", m, as.double(if (hard) lower_bound else NA),
                          as.double(if (hard) upper_bound else NA)), "remove_outliers")
    if (is.null(res)) return(orig(infercnv_obj, out_method, lower_bound, upper_bound))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    if (!is.null(infercnv_obj@.hspike)) {
        futile.logger::flog.info("-mirroring for hspike")
        infercnv_obj@.hspike <- b200_remove_outliers_norm(infercnv_obj@.hspike, out_method, lower_bound, upper_bound)
    }
    infercnv_obj
}

## clear_noise, R/inferCNV_ops.R:2232 (run() step 22 with a numeric noise_filter)
b200_clear_noise <- function(infercnv_obj, threshold, noise_logistic=FALSE) {
    orig <- .icnv_env$orig$clear_noise
    m <- infercnv_obj@expr.data
    if (!.icnv_enabled() || !.icnv_ok(m) || threshold == 0) return(orig(infercnv_obj, threshold, noise_logistic))
    cells <- if (length(infercnv_obj@reference_grouped_cell_indices) > 0)
        as.integer(unlist(infercnv_obj@reference_grouped_cell_indices)) else NULL
    res <- .icnv_try(.Call("icnvR_clear_noise_threshold", m, cells, as.double(threshold), isTRUE(noise_logistic)), "clear_noise_threshold")
    if (is.null(res)) return(orig(infercnv_obj, threshold, noise_logistic))
    infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    infercnv_obj
}

## get_predicted_CNV_regions, R/inferCNV_HMM.R:706-764: consensus state per cell group, run-length regions per
## chromosome and their bounds from ONE library call; the returned list has the reference's shape
## (cell_group_name, cells, gene_regions = named list of per-gene data.frames, cnv_ranges = data.frame), so
## generate_cnv_region_reports (HMM.R:790-869) and everything downstream of its files run unchanged.
b200_get_predicted_CNV_regions <- function(infercnv_obj, by=c("consensus", "subcluster", "cell")) {
    by <- match.arg(by)
    orig <- .icnv_env$orig$get_predicted_CNV_regions
    m <- infercnv_obj@expr.data
    codes <- .icnv_chr_codes(infercnv_obj)
    go <- infercnv_obj@gene_order
    if (!.icnv_enabled() || !.icnv_ok(m) || is.null(codes) || anyNA(go$start) || anyNA(go$stop)) return(orig(infercnv_obj, by))
    futile.logger::flog.info(sprintf("get_predicted_CNV_regions(%s) (B200)", by))
    if (is.null(infercnv_obj@tumor_subclusters)) {
        futile.logger::flog.warn("get_predicted_CNV_regions() - no subclusters defined, resetting reporting mode to consensus")
        by <- "consensus"
    }
    if (by == "consensus") {                                                      # HMM.R:721-723
        cell_groups <- c(infercnv_obj@reference_grouped_cell_indices, infercnv_obj@observation_grouped_cell_indices)
    } else if (by == "subcluster") {                                              # HMM.R:724-725
        cell_groups <- unlist(infercnv_obj@tumor_subclusters[["subclusters"]], recursive=FALSE)
    } else {                                                                      # HMM.R:726-729
        cells <- c(unlist(infercnv_obj@reference_grouped_cell_indices, use.names=FALSE),
                   unlist(infercnv_obj@observation_grouped_cell_indices, use.names=FALSE))
        cell_groups <- as.list(cells)
        names(cell_groups) <- colnames(m)[cells]
    }
    res <- .icnv_try(.Call("icnvR_cnv_regions", m, codes, as.double(go$start), as.double(go$stop),
                          lapply(cell_groups, as.integer)), "cnv_regions")
    if (is.null(res)) return(orig(infercnv_obj, by))
    names(res) <- c("seq", "chr", "first_gene", "last_gene", "state", "start", "end")
    chr_levels <- unique(go$chr)                       # order of appearance = the order of the chromosome ranges
    gene_names <- rownames(go)
    cnv_name <- sprintf("%s-region_%d", as.character(chr_levels[res$chr]), seq_along(res$seq))   # running counter, :758
    if (is.integer(go$start) && is.integer(go$stop)) { res$start <- as.integer(res$start); res$end <- as.integer(res$end) }
    by_group <- split(seq_along(res$seq), factor(res$seq, levels = seq_along(cell_groups)))
    lapply(seq_along(cell_groups), function(k) {
        sel <- by_group[[k]]
        futile.logger::flog.info(sprintf("-processing cell_group_name: %s, size: %d", names(cell_groups)[k], length(cell_groups[[k]])))
        gene_regions <- lapply(sel, function(i) {
            g <- res$first_gene[i]:res$last_gene[i]
            data.frame(state = res$state[i], gene = gene_names[g], chr = go$chr[g], start = go$start[g], end = go$stop[g])
        })
        names(gene_regions) <- cnv_name[sel]
        list(cell_group_name = names(cell_groups)[k],
             cells = colnames(m)[cell_groups[[k]]],
             gene_regions = gene_regions,
             cnv_ranges = data.frame(cnv_name = cnv_name[sel], state = res$state[sel], chr = chr_levels[res$chr[sel]],
                                     start = res$start[sel], end = res$end[sel]))
    })
}

## run() steps 4, 8 .. 12, 14 (and step 17 for analysis_mode = "cells") as ONE library call: one upload of the matrix instead
## of four round trips through the per-step closures above.  infercnv::run() itself calls the steps one by one
## (R/inferCNV_ops.R:771, 865, 911, 952, 1031), so this is for callers that drive the steps themselves, or for a maintainer's
## three-line change in run() (INTEGRATION.md, "fused block").  Input: the object after step 3 (depth-normalised counts, before
## log2) when apply_log = TRUE, after step 4 otherwise; use_bounds / max_centered_threshold / window_length as in run().
## HMM_info (optional): the list .get_HMM() / .i3HMM_get_HMM() returns -> per-cell states are predicted on the block's output
## in the same pass.  Returns list(infercnv_obj = <after step 14>, hmm_obj = <states object or NULL>); NULL when the input is
## outside the GPU path's envelope (the caller then runs the R steps).
infercnvb200_run_smooth_block <- function(infercnv_obj, window_length=101, max_centered_threshold=3, use_bounds=TRUE,
                                          apply_log=TRUE, HMM_info=NULL) {
    m <- infercnv_obj@expr.data
    codes <- .icnv_chr_codes(infercnv_obj)
    if (!.icnv_enabled() || !.icnv_ok(m) || is.null(codes) || window_length < 2 || window_length %% 2 == 0 ||
        is.na(max_centered_threshold) || !is.numeric(max_centered_threshold)) return(NULL)
    futile.logger::flog.info(sprintf("::smooth block (steps 4, 8-12, 14%s) in one pass (B200)", if (is.null(HMM_info)) "" else " + per-cell HMM"))
    refs <- .icnv_ref_groups(infercnv_obj)
    hmm_obj <- NULL
    if (is.null(HMM_info)) {
        res <- .icnv_try(.Call("icnvR_smooth_block", m, codes, refs, isTRUE(apply_log), as.double(max_centered_threshold),
                               as.integer(window_length), isTRUE(use_bounds)), "smooth_block")
        if (is.null(res)) return(NULL)
        infercnv_obj@expr.data <- .icnv_keep_names(res, m)
    } else {
        res <- .icnv_try(.Call("icnvR_smooth_hmm", m, codes, refs, isTRUE(apply_log), as.double(max_centered_threshold),
                               as.integer(window_length), isTRUE(use_bounds), HMM_info[["state_transitions"]], HMM_info[["delta"]],
                               HMM_info[["state_emission_params"]]$mean, as.double(HMM_info[["state_emission_params"]]$sd)),
                         "smooth_hmm")
        if (is.null(res)) return(NULL)
        infercnv_obj@expr.data <- .icnv_keep_names(res[[1]], m)
        hmm_obj <- infercnv_obj
        hmm_obj@expr.data <- .icnv_keep_names(res[[2]], m)
    }
    if (!is.null(infercnv_obj@.hspike)) {
        futile.logger::flog.info("-mirroring for hspike")
        h <- infercnvb200_run_smooth_block(infercnv_obj@.hspike, window_length, max_centered_threshold, use_bounds, apply_log, NULL)
        if (is.null(h)) return(NULL)
        infercnv_obj@.hspike <- h$infercnv_obj
    }
    list(infercnv_obj = infercnv_obj, hmm_obj = hmm_obj)
}

## parallelDist::parallelDist as the reference calls it before every hclust() - parallelDist(t(expr[, cells]), threads = n)
## (R/inferCNV_tumor_subclusters.R:191,411,472,582,609; R/inferCNV_ops.R:1930,3242; R/inferCNV_heatmap.R:719,755,1062,1079).
## Same signature as parallelDist::parallelDist(x, method = "euclidean", diag = FALSE, upper = FALSE, threads = NULL, ...); any
## method but the default one, a non-matrix / non-double / NA-holding x, or a GPU error goes to the original.  `threads` never
## reaches the GPU path.  infercnv imports the function (R/inferCNV_constants.R:27), so the replacement is bound in the
## package's IMPORTS environment (infercnvb200_install); the "dist" object it returns is what hclust() and as.matrix() expect.
b200_parallelDist <- function(x, method = "euclidean", diag = FALSE, upper = FALSE, threads = NULL, ...) {
    orig <- .icnv_env$orig_parallelDist
    ok <- .icnv_enabled() && identical(method, "euclidean") && is.matrix(x) && is.double(x) && !anyNA(x) && nrow(x) >= 2 &&
        nrow(x) <= 65536          # hclust()'s own limit
    if (ok) {
        d <- .icnv_try(.Call("icnvR_pairwise_dist", x), "parallelDist (euclidean)")
        if (!is.null(d)) {
            attributes(d) <- list(Size = nrow(x), Labels = rownames(x), Diag = diag, Upper = upper, method = "euclidean",
                                  call = match.call(), class = "dist")
            return(d)
        }
    }
    orig(x, method = method, diag = diag, upper = upper, threads = threads, ...)
}

infercnvb200_install <- function() {
    fns <- c("subtract_ref_expr_from_obs", "smooth_by_chromosome", "center_cell_expr_across_chromosome",
             "predict_CNV_via_HMM_on_indiv_cells", "predict_CNV_via_HMM_on_tumor_subclusters",
             "predict_CNV_via_HMM_on_whole_tumor_samples", "i3HMM_predict_CNV_via_HMM_on_indiv_cells",
             "apply_median_filtering", "normalize_counts_by_seq_depth", "clear_noise_via_ref_mean_sd",
             "get_predicted_CNV_regions", "remove_outliers_norm", "clear_noise",
             "predict_CNV_via_HMM_on_tumor_subclusters_per_chr", "i3HMM_predict_CNV_via_HMM_on_tumor_subclusters",
             "i3HMM_predict_CNV_via_HMM_on_whole_tumor_samples", "scale_infercnv_expr")
    ns <- asNamespace("infercnv")
    .icnv_env$orig <- lapply(stats::setNames(fns, fns), function(f) get(f, envir = ns))
    for (f in fns) utils::assignInNamespace(f, get(paste0("b200_", f)), ns = "infercnv")
    ## parallelDist is an IMPORT of infercnv: its binding lives in the parent of the namespace ("imports:infercnv")
    imp <- parent.env(ns)
    if (exists("parallelDist", envir = imp, inherits = FALSE)) {
        .icnv_env$orig_parallelDist <- get("parallelDist", envir = imp)
        unlockBinding("parallelDist", imp)
        assign("parallelDist", b200_parallelDist, envir = imp)
        lockBinding("parallelDist", imp)
    }
    invisible(TRUE)
}

infercnvb200_uninstall <- function() {
    for (f in names(.icnv_env$orig)) utils::assignInNamespace(f, .icnv_env$orig[[f]], ns = "infercnv")
    imp <- parent.env(asNamespace("infercnv"))
    if (!is.null(.icnv_env$orig_parallelDist)) {
        unlockBinding("parallelDist", imp)
        assign("parallelDist", .icnv_env$orig_parallelDist, envir = imp)
        lockBinding("parallelDist", imp)
    }
    invisible(TRUE)
}
