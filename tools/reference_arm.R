#!/usr/bin/env Rscript
## reference_arm.R - bench.py's `--impl reference` arm when the box has R: times the REFERENCE's own functions for the hot
## path (source()d from its R/ directory; no package install, SURVEY.md section 8c) on the sample bench.py wrote:
##   subtract_ref_expr_from_obs -> apply_max_threshold_bounds -> smooth_by_chromosome(101) -> center(median) ->
##   subtract_ref_expr_from_obs -> invert_log2, then the per-cell Viterbi (i6 or i3) and, for c4, .median_filter blocks.
## Prints one JSON line {"cell_genes_per_s": ..., "ms_per_step": ...}.  The path ignores num_threads: 1 core.
## NOT EXECUTED in the build image (no R there); bench.py falls back to the C port when this script fails.
args <- commandArgs(trailingOnly = TRUE)
dir <- args[1]
kv <- strsplit(readLines(file.path(dir, "meta.txt")), "=", fixed = TRUE)
meta <- stats::setNames(lapply(kv, `[`, 2), sapply(kv, `[`, 1))
refdir <- meta$rdir
flog.info <- flog.debug <- flog.warn <- flog.error <- function(...) invisible()
C_CHR <- "chr"
infercnv <- methods::setClass("infercnv", slots = c(expr.data = "ANY", count.data = "ANY", gene_order = "data.frame",
                              reference_grouped_cell_indices = "list", observation_grouped_cell_indices = "list",
                              tumor_subclusters = "ANY", options = "list", .hspike = "ANY"))
src <- function(f, drop = NULL) {
    lines <- readLines(file.path(refdir, f))
    if (!is.null(drop)) lines <- lines[!grepl(drop, lines, fixed = TRUE)]
    eval(parse(text = lines), envir = globalenv())
}
src("inferCNV_ops.R"); src("inferCNV_HMM.R", "HiddenMarkov:::makedensity"); src("inferCNV_i3HMM.R"); src("noise_reduction.R")

G <- as.integer(meta$G); C <- as.integer(meta$C); steps <- as.integer(meta$steps)
x <- matrix(readBin(file.path(dir, "x.bin"), "double", n = G * C, size = 8, endian = "little"), nrow = G, ncol = C)
rownames(x) <- paste0("g", seq_len(G)); colnames(x) <- paste0("c", seq_len(C))
chr_len <- scan(file.path(dir, "chr_len.txt"), quiet = TRUE)
gene_order <- data.frame(chr = factor(rep(paste0("chr", seq_along(chr_len)), chr_len), levels = paste0("chr", seq_along(chr_len))),
                         start = seq_len(G), stop = seq_len(G) + 1L, row.names = rownames(x))
read_lists <- function(f) lapply(strsplit(readLines(file.path(dir, f)), " ", fixed = TRUE), as.integer)
refs <- read_lists("refs.txt"); names(refs) <- paste0("ref", seq_along(refs))
obs_cells <- setdiff(seq_len(C), unlist(refs))
obj0 <- new("infercnv", expr.data = x, count.data = x, gene_order = gene_order, reference_grouped_cell_indices = refs,
            observation_grouped_cell_indices = list(obs = obs_cells), tumor_subclusters = NULL, options = list(), .hspike = NULL)
i6 <- list(mean = scan(file.path(dir, "i6_mean.txt"), quiet = TRUE), sd = scan(file.path(dir, "i6_sd.txt"), quiet = TRUE))
t <- 1e-6
step <- function() {
    o <- log2xplus1(obj0)
    o <- subtract_ref_expr_from_obs(o, inv_log = FALSE, use_bounds = TRUE)
    o <- apply_max_threshold_bounds(o, threshold = 3)
    o <- smooth_by_chromosome(o, window_length = 101, smooth_ends = TRUE)
    o <- center_cell_expr_across_chromosome(o, method = "median")
    o <- subtract_ref_expr_from_obs(o, inv_log = FALSE, use_bounds = TRUE)
    o <- invert_log2(o)
    m <- o@expr.data
    chr_of <- as.integer(gene_order$chr)
    if (meta$hmm == "i6") {
        Pi <- matrix(t, 6, 6); diag(Pi) <- 1 - 5 * t; delta <- c(t, t, 1 - 5 * t, t, t, t); pm <- i6
    } else {
        v <- as.vector(m[, unlist(refs)]); mu <- mean(v); sg <- sd(v); d <- abs(qnorm(0.05, 0, sg))
        Pi <- matrix(t, 3, 3); diag(Pi) <- 1 - 5 * t; delta <- c(t, 1 - 5 * t, t); pm <- list(mean = c(mu - d, mu, mu + d), sd = rep(sg, 3))
    }
    for (k in seq_along(chr_len)) {
        idx <- which(chr_of == k)
        for (cc in seq_len(C)) Viterbi.dthmm.adj(list(x = m[idx, cc], Pi = Pi, delta = delta, distn = "norm", pm = pm))
    }
    if (meta$median_filter == "1") {
        for (l in read_lists("lists.txt")) for (k in seq_along(chr_len)) {
            idx <- which(chr_of == k)
            .median_filter(data = m[idx, l, drop = FALSE], window_size = 7, half_window = 3)
        }
    }
    invisible(NULL)
}
t0 <- proc.time()[["elapsed"]]
for (s in seq_len(steps)) step()
dt <- (proc.time()[["elapsed"]] - t0) / steps
cat(sprintf('{"cell_genes_per_s": %.6g, "ms_per_step": %.6g}\n', G * C / dt, dt * 1e3))
