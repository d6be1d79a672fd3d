#!/usr/bin/env Rscript
## make_hmm_fixture.R - run the REFERENCE's own Viterbi.dthmm.adj and .median_filter on inputs exported by
## tools/export_r_fixture_inputs.py and write their results, so that the HMM / median-filter parity of infercnv_b200 is
## pinned by real R output (SURVEY.md section 8c).  Needs only base R + stats: the hot-path functions are source()d from the
## reference checkout with futile.logger stubbed, and the one line of Viterbi.dthmm.adj that touches the HiddenMarkov
## package (makedensity, result unused, R/inferCNV_HMM.R:1109) is dropped from the sourced text.
##
##   python tools/export_r_fixture_inputs.py fixture_dir          # on any box with the repository
##   Rscript tools/make_hmm_fixture.R /path/to/infercnv/R fixture_dir
##   python -m pytest tests/test_r_fixture.py                     # consumes fixture_dir/out (ICNV_R_FIXTURE_DIR)
##
## NOT EXECUTED in the build image (no R there): written against R >= 4.0 semantics; first run should be watched.
args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 2) stop("usage: Rscript make_hmm_fixture.R <reference R dir> <fixture dir>")
refdir <- args[1]; dir <- args[2]
flog.info <- flog.debug <- flog.warn <- flog.error <- function(...) invisible()
C_CHR <- "chr"

source_patched <- function(file, drop_pattern = NULL) {
    lines <- readLines(file)
    if (!is.null(drop_pattern)) lines <- lines[!grepl(drop_pattern, lines, fixed = TRUE)]
    eval(parse(text = lines), envir = globalenv())
}
source_patched(file.path(refdir, "inferCNV_HMM.R"), "HiddenMarkov:::makedensity")   # Viterbi.dthmm.adj, .get_HMM
source_patched(file.path(refdir, "noise_reduction.R"))                              # .median_filter

read_meta <- function(f) { kv <- strsplit(readLines(f), "=", fixed = TRUE); stats::setNames(lapply(kv, `[`, 2), sapply(kv, `[`, 1)) }
meta <- read_meta(file.path(dir, "meta.txt"))
G <- as.integer(meta$G); C <- as.integer(meta$C)
x <- matrix(readBin(file.path(dir, "x.bin"), "double", n = G * C, size = 8, endian = "little"), nrow = G, ncol = C)
chr_len <- scan(file.path(dir, "chr_len.txt"), quiet = TRUE)
chr_of <- rep(seq_along(chr_len), chr_len)
dir.create(file.path(dir, "out"), showWarnings = FALSE)

run_hmm <- function(tag) {
    Pi <- matrix(scan(file.path(dir, paste0(tag, "_Pi.txt")), quiet = TRUE), nrow = as.integer(meta[[paste0(tag, "_m")]]))
    delta <- scan(file.path(dir, paste0(tag, "_delta.txt")), quiet = TRUE)
    pm <- list(mean = scan(file.path(dir, paste0(tag, "_mean.txt")), quiet = TRUE),
               sd = scan(file.path(dir, paste0(tag, "_sd.txt")), quiet = TRUE))
    states <- matrix(-1, nrow = G, ncol = C)
    for (k in seq_along(chr_len)) {
        idx <- which(chr_of == k)
        for (cc in seq_len(C)) {
            obj <- list(x = x[idx, cc], Pi = Pi, delta = delta, distn = "norm", pm = pm)   # what HiddenMarkov::dthmm() returns
            states[idx, cc] <- Viterbi.dthmm.adj(obj)
        }
    }
    writeBin(as.double(states), file.path(dir, "out", paste0(tag, "_states.bin")), size = 8, endian = "little")
}
t0 <- Sys.time()
run_hmm("i6"); run_hmm("i3")
cat(sprintf("Viterbi.dthmm.adj: %d sequences in %.1f s\n", 2L * C * length(chr_len), as.numeric(Sys.time() - t0, units = "secs")))

## .median_filter per (index list x chromosome) block, as apply_median_filtering walks them (window_size 7)
lists <- lapply(strsplit(readLines(file.path(dir, "mf_lists.txt")), " ", fixed = TRUE), as.integer)
res <- x
for (l in lists) for (k in seq_along(chr_len)) {
    idx <- which(chr_of == k)
    res[idx, l] <- .median_filter(data = x[idx, l, drop = FALSE], window_size = 7, half_window = 3)
}
writeBin(as.double(res), file.path(dir, "out", "median_filter.bin"), size = 8, endian = "little")
cat("done\n")
