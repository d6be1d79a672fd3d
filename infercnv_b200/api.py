"""NumPy-facing wrappers of the host-pointer C ABI (the same entry points the R shim binds).

Matrices are (G genes, C cells) float64, Fortran order (R column-major: a cell's genes are
contiguous).  Indices are 0-based.  Every call goes H2D -> sm_100a kernels -> D2H inside the
library; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as ct

import numpy as np

from . import _lib


def _f64(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("expected a genes x cells matrix")
    return np.asfortranarray(a)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data


def chr_ranges(chr_codes):
    """Contiguous row range of every chromosome in order of appearance (rows are pre-sorted by
    chr in the reference, R/inferCNV.R:407-413).  A chromosome that re-appears is an error."""
    codes = np.asarray(chr_codes)
    change = np.flatnonzero(codes[1:] != codes[:-1]) + 1
    starts = np.concatenate([[0], change]).astype(np.int32)
    lens = np.diff(np.concatenate([starts, [len(codes)]])).astype(np.int32)
    if len(set(codes[starts].tolist())) != len(starts):
        raise ValueError("gene_order$chr is not sorted: a chromosome appears in two separate runs")
    return starts, lens


def groups_to_csr(groups):
    off = np.cumsum([0] + [len(g) for g in groups]).astype(np.int32)
    idx = np.concatenate([np.asarray(g, dtype=np.int32) for g in groups]) if len(groups) else np.zeros(1, np.int32)
    return off, _i32(idx)


def init(device: int = 0) -> None:
    _lib.check(_lib.load().icnv_init(int(device)))


def init_devices(device_ids=None) -> int:
    """Single-process multi-GPU (what the R shim does): the host-pointer calls that stream cells shard them over these
    devices.  None = every device present.  Returns the number of devices in use."""
    lib = _lib.load()
    if device_ids is None:
        _lib.check(lib.icnv_init_devices(0, None))
    else:
        ids = np.ascontiguousarray(device_ids, dtype=np.int32)
        _lib.check(lib.icnv_init_devices(len(ids), ids.ctypes.data))
    return int(lib.icnv_devices_in_use())


def shutdown() -> None:
    _lib.load().icnv_shutdown()


def reinit(device: int = 0) -> None:
    """Free everything and initialise again - the ICNV_* tuning switches are read from the environment at icnv_init only."""
    shutdown()
    init(device)


def set_host_threads(n: int) -> None:
    _lib.check(_lib.load().icnv_set_host_threads(int(n)))


def device_count() -> int:
    return int(_lib.load().icnv_device_count())


def launch_count() -> int:
    return int(_lib.load().icnv_launch_count())


def set_hmm_mode(mode) -> None:
    """0 / "exact": reference-order arithmetic; 1 / "fast" / "fast64": certified FP64 pass (default); 2 / "fast32": a
    single-precision pass certified by per-path margins first, then the FP64 pass for what it cannot certify (measured slower
    on the benchmark data, kept as an option).  What no pass certifies is recomputed in reference-order arithmetic."""
    m = {"exact": 0, "fast": 1, "fast64": 1, "fast32": 2}.get(mode, mode)
    _lib.check(_lib.load().icnv_set_hmm_mode(int(m)))


def hmm_rerun_count() -> int:
    return int(_lib.load().icnv_hmm_rerun_count())


def hmm_second_pass_count() -> int:
    """sequences the single-precision pass of the last Viterbi call handed to the FP64 pass"""
    return int(_lib.load().icnv_hmm_second_pass_count())


def ref_means(X, groups, inv_log=False) -> np.ndarray:
    X = _f64(X)
    G, C = X.shape
    off, idx = groups_to_csr(groups)
    M = np.empty((G, len(groups)), dtype=np.float64, order="F")
    _lib.check(_lib.load().icnv_ref_means_f64(_p(X), G, C, _p(off), _p(idx), len(groups), int(bool(inv_log)), _p(M)))
    return M


def subtract_ref(X, means, use_bounds=True) -> np.ndarray:
    X = _f64(X)
    M = _f64(means)
    G, C = X.shape
    if M.shape[0] != G:
        raise ValueError("means must have one row per gene")
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_subtract_ref_f64(_p(X), _p(Y), G, C, _p(M), M.shape[1], int(bool(use_bounds))))
    return Y


def smooth(X, chr_start, chr_len, window_length) -> np.ndarray:
    X = _f64(X)
    G, C = X.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_smooth_f64(_p(X), _p(Y), G, C, _p(cs), _p(cl), len(cs), int(window_length)))
    return Y


def center(X, method="median") -> np.ndarray:
    X = _f64(X)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_center_f64(_p(X), _p(Y), G, C, int(method == "median")))
    return Y


def normalize_counts_by_seq_depth(X, normalize_factor=None) -> np.ndarray:
    X = _f64(X)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    nf = -1.0 if normalize_factor is None else float(normalize_factor)
    _lib.check(_lib.load().icnv_normalize_counts_by_seq_depth_f64(_p(X), _p(Y), G, C, nf))
    return Y


def clear_noise_via_ref_mean_sd(X, cells, sd_amplifier=1.5) -> np.ndarray:
    X = _f64(X)
    G, C = X.shape
    idx = _i32(cells)
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_clear_noise_via_ref_mean_sd_f64(_p(X), _p(Y), G, C, _p(idx), len(idx), float(sd_amplifier)))
    return Y


def log2xplus1(X) -> np.ndarray:
    X = _f64(X)
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_log2xplus1_f64(_p(X), _p(Y), X.size))
    return Y


def invert_log2(X) -> np.ndarray:
    X = _f64(X)
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_invert_log2_f64(_p(X), _p(Y), X.size))
    return Y


def apply_max_threshold_bounds(X, threshold) -> np.ndarray:
    X = _f64(X)
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_apply_max_threshold_bounds_f64(_p(X), _p(Y), X.size, float(threshold)))
    return Y


def smooth_block(X, chr_start, chr_len, ref_groups, apply_log=True, threshold=3.0, window_length=101,
                 use_bounds=True, out=None) -> np.ndarray:
    X = _f64(X)
    G, C = X.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    off, idx = groups_to_csr(ref_groups)
    Y = np.empty_like(X, order="F") if out is None else out
    _lib.check(_lib.load().icnv_smooth_block_f64(_p(X), _p(Y), G, C, _p(cs), _p(cl), len(cs), _p(off), _p(idx),
                                                 len(ref_groups), int(bool(apply_log)), float(threshold),
                                                 int(window_length), int(bool(use_bounds))))
    return Y


def smooth_hmm(X, chr_start, chr_len, ref_groups, Pi, delta, mean, sd, apply_log=True, threshold=3.0, window_length=101,
               use_bounds=True, out=None, out_states=None):
    """Fused smooth block + per-cell HMM (one upload of the matrix).  Returns (Y, states).  States come back as int32
    by default; pass a uint8 `out_states` for the one-byte wire format (icnv_smooth_hmm_u8_f64)."""
    X = _f64(X)
    G, C = X.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    off, idx = groups_to_csr(ref_groups)
    Pi = np.asfortranarray(Pi, dtype=np.float64)
    delta, mean, sd = (np.ascontiguousarray(v, dtype=np.float64) for v in (delta, mean, sd))
    Y = np.empty_like(X, order="F") if out is None else out
    S = np.empty((G, C), dtype=np.int32, order="F") if out_states is None else out_states
    if S.dtype not in (np.int32, np.uint8) or S.shape != (G, C) or not S.flags.f_contiguous:
        raise ValueError("out_states must be a Fortran-ordered (G, C) int32 or uint8 array")
    lib = _lib.load()
    fn = lib.icnv_smooth_hmm_u8_f64 if S.dtype == np.uint8 else lib.icnv_smooth_hmm_f64
    _lib.check(fn(_p(X), _p(Y), _p(S), G, C, _p(cs), _p(cl), len(cs), _p(off), _p(idx), len(ref_groups), int(bool(apply_log)),
                  float(threshold), int(window_length), int(bool(use_bounds)), Pi.shape[0], _p(Pi), _p(delta), _p(mean),
                  _p(sd)))
    return Y, S


def viterbi(X, chr_start, chr_len, Pi, delta, mean, sd, groups=None, want_margins=False, out=None):
    """States (G, C): int32 with -1 = unassigned by default; a uint8 `out` selects the one-byte wire format
    (icnv_viterbi_u8_f64, 255 = unassigned)."""
    X = _f64(X)
    G, C = X.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    K = len(cs)
    Pi = np.asfortranarray(Pi, dtype=np.float64)
    m = Pi.shape[0]
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    sd = np.ascontiguousarray(sd, dtype=np.float64)
    if groups is None:
        off = idx = None
        ng, nseq = 0, C
        if sd.size != m:
            raise ValueError("sd must have m entries in per-cell mode")
    else:
        off, idx = groups_to_csr(groups)
        ng = nseq = len(groups)
        if sd.size == m:
            sd = np.tile(sd, ng)
        if sd.size != m * ng:
            raise ValueError("sd must have m entries per group")
    states = np.empty((G, C), dtype=np.int32, order="F") if out is None else out
    if states.dtype not in (np.int32, np.uint8) or states.shape != (G, C) or not states.flags.f_contiguous:
        raise ValueError("out must be a Fortran-ordered (G, C) int32 or uint8 array")
    margins = np.empty((K, nseq), dtype=np.float64, order="F") if want_margins else None
    lib = _lib.load()
    fn = lib.icnv_viterbi_u8_f64 if states.dtype == np.uint8 else lib.icnv_viterbi_f64
    _lib.check(fn(_p(X), G, C, _p(cs), _p(cl), K, _p(off), _p(idx), ng, m, _p(Pi), _p(delta), _p(mean), _p(sd), _p(states),
                  _p(margins)))
    return (states, margins) if want_margins else states


def median_filter(X, chr_start, chr_len, groups, window_size=7, out=None) -> np.ndarray:
    X = _f64(X)
    G, C = X.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    off, idx = groups_to_csr(groups)
    Y = np.empty_like(X, order="F") if out is None else out
    _lib.check(_lib.load().icnv_median_filter_f64(_p(X), _p(Y), G, C, _p(cs), _p(cl), len(cs), _p(off), _p(idx),
                                                  len(groups), int(window_size)))
    return Y


def mean_sd(X, cells):
    X = _f64(X)
    G, C = X.shape
    idx = _i32(cells)
    mu, sg = ct.c_double(), ct.c_double()
    _lib.check(_lib.load().icnv_mean_sd_f64(_p(X), G, C, _p(idx), len(idx), ct.addressof(mu), ct.addressof(sg)))
    return mu.value, sg.value


def pairwise_dist(X, cells=None, out=None):
    """parallelDist(t(X[, cells])) (euclidean): R's "dist" vector - n (n - 1) / 2 doubles, the strict lower triangle by
    columns - for the listed cells (default: all), R/inferCNV_tumor_subclusters.R:191 and the other hclust() call sites."""
    X = _f64(X)
    G, C = X.shape
    idx = None if cells is None else _i32(cells)
    n = C if idx is None else len(idx)
    n_out = n * (n - 1) // 2
    if out is None:
        out = np.empty(n_out, dtype=np.float64)
    if out.dtype != np.float64 or out.size != n_out or not out.flags.c_contiguous:
        raise ValueError("out must be a contiguous float64 vector of n (n - 1) / 2 entries")
    _lib.check(_lib.load().icnv_pairwise_dist_f64(_p(X), G, C, _p(idx) if idx is not None else None, n, _p(out)))
    return out


def pairwise_dist_rows(x, out=None):
    """parallelDist(x) for x = observations x variables (the t(expr.data[, cells]) the reference passes), without
    transposing it back: icnv_pairwise_dist_rows_f64."""
    x = np.asfortranarray(np.asarray(x, dtype=np.float64))
    if x.ndim != 2:
        raise ValueError("x must be a matrix (observations x variables)")
    n, G = x.shape
    n_out = n * (n - 1) // 2
    if out is None:
        out = np.empty(n_out, dtype=np.float64)
    if out.dtype != np.float64 or out.size != n_out or not out.flags.c_contiguous:
        raise ValueError("out must be a contiguous float64 vector of n (n - 1) / 2 entries")
    _lib.check(_lib.load().icnv_pairwise_dist_rows_f64(_p(x), n, G, _p(out)))
    return out


# ---- CNV region calling on the state matrix (R/inferCNV_HMM.R:706-1087) ---------------------------------------

def _u8(a) -> np.ndarray:
    """State matrix in the one-byte wire format (0..6, 255 = unassigned).  int / float matrices as R holds them
    (-1 = unassigned) are narrowed here; that is marshalling, like the R shim's double -> byte loop."""
    a = np.asarray(a)
    if a.ndim == 1:
        a = a[:, None]
    if a.dtype != np.uint8:
        if np.any(a != np.floor(a)) or a.min() < -1 or a.max() > 254:
            raise ValueError("states must be integers in -1..254")
        a = np.where(a < 0, 255, a).astype(np.uint8)
    return np.asfortranarray(a)


REGION_FIELDS = (("seq", np.int32), ("chr", np.int32), ("first_gene", np.int32), ("last_gene", np.int32),
                 ("state", np.int32), ("start", np.float64), ("end", np.float64))


def _fetch_regions(n: int) -> dict:
    out = {k: np.empty(n, dtype=dt) for k, dt in REGION_FIELDS}
    _lib.check(_lib.load().icnv_cnv_regions_fetch(n, *[_p(out[k]) for k, _ in REGION_FIELDS]))
    return out


def state_consensus(states, groups) -> np.ndarray:
    """Modal state per gene per group, (G, n_grp) uint8 (icnv_state_consensus_u8)."""
    S = _u8(states)
    G, C = S.shape
    off, idx = groups_to_csr(groups)
    cons = np.empty((G, len(groups)), dtype=np.uint8, order="F")
    _lib.check(_lib.load().icnv_state_consensus_u8(_p(S), G, C, _p(off), _p(idx), len(groups), _p(cons)))
    return cons


def cnv_regions(seqs, chr_start, chr_len, gene_start, gene_stop) -> dict:
    """Run-length regions of every column of `seqs` (icnv_cnv_regions_u8 + icnv_cnv_regions_fetch): dict of arrays
    seq, chr, first_gene, last_gene, state, start, end in (sequence, chromosome, position) order."""
    S = _u8(seqs)
    G, n_seq = S.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    gs, ge = (np.ascontiguousarray(v, dtype=np.float64) for v in (gene_start, gene_stop))
    n = ct.c_int64(0)
    _lib.check(_lib.load().icnv_cnv_regions_u8(_p(S), G, n_seq, _p(cs), _p(cl), len(cs), _p(gs), _p(ge), ct.addressof(n)))
    return _fetch_regions(int(n.value))


def predicted_cnv_regions(states, chr_start, chr_len, gene_start, gene_stop, groups, want_consensus=False):
    """Consensus of every cell group + region calling in one upload (icnv_predicted_cnv_regions_u8).
    Returns the region dict, or (regions, consensus) with want_consensus."""
    S = _u8(states)
    G, C = S.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    gs, ge = (np.ascontiguousarray(v, dtype=np.float64) for v in (gene_start, gene_stop))
    off, idx = groups_to_csr(groups)
    cons = np.empty((G, len(groups)), dtype=np.uint8, order="F") if want_consensus else None
    n = ct.c_int64(0)
    _lib.check(_lib.load().icnv_predicted_cnv_regions_u8(_p(S), G, C, _p(cs), _p(cl), len(cs), _p(gs), _p(ge), _p(off), _p(idx),
                                                         len(groups), _p(cons), ct.addressof(n)))
    reg = _fetch_regions(int(n.value))
    return (reg, cons) if want_consensus else reg


# ---- gene filters and counts ingest (run() steps 2-3) -----------------------------------------------------------

def gene_stats(X):
    """Per-gene (sums, n_pos, means) of a dense matrix (icnv_gene_stats_f64): what .below_min_mean_expr_cutoff and
    require_above_min_cells_ref compare with their thresholds."""
    X = _f64(X)
    G, C = X.shape
    sums, means = np.empty(G), np.empty(G)
    n_pos = np.empty(G, dtype=np.int32)
    _lib.check(_lib.load().icnv_gene_stats_f64(_p(X), G, C, _p(sums), _p(n_pos), _p(means)))
    return sums, n_pos, means


def remove_genes(X, keep) -> np.ndarray:
    """X[keep, ] for increasing row indices `keep` (icnv_remove_genes_f64)."""
    X = _f64(X)
    G, C = X.shape
    keep = _i32(keep)
    Y = np.empty((len(keep), C), dtype=np.float64, order="F")
    _lib.check(_lib.load().icnv_remove_genes_f64(_p(X), G, C, _p(keep), len(keep), _p(Y)))
    return Y


def _csc(p, i, x):
    p, i = _i32(p), _i32(i)
    x = np.ascontiguousarray(x, dtype=np.float64)
    if len(i) != len(x) or len(p) < 2 or p[-1] != len(x):
        raise ValueError("inconsistent compressed-column arrays")
    return p, i, x


def csc_gene_stats(p, i, x, n_genes):
    """gene_stats for a compressed-sparse-column counts matrix (dgCMatrix @p / @i / @x)."""
    p, i, x = _csc(p, i, x)
    G, C = int(n_genes), len(p) - 1
    sums, means = np.empty(G), np.empty(G)
    n_pos = np.empty(G, dtype=np.int32)
    _lib.check(_lib.load().icnv_csc_gene_stats_f64(_p(p), _p(i), _p(x), G, C, _p(sums), _p(n_pos), _p(means)))
    return sums, n_pos, means


def csc_normalize(p, i, x, n_genes, keep=None, normalize_factor=None, want_col_sums=False):
    """Dense depth-normalised matrix (kept genes x cells) straight from compressed columns (icnv_csc_normalize_f64)."""
    p, i, x = _csc(p, i, x)
    G, C = int(n_genes), len(p) - 1
    keep = None if keep is None else _i32(keep)
    G_out = G if keep is None else len(keep)
    Y = np.empty((G_out, C), dtype=np.float64, order="F")
    cs = np.empty(C) if want_col_sums else None
    nf = -1.0 if normalize_factor is None else float(normalize_factor)
    _lib.check(_lib.load().icnv_csc_normalize_f64(_p(p), _p(i), _p(x), G, C, _p(keep), 0 if keep is None else len(keep), nf,
                                                  _p(Y), _p(cs)))
    return (Y, cs) if want_col_sums else Y


# ---- outlier clamp and noise clearing (run() steps 16, 22) --------------------------------------------------------

def remove_outliers_norm(X, lower_bound=None, upper_bound=None, want_bounds=False):
    """.remove_outliers_norm (icnv_remove_outliers_norm_f64); a missing bound selects out_method "average_bound"."""
    X = _f64(X)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    b = np.empty(2)
    lo = float("nan") if lower_bound is None else float(lower_bound)
    hi = float("nan") if upper_bound is None else float(upper_bound)
    _lib.check(_lib.load().icnv_remove_outliers_norm_f64(_p(X), _p(Y), G, C, lo, hi, _p(b)))
    return (Y, (float(b[0]), float(b[1]))) if want_bounds else Y


def clear_noise(X, cells, threshold, noise_logistic=False) -> np.ndarray:
    """clear_noise (icnv_clear_noise_f64); cells = reference cells, or None / empty for "all data"."""
    X = _f64(X)
    G, C = X.shape
    idx = None if cells is None or len(cells) == 0 else _i32(cells)
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_clear_noise_f64(_p(X), _p(Y), G, C, _p(idx), 0 if idx is None else len(idx), float(threshold),
                                                int(bool(noise_logistic))))
    return Y


def clear_noise_via_ref_mean_sd_logistic(X, cells, sd_amplifier=1.5) -> np.ndarray:
    X = _f64(X)
    G, C = X.shape
    idx = _i32(cells)
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_clear_noise_via_ref_mean_sd_logistic_f64(_p(X), _p(Y), G, C, _p(idx), len(idx),
                                                                        float(sd_amplifier)))
    return Y


def assign_hmm_states_to_proxy_expr_vals(states, m=6) -> np.ndarray:
    """State matrix -> proxy expression values (icnv_assign_hmm_states_to_proxy_expr_vals_f64)."""
    X = _f64(states)
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_assign_hmm_states_to_proxy_expr_vals_f64(_p(X), _p(Y), X.size, int(m)))
    return Y


# ---- per-chromosome subcluster HMM (R/inferCNV_HMM.R:412-487) -----------------------------------------------------

def viterbi_per_chr(X, chr_start, chr_len, groups_per_chr, Pi, delta, mean, sds) -> np.ndarray:
    """One Viterbi sequence per (chromosome, subcluster of that chromosome) on the subcluster's rowMeans
    (icnv_viterbi_per_chr_u8_f64).  groups_per_chr: per chromosome a list of cell-index arrays; sds: m values per group,
    chromosome-major in the same order.  Returns uint8 states (G, C), 255 where a cell has no subcluster."""
    X = _f64(X)
    G, C = X.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    if len(groups_per_chr) != len(cs):
        raise ValueError("one list of subclusters per chromosome is needed")
    flat = [np.asarray(g) for per in groups_per_chr for g in per]
    chr_off = np.cumsum([0] + [len(per) for per in groups_per_chr]).astype(np.int32)
    off, idx = groups_to_csr(flat)
    Pi = np.asfortranarray(Pi, dtype=np.float64)
    m = Pi.shape[0]
    delta, mean = (np.ascontiguousarray(v, dtype=np.float64) for v in (delta, mean))
    sds = np.ascontiguousarray(sds, dtype=np.float64)
    if sds.size == m:
        sds = np.tile(sds, len(flat))
    if sds.size != m * len(flat):
        raise ValueError("sds must have m entries per (chromosome, subcluster)")
    S = np.empty((G, C), dtype=np.uint8, order="F")
    _lib.check(_lib.load().icnv_viterbi_per_chr_u8_f64(_p(X), G, C, _p(cs), _p(cl), len(cs), _p(chr_off), _p(off), _p(idx), m,
                                                       _p(Pi), _p(delta), _p(mean), _p(sds), _p(S)))
    return S


def apply_state_consensus(states, chr_start, chr_len, groups) -> np.ndarray:
    """Every cell of a group takes the group's consensus state on the genes region calling covers
    (icnv_apply_state_consensus_u8)."""
    S = _u8(states)
    G, C = S.shape
    cs, cl = _i32(chr_start), _i32(chr_len)
    off, idx = groups_to_csr(groups)
    out = np.empty((G, C), dtype=np.uint8, order="F")
    _lib.check(_lib.load().icnv_apply_state_consensus_u8(_p(S), G, C, _p(cs), _p(cl), len(cs), _p(off), _p(idx), len(groups),
                                                         _p(out)))
    return out


def scale_infercnv_expr(X) -> np.ndarray:
    """Per-gene z-scores across the cells (icnv_scale_infercnv_expr_f64)."""
    X = _f64(X)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    _lib.check(_lib.load().icnv_scale_infercnv_expr_f64(_p(X), _p(Y), G, C))
    return Y


def gather_genes(X, idx) -> np.ndarray:
    """X[idx, ] for any row order (icnv_gather_genes_f64)."""
    X = _f64(X)
    G, C = X.shape
    idx = _i32(idx)
    Y = np.empty((len(idx), C), dtype=np.float64, order="F")
    _lib.check(_lib.load().icnv_gather_genes_f64(_p(X), G, C, _p(idx), len(idx), _p(Y)))
    return Y
