"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY - see infercnv_oracle.c header).

Two layers:

* ctypes bindings to ``liboracle.so`` (the C restatement, fast enough for 10^7 cell-genes);
* ``literal_*`` functions: a second, independent line-by-line transcription of the R loops in
  pure Python/NumPy, for small cases only, used to cross-check the C restatement
  (incl. the NA handling of ``.smooth_helper`` that the C twin does not carry).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
import this module.  Arrays are Fortran-ordered (R column-major) float64, shape (G genes, C cells);
all indices 0-based.
"""
from __future__ import annotations

import ctypes as ct
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_d_p = ct.POINTER(ct.c_double)
c_i_p = ct.POINTER(ct.c_int32)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only, no external deps)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "infercnv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib() -> ct.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        src = os.path.join(_HERE, "infercnv_oracle.c")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src) and os.access(_HERE, os.W_OK)):
            try:
                build()
            except (OSError, subprocess.CalledProcessError):
                if not os.path.exists(so):
                    raise
        _LIB = ct.CDLL(so)
        _LIB.orc_pnorm_upper_log.restype = ct.c_double
        _LIB.orc_pnorm_upper_log.argtypes = [ct.c_double]
    return _LIB


def _f(a) -> np.ndarray:
    return np.asfortranarray(a, dtype=np.float64)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_d_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(c_i_p)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def chr_ranges(chr_codes) -> tuple[np.ndarray, np.ndarray]:
    """Contiguous row ranges of each chromosome, in order of first appearance
    (rows are pre-sorted by chr in the reference, R/inferCNV.R:407-413)."""
    codes = np.asarray(chr_codes)
    starts, lens = [], []
    i = 0
    n = len(codes)
    while i < n:
        j = i
        while j < n and codes[j] == codes[i]:
            j += 1
        starts.append(i)
        lens.append(j - i)
        i = j
    return np.array(starts, dtype=np.int32), np.array(lens, dtype=np.int32)


def groups_to_csr(groups) -> tuple[np.ndarray, np.ndarray]:
    off = np.cumsum([0] + [len(g) for g in groups]).astype(np.int32)
    idx = np.concatenate([np.asarray(g, dtype=np.int32) for g in groups]) if len(groups) else np.zeros(0, np.int32)
    return off, _i32(idx)


def max_threads() -> int:
    """Threads OpenMP would use by default (honours OMP_NUM_THREADS)."""
    return int(lib().orc_max_threads())


def num_procs() -> int:
    """Processors of the machine, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)."""
    return int(lib().orc_num_procs())


def synth(G, chr_start, chr_len, cells_global, C_total, seed, nthreads=1) -> np.ndarray:
    """CPU twin of the device workload generator (infercnv_b200/csrc/icnv_synth.cu): (G, n) Fortran array for the given
    GLOBAL cell indices.  Lets bench.py's CPU arm draw its sample without loading the product library."""
    cells = np.ascontiguousarray(cells_global, dtype=np.int64)
    cs, cl = _i32(chr_start), _i32(chr_len)
    X = np.empty((int(G), len(cells)), dtype=np.float64, order="F")
    rc = lib().orc_synth(_dp(X), ct.c_int64(int(G)), cells.ctypes.data_as(ct.POINTER(ct.c_int64)), ct.c_int64(len(cells)),
                         ct.c_int64(int(C_total)), _ip(cs), _ip(cl), len(cs), ct.c_uint64(int(seed)), int(nthreads))
    if rc:
        raise MemoryError("orc_synth")
    return X


def pnorm_upper_log(z) -> np.ndarray:
    f = lib().orc_pnorm_upper_log
    return np.array([f(float(v)) for v in np.ravel(z)]).reshape(np.shape(z))


def normalize_by_seq_depth(X, normalize_factor=None) -> np.ndarray:
    X = _f(X)
    Y = np.empty_like(X, order="F")
    G, C = X.shape
    nf = -1.0 if normalize_factor is None else float(normalize_factor)
    rc = lib().orc_normalize_by_seq_depth(_dp(X), _dp(Y), ct.c_int64(G), ct.c_int64(C), ct.c_double(nf))
    assert rc == 0
    return Y


def log2xplus1(X) -> np.ndarray:
    X = _f(X)
    Y = np.empty_like(X, order="F")
    lib().orc_log2xplus1(_dp(X), _dp(Y), ct.c_int64(X.size))
    return Y


def invert_log2(X) -> np.ndarray:
    X = _f(X)
    Y = np.empty_like(X, order="F")
    lib().orc_invert_log2(_dp(X), _dp(Y), ct.c_int64(X.size))
    return Y


def apply_max_threshold_bounds(X, threshold) -> np.ndarray:
    X = _f(X)
    Y = np.empty_like(X, order="F")
    lib().orc_apply_max_threshold_bounds(_dp(X), _dp(Y), ct.c_int64(X.size), ct.c_double(threshold))
    return Y


def ref_means(X, groups, inv_log=False) -> np.ndarray:
    X = _f(X)
    G, C = X.shape
    off, idx = groups_to_csr(groups)
    M = np.empty((G, len(groups)), dtype=np.float64, order="F")
    rc = lib().orc_ref_means(_dp(X), ct.c_int64(G), ct.c_int64(C), _ip(off), _ip(idx), len(groups),
                             int(bool(inv_log)), _dp(M))
    assert rc == 0
    return M


def subtract_ref(X, means, use_bounds=True) -> np.ndarray:
    X = _f(X)
    M = _f(means)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    lib().orc_subtract_ref(_dp(X), _dp(Y), ct.c_int64(G), ct.c_int64(C), _dp(M), M.shape[1], int(bool(use_bounds)))
    return Y


def smooth_by_chromosome(X, chr_start, chr_len, window, literal=False, nthreads=1) -> np.ndarray:
    X = _f(X)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    cs, cl = _i32(chr_start), _i32(chr_len)
    rc = lib().orc_smooth_by_chromosome(_dp(X), _dp(Y), ct.c_int64(G), ct.c_int64(C), _ip(cs), _ip(cl), len(cs),
                                        int(window), int(bool(literal)), int(nthreads))
    if rc:
        raise ValueError(f"orc_smooth_by_chromosome rc={rc}")
    return Y


def center_columns(X, method="median", nthreads=1) -> np.ndarray:
    X = _f(X)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    lib().orc_center_columns(_dp(X), _dp(Y), ct.c_int64(G), ct.c_int64(C), int(method == "median"), int(nthreads))
    return Y


def median_filter(X, chr_start, chr_len, groups, window_size=7, nthreads=1) -> np.ndarray:
    X = _f(X)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    cs, cl = _i32(chr_start), _i32(chr_len)
    off, idx = groups_to_csr(groups)
    rc = lib().orc_median_filter(_dp(X), _dp(Y), ct.c_int64(G), ct.c_int64(C), _ip(cs), _ip(cl), len(cs), _ip(off),
                                 _ip(idx), len(groups), int(window_size), int(nthreads))
    if rc:
        raise ValueError(f"orc_median_filter rc={rc}")
    return Y


def hmm_params(m: int, t: float = 1e-6):
    """(Pi, delta) literal values of .get_HMM / .i3HMM_get_HMM (R/inferCNV_HMM.R:233-242,
    R/inferCNV_i3HMM.R:108-114): diagonal 1-5t for BOTH models."""
    Pi = np.full((m, m), t, dtype=np.float64, order="F")
    np.fill_diagonal(Pi, 1 - 5 * t)
    delta = np.full(m, t, dtype=np.float64)
    delta[2 if m == 6 else 1] = 1 - 5 * t
    return Pi, delta


def viterbi_seq(x, Pi, delta, mean, sd):
    x = np.ascontiguousarray(x, dtype=np.float64)
    Pi = _f(Pi)
    m = Pi.shape[0]
    st = np.empty(len(x), dtype=np.int32)
    mg = ct.c_double()
    d, mu, s = (np.ascontiguousarray(v, dtype=np.float64) for v in (delta, mean, sd))
    rc = lib().orc_viterbi_seq(_dp(x), ct.c_int64(len(x)), m, _dp(Pi), _dp(d), _dp(mu), _dp(s), _ip(st), ct.byref(mg))
    if rc:
        raise ValueError(f"orc_viterbi_seq rc={rc}")
    return st, mg.value


def viterbi_matrix(X, chr_start, chr_len, Pi, delta, mean, sd, groups=None, nthreads=1, want_margins=False):
    """states int32 (G, C) F-order, 1..m (or -1 for cells outside every group in group mode)."""
    X = _f(X)
    G, C = X.shape
    Pi = _f(Pi)
    m = Pi.shape[0]
    cs, cl = _i32(chr_start), _i32(chr_len)
    K = len(cs)
    d, mu = (np.ascontiguousarray(v, dtype=np.float64) for v in (delta, mean))
    s = np.ascontiguousarray(sd, dtype=np.float64)
    st = np.empty((G, C), dtype=np.int32, order="F")
    if groups is None:
        off, idx, ng = np.zeros(1, np.int32), np.zeros(1, np.int32), 0
        nseq = C
        assert s.size == m
    else:
        off, idx = groups_to_csr(groups)
        ng = len(groups)
        nseq = ng
        if s.size == m:
            s = np.tile(s, ng)
        assert s.size == m * ng
    mg = np.empty((K, nseq), dtype=np.float64, order="F")
    rc = lib().orc_viterbi_matrix(_dp(X), ct.c_int64(G), ct.c_int64(C), _ip(cs), _ip(cl), K, _ip(off), _ip(idx), ng, m,
                                  _dp(Pi), _dp(d), _dp(mu), _dp(s), _ip(st), _dp(mg), int(nthreads))
    if rc:
        raise ValueError(f"orc_viterbi_matrix rc={rc}")
    return (st, mg) if want_margins else st


def mean_sd_over_cells(X, idx):
    X = _f(X)
    idx = _i32(idx)
    mu, sg = ct.c_double(), ct.c_double()
    rc = lib().orc_mean_sd_over_cells(_dp(X), ct.c_int64(X.shape[0]), _ip(idx), ct.c_int64(len(idx)), ct.byref(mu),
                                      ct.byref(sg))
    assert rc == 0
    return mu.value, sg.value


def pairwise_dist(X, cells=None, nthreads=1) -> np.ndarray:
    """R's "dist" vector (strict lower triangle by columns) of the euclidean distances between the listed cells."""
    X = _f(X)
    G, C = X.shape
    idx = None if cells is None else _i32(cells)
    n = C if idx is None else len(idx)
    out = np.empty(n * (n - 1) // 2, dtype=np.float64)
    rc = lib().orc_pairwise_dist(_dp(X), ct.c_int64(G), ct.c_int64(C), _ip(idx) if idx is not None else None, ct.c_int64(n),
                                 _dp(out), int(nthreads))
    assert rc == 0
    return out


def clear_noise_via_ref_mean_sd(X, ref_idx, sd_amplifier=1.5) -> np.ndarray:
    X = _f(X)
    G, C = X.shape
    idx = _i32(ref_idx)
    Y = np.empty_like(X, order="F")
    rc = lib().orc_clear_noise_via_ref_mean_sd(_dp(X), _dp(Y), ct.c_int64(G), ct.c_int64(C), _ip(idx),
                                               ct.c_int64(len(idx)), ct.c_double(sd_amplifier))
    assert rc == 0
    return Y


def smooth_block(X, chr_start, chr_len, ref_groups, apply_log=True, threshold=3.0, window=101, use_bounds=True,
                 nthreads=1) -> np.ndarray:
    """run() steps 4, 8, 9, 10, 11, 12, 14 (R/inferCNV_ops.R:614-1031) on the depth-normalised matrix."""
    X = _f(X)
    G, C = X.shape
    Y = np.empty_like(X, order="F")
    cs, cl = _i32(chr_start), _i32(chr_len)
    off, idx = groups_to_csr(ref_groups)
    rc = lib().orc_smooth_block(_dp(X), _dp(Y), ct.c_int64(G), ct.c_int64(C), _ip(cs), _ip(cl), len(cs), _ip(off),
                                _ip(idx), len(ref_groups), int(bool(apply_log)), ct.c_double(threshold), int(window),
                                int(bool(use_bounds)), int(nthreads))
    if rc:
        raise ValueError(f"orc_smooth_block rc={rc}")
    return Y


def i3_hmm_params(X, cells, i3_p_val=0.05, t=1e-6):
    """R/inferCNV_i3HMM.R:17-80 + :99-156 with use_KS=FALSE (run() default, ops.R:274):
    mean = (mu-d, mu, mu+d), sd = sigma x3, d = |qnorm(p, 0, sigma)|."""
    from scipy.stats import norm

    mu, sigma = mean_sd_over_cells(X, cells)
    d = abs(norm.ppf(i3_p_val, loc=0.0, scale=sigma))
    Pi, delta = hmm_params(3, t)
    return Pi, delta, np.array([mu - d, mu, mu + d]), np.array([sigma] * 3)


# --------------------------------------------------------------------------------------------
# literal pure-Python transcriptions (small cases only)
# --------------------------------------------------------------------------------------------

def literal_smooth_helper(obs_data, window_length):
    """R/inferCNV_ops.R:2483-2532 + :2640-2661, 1-based arithmetic kept, NA = NaN."""
    orig = np.array(obs_data, dtype=np.float64)
    nas = np.isnan(orig)
    obs = orig[~nas]
    n = len(obs)
    end_data = obs.copy()
    tail_length = (window_length - 1) // 2
    if n >= window_length:
        denom = ((window_length - 1) / 2) ** 2 + window_length
        num = list(range(1, tail_length + 1)) + [tail_length + 1] + list(range(tail_length, 0, -1))
        filt = [v / denom for v in num]
        for i in range(n):
            if i + tail_length - (window_length - 1) < 0 or i + tail_length >= n:
                continue
            z = 0.0
            for j in range(window_length):
                z += filt[j] * obs[i + tail_length - j]
            end_data[i] = z
    counts = list(range(1, tail_length + 1)) + [tail_length + 1] + list(range(tail_length, 0, -1))
    iteration_range = tail_length if n > window_length else math.ceil(n / 2)
    for tail_end in range(1, iteration_range + 1):
        end_tail = n - tail_end + 1
        d_left = tail_end - 1
        d_right = min(n - tail_end, tail_length)
        r_left = tail_length - d_left
        r_right = tail_length - d_right
        denominator = (((window_length - 1) / 2) ** 2 + window_length) - (r_left * (r_left + 1)) / 2 - \
            (r_right * (r_right + 1)) / 2
        left = obs[0:tail_end + d_right]
        right = obs[end_tail - d_right - 1:n]
        rng = np.array(counts[tail_length + 1 - d_left - 1:tail_length + 1 + d_right], dtype=np.float64)
        end_data[tail_end - 1] = math.fsum(left * rng) / denominator
        end_data[end_tail - 1] = math.fsum(right * rng[::-1]) / denominator
    orig[~nas] = end_data
    return orig


def literal_viterbi(x, Pi, delta, mean, sd):
    """R/inferCNV_HMM.R:1101-1176 with scipy's log_ndtr in place of nmath pnorm (independent check)."""
    from scipy.special import log_ndtr

    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    if n < 2:
        return np.full(n, 3, dtype=np.int32)
    m = Pi.shape[0]
    sdm = float(np.median(sd))
    mean = np.asarray(mean, dtype=np.float64)
    nu = np.empty((n, m))
    logPi = np.log(Pi)

    def emis(v):
        e = log_ndtr(-np.abs(v - mean) / sdm)
        e = 1.0 / (-1.0 * e)
        e = e / e.sum()
        return np.log(e)

    nu[0] = np.log(delta) + emis(x[0])
    for i in range(1, n):
        nu[i] = (nu[i - 1][:, None] + logPi).max(axis=0) + emis(x[i])
    y = np.empty(n, dtype=np.int32)
    y[n - 1] = int(np.argmax(nu[n - 1]))
    for i in range(n - 2, -1, -1):
        y[i] = int(np.argmax(logPi[:, y[i + 1]] + nu[i]))
    return y + 1
