"""torch-tensor wrappers of the device-pointer C ABI (icnv_dev_*), used by the benchmark and the
multi-GPU driver.  torch is plumbing only (device memory, streams, torch.distributed); every
kernel is in libinfercnv_b200.so.

A device matrix is a torch.float64 tensor of shape (C, G), C-contiguous: that is exactly R's
column-major G x C (a cell's genes contiguous).
"""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np
import torch

from . import _lib, dist as shard
from .api import _i32, groups_to_csr


def _stream_ptr():
    """torch's current stream as the `void *stream` of the device ABI.  torch's default stream is
    the legacy NULL stream (handle 0), which the ABI reserves for "the library's own stream": pass
    cudaStreamLegacy (handle 0x1) for it instead, so our launches stay ordered with torch's ops."""
    h = torch.cuda.current_stream().cuda_stream
    return ct.c_void_p(h if h else 1)


class Engine:
    """One per process / GPU."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        _lib.check(self.lib.icnv_init(self.device))
        self.tdev = torch.device("cuda", self.device)
        self.timing = None   # when set to a list, smooth_block appends (name, start_event, end_event)
        self.collective = True   # False: never enter a collective even if a process group exists
        # pass 2 re-uses pass 1's reference columns instead of recomputing them (ICNV_REF_REUSE=0 turns it off)
        self.reuse_reference_pass = os.environ.get("ICNV_REF_REUSE", "1") != "0"
        self._plan_cache = {}
        self._T = None

    # ---- process-group plumbing -------------------------------------------------------------------------------------
    def _world(self) -> int:
        import torch.distributed as tdist
        return tdist.get_world_size() if (self.collective and tdist.is_available() and tdist.is_initialized()) else 1

    def _all_gather(self, out: torch.Tensor, part: torch.Tensor) -> None:
        """out[(world,) + part.shape] <- every rank's `part`.  NCCL: one ncclAllGather on torch's current stream over
        NVLink.  Any other backend (gloo: several ranks sharing one GPU in the tests, or CPU-only plumbing tests) goes
        through host memory - same values, same order."""
        import torch.distributed as tdist
        if tdist.get_backend() == "nccl":
            tdist.all_gather_into_tensor(out, part)
            return
        host = part.detach().cpu().contiguous()
        parts = [torch.empty_like(host) for _ in range(tdist.get_world_size())]
        tdist.all_gather(parts, host)
        out.copy_(torch.stack(parts).to(out.device))

    def _all_gather_list(self, t: torch.Tensor):
        """every rank's `t` (same shape on all ranks) as a list in rank order"""
        import torch.distributed as tdist
        world = tdist.get_world_size()
        if tdist.get_backend() == "nccl":
            parts = [torch.empty_like(t) for _ in range(world)]
            tdist.all_gather(parts, t)
            return parts
        host = t.detach().cpu().contiguous()
        parts = [torch.empty_like(host) for _ in range(world)]
        tdist.all_gather(parts, host)
        return [x.to(t.device) for x in parts]

    def _all_reduce_sum(self, t: torch.Tensor) -> None:
        import torch.distributed as tdist
        if tdist.get_backend() == "nccl":
            tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
            return
        host = t.detach().cpu().contiguous()
        tdist.all_reduce(host, op=tdist.ReduceOp.SUM)
        t.copy_(host.to(t.device))

    # ---- data ---------------------------------------------------------------------------------------
    def synth(self, G, chr_start, chr_len, cells_global, C_total, seed) -> torch.Tensor:
        """Synthetic depth-normalised expression for the given GLOBAL cell indices (runs of
        consecutive indices are generated with one launch each)."""
        cells_global = np.asarray(cells_global, dtype=np.int64)
        cs, cl = _i32(chr_start), _i32(chr_len)
        X = torch.empty((len(cells_global), G), dtype=torch.float64, device=self.tdev)
        if len(cells_global) == 0:
            return X
        breaks = np.flatnonzero(np.diff(cells_global) != 1) + 1
        starts = np.concatenate([[0], breaks])
        ends = np.concatenate([breaks, [len(cells_global)]])
        for s, e in zip(starts, ends):
            ptr = X.data_ptr() + int(s) * G * 8
            _lib.check(self.lib.icnv_dev_synth_f64(ptr, G, int(cells_global[s]), int(e - s), int(C_total), cs.ctypes.data,
                                                   cl.ctypes.data, len(cs), ct.c_uint64(seed), _stream_ptr()))
        return X

    # ---- building blocks ----------------------------------------------------------------------------
    def group_partial_sums(self, X: torch.Tensor, cells: torch.Tensor, apply_log: bool) -> torch.Tensor:
        """(n_chunks, G) partial sums over chunks of CHUNK list entries (fixed order)."""
        n = int(cells.numel())
        G = X.shape[1]
        n_chunks = (n + shard.CHUNK - 1) // shard.CHUNK
        out = torch.empty((n_chunks, G), dtype=torch.float64, device=self.tdev)
        if n:
            _lib.check(self.lib.icnv_dev_group_partial_sums_f64(X.data_ptr(), G, X.stride(0), cells.data_ptr(), n,
                                                                shard.CHUNK, int(bool(apply_log)), out.data_ptr(),
                                                                _stream_ptr()))
        return out

    def combine_partials(self, partial: torch.Tensor, count: int) -> torch.Tensor:
        G = partial.shape[1]
        out = torch.empty(G, dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.icnv_dev_combine_partials_f64(partial.data_ptr(), G, partial.shape[0], int(count),
                                                          out.data_ptr(), _stream_ptr()))
        return out

    def bounds(self, means: torch.Tensor):
        """means: (n_grp, G) -> lo, hi, mid (each G)."""
        n_grp, G = means.shape
        lo, hi, mid = (torch.empty(G, dtype=torch.float64, device=self.tdev) for _ in range(3))
        _lib.check(self.lib.icnv_dev_bounds_from_means_f64(means.data_ptr(), G, n_grp, lo.data_ptr(), hi.data_ptr(),
                                                           mid.data_ptr(), _stream_ptr()))
        return lo, hi, mid

    def cell_pipeline(self, X, cols, Y, chr_start, chr_len, apply_log, b1, threshold, window, center, b2, apply_exp2,
                      use_bounds=True, err_flag=None):
        G = X.shape[1]
        cs, cl = _i32(chr_start), _i32(chr_len)
        n_cols = int(cols.numel()) if cols is not None else X.shape[0]

        def sel(b):
            if b is None:
                return None, None, None
            lo, hi, mid = b
            return (lo.data_ptr(), hi.data_ptr(), None) if use_bounds else (None, None, mid.data_ptr())

        lo1, hi1, mid1 = sel(b1)
        lo2, hi2, mid2 = sel(b2)
        _lib.check(self.lib.icnv_dev_cell_pipeline_f64(
            X.data_ptr(), G, X.stride(0), cols.data_ptr() if cols is not None else None, n_cols, Y.data_ptr(),
            Y.stride(0), cs.ctypes.data, cl.ctypes.data, len(cs), int(bool(apply_log)), lo1, hi1, mid1, float(threshold),
            int(window), int(center), lo2, hi2, mid2, int(bool(apply_exp2)),
            err_flag.data_ptr() if err_flag is not None else None, _stream_ptr()))

    # ---- the hot path ---------------------------------------------------------------------------------------
    def _reference_bounds(self, X, chr_start, chr_len, ref_groups_local, ref_sizes, max_chunks, apply_log, threshold, window,
                          use_bounds, flag):
        """The two reference-mean steps of the smooth block (ops.R:1678, run() steps 8 and 12): bounds b1 of the log
        values, pass 1 of the cell pipeline over this rank's reference cells (-> T, centred and smoothed), bounds b2 of T.
        Only the reference columns of X are read.  Returns (b1, b2, T, n_ref, ref_leading)."""
        C, G = X.shape
        ref_groups_local = [np.asarray(g, dtype=np.int32) for g in ref_groups_local]
        n_grp = len(ref_groups_local)
        if ref_sizes is None:
            ref_sizes = [len(g) for g in ref_groups_local]
        if max_chunks is None:
            max_chunks = [(len(g) + shard.CHUNK - 1) // shard.CHUNK for g in ref_groups_local]
        key = (tuple(g.tobytes() for g in ref_groups_local), C)
        cached = self._plan_cache.get(key)
        if cached is None:       # device copies of the index lists are built once per partition, not per call
            d_groups = [torch.from_numpy(g).to(self.tdev) for g in ref_groups_local]
            t_lists, pos = [], 0
            for g in ref_groups_local:
                t_lists.append(torch.arange(pos, pos + len(g), dtype=torch.int32, device=self.tdev))
                pos += len(g)
            all_ref = torch.cat(d_groups) if pos else None
            self._plan_cache.clear()
            cached = self._plan_cache[key] = (d_groups, t_lists, all_ref)
        d_groups, t_lists, all_ref = cached

        world = self._world()
        tot = int(sum(max_chunks)) if world > 1 else int(sum((len(g) + shard.CHUNK - 1) // shard.CHUNK for g in ref_groups_local))
        rows = [0]
        for k in range(n_grp):
            rows.append(rows[-1] + (int(max_chunks[k]) if world > 1 else (len(ref_groups_local[k]) + shard.CHUNK - 1) // shard.CHUNK))
        row_off = np.asarray(rows, dtype=np.int32)
        counts = np.asarray([int(v) for v in ref_sizes], dtype=np.int64)
        # persistent exchange buffers: every group's chunk sums are written straight into this rank's block of `packed`
        # (rows past a group's local chunk count are never written and stay zero), ONE all-gather per reference-mean step
        # carries them, and one kernel turns the gathered rows into the bounds - no per-step torch ops in between
        bkey = (G, tot, world, tuple(int(v) for v in row_off), tuple(len(g) for g in ref_groups_local))   # stale rows would poison the padding
        if getattr(self, "_xbuf_key", None) != bkey:
            self._xbuf_key = bkey
            self._packed = torch.zeros((max(tot, 1), G), dtype=torch.float64, device=self.tdev)
            self._gathered = torch.empty((world, max(tot, 1), G), dtype=torch.float64, device=self.tdev) if world > 1 else None
            self._bounds = [torch.empty((3, G), dtype=torch.float64, device=self.tdev) for _ in range(2)]

        def group_bounds(src, lists, log, slot):
            for k in range(n_grp):
                n = int(lists[k].numel())
                if n:
                    _lib.check(self.lib.icnv_dev_group_partial_sums_f64(
                        src.data_ptr(), G, src.stride(0), lists[k].data_ptr(), n, shard.CHUNK, int(bool(log)),
                        self._packed.data_ptr() + 8 * G * int(row_off[k]), _stream_ptr()))
            part = self._packed
            if world > 1:
                self._all_gather(self._gathered, self._packed)
                part = self._gathered
            b = self._bounds[slot]
            _lib.check(self.lib.icnv_dev_bounds_from_partials_f64(part.data_ptr(), G, world, max(tot, 1), n_grp,
                                                                  row_off.ctypes.data, counts.ctypes.data, b[0].data_ptr(),
                                                                  b[1].data_ptr(), b[2].data_ptr(), _stream_ptr()))
            return b[0], b[1], b[2]

        b1 = group_bounds(X, d_groups, apply_log, 0)
        # pass 1: reference cells only, up to the median centring
        n_ref = int(sum(len(g) for g in ref_groups_local))
        ref_leading = n_ref > 0 and np.array_equal(np.concatenate(ref_groups_local), np.arange(n_ref))
        if self._T is None or self._T.shape != (max(n_ref, 1), G):
            self._T = torch.empty((max(n_ref, 1), G), dtype=torch.float64, device=self.tdev)
        T = self._T
        if n_ref:
            self.cell_pipeline(X, all_ref, T, chr_start, chr_len, apply_log, b1, threshold, window, 1, None, False,
                               use_bounds, flag)
        b2 = group_bounds(T, t_lists, False, 1)
        return b1, b2, T, n_ref, ref_leading

    def smooth_block(self, X, chr_start, chr_len, ref_groups_local, ref_sizes=None, max_chunks=None, apply_log=True,
                     threshold=3.0, window=101, use_bounds=True, out=None):
        """run() steps 4, 8-12, 14 on this rank's cells.  ref_groups_local: per reference group the
        LOCAL column indices this rank owns (possibly empty); ref_sizes: GLOBAL group sizes
        (defaults to the local ones = single GPU).  With a process group initialised the
        partial sums of the two reference-mean steps are all-gathered (NCCL)."""
        C, G = X.shape
        Y = torch.empty_like(X) if out is None else out
        flag = torch.zeros(1, dtype=torch.int32, device=self.tdev)
        b1, b2, T, n_ref, ref_leading = self._reference_bounds(X, chr_start, chr_len, ref_groups_local, ref_sizes, max_chunks,
                                                               apply_log, threshold, window, use_bounds, flag)
        # pass 2: every local cell, one read and one write of the matrix
        if self.timing is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.reuse_reference_pass and 0 < n_ref < C and ref_leading:
            # the reference cells are the leading columns (the shard planner's layout): pass 1 already left their centred,
            # smoothed values in T, so they only need the second subtraction and 2^x (stage D, bit-identical to running
            # the whole pipeline again: x - 0.0 == x); the other cells take the full pipeline
            self.cell_pipeline(T[:n_ref], None, Y[:n_ref], chr_start, chr_len, False, None, 0.0, 0, 0, b2, True, use_bounds, flag)
            self.cell_pipeline(X[n_ref:], None, Y[n_ref:], chr_start, chr_len, apply_log, b1, threshold, window, 1, b2, True,
                               use_bounds, flag)
        else:
            self.cell_pipeline(X, None, Y, chr_start, chr_len, apply_log, b1, threshold, window, 1, b2, True, use_bounds, flag)
        if self.timing is not None:
            e1.record()
            self.timing.append(("cell_pipeline_pass2", e0, e1))
        return Y, flag

    def smooth_hmm_host(self, hX, hY, hS, dX, Y, states, chr_start, chr_len, ref_groups_local, ref_sizes, max_chunks, Pi, delta,
                        mean, sd, slab_cells=1024, apply_log=True, threshold=3.0, window=101, use_bounds=True):
        """Smooth block + per-cell HMM from / to pinned HOST tensors hX -> hY (float64) and hS (uint8), all (C, G), with the
        copies pipelined against the kernels over slabs of cells: the multi-rank counterpart of the C host pipeline
        (icnv_smooth_hmm_u8_f64, which knows no collectives).  The reference columns must be the leading local columns (the
        shard planner's layout); they go up first, the reference bounds are formed (two all-gathers across ranks), then
        H2D of slab i+1, pass 2 + Viterbi of slab i and D2H of slab i-1 overlap on three streams.  dX / Y / states are
        (C, G) device work buffers.  Bit-identical to smooth_block + viterbi on the uploaded matrix."""
        C, G = hX.shape
        refs = [np.asarray(g, dtype=np.int32) for g in ref_groups_local]
        n_ref = int(sum(len(g) for g in refs))
        if n_ref and not np.array_equal(np.concatenate(refs), np.arange(n_ref)):
            raise ValueError("smooth_hmm_host: the reference cells must be the leading local columns")
        on_gpu = dX.is_cuda
        flag = torch.zeros(1, dtype=torch.int32, device=self.tdev)
        if n_ref:
            dX[:n_ref].copy_(hX[:n_ref], non_blocking=True)
        b1, b2, _, _, _ = self._reference_bounds(dX, chr_start, chr_len, refs, ref_sizes, max_chunks, apply_log, threshold, window,
                                                 use_bounds, flag)
        if on_gpu:
            if getattr(self, "_copy_streams", None) is None:
                self._copy_streams = (torch.cuda.Stream(), torch.cuda.Stream())
            s_in, s_out = self._copy_streams
            main = torch.cuda.current_stream()
            s_in.wait_stream(main)     # the work buffers may still be in use by what was queued before this call
            s_out.wait_stream(main)
        flags = [flag]
        for c0 in range(0, C, slab_cells):
            c1 = min(C, c0 + slab_cells)
            u0 = max(c0, n_ref)        # the reference columns are already on the device
            if on_gpu:
                ev_in = torch.cuda.Event()
                with torch.cuda.stream(s_in):
                    if u0 < c1:
                        dX[u0:c1].copy_(hX[u0:c1], non_blocking=True)
                    ev_in.record(s_in)
                main.wait_event(ev_in)
            elif u0 < c1:
                dX[u0:c1].copy_(hX[u0:c1])
            self.cell_pipeline(dX[c0:c1], None, Y[c0:c1], chr_start, chr_len, apply_log, b1, threshold, window, 1, b2, True,
                               use_bounds, flag)
            _, f2 = self.viterbi(Y[c0:c1], chr_start, chr_len, Pi, delta, mean, sd, out=states[c0:c1])
            flags.append(f2)
            if on_gpu:
                ev_c = torch.cuda.Event()
                ev_c.record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_c)
                    hY[c0:c1].copy_(Y[c0:c1], non_blocking=True)
                    hS[c0:c1].copy_(states[c0:c1], non_blocking=True)
            else:
                hY[c0:c1].copy_(Y[c0:c1])
                hS[c0:c1].copy_(states[c0:c1])
        if on_gpu:
            main.wait_stream(s_out)
            torch.cuda.synchronize()
        return flags

    def viterbi(self, X, chr_start, chr_len, Pi, delta, mean, sd, out=None, want_margins=False):
        """Per-cell i6 / i3 Viterbi on device data -> uint8 states (C, G)."""
        C, G = X.shape
        cs, cl = _i32(chr_start), _i32(chr_len)
        Pi = np.asfortranarray(Pi, dtype=np.float64)
        m = Pi.shape[0]
        delta, mean, sd = (np.ascontiguousarray(v, dtype=np.float64) for v in (delta, mean, sd))
        st = torch.empty((C, G), dtype=torch.uint8, device=self.tdev) if out is None else out
        mg = torch.empty((C, len(cs)), dtype=torch.float64, device=self.tdev) if want_margins else None
        flag = torch.zeros(1, dtype=torch.int32, device=self.tdev)
        _lib.check(self.lib.icnv_dev_viterbi_f64(X.data_ptr(), G, C, cs.ctypes.data, cl.ctypes.data, len(cs), m,
                                                 Pi.ctypes.data, delta.ctypes.data, mean.ctypes.data, sd.ctypes.data, 0,
                                                 st.data_ptr(), mg.data_ptr() if mg is not None else None,
                                                 flag.data_ptr(), _stream_ptr()))
        return (st, flag, mg) if want_margins else (st, flag)

    def viterbi_groups(self, X, chr_start, chr_len, Pi, delta, mean, sds, groups_local, group_sizes=None, max_chunks=None, out=None):
        """Group modes of the HMM (predict_CNV_via_HMM_on_tumor_subclusters / _on_whole_tumor_samples and the i3 twins,
        R/inferCNV_HMM.R:345-408, 509-567): per group x = rowMeans(X[, group]) (:383), ONE trace per (group, chromosome),
        written to every cell of the group; cells in no group stay 255.  groups_local: per group this rank's LOCAL columns,
        a contiguous slice of the group's list cut at multiples of CHUNK (what plan_shards does for the groups it is given) or
        the whole group; group_sizes / max_chunks: the global sizes and the largest per-rank chunk count (defaults: this
        rank holds every group whole).  sds: m values per group (.get_state_emission_params), or m values for all.
        Across ranks the groups' chunk sums are all-gathered (32 groups per exchange) and combined in list order, every rank
        runs the few group sequences itself and scatters the states to its own cells: identical bits for any rank count."""
        C, G = X.shape
        cs, cl = _i32(chr_start), _i32(chr_len)
        Pi = np.asfortranarray(Pi, dtype=np.float64)
        m = Pi.shape[0]
        delta, mean = (np.ascontiguousarray(v, dtype=np.float64) for v in (delta, mean))
        groups_local = [np.asarray(g, dtype=np.int32) for g in groups_local]
        n_grp = len(groups_local)
        sizes = [len(g) for g in groups_local] if group_sizes is None else [int(v) for v in group_sizes]
        world = self._world()
        if max_chunks is None:
            if world > 1 and group_sizes is not None:
                raise ValueError("viterbi_groups: max_chunks is needed when groups are spread over ranks")
            max_chunks = [(len(g) + shard.CHUNK - 1) // shard.CHUNK for g in groups_local]
        sds = np.asarray(sds, dtype=np.float64).reshape(-1)
        if sds.size == m:
            sds = np.tile(sds, n_grp)
        sd_med = np.median(sds.reshape(n_grp, m), axis=1)          # object$pm$sd = median(object$pm$sd), HMM.R:1122
        means = torch.empty((n_grp, G), dtype=torch.float64, device=self.tdev)
        for k0 in range(0, n_grp, 32):
            k1 = min(n_grp, k0 + 32)
            rows = np.concatenate([[0], np.cumsum([int(max_chunks[k]) for k in range(k0, k1)])]).astype(np.int32)
            tot = max(int(rows[-1]), 1)
            packed = torch.zeros((tot, G), dtype=torch.float64, device=self.tdev)
            for k in range(k0, k1):
                n = len(groups_local[k])
                if n:
                    idx = torch.as_tensor(groups_local[k], device=self.tdev)
                    _lib.check(self.lib.icnv_dev_group_partial_sums_f64(X.data_ptr(), G, X.stride(0), idx.data_ptr(), n, shard.CHUNK, 0,
                                                                        packed.data_ptr() + 8 * G * int(rows[k - k0]), _stream_ptr()))
            part = packed
            if world > 1:
                part = torch.empty((world, tot, G), dtype=torch.float64, device=self.tdev)
                self._all_gather(part, packed)
            counts = np.asarray(sizes[k0:k1], dtype=np.int64)
            _lib.check(self.lib.icnv_dev_means_from_partials_f64(part.data_ptr(), G, world, tot, k1 - k0, rows.ctypes.data,
                                                                 counts.ctypes.data, means[k0].data_ptr(), _stream_ptr()))
        d_sd = torch.as_tensor(sd_med, device=self.tdev)
        gst = torch.empty((n_grp, G), dtype=torch.uint8, device=self.tdev)
        flag = torch.zeros(1, dtype=torch.int32, device=self.tdev)
        _lib.check(self.lib.icnv_dev_viterbi_f64(means.data_ptr(), G, n_grp, cs.ctypes.data, cl.ctypes.data, len(cs), m, Pi.ctypes.data,
                                                 delta.ctypes.data, mean.ctypes.data, d_sd.data_ptr(), 1, gst.data_ptr(), None,
                                                 flag.data_ptr(), _stream_ptr()))
        grp_of = np.full(C, -1, dtype=np.int32)
        for k, g in enumerate(groups_local):
            grp_of[g] = k
        d_grp_of = torch.as_tensor(grp_of, device=self.tdev)
        st = torch.empty((C, G), dtype=torch.uint8, device=self.tdev) if out is None else out
        _lib.check(self.lib.icnv_dev_scatter_group_states_u8(gst.data_ptr(), G, C, d_grp_of.data_ptr(), st.data_ptr(), _stream_ptr()))
        return st, flag

    def pairwise_dist(self, X, cells=None, out=None):
        """Euclidean distances between the listed local cells (rows of X; default all) as R's "dist" vector, on the device
        (`hclust(parallelDist(t(expr[, cells])))`, R/inferCNV_tumor_subclusters.R:191): n (n - 1) / 2 doubles."""
        G = X.shape[1]
        idx = None if cells is None else torch.as_tensor(np.asarray(cells, dtype=np.int32), device=self.tdev)
        n = X.shape[0] if idx is None else int(idx.numel())
        if out is None:
            out = torch.empty(n * (n - 1) // 2, dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.icnv_dev_pairwise_dist_f64(X.data_ptr(), G, X.stride(0), idx.data_ptr() if idx is not None else None,
                                                       n, out.data_ptr(), _stream_ptr()))
        return out

    def mean_sd(self, X, groups_local):
        """mu / sigma over all values of the listed cells across ALL ranks (.i3HMM_get_sd_trend_by_num_cells_fit,
        R/inferCNV_i3HMM.R:17-30).  groups_local: per group, this rank's LOCAL columns (the planner's slices).
        Per-cell (sum, sd) are computed on each GPU, all-gathered group by group in rank order - which restores
        the global list order - and combined by the routine the single-GPU entry point uses: identical bits for
        any rank count."""
        import torch.distributed as tdist
        G = X.shape[1]
        multi = self.collective and tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1
        world = tdist.get_world_size() if multi else 1
        lens = [len(g) for g in groups_local]
        cat = np.concatenate([np.asarray(g, dtype=np.int32) for g in groups_local]) if sum(lens) else np.zeros(0, np.int32)
        n = len(cat)
        stats = torch.zeros((2, max(n, 1)), dtype=torch.float64, device=self.tdev)
        if n:
            idx = torch.as_tensor(cat, device=self.tdev)
            _lib.check(self.lib.icnv_dev_column_stats_f64(X.data_ptr(), G, idx.data_ptr(), n, stats[0].data_ptr(),
                                                          stats[1].data_ptr(), _stream_ptr()))
        if not multi:
            allstats = stats[:, :n].cpu().numpy()
        else:
            mine = torch.tensor(lens, dtype=torch.int64, device=self.tdev)
            all_lens = [t.cpu().numpy() for t in self._all_gather_list(mine)]
            nmax = int(max(int(l.sum()) for l in all_lens))
            padded = torch.zeros((2, max(nmax, 1)), dtype=torch.float64, device=self.tdev)
            padded[:, :n] = stats[:, :n]
            gathered = [g.cpu().numpy() for g in self._all_gather_list(padded)]
            cols = []
            for k in range(len(lens)):          # group-major, rank-minor = the global list order
                for r in range(world):
                    off = int(all_lens[r][:k].sum())
                    cols.append(gathered[r][:, off:off + int(all_lens[r][k])])
            allstats = np.concatenate(cols, axis=1)
        sums, sds = np.ascontiguousarray(allstats[0]), np.ascontiguousarray(allstats[1])
        mu, sg = ct.c_double(), ct.c_double()
        self.lib.icnv_combine_cell_stats(sums.ctypes.data, sds.ctypes.data, len(sums), G, ct.addressof(mu), ct.addressof(sg))
        return mu.value, sg.value

    def median_filter(self, X, chr_start, chr_len, groups_local, window_size=7, out=None):
        C, G = X.shape
        cs, cl = _i32(chr_start), _i32(chr_len)
        off, idx = groups_to_csr(groups_local)
        Y = torch.empty_like(X) if out is None else out
        _lib.check(self.lib.icnv_dev_median_filter_f64(X.data_ptr(), Y.data_ptr(), G, C, cs.ctypes.data, cl.ctypes.data,
                                                       len(cs), off.ctypes.data, idx.ctypes.data, len(groups_local),
                                                       int(window_size), _stream_ptr()))
        return Y

    def median_filter_sharded(self, Xext, n_local, whole_lists, split_slices, chr_start, chr_len, window_size=7, out=None):
        """apply_median_filtering (R/noise_reduction.R:43-113) on a cell shard.  Xext: (n_local + n_scratch, G) - the
        shard's columns followed by scratch rows for the halo columns (n_scratch >= 2 r per split list, r = (window_size
        + 1) / 2).  whole_lists: index lists (LOCAL columns) that live entirely on this rank (tumour subclusters - the
        planner keeps them whole).  split_slices: for lists cut over ranks in list order (the reference groups, cut at
        chunk boundaries by plan_shards), this rank's contiguous slice as LOCAL columns (possibly empty).  The r entries
        either side of a slice are fetched from the neighbouring ranks (one all-gather of every rank's first / last r
        columns per split list) into the scratch rows; their own outputs land in scratch rows of `out` and are
        meaningless.  The first n_local rows of the result equal the single-GPU result bit for bit."""
        import torch.distributed as tdist
        r = (int(window_size) + 1) // 2
        G = Xext.shape[1]
        world = self._world()
        Y = torch.empty_like(Xext) if out is None else out
        lists = [np.asarray(v, dtype=np.int32) for v in whole_lists]
        n_split = len(split_slices)
        if world == 1 or n_split == 0:
            lists += [np.asarray(v, dtype=np.int32) for v in split_slices if len(v)]
            return self.median_filter(Xext, chr_start, chr_len, lists, window_size, out=Y)
        if Xext.shape[0] < n_local + 2 * r * n_split:
            raise ValueError("median_filter_sharded: not enough scratch rows behind the shard")
        rank = tdist.get_rank()
        # every rank's first / last r columns of every split list, and the slice lengths
        edge = torch.zeros((n_split, 2, r, G), dtype=torch.float64, device=self.tdev)
        lens = torch.zeros(n_split, dtype=torch.float64, device=self.tdev)
        for k, sl in enumerate(split_slices):
            n = len(sl)
            lens[k] = n
            if n:
                idx = torch.as_tensor(np.asarray(sl, dtype=np.int64), device=self.tdev)
                h = min(r, n)
                edge[k, 0, :h] = Xext[idx[:h]]
                edge[k, 1, r - h:] = Xext[idx[n - h:]]
        all_edge = torch.empty((world,) + tuple(edge.shape), dtype=torch.float64, device=self.tdev)
        all_lens = torch.empty((world, n_split), dtype=torch.float64, device=self.tdev)
        self._all_gather(all_edge, edge)
        self._all_gather(all_lens, lens)
        lens_h = all_lens.cpu().numpy().astype(np.int64)
        pos = n_local
        for k, sl in enumerate(split_slices):
            if not len(sl):
                continue
            prev, nxt = shard.halo_sources(lens_h[:, k], rank, r)
            ext = []
            for q, cnt in prev:          # the last cnt columns of rank q's slice sit at the end of its tail block
                Xext[pos:pos + cnt] = all_edge[q, k, 1, r - cnt:]
                ext += list(range(pos, pos + cnt))
                pos += cnt
            ext += [int(v) for v in sl]
            for q, cnt in nxt:
                Xext[pos:pos + cnt] = all_edge[q, k, 0, :cnt]
                ext += list(range(pos, pos + cnt))
                pos += cnt
            lists.append(np.asarray(ext, dtype=np.int32))
        return self.median_filter(Xext, chr_start, chr_len, lists, window_size, out=Y)

    # ---- CNV region calling on device-resident states (R/inferCNV_HMM.R:706-1087) --------------------------------------
    def state_counts(self, S: torch.Tensor, groups_local) -> torch.Tensor:
        """(n_grp, G, 8) int32 counts of the listed LOCAL cells' states per gene (slot 0: unassigned, v + 1: state v)."""
        C, G = S.shape
        off, idx = groups_to_csr(groups_local)
        n_grp = len(groups_local)
        counts = torch.empty((n_grp, G, 8), dtype=torch.int32, device=self.tdev)
        d_idx = torch.as_tensor(idx if len(idx) else np.zeros(1, np.int32), device=self.tdev)   # a rank may own no listed cell
        flag = torch.zeros(1, dtype=torch.int32, device=self.tdev)
        _lib.check(self.lib.icnv_dev_state_counts_u8(S.data_ptr(), G, S.stride(0), d_idx.data_ptr(), off.ctypes.data, n_grp,
                                                     counts.data_ptr(), flag.data_ptr(), _stream_ptr()))
        if int(flag.item()) & 4:
            raise ValueError("state outside 0..6 / 255 in the state matrix")
        return counts

    def state_consensus(self, S: torch.Tensor, groups_local) -> torch.Tensor:
        """.get_state_consensus for every group -> (n_grp, G) uint8.  groups_local: per group, this rank's LOCAL
        columns (possibly empty).  With a process group the integer counts are summed over ranks (all-reduce), so
        the consensus is the same for any partition of a group's cells."""
        import torch.distributed as tdist
        counts = self.state_counts(S, groups_local)
        if self.collective and tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1:
            self._all_reduce_sum(counts)
        n_grp, G, _ = counts.shape
        cons = torch.empty((n_grp, G), dtype=torch.uint8, device=self.tdev)
        _lib.check(self.lib.icnv_dev_consensus_from_counts(counts.data_ptr(), G, n_grp, cons.data_ptr(), _stream_ptr()))
        return cons

    def cnv_regions(self, seqs: torch.Tensor, chr_start, chr_len, gene_start, gene_stop, cols=None) -> dict:
        """Run-length regions + bounds of the rows of `seqs` ((n_seq, G) uint8: consensus sequences, or the state
        matrix itself with `cols` = the cells to report, by = "cell").  Returns host arrays (api.REGION_FIELDS)."""
        from .api import REGION_FIELDS
        n_rows, G = seqs.shape
        cs, cl = _i32(chr_start), _i32(chr_len)
        gs, ge = (np.ascontiguousarray(v, dtype=np.float64) for v in (gene_start, gene_stop))
        d_cols = torch.as_tensor(np.asarray(cols, dtype=np.int32), device=self.tdev) if cols is not None else None
        n_seq = int(d_cols.numel()) if d_cols is not None else n_rows
        n = ct.c_int64(0)
        _lib.check(self.lib.icnv_dev_cnv_regions_u8(seqs.data_ptr(), G, seqs.stride(0), n_seq,
                                                    d_cols.data_ptr() if d_cols is not None else None, cs.ctypes.data,
                                                    cl.ctypes.data, len(cs), gs.ctypes.data, ge.ctypes.data,
                                                    ct.addressof(n), _stream_ptr()))
        out = {k: np.empty(int(n.value), dtype=dt) for k, dt in REGION_FIELDS}
        _lib.check(self.lib.icnv_cnv_regions_fetch(int(n.value), *[out[k].ctypes.data for k, _ in REGION_FIELDS]))
        return out

    def launch_count(self) -> int:
        return int(self.lib.icnv_launch_count())
