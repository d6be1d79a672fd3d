#!/usr/bin/env python
"""bench.py - cells x genes / s through the smooth block + i6 HMM (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells C] [--genes G]

One "step" = one pass of the hot path over the workload: run() steps 4, 8-12, 14 (fused smooth
block) followed by step 17 (per-cell 6-state Viterbi) on a synthetic depth-normalised matrix.

Workload at N = 1: BASELINE.json configs[1] - 10 000 cells x 10 000 genes, 22 chromosomes,
window 101, i6 HMM (0.8 GB of float64 per pass, far larger than the 126 MB L2, so consecutive
timed steps cannot hit in cache).  N > 1 (torchrun, one rank per GPU): weak scaling, every rank
holds 10 000 cells of an N x 10 000-cell run; cells are sharded, the only exchange is the NCCL
all-gather of the reference-mean partial sums.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over
ranks); `e2e` = the same metric through the host-pointer C ABI with pinned host buffers, H2D and
D2H inside the timed region; `roofline` = the dominant kernel (Viterbi) against measured HBM
bandwidth; `cpu_baseline` = the C oracle port on the host cores over a bounded sample.

`--impl reference` times the CPU restatement of the reference's algorithm (oracle/; the reference
itself is interpreted R and R is not installable in this image) with all host threads on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cells x genes / s (smooth block + i6 HMM)"
UNIT = "cell-genes/s"
CHR_TEMPLATE = [852, 615, 535, 288, 420, 453, 458, 297, 349, 363, 514, 472, 162, 301, 274, 397, 546, 126, 545, 239, 90,
                212]  # oligodendroglioma example, SURVEY section 8(d)
I6_MEAN = np.array([0.41234766, 0.84075773, 1.01693983, 1.12238786, 1.23842619, 1.44298781])
I6_SD = np.array([0.028893, 0.164549, 0.105553, 0.190574, 0.244093, 0.290072])
SEED = 20260922 + 1
# algorithmic HBM bytes per cell-gene (SURVEY section 8d, FP64 parity mode)
BYTES_SMOOTH = 16.0          # + 16 per reference-cell gene for the two reference pre-passes
BYTES_HMM = 9.0              # 8 read + 1 state byte written


def chr_layout(G: int):
    t = np.array(CHR_TEMPLATE, dtype=np.float64)
    lens = np.floor(t * G / t.sum()).astype(np.int64)
    lens[0] += G - lens.sum()
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return starts.astype(np.int32), lens.astype(np.int32)


def ref_groups_global(C_total: int):
    """first 10 % of the cells are reference cells in two groups (6 % / 4 %)."""
    a, b = int(round(0.06 * C_total)), int(round(0.10 * C_total))
    return [np.arange(0, a, dtype=np.int64), np.arange(a, b, dtype=np.int64)]


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), or None."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Clock / throttle-reason sampling during the timed region (B200_PROFILING.md clocks line).

    Two sources run side by side from before the warm-up until after the timed region: an `nvidia-smi -lms 20`
    child writing CSV to a temp file, and an NVML polling thread in this process (pynvml, same counters).  Samples
    are selected by wall-clock stamp inside [t0, t1]; nvidia-smi rows win when any landed there, else the NVML
    thread's, else everything captured under the (identical) warm-up load - `window` / `source` say which."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    NVML_BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int):
        self.proc = None
        self.path = None
        self.gpu = gpu_index
        self.nvml_rows = []          # (stamp, sm_mhz, sm_max_mhz, reason mask)
        self._stop = None
        self._thread = None

    def _nvml_loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop.is_set():
                try:
                    self.nvml_rows.append((time.time(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), mx,
                                           int(reasons(h))))
                except Exception:
                    pass
                self._stop.wait(0.02)
        except Exception:
            return

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        try:
            import threading
            self._stop = threading.Event()
            self._thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        """t0 / t1: wall-clock bounds (time.time()) of the timed region.  The nvidia-smi child is left running for
        up to a second past t1 so that its stdio buffer holding the timed region's rows reaches the file before it
        is terminated."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if t1 is not None:
            time.sleep(max(0.0, min(1.0, t1 + 1.0 - time.time())))
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        smi = []                     # (stamp, sm, max, [reason names])
        if self.proc is not None:
            try:
                self.proc.terminate()      # exact PID we started
                self.proc.wait(timeout=5)
            except Exception:
                pass
            try:
                for line in open(self.path):
                    r = line.strip().split(", ")
                    if len(r) < 9:
                        continue
                    try:
                        smi.append((self._stamp(r[0]), float(r[1]), float(r[2]),
                                    [n for n, v in zip(self.NAMES, r[5:9]) if v.strip().lower().startswith("active")]))
                    except ValueError:
                        continue
                os.unlink(self.path)
            except Exception:
                pass
        nvml = [(t, sm, mx, [n for n, b in self.NVML_BITS.items() if mask & b]) for t, sm, mx, mask in self.nvml_rows]

        def inside(rows):
            if t0 is None or t1 is None:
                return rows
            return [r for r in rows if r[0] is not None and t0 <= r[0] <= t1]
        for rows, source, window in ((inside(smi), "nvidia-smi", "timed region"),
                                     (inside(nvml), "nvml", "timed region"),
                                     (smi, "nvidia-smi", "warm-up + timed region (no sample was stamped inside the timed region)"),
                                     (nvml, "nvml", "warm-up + timed region (no sample was stamped inside the timed region)")):
            if rows:
                out = {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
                       "reasons": sorted({n for r in rows for n in r[3]}), "samples": len(rows), "window": window,
                       "source": source}
                break
        return out


def run_reference(args):
    """CPU arm: the oracle port (all host threads) on a bounded, stratified sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    G = args.genes
    C_total = args.cells * max(1, args.gpus)
    n_sample = args.ref_sample_cells
    cs, cl = chr_layout(G)
    stride = max(1, C_total // n_sample)
    cells = np.arange(0, C_total, stride, dtype=np.int64)[:n_sample]
    X = sample_matrix(G, cs, cl, cells, C_total)
    refs_g = ref_groups_global(C_total)
    ref_local = [np.flatnonzero(np.isin(cells, g)).astype(np.int32) for g in refs_g]
    ref_local = [g for g in ref_local if len(g)]
    Pi, delta = orc.hmm_params(6)
    nt = orc.max_threads()

    def step():
        S = orc.smooth_block(X, cs, cl, ref_local, apply_log=True, threshold=3.0, window=101, nthreads=nt)
        orc.viterbi_matrix(S, cs, cl, Pi, delta, I6_MEAN, I6_SD, nthreads=nt)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = G * len(cells) / dt
    sample = f"{len(cells)} cells (every {stride}th of {C_total}) x {G} genes per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, C_total),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nt, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = interpreted R (not installable here); this arm is the C restatement in oracle/ "
                "with OpenMP over cells - a best-case CPU line, the R path itself is single-threaded",
    }))


def sample_matrix(G, cs, cl, cells, C_total):
    """Workload values for the given global cells as a host (G, n) Fortran array.  Uses the same
    counter-based generator as the GPU arm when a GPU is present (data prep, not timed)."""
    try:
        import torch
        if torch.cuda.is_available():
            from infercnv_b200.device import Engine
            eng = Engine(int(os.environ.get("LOCAL_RANK", "0")))
            X = eng.synth(G, cs, cl, cells, C_total, SEED)
            torch.cuda.synchronize()
            return np.asfortranarray(X.cpu().numpy().T)
    except Exception as e:  # pragma: no cover
        print(f"[bench] GPU generator unavailable ({e}); using the NumPy stand-in", file=sys.stderr)
    rng = np.random.default_rng(SEED)
    m_g = rng.lognormal(0.5, 1.0, size=(G, 1))
    f_c = rng.lognormal(0.0, 0.2, size=(1, len(cells)))
    lam = rng.gamma(10.0, (m_g * f_c) / 10.0)
    return np.asfortranarray(rng.poisson(lam).astype(np.float64))


def workload_config(args, C_total):
    return {"workload": f"synthetic {C_total} cells x {args.genes} genes, 22 chromosomes, window 101, i6 HMM "
                        f"(BASELINE configs[1] per GPU)",
            "cells_per_gpu": args.cells, "genes": args.genes, "window_length": 101, "hmm": "i6 per cell, t=1e-6",
            "reference_cells": "first 10 % in two groups (6 % / 4 %)", "sharding": f"cells over {args.gpus} GPU(s)",
            "l2": "inputs (%.2f GB/GPU/pass) %s the 126 MB L2; no explicit flush" % (
                args.cells * args.genes * 8 / 1e9, "exceed" if args.cells * args.genes * 8 > 126e6 else "do NOT exceed")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=10000, help="cells per GPU")
    ap.add_argument("--genes", type=int, default=10000)
    ap.add_argument("--ref-sample-cells", type=int, default=512)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from infercnv_b200 import api, dist as shard
    from infercnv_b200.device import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    eng = Engine(local_rank)
    sampler = ClockSampler(local_rank)   # started now: nvidia-smi takes longer to come up than a short timed region lasts
    if rank == 0:
        sampler.start()
    G = args.genes
    C_total = args.cells * world
    cs, cl = chr_layout(G)
    refs_g = ref_groups_global(C_total)
    plan = shard.plan_shards(C_total, refs_g, world)[rank]
    cells = plan.local_cells
    C_local = len(cells)
    X = eng.synth(G, cs, cl, cells, C_total, SEED)
    ref_local = plan.local_ref_groups()
    from infercnv_b200.ops import CNV_LEVELS, get_HMM
    Pi, delta, _, _ = get_HMM({k: {"mean": m, "sd": s} for k, m, s in zip(CNV_LEVELS, I6_MEAN, I6_SD)}, 1e-6)
    Y = torch.empty_like(X)
    states = torch.empty((C_local, G), dtype=torch.uint8, device=X.device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    flags = []

    def step(record=None):
        a, b, c = (ev(), ev(), ev()) if record is not None else (None, None, None)
        if record is not None:
            a.record()
        _, f1 = eng.smooth_block(X, cs, cl, ref_local, plan.ref_sizes, plan.max_chunks, apply_log=True, threshold=3.0,
                                 window=101, use_bounds=True, out=Y)
        if record is not None:
            b.record()
        _, f2 = eng.viterbi(Y, cs, cl, Pi, delta, I6_MEAN, I6_SD, out=states)
        if record is not None:
            c.record()
            record.append((a, b, c))
        flags.append((f1, f2))

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.launch_count()
    rec = []
    eng.timing = []
    barrier()
    wall0 = time.time()
    t_start, t_end = ev(), ev()
    t_start.record()
    for _ in range(args.steps):
        step(rec)
    t_end.record()
    barrier()
    wall1 = time.time()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_total = t_start.elapsed_time(t_end)
    ms_step = ms_total / args.steps
    ms_smooth = float(np.mean([a.elapsed_time(b) for a, b, _ in rec]))
    ms_hmm = float(np.mean([b.elapsed_time(c) for _, b, c in rec]))
    ms_pass2 = float(np.mean([a.elapsed_time(b) for _, a, b in eng.timing]))
    eng.timing = None
    for f1, f2 in flags:
        if int(f1.item()) or int(f2.item()):
            raise SystemExit("non-finite / underflow flag raised during the benchmark")
    smin, smax = int(states.min().item()), int(states.max().item())
    assert 1 <= smin and smax <= 6, (smin, smax)

    # ---- end to end through the host-facing API, pinned host buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hX = torch.empty((C_local, G), dtype=torch.float64, pin_memory=True)
        hX.copy_(X)
        hY = torch.empty((C_local, G), dtype=torch.float64, pin_memory=True)
        hS = torch.empty((C_local, G), dtype=torch.uint8, pin_memory=True)   # one byte per state, as the R shim asks for
        torch.cuda.synchronize()
        xn, yn, sn = (t.numpy().T for t in (hX, hY, hS))   # (G, C) Fortran views of the pinned buffers
        if world == 1:
            off, idx = api.groups_to_csr(ref_local)

            def e2e_step():
                api.smooth_block(xn, cs, cl, ref_local, apply_log=True, threshold=3.0, window_length=101, out=yn)
                api.viterbi(yn, cs, cl, Pi, delta, I6_MEAN, I6_SD, out=sn)
            h2d = 2 * C_local * G * 8
            d2h = C_local * G * 8 + C_local * G

            def fused_step():
                api.smooth_hmm(xn, cs, cl, ref_local, Pi, delta, I6_MEAN, I6_SD, out=yn, out_states=sn)
        else:
            fused_step = None
            dX = torch.empty_like(X)
            hS8 = torch.empty((C_local, G), dtype=torch.uint8, pin_memory=True)

            # ICNV_BENCH_E2E_PIPELINE=1: copies pipelined against the kernels over cell slabs (Engine.smooth_hmm_host, written
            # after the last GPU session: opt-in until it has been timed); default: upload, compute, download in sequence
            pipelined = os.environ.get("ICNV_BENCH_E2E_PIPELINE", "0") == "1"

            def e2e_step():
                if pipelined:
                    eng.smooth_hmm_host(hX, hY, hS8, dX, Y, states, cs, cl, ref_local, plan.ref_sizes, plan.max_chunks, Pi, delta,
                                        I6_MEAN, I6_SD)
                    return
                dX.copy_(hX, non_blocking=True)
                eng.smooth_block(dX, cs, cl, ref_local, plan.ref_sizes, plan.max_chunks, out=Y)
                eng.viterbi(Y, cs, cl, Pi, delta, I6_MEAN, I6_SD, out=states)
                hY.copy_(Y, non_blocking=True)
                hS8.copy_(states, non_blocking=True)
                torch.cuda.synchronize()
            h2d = C_local * G * 8
            d2h = C_local * G * 8 + C_local * G
        n_e2e = max(2, min(args.steps, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=X.device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": G * C_total / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(dt.item()) * 1e3, "steps": n_e2e,
               "api": "icnv_smooth_block_f64 + icnv_viterbi_u8_f64 (host pointers, two calls as the R shim makes them)"
                      if world == 1 else ("Engine.smooth_hmm_host (slab-pipelined copies) with pinned host tensors" if pipelined
                                          else "Engine.smooth_block/viterbi with pinned host tensors")}
        if fused_step is not None:   # one upload instead of two: the optional fused entry point
            fused_step()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                fused_step()
            dtf = (time.perf_counter() - t0) / n_e2e
            e2e["fused_call"] = {"value": G * C_total / dtf, "ms_per_step": dtf * 1e3, "api": "icnv_smooth_hmm_u8_f64",
                                 "h2d_bytes_per_step": int(C_local * G * 8 * 1.1),
                                 "d2h_bytes_per_step": int(C_local * G * 9)}

    # ---- max over ranks ----------------------------------------------------------------------------------------
    t = torch.tensor([ms_step, ms_smooth, ms_hmm, ms_pass2], dtype=torch.float64, device=X.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_smooth, ms_hmm, ms_pass2 = (float(v) for v in t.tolist())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        n_ref_local = sum(len(g) for g in ref_local)
        hmm_bytes = BYTES_HMM * C_local * G
        smooth_bytes = BYTES_SMOOTH * C_local * G + 16.0 * n_ref_local * G
        pass2_bytes = BYTES_SMOOTH * C_local * G
        ach_p2 = pass2_bytes / (ms_pass2 * 1e-3) / 1e9
        ach_hmm = hmm_bytes / (ms_hmm * 1e-3) / 1e9
        reruns = int(api.hmm_rerun_count())
        out = {
            "metric": METRIC, "value": G * C_total / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, C_total),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            # dominant kernel of the step: the fused per-cell pipeline over all cells (pass 2)
            "roofline": {"kernel": "cell_pipeline3_kernel pass 2 (%.0f %% of the step)" % (100 * ms_pass2 / ms_step),
                         "bound": "hbm", "achieved": ach_p2, "peak": peak, "unit": "GB/s", "frac": ach_p2 / peak,
                         "traffic": ncu_traffic("cell_pipeline_pass2") if (C_local, G) == (10000, 10000) else None,
                         "traffic_source": "profiles/r01_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch of this workload; "
                                           "captured before the kernel's instruction diet - the memory traffic, one read and one write of the matrix, is unchanged by it)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": pass2_bytes,
                         "ms_per_launch": ms_pass2,
                         "note": "16 B per cell-gene (one FP64 read, one FP64 write); instruction-issue bound, see DESIGN.md"},
            "roofline_hmm": {"kernel": "viterbi_fast_kernel<6> + exact re-run list (%.0f %% of the step)" % (100 * ms_hmm / ms_step),
                             "bound": "hbm", "achieved": ach_hmm, "peak": peak, "unit": "GB/s", "frac": ach_hmm / peak,
                             "traffic": ncu_traffic("viterbi_fast") if (C_local, G) == (10000, 10000) else None,
                             "algorithmic_bytes_per_launch": hmm_bytes, "ms_per_launch": ms_hmm,
                             "sequences_rerun_in_reference_order_arithmetic": reruns,
                             "sequences": int(C_local * len(cs)),
                             "note": "9 B per cell-gene (8 read + 1 state byte); instruction-issue / shared-memory-table bound, see DESIGN.md"},
            "roofline_smooth_block": {"kernels": "group means + cell_pipeline pass 1 (reference cells) + pass 2",
                                      "bound": "hbm", "achieved": smooth_bytes / (ms_smooth * 1e-3) / 1e9, "peak": peak,
                                      "unit": "GB/s", "frac": smooth_bytes / (ms_smooth * 1e-3) / 1e9 / peak,
                                      "algorithmic_bytes_per_step": smooth_bytes, "ms_per_step": ms_smooth},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, G, cs, cl, C_total)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, G, cs, cl, C_total):
    """Oracle port on the host cores over a bounded stratified sample (rank 0, N = 1 only)."""
    from oracle import oracle as orc
    n = args.ref_sample_cells
    stride = max(1, C_total // n)
    cells = np.arange(0, C_total, stride, dtype=np.int64)[:n]
    X = sample_matrix(G, cs, cl, cells, C_total)
    refs_g = ref_groups_global(C_total)
    ref_local = [np.flatnonzero(np.isin(cells, g)).astype(np.int32) for g in refs_g]
    ref_local = [g for g in ref_local if len(g)]
    Pi, delta = orc.hmm_params(6)
    nt = orc.max_threads()
    reps, t_total = 0, 0.0
    while t_total < 10.0 and reps < 50:
        t0 = time.perf_counter()
        S = orc.smooth_block(X, cs, cl, ref_local, nthreads=nt)
        orc.viterbi_matrix(S, cs, cl, Pi, delta, I6_MEAN, I6_SD, nthreads=nt)
        t_total += time.perf_counter() - t0
        reps += 1
    return {"value": reps * G * len(cells) / t_total, "unit": UNIT, "cores": nt, "kind": "port",
            "sample": f"{len(cells)} cells (every {stride}th of {C_total}) x {G} genes, {reps} passes, {t_total:.1f} s"}


if __name__ == "__main__":
    main()
