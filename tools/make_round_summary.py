#!/usr/bin/env python
"""Turn one run of tools/final_profile.sh (gpurun_out/<tag>_*) into the committed evidence under profiles/:

    python tools/make_round_summary.py <tag> [<round label, default r02>]

  profiles/<r>_<kernel>_ncu_summary.csv + _hotspots.txt   (tools/profile_digest.py on every <tag>_prof_*.ncu-rep)
  profiles/<r>_traffic.json                                DRAM bytes per launch of the hot kernels (what bench.py's `traffic` reads)
  profiles/<r>_bench_<config>.json, <r>_bench_reference_c3.json, <r>_secondary_kernels.json, <r>_scaling_probe.txt,
  profiles/<r>_launches_c3.csv, <r>_pytest_gpu.txt         copies of the session's outputs
  profiles/<r>_sass_excerpts.txt                           cuobjdump lines that show the bulk-copy / mbarrier / cp.async instructions
  profiles/<r>_summary.md                                  one table per configuration
Runs here, without a GPU."""
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
R = sys.argv[2] if len(sys.argv) > 2 else "r02"
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")


def first_json(path):
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                pass
    return None


def digest(rep, name):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_digest.py"), rep, os.path.join(PR, f"{R}_{name}")],
                         capture_output=True, text=True)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line)
    print(f"digest of {rep} failed: {out.stderr[-300:]}", file=sys.stderr)
    return None


reps = {"cellpipe": "cell_pipeline_pass2", "vfast": "viterbi_fast", "medfilt": "median_filter_1e8", "medfilt_c4": "median_filter",
        "dist": "pairwise_dist", "cellpipe4_c5": "cell_pipeline4_c5_8000_cells"}
dig = {}
for short, key in reps.items():
    rep = os.path.join(GO, f"{tag}_prof_{short}.ncu-rep")
    if os.path.exists(rep):
        d = digest(rep, short)
        if d:
            dig[key] = d
traffic = {"_source": f"ncu --set full --clock-control none, one launch each, session {tag} (tools/final_profile.sh); bytes = "
                      "dram__bytes_read.sum + dram__bytes_write.sum",
           "c3": {k: dig[k] for k in ("cell_pipeline_pass2", "viterbi_fast") if k in dig},
           "c4": {k: dig[k] for k in ("cell_pipeline_pass2", "median_filter") if k in dig},
           "c2_size": {k: dig[k] for k in ("median_filter_1e8", "pairwise_dist") if k in dig},
           "c5_8000_cells": {k: dig[k] for k in ("cell_pipeline4_c5_8000_cells",) if k in dig}}
json.dump(traffic, open(os.path.join(PR, f"{R}_traffic.json"), "w"), indent=1)

for src, dst in [(f"{tag}_bench_{c}.json", f"{R}_bench_{c}.json") for c in ("c2", "c3", "c4", "c5")] + \
        [(f"{tag}_bench_reference_c3.json", f"{R}_bench_reference_c3.json"), (f"{tag}_secondary_kernels.json", f"{R}_secondary_kernels.json"),
         (f"{tag}_scaling_probe.txt", f"{R}_scaling_probe.txt"), (f"{tag}_launches_c3.csv", f"{R}_launches_c3.csv"),
         (f"{tag}_pytest_gpu.log", f"{R}_pytest_gpu.txt")]:
    if os.path.exists(os.path.join(GO, src)):
        shutil.copy(os.path.join(GO, src), os.path.join(PR, dst))

# SASS evidence of the asynchronous copy instructions, from the library that was measured
so = os.path.join(ROOT, "infercnv_b200", "libinfercnv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keep, fn = [], ""
for line in sass.splitlines():
    if "Function :" in line:
        fn = line.strip()
    if any(t in line for t in ("UBLKCP", "SYNCS", "LDGSTS", "UTMA")):
        keep.append((fn, line.split("/*")[1].split("*/")[1].strip() if line.count("/*") >= 2 else line.strip()))
with open(os.path.join(PR, f"{R}_sass_excerpts.txt"), "w") as f:
    f.write("# cuobjdump -sass infercnv_b200/libinfercnv_b200.so: bulk copy (UBLKCP), mbarrier (SYNCS) and cp.async (LDGSTS) instructions per kernel\n")
    last, n = None, 0
    for fn, ins in keep:
        if fn != last:
            f.write(f"\n{fn}\n")
            last, n = fn, 0
        if n < 6:
            f.write(f"    {ins}\n")
        n += 1

# summary table
lines = [f"# Round 2 measurements (session `{tag}`, one B200; every file named here is in `profiles/`)", ""]
for c in ("c3", "c2", "c4", "c5"):
    d = first_json(os.path.join(PR, f"{R}_bench_{c}.json"))
    if not d:
        continue
    st = d.get("stage_ms", {})
    lines += [f"## {c}: {d['config'].get('workload', '')}", "",
              f"* device-resident: **{d['ms_per_step']:.2f} ms per step = {d['value']:.3e} {d['unit']}** "
              f"(smooth block {st.get('smooth_block')}, HMM {st.get('hmm')}, median filter {st.get('median_filter')} ms); "
              f"clocks {d.get('clocks', {}).get('sm_mhz')} MHz, reasons {d.get('clocks', {}).get('reasons')}"]
    for k in ("roofline_cell_pipeline", "roofline_hmm", "roofline_median_filter"):
        r = d.get(k)
        if r:
            lines.append(f"* {r['kernel']}: {r['ms_per_launch']:.3f} ms, {r['achieved']:.0f} GB/s algorithmic = **{r['frac']:.3f}** of {r['peak']} GB/s "
                         f"({r['peak_source']}); DRAM traffic per launch {r.get('traffic')}")
    e = d.get("e2e")
    if e:
        lines.append(f"* end to end through the host ABI ({e.get('host_memory')}): {e['ms_per_step']:.1f} ms = {e['value']:.3e} {d['unit']} "
                     f"({e['h2d_bytes_per_step'] / 1e9:.2f} GB in, {e['d2h_bytes_per_step'] / 1e9:.2f} GB out per step)")
    cb = d.get("cpu_baseline")
    if cb:
        lines.append(f"* CPU arm ({cb.get('kind')}, {cb.get('cores')} threads, {cb.get('sample')}): {cb['value']:.3e} {cb['unit']}")
    lines.append("")
open(os.path.join(PR, f"{R}_summary.md"), "w").write("\n".join(lines))
print(open(os.path.join(PR, f"{R}_summary.md")).read()[:1500])
