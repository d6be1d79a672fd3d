#!/usr/bin/env python
"""Turn the files a `bash tools/_final_profile.sh` run left in gpurun_out/ into the committed
summaries under profiles/ (bench lines, launch-share table, ncu --set full digests, hot lines,
DRAM traffic per launch).  Run here (no GPU needed: ncu -i reads the .ncu-rep)."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "r01"
TRAFFIC_ONLY = "--traffic-only" in sys.argv   # on the GPU box, before bench.py: it quotes profiles/<tag>_traffic.json
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__inst_executed.sum.per_cycle_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem']


def ncu(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


traffic = {}
for f, name in (("cellpipe", "cell_pipeline_pass2"), ("vfast", "viterbi_fast"), ("medfilt", "median_filter_w7")):
    rep = os.path.join(GO, f"{TAG}_prof_{f}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = list(csv.reader(ncu([rep, "--page", "raw", "--csv"]).splitlines()))
    hh, uu, vv = raw[0], raw[1], raw[2]
    d = {}
    with open(os.path.join(PR, f"{TAG}_{f}_ncu_summary.csv"), "w") as out:
        out.write("metric,unit,value\n")
        for i, n in enumerate(hh):
            if n in KEEP or ("issue_stalled" in n and "per_issue_active" in n):
                out.write(f"{n},{uu[i]},{vv[i]}\n")
                d[n] = (uu[i], vv[i])
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    b = sum(float(d[k][1]) * mult[d[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    traffic[name] = {"dram_bytes_per_launch": b, "ncu_ms": float(d["gpu__time_duration.sum"][1]),
                     "source": f"profiles/{TAG}_{f}_ncu_summary.csv (ncu --set full, one launch)"}
    src = ncu([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    tmp = os.path.join(GO, f"_{f}_src.csv")
    open(tmp, "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(PR, "src_hotspots.py"), tmp, "30"], capture_output=True, text=True).stdout
    open(os.path.join(PR, f"{TAG}_{f}_hotspots.txt"), "w").write(hot)
json.dump(traffic, open(os.path.join(PR, f"{TAG}_traffic.json"), "w"), indent=1)
if TRAFFIC_ONLY:
    sys.exit(0)
for f in ("bench", "bench_reference"):
    if os.path.exists(os.path.join(GO, f"{TAG}_{f}.json")):
        shutil.copy(os.path.join(GO, f"{TAG}_{f}.json"), os.path.join(PR, f"{TAG}_{f}.json"))
shutil.copy(os.path.join(GO, f"{TAG}_launches.csv"), os.path.join(PR, f"{TAG}_launches.csv"))

rows = list(csv.reader(open(os.path.join(GO, f"{TAG}_launches.csv"))))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e6
ours = {k: v for k, v in agg.items() if "icnv::" in k and "synth" not in k}
tot = sum(v[1] for v in ours.values())
with open(os.path.join(PR, f"{TAG}_launches_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline\n"
            "(5 steps captured = 3 warm-up + 2 timed; per-launch times are cold-cache and serialised: compare SHARES)\n\n"
            f"{'total ms':>10s} {'launches':>8s} {'ms/launch':>10s} {'share':>7s}  kernel\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = f"{100 * t / tot:6.1f}%" if k in ours else "   n/a "
        f.write(f"{t:10.3f} {n:8d} {t / n:10.4f} {share}  {k[:120]}\n")

b = json.load(open(os.path.join(PR, f"{TAG}_bench.json")))
print("value", b["value"], "ms/step", b["ms_per_step"], "e2e", b["e2e"]["value"], b["e2e"]["ms_per_step"], "fused", b["e2e"]["fused_call"]["ms_per_step"])
print("roofline", b["roofline"]["frac"], b["roofline"]["ms_per_launch"], "hmm", b["roofline_hmm"]["frac"], b["roofline_hmm"]["ms_per_launch"])
print("cpu", b["cpu_baseline"]["value"], "ref arm", json.load(open(os.path.join(PR, f"{TAG}_bench_reference.json")))["value"])
print(traffic)
print(open(os.path.join(PR, f"{TAG}_launches_summary.txt")).read()[:1200])
