#!/usr/bin/env python
"""Digest of one `ncu --set full --import-source on` report into the committed evidence under profiles/:

    python tools/profile_digest.py gpurun_out/X.ncu-rep profiles/r02_<name>

writes <prefix>_ncu_summary.csv (time, DRAM bytes, issue / pipe utilisation, occupancy limiters, shared-memory wavefronts
and bank conflicts, every warp-stall reason per issued instruction), <prefix>_hotspots.txt (source lines by warp-stall
samples and executed instructions) and prints the DRAM bytes per launch (for profiles/r02_traffic.json).
Runs here, without a GPU: `ncu -i` only reads the report."""
import csv
import json
import os
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active']
MULT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def ncu(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(rep, prefix):
    raw = list(csv.reader(ncu([rep, "--page", "raw", "--csv"]).splitlines()))
    hh, uu, vv = raw[0], raw[1], raw[2]
    d = {}
    with open(prefix + "_ncu_summary.csv", "w") as out:
        out.write("metric,unit,value\n")
        for i, n in enumerate(hh):
            if n in KEEP or ("issue_stalled" in n and "per_issue_active" in n):
                out.write(f'{n},{uu[i]},"{vv[i]}"\n')
                d[n] = (uu[i], vv[i])
    dram = sum(float(d[k][1]) * MULT[d[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    src = ncu([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    tmp = prefix + "_src.tmp.csv"
    open(tmp, "w").write(src)
    here = os.path.dirname(os.path.abspath(__file__))
    hot = subprocess.run([sys.executable, os.path.join(os.path.dirname(here), "profiles", "src_hotspots.py"), tmp, "40"],
                         capture_output=True, text=True).stdout
    os.unlink(tmp)
    open(prefix + "_hotspots.txt", "w").write(f"# {d['Kernel Name'][1]}  ({os.path.basename(rep)})\n"
                                              "# note: ncu attributes an inlined function's instructions to its own line AND to the line that\n"
                                              "# inlines it; the percentages are shares of that (double-counting) total\n" + hot)
    unit, val = d["gpu__time_duration.sum"]
    ms = float(val) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(unit, 1.0)
    print(json.dumps({"kernel": d["Kernel Name"][1], "dram_bytes_per_launch": dram, "ncu_ms": ms,
                      "source": os.path.basename(prefix) + "_ncu_summary.csv (ncu --set full, one launch)"}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
