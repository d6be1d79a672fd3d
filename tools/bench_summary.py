#!/usr/bin/env python
"""One line per bench JSON file: python tools/bench_summary.py gpurun_out/r02d_bench_*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "no json:", e)
        continue
    if d.get("impl") == "reference":
        print(f, "REF value %.3e" % d["value"], d["cpu_baseline"])
        continue
    st = d["stage_ms"]
    print(f, "value %.3e step %.3f ms | smooth %.3f hmm %.3f mf %s | pass2 %.3f ms frac %.3f | hmm frac %.3f | reruns %s second pass %s" % (
        d["value"], d["ms_per_step"], st["smooth_block"], st["hmm"], st["median_filter"], d["roofline_cell_pipeline"]["ms_per_launch"],
        d["roofline_cell_pipeline"]["frac"], d["roofline_hmm"]["frac"], d["roofline_hmm"].get("sequences_rerun_in_reference_order_arithmetic"), d["roofline_hmm"].get("sequences_second_pass_fp64")))
    e = d.get("e2e")
    if e:
        print("     e2e %.3e  %.1f ms  (%s)" % (e["value"], e["ms_per_step"], e.get("host_memory")))
    if "cpu_baseline" in d:
        print("     cpu %.3e on %s threads" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
    if "roofline_median_filter" in d:
        print("     median filter %.2f ms frac %.4f" % (d["roofline_median_filter"]["ms_per_launch"], d["roofline_median_filter"]["frac"]))
