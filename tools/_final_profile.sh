set -x
mkdir -p gpurun_out
# 1. ncu --set full captures of the two hot kernels (one launch each), then the DRAM traffic bench.py quotes
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cell_pipeline3 -s 7 -c 1 -f -o gpurun_out/r01_prof_cellpipe python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast -s 3 -c 1 -f -o gpurun_out/r01_prof_vfast python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 200 python tools/summarise_profiles.py r01 --traffic-only > gpurun_out/summarise.log 2>&1
# 2. the bench line (never under a profiler)
timeout 400 python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/bench.err
# 3. launch list of the same command
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# 4. secondary kernel: median filter
timeout 300 ncu --set full --clock-control none --import-source on -k regex:median_filter_select -c 1 -f -o gpurun_out/r01_prof_medfilt python tools/bench_extra.py > /dev/null 2>&1
timeout 100 python tools/pcie_probe.py > gpurun_out/r01_pcie_probe.json 2>&1
ls -la gpurun_out | tail -12
