set -x
mkdir -p gpurun_out
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/ref.err
timeout 300 python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cell_pipeline3 -s 7 -c 1 -o gpurun_out/r01_prof_cellpipe python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast -s 3 -c 1 -o gpurun_out/r01_prof_vfast python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -8
