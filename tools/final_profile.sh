# The standing measurement set of a round (bash tools/final_profile.sh <tag> under gpurun, one GPU, ~6 min): the bench line of
# every configuration, the CPU arm, the launch list of the default bench command, and one `ncu --set full` capture of each
# hot kernel on the DEFAULT configuration (c3).  tools/profile_digest.py turns the captures into profiles/<tag>_*.
tag=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; tail -c 400 gpurun_out/${tag}_bench_c3.json; tail -3 gpurun_out/${tag}_bench_c3.err
timeout 600 python bench.py --config c2 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err; tail -3 gpurun_out/${tag}_bench_c2.err
timeout 900 python bench.py --config c4 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; tail -3 gpurun_out/${tag}_bench_c4.err
timeout 900 python bench.py --config c5 --no-e2e > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; tail -3 gpurun_out/${tag}_bench_c5.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_reference_c3.json 2>&1
timeout 300 python tools/bench_extra.py > gpurun_out/${tag}_secondary_kernels.json 2> gpurun_out/${tag}_secondary.err
B="python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c3.csv $B > /dev/null 2>&1
# per step of c3: cell pipeline launches = pass 1 (reference cells), stage D of the reference cells, pass 2 -> the 9th is a pass 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_pipeline -s 8 -c 1 -f -o gpurun_out/${tag}_prof_cellpipe $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast -s 3 -c 1 -f -o gpurun_out/${tag}_prof_vfast $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:median_filter_merge -c 1 -f -o gpurun_out/${tag}_prof_medfilt python tools/bench_extra.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:median_filter_merge -s 3 -c 1 -f -o gpurun_out/${tag}_prof_medfilt_c4 python bench.py --config c4 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairwise_dist -c 1 -f -o gpurun_out/${tag}_prof_dist python tools/bench_extra.py > /dev/null 2>&1
timeout 600 python tools/scaling_probe.py > gpurun_out/${tag}_scaling_probe.txt 2>&1
B5="python bench.py --config c5 --cells 8000 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_pipeline4 -s 8 -c 1 -f -o gpurun_out/${tag}_prof_cellpipe4_c5 $B5 > /dev/null 2>&1
ls -la gpurun_out | tail -14
