# One GPU session (bash tools/gpu_call.sh <tag> under gpurun): GPU parity suite, bench lines, the CPU arm, the launch list
# and ncu --set full captures of the hot kernels.  Outputs land in gpurun_out/<tag>_*.
tag=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --config c2 --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err; tail -c 300 gpurun_out/${tag}_bench_c2.json; tail -5 gpurun_out/${tag}_bench_c2.err
ICNV_CELL_KERNEL=3 timeout 600 python bench.py --config c2 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c2_v3.json 2> gpurun_out/${tag}_bench_c2_v3.err
ICNV_CELL_NT=1024 timeout 600 python bench.py --config c2 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c2_nt1024.json 2> gpurun_out/${tag}_bench_c2_nt1024.err
timeout 900 python bench.py > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; tail -c 1500 gpurun_out/${tag}_bench_c3.json; tail -5 gpurun_out/${tag}_bench_c3.err
timeout 900 python bench.py --config c5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; tail -c 600 gpurun_out/${tag}_bench_c5.json; tail -5 gpurun_out/${tag}_bench_c5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_c3.json 2>&1; cat gpurun_out/${tag}_bench_reference_c3.json | cut -c1-300
B="python bench.py --config c2 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c2.csv $B > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cell_pipeline -s 8 -c 1 -f -o gpurun_out/${tag}_prof_cellpipe $B > /dev/null 2>&1
ls -la gpurun_out | tail -12
