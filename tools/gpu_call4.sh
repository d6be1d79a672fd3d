# One GPU session: ncu capture (full set, source) of the default cell pipeline kernel (pass 2) on c2
tag=${1:-r02}
set -x
mkdir -p gpurun_out
B="python bench.py --config c2 --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_pipeline -s 8 -c 1 -f -o gpurun_out/${tag}_prof_cellpipe_c2 $B > /dev/null 2>&1
ls -la gpurun_out | tail -3
