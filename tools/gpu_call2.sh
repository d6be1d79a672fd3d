# One GPU session (bash tools/gpu_call2.sh <tag> under gpurun): parity tests, secondary kernels incl. the distance kernel,
# the fixed-cost probe of small shards, one ncu capture of the distance kernel
tag=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python tools/bench_extra.py > gpurun_out/${tag}_secondary_kernels.json 2> gpurun_out/${tag}_secondary.err; cut -c1-200 gpurun_out/${tag}_secondary_kernels.json; tail -3 gpurun_out/${tag}_secondary.err
timeout 600 python tools/scaling_probe.py > gpurun_out/${tag}_scaling_probe.txt 2>&1; tail -8 gpurun_out/${tag}_scaling_probe.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pairwise_dist -c 1 -f -o gpurun_out/${tag}_prof_dist python tools/bench_extra.py > /dev/null 2>&1
ls -la gpurun_out | tail -5
