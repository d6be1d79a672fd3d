# One GPU session (bash tools/gpu_call3.sh <tag> under gpurun): parity tests, then the HMM passes A/B (FP32 first pass at 16 / 20 /
# 24 warps per SM against the FP64 pass) on c2 and c3, ncu capture of the FP32 pass
tag=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${tag}_pytest_gpu.log
for v in "2 16" "2 24" "1 16"; do set -- $v; ICNV_HMM_MODE=$1 ICNV_VFAST_WARPS=$2 timeout 600 python bench.py --config c2 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c2_mode$1_w$2.json 2>/dev/null; python tools/bench_summary.py gpurun_out/${tag}_bench_c2_mode$1_w$2.json; done
timeout 900 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; python tools/bench_summary.py gpurun_out/${tag}_bench_c3.json; tail -3 gpurun_out/${tag}_bench_c3.err
timeout 900 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; python tools/bench_summary.py gpurun_out/${tag}_bench_c4.json
timeout 300 python tools/bench_extra.py > gpurun_out/${tag}_secondary_kernels.json 2> gpurun_out/${tag}_secondary.err; cut -c1-200 gpurun_out/${tag}_secondary_kernels.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast32 -s 3 -c 1 -f -o gpurun_out/${tag}_prof_vfast32 python bench.py --config c2 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
timeout 600 python tools/scaling_probe.py > gpurun_out/${tag}_scaling_probe.txt 2>&1; tail -3 gpurun_out/${tag}_scaling_probe.txt
ls -la gpurun_out | tail -5
