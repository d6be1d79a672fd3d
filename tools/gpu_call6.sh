# One GPU session: parity tests, c2 / c3 bench lines, ncu capture of the Viterbi kernel on c3 (DRAM traffic of the backpointer ring)
tag=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --config c2 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c2.json 2>/dev/null; python tools/bench_summary.py gpurun_out/${tag}_bench_c2.json
timeout 900 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c3.json 2>/dev/null; python tools/bench_summary.py gpurun_out/${tag}_bench_c3.json
B="python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast -s 3 -c 1 -f -o gpurun_out/${tag}_prof_vfast $B > /dev/null 2>&1
