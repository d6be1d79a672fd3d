# compute-sanitizer passes over the parity tests of the kernels written this round (memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards).  bash tools/gpu_sanitize.sh <tag> under gpurun, one GPU.
tag=${1:-r02}
set -x
mkdir -p gpurun_out
K="pairwise or median_filter or single_precision or exact_ties or known_answers or padded_q"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/${tag}_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/${tag}_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pairwise or median_filter_shared or single_precision" > gpurun_out/${tag}_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/${tag}_racecheck.txt
