# First GPU call of round 2 (bash tools/_r02_first_call.sh under gpurun, one GPU): everything written after the
# round-1 GPU budget ran out gets its first run here, then the standing profile set is refreshed.
set -x
mkdir -p gpurun_out
# 1. parity: the whole GPU suite (new files: test_gpu_widen_regions.py, test_gpu_widen_ingest.py, test_gpu_widen_denoise.py)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log
# 2. memory safety of the new kernels on small cases
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_widen_regions.py tests/test_gpu_widen_ingest.py tests/test_gpu_widen_denoise.py -m gpu -x -q -k "not full_size" > gpurun_out/r02_memcheck.log 2>&1; tail -5 gpurun_out/r02_memcheck.log
# 3. the bench line (never under a profiler) and the secondary kernels incl. K7-K9
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 600 gpurun_out/r02_bench.json
timeout 300 python tools/bench_extra.py > gpurun_out/r02_secondary_kernels.json 2> gpurun_out/r02_secondary.err; cat gpurun_out/r02_secondary_kernels.json
# 3b. A/B of the reference-column reuse in pass 2 (default on since the end of round 1, never timed)
ICNV_REF_REUSE=0 timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_no_ref_reuse.json 2>/dev/null; tail -c 300 gpurun_out/r02_bench_no_ref_reuse.json
# 4. launch list of the bench command
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# 5. ncu --set full of the two hot kernels and the region kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cell_pipeline3 -s 7 -c 1 -f -o gpurun_out/r02_prof_cellpipe python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast -s 3 -c 1 -f -o gpurun_out/r02_prof_vfast python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"state_counts|region_count|region_emit|region_finish" -c 6 -f -o gpurun_out/r02_prof_regions python tools/bench_extra.py > /dev/null 2>&1
ls -la gpurun_out | tail -15
