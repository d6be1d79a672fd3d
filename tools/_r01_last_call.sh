# Last GPU call of round 1 (1.5 GPU-minutes left): the widened rows' parity tests that need no torch import, then
# the K7-K9 timings if any time remains.  Every line writes its own log so a cut-off call still brings results back.
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_widen_ingest.py tests/test_gpu_widen_denoise.py tests/test_gpu_widen_elementwise.py tests/test_gpu_widen_hmm_per_chr.py -m gpu -x -q -k "not full_size" > gpurun_out/r01b_widen_pytest.log 2>&1; tail -3 gpurun_out/r01b_widen_pytest.log
timeout 60 python -m pytest tests/test_gpu_widen_regions.py -m gpu -x -q -k "not full_size" > gpurun_out/r01b_widen_regions_pytest.log 2>&1; tail -3 gpurun_out/r01b_widen_regions_pytest.log
timeout 120 python tools/bench_extra.py > gpurun_out/r01b_secondary_kernels.json 2> gpurun_out/r01b_secondary.err; cat gpurun_out/r01b_secondary_kernels.json
