# First GPU call of round 2 (bash tools/_r02_first_call.sh under gpurun, one GPU, ~8 min): everything written after the
# round-1 GPU budget ran out gets its first timed run here (it was verified bit-for-bit under the host emulation only),
# each change behind its own switch so the A/B is one environment variable, then the standing profile set is refreshed.
set -x
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline"
# 1. parity: the whole GPU suite (new since the last full GPU run: padded-Q identity, grouped element-wise slow paths)
ICNV_TEST_PIPELINED_HOST=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log
# 2. the bench line of the default path (never under a profiler)
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 600 gpurun_out/r02_bench.json
# 3. A/B of what changed since profiles/r01_bench.json (device-resident value only; every line is one JSON record)
#    a. cell pipeline: padded Q layout off (the measured ping-pong layout + the new stage A / D groups)
ICNV_CELL_PADQ=0 timeout 300 $B > gpurun_out/r02_ab_padq0.json 2>/dev/null
#    a2. counted scan loops instead of the fully unrolled slices
ICNV_CELL_LFIX=0 timeout 300 $B > gpurun_out/r02_ab_lfix0.json 2>/dev/null
#    a3. 512 threads per CTA (21 genes per thread, ~105 registers) instead of 1024 x 11
ICNV_CELL_NT=512 timeout 300 $B > gpurun_out/r02_ab_nt512.json 2>/dev/null
#    b. reference-column reuse in pass 2 off (default on since the end of round 1)
ICNV_REF_REUSE=0 timeout 300 $B > gpurun_out/r02_ab_refreuse0.json 2>/dev/null
#    c. fast Viterbi occupancy variants (default 16 warps per CTA at 128 registers)
ICNV_VFAST_WARPS=20 timeout 300 $B > gpurun_out/r02_ab_vfast20.json 2>/dev/null
ICNV_VFAST_WARPS=24 timeout 300 $B > gpurun_out/r02_ab_vfast24.json 2>/dev/null
#    d. host pipeline slab size (default 1024 cells; fill + drain of the PCIe pipeline is one slab each way)
for s in 1024 512 256; do ICNV_SLAB_CELLS=$s timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02_ab_slab$s.json 2>/dev/null; done
for f in gpurun_out/r02_ab_*.json; do echo "$f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('roofline',{}).get('ms_per_launch'), d.get('roofline_hmm',{}).get('ms_per_launch'), d.get('e2e',{}).get('ms_per_step'))")"; done
# 4. secondary kernels incl. K7-K9
timeout 300 python tools/bench_extra.py > gpurun_out/r02_secondary_kernels.json 2> gpurun_out/r02_secondary.err; cat gpurun_out/r02_secondary_kernels.json
# 5. launch list of the bench command
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $B --steps 2 --warmup 3 > /dev/null 2>&1
# 6. ncu --set full of the two hot kernels and the region kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cell_pipeline3 -s 7 -c 1 -f -o gpurun_out/r02_prof_cellpipe $B --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viterbi_fast -s 3 -c 1 -f -o gpurun_out/r02_prof_vfast $B --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"state_counts|region_count|region_emit|region_finish" -c 6 -f -o gpurun_out/r02_prof_regions python tools/bench_extra.py > /dev/null 2>&1
# 7. memory safety of the changed kernels on small cases
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "padded_q or slow_paths or golden or viterbi_cells" > gpurun_out/r02_memcheck.log 2>&1; tail -5 gpurun_out/r02_memcheck.log
ls -la gpurun_out | tail -20
# 8. (needs --gpus 2; run separately) N = 2: default and slab-pipelined e2e
#    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02_bench_n2.json
#    ICNV_BENCH_E2E_PIPELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02_bench_n2_pipelined.json
# 9. median filter: counting selection (default, now with four pivots per round) against the key network for full windows
ICNV_MF_KERNEL=2 timeout 300 python tools/bench_extra.py > gpurun_out/r02_secondary_kernels_mfnet.json 2>/dev/null; python -c "import json; a=json.load(open('gpurun_out/r02_secondary_kernels.json')); b=json.load(open('gpurun_out/r02_secondary_kernels_mfnet.json')); print('median filter ms: counting', a['median_filter_w7']['ms'], ' key network', b['median_filter_w7']['ms'])"
