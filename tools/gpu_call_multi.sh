# Multi-GPU session (bash tools/gpu_call_multi.sh <tag> <N> under `gpurun --gpus N`): NCCL bitwise check, one process on N GPUs
# through the C ABI, and the strong-scaling bench lines of c3 (and c4 / c5 at N = 8).
tag=${1:-r02}; N=${2:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29501 tools/check_multigpu.py > gpurun_out/${tag}_check_multigpu_n$N.txt 2>&1; tail -8 gpurun_out/${tag}_check_multigpu_n$N.txt
timeout 600 python tools/check_init_devices.py > gpurun_out/${tag}_check_init_devices_n$N.txt 2>&1; tail -3 gpurun_out/${tag}_check_init_devices_n$N.txt
timeout 900 $TR --master-port 29502 bench.py --gpus $N > gpurun_out/${tag}_bench_c3_n$N.json 2> gpurun_out/${tag}_bench_c3_n$N.err; tail -c 300 gpurun_out/${tag}_bench_c3_n$N.json; tail -3 gpurun_out/${tag}_bench_c3_n$N.err
timeout 900 python bench.py --gpus 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_c3_n1_samebox.json 2> /dev/null
if [ "$N" = "8" ]; then
  timeout 900 $TR --master-port 29503 bench.py --gpus $N --config c4 > gpurun_out/${tag}_bench_c4_n$N.json 2> gpurun_out/${tag}_bench_c4_n$N.err; tail -3 gpurun_out/${tag}_bench_c4_n$N.err
  timeout 900 $TR --master-port 29504 bench.py --gpus $N --config c5 --no-e2e > gpurun_out/${tag}_bench_c5_n$N.json 2> gpurun_out/${tag}_bench_c5_n$N.err; tail -3 gpurun_out/${tag}_bench_c5_n$N.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29505 tools/check_multigpu.py > gpurun_out/${tag}_check_multigpu_n4.txt 2>&1; grep "check_multigpu. world" gpurun_out/${tag}_check_multigpu_n4.txt
  for n in 4; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-e2e > gpurun_out/${tag}_bench_c3_n$n.json 2> /dev/null; done
fi
ls -la gpurun_out | tail -8
