"""Cell sharding over 2 / 4 / 8 ranks gives bit-identical results to one rank (tools/check_multigpu.py): smooth block with
its all-gathered reference chunk sums, per-cell HMM, i3 mu / sigma, state consensus, and the median filter with subclusters
kept whole and the reference groups' 4-cell halos exchanged.  On a one-GPU box the ranks share cuda:0 and the collectives go
through gloo / host memory (Engine._all_gather) - the arithmetic and its order are the same as over NCCL; with >= world GPUs
the same script runs over NCCL (tools/gpu_call_multi.sh)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_is_bitwise_equal_to_one_rank(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, ICNV_DIST_BACKEND="gloo", ICNV_ONE_GPU="1", ICNV_CHECK_GENES="2600", ICNV_CHECK_CELLS="700",
               OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_multigpu.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    tail = r.stdout[-2500:] + r.stderr[-2500:]
    assert r.returncode == 0 and "BITWISE EQUAL" in r.stdout, tail
