#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multigpu.py : the sharded smooth block + HMM must be
BITWISE equal to the single-GPU result (rank 0 recomputes the whole run alone and compares)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from infercnv_b200 import dist as shard  # noqa: E402
from infercnv_b200.device import Engine  # noqa: E402
from infercnv_b200.ops import CNV_LEVELS, get_HMM  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(local)
G, C_total = 10000, 1500 * world + 37
cs, cl = bench.chr_layout(G)
refs = bench.ref_groups_global(C_total)
Pi, delta, _, _ = get_HMM({k: {"mean": m, "sd": s} for k, m, s in zip(CNV_LEVELS, bench.I6_MEAN, bench.I6_SD)}, 1e-6)
plan = shard.plan_shards(C_total, refs, world)[rank]
X = eng.synth(G, cs, cl, plan.local_cells, C_total, bench.SEED)
Y, f = eng.smooth_block(X, cs, cl, plan.local_ref_groups(), plan.ref_sizes, plan.max_chunks)
S, f2 = eng.viterbi(Y, cs, cl, Pi, delta, bench.I6_MEAN, bench.I6_SD)
mu_d, sg_d = eng.mean_sd(Y, plan.local_ref_groups())        # i3 parameters across ranks
# CNV region consensus of two "sample" groups whose cells are spread over all ranks: integer counts, all-reduced
obs_global = [np.arange(int(0.10 * C_total), int(0.55 * C_total)), np.arange(int(0.55 * C_total), C_total)]
pos_of = {int(c): i for i, c in enumerate(plan.local_cells)}
obs_local = [np.array([pos_of[int(c)] for c in g if int(c) in pos_of], dtype=np.int32) for g in obs_global]
cons_d = eng.state_consensus(S, obs_local)
torch.cuda.synchronize()
assert int(f.item()) == 0 and int(f2.item()) == 0
# gather every rank's rows on rank 0 (variable sizes -> pad)
n_local = torch.tensor([X.shape[0]], device=X.device)
sizes = [torch.zeros_like(n_local) for _ in range(world)]
dist.all_gather(sizes, n_local)
nmax = int(max(s.item() for s in sizes))
Yp = torch.zeros((nmax, G), dtype=torch.float64, device=X.device); Yp[: X.shape[0]] = Y
Sp = torch.zeros((nmax, G), dtype=torch.uint8, device=X.device); Sp[: X.shape[0]] = S
Ys = [torch.zeros_like(Yp) for _ in range(world)]
Ss = [torch.zeros_like(Sp) for _ in range(world)]
dist.all_gather(Ys, Yp)
dist.all_gather(Ss, Sp)
ok = True
if rank == 0:
    plans = shard.plan_shards(C_total, refs, world)
    p1 = shard.plan_shards(C_total, refs, 1)[0]
    X1 = eng.synth(G, cs, cl, p1.local_cells, C_total, bench.SEED)
    eng.collective = False   # rank 0 alone: the other ranks are not in this computation
    Y1, _ = eng.smooth_block(X1, cs, cl, p1.local_ref_groups(), p1.ref_sizes, [(len(g) + 31) // 32 for g in refs])
    S1, _ = eng.viterbi(Y1, cs, cl, Pi, delta, bench.I6_MEAN, bench.I6_SD)
    mu_1, sg_1 = eng.mean_sd(Y1, p1.local_ref_groups())
    torch.cuda.synchronize()
    pos1 = {int(c): i for i, c in enumerate(p1.local_cells)}
    bad_y = bad_s = 0
    for r, p in enumerate(plans):
        idx = torch.tensor([pos1[int(c)] for c in p.local_cells], device=X.device)
        bad_y += int((Ys[r][: len(idx)] != Y1[idx]).sum().item())
        bad_s += int((Ss[r][: len(idx)] != S1[idx]).sum().item())
    cons_1 = eng.state_consensus(S1, [np.array([pos1[int(c)] for c in g], dtype=np.int32) for g in obs_global])
    bad_c = int((cons_1 != cons_d).sum().item())
    print(f"[check_multigpu] group consensus states differing from the 1-GPU run: {bad_c}")
    ok = bad_y == 0 and bad_s == 0 and mu_d == mu_1 and sg_d == sg_1 and bad_c == 0
    print(f"[check_multigpu] i3 mu/sigma over the reference cells: {mu_d!r}, {sg_d!r} (1-GPU: {mu_1!r}, {sg_1!r})")
    print(f"[check_multigpu] world={world} cells={C_total}: smoothed values differing from 1-GPU run: {bad_y}; "
          f"states differing: {bad_s}  -> {'BITWISE EQUAL' if ok else 'MISMATCH'}")
# ---- configs[3] shape: median filter over tumour subclusters, every index list whole on one rank (plan_list_shards) ----
rng = np.random.default_rng(7)
lists, pos = [], 0
while pos < C_total:
    n = int(rng.integers(50, 501))
    lists.append(np.arange(pos, min(C_total, pos + n)))
    pos += n
lplan = shard.plan_list_shards(lists, world)[rank]
Xl = eng.synth(G, cs, cl, lplan.cells, C_total, bench.SEED)
eng.collective = False            # the smooth block below is rank-local on purpose (its own reference cells): no collective
Yl, _ = eng.smooth_block(Xl, cs, cl, [np.arange(0, min(64, len(lplan.cells)))])
Fl = eng.median_filter(Yl, cs, cl, lplan.local_lists(lists), 7)
torch.cuda.synchronize()
ok_mf = True
if len(lplan.cells):
    # the same lists filtered one by one give the same bytes: a list's result depends on its own cells only
    k = lplan.list_ids[len(lplan.list_ids) // 2]
    loc = lplan.local_lists(lists)[lplan.list_ids.index(k)]
    one = eng.median_filter(Yl[torch.as_tensor(loc, device=Yl.device).long()].contiguous(), cs, cl, [np.arange(len(loc))], 7)
    ok_mf = bool(torch.equal(one, Fl[torch.as_tensor(loc, device=Yl.device).long()]))
flag = torch.tensor([1 if ok_mf else 0], device=X.device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"[check_multigpu] median filter over {len(lists)} subclusters sharded as whole lists: "
          f"{'per-list results independent of the sharding' if int(flag.item()) else 'MISMATCH'}")
ok = ok and bool(int(flag.item()))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
