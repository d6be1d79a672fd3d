#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multigpu.py : the cell-sharded run must be BITWISE equal to the single-GPU run
(rank 0 recomputes the whole run alone and compares): smooth block (all-gathered reference chunk sums), per-cell i6 HMM,
i3 mu / sigma, group consensus of the states (all-reduced counts) and the median filter (subclusters whole on a rank,
reference groups cut over ranks with their 4-cell halos exchanged).

    ICNV_DIST_BACKEND=gloo ICNV_ONE_GPU=1   every rank on cuda:0, collectives through host memory - how the GPU test suite
                                            runs it on a one-GPU box (tests/test_gpu_multirank.py); default: NCCL, one GPU per rank
    ICNV_CHECK_GENES / ICNV_CHECK_CELLS     problem size (genes, cells per rank); default 10000 / 1500
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from infercnv_b200 import dist as shard  # noqa: E402
from infercnv_b200.device import Engine  # noqa: E402
from infercnv_b200.hmm import CNV_LEVELS, get_HMM  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
backend = os.environ.get("ICNV_DIST_BACKEND", "nccl")
if os.environ.get("ICNV_ONE_GPU", "0") == "1":
    local = 0
torch.cuda.set_device(local)
if backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group(backend)
eng = Engine(local)
G = int(os.environ.get("ICNV_CHECK_GENES", "10000"))
C_total = int(os.environ.get("ICNV_CHECK_CELLS", "1500")) * world + 37
cs, cl = bench.chr_layout(G)
refs = bench.ref_groups_global(C_total)
atoms = bench.subclusters_global(C_total, 7)
Pi, delta, _, _ = get_HMM({k: {"mean": m, "sd": s} for k, m, s in zip(CNV_LEVELS, bench.I6_MEAN, bench.I6_SD)}, 1e-6)
plan = shard.plan_shards(C_total, refs, world, other_atoms=atoms)[rank]
X = eng.synth(G, cs, cl, plan.local_cells, C_total, bench.SEED)
C_local = X.shape[0]
ref_local = plan.local_ref_groups()
n_ref_local = int(sum(len(g) for g in ref_local))
Yext = torch.empty((C_local + 8 * len(ref_local), G), dtype=torch.float64, device=X.device)
Y = Yext[:C_local]
_, f = eng.smooth_block(X, cs, cl, ref_local, plan.ref_sizes, plan.max_chunks, out=Y)
S, f2 = eng.viterbi(Y, cs, cl, Pi, delta, bench.I6_MEAN, bench.I6_SD)
mu_d, sg_d = eng.mean_sd(Y, ref_local)        # i3 parameters across ranks
# CNV region consensus of two "sample" groups whose cells are spread over all ranks: integer counts, all-reduced
obs_global = [np.arange(int(0.10 * C_total), int(0.55 * C_total)), np.arange(int(0.55 * C_total), C_total)]
pos_of = {int(c): i for i, c in enumerate(plan.local_cells)}
obs_local = [np.array([pos_of[int(c)] for c in g if int(c) in pos_of], dtype=np.int32) for g in obs_global]
cons_d = eng.state_consensus(S, obs_local)
# median filter: this rank's whole subclusters + its slices of the reference groups (halos from the neighbours)
oc = plan.other_cells
sub_local = [np.arange(n_ref_local + int(np.searchsorted(oc, a[0])), n_ref_local + int(np.searchsorted(oc, a[0])) + len(a), dtype=np.int32)
             for a in atoms if len(oc) and oc[0] <= a[0] <= oc[-1]]
Fext = eng.median_filter_sharded(Yext, C_local, sub_local, ref_local, cs, cl, 7)
F = Fext[:C_local]
# group-mode HMM ("samples"): the reference groups and the observation cells as three groups, each cut over the ranks at chunk
# boundaries; rowMeans from all-gathered chunk sums, the traces scattered to the local cells
all_groups = refs + [np.arange(int(0.10 * C_total), C_total)]
gplan = shard.plan_shards(C_total, all_groups, world)[rank]
Xg = eng.synth(G, cs, cl, gplan.local_cells, C_total, bench.SEED)
Yg, _ = eng.smooth_block(Xg, cs, cl, gplan.local_ref_groups()[:2], gplan.ref_sizes[:2], gplan.max_chunks[:2])
sds_g = np.concatenate([bench.I6_SD * len(g) ** -0.5 for g in all_groups])
Sg, _ = eng.viterbi_groups(Yg, cs, cl, Pi, delta, bench.I6_MEAN, sds_g, gplan.local_ref_groups(), gplan.ref_sizes, gplan.max_chunks)
torch.cuda.synchronize()
assert int(f.item()) == 0 and int(f2.item()) == 0
print(f"[check_multigpu] rank {rank}: sharded run done ({C_local} local cells)", file=sys.stderr, flush=True)


def gather_rows(t):
    """every rank's rows on every rank (variable sizes: padded), through host memory when the backend is not NCCL"""
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device if backend == "nccl" else "cpu")
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    nmax = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((nmax, t.shape[1]), dtype=t.dtype, device=t.device if backend == "nccl" else "cpu")
    pad[: t.shape[0]] = t if backend == "nccl" else t.cpu()
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return [o.to(t.device) for o in outs]


Ys, Ss, Fs, Gs = gather_rows(Y), gather_rows(S), gather_rows(F), gather_rows(Sg)
ok = True
if rank == 0:
    plans = shard.plan_shards(C_total, refs, world, other_atoms=atoms)
    p1 = shard.plan_shards(C_total, refs, 1)[0]
    X1 = eng.synth(G, cs, cl, p1.local_cells, C_total, bench.SEED)
    eng.collective = False   # rank 0 alone: the other ranks are not in this computation
    Y1, _ = eng.smooth_block(X1, cs, cl, p1.local_ref_groups(), p1.ref_sizes, [(len(g) + 31) // 32 for g in refs])
    S1, _ = eng.viterbi(Y1, cs, cl, Pi, delta, bench.I6_MEAN, bench.I6_SD)
    mu_1, sg_1 = eng.mean_sd(Y1, p1.local_ref_groups())
    pos1 = {int(c): i for i, c in enumerate(p1.local_cells)}
    lists1 = [np.array([pos1[int(c)] for c in a], dtype=np.int32) for a in atoms] + [np.array([pos1[int(c)] for c in g], dtype=np.int32) for g in refs]
    F1 = eng.median_filter(Y1, cs, cl, lists1, 7)
    torch.cuda.synchronize()
    bad_y = bad_s = bad_f = 0
    for r, p in enumerate(plans):
        idx = torch.tensor([pos1[int(c)] for c in p.local_cells], device=X.device)
        bad_y += int((Ys[r][: len(idx)] != Y1[idx]).sum().item())
        bad_s += int((Ss[r][: len(idx)] != S1[idx]).sum().item())
        bad_f += int((Fs[r][: len(idx)] != F1[idx]).sum().item())
    cons_1 = eng.state_consensus(S1, [np.array([pos1[int(c)] for c in g], dtype=np.int32) for g in obs_global])
    bad_c = int((cons_1 != cons_d).sum().item())
    print(f"[check_multigpu] group consensus states differing from the 1-GPU run: {bad_c}")
    print(f"[check_multigpu] median filter ({len(atoms)} subclusters whole per rank, {len(refs)} reference groups cut over ranks with halos): "
          f"values differing from the 1-GPU run: {bad_f}")
    g1 = shard.plan_shards(C_total, all_groups, 1)[0]
    Xg1 = eng.synth(G, cs, cl, g1.local_cells, C_total, bench.SEED)
    Yg1, _ = eng.smooth_block(Xg1, cs, cl, g1.local_ref_groups()[:2])
    Sg1, _ = eng.viterbi_groups(Yg1, cs, cl, Pi, delta, bench.I6_MEAN, sds_g, g1.local_ref_groups())
    posg = {int(c): i for i, c in enumerate(g1.local_cells)}
    bad_g = 0
    for r, p in enumerate(shard.plan_shards(C_total, all_groups, world)):
        idx = torch.tensor([posg[int(c)] for c in p.local_cells], device=X.device)
        bad_g += int((Gs[r][: len(idx)] != Sg1[idx]).sum().item())
    print(f"[check_multigpu] group-mode HMM (3 groups cut over the ranks): states differing from the 1-GPU run: {bad_g}")
    ok = bad_y == 0 and bad_s == 0 and bad_f == 0 and bad_g == 0 and mu_d == mu_1 and sg_d == sg_1 and bad_c == 0
    print(f"[check_multigpu] i3 mu/sigma over the reference cells: {mu_d!r}, {sg_d!r} (1-GPU: {mu_1!r}, {sg_1!r})")
    print(f"[check_multigpu] world={world} backend={backend} cells={C_total} genes={G}: smoothed values differing from 1-GPU run: {bad_y}; "
          f"states differing: {bad_s}  -> {'BITWISE EQUAL' if ok else 'MISMATCH'}")
flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=X.device if backend == "nccl" else "cpu")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
