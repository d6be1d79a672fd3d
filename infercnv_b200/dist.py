"""Cell sharding for multi-GPU runs: one process per GPU, cells partitioned, no data-path
collective except the all-gather of the per-gene partial sums behind the two reference-mean steps
(subtract_ref_expr_from_obs before and after smoothing, R/inferCNV_ops.R:771, :952).

Determinism: group means are summed as fixed chunks of CHUNK consecutive list entries (K1 in
csrc/icnv_smooth.cu) and the chunks are combined in global list order.  The planner below cuts
every reference group at multiples of CHUNK, so the chunk contents and their order are the same
for 1, 2, 4 or 8 ranks and the means - hence the whole output - are bit-identical.

Pure host logic + torch.distributed plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CHUNK = 32  # must match dev_group_means() in csrc/icnv_api.cu


@dataclass
class ShardPlan:
    """What one rank owns.  All indices are GLOBAL cell indices."""
    rank: int
    world: int
    ref_slices: list[np.ndarray]   # per reference group: this rank's part of the group's list
    other_cells: np.ndarray        # non-reference cells of this rank
    ref_sizes: list[int]           # global size of every reference group
    max_chunks: list[int]          # per group: max over ranks of the local chunk count (padding)

    @property
    def local_cells(self) -> np.ndarray:
        """Global indices of the local columns, reference slices first (in group order)."""
        parts = list(self.ref_slices) + [self.other_cells]
        return np.concatenate(parts) if parts else np.zeros(0, np.int64)

    def local_ref_groups(self) -> list[np.ndarray]:
        """Reference groups as LOCAL column indices (columns ordered as local_cells)."""
        out, pos = [], 0
        for s in self.ref_slices:
            out.append(np.arange(pos, pos + len(s), dtype=np.int32))
            pos += len(s)
        return out


def split_chunk_aligned(n: int, world: int, chunk: int = CHUNK) -> list[tuple[int, int]]:
    """Cut [0, n) into `world` contiguous ranges whose boundaries are multiples of `chunk`
    (the last range takes the ragged tail).  Ranges may be empty when n is small."""
    n_chunks = (n + chunk - 1) // chunk
    base, extra = divmod(n_chunks, world)
    out, pos = [], 0
    for r in range(world):
        c = base + (1 if r < extra else 0)
        lo = min(pos * chunk, n)
        hi = min((pos + c) * chunk, n)
        out.append((lo, hi))
        pos += c
    return out


def plan_shards(n_cells: int, ref_groups: list[np.ndarray], world: int, other_atoms=None) -> list[ShardPlan]:
    """Partition the cells of a run over `world` ranks.  other_atoms (optional): index lists that partition the
    non-reference cells and must each stay whole on one rank (tumour subclusters: a subcluster's cell order is the
    median filter's window, R/noise_reduction.R:60-75); consecutive atoms go to consecutive ranks."""
    ref_groups = [np.asarray(g, dtype=np.int64) for g in ref_groups]
    is_ref = np.zeros(n_cells, dtype=bool)
    for g in ref_groups:
        is_ref[g] = True
    others = np.flatnonzero(~is_ref)
    cuts = [split_chunk_aligned(len(g), world) for g in ref_groups]
    # the non-reference cells even out what the chunk-aligned reference cuts left uneven
    ref_count = [sum(c[r][1] - c[r][0] for c in cuts) for r in range(world)]
    base, extra = divmod(n_cells, world)
    want = [max(0, base + (1 if r < extra else 0) - ref_count[r]) for r in range(world)]
    scale_fix = len(others) - sum(want)
    r = 0
    while scale_fix != 0:                      # rounding leftovers (also when a rank is reference-only)
        step = 1 if scale_fix > 0 else -1
        if want[r % world] + step >= 0:
            want[r % world] += step
            scale_fix -= step
        r += 1
    if other_atoms is None:
        bounds = np.concatenate([[0], np.cumsum(want)])
        other_parts = [others[bounds[r]:bounds[r + 1]] for r in range(world)]
    else:
        atoms = [np.asarray(a, dtype=np.int64) for a in other_atoms]
        if not np.array_equal(np.sort(np.concatenate(atoms)) if atoms else np.zeros(0, np.int64), others):
            raise ValueError("other_atoms must partition the non-reference cells")
        other_parts, a = [], 0
        target = np.cumsum(want)
        done = 0
        for r in range(world):
            mine = []
            while a < len(atoms) and (r == world - 1 or abs(done + len(atoms[a]) - target[r]) <= abs(done - target[r])):
                mine.append(atoms[a])
                done += len(atoms[a])
                a += 1
            other_parts.append(np.concatenate(mine) if mine else np.zeros(0, np.int64))
    max_chunks = [max((hi - lo + CHUNK - 1) // CHUNK for lo, hi in c) for c in cuts]
    plans = []
    for r in range(world):
        plans.append(ShardPlan(
            rank=r, world=world,
            ref_slices=[g[cuts[k][r][0]:cuts[k][r][1]] for k, g in enumerate(ref_groups)],
            other_cells=other_parts[r],
            ref_sizes=[len(g) for g in ref_groups],
            max_chunks=max_chunks))
    return plans


@dataclass
class ListShardPlan:
    """Whole index lists (tumour subclusters, reference groups) owned by one rank: the partition for the steps whose
    unit is a list in its own cell order - apply_median_filtering's blocks (R/noise_reduction.R:43-89) and the
    group-mode HMM's rowMeans (R/inferCNV_HMM.R:345-408, 509-567)."""
    rank: int
    world: int
    list_ids: list[int]            # which of the caller's lists this rank owns, in the caller's order
    cells: np.ndarray              # GLOBAL cell indices of the local columns: the owned lists back to back

    def local_lists(self, lists) -> list[np.ndarray]:
        """The owned lists as LOCAL column indices (columns ordered as `cells`)."""
        out, pos = [], 0
        for k in self.list_ids:
            n = len(lists[k])
            out.append(np.arange(pos, pos + n, dtype=np.int32))
            pos += n
        return out


def plan_list_shards(lists: list[np.ndarray], world: int) -> list[ListShardPlan]:
    """Assign every list WHOLE to one rank (a list is never split: its cells' order is part of the median filter's
    window and of nothing else, so no halo exchange is needed), largest first onto the least loaded rank; ties go to
    the lower rank, so the plan is a pure function of the list sizes.  Cells in no list are in no shard (the median
    filter copies them, the group HMM leaves them at -1: the caller keeps them)."""
    lists = [np.asarray(v, dtype=np.int64) for v in lists]
    seen = np.concatenate(lists) if lists else np.zeros(0, np.int64)
    if len(np.unique(seen)) != len(seen):
        raise ValueError("index lists overlap: a cell would live on two ranks")
    load = [0] * world
    owner = [0] * len(lists)
    for k in sorted(range(len(lists)), key=lambda k: (-len(lists[k]), k)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[k] = r
        load[r] += len(lists[k])
    plans = []
    for r in range(world):
        ids = [k for k in range(len(lists)) if owner[k] == r]
        cells = np.concatenate([lists[k] for k in ids]) if ids else np.zeros(0, np.int64)
        plans.append(ListShardPlan(rank=r, world=world, list_ids=ids, cells=cells))
    return plans


def halo_sources(slice_lens, rank: int, r: int):
    """A list whose entries are spread over ranks as contiguous slices (slice_lens[q] entries on rank q, list order =
    rank order): which entries does `rank` need from its neighbours so that every window of 2r + 1 consecutive list
    entries around its own entries is complete?  Returns (prev, nxt): lists of (src_rank, k) meaning "the last k entries
    of src_rank's slice" (prev, in list order) and "the first k entries" (nxt).  Ranks holding fewer than r entries are
    walked through.  A rank with an empty slice needs nothing."""
    if slice_lens[rank] == 0:
        return [], []
    prev, need = [], r
    q = rank - 1
    while need > 0 and q >= 0:
        k = min(need, int(slice_lens[q]))
        if k:
            prev.insert(0, (q, k))
            need -= k
        q -= 1
    nxt, need = [], r
    q = rank + 1
    while need > 0 and q < len(slice_lens):
        k = min(need, int(slice_lens[q]))
        if k:
            nxt.append((q, k))
            need -= k
        q += 1
    return prev, nxt


def allgather_partials(local, max_chunks: int):
    """All-gather one group's partial sums.  `local` is a torch tensor (n_local_chunks, G) on the
    device of the process group's backend.  Returns (world * max_chunks, G): rank-major, each
    rank's block zero-padded to `max_chunks` rows (adding 0.0 chunks does not change a sum), or
    `local` itself when no process group is initialised (single GPU)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    G = local.shape[1]
    padded = torch.zeros((max_chunks, G), dtype=local.dtype, device=local.device)
    if local.shape[0]:
        padded[: local.shape[0]] = local
    out = torch.empty((world * max_chunks, G), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded) if hasattr(dist, "all_gather_into_tensor") and local.is_cuda else \
        _allgather_list(out, padded, world)
    return out


def _allgather_list(out, padded, world):
    import torch
    import torch.distributed as dist

    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out.copy_(torch.cat(parts, dim=0))
