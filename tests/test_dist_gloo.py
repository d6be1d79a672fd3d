"""Host-side logic of the multi-GPU path on CPU: the shard planner and the all-gather of the
reference-mean partial sums (torch.distributed, gloo, world_size 2).  The per-chunk sums are
emulated with NumPy here (on the GPU they come from group_partial_sums_kernel); what is under test
is that cutting the reference groups at chunk boundaries + rank-major gathering + zero padding
reproduces the single-process summation order bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from infercnv_b200 import dist as shard


def _chunk_sums(X, cells):
    """partial[q] = sum over chunk q (CHUNK list entries, in order) - float64, sequential adds."""
    n = len(cells)
    nq = (n + shard.CHUNK - 1) // shard.CHUNK
    out = np.zeros((nq, X.shape[1]))
    for q in range(nq):
        s = np.zeros(X.shape[1])
        for c in cells[q * shard.CHUNK:(q + 1) * shard.CHUNK]:
            s = s + X[c]
        out[q] = s
    return out


def _combine(part, count):
    s = np.zeros(part.shape[1])
    for q in range(part.shape[0]):
        s = s + part[q]
    return s / count


def test_plan_covers_every_cell_once_and_cuts_refs_on_chunk_boundaries():
    C = 10007
    refs = [np.arange(0, 601), np.arange(650, 1049)]
    for world in (1, 2, 4, 8):
        plans = shard.plan_shards(C, refs, world)
        allc = np.concatenate([p.local_cells for p in plans])
        assert sorted(allc.tolist()) == list(range(C))
        for k, g in enumerate(refs):
            pos = 0
            for p in plans:
                sl = p.ref_slices[k]
                np.testing.assert_array_equal(sl, g[pos:pos + len(sl)])
                assert pos % shard.CHUNK == 0
                pos += len(sl)
            assert pos == len(g)
        sizes = [len(p.local_cells) for p in plans]
        assert max(sizes) - min(sizes) <= 1
        # local reference groups index the local column order
        p0 = plans[0]
        for k, loc in enumerate(p0.local_ref_groups()):
            np.testing.assert_array_equal(p0.local_cells[loc], p0.ref_slices[k])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    C, G = 1000, 37
    X = rng.normal(size=(C, G)) * 10.0 ** rng.integers(-3, 4, size=(C, 1))
    refs = [np.arange(0, 333), np.arange(400, 471)]
    plan = shard.plan_shards(C, refs, world)[rank]
    means = []
    for k, g in enumerate(refs):
        local = torch.from_numpy(_chunk_sums(X, plan.ref_slices[k]))
        gathered = shard.allgather_partials(local, plan.max_chunks[k])
        means.append(_combine(gathered.numpy(), plan.ref_sizes[k]))
    single = [_combine(_chunk_sums(X, g), len(g)) for g in refs]
    ok = all(np.array_equal(a, b) for a, b in zip(means, single))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allgather_of_partial_sums_is_bitwise_equal_to_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(0, True), (1, True)]


def test_list_shards_keep_every_list_whole_and_balance_the_load():
    rng = np.random.default_rng(0)
    C = 20000
    perm = rng.permutation(C)
    lists, pos = [], 0
    while pos < C - 600:
        n = int(rng.integers(50, 501))
        lists.append(perm[pos:pos + n])
        pos += n                                            # the last cells are in no list
    for world in (1, 2, 4, 8):
        plans = shard.plan_list_shards(lists, world)
        owned = sorted(k for p in plans for k in p.list_ids)
        assert owned == list(range(len(lists)))             # every list on exactly one rank
        for p in plans:
            assert p.list_ids == sorted(p.list_ids)
            np.testing.assert_array_equal(p.cells, np.concatenate([lists[k] for k in p.list_ids]))
            for k, loc in zip(p.list_ids, p.local_lists(lists)):
                np.testing.assert_array_equal(p.cells[loc], lists[k])      # local columns keep the list's own order
        loads = [len(p.cells) for p in plans]
        assert max(loads) - min(loads) <= 500               # within one list of each other
        assert shard.plan_list_shards(lists, world)[0].list_ids == plans[0].list_ids      # deterministic
    with pytest.raises(ValueError):
        shard.plan_list_shards([np.array([1, 2, 3]), np.array([3, 4])], 2)
