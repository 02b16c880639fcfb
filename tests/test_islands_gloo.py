"""Multi-GPU host logic on CPU: island partition across ranks, checked with a world_size-2 gloo group
(one process per rank, rendezvous on 127.0.0.1)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from phyx_b200 import islands, scenes


def test_islands_of_multi_island_scene():
    sc = scenes.make("islands_64x20")
    ids, count = islands.find_islands(sc)
    assert count == 64 and ids[0] == -1                      # ground is static, 64 separate pyramids
    per = np.bincount(ids[ids >= 0])
    assert np.all(per == per[0]) and per[0] == 20 * 21 // 2
    one, n1 = islands.find_islands(scenes.make("pyramid_1k"))
    assert n1 == 1                                           # a single pyramid is one island


def test_partition_is_a_disjoint_cover():
    sc = scenes.make("islands_64x20")
    for ws in (1, 2, 4, 8):
        parts = islands.partition(sc, ws)
        dyn = [p[sc[p, 5] == 0] for p in parts]
        allb = np.concatenate(dyn)
        assert np.unique(allb).size == allb.size == int((sc[:, 5] == 0).sum())
        assert all(0 in p for p in parts)                    # ground replicated
        counts = [d.size for d in dyn]
        assert max(counts) - min(counts) <= 210              # balanced to within one island
        # contiguous in x: rank r's bodies lie left of rank r+1's
        for a, b in zip(dyn[:-1], dyn[1:]):
            assert sc[a, 0].max() < sc[b, 0].min()


def _worker(rank, world_size, port, out):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    sc = scenes.make("islands_64x20")
    mine, idx = islands.rank_scene(sc, rank, world_size)
    dyn = idx[sc[idx, 5] == 0]
    # every rank reports how many dynamic bodies it owns and a checksum of their indices
    t = torch.tensor([dyn.size, int(dyn.sum())], dtype=torch.int64)
    gathered = [torch.zeros_like(t) for _ in range(world_size)]
    dist.all_gather(gathered, t)
    total = torch.stack(gathered).sum(dim=0)
    # the bench's timing reduction: max over ranks
    ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        np.save(out, np.array([int(total[0]), int(total[1]), int(ms.item())]))
    dist.destroy_process_group()


def test_partition_agrees_across_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "total.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    total = np.load(out)
    sc = scenes.make("islands_64x20")
    dyn = np.nonzero(sc[:, 5] == 0)[0]
    assert total[0] == dyn.size and total[1] == int(dyn.sum())   # disjoint cover, agreed by both ranks
    assert total[2] == 2                                          # max-over-ranks reduction
