"""Host logic of the partitioned solve on CPU: the numpy model of the partition plan, the static-body rewrite
of the sequential equivalent, and the IPC-handle exchange of attach_process_group on a world_size-2 gloo
group (one process per rank, rendezvous on 127.0.0.1) with a stand-in for the device context."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from phyx_b200 import partition
from phyx_b200 import types as T


def test_plan_model_cuts_balance_and_classify():
    # a chain of 12 dynamic rows, manifolds between neighbours, plus ground contacts (row -1)
    r1 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, -1, -1, 11])
    r2 = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0, 11, -1])
    cuts, cls, boundary = partition.plan_model_n(r1, r2, 2, 12)
    assert cuts[0] == 0 and cuts[2] == 12 and 0 < cuts[1] < 12
    # exactly one neighbour manifold straddles the cut, and its two rows are the boundary
    assert int((cls == 2).sum()) == 1
    k = int(np.nonzero(cls == 2)[0][0])
    assert set(boundary.tolist()) == {int(r1[k]), int(r2[k])} and r1[k] == cuts[1] - 1
    # ground contacts belong to the rank of their dynamic body
    assert cls[11] == 0 and cls[12] == 1 and cls[13] == 1
    # balanced: each rank is home to about half of the manifolds
    home = np.where(r1 < 0, r2, np.where(r2 < 0, r1, np.minimum(r1, r2)))
    left = int((home < cuts[1]).sum())
    assert abs(left - (len(r1) - left)) <= 2
    # four ranks: monotone cuts, every interior manifold inside its rank's rows
    cuts4, cls4, b4 = partition.plan_model_n(r1, r2, 4, 12)
    assert np.all(np.diff(cuts4) >= 0)
    for m in range(len(r1)):
        rows = [r for r in (r1[m], r2[m]) if r >= 0]
        if cls4[m] < 4:
            assert all(cuts4[cls4[m]] <= r < cuts4[cls4[m] + 1] for r in rows)
        else:
            assert len({int(np.searchsorted(cuts4[1:4], r, side="right")) for r in rows}) == 2


def test_plan_model_degenerate_inputs():
    cuts, cls, boundary = partition.plan_model_n(np.zeros(0, int), np.zeros(0, int), 3, 5)
    assert cuts.tolist() == [0, 0, 0, 5] and cls.size == 0 and boundary.size == 0
    # only static-static manifolds: class 0, no boundary
    cuts, cls, boundary = partition.plan_model_n(np.array([-1, -1]), np.array([-1, -1]), 2, 4)
    assert cls.tolist() == [0, 0] and boundary.size == 0


def test_sequential_equivalent_gives_every_rank_its_own_static_bodies():
    bodies = np.zeros(5, dtype=T.RIGID_BODY)
    bodies["invMass"][1:] = 1.0
    bodies["invInertia"][1:] = 1.0                      # body 0 is static
    joints = np.zeros(6, dtype=T.CONTACT_JOINT)
    joints["body1Index"] = [0, 1, 0, 3, 2, 0]
    joints["body2Index"] = [1, 2, 3, 4, 3, 4]
    # slots: rank 0 = joints 0,1 ; rank 1 = joints 2,3,5 ; cut = joint 4
    slots = np.array([0, -1, 1, -1, 2, -1, 3, -1, 5, -1, 4, -1], dtype=np.int32)
    cls_start = np.array([0, 4, 10, 12])
    ob, oj = partition.sequential_equivalent(bodies, joints, slots, cls_start, 2)
    assert ob.shape[0] == 6 and ob["invMass"][5] == 0
    assert oj["body1Index"].tolist() == [0, 1, 5, 3, 2, 5]   # rank 1's ground contacts use the copy
    assert oj["body2Index"].tolist() == joints["body2Index"].tolist()


class _FakeContext:
    """Stands in for capi.Context in the handshake: records what it was given."""

    def __init__(self):
        self.created, self.attached = None, None

    def partition_create(self, rank, ranks, boundary, bulk):
        self.created = (rank, ranks, boundary, bulk)
        return bytes([65 + rank]) * 64, 0

    def partition_attach(self, ranks, ipc_handles=None, local_pointers=None, peer_devices=None):
        self.attached = (ranks, list(ipc_handles))


def _worker(rank, world_size, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    ctx = _FakeContext()
    r, n = partition.attach_process_group(ctx, (128, 4096))
    ok = (r, n) == (rank, world_size) and ctx.created == (rank, world_size, 128, 4096)
    ok = ok and ctx.attached[0] == world_size and ctx.attached[1] == [bytes([65 + q]) * 64 for q in range(world_size)]
    # every rank derives the same plan from the same (replicated) inputs
    rng = np.random.default_rng(7)
    r1 = rng.integers(-1, 400, 3000)
    r2 = rng.integers(0, 400, 3000)
    cuts, cls, boundary = partition.plan_model_n(r1, r2, world_size, 400)
    digest = [int(cuts.sum()), int(cls.sum()), int(boundary.sum())]
    gathered = [None] * world_size
    dist.all_gather_object(gathered, digest)
    ok = ok and all(g == digest for g in gathered)
    dist.barrier()
    np.save(out + f".{rank}.npy", np.array([int(ok)]))
    dist.destroy_process_group()


def test_ipc_handshake_over_gloo(tmp_path):
    import multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "ok")
    ctxm = mp.get_context("spawn")
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(2):
        assert int(np.load(out + f".{r}.npy")[0]) == 1


def test_plan_model_properties_on_random_contact_graphs():
    """Whatever the contact graph: cuts are monotone and cover all rows, an interior manifold's dynamic rows lie inside
    its rank's range, a cut manifold's rows lie in two different ranges, the boundary is exactly the rows of the cut
    manifolds, and no rank is home to more than its share (+ the manifolds of one row) of the manifolds."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(2, 8), st.integers(1, 300), st.integers(0, 900), st.integers(0, 2**31 - 1))
    def check(ranks, nb, m, seed):
        rng = np.random.default_rng(seed)
        r1 = rng.integers(-1, nb, m)
        near = np.clip(r1 + rng.integers(-3, 4, m), -1, nb - 1)            # contacts are local in sorted-x order
        r2 = np.where(rng.random(m) < 0.8, near, rng.integers(-1, nb, m))
        cuts, cls, boundary = partition.plan_model_n(r1, r2, ranks, nb)
        assert cuts[0] == 0 and cuts[ranks] == nb and np.all(np.diff(cuts) >= 0)
        rank_of = lambda r: np.searchsorted(cuts[1:ranks], r, side="right")
        want_boundary = set()
        for k in range(m):
            rows = [int(r) for r in (r1[k], r2[k]) if r >= 0]
            if cls[k] < ranks:
                assert all(rank_of(r) == cls[k] for r in rows) or not rows
            else:
                assert len(rows) == 2 and rank_of(rows[0]) != rank_of(rows[1])
                want_boundary.update(rows)
        assert set(boundary.tolist()) == want_boundary
        home = np.where(r1 < 0, r2, np.where(r2 < 0, r1, np.minimum(r1, r2)))
        home = home[home >= 0]
        if home.size:
            per_row = np.bincount(home, minlength=nb).max()
            share = -(-home.size // ranks)
            for q in range(ranks):
                assert ((home >= cuts[q]) & (home < cuts[q + 1])).sum() <= share + per_row

    check()
