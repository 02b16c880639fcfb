"""The multi-rank protocol of the partitioned solve on CPU: every rank is the oracle's literal
restatement (oracle pxo_rank_*: own copy of rows and accumulators, interior passes, boundary rows in and out, cut
passes), the transport is first a plain loop (3 ranks in one process), then torch.distributed on a world_size-2
gloo group (one process per rank, rendezvous on 127.0.0.1) — the same order of operations the devices and
phyx_b200/partition.py use, with numpy arrays instead of NVLink stores.

Claim checked, bit for bit: interior passes in parallel + exchange of the boundary rows + cut passes on every rank
+ OR-ed early-out == ONE sequential sweep over the class-major slot order with per-rank static bodies
(partition.sequential_equivalent), and all ranks end with the same state."""
import os
import socket

import numpy as np
import pytest

from conftest import assert_records_equal, golden
from phyx_b200 import partition

VEL_FIELDS = ("velocity", "angularVelocity", "displacingVelocity", "displacingAngularVelocity")
ITERS = (20, 20)


def build_problem(ranks, fixture="solve_pyramid_1k_s30.npz"):
    """Partition plan and class-major schedule for a golden solve input, in numpy: rows = bodies by min x, units =
    joints, colours by first fit inside each class."""
    from oracle import oraclepy

    g = golden(fixture)
    bodies, joints, cps = g["bodies"], g["joints"], g["contact_points"]
    n = bodies.shape[0]
    static = (bodies["invMass"] == 0) & (bodies["invInertia"] == 0)
    order = np.argsort(bodies["aabb_min"][:, 0], kind="stable")
    row_of = np.empty(n, np.int64)
    row_of[order] = np.arange(n)
    b1, b2 = joints["body1Index"].astype(np.int64), joints["body2Index"].astype(np.int64)
    r1, r2 = np.where(static[b1], -1, row_of[b1]), np.where(static[b2], -1, row_of[b2])
    cuts, cls, boundary_rows = partition.plan_model_n(r1, r2, ranks, n)
    owner = np.where(static, -1, np.searchsorted(cuts[1:ranks], row_of, side="right"))
    # first fit per class over the joints in index order
    used = {}
    colour = np.zeros(joints.shape[0], np.int64)
    for j in range(joints.shape[0]):
        bs = [b for b in (int(b1[j]), int(b2[j])) if not static[b]]
        taken = set().union(*[used.setdefault((int(cls[j]), b), set()) for b in bs]) if bs else set()
        c = 0
        while c in taken:
            c += 1
        colour[j] = c
        for b in bs:
            used[(int(cls[j]), b)].add(c)
    slots, levels, level_class, cls_start = [], [], [], []
    for q in range(ranks + 1):
        cls_start.append(len(slots))
        for c in range(int(colour.max()) + 1):
            members = np.nonzero((cls == q) & (colour == c))[0]
            if members.size == 0:
                continue
            start = len(slots)
            slots.extend(members.tolist())
            levels.append((start, start, len(slots)))            # 1-wide units only
            level_class.append(q)
            slots.extend([-1] * (-len(slots) % 8))
    cls_start.append(len(slots))
    slots = np.asarray(slots, np.int32)
    levels = np.asarray(levels, dtype=oraclepy.LEVEL)
    boundary = order[boundary_rows]                               # body ids of the boundary rows
    return dict(bodies=bodies, joints=joints, cps=cps, slots=slots, levels=levels, level_class=np.asarray(level_class, np.int32),
                cls=cls, cls_start=np.asarray(cls_start), owner=owner, boundary=boundary, ranks=ranks)


def sequential(pb):
    from oracle import oraclepy

    ob, oj = partition.sequential_equivalent(pb["bodies"], pb["joints"], pb["slots"], pb["cls_start"], pb["ranks"])
    ob, oj, ran = oraclepy.solve_scheduled(ob, oj, pb["cps"], pb["slots"], pb["levels"], iters=ITERS)
    return ob[:pb["bodies"].shape[0]], oj, ran


def passes():
    yield -1, 0
    for it in range(ITERS[0]):
        yield 0, it
    for it in range(ITERS[1]):
        yield 1, it


def test_three_ranks_in_one_process_equal_the_sequential_sweep(oracle):
    pb = build_problem(3)
    R = pb["ranks"]
    assert (pb["cls"] == R).sum() > 0 and all((pb["cls"] == q).sum() > 0 for q in range(R))
    ranks = [oracle.Rank(pb["bodies"], pb["joints"], pb["cps"], pb["slots"], pb["levels"], pb["level_class"], q, R) for q in range(R)]
    mine = [pb["boundary"][pb["owner"][pb["boundary"]] == q] for q in range(R)]
    stop, ran = [False, False], [0, 0]
    for phase, it in passes():
        if phase >= 0 and stop[phase]:
            continue
        flags = [r.run(phase, it, cut=False) for r in ranks]
        sent = [ranks[q].get_rows(phase, mine[q]) for q in range(R)]
        for p in range(R):
            for q in range(R):
                if p != q:
                    ranks[p].set_rows(phase, mine[q], sent[q])
        cut_flags = [r.run(phase, it, cut=True) for r in ranks]
        assert len(set(cut_flags)) == 1                         # every rank computes the same cut pass
        if phase >= 0:
            ran[phase] = it + 1
            stop[phase] = not (any(flags) or cut_flags[0])
    # end of solve: everybody's own rows and the accumulators of its own joints go to rank 0
    for q in range(1, R):
        own = np.nonzero(pb["owner"] == q)[0]
        for phase in (0, 1):
            ranks[0].set_rows(phase, own, ranks[q].get_rows(phase, own))
        js = np.nonzero(pb["cls"] == q)[0]
        ranks[0].set_acc(js, ranks[q].get_acc(js))
    b, j = ranks[0].finish()
    ob, oj, want_ran = sequential(pb)
    assert tuple(ran) == want_ran
    assert_records_equal(j, oj, ("normalImpulse", "frictionImpulse"), what="joints")
    assert_records_equal(b, ob, VEL_FIELDS, what="bodies")
    for r in ranks:
        r.close()


def _worker(rank, world_size, port, out):
    import torch.distributed as dist
    from oracle import oraclepy

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    pb = build_problem(world_size)                                # replicated: every process derives the same plan
    me = oraclepy.Rank(pb["bodies"], pb["joints"], pb["cps"], pb["slots"], pb["levels"], pb["level_class"], rank, world_size)
    mine = [pb["boundary"][pb["owner"][pb["boundary"]] == q] for q in range(world_size)]
    stop, ran = [False, False], [0, 0]
    for phase, it in passes():
        if phase >= 0 and stop[phase]:
            continue
        flag = me.run(phase, it, cut=False)
        got = [None] * world_size
        dist.all_gather_object(got, (me.get_rows(phase, mine[rank]), flag))     # the boundary exchange + the interior flag
        for q in range(world_size):
            if q != rank:
                me.set_rows(phase, mine[q], got[q][0])
        cut_flag = me.run(phase, it, cut=True)
        if phase >= 0:
            ran[phase] = it + 1
            stop[phase] = not (any(g[1] for g in got) or cut_flag)
    own = np.nonzero(pb["owner"] == rank)[0]
    js = np.nonzero(pb["cls"] == rank)[0]
    final = [None] * world_size
    dist.all_gather_object(final, (own, me.get_rows(0, own), me.get_rows(1, own), js, me.get_acc(js)))   # end-of-solve exchange
    for q in range(world_size):
        if q != rank:
            o, r0, r1, jq, acc = final[q]
            me.set_rows(0, o, r0)
            me.set_rows(1, o, r1)
            me.set_acc(jq, acc)
    b, j = me.finish()
    ob, oj, want_ran = sequential(pb)
    ok = tuple(ran) == want_ran
    ok = ok and all(np.array_equal(j[f].view(np.uint32), oj[f].view(np.uint32)) for f in ("normalImpulse", "frictionImpulse"))
    ok = ok and all(np.array_equal(np.ascontiguousarray(b[f]).view(np.uint32), np.ascontiguousarray(ob[f]).view(np.uint32)) for f in VEL_FIELDS)
    dist.barrier()
    np.save(out + f".{rank}.npy", np.array([int(ok), ran[0], ran[1]]))
    dist.destroy_process_group()


def test_two_ranks_over_gloo_equal_the_sequential_sweep(oracle, tmp_path):
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
        p.join(timeout=300)
        assert p.exitcode == 0
    res = [np.load(out + f".{r}.npy") for r in range(2)]
    assert all(int(x[0]) == 1 for x in res), res                  # both ranks: identical to the sequential sweep
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]      # and they stopped in the same iterations
