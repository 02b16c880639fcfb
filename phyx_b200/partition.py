"""One world over several devices: host side of the partitioned solve (include/phyx_b200.h
"one world over several devices"; SURVEY.md §8e "an island that spans devices").

Every rank holds the whole world and runs the collider stages redundantly; Solver::SolveJoints is split
by solver row.  This module holds

* the plumbing that connects the ranks' exchange buffers: `LocalGroup` for ranks living in ONE process
  (device pointers are handed over directly; what the single-GPU tests use, and what a single host
  process driving several devices would use) and `attach_process_group` for one process per GPU
  (CUDA IPC handles travel through torch.distributed; the data path itself never touches it);
* `ReplicatedWorld`, the per-rank World::Update loop on top of it;
* `plan_model`, a numpy statement of the partition plan the device builds (row cuts balanced by
  manifold count, interior / cut classes, boundary rows), used to check the device's plan;
* `sequential_equivalent`, the rewrite under which a plain sequential sweep over the partitioned slot
  order reproduces the partitioned execution exactly (every rank its own copy of each static body).
"""
import numpy as np

from . import capi, scenes
from . import types as T


# ---------------------------------------------------------------------------------------------------
def plan_model(row1, row2, ranks):
    """row1/row2: solver row of each coloured manifold's bodies, -1 for a static body.
    Returns (cuts[ranks+1], cls[manifolds] with `ranks` = cut, boundary rows sorted)."""
    row1, row2 = np.asarray(row1, np.int64), np.asarray(row2, np.int64)
    nb = int(max(row1.max(initial=-1), row2.max(initial=-1)) + 1)
    return plan_model_n(row1, row2, ranks, nb)


def plan_model_n(row1, row2, ranks, nb):
    row1, row2 = np.asarray(row1, np.int64), np.asarray(row2, np.int64)
    both = (row1 >= 0) & (row2 >= 0)
    home = np.where(row1 < 0, row2, np.where(row2 < 0, row1, np.minimum(row1, row2)))
    hist = np.bincount(home[home >= 0], minlength=nb)
    prefix = np.concatenate([[0], np.cumsum(hist)])          # prefix[r] = manifolds whose home row is < r
    total = int(prefix[-1])
    cuts = np.zeros(ranks + 1, np.int64)
    cuts[ranks] = nb
    for q in range(1, ranks):
        target = (total * q + ranks - 1) // ranks
        cuts[q] = int(np.searchsorted(prefix, target, side="left"))
    inner = cuts[1:ranks]

    def rank_of(rows):
        return np.searchsorted(inner, rows, side="right")

    k1, k2 = rank_of(row1), rank_of(row2)
    cls = np.where(both, np.where(k1 == k2, k1, ranks), np.where(row1 >= 0, k1, np.where(row2 >= 0, k2, 0)))
    cut = both & (k1 != k2)
    boundary = np.unique(np.concatenate([row1[cut], row2[cut]]))
    return cuts, cls, boundary


def sequential_equivalent(bodies, joints, slots, class_slot_start, ranks):
    """The partitioned solve equals ONE sequential sweep over the slot order, except that a static body's
    lastIteration (Solver.cpp:790-798, 903-910) is tracked per rank.  Returns (bodies', joints') in which the
    joints of rank q > 0 reference rank q's own copy of each static body (copies appended after the originals);
    a plain sequential sweep of that problem in slot order is what the devices compute."""
    bodies = np.asarray(bodies)
    joints = np.array(joints, copy=True)
    static = np.nonzero((bodies["invMass"] == 0) & (bodies["invInertia"] == 0))[0]
    n, ns = bodies.shape[0], static.size
    if ns == 0:
        return bodies.copy(), joints
    clone_of = np.full(n, -1, np.int64)
    clone_of[static] = np.arange(ns)
    out = np.concatenate([bodies] + [bodies[static]] * (ranks - 1))
    slots = np.asarray(slots)
    for q in range(1, ranks):
        js = slots[class_slot_start[q]:class_slot_start[q + 1]]
        js = js[js >= 0]
        for f in ("body1Index", "body2Index"):
            b = joints[f][js]
            is_static = clone_of[b] >= 0
            joints[f][js] = np.where(is_static, n + (q - 1) * ns + clone_of[b], b)
    return out, joints


# ---------------------------------------------------------------------------------------------------
def default_capacities(bodies, ranks):
    """Exchange buffer sizes for a world of `bodies` bodies: boundary rows per peer and pass, bulk bytes per peer
    (32 B per body + 8 B per slot, slots <= ~6 per body incl. padding, if ONE rank owned everything)."""
    boundary = max(4096, int(bodies) // 4)
    bulk = int(bodies) * 32 + (6 * int(bodies) + 64 * 64 * (ranks + 1)) * 8
    return boundary, bulk


class LocalGroup:
    """`ranks` contexts in this process forming one partition (all on `devices[k]`, default device 0)."""

    def __init__(self, contexts, devices=None, capacities=None):
        self.ctx = list(contexts)
        self.ranks = len(self.ctx)
        self.devices = list(devices) if devices is not None else [0] * self.ranks
        n = max(c.l.phyx_b200_body_count(c.h) for c in self.ctx)
        boundary, bulk = capacities or default_capacities(max(n, 1), self.ranks)
        ptrs = []
        for k, c in enumerate(self.ctx):
            _, p = c.partition_create(k, self.ranks, boundary, bulk)
            ptrs.append(p)
        for c in self.ctx:
            c.partition_attach(self.ranks, local_pointers=ptrs, peer_devices=self.devices)

    def solve(self, iters=(20, 20)):
        import ctypes as C

        l = self.ctx[0].l
        handles = (C.c_void_p * self.ranks)(*[c.h for c in self.ctx])
        cfg = capi.SolveConfig(iters[0], iters[1], capi.SCHEDULE_COLOUR, 0)
        stats = (capi.SolveStats * self.ranks)()
        self.ctx[0]._check(l.phyx_b200_solve_partitioned_group(handles, self.ranks, C.byref(cfg), stats))
        return list(stats)

    def close(self):
        for c in self.ctx:
            c.partition_destroy()


def attach_process_group(ctx, capacities, group=None):
    """One process per GPU: create this rank's exchange buffer, swap CUDA IPC handles with the other ranks of the
    torch.distributed group (object all-gather; bootstrap only) and open theirs.  Returns (rank, ranks).
    Every rank reaches the all-gather even if its own buffer could not be created, so a failure raises on all
    ranks instead of leaving the others waiting."""
    import torch.distributed as dist

    rank, ranks = dist.get_rank(group), dist.get_world_size(group)
    handle, error = None, None
    try:
        handle, _ = ctx.partition_create(rank, ranks, capacities[0], capacities[1])
    except Exception as e:  # noqa: BLE001 - reported to every rank below
        error = f"rank {rank}: {e}"
    handles = [None] * ranks
    dist.all_gather_object(handles, handle if error is None else {"error": error}, group=group)
    bad = [h["error"] for h in handles if isinstance(h, dict)]
    if not bad:
        try:
            ctx.partition_attach(ranks, ipc_handles=handles)
        except Exception as e:  # noqa: BLE001
            error = f"rank {rank}: {e}"
    outcome = [None] * ranks
    dist.all_gather_object(outcome, error, group=group)
    bad += [o for o in outcome if o]
    if bad:
        raise capi.PhyxError("partition set-up failed: " + "; ".join(sorted(set(bad))))
    return rank, ranks


class ReplicatedWorld:
    """World::Update with the solve shared between the ranks: every stage but SolveJoints runs redundantly on each
    rank's replica (reference src/World.cpp:19-37), SolveJoints is `solve` (a callable returning solve stats)."""

    def __init__(self, ctx, bodies):
        self.ctx = ctx
        ctx.upload_bodies(bodies)

    def stages_before_solve(self):
        c = self.ctx
        c.integrate_velocity(scenes.DT, scenes.GRAVITY)
        c.update_broadphase()
        bp = c.update_pairs()
        c.update_manifolds()
        c.pack_manifolds()
        c.refresh_contact_joints()
        return bp

    def step(self, iters=(20, 20)):
        bp = self.stages_before_solve()
        st = self.ctx.solve_partitioned(iters)
        self.ctx.integrate_position(scenes.DT)
        return bp, st


def body_records(scene, device=0):
    """RigidBody records (reference AddBody semantics) for a scene array, via the host mirror."""
    from . import world

    w = world.World(scene, device=device, mirror_contents=False)
    b = np.array(w.bodies(), dtype=T.RIGID_BODY, copy=True)
    w.close()
    return b


def run_spanning(scene_name, settle, steps, device, iters=(20, 20), group=None, check=False):
    """ONE world over all ranks of the torch.distributed group (one process per GPU): every rank steps its replica,
    the solve is partitioned.  Returns a dict on every rank (timings are this rank's; `ms_per_step` is the max over
    ranks of the CUDA-event time of the timed steps).  check: rank 0 repeats the run with all ranks as contexts on
    its own device and requires bit-identical bodies; it also times the one-device solve on the same states."""
    import hashlib

    import torch
    import torch.distributed as dist

    def digest(b):
        h = hashlib.sha256()
        for f in ("pos", "xVector", "velocity", "angularVelocity"):
            h.update(np.ascontiguousarray(b[f]).tobytes())
        return h.hexdigest()

    rank, ranks = dist.get_rank(group), dist.get_world_size(group)
    bodies = body_records(scenes.make(scene_name), device=device)
    n = bodies.shape[0]
    ctx = capi.Context(device)
    world = ReplicatedWorld(ctx, bodies)
    attach_process_group(ctx, default_capacities(n, ranks), group=group)
    for _ in range(settle):
        world.step(iters)
    ctx.synchronize()
    dist.barrier(group=group)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    stats = [world.step(iters)[1] for _ in range(steps)]
    e1.record(stream)
    ctx.synchronize()
    dist.barrier(group=group)
    ms = torch.tensor([e0.elapsed_time(e1) / max(steps, 1)], dtype=torch.float64, device=f"cuda:{device}" if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
    digests = [None] * ranks
    dist.all_gather_object(digests, digest(ctx.download_bodies()), group=group)
    cuts, bstart, cls = ctx.partition_plan(ranks)
    last = stats[-1]
    out = {
        "scene": scene_name, "bodies": int(n), "joints": int(last.joints), "ranks": ranks, "steps": steps, "settle": settle,
        "ms_per_step": float(ms.item()),
        "constraint_iterations_per_sec": float(np.mean([s.joints for s in stats])) * sum(iters) / (float(ms.item()) * 1e-3),
        "replicas_identical": len(set(digests)) == 1,
        "solve_ms": float(np.mean([s.ms_total for s in stats])), "passes_ms": float(np.mean([s.ms_iterations for s in stats])),
        "schedule_ms": float(np.mean([s.ms_schedule for s in stats])), "end_exchange_and_finish_ms": float(np.mean([s.ms_finish for s in stats])),
        "iterations_run": [int(last.contactIterationsRun), int(last.penetrationIterationsRun)],
        "row_cuts": cuts.tolist(), "boundary_rows": int(bstart[-1]), "cut_manifolds": int(cls[-1] - cls[-2]) // 2, "slots": int(cls[-1]),
        "exchange": "peer stores over CUDA IPC mapped buffers + sequence flags, once per pass (warm start, impulse and displacement iterations)",
    }
    if check and rank == 0:
        ctxs = [capi.Context(device) for _ in range(ranks)]
        worlds = [ReplicatedWorld(c, bodies) for c in ctxs]
        grp = LocalGroup(ctxs, devices=[device] * ranks)
        one = ReplicatedWorld(capi.Context(device), bodies)
        one_ms = []
        for step in range(settle + steps):
            for w in worlds:
                w.stages_before_solve()
            grp.solve(iters)
            for c in ctxs:
                c.integrate_position(scenes.DT)
            one.stages_before_solve()
            st = one.ctx.solve_resident(iters=iters, schedule=capi.SCHEDULE_COLOUR)
            one.ctx.integrate_position(scenes.DT)
            if step >= settle:
                one_ms.append(st.ms_total)
        out["matches_in_process_group"] = digest(ctxs[0].download_bodies()) == digests[0]
        out["one_device_solve_ms"] = float(np.mean(one_ms))
        grp.close()
        for c in ctxs:
            c.close()
        one.ctx.close()
    dist.barrier(group=group)
    ctx.partition_destroy()
    ctx.close()
    return out
