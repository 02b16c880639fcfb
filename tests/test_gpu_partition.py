"""One world over several ranks (the island that spans devices): the partitioned solve.

These tests run every rank as its own context on cuda:0 (phyx_b200.partition.LocalGroup): the kernels, the
peer stores into the other rank's exchange buffer, the sequence flags and the end-of-solve exchange are
the ones a multi-GPU run uses; only the way the buffers are introduced to each other (pointers instead of
CUDA IPC handles) and the ordering of the launches (events instead of in-kernel waits) differ.

Bar: bit-exact.  The partitioned execution must equal the oracle's plain sequential sweep over the slot
order the devices report, with each rank tracking static bodies' lastIteration on its own copy
(partition.sequential_equivalent), and every rank must end with identical state.
"""
import numpy as np
import pytest

from conftest import assert_records_equal
from phyx_b200 import capi, partition, scenes

pytestmark = pytest.mark.gpu

VEL_FIELDS = ("velocity", "angularVelocity", "displacingVelocity", "displacingAngularVelocity")


def _row_of(ctx, n):
    entries = ctx.download_broadphase()
    row_of = np.empty(n, np.int64)
    row_of[entries["index"].astype(np.int64)] = np.arange(n)
    return row_of


def _check_plan(ctx, ranks, slots, levels, joints, bodies):
    """The device's plan against the numpy model: cuts balance the manifolds, classes follow the cuts, every
    level lives inside one class, the boundary rows are the rows of the cut manifolds."""
    n = bodies.shape[0]
    cuts, bstart, cls_start = ctx.partition_plan(ranks)
    row_of = _row_of(ctx, n)
    static = (bodies["invMass"] == 0) & (bodies["invInertia"] == 0)
    first = np.nonzero(slots[0::2] >= 0)[0] * 2                       # first slot of every manifold
    j = slots[first]
    b1, b2 = joints["body1Index"][j], joints["body2Index"][j]
    r1 = np.where(static[b1], -1, row_of[b1])
    r2 = np.where(static[b2], -1, row_of[b2])
    want_cuts, want_cls, want_boundary = partition.plan_model_n(r1, r2, ranks, n)
    assert np.array_equal(cuts, want_cuts), (cuts, want_cuts)
    got_cls = np.searchsorted(cls_start[1:], first, side="right")
    assert np.array_equal(got_cls, want_cls)
    for q in range(ranks):
        assert bstart[q + 1] - bstart[q] == int(((want_boundary >= cuts[q]) & (want_boundary < cuts[q + 1])).sum())
    for lv in levels:
        q0, q1 = np.searchsorted(cls_start[1:], [lv["start"], lv["end"] - 1], side="right")
        assert q0 == q1
    per_rank = np.bincount(want_cls, minlength=ranks + 1)
    return per_rank


@pytest.mark.parametrize("scene,ranks,steps,iters", [
    ("pyramid_1k", 2, 16, (20, 20)),
    ("pyramid_1k", 3, 9, (20, 20)),
    ("stack_1k", 2, 12, (20, 20)),
    ("platforms_400", 2, 30, (8, 4)),        # many static bodies: per-rank lastIteration words
    ("islands_64x20", 4, 8, (20, 20)),
])
def test_partitioned_solve_equals_the_sequential_sweep(oracle, scene, ranks, steps, iters):
    from test_gpu_hotpath import check_schedule

    bodies = partition.body_records(scenes.make(scene))
    n = bodies.shape[0]
    ctxs = [capi.Context(0) for _ in range(ranks)]
    worlds = [partition.ReplicatedWorld(c, bodies) for c in ctxs]
    group = partition.LocalGroup(ctxs)
    checked = 0
    for step in range(steps):
        for w in worlds:
            w.stages_before_solve()
        check = step % 4 == 3 or step == steps - 1
        if check:
            b0, j0, cp = ctxs[0].download_bodies(), ctxs[0].download_joints(), ctxs[0].download_contact_points()
        stats = group.solve(iters)
        if check and j0.shape[0]:
            slots, levels = ctxs[0].get_schedule()
            check_schedule(slots, levels, j0, b0)
            per_rank = _check_plan(ctxs[0], ranks, slots, levels, j0, b0)
            if scene == "pyramid_1k":
                assert per_rank[ranks] > 0                      # the pyramid is one island: something must straddle the cut
                assert per_rank[:ranks].min() > 0.5 * per_rank[:ranks].max()
            # every rank ends with the same state
            jr, br = [c.download_joints() for c in ctxs], [c.download_bodies() for c in ctxs]
            for k in range(1, ranks):
                assert_records_equal(jr[k], jr[0], what=f"step {step} joints of rank {k}")
                assert_records_equal(br[k], br[0], VEL_FIELDS, what=f"step {step} bodies of rank {k}")
            # ... which is the sequential sweep over the reported slot order
            _, _, cls_start = ctxs[0].partition_plan(ranks)
            ob, oj = partition.sequential_equivalent(b0, j0, slots, cls_start, ranks)
            ob, oj, ran = oracle.solve_scheduled(ob, oj, cp, slots, levels, iters=iters)
            assert (stats[0].contactIterationsRun, stats[0].penetrationIterationsRun) == ran
            assert_records_equal(jr[0], oj, ("normalImpulse", "frictionImpulse"), what=f"step {step} joints")
            assert_records_equal(br[0], ob[:n], VEL_FIELDS, what=f"step {step} bodies")
            checked += 1
        for c in ctxs:
            c.integrate_position(scenes.DT)
    assert checked >= 2
    group.close()
    for c in ctxs:
        c.close()


def test_partitioned_world_stays_close_to_one_device():
    """Another relaxation order of the same joints: as close to the one-device colour schedule as the reference's
    own solve modes are to each other (2-4e-3 of the scene size on this scene after 100 steps)."""
    bodies = partition.body_records(scenes.make("pyramid_1k"))
    one = capi.Context(0)
    one.upload_bodies(bodies)
    ctxs = [capi.Context(0) for _ in range(2)]
    worlds = [partition.ReplicatedWorld(c, bodies) for c in ctxs]
    group = partition.LocalGroup(ctxs)
    for _ in range(60):
        one.integrate_velocity(scenes.DT, scenes.GRAVITY)
        one.update_broadphase()
        one.update_pairs()
        one.update_manifolds()
        one.pack_manifolds()
        one.refresh_contact_joints()
        one.solve_resident(schedule=capi.SCHEDULE_COLOUR)
        one.integrate_position(scenes.DT)
        for w in worlds:
            w.stages_before_solve()
        group.solve()
        for c in ctxs:
            c.integrate_position(scenes.DT)
    a, b, b1 = one.download_bodies(), ctxs[0].download_bodies(), ctxs[1].download_bodies()
    assert_records_equal(b, b1, ("pos", "velocity", "xVector"), what="replicas")
    extent = float(np.abs(a["pos"][1:]).max())
    dev = float(np.abs(a["pos"] - b["pos"]).max()) / extent
    assert dev < 2e-2, dev
    group.close()


def test_partition_errors_are_reported():
    bodies = partition.body_records(scenes.make("pyramid_10"))
    c = capi.Context(0)
    w = partition.ReplicatedWorld(c, bodies)
    w.stages_before_solve()
    with pytest.raises(capi.PhyxError):
        c.solve_partitioned()                                   # not partitioned
    c.partition_create(0, 2, 16, 1 << 16)
    with pytest.raises(capi.PhyxError):
        c.solve_partitioned()                                   # rank 1 not attached
    with pytest.raises(capi.PhyxError):
        c.partition_create(0, 2, 16, 1 << 16)                   # twice
    c.partition_destroy()
    with pytest.raises(capi.PhyxError):
        c.partition_create(3, 2, 16, 1 << 16)                   # rank out of range
    # capacity too small for the boundary rows
    ctxs = [capi.Context(0) for _ in range(2)]
    big = partition.body_records(scenes.make("pyramid_1k"))
    worlds = [partition.ReplicatedWorld(x, big) for x in ctxs]
    group = partition.LocalGroup(ctxs, capacities=(2, 1 << 22))
    for _ in range(3):
        for x in worlds:
            x.stages_before_solve()
        try:
            group.solve()
        except capi.PhyxError as e:
            assert "capacity" in str(e)
            break
        for x in ctxs:
            x.integrate_position(scenes.DT)
    else:
        raise AssertionError("a 2-row boundary capacity must be too small for a pyramid cut in two")
    # one device solve still works on a context after partition_destroy
    w.stages_before_solve()
    st = c.solve_resident(schedule=capi.SCHEDULE_COLOUR)
    assert st.joints > 0


def _run_group(scene, ranks, steps, iters=(20, 20)):
    bodies = partition.body_records(scene)
    ctxs = [capi.Context(0) for _ in range(ranks)]
    worlds = [partition.ReplicatedWorld(c, bodies) for c in ctxs]
    group = partition.LocalGroup(ctxs)
    stats = None
    for _ in range(steps):
        for w in worlds:
            w.stages_before_solve()
        stats = group.solve(iters)
        for c in ctxs:
            c.integrate_position(scenes.DT)
    out = [c.download_bodies() for c in ctxs]
    joints = [c.download_joints() for c in ctxs]
    group.close()
    for c in ctxs:
        c.close()
    return out, joints, stats


def test_more_ranks_than_work():
    """8 ranks on a 55-box pyramid: several ranks own no manifold at all, some own no boundary row; the passes
    and the exchanges must still line up (every rank sends, possibly nothing) and the replicas stay identical."""
    bodies, joints, stats = _run_group(scenes.make("pyramid_10"), 8, 25)
    assert stats[0].joints > 0
    for k in range(1, 8):
        assert_records_equal(bodies[k], bodies[0], ("pos", "velocity", "angularVelocity", "xVector"), what=f"rank {k}")
        assert_records_equal(joints[k], joints[0], what=f"joints of rank {k}")
        assert (stats[k].contactIterationsRun, stats[k].penetrationIterationsRun) == (stats[0].contactIterationsRun, stats[0].penetrationIterationsRun)


def test_world_without_contacts_and_one_sided_world():
    # free fall, no ground: no manifold ever, the partitioned solve is a no-op on every rank
    free = np.array([[0, 100, 0, 10, 5, 0], [50, 100, 0.3, 10, 5, 0], [120, 140, 0, 10, 5, 0]], dtype=np.float32)
    bodies, joints, stats = _run_group(free, 2, 5)
    assert stats[0].joints == 0 and joints[0].shape[0] == 0
    assert_records_equal(bodies[1], bodies[0], ("pos", "velocity"), what="free fall")
    assert np.all(bodies[0]["velocity"][:, 1] < 0)
    # all contacts on the far left, a lone falling box far right: every manifold is interior to rank 0 or touches
    # only the ground, the cut class is empty
    left = scenes.stack(3, 6)
    lone = np.array([[5000, 400, 0, 10, 5, 0]], dtype=np.float32)
    bodies, joints, stats = _run_group(np.concatenate([left, lone]), 3, 20)
    assert stats[0].joints > 0
    for k in (1, 2):
        assert_records_equal(bodies[k], bodies[0], ("pos", "velocity", "angularVelocity"), what=f"rank {k}")


def test_partitioned_solve_at_100k_keeps_replicas_identical_and_exact(oracle):
    """configs[1] size: 100 001 bodies, 2 ranks, a few steps; last step against the oracle."""
    bodies = partition.body_records(scenes.make("stack_100k"))
    n = bodies.shape[0]
    ctxs = [capi.Context(0) for _ in range(2)]
    worlds = [partition.ReplicatedWorld(c, bodies) for c in ctxs]
    group = partition.LocalGroup(ctxs)
    for step in range(6):
        for w in worlds:
            w.stages_before_solve()
        if step == 5:
            b0, j0, cp = ctxs[0].download_bodies(), ctxs[0].download_joints(), ctxs[0].download_contact_points()
        stats = group.solve()
        if step < 5:
            for c in ctxs:
                c.integrate_position(scenes.DT)
    slots, levels = ctxs[0].get_schedule()
    _, _, cls_start = ctxs[0].partition_plan(2)
    b = [c.download_bodies() for c in ctxs]
    assert_records_equal(b[1], b[0], VEL_FIELDS, what="replicas")
    ob, oj = partition.sequential_equivalent(b0, j0, slots, cls_start, 2)
    ob, oj, ran = oracle.solve_scheduled(ob, oj, cp, slots, levels)
    assert (stats[0].contactIterationsRun, stats[0].penetrationIterationsRun) == ran
    assert_records_equal(ctxs[0].download_joints(), oj, ("normalImpulse", "frictionImpulse"), what="joints")
    assert_records_equal(b[0], ob[:n], VEL_FIELDS, what="bodies")
    group.close()


def test_partitioned_layout_runs_on_one_device_too(oracle):
    """A partitioned context can still run the whole schedule alone (phyx_b200_solve_resident): the class-major
    slot order is then one long level table on one device, and the result is the plain sequential sweep over it
    (one device, one copy of every static body: no rewrite needed)."""
    from test_gpu_hotpath import check_schedule

    bodies = partition.body_records(scenes.make("pyramid_1k"))
    ctxs = [capi.Context(0) for _ in range(2)]
    worlds = [partition.ReplicatedWorld(c, bodies) for c in ctxs]
    group = partition.LocalGroup(ctxs)
    for step in range(6):
        for w in worlds:
            w.stages_before_solve()
        if step < 5:
            group.solve()
            for c in ctxs:
                c.integrate_position(scenes.DT)
    c = ctxs[0]
    b0, j0, cp = c.download_bodies(), c.download_joints(), c.download_contact_points()
    st = c.solve_resident(schedule=capi.SCHEDULE_COLOUR)
    slots, levels = c.get_schedule()
    check_schedule(slots, levels, j0, b0)
    _, _, cls_start = c.partition_plan(2)
    assert cls_start[3] == slots.shape[0] and cls_start[2] > cls_start[1] > 0
    ob, oj, ran = oracle.solve_scheduled(b0, j0, cp, slots, levels)
    assert (st.contactIterationsRun, st.penetrationIterationsRun) == ran
    assert_records_equal(c.download_joints(), oj, what="joints")
    assert_records_equal(c.download_bodies(), ob, VEL_FIELDS, what="bodies")
    group.close()
