"""The resident collider stages (UpdatePairs, UpdateManifolds, PackManifolds, RefreshContactJoints on
the device) against the reference, stage by stage: the reference's state before a stage is pushed
to the device, the stage runs there, and the arrays read back must be bit-identical to what the
reference's own stage function leaves behind (order included)."""
import numpy as np
import pytest

from conftest import assert_records_equal, oracle_on_device_schedule
from phyx_b200 import capi, scenes, types as T, world

pytestmark = pytest.mark.gpu

CP_FIELDS = ("delta1", "delta2", "normal", "isMerged", "isNewlyCreated", "solverIndex")


def live_points(manifolds, cps):
    idx = np.concatenate([np.arange(2 * m, 2 * m + c) for m, c in enumerate(manifolds["pointCount"])]) if manifolds.shape[0] else np.zeros(0, dtype=int)
    return cps[idx.astype(int)]


def compare_collider(ctx, r, what):
    m = ctx.download_manifolds()
    assert_records_equal(m, r.manifolds(), what=f"{what}: manifolds")
    assert_records_equal(live_points(m, ctx.download_contact_points()), live_points(r.manifolds(), r.contact_points()), CP_FIELDS, what=f"{what}: contact points")


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("scene,steps", [("pyramid_10", (0, 1, 2, 7, 30)), ("pyramid_1k", (0, 1, 3, 25)), ("stack_1k", (0, 2, 40, 41)), ("islands_8x10", (0, 5)),
                                          ("tumble_300", (0, 10, 35, 36, 80)), ("clump_300", (0, 1))])
def test_each_resident_stage_matches_reference(ctx, ref, scene, steps):
    r = ref.RefWorld(scenes.make(scene), "strict")
    for step in range(max(steps) + 1):
        if step not in steps:
            r.step()
            continue
        # state before the step -> device
        ctx.upload_bodies(r.bodies())
        ctx.upload_collider(r.manifolds(), r.contact_points() if len(r.contact_points()) == 2 * len(r.manifolds()) else np.zeros(2 * len(r.manifolds()), dtype=T.CONTACT_POINT), r.joints())
        r.step_staged(mask=0x01)
        ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
        r.step_staged(mask=0x02)
        ctx.update_broadphase()
        r.step_staged(mask=0x04 | ref.SAFE_PAIRS)
        bp = ctx.update_pairs()
        m = ctx.download_manifolds()
        assert_records_equal(m, r.manifolds(), what=f"step {step} UpdatePairs")
        tests, pairs = r.count_sweep()
        assert (bp.tests, bp.pairs) == (tests, pairs)
        r.step_staged(mask=0x08)
        ctx.update_manifolds()
        compare_collider(ctx, r, f"step {step} UpdateManifolds")
        r.step_staged(mask=0x10)
        ctx.pack_manifolds()
        compare_collider(ctx, r, f"step {step} PackManifolds")
        r.step_staged(mask=0x20)
        ctx.refresh_contact_joints()
        compare_collider(ctx, r, f"step {step} RefreshContactJoints")
        assert_records_equal(ctx.download_joints(), r.joints(), what=f"step {step} joints after refresh")
        r.step_staged(mask=0x40)
        ctx.solve_resident(schedule=capi.SCHEDULE_REPLAY_AVX2)
        assert_records_equal(ctx.download_joints(), r.joints(), what=f"step {step} joints after solve")
        r.step_staged(mask=0x80)
        ctx.integrate_position(scenes.DT)
        assert_records_equal(ctx.download_bodies(), r.bodies(), ("pos", "xVector", "yVector", "velocity", "angularVelocity", "aabb_min", "aabb_max"), what=f"step {step} bodies")


def test_resident_stepping_without_host_round_trips(ref):
    """100 steps driven only by the resident stage calls (nothing uploaded after step 0) end in the
    reference's state, bit for bit."""
    sc = scenes.make("pyramid_1k")
    r = ref.RefWorld(sc, "strict")
    w = world.World(sc)
    w.step_staged(mask=0)  # no-op; creates nothing yet
    ctx = w.context()
    ctx.upload_bodies(w.bodies())
    for _ in range(100):
        r.step()
        ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
        ctx.update_broadphase()
        ctx.update_pairs()
        ctx.update_manifolds()
        ctx.pack_manifolds()
        ctx.refresh_contact_joints()
        ctx.solve_resident(schedule=capi.SCHEDULE_REPLAY_AVX2)
        ctx.integrate_position(scenes.DT)
    assert_records_equal(ctx.download_bodies(), r.bodies(), ("pos", "xVector", "yVector", "velocity", "angularVelocity"), what="bodies")
    assert_records_equal(ctx.download_joints(), r.joints(), what="joints")
    assert_records_equal(ctx.download_manifolds(), r.manifolds(), what="manifolds")


def test_reset_world_clears_device_caches(ref):
    sc = scenes.make("pyramid_10")
    w = world.World(sc)
    for _ in range(5):
        w.step()
    assert len(w.joints()) > 0
    w.reset_world()
    w.add_scene(sc)
    r = ref.RefWorld(sc, "strict")
    for _ in range(5):
        w.step()
        r.step()
    assert_records_equal(w.joints(), r.joints(), what="joints after reset")
    assert_records_equal(w.bodies(), r.bodies(), ("pos", "velocity"), what="bodies after reset")


def test_mirrors_can_be_switched_off(ref):
    sc = scenes.make("pyramid_10")
    w = world.World(sc, mirror_contents=False)
    r = ref.RefWorld(sc, "strict")
    for _ in range(10):
        w.step()
        r.step()
    assert len(w.joints()) == len(r.joints()) and len(w.manifolds()) == len(r.manifolds())   # sizes stay current
    assert_records_equal(w.bodies(), r.bodies(), ("pos", "velocity"), what="bodies")


@pytest.mark.parametrize("scene,steps", [("stack_1k", 60), ("pyramid_1k", 25)])
def test_incremental_colouring_stays_valid_and_exact(oracle, scene, steps):
    """Colour mode over many resident steps: the manifold cache carries colours from step to step and only
    manifolds without one are coloured.  Every step the schedule must still be a proper colouring, and the solve
    must equal the oracle's sequential sweep in slot order, bit for bit."""
    from test_gpu_hotpath import check_schedule, VEL_FIELDS

    w = world.World(scenes.make(scene))
    ctx = w.context()
    ctx.upload_bodies(w.bodies())
    rounds = []
    for step in range(steps):
        ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
        ctx.update_broadphase()
        ctx.update_pairs()
        ctx.update_manifolds()
        ctx.pack_manifolds()
        ctx.refresh_contact_joints()
        b0, j0, cp = ctx.download_bodies(), ctx.download_joints(), ctx.download_contact_points()
        st = ctx.solve_resident(schedule=capi.SCHEDULE_COLOUR)
        rounds.append(st.colourRounds)
        slots, levels = ctx.get_schedule()
        check_schedule(slots, levels, j0, b0)
        if step % 7 == 0 or step == steps - 1:
            ob, oj, ran, _, _ = oracle_on_device_schedule(ctx, oracle, b0, j0, cp)
            assert (st.contactIterationsRun, st.penetrationIterationsRun) == ran
            assert_records_equal(ctx.download_joints(), oj, what=f"step {step} joints")
            assert_records_equal(ctx.download_bodies(), ob, VEL_FIELDS, what=f"step {step} bodies")
        ctx.integrate_position(scenes.DT)
    # after the first (full) build, later steps only colour the few new joints
    assert min(rounds[1:]) < rounds[0]


def test_unit_colouring_is_priority_first_fit_over_manifolds():
    """Resident colour schedule: the unit is the manifold.  A full build must equal sequential first-fit over
    the manifolds in priority order (deterministic), and lay the joints of a manifold out as slot pairs."""
    from test_gpu_hotpath import check_schedule, _mix32

    w = world.World(scenes.make("pyramid_1k"))
    ctx = w.context()
    ctx.solve_tuning(strips=-1)   # colour-major layout: level index = colour
    ctx.upload_bodies(w.bodies())
    ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
    ctx.update_broadphase()
    ctx.update_pairs()
    ctx.update_manifolds()
    ctx.pack_manifolds()
    ctx.refresh_contact_joints()
    bodies, joints, man = ctx.download_bodies(), ctx.download_joints(), ctx.download_manifolds()
    ctx.solve_resident(schedule=capi.SCHEDULE_COLOUR)
    slots, levels = ctx.get_schedule()
    check_schedule(slots, levels, joints, bodies)
    assert np.all(levels["grouped_end"] < 0)
    got = np.full(man.shape[0], -1)
    for c, lv in enumerate(levels):
        a = slots[lv["start"]:lv["end"]:2]
        b = slots[lv["start"] + 1:lv["end"]:2]
        m = joints["contactPointIndex"][a] // 2
        assert np.all(np.diff(m) > 0)                                  # manifold order inside a colour
        assert np.array_equal(joints["contactPointIndex"][a], 2 * m)   # first point, then second
        assert np.array_equal(joints["contactPointIndex"][b[b >= 0]], 2 * m[b >= 0] + 1)
        assert np.array_equal(b >= 0, man["pointCount"][m] > 1)
        got[m] = c
    static = (bodies["invMass"] == 0) & (bodies["invInertia"] == 0)
    used = {}
    want = np.full(man.shape[0], -1)
    for m in sorted(range(man.shape[0]), key=lambda m: (_mix32(m), m)):
        if man["pointCount"][m] == 0:
            continue
        bs = [b for b in (int(man["body1Index"][m]), int(man["body2Index"][m])) if not static[b]]
        taken = set().union(*[used.get(b, set()) for b in bs]) if bs else set()
        c = 0
        while c in taken:
            c += 1
        want[m] = c
        for b in bs:
            used.setdefault(b, set()).add(c)
    assert np.array_equal(got, want)


def test_more_than_64_colours_falls_back_to_the_host_builder(oracle):
    """clump_300: hundreds of joints per body, far beyond the device colouring's 64 colours."""
    from test_gpu_hotpath import check_schedule, VEL_FIELDS

    w = world.World(scenes.make("clump_300"))
    ctx = w.context()
    ctx.upload_bodies(w.bodies())
    ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
    ctx.update_broadphase()
    bp = ctx.update_pairs()
    assert bp.pairs > 40000
    ctx.update_manifolds()
    ctx.pack_manifolds()
    ctx.refresh_contact_joints()
    b0, j0, cp = ctx.download_bodies(), ctx.download_joints(), ctx.download_contact_points()
    st = ctx.solve_resident(schedule=capi.SCHEDULE_COLOUR, iters=(4, 2))
    assert st.levels > 64 and st.colourRounds == 0
    slots, levels = ctx.get_schedule()
    check_schedule(slots, levels, j0, b0)
    ob, oj, ran = oracle.solve_scheduled(b0, j0, cp, slots, levels, iters=(4, 2))
    assert_records_equal(ctx.download_joints(), oj, what="joints")
    assert_records_equal(ctx.download_bodies(), ob, VEL_FIELDS, what="bodies")


def test_tall_static_wall_and_thick_ground_do_not_lose_contacts(ctx, ref):
    """A body whose y reach is thousands of cells high (a wall, a slab with half-height 2000) makes the sweep's cell index
    exceed the int range before it is clamped (advisor finding, round 1): the pairs must still equal the reference's."""
    rows = [(0.0, -1990.0, 0.0, 1e6, 2000.0, 1.0)]                      # thick ground, top at y = 10
    rows += [(-500.0, 3000.0, 0.0, 10.0, 3000.0, 1.0)]                  # tall wall
    rows += [(-479.0 + 21.0 * i, 15.0, 0.0, 10.0, 5.0, 0.0) for i in range(40)]
    rows += [(-479.0 + 21.0 * i, 25.2, 0.0, 10.0, 5.0, 0.0) for i in range(40)]
    sc = np.asarray(rows, dtype=np.float32)
    r = ref.RefWorld(sc, "strict")
    w = world.World(sc)
    c2 = w.context()
    c2.upload_bodies(w.bodies())
    for step in range(6):
        r.step_staged(mask=0x07 | ref.SAFE_PAIRS)
        c2.integrate_velocity(scenes.DT, scenes.GRAVITY)
        c2.update_broadphase()
        bp = c2.update_pairs()
        tests, pairs = r.count_sweep()
        assert (bp.tests, bp.pairs) == (tests, pairs), f"step {step}"
        assert_records_equal(c2.download_manifolds(), r.manifolds(), what=f"step {step} manifolds")
        r.step_staged(mask=0xF8)
        c2.update_manifolds()
        c2.pack_manifolds()
        c2.refresh_contact_joints()
        c2.solve_resident(schedule=capi.SCHEDULE_REPLAY_AVX2)
        c2.integrate_position(scenes.DT)
    assert_records_equal(c2.download_bodies(), r.bodies(), ("pos", "velocity"), what="bodies")
    w.close()
