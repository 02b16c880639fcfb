"""phyx_b200_world_step (World::Update as one call, reference src/World.cpp:19-37) against the eight stage functions: the
deferred step keeps the counts on the device, sizes everything by predicted bounds and reads back once; its results must be
those of the stage path BIT FOR BIT (bodies, manifolds, contact points, joints incl. order), whether a step runs through,
stops on the device and is finished by the stage functions, or is not eligible at all."""
import numpy as np
import pytest

from conftest import assert_records_equal
from phyx_b200 import capi, scenes, world

pytestmark = pytest.mark.gpu


def stage_step(ctx, iters=(20, 20)):
    ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
    ctx.update_broadphase()
    ctx.update_pairs()
    ctx.update_manifolds()
    ctx.pack_manifolds()
    ctx.refresh_contact_joints()
    st = ctx.solve_resident(schedule=capi.SCHEDULE_COLOUR, iters=iters)
    ctx.integrate_position(scenes.DT)
    return st


def compare(a, b, what):
    assert a.collider_counts() == b.collider_counts(), what
    assert_records_equal(a.download_bodies(), b.download_bodies(), what=f"{what} bodies")
    assert_records_equal(a.download_manifolds(), b.download_manifolds(), what=f"{what} manifolds")
    assert_records_equal(a.download_contact_points(), b.download_contact_points(), what=f"{what} contact points")
    assert_records_equal(a.download_joints(), b.download_joints(), what=f"{what} joints")


def run_pair(scene, steps, mode, check_every=0, iters=(20, 20)):
    sc = scenes.make(scene)
    wa, wb = world.World(sc, mirror_contents=False), world.World(sc, mirror_contents=False)
    a, b = wa.context(), wb.context()
    a.upload_bodies(wa.bodies())
    b.upload_bodies(wb.bodies())
    b.step_mode(mode)
    # cuts balanced by MEASURED strip cost make the slot order of a wide island depend on timing: two contexts then give two
    # valid, different sweeps (each checked against the oracle in test_gpu_strips.py).  Predicted work only: reproducible.
    a.strip_feedback(False)
    b.strip_feedback(False)
    infos = []
    for step in range(steps):
        sa = stage_step(a, iters)
        sb, bp, info = b.world_step(scenes.DT, scenes.GRAVITY, iters=iters)
        infos.append(info.as_dict())
        assert (sa.contactIterationsRun, sa.penetrationIterationsRun, sa.joints, sa.kernelForm) == \
            (sb.contactIterationsRun, sb.penetrationIterationsRun, sb.joints, sb.kernelForm), f"{scene} step {step}: {infos[-1]}"
        assert list(sa.activeJointIterations) == list(sb.activeJointIterations), f"{scene} step {step}"
        if check_every and step % check_every == check_every - 1:
            compare(a, b, f"{scene} step {step}")
    compare(a, b, f"{scene} after {steps} steps")
    wa.close()
    wb.close()
    return infos


@pytest.mark.parametrize("scene,steps", [
    ("pyramid_1k", 40), ("stack_1k", 60), ("islands_64x20", 30), ("tumble_300", 80), ("pyramid_10k", 25), ("stack_10k", 25), ("wall_18k", 15),
])
def test_world_step_equals_the_stage_functions(scene, steps):
    infos = run_pair(scene, steps, 1, check_every=10)
    deferred = sum(i["deferred"] for i in infos)
    assert infos[0]["deferred"] == 0, "the first step of a world has nothing to predict from"
    assert deferred >= steps // 2, f"{scene}: only {deferred} of {steps} steps ran deferred: {[(i['stopStage'], i['stopReason']) for i in infos]}"


@pytest.mark.parametrize("scene,steps", [("tumble_300", 80), ("pyramid_1k", 30), ("stack_10k", 20)])
def test_stopped_steps_are_finished_by_the_stage_functions(scene, steps):
    """Bounds without headroom: every count that grows stops the step on the device; the result must not change."""
    infos = run_pair(scene, steps, 2, check_every=5)
    stops = [(i["stopStage"], i["stopReason"]) for i in infos if i["stopStage"]]
    assert stops, "tight bounds never stopped a step: the resume path was not exercised"
    assert any(i["deferred"] for i in infos)


def test_steady_state_steps_are_replayed_as_a_graph():
    """Same bounds and buffers as the step before: the ~110 stream operations of a step become one graph launch; results
    stay those of the stage path."""
    infos = run_pair("stack_10k", 40, 1, check_every=8)
    replays = sum(i["graphReplay"] for i in infos)
    assert replays >= 10, [(i["deferred"], i["graphReplay"], i["stopStage"], i["stopReason"]) for i in infos]
    infos = run_pair("stack_10k", 12, 3)
    assert not any(i["graphReplay"] for i in infos) and any(i["deferred"] for i in infos)


def test_step_mode_zero_is_the_stage_path():
    infos = run_pair("pyramid_1k", 8, 0)
    assert not any(i["deferred"] for i in infos)


def test_world_step_at_100k_bodies():
    """BASELINE configs[1] size, what bench.py times."""
    infos = run_pair("stack_100k", 14, 1)
    assert sum(i["deferred"] for i in infos) >= 8, [(i["stopStage"], i["stopReason"]) for i in infos]


def test_counts_and_statistics_come_home():
    w = world.World(scenes.make("pyramid_1k"), mirror_contents=False)
    ctx = w.context()
    ctx.upload_bodies(w.bodies())
    for _ in range(12):
        st, bp, info = ctx.world_step(scenes.DT, scenes.GRAVITY)
    assert info.deferred == 1
    m, p, j = ctx.collider_counts()
    assert (info.manifolds, info.contactPoints, info.joints) == (m, p, j) and p == 2 * m
    assert st.joints == j and st.slots >= j and st.kernelForm == 3 and st.contactIterationsRun >= 1
    assert bp.pairs >= m and bp.tests >= bp.pairs
    assert st.ms_iterations > 0 and bp.ms_sweep > 0
    # the stage functions continue from a deferred step's state
    st2 = stage_step(ctx)
    assert st2.joints == ctx.collider_counts()[2]
    w.close()
