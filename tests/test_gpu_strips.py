"""The strip-local solve (phyx_b200/csrc/strips.cu, kernelForm 3) and the two grid-barrier forms (1 streaming, 2 record)
of the resident pipeline against the oracle: the device's own slot order is replayed sequentially on the CPU and bodies +
cached impulses must agree BIT FOR BIT.  Covers the Solver::SolveJointIsland loops of the reference
(src/Solver.cpp:130-215) in the throughput (colour) mode, at BASELINE configs[0] / configs[1] sizes."""
import numpy as np
import pytest

from conftest import assert_records_equal, oracle_on_device_schedule
from phyx_b200 import capi, scenes, world

pytestmark = pytest.mark.gpu

VEL_FIELDS = ("velocity", "angularVelocity", "displacingVelocity", "displacingAngularVelocity")


def stages_before_solve(ctx):
    ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
    ctx.update_broadphase()
    ctx.update_pairs()
    ctx.update_manifolds()
    ctx.pack_manifolds()
    ctx.refresh_contact_joints()


def run_and_check(ctx, oracle, steps, check_at, form, iters=(20, 20), what=""):
    from test_gpu_hotpath import check_schedule

    seen = []
    for step in range(steps):
        stages_before_solve(ctx)
        checking = step in check_at
        if checking:
            b0, j0, cp = ctx.download_bodies(), ctx.download_joints(), ctx.download_contact_points()
        st = ctx.solve_resident(schedule=capi.SCHEDULE_COLOUR, iters=iters)
        seen.append(st.kernelForm)
        if form:
            assert st.kernelForm == form, f"{what} step {step}: kernel form {st.kernelForm}, wanted {form}"
        if checking:
            ob, oj, ran, slots, levels = oracle_on_device_schedule(ctx, oracle, b0, j0, cp, iters=iters)
            check_schedule(slots, levels, j0, b0)
            assert (st.contactIterationsRun, st.penetrationIterationsRun) == ran, f"{what} step {step}"
            assert_records_equal(ctx.download_joints(), oj, what=f"{what} step {step} joints")
            assert_records_equal(ctx.download_bodies(), ob, VEL_FIELDS, what=f"{what} step {step} bodies")
        ctx.integrate_position(scenes.DT)
    return seen


@pytest.mark.parametrize("scene,strips,steps,check_at", [
    ("pyramid_1k", 1, 12, (0, 1, 11)),
    ("pyramid_1k", 2, 12, (0, 5, 11)),
    ("pyramid_1k", 5, 30, (0, 7, 29)),
    ("stack_1k", 4, 45, (0, 20, 44)),
    ("stack_1k", 0, 10, (0, 9)),
    ("islands_8x10", 3, 12, (0, 11)),
    ("islands_64x20", 16, 8, (0, 7)),
    ("tumble_300", 2, 40, (0, 15, 39)),
    ("pyramid_10k", 0, 8, (0, 7)),
    ("pyramid_10k", 12, 6, (5,)),
    ("stack_10k", 0, 8, (7,)),
    ("wall_18k", 60, 8, (0, 7)),       # one wide island: every strip boundary cuts manifolds
    ("wall_18k", 200, 8, (0, 7)),      # more strips than SMs: several strips per CTA, rows staged per visit
    ("stack_10k", 333, 6, (5,)),       # the same without cut sets
])
def test_strip_solve_equals_oracle_on_its_slot_order(oracle, scene, strips, steps, check_at):
    w = world.World(scenes.make(scene))
    ctx = w.context()
    ctx.solve_tuning(kernel_form=3, strips=strips)
    ctx.upload_bodies(w.bodies())
    run_and_check(ctx, oracle, steps, check_at, 3, what=f"{scene} strips={strips}")
    plan = ctx.strip_plan()
    assert plan["usable"] == 1 and plan["strips"] == (strips or plan["strips"])
    w.close()


@pytest.mark.parametrize("form", [1, 2])
@pytest.mark.parametrize("scene,steps,check_at", [("pyramid_1k", 10, (0, 9)), ("stack_1k", 42, (0, 41)), ("pyramid_10k", 6, (5,))])
def test_grid_barrier_forms_equal_oracle(oracle, form, scene, steps, check_at):
    """k_solve_pairs (1) and k_solve_pairs2 (2) on the colour-major layout, each forced through the ABI."""
    w = world.World(scenes.make(scene))
    ctx = w.context()
    ctx.solve_tuning(kernel_form=form)
    ctx.upload_bodies(w.bodies())
    run_and_check(ctx, oracle, steps, check_at, form, what=f"{scene} form={form}")
    assert ctx.strip_plan()["strips"] == 0
    w.close()


def test_forms_one_and_two_are_bit_identical():
    sc = scenes.make("pyramid_10k")
    out = []
    for form in (1, 2):
        w = world.World(sc)
        ctx = w.context()
        ctx.solve_tuning(kernel_form=form)
        ctx.upload_bodies(w.bodies())
        for _ in range(15):
            stages_before_solve(ctx)
            assert ctx.solve_resident(schedule=capi.SCHEDULE_COLOUR).kernelForm == form
            ctx.integrate_position(scenes.DT)
        out.append((ctx.download_bodies(), ctx.download_joints()))
        w.close()
    assert_records_equal(out[0][0], out[1][0], ("pos", "velocity", "angularVelocity"), what="bodies")
    assert_records_equal(out[0][1], out[1][1], what="joints")


@pytest.mark.parametrize("scene,settle", [("stack_100k", 12), ("pyramid_100k", 8)])
def test_benched_mode_equals_oracle_at_100k(oracle, scene, settle):
    """BASELINE configs[1] size: the default kernel choice (what bench.py times) against the oracle on the device's
    own schedule, on a state reached after `settle` steps."""
    w = world.World(scenes.make(scene), mirror_contents=False)
    ctx = w.context()
    ctx.upload_bodies(w.bodies())
    forms = run_and_check(ctx, oracle, settle + 1, (settle,), 0, what=scene)
    assert forms[-1] == 3, "the default form of the resident pipeline is the strip-local kernel"
    w.close()


def test_rejected_layout_falls_back_to_the_grid_barrier_forms(oracle):
    """A dynamic plank on a row of boxes touches bodies far apart in sorted-x order: with narrow strips its manifolds
    span non-adjacent strips, the strip layout is rejected and the colour-major layout runs (still exact)."""
    n = 600
    rows = [(0.0, 0.0, 0.0, 1e7, 10.0, 1.0)]
    rows += [((i - n / 2) * 21.0, 15.0, 0.0, 10.0, 5.0, 0.0) for i in range(n)]
    rows += [(0.0, 24.9, 0.0, 126.0, 5.0, 0.0)]   # the plank: 12 boxes long, spans several strips
    for layer in range(3):
        rows += [((i - n / 2) * 21.0, 34.8 + 9.9 * layer, 0.0, 10.0, 5.0, 0.0) for i in range(n)]
    w = world.World(np.asarray(rows, dtype=np.float32))
    ctx = w.context()
    ctx.solve_tuning(strips=120)
    ctx.upload_bodies(w.bodies())
    forms = run_and_check(ctx, oracle, 6, (2, 5), 0, what="plank")
    assert all(f in (1, 2) for f in forms[2:]), forms
    assert ctx.strip_plan()["rejected"] & 1
    ctx.solve_tuning(kernel_form=3, strips=120)
    stages_before_solve(ctx)
    with pytest.raises(capi.PhyxError):
        ctx.solve_resident(schedule=capi.SCHEDULE_COLOUR)
    w.close()
