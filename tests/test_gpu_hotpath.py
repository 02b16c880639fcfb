"""Parity of the CUDA hot path (through the C ABI) with the oracle / the reference.

Every comparison is BIT-EXACT (float bit patterns) unless a tolerance is written next to it.
  * golden fixtures (tests/golden, produced by the unmodified reference) pin the small cases;
  * the oracle restatement replays the device's own schedule for colour-mode solves;
  * the live reference (oracle/_ref, travels with the repo) provides larger captured inputs.
"""
import numpy as np
import pytest

from conftest import assert_records_equal, golden
from phyx_b200 import capi, scenes, types as T

pytestmark = pytest.mark.gpu

SOLVE_FILES = ["solve_pyramid_10_s0.npz", "solve_pyramid_10_s5.npz", "solve_pyramid_10_s30.npz", "solve_pyramid_1k_s30.npz",
               "solve_stack_1k_s0.npz", "solve_stack_1k_s40.npz"]
STAGE_FILES = [f.replace("solve_", "stages_") for f in SOLVE_FILES]
VEL_FIELDS = ("velocity", "angularVelocity", "displacingVelocity", "displacingAngularVelocity")


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", STAGE_FILES)
def test_integrate_and_broadphase_match_reference(ctx, name):
    g = golden(name)
    ctx.upload_bodies(g["bodies_start"])
    ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
    b = ctx.download_bodies()
    assert_records_equal(b, g["bodies_after_velocity"], T.BODY_STATE_FIELDS, what="IntegrateVelocity")
    ctx.update_broadphase()
    assert_records_equal(ctx.download_broadphase(), g["broadphase"], what="UpdateBroadphase")
    pairs, stats = ctx.sweep_pairs()
    assert np.array_equal(pairs, g["pairs"])
    assert stats.pairs == pairs.shape[0] and stats.tests >= stats.pairs
    ctx.upload_bodies(g["bodies_solved"])
    ctx.integrate_position(scenes.DT)
    assert_records_equal(ctx.download_bodies(), g["bodies_end"], T.BODY_STATE_FIELDS, what="IntegratePosition")


@pytest.mark.parametrize("name", SOLVE_FILES)
@pytest.mark.parametrize("tag,schedule", [("avx2", capi.SCHEDULE_REPLAY_AVX2), ("sse2", capi.SCHEDULE_REPLAY_SSE2), ("scalar", capi.SCHEDULE_REPLAY_SCALAR)])
@pytest.mark.parametrize("flags", [capi.SOLVE_STATIC_DEPS, 0])
def test_replay_solve_is_bit_equal_to_reference(ctx, name, tag, schedule, flags):
    """Dependency-level replay of the reference's own joint order == the reference's sequential
    SIMD loop, bit for bit (bodies and cached impulses)."""
    g = golden(name)
    ctx.upload_bodies(g["bodies"])
    j, stats = ctx.solve_joints(g["joints"], g["contact_points"], schedule=schedule, flags=flags)
    b = ctx.download_bodies()
    assert_records_equal(j, g[f"joints_{tag}"], what="joints")
    assert_records_equal(b, g[f"bodies_{tag}"], VEL_FIELDS, what="bodies")


@pytest.mark.parametrize("name", SOLVE_FILES)
def test_colour_solve_is_bit_equal_to_oracle_on_same_schedule(ctx, oracle, name):
    g = golden(name)
    ctx.upload_bodies(g["bodies"])
    j, stats = ctx.solve_joints(g["joints"], g["contact_points"], schedule=capi.SCHEDULE_COLOUR)
    b = ctx.download_bodies()
    slots, levels = ctx.get_schedule()
    check_schedule(slots, levels, g["joints"], g["bodies"])
    ob, oj, ran = oracle.solve_scheduled(g["bodies"], g["joints"], g["contact_points"], slots, levels)
    assert (stats.contactIterationsRun, stats.penetrationIterationsRun) == ran
    assert_records_equal(j, oj, what="joints")
    assert_records_equal(b, ob, VEL_FIELDS, what="bodies")


def check_schedule(slots, levels, joints, bodies):
    """Every joint exactly once; inside a level no dynamic body appears twice.  A paired level
    (grouped_end < 0, manifold units) holds the two joints of one body pair in slots 2u, 2u+1: the pair
    counts as one unit."""
    used = slots[slots >= 0]
    assert sorted(used.tolist()) == list(range(joints.shape[0]))
    static = (bodies["invMass"] == 0) & (bodies["invInertia"] == 0)
    for lv in levels:
        s = slots[lv["start"]:lv["end"]]
        if lv["grouped_end"] < 0:
            assert lv["start"] % 2 == 0 and s.size % 2 == 0
            a, b = s[0::2], s[1::2]
            assert np.all(a >= 0)
            two = b >= 0
            assert np.array_equal(joints["body1Index"][a[two]], joints["body1Index"][b[two]])
            assert np.array_equal(joints["body2Index"][a[two]], joints["body2Index"][b[two]])
            s = a
        s = s[s >= 0]
        bs = np.concatenate([joints["body1Index"][s], joints["body2Index"][s]])
        bs = bs[~static[bs]]
        assert np.unique(bs).size == bs.size


def test_solve_edge_cases(ctx):
    g = golden("solve_pyramid_10_s0.npz")
    ctx.upload_bodies(g["bodies"])
    j, stats = ctx.solve_joints(g["joints"][:0], g["contact_points"])
    assert j.shape[0] == 0 and stats.joints == 0
    assert_records_equal(ctx.download_bodies(), g["bodies"], VEL_FIELDS, what="no joints")
    # zero iterations: only the warm start (PreStepJoints) is applied
    j, stats = ctx.solve_joints(g["joints"], g["contact_points"], iters=(0, 0), schedule=capi.SCHEDULE_REPLAY_AVX2)
    assert stats.contactIterationsRun == 0 and stats.penetrationIterationsRun == 0
    # bad index -> error, not a crash
    bad = g["joints"].copy()
    bad["body1Index"][0] = 10**6
    with pytest.raises(capi.PhyxError):
        ctx.solve_joints(bad, g["contact_points"])
    # empty world
    ctx.upload_bodies(np.zeros(0, dtype=T.RIGID_BODY))
    ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
    ctx.update_broadphase()
    pairs, _ = ctx.sweep_pairs()
    assert pairs.shape[0] == 0


def _capture(ref, scene, step):
    w = ref.RefWorld(scenes.make(scene), "strict")
    for _ in range(step):
        w.step()
    w.step_staged(mask=0x3F | ref.SAFE_PAIRS)
    return w.bodies(), w.joints(), w.contact_points()


@pytest.mark.parametrize("scene,step", [("pyramid_10k", 12), ("stack_10k", 12), ("islands_64x20", 12)])
def test_larger_scenes_against_live_reference(ctx, oracle, ref, scene, step):
    b0, j0, cp = _capture(ref, scene, step)
    # replay vs the reference itself
    rb, rj, _ = ref.solve_joints(b0, j0, cp, solve=T.SOLVE_AVX2)
    ctx.upload_bodies(b0)
    j, stats = ctx.solve_joints(j0, cp, schedule=capi.SCHEDULE_REPLAY_AVX2)
    assert_records_equal(j, rj, what="replay joints")
    assert_records_equal(ctx.download_bodies(), rb, VEL_FIELDS, what="replay bodies")
    # colour vs the oracle on the device's schedule
    ctx.upload_bodies(b0)
    j, stats = ctx.solve_joints(j0, cp, schedule=capi.SCHEDULE_COLOUR)
    slots, levels = ctx.get_schedule()
    check_schedule(slots, levels, j0, b0)
    ob, oj, ran = oracle.solve_scheduled(b0, j0, cp, slots, levels)
    assert_records_equal(j, oj, what="colour joints")
    assert_records_equal(ctx.download_bodies(), ob, VEL_FIELDS, what="colour bodies")
    # broadphase + integration on the same state
    ctx.upload_bodies(b0)
    ctx.update_broadphase()
    assert_records_equal(ctx.download_broadphase(), oracle.update_broadphase(b0), what="broadphase")
    pairs, _ = ctx.sweep_pairs()
    assert np.array_equal(pairs, oracle.sweep_pairs(oracle.update_broadphase(b0))[0])
    ctx.integrate_position(scenes.DT)
    assert_records_equal(ctx.download_bodies(), oracle.integrate_position(b0, scenes.DT), T.BODY_STATE_FIELDS, what="IntegratePosition")


def test_radix_sort_is_stable_on_ties_and_ragged_sizes(ctx, oracle):
    """Sizes around tile boundaries, heavy ties (whole columns share min-x), negative / zero keys."""
    for n in (1, 2, 31, 33, 4095, 4096, 4097, 70001):
        b = np.zeros(n, dtype=T.RIGID_BODY)
        i = np.arange(n)
        x = ((i * 7919) % 257 - 128).astype(np.float32) * np.float32(0.5)
        x[::5] = -0.0
        b["aabb_min"][:, 0] = x
        b["aabb_max"][:, 0] = x + 1.0
        b["aabb_min"][:, 1] = (i % 1013).astype(np.float32)
        b["aabb_max"][:, 1] = (i % 1013).astype(np.float32) + 0.75
        ctx.upload_bodies(b)
        ctx.update_broadphase()
        assert_records_equal(ctx.download_broadphase(), oracle.update_broadphase(b), what=f"n={n}")
        pairs, stats = ctx.sweep_pairs()
        want, tests = oracle.sweep_pairs(oracle.update_broadphase(b))
        assert np.array_equal(pairs, want) and stats.tests == tests


def _mix32(x):
    x = np.uint32(x)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x85EBCA6B)
        x ^= x >> np.uint32(13)
        x *= np.uint32(0xC2B2AE35)
        x ^= x >> np.uint32(16)
    return int(x)


@pytest.mark.parametrize("name", ["solve_pyramid_1k_s30.npz", "solve_stack_1k_s40.npz"])
def test_device_colouring_is_priority_first_fit(ctx, name):
    """The device colouring (Jones-Plassmann rounds) must equal sequential first-fit over the joints
    visited in priority order: deterministic, independent of thread timing."""
    g = golden(name)
    joints, bodies = g["joints"], g["bodies"]
    ctx.upload_bodies(bodies)
    ctx.solve_joints(joints, g["contact_points"], schedule=capi.SCHEDULE_COLOUR)
    slots, levels = ctx.get_schedule()
    check_schedule(slots, levels, joints, bodies)
    got = np.full(joints.shape[0], -1)
    for c, lv in enumerate(levels):
        s = slots[lv["start"]:lv["end"]]
        assert np.all(s >= 0) and np.all(np.diff(s) > 0)  # colour-major, joint order inside a colour
        got[s] = c
    static = (bodies["invMass"] == 0) & (bodies["invInertia"] == 0)
    used = {}
    want = np.zeros(joints.shape[0], dtype=int)
    for j in sorted(range(joints.shape[0]), key=lambda j: (_mix32(j), j)):
        bs = [b for b in (int(joints["body1Index"][j]), int(joints["body2Index"][j])) if not static[b]]
        taken = set().union(*[used.get(b, set()) for b in bs]) if bs else set()
        c = 0
        while c in taken:
            c += 1
        want[j] = c
        for b in bs:
            used.setdefault(b, set()).add(c)
    assert np.array_equal(got, want)


def test_host_colouring_cross_check(ctx, oracle):
    g = golden("solve_pyramid_1k_s30.npz")
    ctx.upload_bodies(g["bodies"])
    j, stats = ctx.solve_joints(g["joints"], g["contact_points"], schedule=capi.SCHEDULE_COLOUR, flags=capi.SOLVE_HOST_COLOURING)
    slots, levels = ctx.get_schedule()
    check_schedule(slots, levels, g["joints"], g["bodies"])
    ob, oj, ran = oracle.solve_scheduled(g["bodies"], g["joints"], g["contact_points"], slots, levels)
    assert_records_equal(j, oj, what="joints")
    assert_records_equal(ctx.download_bodies(), ob, VEL_FIELDS, what="bodies")
