"""Whole-step parity: the C++ host mirror's World::Update (GPU hot path behind the reference's own
API) against the reference's World stepped on the CPU, same scene, same Configuration.

Solve_AVX2 / Solve_SSE2 / Solve_Scalar replay the reference's joint order, so the trajectories are
compared bit for bit; the north-star tolerance (1e-4 relative on position / velocity after 100
steps) is asserted as well, and the measured deviation is printed."""
import numpy as np
import pytest

from conftest import assert_records_equal
from phyx_b200 import capi, scenes, types as T, world

pytestmark = pytest.mark.gpu

STATE = ("pos", "xVector", "yVector", "velocity", "angularVelocity", "aabb_min", "aabb_max", "geom_pos")


def rel_dev(a, b):
    scale = max(float(np.abs(b["pos"]).max()), 1.0)
    dpos = float(np.abs(a["pos"][1:] - b["pos"][1:]).max()) / scale
    vscale = max(float(np.abs(b["velocity"]).max()), 1.0)
    dvel = float(np.abs(a["velocity"] - b["velocity"]).max()) / vscale
    return dpos, dvel


@pytest.mark.parametrize("scene,steps,mode,flags", [
    ("pyramid_10", 100, T.SOLVE_AVX2, capi.SOLVE_STATIC_DEPS),
    ("pyramid_1k", 100, T.SOLVE_AVX2, capi.SOLVE_STATIC_DEPS),
    ("pyramid_1k", 100, T.SOLVE_AVX2, 0),
    ("pyramid_1k", 40, T.SOLVE_SSE2, 0),
    ("pyramid_1k", 40, T.SOLVE_SCALAR, 0),
    ("stack_1k", 100, T.SOLVE_AVX2, 0),
    ("islands_8x10", 100, T.SOLVE_AVX2, 0),
    ("tumble_300", 150, T.SOLVE_AVX2, 0),      # rotated boxes: every narrowphase branch, manifold churn
    ("tumble_300", 60, T.SOLVE_SCALAR, 0),
    ("tumble_3k", 80, T.SOLVE_AVX2, 0),
    ("platforms_400", 200, T.SOLVE_AVX2, 0),   # bodies with invMass = 0 only (not static): many joints on one dynamic body
    ("pyramid_10k", 100, T.SOLVE_AVX2, 0),     # "after 100 steps" at 10 k bodies (1150 dependency levels per pass)
])
def test_world_update_tracks_reference(ref, scene, steps, mode, flags):
    sc = scenes.make(scene)
    r = ref.RefWorld(sc, "strict")
    w = world.World(sc, solve_flags=flags)
    wakes = 0
    first_diff = None
    for step in range(steps):
        r.step(solve=mode)
        w.step(solve=mode)
        wakes += w.solve_stats().wakePasses
        if first_diff is None and not np.array_equal(r.bodies()["pos"].view(np.uint32), w.bodies()["pos"].view(np.uint32)):
            first_diff = step
    rb, wb = r.bodies(), w.bodies()
    dpos, dvel = rel_dev(wb, rb)
    print(f"\n{scene} mode={mode} flags={flags}: {steps} steps, joints {len(w.joints())}/{len(r.joints())}, "
          f"rel dpos {dpos:.3e} dvel {dvel:.3e}, wake passes {wakes}, first bit difference at step {first_diff}")
    assert len(w.joints()) == len(r.joints()) and len(w.manifolds()) == len(r.manifolds())
    assert dpos <= 1e-4 and dvel <= 1e-4  # north-star tolerance
    assert_records_equal(wb, rb, STATE, what="bodies after N steps")
    assert_records_equal(w.joints(), r.joints(), what="joint cache")
    assert_records_equal(w.manifolds(), r.manifolds(), what="manifolds")


def test_public_stage_functions_are_drop_in(ref):
    """Calling the eight stage functions one by one (as a reference user may) equals Update."""
    sc = scenes.make("pyramid_10")
    a, b, r = world.World(sc), world.World(sc), ref.RefWorld(sc, "strict")
    for _ in range(10):
        a.step()
        for bit in range(8):
            b.step_staged(mask=1 << bit)
        r.step()
    assert_records_equal(a.bodies(), b.bodies(), STATE, what="staged vs Update")
    assert_records_equal(a.bodies(), r.bodies(), STATE, what="vs reference")
    assert_records_equal(b.broadphase(), r.broadphase(), what="collider.broadphase mirror")


def test_caller_edits_between_steps_are_honoured(ref):
    """The demo writes bodies[i].acceleration between steps and flips bodies to static after
    AddBody (reference src/main.cpp:91-93,337-346)."""
    sc = scenes.make("pyramid_10")
    w, r = world.World(sc), ref.RefWorld(sc, "strict")
    for step in range(20):
        for sim in (w, r):
            b = sim.bodies()
            b["acceleration"][5] = (300.0, 50.0)
            if step == 10:
                b["invMass"][7] = 0.0
                b["invInertia"][7] = 0.0
            sim.set_bodies(b)
            sim.step()
    assert_records_equal(w.bodies(), r.bodies(), STATE, what="bodies")


def test_throughput_mode_stays_physical(ref):
    """Solve_B200 (device colouring) applies the same impulses in a different Gauss-Seidel order:
    not bit-comparable with the reference, but it must stay as close to it as the reference's own
    Scalar and SSE2 modes are (~4e-3 of scene size on this scene)."""
    sc = scenes.make("pyramid_1k")
    r = ref.RefWorld(sc, "strict")
    rs = ref.RefWorld(sc, "strict")
    w = world.World(sc)
    for _ in range(100):
        r.step(solve=T.SOLVE_AVX2)
        rs.step(solve=T.SOLVE_SCALAR)
        w.step(solve=world.SOLVE_B200)
    own_spread = rel_dev(rs.bodies(), r.bodies())[0]
    ours = rel_dev(w.bodies(), r.bodies())[0]
    print(f"\ncolour-vs-AVX2 rel dpos {ours:.3e}; reference Scalar-vs-AVX2 {own_spread:.3e}")
    assert ours < max(3 * own_spread, 1e-2)
    assert abs(len(w.joints()) - len(r.joints())) < 0.02 * len(r.joints())


@pytest.mark.parametrize("scene,steps", [("stack_100k", 12), ("pyramid_100k", 12)])
def test_replay_parity_at_100k(ref, scene, steps):
    """BASELINE configs[1] scale: the replay of the reference's AVX2 order stays bit-identical to the
    reference on 100 k bodies (300-400 k joints), collider arrays included."""
    sc = scenes.make(scene)
    r = ref.RefWorld(sc, "strict")
    w = world.World(sc)
    levels = 0
    for _ in range(steps):
        r.step(solve=T.SOLVE_AVX2)
        w.step(solve=T.SOLVE_AVX2)
        levels = max(levels, w.solve_stats().levels)
    rb, wb = r.bodies(), w.bodies()
    dpos, dvel = rel_dev(wb, rb)
    print(f"\n{scene}: {steps} steps, {len(w.joints())} joints, {levels} dependency levels, rel dpos {dpos:.3e} dvel {dvel:.3e}")
    assert_records_equal(wb, rb, STATE, what="bodies")
    assert_records_equal(w.joints(), r.joints(), what="joints")
    assert_records_equal(w.manifolds(), r.manifolds(), what="manifolds")


def test_full_pipeline_at_1m_matches_reference_for_three_steps(ref):
    """BASELINE configs[2] scale (1 M-box pyramid, ~4 M joints): three replay steps, bit-identical."""
    sc = scenes.make("pyramid_1m")
    r = ref.RefWorld(sc, "strict")
    w = world.World(sc, mirror_contents=False)
    for _ in range(3):
        r.step(solve=T.SOLVE_AVX2)
        w.step(solve=T.SOLVE_AVX2)
    rb, wb = r.bodies(), w.bodies()
    assert len(w.joints()) == len(r.joints()) and len(w.manifolds()) == len(r.manifolds())
    assert_records_equal(wb, rb, STATE, what="bodies")
    ctx = w.context()
    assert_records_equal(ctx.download_joints(), r.joints(), what="joints")


def test_lazy_bodies_contract_gives_the_same_trajectory():
    """The host mirror's opt-in World::bodies contract (upload only declared edits, download only on demand) must not
    change a single bit of the simulation, including an edit made between two Updates, and must stop moving the body
    array over PCIe on every Update."""
    sc = scenes.make("pyramid_1k")
    a, b = world.World(sc), world.World(sc, lazy_bodies=True)
    for step in range(30):
        if step == 10:
            for w in (a, b):
                bodies = w.bodies()
                bodies["acceleration"][5] = (3000.0, 500.0)
                w.set_bodies(bodies)
        if step == 20:
            assert_records_equal(b.bodies(), a.bodies(), STATE, what="bodies at step 20")   # a read in the middle syncs
        a.step(solve=world.SOLVE_B200)
        b.step(solve=world.SOLVE_B200)
    assert_records_equal(b.bodies(), a.bodies(), STATE, what="bodies")
    assert_records_equal(b.joints(), a.joints(), what="joints")
    assert a.bodies()["velocity"][5][0] != 0
