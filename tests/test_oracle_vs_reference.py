"""The C restatement against the LIVE reference (oracle/_ref/libphyx_ref_strict.so, the
unmodified reference compiled by oracle/Makefile) on scenes stepped here, beyond what the
committed fixtures cover.  Also re-derives the trajectory hashes so compiler/oracle drift is
detected.  CPU only; skipped when oracle/_ref has not been built."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_records_equal
from phyx_b200 import scenes, types as T


@pytest.mark.parametrize("scene,steps", [("pyramid_1k", (0, 7, 60)), ("stack_1k", (3, 50)), ("islands_8x10", (0, 20))])
def test_solve_and_stages_on_live_reference(oracle, ref, scene, steps):
    w = ref.RefWorld(scenes.make(scene), "strict")
    for step in range(max(steps) + 1):
        if step not in steps:
            w.step()
            continue
        b_start = w.bodies()
        w.step_staged(mask=0x01)
        b_vel = w.bodies()
        assert_records_equal(oracle.integrate_velocity(b_start, scenes.DT, scenes.GRAVITY), b_vel, T.BODY_STATE_FIELDS, what="IntegrateVelocity")
        w.step_staged(mask=0x3E | ref.SAFE_PAIRS)
        e = oracle.update_broadphase(b_vel)
        assert_records_equal(e, w.broadphase(), what="UpdateBroadphase")
        assert np.array_equal(oracle.sweep_pairs(e)[0], ref.all_pairs(b_vel))
        b0, j0, cp = w.bodies(), w.joints(), w.contact_points()
        for mode, group in ((T.SOLVE_AVX2, 8), (T.SOLVE_SSE2, 4), (T.SOLVE_SCALAR, 1)):
            rb, rj, ridx = ref.solve_joints(b0, j0, cp, solve=mode)
            ob, oj, oidx, _ = oracle.solve_joints(b0, j0, cp, group=group)
            assert np.array_equal(ridx, oidx)
            assert_records_equal(oj, rj, what=f"joints N={group}")
            assert_records_equal(ob, rb, T.BODY_STATE_FIELDS, what=f"bodies N={group}")
        w.step_staged(mask=0x40)
        b_solved = w.bodies()
        w.step_staged(mask=0x80)
        assert_records_equal(oracle.integrate_position(b_solved, scenes.DT), w.bodies(), T.BODY_STATE_FIELDS, what="IntegratePosition")


def test_reference_trajectory_hashes_are_stable(ref):
    """The reference rebuilt today still produces the committed hashes (guards compiler drift)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    with open(os.path.join(GOLDEN, "trajectory.json")) as f:
        want = json.load(f)
    w = ref.RefWorld(scenes.make("pyramid_1k"), "strict")
    for step in range(1, 11):
        w.step(solve=T.SOLVE_AVX2)
        if step in (1, 10):
            assert mg.fnv1a(mg.state_bytes(w.bodies())) == want["pyramid_1k/avx2"][str(step)]["hash"]
    assert len(w.joints()) == want["pyramid_1k/avx2"]["10"]["joints"]


def test_fast_and_strict_builds_stay_close(ref):
    """the reference's own -ffast-math build drifts ~1e-4 abs from strict on the
    1 k pyramid after 100 steps; this is the noise floor parity claims are read against."""
    sc = scenes.make("pyramid_1k")
    a, b = ref.RefWorld(sc, "strict"), ref.RefWorld(sc, "fast")
    for _ in range(100):
        a.step()
        b.step()
    assert np.abs(a.bodies()["pos"] - b.bodies()["pos"]).max() < 5e-3
