"""The C restatement (oracle/phyx_oracle.c) against the committed golden fixtures, which were
produced by the unmodified reference (tests/golden/make_golden.py).  Bit-exact: every comparison
is on float bit patterns.  CPU only."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_records_equal, golden
from phyx_b200 import scenes, types as T

SOLVE_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "solve_*.npz")))
STAGE_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "stages_*.npz")))


def test_fixture_inventory():
    assert len(SOLVE_FILES) >= 6 and len(STAGE_FILES) >= 6
    assert os.path.exists(os.path.join(GOLDEN, "trajectory.json"))


@pytest.mark.parametrize("name", SOLVE_FILES)
@pytest.mark.parametrize("tag,group", [("avx2", 8), ("sse2", 4), ("scalar", 1)])
def test_solve_joints_matches_reference(oracle, name, tag, group):
    g = golden(name)
    b, j, idx, ran = oracle.solve_joints(g["bodies"], g["joints"], g["contact_points"], group=group)
    assert np.array_equal(idx, g[f"order_{tag}"]), "PrepareIndices order"
    assert_records_equal(j, g[f"joints_{tag}"], what="joints")
    assert_records_equal(b, g[f"bodies_{tag}"], T.BODY_STATE_FIELDS, what="bodies")
    assert 1 <= ran[0] <= 20 and 1 <= ran[1] <= 20


@pytest.mark.parametrize("name", SOLVE_FILES)
def test_prepare_indices_groups_are_independent(oracle, name):
    g = golden(name)
    j = g["joints"]
    off, idx = oracle.prepare_indices(j, g["bodies"].shape[0], 8)
    assert off % 8 == 0 and sorted(idx.tolist()) == list(range(j.shape[0]))
    grp = idx[:off].reshape(-1, 8)
    bodies = np.stack([j["body1Index"][grp], j["body2Index"][grp]], axis=2).reshape(grp.shape[0], 16)
    assert all(len(set(row.tolist())) == 16 for row in bodies)


@pytest.mark.parametrize("name", STAGE_FILES)
def test_stage_functions_match_reference(oracle, name):
    g = golden(name)
    b = oracle.integrate_velocity(g["bodies_start"], scenes.DT, scenes.GRAVITY)
    assert_records_equal(b, g["bodies_after_velocity"], T.BODY_STATE_FIELDS, what="IntegrateVelocity")
    e = oracle.update_broadphase(g["bodies_after_velocity"])
    assert_records_equal(e, g["broadphase"], what="UpdateBroadphase")
    pairs, tests = oracle.sweep_pairs(e)
    assert np.array_equal(pairs, g["pairs"]) and tests >= pairs.shape[0]
    b = oracle.integrate_position(g["bodies_solved"], scenes.DT)
    assert_records_equal(b, g["bodies_end"], T.BODY_STATE_FIELDS, what="IntegratePosition")


def test_radix_known_answers(oracle):
    g = golden("radix.npz")
    keys = np.array([oracle.radix_float(v) for v in g["floats"]], dtype=np.uint32)
    assert np.array_equal(keys, g["keys"])
    out = oracle.radix_sort3(g["sort_in"])
    assert np.array_equal(out, g["sort_out"])
    # stability: equal keys keep index order
    k, i = out[:, 0].astype(np.int64), out[:, 1].astype(np.int64)
    assert np.all((np.diff(k) > 0) | ((np.diff(k) == 0) & (np.diff(i) > 0)))


def test_edge_cases(oracle):
    empty_b = np.zeros(0, dtype=T.RIGID_BODY)
    assert oracle.update_broadphase(empty_b).shape[0] == 0
    assert oracle.sweep_pairs(np.zeros(0, dtype=T.BROADPHASE_ENTRY))[0].shape[0] == 0
    assert oracle.radix_sort3(np.zeros(0, dtype=np.uint32)).shape[0] == 0
    g = golden("solve_pyramid_10_s0.npz")
    b, j, idx, ran = oracle.solve_joints(g["bodies"], g["joints"][:0], g["contact_points"], group=8)
    assert_records_equal(b, g["bodies"], ("velocity", "angularVelocity", "displacingVelocity"), what="no joints")
    # fewer joints than one SIMD group: everything goes down the scalar tail
    off, idx = oracle.prepare_indices(g["joints"][:5], g["bodies"].shape[0], 8)
    assert off == 0 and idx.tolist() == [0, 1, 2, 3, 4]


def test_trajectory_file_shape():
    with open(os.path.join(GOLDEN, "trajectory.json")) as f:
        t = json.load(f)
    assert "pyramid_1k/avx2" in t and set(t["pyramid_1k/avx2"]) == {"1", "10", "100"}
