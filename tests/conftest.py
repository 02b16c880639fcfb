import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def bits(a):
    """Bit pattern view for exact float comparison."""
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_records_equal(a, b, fields=None, what=""):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    for f in fields or a.dtype.names:
        if f.startswith("_"):
            continue
        x, y = bits(a[f]), bits(b[f])
        if not np.array_equal(x, y):
            bad = np.nonzero((x != y).reshape(x.shape[0], -1).any(axis=1))[0]
            raise AssertionError(f"{what}: field {f} differs in {bad.size} rows, first {bad[:5]}: {a[f][bad[:3]]} vs {b[f][bad[:3]]}")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oraclepy

    oraclepy.build()
    return oraclepy


@pytest.fixture(scope="session")
def ref():
    from oracle import refpy

    if not refpy.available("strict"):
        pytest.skip("oracle/_ref not built (make -C oracle ref needs /root/reference)")
    return refpy


def oracle_on_device_schedule(ctx, oracle, bodies, joints, contact_points, iters=(20, 20)):
    """The oracle's sequential sweep over the slot order the device used for its last colour-mode solve.  With the
    strip layout every interior class keeps its own lastIteration word per static body (as the partitioned solve does
    per rank): the sweep runs on the equivalent problem in which class k > 0 references its own copy of each static
    body (phyx_b200.partition.sequential_equivalent).  Returns (bodies, joints, iterations run, slots, levels)."""
    from phyx_b200 import partition

    slots, levels = ctx.get_schedule()
    plan = ctx.strip_plan()
    n = bodies.shape[0]
    if plan["strips"] > 1:
        b2, j2 = partition.sequential_equivalent(bodies, joints, slots, plan["class_slot_start"], plan["strips"])
        ob, oj, ran = oracle.solve_scheduled(b2, j2, contact_points, slots, levels, iters=iters)
        oj = oj.copy()
        oj["body1Index"], oj["body2Index"] = joints["body1Index"], joints["body2Index"]
        ob = ob[:n]
    else:
        ob, oj, ran = oracle.solve_scheduled(bodies, joints, contact_points, slots, levels, iters=iters)
    return ob, oj, ran, slots, levels
