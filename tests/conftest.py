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


def private_statics(bodies, joints):
    """The equivalent problem the strip-local kernel solves: every (dynamic body, static body) pair gets its own copy of the
    static body, so a static body's lastIteration (reference src/Solver.cpp:790-798, 903-910) no longer couples joints of
    different dynamic bodies (phyx_b200/csrc/strips.cu, "Static bodies").  Returns (bodies', joints')."""
    bodies = np.asarray(bodies)
    joints = np.array(joints, copy=True)
    static = (bodies["invMass"] == 0) & (bodies["invInertia"] == 0)
    b1, b2 = joints["body1Index"].astype(np.int64), joints["body2Index"].astype(np.int64)
    n = bodies.shape[0]
    clones = []
    seen = {}
    for side, other in (("body1Index", b2), ("body2Index", b1)):
        mine = joints[side].astype(np.int64)
        for j in np.nonzero(static[mine] & ~static[other])[0]:
            key = (int(other[j]), int(mine[j]))
            if key not in seen:
                seen[key] = n + len(clones)
                clones.append(int(mine[j]))
            joints[side][j] = seen[key]
    out = np.concatenate([bodies, bodies[np.asarray(clones, dtype=np.int64)]]) if clones else bodies.copy()
    return out, joints


def oracle_on_device_schedule(ctx, oracle, bodies, joints, contact_points, iters=(20, 20)):
    """The oracle's sequential sweep over the slot order the device used for its last colour-mode solve.  The strip-local
    kernel tracks a static body's lastIteration per dynamic partner: there the sweep runs on the equivalent problem of
    private_statics().  Returns (bodies, joints, iterations run, slots, levels)."""
    slots, levels = ctx.get_schedule()
    plan = ctx.strip_plan()
    n = bodies.shape[0]
    if plan["strips"] >= 1:
        b2, j2 = private_statics(bodies, joints)
        ob, oj, ran = oracle.solve_scheduled(b2, j2, contact_points, slots, levels, iters=iters)
        oj = oj.copy()
        oj["body1Index"], oj["body2Index"] = joints["body1Index"], joints["body2Index"]
        ob = ob[:n]
    else:
        ob, oj, ran = oracle.solve_scheduled(bodies, joints, contact_points, slots, levels, iters=iters)
    return ob, oj, ran, slots, levels
