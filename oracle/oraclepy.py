"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/libphyx_oracle.so (the plain-C restatement
of the reference hot path, oracle/phyx_oracle.c).  Only tests/, ``__graft_entry__.smoke()`` and
bench.py's ``cpu_baseline`` leg may import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from phyx_b200 import types as T

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LEVEL = np.dtype([("start", np.int32), ("grouped_end", np.int32), ("end", np.int32)])


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(HERE, "libphyx_oracle.so")
    if not os.path.exists(path):
        build()
    l = C.CDLL(path)
    vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    l.pxo_integrate_velocity.argtypes = [vp, i32, f32, f32]
    l.pxo_integrate_position.argtypes = [vp, i32, f32]
    l.pxo_radix_float.argtypes = [f32]
    l.pxo_radix_float.restype = C.c_uint32
    l.pxo_radix_sort3.argtypes = [vp, i32]
    l.pxo_update_broadphase.argtypes = [vp, i32, vp]
    l.pxo_sweep_pairs.argtypes = [vp, i32, vp, i64, vp]
    l.pxo_sweep_pairs.restype = i64
    l.pxo_prepare_indices.argtypes = [vp, i32, i32, i32, vp]
    l.pxo_prepare_indices.restype = i32
    l.pxo_solve_joints.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, vp, vp]
    l.pxo_solve_scheduled.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, vp]
    l.pxo_rank_create.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, i32, i32]
    l.pxo_rank_create.restype = vp
    l.pxo_rank_destroy.argtypes = [vp]
    l.pxo_rank_destroy.restype = None
    l.pxo_rank_pass.argtypes = [vp, i32, i32, i32]
    l.pxo_rank_pass.restype = i32
    for name in ("pxo_rank_get_rows", "pxo_rank_set_rows"):
        getattr(l, name).argtypes = [vp, i32, vp, i32, vp]
        getattr(l, name).restype = None
    for name in ("pxo_rank_get_acc", "pxo_rank_set_acc"):
        getattr(l, name).argtypes = [vp, vp, i32, vp]
        getattr(l, name).restype = None
    l.pxo_rank_finish.argtypes = [vp, vp, vp]
    l.pxo_rank_finish.restype = None
    _LIB = l
    return l


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def integrate_velocity(bodies, dt, gravity):
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    lib().pxo_integrate_velocity(_p(b), b.shape[0], dt, gravity)
    return b


def integrate_position(bodies, dt):
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    lib().pxo_integrate_position(_p(b), b.shape[0], dt)
    return b


def radix_float(x):
    return int(lib().pxo_radix_float(float(x)))


def radix_sort3(values):
    v = np.asarray(values, dtype=np.uint32)
    e = np.empty((v.shape[0], 2), dtype=np.uint32)
    e[:, 0] = v
    e[:, 1] = np.arange(v.shape[0], dtype=np.uint32)
    lib().pxo_radix_sort3(_p(e), v.shape[0])
    return e


def update_broadphase(bodies):
    b = np.ascontiguousarray(bodies, dtype=T.RIGID_BODY)
    out = np.zeros(b.shape[0], dtype=T.BROADPHASE_ENTRY)
    lib().pxo_update_broadphase(_p(b), b.shape[0], _p(out))
    return out


def sweep_pairs(entries, capacity=None):
    e = np.ascontiguousarray(entries, dtype=T.BROADPHASE_ENTRY)
    cap = capacity or 16 * e.shape[0] + 1024
    out = np.zeros((cap, 2), dtype=np.int32)
    tests = C.c_longlong(0)
    n = lib().pxo_sweep_pairs(_p(e), e.shape[0], _p(out), cap, C.byref(tests))
    assert n <= cap
    return out[:n], int(tests.value)


def prepare_indices(joints, nbodies, group=8):
    j = np.ascontiguousarray(joints, dtype=T.CONTACT_JOINT)
    idx = np.zeros(j.shape[0], dtype=np.int32)
    off = lib().pxo_prepare_indices(_p(j), j.shape[0], nbodies, group, _p(idx))
    return off, idx


def solve_joints(bodies, joints, contact_points, group=8, iters=(20, 20)):
    """Reference-order solve (Island_Single, SIMD width `group`). Returns (bodies, joints, joint_index, iters_run)."""
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    j = np.array(joints, dtype=T.CONTACT_JOINT, copy=True)
    cp = np.ascontiguousarray(contact_points, dtype=T.CONTACT_POINT)
    idx = np.zeros(j.shape[0], dtype=np.int32)
    ran = np.zeros(2, dtype=np.int32)
    lib().pxo_solve_joints(_p(b), b.shape[0], _p(j), j.shape[0], _p(cp), group, iters[0], iters[1], _p(idx), _p(ran))
    return b, j, idx, (int(ran[0]), int(ran[1]))


def solve_scheduled(bodies, joints, contact_points, slots, levels, iters=(20, 20)):
    """Sequential solve in slot order. Returns (bodies, joints, iters_run)."""
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    j = np.array(joints, dtype=T.CONTACT_JOINT, copy=True)
    cp = np.ascontiguousarray(contact_points, dtype=T.CONTACT_POINT)
    s = np.ascontiguousarray(slots, dtype=np.int32)
    lv = np.ascontiguousarray(levels, dtype=LEVEL)
    ran = np.zeros(2, dtype=np.int32)
    lib().pxo_solve_scheduled(_p(b), b.shape[0], _p(j), j.shape[0], _p(cp), _p(s), _p(lv), lv.shape[0], iters[0], iters[1], _p(ran))
    return b, j, (int(ran[0]), int(ran[1]))


class Rank:
    """One rank of the partitioned solve, executed literally on the CPU (phyx_oracle.c pxo_rank_*): own copy of the
    rows and accumulators, interior passes, boundary rows in and out, cut passes."""

    def __init__(self, bodies, joints, contact_points, slots, levels, level_class, rank, ranks):
        self.b = np.ascontiguousarray(bodies, dtype=T.RIGID_BODY)
        self.j = np.ascontiguousarray(joints, dtype=T.CONTACT_JOINT)
        cp = np.ascontiguousarray(contact_points, dtype=T.CONTACT_POINT)
        s = np.ascontiguousarray(slots, dtype=np.int32)
        lv = np.ascontiguousarray(levels, dtype=LEVEL)
        lc = np.ascontiguousarray(level_class, dtype=np.int32)
        self.h = lib().pxo_rank_create(_p(self.b), self.b.shape[0], _p(self.j), self.j.shape[0], _p(cp), _p(s), s.shape[0], _p(lv), _p(lc), lv.shape[0], rank, ranks)

    def run(self, phase, it, cut):
        return bool(lib().pxo_rank_pass(self.h, phase, it, 1 if cut else 0))

    def get_rows(self, phase, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.zeros((ids.shape[0], 4), dtype=np.float32)
        lib().pxo_rank_get_rows(self.h, phase, _p(ids), ids.shape[0], _p(out))
        return out

    def set_rows(self, phase, ids, rows):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        lib().pxo_rank_set_rows(self.h, phase, _p(ids), ids.shape[0], _p(rows))

    def get_acc(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.zeros((ids.shape[0], 2), dtype=np.float32)
        lib().pxo_rank_get_acc(self.h, _p(ids), ids.shape[0], _p(out))
        return out

    def set_acc(self, ids, acc):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        acc = np.ascontiguousarray(acc, dtype=np.float32)
        lib().pxo_rank_set_acc(self.h, _p(ids), ids.shape[0], _p(acc))

    def finish(self):
        b, j = self.b.copy(), self.j.copy()
        lib().pxo_rank_finish(self.h, _p(b), _p(j))
        return b, j

    def close(self):
        if self.h:
            lib().pxo_rank_destroy(self.h)
            self.h = None
