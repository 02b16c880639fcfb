"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/_ref/libphyx_ref_{strict,fast}.so, the
UNMODIFIED reference compiled by oracle/Makefile (see oracle/ref_harness.cpp).

Only tests/, the golden-fixture generator, ``__graft_entry__.smoke()`` and bench.py's
``cpu_baseline`` / ``--impl reference`` legs may import this module.  It never reads
/root/reference at run time: it loads the prebuilt shared objects that travel with the repo
snapshot.
"""
import ctypes as C
import os

import numpy as np

from phyx_b200 import types as T

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def available(flavour="strict"):
    return os.path.exists(os.path.join(HERE, "_ref", f"libphyx_ref_{flavour}.so"))


def lib(flavour="strict"):
    if flavour in _LIBS:
        return _LIBS[flavour]
    path = os.path.join(HERE, "_ref", f"libphyx_ref_{flavour}.so")
    # RTLD_LOCAL: the two flavours define the same symbols and must not interpose each other.
    l = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    l.ref_build_flavour.restype = C.c_char_p
    l.ref_hardware_concurrency.restype = i32
    l.ref_world_create.restype = vp
    l.ref_world_create.argtypes = [i32]
    l.ref_world_destroy.argtypes = [vp]
    l.ref_set_gravity.argtypes = [vp, f32]
    l.ref_add_body.argtypes = [vp, f32, f32, f32, f32, f32, i32]
    l.ref_add_body.restype = i32
    l.ref_add_bodies.argtypes = [vp, vp, i32]
    for name in ("ref_body_count", "ref_joint_count", "ref_manifold_count", "ref_contact_point_count"):
        getattr(l, name).argtypes = [vp]
        getattr(l, name).restype = i32
    for name in ("ref_get_bodies", "ref_get_joints", "ref_get_manifolds", "ref_get_contact_points"):
        getattr(l, name).argtypes = [vp, vp]
    l.ref_set_bodies.argtypes = [vp, vp, i32]
    l.ref_get_broadphase.argtypes = [vp, vp]
    l.ref_get_broadphase.restype = i32
    l.ref_get_joint_index.argtypes = [vp, vp]
    l.ref_get_joint_index.restype = i32
    l.ref_step.argtypes = [vp, f32, i32, i32, i32, i32]
    l.ref_step_staged.argtypes = [vp, f32, i32, i32, i32, i32, i32]
    l.ref_get_stage_ms.argtypes = [vp, vp]
    l.ref_reset_stage_ms.argtypes = [vp]
    l.ref_count_sweep.argtypes = [vp, vp]
    l.ref_prepare_indices.argtypes = [vp, i32, vp]
    l.ref_prepare_indices.restype = i32
    l.ref_gather_islands.argtypes = [vp, i32, vp, vp, vp]
    l.ref_solve_joints.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]
    l.ref_update_broadphase.argtypes = [vp, i32, vp]
    l.ref_all_pairs.argtypes = [vp, i32, vp, i32]
    l.ref_all_pairs.restype = i32
    l.ref_integrate_velocity.argtypes = [vp, i32, f32, f32]
    l.ref_integrate_position.argtypes = [vp, i32, f32]
    l.ref_radix_sort3.argtypes = [vp, i32]
    l.ref_radix_float.argtypes = [f32]
    l.ref_radix_float.restype = C.c_uint32
    _LIBS[flavour] = l
    return l


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


STAGES = (
    "IntegrateVelocity",
    "UpdateBroadphase",
    "UpdatePairs",
    "UpdateManifolds",
    "PackManifolds",
    "RefreshContactJoints",
    "SolveJoints",
    "IntegratePosition",
    "PrepareIndices",
)
ALL_STAGES = 0xFF
SAFE_PAIRS = 0x100  # route UpdatePairs through UpdatePairsParallel (see ref_harness.cpp)


class RefWorld:
    """The reference's World driven through its public API."""

    def __init__(self, scene=None, flavour="strict", workers=0, gravity=-200.0):
        self.l = lib(flavour)
        self.h = self.l.ref_world_create(workers)
        self.l.ref_set_gravity(self.h, gravity)
        if scene is not None:
            self.add_scene(scene)

    def __del__(self):
        try:
            if self.h:
                self.l.ref_world_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add_scene(self, scene):
        rows = np.ascontiguousarray(scene, dtype=np.float32)
        self.l.ref_add_bodies(self.h, _p(rows), rows.shape[0])

    def step(self, dt=1.0 / 60.0, solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE, iters=(20, 20), safe_pairs=True):
        """One World::Update.  safe_pairs=True runs the same eight stages through the public stage
        functions but takes UpdatePairsParallel (identical pair order with 0 workers, immune to the
        reference's DenseHash tombstone bug); safe_pairs=False calls World::Update itself."""
        if safe_pairs:
            self.l.ref_step_staged(self.h, dt, solve, island, iters[0], iters[1], ALL_STAGES | SAFE_PAIRS)
        else:
            self.l.ref_step(self.h, dt, solve, island, iters[0], iters[1])

    def step_staged(self, dt=1.0 / 60.0, solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE, iters=(20, 20), mask=ALL_STAGES):
        self.l.ref_step_staged(self.h, dt, solve, island, iters[0], iters[1], mask)

    def stage_ms(self):
        out = np.zeros(9, dtype=np.float64)
        self.l.ref_get_stage_ms(self.h, _p(out))
        return dict(zip(STAGES, out.tolist()))

    def reset_stage_ms(self):
        self.l.ref_reset_stage_ms(self.h)

    def bodies(self):
        out = np.zeros(self.l.ref_body_count(self.h), dtype=T.RIGID_BODY)
        self.l.ref_get_bodies(self.h, _p(out))
        return out

    def set_bodies(self, bodies):
        bodies = np.ascontiguousarray(bodies, dtype=T.RIGID_BODY)
        self.l.ref_set_bodies(self.h, _p(bodies), bodies.shape[0])

    def joints(self):
        out = np.zeros(self.l.ref_joint_count(self.h), dtype=T.CONTACT_JOINT)
        self.l.ref_get_joints(self.h, _p(out))
        return out

    def manifolds(self):
        out = np.zeros(self.l.ref_manifold_count(self.h), dtype=T.MANIFOLD)
        self.l.ref_get_manifolds(self.h, _p(out))
        return out

    def contact_points(self):
        out = np.zeros(self.l.ref_contact_point_count(self.h), dtype=T.CONTACT_POINT)
        self.l.ref_get_contact_points(self.h, _p(out))
        return out

    def broadphase(self):
        n = self.l.ref_get_broadphase(self.h, None)
        out = np.zeros(n, dtype=T.BROADPHASE_ENTRY)
        self.l.ref_get_broadphase(self.h, _p(out))
        return out

    def joint_index(self):
        n = self.l.ref_get_joint_index(self.h, None)
        out = np.zeros(n, dtype=np.int32)
        self.l.ref_get_joint_index(self.h, _p(out))
        return out

    def gather_islands(self, group=8):
        """Solver::GatherIslands on the current joints: (island per body, coalesced group per body, (islands, islandCount, islandMaxSize))."""
        n = self.l.ref_body_count(self.h)
        isl, grp, out3 = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(3, np.int32)
        self.l.ref_gather_islands(self.h, group, _p(isl), _p(grp), _p(out3))
        return isl, grp, tuple(int(v) for v in out3)

    def count_sweep(self):
        out = np.zeros(2, dtype=np.int64)
        self.l.ref_count_sweep(self.h, _p(out))
        return int(out[0]), int(out[1])

    def prepare_indices(self, group=8):
        out = np.zeros(self.l.ref_joint_count(self.h), dtype=np.int32)
        off = self.l.ref_prepare_indices(self.h, group, _p(out))
        return off, out


# ---- function-level calls on caller arrays --------------------------------------------------


def solve_joints(bodies, joints, contact_points, solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE, iters=(20, 20), workers=0, flavour="strict"):
    """Reference Solver::SolveJoints on copies of the arrays; returns (bodies, joints, joint_index)."""
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    j = np.array(joints, dtype=T.CONTACT_JOINT, copy=True)
    cp = np.ascontiguousarray(contact_points, dtype=T.CONTACT_POINT)
    idx = np.zeros(j.shape[0], dtype=np.int32)
    lib(flavour).ref_solve_joints(_p(b), b.shape[0], _p(j), j.shape[0], _p(cp), cp.shape[0], solve, island, iters[0], iters[1], workers, _p(idx))
    return b, j, idx


def update_broadphase(bodies, flavour="strict"):
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    out = np.zeros(b.shape[0], dtype=T.BROADPHASE_ENTRY)
    lib(flavour).ref_update_broadphase(_p(b), b.shape[0], _p(out))
    return out


def all_pairs(bodies, capacity=None, flavour="strict"):
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    cap = capacity or 16 * b.shape[0] + 1024
    out = np.zeros((cap, 2), dtype=np.int32)
    n = lib(flavour).ref_all_pairs(_p(b), b.shape[0], _p(out), cap)
    assert n <= cap
    return out[:n]


def integrate_velocity(bodies, dt, gravity, flavour="strict"):
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    lib(flavour).ref_integrate_velocity(_p(b), b.shape[0], dt, gravity)
    return b


def integrate_position(bodies, dt, flavour="strict"):
    b = np.array(bodies, dtype=T.RIGID_BODY, copy=True)
    lib(flavour).ref_integrate_position(_p(b), b.shape[0], dt)
    return b


def radix_sort3(values, flavour="strict"):
    """Reference radixSort3 over {value, index=i} entries; returns the sorted (value, index) array."""
    v = np.asarray(values, dtype=np.uint32)
    e = np.empty((v.shape[0], 2), dtype=np.uint32)
    e[:, 0] = v
    e[:, 1] = np.arange(v.shape[0], dtype=np.uint32)
    lib(flavour).ref_radix_sort3(_p(e), v.shape[0])
    return e


def radix_float(x, flavour="strict"):
    return int(lib(flavour).ref_radix_float(float(x)))
