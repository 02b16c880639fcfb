"""Python handle on the C++ host mirror (phyx_b200/host: World / Collider / Solver with the
reference's own surface) through its C wrapper, phyx_b200/libphyx_b200_host.so.

`World.step()` is `World::Update(queue, dt, configuration)`: integration, broadphase and the
contact solve run in the sm_100a kernels behind the C ABI; there is no CPU implementation of
those stages to fall back to.
"""
import ctypes as C
import os

import numpy as np

from . import capi
from . import types as T

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, "libphyx_b200_host.so")

SOLVE_B200 = 3  # Configuration::Solve_B200 (device colouring); 0..2 replay the reference's modes

STAGES = ("IntegrateVelocity", "UpdateBroadphase", "UpdatePairs", "UpdateManifolds", "PackManifolds", "RefreshContactJoints", "SolveJoints", "IntegratePosition")

_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    capi.load()  # libphyx_b200.so first (same directory, also found through the rpath)
    if not os.path.exists(HOST_LIB_PATH):
        raise capi.PhyxError(f"{HOST_LIB_PATH} is missing: build it with `make -C phyx_b200/host`")
    l = C.CDLL(HOST_LIB_PATH)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    l.phyxw_create.restype = vp
    l.phyxw_create.argtypes = [i32]
    l.phyxw_destroy.argtypes = [vp]
    l.phyxw_set_gravity.argtypes = [vp, f32]
    l.phyxw_set_solve_flags.argtypes = [vp, i32]
    l.phyxw_add_body.argtypes = [vp, f32, f32, f32, f32, f32, i32]
    l.phyxw_add_bodies.argtypes = [vp, vp, i32]
    l.phyxw_step.argtypes = [vp, f32, i32, i32, i32, i32]
    l.phyxw_step_staged.argtypes = [vp, f32, i32, i32, i32, i32, i32]
    for n in ("body", "joint", "manifold", "contact_point", "broadphase"):
        getattr(l, f"phyxw_{n}_count").argtypes = [vp]
    for n in ("bodies", "joints", "manifolds", "contact_points", "broadphase", "stage_ms", "solve_stats", "broadphase_stats"):
        getattr(l, f"phyxw_get_{n}").argtypes = [vp, vp]
    l.phyxw_set_bodies.argtypes = [vp, vp, i32]
    l.phyxw_reset_stage_ms.argtypes = [vp]
    l.phyxw_set_mirror_contents.argtypes = [vp, i32]
    l.phyxw_set_lazy_bodies.argtypes = [vp, i32]
    l.phyxw_sync_bodies.argtypes = [vp]
    l.phyxw_get_sync_ms.argtypes = [vp]
    l.phyxw_get_sync_ms.restype = C.c_double
    l.phyxw_get_step_ms.argtypes = [vp]
    l.phyxw_get_step_ms.restype = C.c_double
    l.phyxw_set_fused_update.argtypes = [vp, i32]
    l.phyxw_last_step_deferred.argtypes = [vp]
    l.phyxw_last_step_deferred.restype = i32
    l.phyxw_reset_world.argtypes = [vp]
    l.phyxw_context.argtypes = [vp]
    l.phyxw_get_island_counts.argtypes = [vp, vp]
    l.phyxw_context.restype = vp
    _LIB = l
    return l


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class World:
    def __init__(self, scene=None, device=0, gravity=-200.0, solve_flags=0, mirror_contents=True, lazy_bodies=False):
        """lazy_bodies: the opt-in World::bodies contract of the host mirror (upload only what set_bodies / AddBody
        changed, download only when bodies() is read); default off = the reference's semantics."""
        self.l = load()
        self.h = self.l.phyxw_create(device)
        self.l.phyxw_set_gravity(self.h, gravity)
        self.l.phyxw_set_solve_flags(self.h, solve_flags)
        self.l.phyxw_set_mirror_contents(self.h, int(mirror_contents))
        self.l.phyxw_set_lazy_bodies(self.h, int(lazy_bodies))
        if scene is not None:
            self.add_scene(scene)

    def close(self):
        if getattr(self, "h", None):
            self.l.phyxw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_scene(self, scene):
        rows = np.ascontiguousarray(scene, dtype=np.float32)
        self.l.phyxw_add_bodies(self.h, _p(rows), rows.shape[0])

    def step(self, dt=1.0 / 60.0, solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE, iters=(20, 20)):
        self.l.phyxw_step(self.h, dt, solve, island, iters[0], iters[1])

    def step_staged(self, dt=1.0 / 60.0, solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE, iters=(20, 20), mask=0xFF):
        self.l.phyxw_step_staged(self.h, dt, solve, island, iters[0], iters[1], mask)

    def _get(self, what, dtype):
        out = np.zeros(getattr(self.l, f"phyxw_{what}_count")(self.h), dtype=dtype)
        getattr(self.l, f"phyxw_get_{what}s" if what != "body" else "phyxw_get_bodies")(self.h, _p(out))
        return out

    def bodies(self):
        return self._get("body", T.RIGID_BODY)

    def set_bodies(self, bodies):
        b = np.ascontiguousarray(bodies, dtype=T.RIGID_BODY)
        self.l.phyxw_set_bodies(self.h, _p(b), b.shape[0])

    def joints(self):
        return self._get("joint", T.CONTACT_JOINT)

    def manifolds(self):
        return self._get("manifold", T.MANIFOLD)

    def contact_points(self):
        return self._get("contact_point", T.CONTACT_POINT)

    def broadphase(self):
        out = np.zeros(self.l.phyxw_broadphase_count(self.h), dtype=T.BROADPHASE_ENTRY)
        self.l.phyxw_get_broadphase(self.h, _p(out))
        return out

    def stage_ms(self):
        out = np.zeros(8, dtype=np.float64)
        self.l.phyxw_get_stage_ms(self.h, _p(out))
        d = dict(zip(STAGES, out.tolist()))
        d["sync"] = float(self.l.phyxw_get_sync_ms(self.h))
        step = float(self.l.phyxw_get_step_ms(self.h))
        if step > 0.0:
            d["WorldStep"] = step   # World::Update took the fused path (one C-ABI call for the eight stages)
        return d

    def set_fused_update(self, on):
        """World::Update through phyx_b200_world_step (default) or through the eight stage calls: same results."""
        self.l.phyxw_set_fused_update(self.h, int(on))

    def last_step_deferred(self):
        return bool(self.l.phyxw_last_step_deferred(self.h))

    def reset_world(self):
        self.l.phyxw_reset_world(self.h)

    def context(self):
        """A capi.Context view of this world's device context (not owned)."""
        ctx = capi.Context.__new__(capi.Context)
        ctx.l = capi.load()
        ctx.h = C.c_void_p(self.l.phyxw_context(self.h))
        ctx.close = lambda: None  # borrowed: the World owns the context
        return ctx

    def reset_stage_ms(self):
        self.l.phyxw_reset_stage_ms(self.h)

    def island_counts(self):
        """(Solver::islandCount, Solver::islandMaxSize) as left by the last SolveJoints"""
        out = (C.c_int * 2)()
        self.l.phyxw_get_island_counts(self.h, out)
        return int(out[0]), int(out[1])

    def solve_stats(self):
        s = capi.SolveStats()
        self.l.phyxw_get_solve_stats(self.h, C.byref(s))
        return s

    def broadphase_stats(self):
        s = capi.BroadphaseStats()
        self.l.phyxw_get_broadphase_stats(self.h, C.byref(s))
        return s
