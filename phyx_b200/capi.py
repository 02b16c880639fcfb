"""ctypes binding of the C ABI in include/phyx_b200.h (phyx_b200/libphyx_b200.so).

This is the only way Python reaches the hot path: there is no Python or CPU implementation behind
it.  If the shared library has not been built, or no B200 is present, the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import types as T

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libphyx_b200.so")

SCHEDULE_COLOUR, SCHEDULE_REPLAY_AVX2, SCHEDULE_REPLAY_SSE2, SCHEDULE_REPLAY_SCALAR = 0, 1, 2, 3
SOLVE_STATIC_DEPS, SOLVE_KEEP_SCHEDULE, SOLVE_HOST_COLOURING = 1, 2, 4

LEVEL = np.dtype([("start", np.int32), ("grouped_end", np.int32), ("end", np.int32)])

# every symbol include/phyx_b200.h declares (tests/test_abi.py checks the header against this list
# and the built library against both)
EXPORTS = (
    "phyx_b200_create",
    "phyx_b200_destroy",
    "phyx_b200_last_error",
    "phyx_b200_version",
    "phyx_b200_launch_count",
    "phyx_b200_alloc_stats",
    "phyx_b200_stream",
    "phyx_b200_synchronize",
    "phyx_b200_upload_bodies",
    "phyx_b200_upload_bodies_async",
    "phyx_b200_download_bodies",
    "phyx_b200_body_count",
    "phyx_b200_host_register",
    "phyx_b200_host_unregister",
    "phyx_b200_dynamic_extent",
    "phyx_b200_integrate_velocity",
    "phyx_b200_integrate_position",
    "phyx_b200_update_broadphase",
    "phyx_b200_download_broadphase",
    "phyx_b200_sweep_pairs",
    "phyx_b200_solve_joints",
    "phyx_b200_get_schedule",
    "phyx_b200_stage_joints",
    "phyx_b200_solve_staged",
    "phyx_b200_fetch_joints",
    "phyx_b200_snapshot_bodies",
    "phyx_b200_restore_bodies",
    "phyx_b200_sweep_pairs_resident",
    "phyx_b200_update_pairs",
    "phyx_b200_update_manifolds",
    "phyx_b200_pack_manifolds",
    "phyx_b200_refresh_contact_joints",
    "phyx_b200_solve_resident",
    "phyx_b200_world_step",
    "phyx_b200_step_mode",
    "phyx_b200_reset_collider",
    "phyx_b200_collider_counts",
    "phyx_b200_download_manifolds",
    "phyx_b200_download_contact_points",
    "phyx_b200_download_joints",
    "phyx_b200_upload_collider",
    "phyx_b200_solve_tuning",
    "phyx_b200_strip_plan",
    "phyx_b200_strip_feedback",
    "phyx_b200_strip_trace",
    "phyx_b200_build_islands",
    "phyx_b200_download_islands",
    "phyx_b200_island_partition",
    "phyx_b200_download_body_owners",
    "phyx_b200_island_exchange_words",
    "phyx_b200_island_pack",
    "phyx_b200_island_unpack",
    "phyx_b200_partition_create",
    "phyx_b200_partition_attach",
    "phyx_b200_partition_destroy",
    "phyx_b200_partition_plan",
    "phyx_b200_solve_partitioned",
    "phyx_b200_solve_partitioned_group",
)


class SolveConfig(C.Structure):
    _fields_ = [
        ("contactIterationsCount", C.c_int32),
        ("penetrationIterationsCount", C.c_int32),
        ("schedule", C.c_int32),
        ("flags", C.c_int32),
    ]


class SolveStats(C.Structure):
    _fields_ = [
        ("joints", C.c_int32),
        ("slots", C.c_int32),
        ("levels", C.c_int32),
        ("contactIterationsRun", C.c_int32),
        ("penetrationIterationsRun", C.c_int32),
        ("wakePasses", C.c_int32),
        ("colourRounds", C.c_int32),
        ("kernelForm", C.c_int32),
        ("activeJointIterations", C.c_int64 * 2),
        ("ms_schedule", C.c_float),
        ("ms_refresh", C.c_float),
        ("ms_iterations", C.c_float),
        ("ms_finish", C.c_float),
        ("ms_total", C.c_float),
        ("ms_h2d", C.c_float),
        ("ms_d2h", C.c_float),
    ]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "activeJointIterations" else getattr(self, k)) for k, _ in self._fields_}


class BroadphaseStats(C.Structure):
    _fields_ = [("tests", C.c_int64), ("pairs", C.c_int64), ("ms_sort", C.c_float), ("ms_sweep", C.c_float), ("ms_total", C.c_float)]


class StepInfo(C.Structure):
    _fields_ = [
        ("deferred", C.c_int32),
        ("stopStage", C.c_int32),
        ("stopReason", C.c_int32),
        ("manifolds", C.c_int32),
        ("contactPoints", C.c_int32),
        ("joints", C.c_int32),
        ("newPairs", C.c_int32),
        ("jointsCreated", C.c_int32),
        ("jointsDeleted", C.c_int32),
        ("graphReplay", C.c_int32),
        ("graphStatus", C.c_int32),
        ("pad_", C.c_int32),
        ("pairs", C.c_int64),
        ("tests", C.c_int64),
        ("deferredSteps", C.c_int64),
        ("deferredStops", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ }


class PhyxError(RuntimeError):
    pass


_LIB = None


def load():
    """Load libphyx_b200.so.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise PhyxError(f"{LIB_PATH} is missing: build it with `make -C phyx_b200/csrc` (there is no CPU fallback)")
    l = C.CDLL(LIB_PATH)
    vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    l.phyx_b200_last_error.restype = C.c_char_p
    l.phyx_b200_version.restype = C.c_char_p
    l.phyx_b200_create.argtypes = [i32, C.POINTER(vp)]
    l.phyx_b200_destroy.argtypes = [vp]
    l.phyx_b200_destroy.restype = None
    l.phyx_b200_launch_count.argtypes = [vp]
    l.phyx_b200_launch_count.restype = i64
    l.phyx_b200_alloc_stats.argtypes = [C.POINTER(i64), C.POINTER(C.c_double)]
    l.phyx_b200_alloc_stats.restype = None
    l.phyx_b200_stream.argtypes = [vp]
    l.phyx_b200_stream.restype = vp
    l.phyx_b200_synchronize.argtypes = [vp]
    l.phyx_b200_upload_bodies.argtypes = [vp, vp, i32]
    l.phyx_b200_upload_bodies_async.argtypes = [vp, vp, i32]
    l.phyx_b200_download_bodies.argtypes = [vp, vp, i32]
    l.phyx_b200_body_count.argtypes = [vp]
    l.phyx_b200_host_register.argtypes = [vp, vp, C.c_size_t]
    l.phyx_b200_host_unregister.argtypes = [vp, vp]
    l.phyx_b200_dynamic_extent.argtypes = [vp, vp]
    l.phyx_b200_integrate_velocity.argtypes = [vp, f32, f32]
    l.phyx_b200_integrate_position.argtypes = [vp, f32]
    l.phyx_b200_update_broadphase.argtypes = [vp]
    l.phyx_b200_download_broadphase.argtypes = [vp, vp, i32]
    l.phyx_b200_sweep_pairs.argtypes = [vp, vp, i64, C.POINTER(i64), C.POINTER(BroadphaseStats)]
    l.phyx_b200_sweep_pairs_resident.argtypes = [vp, C.POINTER(BroadphaseStats)]
    l.phyx_b200_solve_joints.argtypes = [vp, vp, i32, vp, i32, C.POINTER(SolveConfig), C.POINTER(SolveStats)]
    l.phyx_b200_get_schedule.argtypes = [vp, vp, i32, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    l.phyx_b200_stage_joints.argtypes = [vp, vp, i32, vp, i32]
    l.phyx_b200_solve_staged.argtypes = [vp, C.POINTER(SolveConfig), C.POINTER(SolveStats)]
    l.phyx_b200_fetch_joints.argtypes = [vp, vp, i32]
    l.phyx_b200_snapshot_bodies.argtypes = [vp]
    l.phyx_b200_restore_bodies.argtypes = [vp]
    l.phyx_b200_update_pairs.argtypes = [vp, C.POINTER(BroadphaseStats)]
    l.phyx_b200_update_manifolds.argtypes = [vp]
    l.phyx_b200_pack_manifolds.argtypes = [vp]
    l.phyx_b200_refresh_contact_joints.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    l.phyx_b200_solve_resident.argtypes = [vp, C.POINTER(SolveConfig), C.POINTER(SolveStats)]
    l.phyx_b200_world_step.argtypes = [vp, f32, f32, C.POINTER(SolveConfig), C.POINTER(SolveStats), C.POINTER(BroadphaseStats), C.POINTER(StepInfo)]
    l.phyx_b200_step_mode.argtypes = [vp, i32]
    l.phyx_b200_strip_feedback.argtypes = [vp, i32]
    l.phyx_b200_reset_collider.argtypes = [vp]
    l.phyx_b200_collider_counts.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    l.phyx_b200_download_manifolds.argtypes = [vp, vp, i32]
    l.phyx_b200_download_contact_points.argtypes = [vp, vp, i32]
    l.phyx_b200_download_joints.argtypes = [vp, vp, i32]
    l.phyx_b200_upload_collider.argtypes = [vp, vp, i32, vp, vp, i32]
    l.phyx_b200_solve_tuning.argtypes = [vp, i32, i32]
    l.phyx_b200_strip_plan.argtypes = [vp, C.POINTER(i32), vp, vp, i32, vp]
    l.phyx_b200_strip_trace.argtypes = [vp, i32, vp, i64, C.POINTER(i32)]
    l.phyx_b200_build_islands.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    l.phyx_b200_download_islands.argtypes = [vp, vp, vp, i32]
    l.phyx_b200_island_partition.argtypes = [vp, i32, i32]
    l.phyx_b200_download_body_owners.argtypes = [vp, vp, i32]
    l.phyx_b200_island_exchange_words.argtypes = [vp, C.POINTER(i64)]
    l.phyx_b200_island_pack.argtypes = [vp, vp]
    l.phyx_b200_island_unpack.argtypes = [vp, vp]
    l.phyx_b200_partition_create.argtypes = [vp, i32, i32, i32, C.c_size_t, vp, C.POINTER(vp)]
    l.phyx_b200_partition_attach.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(i32)]
    l.phyx_b200_partition_destroy.argtypes = [vp]
    l.phyx_b200_partition_plan.argtypes = [vp, vp, vp, vp]
    l.phyx_b200_solve_partitioned.argtypes = [vp, C.POINTER(SolveConfig), C.POINTER(SolveStats)]
    l.phyx_b200_solve_partitioned_group.argtypes = [C.POINTER(vp), i32, C.POINTER(SolveConfig), C.POINTER(SolveStats)]
    _LIB = l
    return l


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One device context = one World's state in HBM (phyx_b200_ctx)."""

    def __init__(self, device=0):
        self.l = load()
        h = C.c_void_p()
        self._check(self.l.phyx_b200_create(device, C.byref(h)))
        self.h = h

    def _check(self, status):
        if status != 0:
            raise PhyxError(f"phyx_b200 status {status}: {self.l.phyx_b200_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.l.phyx_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- bodies ----
    def upload_bodies(self, bodies):
        b = np.ascontiguousarray(bodies, dtype=T.RIGID_BODY)
        self._check(self.l.phyx_b200_upload_bodies(self.h, _p(b), b.shape[0]))

    def download_bodies(self, out=None):
        n = self.l.phyx_b200_body_count(self.h)
        out = np.zeros(n, dtype=T.RIGID_BODY) if out is None else out
        self._check(self.l.phyx_b200_download_bodies(self.h, _p(out), n))
        return out

    def dynamic_extent(self):
        """(min.x, min.y, max.x, max.y) over the dynamic bodies' AABBs"""
        out = np.zeros(4, np.float32)
        self._check(self.l.phyx_b200_dynamic_extent(self.h, _p(out)))
        return out

    def integrate_velocity(self, dt, gravity):
        self._check(self.l.phyx_b200_integrate_velocity(self.h, dt, gravity))

    def integrate_position(self, dt):
        self._check(self.l.phyx_b200_integrate_position(self.h, dt))

    def snapshot_bodies(self):
        self._check(self.l.phyx_b200_snapshot_bodies(self.h))

    def restore_bodies(self):
        self._check(self.l.phyx_b200_restore_bodies(self.h))

    # ---- broadphase ----
    def update_broadphase(self):
        self._check(self.l.phyx_b200_update_broadphase(self.h))

    def download_broadphase(self):
        n = self.l.phyx_b200_body_count(self.h)
        out = np.zeros(n, dtype=T.BROADPHASE_ENTRY)
        self._check(self.l.phyx_b200_download_broadphase(self.h, _p(out), n))
        return out

    def sweep_pairs(self, capacity=None):
        n = self.l.phyx_b200_body_count(self.h)
        cap = capacity if capacity is not None else 8 * n + 1024
        stats = BroadphaseStats()
        while True:
            out = np.zeros((cap, 2), dtype=np.int32)
            cnt = C.c_int64(0)
            st = self.l.phyx_b200_sweep_pairs(self.h, _p(out), cap, C.byref(cnt), C.byref(stats))
            if st == 4 and capacity is None:  # PHYX_B200_ERR_CAPACITY: retry with the reported size
                cap = int(cnt.value)
                continue
            self._check(st)
            return out[: cnt.value], stats

    def sweep_pairs_resident(self):
        stats = BroadphaseStats()
        self._check(self.l.phyx_b200_sweep_pairs_resident(self.h, C.byref(stats)))
        return stats

    # ---- solve ----
    def solve_joints(self, joints, contact_points, iters=(20, 20), schedule=SCHEDULE_COLOUR, flags=0):
        """Solver::SolveJoints on the resident bodies.  Returns (joints_out, stats)."""
        j = np.array(joints, dtype=T.CONTACT_JOINT, copy=True)
        cp = np.ascontiguousarray(contact_points, dtype=T.CONTACT_POINT)
        cfg = SolveConfig(iters[0], iters[1], schedule, flags)
        stats = SolveStats()
        self._check(self.l.phyx_b200_solve_joints(self.h, _p(j), j.shape[0], _p(cp), cp.shape[0], C.byref(cfg), C.byref(stats)))
        return j, stats

    def stage_joints(self, joints, contact_points):
        j = np.ascontiguousarray(joints, dtype=T.CONTACT_JOINT)
        cp = np.ascontiguousarray(contact_points, dtype=T.CONTACT_POINT)
        self._check(self.l.phyx_b200_stage_joints(self.h, _p(j), j.shape[0], _p(cp), cp.shape[0]))
        self._staged = j.shape[0]

    def solve_staged(self, iters=(20, 20), schedule=SCHEDULE_COLOUR, flags=0):
        cfg = SolveConfig(iters[0], iters[1], schedule, flags)
        stats = SolveStats()
        self._check(self.l.phyx_b200_solve_staged(self.h, C.byref(cfg), C.byref(stats)))
        return stats

    def fetch_joints(self):
        out = np.zeros(self._staged, dtype=T.CONTACT_JOINT)
        self._check(self.l.phyx_b200_fetch_joints(self.h, _p(out), out.shape[0]))
        return out

    # ---- resident collider stages ----
    def update_pairs(self):
        stats = BroadphaseStats()
        self._check(self.l.phyx_b200_update_pairs(self.h, C.byref(stats)))
        return stats

    def update_manifolds(self):
        self._check(self.l.phyx_b200_update_manifolds(self.h))

    def pack_manifolds(self):
        self._check(self.l.phyx_b200_pack_manifolds(self.h))

    def refresh_contact_joints(self):
        m, cr, d = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._check(self.l.phyx_b200_refresh_contact_joints(self.h, C.byref(m), C.byref(cr), C.byref(d)))
        return m.value, cr.value, d.value

    def solve_resident(self, iters=(20, 20), schedule=SCHEDULE_COLOUR, flags=0):
        cfg = SolveConfig(iters[0], iters[1], schedule, flags)
        stats = SolveStats()
        self._check(self.l.phyx_b200_solve_resident(self.h, C.byref(cfg), C.byref(stats)))
        return stats

    def world_step(self, dt, gravity=0.0, iters=(20, 20), schedule=SCHEDULE_COLOUR, flags=0):
        """World::Update as one call (reference src/World.cpp:19-37); returns (solve stats, broadphase stats, step info)."""
        cfg = SolveConfig(iters[0], iters[1], schedule, flags)
        stats, bp, info = SolveStats(), BroadphaseStats(), StepInfo()
        self._check(self.l.phyx_b200_world_step(self.h, dt, gravity, C.byref(cfg), C.byref(stats), C.byref(bp), C.byref(info)))
        return stats, bp, info

    def step_mode(self, deferred):
        self._check(self.l.phyx_b200_step_mode(self.h, int(deferred)))

    def reset_collider(self):
        self._check(self.l.phyx_b200_reset_collider(self.h))

    def collider_counts(self):
        m, p, j = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._check(self.l.phyx_b200_collider_counts(self.h, C.byref(m), C.byref(p), C.byref(j)))
        return m.value, p.value, j.value

    def download_manifolds(self):
        out = np.zeros(self.collider_counts()[0], dtype=T.MANIFOLD)
        self._check(self.l.phyx_b200_download_manifolds(self.h, _p(out), out.shape[0]))
        return out

    def download_contact_points(self):
        out = np.zeros(self.collider_counts()[1], dtype=T.CONTACT_POINT)
        self._check(self.l.phyx_b200_download_contact_points(self.h, _p(out), out.shape[0]))
        return out

    def download_joints(self):
        out = np.zeros(self.collider_counts()[2], dtype=T.CONTACT_JOINT)
        self._check(self.l.phyx_b200_download_joints(self.h, _p(out), out.shape[0]))
        return out

    def upload_collider(self, manifolds, contact_points, joints):
        m = np.ascontiguousarray(manifolds, dtype=T.MANIFOLD)
        cp = np.ascontiguousarray(contact_points, dtype=T.CONTACT_POINT)
        j = np.ascontiguousarray(joints, dtype=T.CONTACT_JOINT)
        assert cp.shape[0] == 2 * m.shape[0]
        self._check(self.l.phyx_b200_upload_collider(self.h, _p(m), m.shape[0], _p(cp), _p(j), j.shape[0]))

    def get_schedule(self):
        ns, nl = C.c_int32(0), C.c_int32(0)
        self._check(self.l.phyx_b200_get_schedule(self.h, None, 0, None, 0, C.byref(ns), C.byref(nl)))
        slots = np.zeros(ns.value, dtype=np.int32)
        levels = np.zeros(nl.value, dtype=LEVEL)
        self._check(self.l.phyx_b200_get_schedule(self.h, _p(slots), ns.value, _p(levels), nl.value, C.byref(ns), C.byref(nl)))
        return slots, levels

    def solve_tuning(self, kernel_form=0, strips=0):
        """kernel_form: 0 choose, 1 streaming, 2 record form, 3 strip-local (required); strips: 0 choose, -1 never, n strips."""
        self._check(self.l.phyx_b200_solve_tuning(self.h, kernel_form, strips))

    def strip_feedback(self, measured):
        """measured=False: strip cuts from predicted work only (bit-reproducible runs); True (default): also from measured cost."""
        self._check(self.l.phyx_b200_strip_feedback(self.h, int(bool(measured))))

    def strip_plan(self):
        """Strip layout of the last solve: dict(strips, cuts, class_slot_start, info...) or None if it did not use strips."""
        n = C.c_int32(0)
        info = np.zeros(8, np.int32)
        self._check(self.l.phyx_b200_strip_plan(self.h, C.byref(n), None, None, 0, _p(info)))
        names = ("usable", "rejected", "max_strip_rows", "max_cut_rows", "max_bin", "last_reject", "colours", "cut_manifolds")
        out = dict(zip(names, (int(v) for v in info)))
        out["strips"] = n.value
        if n.value == 0:
            return out
        cuts, cls = np.zeros(2 * n.value + 1, np.int32), np.zeros(2 * n.value + 1, np.int32)
        self._check(self.l.phyx_b200_strip_plan(self.h, C.byref(n), _p(cuts), _p(cls), cuts.shape[0], _p(info)))
        out["cuts"] = cuts[: n.value + 1].copy()
        out["class_slot_start"] = cls
        return out

    def strip_trace(self, passes=-1, fetch=False):
        """passes >= 0: record that many passes per CTA in the following solves; fetch: stamps [strips][passes][8] (ns) of the last solve."""
        out = None
        tp = getattr(self, "_trace_passes", 0)
        if fetch and tp > 0:
            n = C.c_int32(0)
            buf = np.zeros(1024 * tp * 8, np.uint64)
            self._check(self.l.phyx_b200_strip_trace(self.h, -1, _p(buf), buf.shape[0], C.byref(n)))
            if n.value:
                out = buf[: n.value * tp * 8].reshape(n.value, tp, 8)
        if passes >= 0:
            self._check(self.l.phyx_b200_strip_trace(self.h, passes, None, 0, None))
            self._trace_passes = passes
        return out

    # ---- islands (Solver::GatherIslands) and the island-parallel solve over several devices (phyx_b200/islands.py) ----
    def build_islands(self):
        """Returns (islandCount, islandMaxSize, islands before coalescing)."""
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._check(self.l.phyx_b200_build_islands(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def download_islands(self):
        """(island of every body, coalesced group of every body); -1 for static bodies."""
        n = self.l.phyx_b200_body_count(self.h)
        isl, grp = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._check(self.l.phyx_b200_download_islands(self.h, _p(isl), _p(grp), n))
        return isl, grp

    def island_partition(self, rank, ranks):
        self._check(self.l.phyx_b200_island_partition(self.h, rank, ranks))

    def island_exchange_words(self):
        w = C.c_int64(0)
        self._check(self.l.phyx_b200_island_exchange_words(self.h, C.byref(w)))
        return int(w.value)

    def island_pack(self, device_pointer):
        self._check(self.l.phyx_b200_island_pack(self.h, C.c_void_p(device_pointer)))

    def island_unpack(self, device_pointer):
        self._check(self.l.phyx_b200_island_unpack(self.h, C.c_void_p(device_pointer)))

    # ---- one world over several devices (phyx_b200/partition.py drives these) ----
    def partition_create(self, rank, ranks, boundary_capacity, bulk_bytes):
        """Returns (ipc_handle: 64 bytes, device pointer of this rank's exchange buffer)."""
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        self._check(self.l.phyx_b200_partition_create(self.h, rank, ranks, boundary_capacity, bulk_bytes, handle, C.byref(ptr)))
        return handle.raw, int(ptr.value or 0)

    def partition_attach(self, ranks, ipc_handles=None, local_pointers=None, peer_devices=None):
        hbuf = C.create_string_buffer(b"".join(ipc_handles), 64 * ranks) if ipc_handles is not None else None
        ptrs = (C.c_void_p * ranks)(*local_pointers) if local_pointers is not None else None
        devs = (C.c_int * ranks)(*peer_devices) if peer_devices is not None else None
        self._check(self.l.phyx_b200_partition_attach(self.h, hbuf, ptrs, devs))

    def partition_destroy(self):
        self._check(self.l.phyx_b200_partition_destroy(self.h))

    def partition_plan(self, ranks):
        cuts, bstart, cls = np.zeros(ranks + 1, np.int32), np.zeros(ranks + 1, np.int32), np.zeros(ranks + 2, np.int32)
        self._check(self.l.phyx_b200_partition_plan(self.h, _p(cuts), _p(bstart), _p(cls)))
        return cuts, bstart, cls

    def solve_partitioned(self, iters=(20, 20)):
        cfg = SolveConfig(iters[0], iters[1], SCHEDULE_COLOUR, 0)
        stats = SolveStats()
        self._check(self.l.phyx_b200_solve_partitioned(self.h, C.byref(cfg), C.byref(stats)))
        return stats

    def launch_count(self):
        return int(self.l.phyx_b200_launch_count(self.h))

    def alloc_stats(self):
        """(device allocations made by this process so far, host milliseconds spent in them)"""
        n, ms = C.c_int64(0), C.c_double(0.0)
        self.l.phyx_b200_alloc_stats(C.byref(n), C.byref(ms))
        return int(n.value), float(ms.value)

    def stream(self):
        return int(self.l.phyx_b200_stream(self.h) or 0)

    def synchronize(self):
        self._check(self.l.phyx_b200_synchronize(self.h))
