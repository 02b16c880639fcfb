"""bench.py — the hot path of World::Update() on B200, measured per the round contract.

A "step" is one pass of the hot path over one scene state: IntegrateVelocity -> UpdateBroadphase ->
sweep (UpdatePairs) -> SolveJoints (schedule, refresh, warm start, I impulse + D displacement
iterations, finish) -> IntegratePosition, on the BASELINE.json workload "1M-box pyramid, 20+20
iterations" (configs[2], the configuration the metric is quoted on; it fits one GPU).

  value  constraint-iterations/s = joints x (I + D) x steps / time of the timed region, with every
         input already resident in HBM (body SoA snapshot, staged joints and contact points);
  e2e    the same metric through the C ABI with HOST buffers: every step uploads the body array,
         joints and contact points from pinned memory and reads bodies, pairs and cached impulses
         back;
  roofline / cpu_baseline / clocks / gpu_launches: see the keys' comments below and DESIGN.md.

`--impl reference` times the reference's own CPU implementation of the same stages (oracle/_ref,
the unmodified reference compiled by oracle/Makefile) on the box's host cores.

Multi-GPU (`--gpus N` under torchrun): the path shards by island with no data-path collective, so
every rank owns an independent 1M-box pyramid on its own GPU (weak scaling); the barrier and the
max-over-ranks time come from torch.distributed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITERS = (20, 20)
BYTES_IMPULSE, BYTES_DISPLACEMENT, BYTES_PRESTEP = 196, 136, 128  # SURVEY.md §8(d), per joint-iteration
STAGE_KEYS = ("IntegrateVelocity", "UpdateBroadphase", "UpdatePairs", "SolveJoints", "IntegratePosition")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="pyramid_1m")
    ap.add_argument("--settle", type=int, default=4, help="untimed World::Update steps that build the contact state")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
def reference_stage_run(scene, settle, steps, warmup, workers=None):
    """The reference's own implementation of the timed stages (oracle/_ref fast build = the reference
    Makefile's flags), Solve_AVX2 + Island_SingleSloppy on all host cores: the "AVX2 multicore path"."""
    from oracle import refpy
    from phyx_b200 import types as T

    cores = os.cpu_count() or 1
    workers = max(cores - 1, 1) if workers is None else workers
    w = refpy.RefWorld(scene, "fast", workers=workers)
    for _ in range(settle + warmup):
        w.step_staged(solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE_SLOPPY, iters=ITERS, mask=refpy.ALL_STAGES | refpy.SAFE_PAIRS)
    w.reset_stage_ms()
    joint_iters = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step_staged(solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE_SLOPPY, iters=ITERS, mask=refpy.ALL_STAGES | refpy.SAFE_PAIRS)
        joint_iters += len(w.joints()) * sum(ITERS)
    wall = time.perf_counter() - t0
    ms = w.stage_ms()
    hot_ms = sum(ms[k] for k in STAGE_KEYS)
    tests, pairs = w.count_sweep()
    return {
        "value": joint_iters / (hot_ms * 1e-3),
        "ms_per_step": hot_ms / steps,
        "cores": workers + 1,
        "joints": len(w.joints()),
        "stage_ms_per_step": {k: ms[k] / steps for k in ms if k != "PrepareIndices"},
        "full_update_steps_per_s": steps / wall,
        "broadphase_pairs_per_s": pairs / ((ms["UpdateBroadphase"] + ms["UpdatePairs"]) / steps * 1e-3),
    }


def run_reference(args, rank, world_size):
    if rank != 0:
        return
    from oracle import refpy
    from phyx_b200 import scenes

    if not refpy.available("fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libphyx_ref_fast.so was not built (needs /root/reference at build time)"}))
        return
    scene = scenes.make(args.scene)
    steps = min(args.steps, 3)  # bounded sample: the reference needs seconds per 1M-box step
    r = reference_stage_run(scene, args.settle, steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "constraint_iterations_per_sec", "value": r["value"], "unit": "constraint-iterations/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}: {scene.shape[0]} bodies, {r['joints']} joints, {ITERS[0]}+{ITERS[1]} iterations; timed stages: " + ", ".join(STAGE_KEYS),
                   "mode": "Solve_AVX2 / Island_SingleSloppy, reference Makefile flags (-O3 -ffast-math -mavx2 -mfma)"},
        "cpu_baseline": {"value": r["value"], "unit": "constraint-iterations/s", "cores": r["cores"], "kind": "reference",
                         "sample": f"{steps} steps of the same 1M-box workload after {args.settle}+{min(args.warmup, 1)} untimed steps"},
        "e2e": {"value": r["value"], "unit": "constraint-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_ms_per_step": r["stage_ms_per_step"], "steps_per_sec": r["full_update_steps_per_s"], "broadphase_pairs_per_sec": r["broadphase_pairs_per_s"],
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world_size, local_rank):
    import torch

    from phyx_b200 import capi, scenes, types as T, world

    dist = None
    if world_size > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    # ---- workload: settle the scene with full World::Update steps (untimed), capture the solver inputs
    scene = scenes.make(args.scene)
    w = world.World(scene, device=local_rank)
    for _ in range(args.settle):
        w.step(solve=world.SOLVE_B200, iters=ITERS)
    bodies0 = w.bodies()          # host AoS at the start of the next step ...
    w.step_staged(solve=world.SOLVE_B200, iters=ITERS, mask=0x3F)  # ... and that step's IntegrateVelocity .. RefreshContactJoints
    joints0, cps0 = w.joints(), w.contact_points()
    manifolds = len(w.manifolds())
    w.close()
    nb, nj, ncp = bodies0.shape[0], joints0.shape[0], cps0.shape[0]

    ctx = capi.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    ctx.upload_bodies(bodies0)
    ctx.stage_joints(joints0, cps0)
    ctx.snapshot_bodies()

    def resident_step():
        ctx.restore_bodies()                      # d2d: identical work every step (bodies + cached impulses)
        ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
        ctx.update_broadphase()
        bp = ctx.sweep_pairs_resident()
        st = ctx.solve_staged(iters=ITERS, schedule=capi.SCHEDULE_COLOUR)
        ctx.integrate_position(scenes.DT)
        return bp, st

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    for _ in range(max(args.warmup, 3)):
        resident_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    solve_ms, iter_ms, sched_ms, ran = [], [], [], []
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            bp, st = resident_step()
            solve_ms.append(st.ms_total)
            iter_ms.append(st.ms_iterations)
            sched_ms.append(st.ms_schedule)
            ran.append((st.contactIterationsRun, st.penetrationIterationsRun))
        e1.record(stream)
        barrier()
    launches = ctx.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms_total], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world_size * nj * sum(ITERS) / (ms_step * 1e-3)

    # ---- e2e: the same step through the C ABI with host buffers (pinned), copies inside the timed region
    pb = torch.empty(bodies0.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(T.RIGID_BODY)
    pj = torch.empty(joints0.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(T.CONTACT_JOINT)
    pc = torch.empty(cps0.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(T.CONTACT_POINT)
    pp = torch.empty(8 * (4 * nb + 1024), dtype=torch.uint8, pin_memory=True).numpy().view(np.int32).reshape(-1, 2)
    pc[:] = cps0
    import ctypes as C

    def e2e_step():
        pb[:] = bodies0      # the caller's state for this step (host-side refresh of the pinned buffers
        pj[:] = joints0      # is part of what a host application does between Update calls)
        ctx.upload_bodies(pb)
        ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
        ctx.update_broadphase()
        cnt = C.c_int64(0)
        stats = capi.BroadphaseStats()
        ctx._check(ctx.l.phyx_b200_sweep_pairs(ctx.h, pp.ctypes.data_as(C.c_void_p), pp.shape[0], C.byref(cnt), C.byref(stats)))
        cfg = capi.SolveConfig(ITERS[0], ITERS[1], capi.SCHEDULE_COLOUR, 0)
        sst = capi.SolveStats()
        ctx._check(ctx.l.phyx_b200_solve_joints(ctx.h, pj.ctypes.data_as(C.c_void_p), nj, pc.ctypes.data_as(C.c_void_p), ncp, C.byref(cfg), C.byref(sst)))
        ctx.integrate_position(scenes.DT)
        ctx.download_bodies(pb)
        return int(cnt.value)

    e2e_steps = max(2, min(args.steps, 5))
    npairs = e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        npairs = e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_ms], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = bodies0.nbytes + joints0.nbytes + cps0.nbytes
    d2h = bodies0.nbytes + joints0.nbytes + npairs * 8

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (k_solve: prestep + all iterations, one launch per step)
    it_i = float(np.mean([r[0] for r in ran]))
    it_d = float(np.mean([r[1] for r in ran]))
    alg_bytes = nj * (BYTES_PRESTEP + it_i * BYTES_IMPULSE + it_d * BYTES_DISPLACEMENT)
    k_ms = float(np.mean(iter_ms))
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k_solve_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import refpy

        if refpy.available("fast"):
            r = reference_stage_run(scene, args.settle, 2, 1)
            cpu = {"value": r["value"], "unit": "constraint-iterations/s", "cores": r["cores"], "kind": "reference",
                   "sample": "2 steps of the same 1M-box workload (Solve_AVX2 / Island_SingleSloppy, reference Makefile flags), same timed stages",
                   "ms_per_step": r["ms_per_step"], "stage_ms_per_step": r["stage_ms_per_step"]}
        else:
            cpu = {"value": None, "unit": "constraint-iterations/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}

    line = {
        "metric": "constraint_iterations_per_sec", "value": value, "unit": "constraint-iterations/s", "n_gpus": world_size,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}: {nb} bodies, {manifolds} manifolds, {nj} joints per GPU, {ITERS[0]}+{ITERS[1]} iterations (nominal); step = "
                               "IntegrateVelocity + UpdateBroadphase + sweep + SolveJoints(colour schedule) + IntegratePosition",
                   "inputs": "larger than L2: joint streams 4M x 80 B per impulse iteration; body + joint state restored from a device snapshot every step",
                   "parallelism": f"island-parallel x{world_size} (no data-path collective)"},
        "e2e": {"value": world_size * nj * sum(ITERS) / (e2e_ms * 1e-3), "unit": "constraint-iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": "k_solve (warm start + impulse + displacement iterations, persistent)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "iterations_run": [it_i, it_d]},
        "cpu_baseline": cpu,
        "solve_ms_per_step": {"total": float(np.mean(solve_ms)), "schedule": float(np.mean(sched_ms)), "iterations_kernel": k_ms},
        "solve_only_constraint_iterations_per_sec": world_size * nj * sum(ITERS) / (float(np.mean(solve_ms)) * 1e-3),
        "steps_per_sec": world_size * 1e3 / ms_step,
        "broadphase": {"pairs": int(bp.pairs), "tests": int(bp.tests)},
    }
    print(json.dumps(line))
    ctx.close()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world_size)
    else:
        run_ours(args, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
