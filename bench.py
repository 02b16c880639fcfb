"""bench.py — World::Update() on B200, measured per the round contract.

A "step" is one World::Update over the BASELINE.json workload "1M-box pyramid, full pipeline
(radix broadphase + coloured solve + integrate), 20+20 iterations" (configs[2], the configuration
the metric is quoted on; it fits one GPU): IntegrateVelocity -> UpdateBroadphase -> UpdatePairs ->
UpdateManifolds -> PackManifolds -> RefreshContactJoints -> SolveJoints -> IntegratePosition, every
stage a sm_100a kernel behind the C ABI.  The scene is first advanced `--settle` steps (untimed) so
the contact state is the one the reference's own timing method uses (steps after step 30).

  value  constraint-iterations/s = sum over the timed steps of joints x (I + D) / time, the whole
         state resident in HBM (the eight stage calls of the C ABI, no host I/O);
  e2e    the same metric through the reference-facing call, World::Update of the host mirror, with
         HOST buffers: every step uploads World::bodies (page-locked in place) and reads it back;
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".

`--impl reference` times the reference's own World::Update stages (oracle/_ref: the unmodified
reference compiled by oracle/Makefile with the reference Makefile's flags), Solve_AVX2 /
Island_SingleSloppy on all host cores, on the same scene and metric.

Multi-GPU (`--gpus N` under torchrun): the path shards by island with no data-path collective, so
every rank owns an independent 1M-box world on its own GPU (weak scaling); the barrier and the
max-over-ranks time come from torch.distributed.  After that timed run the ranks also step ONE
1M-box world together (an island that spans devices: partitioned solve) and report
it under "spanning"; it does not enter `value`.
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
# algorithmic bytes per joint-iteration of the reference's loops (a10 / a11 / a9), per body of the radix
# sort (a3: 8 histogram + 3 x 16 scatter), per sweep test (a4).  BYTES_SKIP (two indices + two body rows for a joint the
# lastIteration test skips) is NOT part of §8(d): it only enters the separately reported "with_skips" figure.
BYTES_IMPULSE, BYTES_DISPLACEMENT, BYTES_PRESTEP, BYTES_SKIP = 196, 136, 128, 40
BYTES_RADIX_PER_BODY, BYTES_SWEEP_TEST = 56, 20
KERNEL_FORMS = {0: "k_solve (joint units)", 1: "k_solve_pairs (manifold units, streaming)", 2: "k_solve_pairs2 (manifold units, record form)",
                3: "k_solve_strips (strip-local: rows in shared memory via bulk TMA, neighbour flags)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default=None, help="default: pyramid_1m on one GPU (BASELINE configs[2]); islands_1m on several (configs[3])")
    ap.add_argument("--settle", type=int, default=30, help="untimed World::Update steps that build the contact state")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--step-mode", type=int, default=1, choices=[0, 1, 3],
                    help="phyx_b200_step_mode: 1 deferred steps with CUDA-graph replay (default), 3 deferred without graphs, 0 stage path")
    ap.add_argument("--stage-calls", action="store_true",
                    help="time the eight stage functions of the C ABI instead of phyx_b200_world_step (same results, one read-back per stage)")
    ap.add_argument("--no-parity", dest="parity", action="store_false",
                    help="skip the parity_mode section (replay mode timed on the workload; colour-mode deviation from the reference after 100 steps at 100 k bodies)")
    ap.set_defaults(parity=True)
    ap.add_argument("--no-spanning", dest="spanning", action="store_false",
                    help="with --gpus N > 1 the bench additionally runs ONE world of the scene over all N GPUs (partitioned solve, boundary rows over NVLink "
                         "peer memory) after the timed weak-scaling run and reports it under \"spanning\"; this switches that off")
    ap.add_argument("--spanning", dest="spanning", action="store_true", help=argparse.SUPPRESS)
    ap.set_defaults(spanning=True)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class NearGpu:
    """Run this process on the cores of the GPU's NUMA node while the GPU arm is measured: World::bodies is first touched (and
    page-locked) there, so the 2 x 128 MB per step of the end-to-end path cross ONE socket's PCIe root instead of the
    inter-socket link (seen as 2.3 against 3.0 ms per copy between boxes).  restore() gives all cores back for the CPU legs
    (cpu_baseline, parity), which use every core."""

    def __init__(self, index):
        self.all = None
        self.info = "not bound"
        try:
            import pynvml

            self.all = os.sched_getaffinity(0)
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(ClockSampler._physical_index(index))
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            near = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & self.all
            if near and near != self.all:
                os.sched_setaffinity(0, near)
                self.info = f"{len(near)} of {len(self.all)} cores (NVML cpu affinity of the device)"
            elif near:
                self.info = "device is near every core"
        except Exception as e:  # noqa: BLE001 - a missing NVML or a container that forbids it must not stop the bench
            self.info = f"not bound ({type(e).__name__})"

    def restore(self):
        try:
            if self.all:
                os.sched_setaffinity(0, self.all)
        except Exception:  # noqa: BLE001
            pass


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region: the counters of the B200_PROFILING.md
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.*` line, read through
    NVML (the library nvidia-smi itself prints from) by a thread of this process every 20 ms, started BEFORE
    the warm-up steps: the first query on a freshly started box (and the start-up of an nvidia-smi process)
    holds the driver for ~0.16 s, which landed in the first timed step when the sampler was started with the
    timed region (11.4 instead of 3.5 ms per step over 20 steps).  The subprocess form is the fallback when
    pynvml is missing."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.handle = index, [], None, None, None
        self.recording, self.stop = False, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = "NVML (pynvml), 20 ms period"
            self.thread = threading.Thread(target=self._poll, daemon=True)
        except Exception:
            self.nvml = None
            self.source = "nvidia-smi -lms 100"
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except OSError:
                self.proc = None
            self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _poll(self):
        n = self.nvml
        while not self.stop:
            # query all the time, keep the samples of the timed region only: the FIRST query of a process can
            # block the driver for ~0.16 s on a freshly started box (seen as one 160 ms UpdatePairs call in the
            # first timed step), so it has to happen during the warm-up steps
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                if self.recording:
                    self.rows.append((sm, self.mx, mask))
            except Exception:
                pass
            time.sleep(0.02)

    def _read(self):
        if not self.proc:
            return
        for line in self.proc.stdout:
            if not self.recording:
                continue
            r = [x.strip() for x in line.split(",")]
            try:
                mask = 0
                for (name, bit), v in zip(self.REASONS, r[3:7]):
                    if v.lower().startswith("active"):
                        mask |= bit
                self.rows.append((float(r[0]), float(r[1]), mask))
            except (ValueError, IndexError):
                continue

    def __enter__(self):
        self.rows.clear()
        self.recording = True
        return self

    def __exit__(self, *a):
        if len(self.rows) < 2:
            time.sleep(0.05)   # a very short timed region: take the samples right behind it (clocks do not drop within 50 ms)
        self.recording = False

    def close(self):
        self.stop = True
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = max([r[1] for r in self.rows], default=0.0)
        reasons = sorted({name for r in self.rows for name, bit in self.REASONS if r[2] & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm), "source": self.source}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
def reference_run(scene, settle, steps, warmup, workers=None, count_executed=False):
    """The reference's own World::Update stages (oracle/_ref fast build = the reference Makefile's
    flags), Solve_AVX2 + Island_SingleSloppy on all host cores: its "AVX2 multicore path"."""
    from oracle import refpy
    from phyx_b200 import types as T

    cores = os.cpu_count() or 1
    workers = max(cores - 1, 1) if workers is None else workers
    w = refpy.RefWorld(scene, "fast", workers=workers)
    mask = refpy.ALL_STAGES | refpy.SAFE_PAIRS
    for _ in range(settle + warmup):
        w.step_staged(solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE_SLOPPY, iters=ITERS, mask=mask)
    w.reset_stage_ms()
    joint_iters = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step_staged(solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE_SLOPPY, iters=ITERS, mask=mask)
        joint_iters += len(w.joints()) * sum(ITERS)
    wall = time.perf_counter() - t0
    ms = w.stage_ms()
    tests, pairs = w.count_sweep()
    executed = None
    if count_executed:
        # Iteration counts the reference loop runs (Solver.cpp:175-211 breaks after the first non-productive iteration).  The
        # reference exposes no counter, so they come from the oracle's bit-exact restatement of Solve_AVX2 / Island_Single on
        # the state the timed steps ended in (stages 1-6 of one more step, then the restated SolveJoints).
        from oracle import oraclepy

        w.step_staged(solve=T.SOLVE_AVX2, island=T.ISLAND_SINGLE_SLOPPY, iters=ITERS, mask=0x3F | refpy.SAFE_PAIRS)
        _, _, _, ran = oraclepy.solve_joints(w.bodies(), w.joints(), w.contact_points(), group=8, iters=ITERS)
        executed = [int(ran[0]), int(ran[1])]
    return {
        "executed_iterations": executed,
        "value": joint_iters / wall,
        "ms_per_step": wall * 1e3 / steps,
        "cores": workers + 1,
        "joints": len(w.joints()),
        "stage_ms_per_step": {k: ms[k] / steps for k in ms if k != "PrepareIndices"},
        "solve_only": joint_iters / (ms["SolveJoints"] * 1e-3),
        "steps_per_s": steps / wall,
        "broadphase_pairs_per_s": pairs * steps / ((ms["UpdateBroadphase"] + ms["UpdatePairs"]) * 1e-3),
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
    steps = min(max(args.steps, 10), 12)  # bounded sample (the reference needs ~0.3 s per 1M-box step on 16 cores), at least 10 timed steps
    warm = min(max(args.warmup, 1), 3)
    r = reference_run(scene, args.settle, steps, warm, count_executed=True)
    sample = f"{steps} World::Update steps of the same workload after {args.settle}+{warm} untimed steps"
    line = {
        "impl": "reference", "metric": "constraint_iterations_per_sec", "value": r["value"], "unit": "constraint-iterations/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}: {scene.shape[0]} bodies, {r['joints']} joints, {ITERS[0]}+{ITERS[1]} iterations (nominal); step = World::Update (8 stages)",
                   "mode": "Solve_AVX2 / Island_SingleSloppy, reference Makefile flags (-O3 -ffast-math -mavx2 -mfma)"},
        "cpu_baseline": {"value": r["value"], "unit": "constraint-iterations/s", "cores": r["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": r["value"], "unit": "constraint-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_ms_per_step": r["stage_ms_per_step"], "steps_per_sec": r["steps_per_s"], "broadphase_pairs_per_sec": r["broadphase_pairs_per_s"],
        "solve_only_constraint_iterations_per_sec": r["solve_only"],
        "executed_iterations": r["executed_iterations"],
        "executed_constraint_iterations_per_sec": (r["joints"] * sum(r["executed_iterations"]) / (r["ms_per_step"] * 1e-3)) if r["executed_iterations"] else None,
        "counting": "value = joints x configured iterations (20 + 20, nominal) / time, the same convention as the GPU arm; executed_* uses the iterations the "
                    "loops actually run before the productive early-out",
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def rel_dev(a, b):
    """max |delta pos| / scene size, max |delta velocity| / max |velocity| (the measure tests/test_gpu_world.py uses)"""
    scale = max(float(np.abs(b["pos"]).max()), 1.0)
    dpos = float(np.abs(a["pos"][1:] - b["pos"][1:]).max()) / scale
    vscale = max(float(np.abs(b["velocity"]).max()), 1.0)
    dvel = float(np.abs(a["velocity"] - b["velocity"]).max()) / vscale
    return dpos, dvel


def parity_section(args, w, local_rank):
    """Which mode produced which number (hard part 1).  `value` / `e2e` are the throughput mode (Solve_B200:
    device colouring, strip-local kernel): the same joints relaxed in another Gauss-Seidel order, bit-exact against the oracle
    on that order (tests), not bit-comparable with the reference.  The parity mode (Solve_AVX2: replay of the reference's
    order as dependency levels) IS bit-identical to the reference (tests) and is timed here on the bench workload; and the
    throughput mode's distance from the reference after 100 steps is measured at BASELINE configs[1] size next to the
    reference's own Scalar-vs-AVX2 spread."""
    from oracle import refpy
    from phyx_b200 import scenes, world
    from phyx_b200 import types as T

    out = {"benched_mode": "Solve_B200 (throughput): colour schedule, strip-local kernel; bit-exact vs the oracle on its own slot order, NOT bit-comparable with the reference",
           "parity_mode": "Solve_AVX2 (replay): bit-identical to the reference's AVX2 / Island_Single path (tests/test_gpu_world.py)"}
    # (1) the replay mode on the bench workload, continuing from the settled state
    t0 = time.perf_counter()
    w.step(solve=T.SOLVE_AVX2, iters=ITERS)   # builds the replay schedule for this joint graph
    first = time.perf_counter() - t0
    ms = []
    while len(ms) < 2 and sum(ms) * 1e-3 + first < 60.0:
        t0 = time.perf_counter()
        w.step(solve=T.SOLVE_AVX2, iters=ITERS)
        ms.append((time.perf_counter() - t0) * 1e3)
    st = w.solve_stats()
    out["replay_on_workload"] = {"scene": args.scene, "first_step_ms": first * 1e3, "ms_per_step": float(np.mean(ms)) if ms else None, "steps": len(ms),
                                 "levels": int(st.levels), "joints": int(st.joints), "solve_ms": float(st.ms_total), "schedule_ms": float(st.ms_schedule),
                                 "note": "World::Update through the host mirror (e2e path); the reference's greedy PrepareIndices order is restated on the host each step, "
                                         "its dependency levels are deep because SIMD groups couple unrelated stacks: a parity mode, not a throughput mode"}
    # (2) throughput mode vs the reference after 100 steps, 100 k bodies
    if refpy.available("strict"):
        name = "stack_100k"
        sc = scenes.make(name)
        t0 = time.perf_counter()
        ra, rs = refpy.RefWorld(sc, "strict"), refpy.RefWorld(sc, "strict")
        g = world.World(sc, device=local_rank, mirror_contents=False)
        steps = 100
        for _ in range(steps):
            ra.step(solve=T.SOLVE_AVX2)
            rs.step(solve=T.SOLVE_SCALAR)
            g.step(solve=world.SOLVE_B200)
        ours, spread = rel_dev(g.bodies(), ra.bodies()), rel_dev(rs.bodies(), ra.bodies())
        out["throughput_mode_vs_reference"] = {"scene": name, "steps": steps, "ours_vs_ref_avx2": {"pos": ours[0], "vel": ours[1]},
                                               "ref_scalar_vs_ref_avx2": {"pos": spread[0], "vel": spread[1]},
                                               "measure": "max |delta pos| / scene size, max |delta velocity| / max |velocity|; reference = oracle/_ref strict build, Island_Single, 1 core",
                                               "seconds": time.perf_counter() - t0}
        g.close()
    return out


# ---------------------------------------------------------------------------------------------------
def run_island_parallel(args, rank, world_size, local_rank):
    """--gpus N > 1: ONE world (default islands_1m, BASELINE configs[3]) stepped by all N ranks together, the solve split BY
    ISLAND (phyx_b200/islands.py, csrc/islands.cu): strong scaling of one scene.  Every rank holds the whole world and runs the
    collider stages redundantly; a rank relaxes only its own islands (no communication inside the solve) and one integer-sum
    all-reduce over NCCL per step merges the results.  `value` = joints x (20 + 20) of that ONE world / max-over-ranks time."""
    import hashlib

    import torch
    import torch.distributed as dist

    from phyx_b200 import capi, islands, partition, scenes

    NearGpu(local_rank)   # every rank on the cores next to its own GPU (host buffers of the end-to-end leg are first touched there)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(local_rank)

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def digest(b):
        h = hashlib.sha256()
        for f in ("pos", "xVector", "velocity", "angularVelocity"):
            h.update(np.ascontiguousarray(b[f]).tobytes())
        return h.hexdigest()

    scene = scenes.make(args.scene)
    bodies = partition.body_records(scene, device=local_rank)
    nb = bodies.shape[0]
    ctx = capi.Context(local_rank)
    ipw = islands.IslandParallelWorld(ctx, bodies, local_rank)
    for _ in range(args.settle):
        ipw.step(ITERS)
    clocks = ClockSampler(local_rank)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        ipw.step(ITERS)
    barrier()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    ipw.timing.update({"stages_before_solve_ms": 0.0, "solve_ms": 0.0, "exchange_ms": 0.0, "steps": 0})
    with clocks:
        e0.record(stream)
        stats = [ipw.step(ITERS) for _ in range(args.steps)]
        e1.record(stream)
        barrier()
    clocks.close()
    launches = ctx.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    joint_iters = sum(st.joints for _, st in stats) * sum(ITERS)
    value = joint_iters / (ms_total * 1e-3)
    state = digest(ctx.download_bodies())

    # e2e: the same step with HOST buffers: every rank uploads World::bodies (pinned) and reads it back, every step
    host = torch.from_numpy(ctx.download_bodies().view(np.uint8).reshape(-1)).pin_memory()
    host_np = host.numpy().view(bodies.dtype)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    e2e_ji = 0
    for _ in range(e2e_steps):
        ctx.upload_bodies(host_np)
        _, st = ipw.step(ITERS)
        ctx.download_bodies(out=host_np)
        e2e_ji += st.joints * sum(ITERS)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps

    plan = ctx.strip_plan()
    mine = {"rank": rank, "strips": int(plan["strips"]), "cut_manifolds": int(plan["cut_manifolds"]), "manifolds_relaxed": int(np.mean([st.slots for _, st in stats])) // 2, "kernel_ms": float(np.mean([st.ms_iterations for _, st in stats])),
            "solve_ms": float(np.mean([st.ms_total for _, st in stats])), "phases_ms": ipw.mean_timing(), "schedule_ms": float(np.mean([st.ms_schedule for _, st in stats])), "kernel_form": int(stats[-1][1].kernelForm), "state": state,
            "relaxed": [float(np.mean([st.activeJointIterations[0] for _, st in stats])), float(np.mean([st.activeJointIterations[1] for _, st in stats]))]}
    per_rank = [None] * world_size
    dist.all_gather_object(per_rank, mine)

    # the same world on ONE device, same number of steps: the island-parallel result must be bit-identical (rank 0 checks)
    matches = None
    if rank == 0:
        one = capi.Context(local_rank)
        one.upload_bodies(bodies)
        for _ in range(args.settle + warm + args.steps):
            islands.stages_before_solve(one)
            one.solve_resident(iters=ITERS, schedule=capi.SCHEDULE_COLOUR)
            one.integrate_position(scenes.DT)
        matches = digest(one.download_bodies()) == state
        one.close()

    # ---- the same scene SHARDED by island: every rank simulates only its own islands (all stages scale, no exchange of state)
    sharded = None
    try:
        sctx = capi.Context(local_rank)
        sw = islands.ShardedWorld(sctx, bodies, local_rank)
        for _ in range(args.settle):
            sw.step(ITERS)
        for _ in range(warm):
            sw.step(ITERS)
        barrier()
        sstream = torch.cuda.ExternalStream(sctx.stream(), device=local_rank)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(sstream)
        sstats = [sw.step(ITERS) for _ in range(args.steps)]
        s1.record(sstream)
        barrier()
        s_ms = max_over_ranks(s0.elapsed_time(s1)) / args.steps
        sj = torch.tensor([float(sum(st.joints for _, st in sstats))], device=dev, dtype=torch.float64)
        dist.all_reduce(sj, op=dist.ReduceOp.SUM)
        shard_info = [None] * world_size
        dist.all_gather_object(shard_info, {"rank": rank, "bodies": int(sw.global_index.shape[0]), "dynamic_bodies_owned": sw.dynamic_owned,
                                            "joints": int(sstats[-1][1].joints), "ms_per_step": s0.elapsed_time(s1) / args.steps,
                                            "steps_total": sw.steps, "deferred_steps_total": sw.deferred_steps, "graph_replays_total": sw.graph_replays})
        # e2e of the sharded run: this rank's bodies up and down every step
        shost = torch.from_numpy(sctx.download_bodies().view(np.uint8).reshape(-1)).pin_memory()
        shost_np = shost.numpy().view(bodies.dtype)
        barrier()
        t0 = time.perf_counter()
        s_e2e_ji = 0
        for _ in range(e2e_steps):
            sctx.upload_bodies(shost_np)
            _, st = sw.step(ITERS)
            sctx.download_bodies(out=shost_np)
            s_e2e_ji += st.joints * sum(ITERS)
        barrier()
        s_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        sje = torch.tensor([float(s_e2e_ji)], device=dev, dtype=torch.float64)
        dist.all_reduce(sje, op=dist.ReduceOp.SUM)
        shard_bytes = torch.tensor([float(shost_np.shape[0] * 128)], device=dev, dtype=torch.float64)
        dist.all_reduce(shard_bytes, op=dist.ReduceOp.MAX)
        spans = sw.check_apart()
        merged = sw.gather_bodies(nb)
        sharded = {"ms_per_step": s_ms, "value": float(sj.item()) * sum(ITERS) / (s_ms * args.steps * 1e-3), "per_rank": shard_info,
                   "islands": {"groups": sw.island_counts[0], "largest_group_joints": sw.island_counts[1], "islands": sw.island_counts[2]},
                   "independence_check": f"all-gather of the ranks' x-extents every {sw.check_every} steps; last: {len(spans)} disjoint spans, margin {sw.margin}",
                   "merged_state_finite": bool(np.isfinite(merged["pos"]).all()),
                   "e2e": {"value": float(sje.item()) / (s_e2e_ms * e2e_steps * 1e-3), "ms_per_step": s_e2e_ms, "h2d": int(shard_bytes.item()), "d2h": int(shard_bytes.item())},
                   "note": "each rank steps only the bodies of its own islands (+ the static bodies): all eight stages scale, no state is exchanged; not bit-identical to the "
                           "one-device run (indices, hence colouring priorities, differ per shard): the replicated island_parallel path above is the bit-identical one"}
        sctx.close()
    except Exception as e:  # noqa: BLE001
        sharded = {"error": str(e)}

    spanning = None
    if args.spanning:
        try:
            spanning = partition.run_spanning("pyramid_1m", args.settle, max(3, min(args.steps, 10)), local_rank, iters=ITERS, check=True)
        except Exception as e:  # noqa: BLE001
            spanning = {"error": str(e)}
    dist.barrier()
    dist.destroy_process_group()
    if rank != 0:
        return
    jm = float(np.mean([st.joints for _, st in stats]))
    peak, peak_src = measured_peak()
    k_ms = max(r["kernel_ms"] for r in per_rank)
    act = [sum(r["relaxed"][0] for r in per_rank), sum(r["relaxed"][1] for r in per_rank)]
    alg_bytes = jm * BYTES_PRESTEP + act[0] * BYTES_IMPULSE + act[1] * BYTES_DISPLACEMENT
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    words = nb * 8 + int(jm) * 2
    replicated = {"ms_per_step": ms_step, "value": value, "per_rank": [{k: v for k, v in r.items() if k != "state"} for r in per_rank],
                  "replicas_identical": len({r["state"] for r in per_rank}) == 1, "matches_one_device_run_bit_for_bit": matches,
                  "bit_identity_condition": "holds when every rank ran the strip-local kernel (kernel_form 3) with no manifold across a strip cut (cut_manifolds 0, also on the one-device "
                                            "run): then the order in which an island's manifolds are relaxed does not depend on the partition (tests/test_gpu_islands.py)",
                  "exchange_bytes_per_step_per_rank": words * 4,
                  "exchange": "NCCL all-reduce (int32 SUM, one non-zero term per word) of [velocity rows | displacement rows | cached impulses]",
                  "e2e_ms_per_step": e2e_ms,
                  "note": "every rank holds the whole world and runs the collider stages redundantly; only SolveJoints is split by island: bit-identical to one device, "
                          "but the replicated stages bound its speed-up"}
    use_sharded = isinstance(sharded, dict) and "error" not in sharded
    head_ms = sharded["ms_per_step"] if use_sharded else ms_step
    head_value = sharded["value"] if use_sharded else value
    head_e2e = sharded["e2e"] if use_sharded else {"value": e2e_ji / (e2e_ms * e2e_steps * 1e-3), "ms_per_step": e2e_ms, "h2d": int(nb * 128), "d2h": int(nb * 128)}
    line = {
        "metric": "constraint_iterations_per_sec", "value": head_value, "unit": "constraint-iterations/s", "n_gpus": world_size, "steps": args.steps, "warmup": warm,
        "ms_per_step": head_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}: ONE world of {nb} bodies, {int(jm)} joints after {args.settle} settle steps, {ITERS[0]}+{ITERS[1]} iterations (nominal); step = World::Update "
                               f"(8 stages, colour schedule), the scene split BY ISLAND over {world_size} GPUs",
                   "inputs": "larger than L2; consecutive simulation steps (state evolves, nothing is replayed)",
                   "parallelism": (f"island-parallel x{world_size}, sharded: islands from the device union-find (csrc/islands.cu), contiguous runs of island groups with equal joint counts per rank, "
                                   "each rank steps only its own islands' bodies; no data-path collective (an all-gather of the ranks' x-extents every 8 steps checks that the shards stay apart)")
                   if use_sharded else f"island-parallel x{world_size}, replicated world, SolveJoints split by island, one NCCL all-reduce per step"},
        "e2e": {"value": head_e2e["value"], "unit": "constraint-iterations/s", "h2d_bytes_per_step": head_e2e["h2d"], "d2h_bytes_per_step": head_e2e["d2h"],
                "ms_per_step": head_e2e["ms_per_step"], "steps": e2e_steps,
                "call": "per rank: upload this rank's World::bodies (pinned), World::Update stages, download them" + (" (bytes: the largest shard)" if use_sharded else "")},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": f"{KERNEL_FORMS[per_rank[0]['kernel_form']]}, one launch per rank and step over the rank's own islands (measured on the replicated run)", "achieved": achieved,
                     "peak": peak * world_size, "unit": "GB/s", "frac": achieved / (peak * world_size), "traffic": None, "peak_source": peak_src + f" x {world_size} GPUs",
                     "formula": "SURVEY 8(d) bytes of the relaxed joint-iterations of all ranks / the slowest rank's kernel time", "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms},
        "cpu_baseline": None,
        "island_parallel_sharded": sharded,
        "island_parallel_replicated": replicated,
        "spanning": spanning,
        "executed_constraint_iterations_per_sec": jm * float(np.mean([st.contactIterationsRun + st.penetrationIterationsRun for _, st in stats])) / (head_ms * 1e-3),
        "steps_per_sec": 1e3 / head_ms,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world_size, local_rank):
    import torch

    from phyx_b200 import capi, scenes, world

    dist = None
    if world_size > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload: the scene advanced `settle` steps by full World::Update calls (untimed)
    near = NearGpu(local_rank)
    scene = scenes.make(args.scene)
    w = world.World(scene, device=local_rank, mirror_contents=False)
    for _ in range(args.settle):
        w.step(solve=world.SOLVE_B200, iters=ITERS)
    ctx = w.context()
    ctx.step_mode(args.step_mode)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    nb = scene.shape[0]

    stage_max = {}
    stage_wall = {}   # host wall time per stage call of the timed resident loop (stages that read counts back wait for the GPU)

    def timed(name, fn, *a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        ms = (time.perf_counter() - t) * 1e3
        stage_wall[name] = stage_wall.get(name, 0.0) + ms
        stage_max[name] = max(stage_max.get(name, 0.0), ms)
        return r

    step_infos = []

    def resident_step():
        # World::Update as ONE C-ABI call (phyx_b200_world_step): the eight stages with the counts kept on the device, one
        # read-back per step.  --stage-calls times the eight stage functions instead (same results, a read-back per stage).
        if not args.stage_calls:
            st, bp, info = timed("WorldStep", ctx.world_step, scenes.DT, scenes.GRAVITY, iters=ITERS, schedule=capi.SCHEDULE_COLOUR)
            step_infos.append((info.deferred, info.stopStage, info.stopReason, info.graphReplay))
            return bp, st
        return stage_calls_step()

    def stage_calls_step():
        timed("IntegrateVelocity", ctx.integrate_velocity, scenes.DT, scenes.GRAVITY)
        timed("UpdateBroadphase", ctx.update_broadphase)
        bp = timed("UpdatePairs", ctx.update_pairs)
        timed("UpdateManifolds", ctx.update_manifolds)
        timed("PackManifolds", ctx.pack_manifolds)
        timed("RefreshContactJoints", ctx.refresh_contact_joints)
        st = timed("SolveJoints", ctx.solve_resident, iters=ITERS, schedule=capi.SCHEDULE_COLOUR)
        timed("IntegratePosition", ctx.integrate_position, scenes.DT)
        return bp, st

    clocks = ClockSampler(local_rank)   # started before the warm-up so that its start-up cost is not in the timed region
    for _ in range(max(args.warmup, 3)):
        resident_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    stats = []
    stage_wall.clear()
    stage_max.clear()
    step_infos.clear()
    allocs0 = ctx.alloc_stats()
    with clocks:
        e0.record(stream)
        for _ in range(args.steps):
            stats.append(resident_step())
        e1.record(stream)
        barrier()
    clocks.close()
    allocs1 = ctx.alloc_stats()
    launches = ctx.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    joint_iters = sum(st.joints for _, st in stats) * sum(ITERS)
    value = world_size * joint_iters / (ms_total * 1e-3)
    nj = stats[-1][1].joints
    manifolds = ctx.collider_counts()[0]
    plan = ctx.strip_plan()
    # the same steps through the eight stage functions (a read-back per stage), for the per-stage host wall times
    step_wall = dict(stage_wall)
    step_infos_timed = list(step_infos)
    stage_wall.clear()
    stage_max.clear()
    stage_steps = 0
    stage_calls_ms = None
    if not args.stage_calls:
        stage_steps = max(3, min(args.steps, 10))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_calls_step()
        stage_wall.clear()
        stage_max.clear()
        torch.cuda.synchronize(local_rank)
        s0.record(stream)
        for _ in range(stage_steps):
            stage_calls_step()
        s1.record(stream)
        torch.cuda.synchronize(local_rank)
        stage_calls_ms = s0.elapsed_time(s1) / stage_steps
    else:
        stage_steps = args.steps

    # ---- e2e: World::Update through the host mirror, World::bodies uploaded and read back every step
    e2e_steps = max(3, min(args.steps, 10))
    w.step(solve=world.SOLVE_B200, iters=ITERS)
    barrier()
    t0 = time.perf_counter()
    e2e_joint_iters = 0
    for _ in range(e2e_steps):
        w.step(solve=world.SOLVE_B200, iters=ITERS)
        e2e_joint_iters += w.solve_stats().joints * sum(ITERS)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_value = world_size * e2e_joint_iters / (e2e_ms * e2e_steps * 1e-3)
    stage_ms_e2e = None
    if rank == 0:
        w.reset_stage_ms()
        w.step(solve=world.SOLVE_B200, iters=ITERS)
        stage_ms_e2e = {k: round(v, 3) for k, v in w.stage_ms().items()}

    near.restore()
    parity = None
    if args.parity and world_size == 1:
        try:
            parity = parity_section(args, w, local_rank)
        except Exception as e:  # noqa: BLE001 - the main line must still be printed
            parity = {"error": str(e)}

    # ---- optional: ONE world over all ranks (an island that spans devices), strong scaling of the solve
    spanning = None
    if args.spanning and dist is not None:
        from phyx_b200 import partition

        try:
            spanning = partition.run_spanning(args.scene, args.settle, max(3, min(args.steps, 10)), local_rank, iters=ITERS)
        except Exception as e:  # noqa: BLE001 - the main line must still be printed
            spanning = {"error": str(e)}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel (the iteration kernel: warm start + all iterations, one launch per step)
    ran_i = float(np.mean([st.contactIterationsRun for _, st in stats]))
    ran_d = float(np.mean([st.penetrationIterationsRun for _, st in stats]))
    act_i = float(np.mean([st.activeJointIterations[0] for _, st in stats]))
    act_d = float(np.mean([st.activeJointIterations[1] for _, st in stats]))
    jm = float(np.mean([st.joints for _, st in stats]))
    # bytes the reference's algorithm moves for the joint-iterations that are relaxed (and the warm start)
    alg_bytes = jm * BYTES_PRESTEP + act_i * BYTES_IMPULSE + act_d * BYTES_DISPLACEMENT
    # not §8(d), reported separately: plus two indices and two body rows for every joint-iteration the lastIteration test skips
    alg_bytes_with_skips = alg_bytes + (jm * ran_i - act_i) * BYTES_SKIP + (jm * ran_d - act_d) * BYTES_SKIP
    nominal_bytes = jm * (BYTES_PRESTEP + ran_i * BYTES_IMPULSE + ran_d * BYTES_DISPLACEMENT)
    k_ms = float(np.mean([st.ms_iterations for _, st in stats]))
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    forms = {}
    for _, st in stats:
        forms[int(st.kernelForm)] = forms.get(int(st.kernelForm), 0) + 1
    # DRAM traffic (ncu dram__bytes_read + write of one launch on this workload), per kernel form that actually ran,
    # weighted by how often each form ran; null when a form that ran has no capture
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "k_solve_traffic.json")
    if os.path.exists(tpath) and args.scene == "pyramid_1m":
        by_form = json.load(open(tpath)).get("by_form", {})
        if all(str(f) in by_form for f in forms):
            traffic = sum(by_form[str(f)]["dram_bytes_per_launch"] * n for f, n in forms.items()) / sum(forms.values())
            traffic_src = {str(f): by_form[str(f)]["source"] for f in forms}
    # radix sort of the broadphase (north_star: "achieved HBM GB/s for the solve and radix kernels"): key build + 3 LSD passes
    sort_ms = float(np.mean([bp.ms_sort for bp, _ in stats]))
    sweep_ms = float(np.mean([bp.ms_sweep for bp, _ in stats]))
    pairs_m = float(np.mean([bp.pairs for bp, _ in stats]))
    tests_m = float(np.mean([bp.tests for bp, _ in stats]))
    radix_achieved = nb * BYTES_RADIX_PER_BODY / (sort_ms * 1e-3) / 1e9 if sort_ms > 0 else None

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import refpy

        if refpy.available("fast"):
            r = reference_run(scene, args.settle, 3, 1)
            cpu = {"value": r["value"], "unit": "constraint-iterations/s", "cores": r["cores"], "kind": "reference", "broadphase_pairs_per_sec": r["broadphase_pairs_per_s"],
                   "sample": f"3 World::Update steps of the same workload after {args.settle}+1 untimed steps (Solve_AVX2 / Island_SingleSloppy, reference Makefile flags)",
                   "ms_per_step": r["ms_per_step"], "stage_ms_per_step": r["stage_ms_per_step"], "solve_only": r["solve_only"]}
        else:
            cpu = {"value": None, "unit": "constraint-iterations/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}

    bp_last = stats[-1][0]
    solve_ms = float(np.mean([st.ms_total for _, st in stats]))
    line = {
        "metric": "constraint_iterations_per_sec", "value": value, "unit": "constraint-iterations/s", "n_gpus": world_size,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}: {nb} bodies, {manifolds} manifolds, {nj} joints per GPU after {args.settle} settle steps, "
                               f"{ITERS[0]}+{ITERS[1]} iterations (nominal); step = World::Update (8 stages, colour schedule)",
                   "inputs": "larger than L2: the packed joint streams alone are ~0.3 GB per impulse iteration; consecutive simulation steps (state evolves, nothing is replayed)",
                   "parallelism": f"island-parallel x{world_size} (no data-path collective)"},
        "e2e": {"value": e2e_value, "unit": "constraint-iterations/s", "h2d_bytes_per_step": int(nb * 128), "d2h_bytes_per_step": int(nb * 128),
                "ms_per_step": e2e_ms, "steps": e2e_steps, "call": "World::Update (host mirror), bodies page-locked in place, collider mirrors: sizes only",
                "stage_wall_ms": stage_ms_e2e, "host_cores": near.info},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": "; ".join(f"{KERNEL_FORMS[f]} x{n}" for f, n in sorted(forms.items())) + " (warm start + impulse + displacement iterations, one persistent launch per step)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "formula": "SURVEY 8(d): joints x 128 B (warm start) + relaxed impulse joint-iterations x 196 B + relaxed displacement joint-iterations x 136 B, / kernel_ms",
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "iterations_run": [ran_i, ran_d], "active_joint_iterations": [act_i, act_d], "joints": jm,
                     "kernel_forms_launched": {str(f): n for f, n in sorted(forms.items())},
                     "with_skips": {"note": "NOT the 8(d) figure: adds 40 B (two indices, two body rows) per joint-iteration the lastIteration test skips",
                                    "bytes_per_launch": alg_bytes_with_skips, "achieved": alg_bytes_with_skips / (k_ms * 1e-3) / 1e9, "frac": alg_bytes_with_skips / (k_ms * 1e-3) / 1e9 / peak},
                     "nominal_bytes_per_launch": nominal_bytes},
        "roofline_radix": {"bound": "hbm", "kernel": "k_make_keys + 3 x (k_radix_hist, scan, k_radix_scatter): radixFloat keys, 11/11/10-bit LSD passes", "achieved": radix_achieved,
                           "peak": peak, "unit": "GB/s", "frac": (radix_achieved / peak) if radix_achieved else None, "ms": sort_ms,
                           "formula": "SURVEY 8(d): 56 B per body (8 histogram + 3 x 16 scatter) x bodies / CUDA-event time of the sort",
                           "note": f"{nb * 8 / 1e6:.0f} MB of keys: the passes run out of L2, the sort is launch- and latency-bound, not HBM-bound"},
        "broadphase_pairs_per_sec": world_size * pairs_m / ((sort_ms + sweep_ms) * 1e-3) if sort_ms + sweep_ms > 0 else None,
        "broadphase_tests_per_sec": world_size * tests_m / ((sort_ms + sweep_ms) * 1e-3) if sort_ms + sweep_ms > 0 else None,
        "executed_constraint_iterations_per_sec": world_size * jm * (ran_i + ran_d) / (ms_step * 1e-3),
        "relaxed_constraint_iterations_per_sec": world_size * (act_i + act_d) / (ms_step * 1e-3),
        "counting": "value = joints x configured iterations (20 + 20, nominal) / time, the convention of both arms; executed_* = joints x iterations run before the "
                    "productive early-out; relaxed_* = joint-iterations that pass the lastIteration test (counted on the device)",
        "cpu_baseline": cpu,
        "solve_ms_per_step": {"total": solve_ms, "schedule": float(np.mean([st.ms_schedule for _, st in stats])), "refresh": float(np.mean([st.ms_refresh for _, st in stats])),
                              "iterations_kernel": k_ms, "colour_rounds": int(stats[-1][1].colourRounds), "colours": int(stats[-1][1].levels)},
        "solve_only_constraint_iterations_per_sec": world_size * jm * sum(ITERS) / (solve_ms * 1e-3),
        "spanning": spanning,
        "step_call": {
            "api": "eight stage functions" if args.stage_calls else "phyx_b200_world_step",
            "deferred_steps": sum(1 for d, _, _, _ in step_infos_timed if d),
            "graph_replays": sum(1 for _, _, _, g in step_infos_timed if g),
            "stopped_steps": [[stg, why] for d, stg, why, _ in step_infos_timed if stg],
            "host_wall_ms_per_step": round(step_wall.get("WorldStep", 0.0) / args.steps, 3) if not args.stage_calls else None,
            "stage_calls_ms_per_step": stage_calls_ms,
            "note": "deferred = counts on the device, one read-back per step; graph replay = the whole step is one CUDA-graph launch (same bounds and buffers as the step before); a stopped step is finished by the stage functions (reasons: include/phyx_b200.h)",
        },
        "strip_plan": {k: v for k, v in plan.items() if k in ("strips", "usable", "rejected", "last_reject", "max_strip_rows", "max_cut_rows", "max_bin", "colours", "cut_manifolds")},
        "resident_stage_wall_ms": {k: round(v / max(stage_steps, 1), 3) for k, v in stage_wall.items()},
        "resident_stage_wall_ms_max": {k: round(v, 3) for k, v in stage_max.items()},
        "device_allocations_in_timed_region": {"count": allocs1[0] - allocs0[0], "host_ms": round(allocs1[1] - allocs0[1], 3)},
        "steps_per_sec": world_size * 1e3 / ms_step,
        "broadphase": {"pairs": int(bp_last.pairs), "tests": int(bp_last.tests), "sort_ms": sort_ms, "sweep_and_cache_filter_ms": sweep_ms},
        "parity_mode": parity,
    }
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.scene is None:
        # one GPU: the configuration the metric is quoted on; several: ONE multi-island world split by island (configs[3]),
        # for both arms, so that the driver's per-N ratio compares like with like
        args.scene = "pyramid_1m" if max(world_size, args.gpus) == 1 else "islands_1m"
    if args.impl == "reference":
        run_reference(args, rank, world_size)
    elif world_size > 1:
        run_island_parallel(args, rank, world_size, local_rank)
    else:
        run_ours(args, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
