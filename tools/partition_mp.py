"""One world over N GPUs, one process per GPU (launch under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/partition_mp.py --scene pyramid_1m --steps 20 --settle 30 [--check]

Every rank steps its replica of the world (all stages but the solve redundantly) and calls the collective
phyx_b200_solve_partitioned; boundary rows travel between the GPUs by peer stores (CUDA IPC mapped buffers),
torch.distributed (gloo) is used for the handle exchange, the barriers and the max-over-ranks only.
--check: rank 0 repeats the run with all ranks as contexts on ITS device (partition.LocalGroup) and requires
bit-identical bodies; it also times the one-device solve of the same state for comparison.
Prints one JSON line on rank 0.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from phyx_b200 import capi, partition, scenes  # noqa: E402

ITERS = (20, 20)


def digest(bodies):
    h = hashlib.sha256()
    for f in ("pos", "xVector", "velocity", "angularVelocity"):
        h.update(np.ascontiguousarray(bodies[f]).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="pyramid_1k")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--settle", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()

    import torch.distributed as dist

    rank, ranks, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")
    bodies = partition.body_records(scenes.make(args.scene), device=local)
    n = bodies.shape[0]
    ctx = capi.Context(local)
    world = partition.ReplicatedWorld(ctx, bodies)
    partition.attach_process_group(ctx, partition.default_capacities(n, ranks))

    def run(steps):
        out = []
        for _ in range(steps):
            t0 = time.perf_counter()
            bp, st = world.step(ITERS)
            ctx.synchronize()
            out.append(((time.perf_counter() - t0) * 1e3, st))
        return out

    run(args.settle)
    dist.barrier()
    t0 = time.perf_counter()
    timed = run(args.steps)
    dist.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    mine = ctx.download_bodies()
    digests = [None] * ranks
    dist.all_gather_object(digests, digest(mine))
    per_rank = [None] * ranks
    st_last = timed[-1][1]
    dist.all_gather_object(per_rank, {"solve_ms": float(np.mean([s.ms_total for _, s in timed])), "iterations_ms": float(np.mean([s.ms_iterations for _, s in timed])),
                                      "exchange_finish_ms": float(np.mean([s.ms_finish for _, s in timed])), "schedule_ms": float(np.mean([s.ms_schedule for _, s in timed])),
                                      "active": [int(st_last.activeJointIterations[0]), int(st_last.activeJointIterations[1])],
                                      "iterations_run": [int(st_last.contactIterationsRun), int(st_last.penetrationIterationsRun)]})
    line = None
    if rank == 0:
        cuts, bstart, cls = ctx.partition_plan(ranks)
        line = {"scene": args.scene, "bodies": int(n), "joints": int(st_last.joints), "ranks": ranks, "steps": args.steps, "settle": args.settle,
                "ms_per_step_wall": wall_ms, "replicas_identical": len(set(digests)) == 1, "per_rank": per_rank,
                "row_cuts": cuts.tolist(), "boundary_rows": int(bstart[-1]), "cut_slots": int(cls[-1] - cls[-2]), "slots": int(cls[-1]),
                "levels": int(st_last.levels)}
    if args.check and rank == 0:
        # the same run with every rank as a context on this device, and the one-device solve for timing
        ctxs = [capi.Context(local) for _ in range(ranks)]
        worlds = [partition.ReplicatedWorld(c, bodies) for c in ctxs]
        group = partition.LocalGroup(ctxs, devices=[local] * ranks)
        one = capi.Context(local)
        one.upload_bodies(bodies)
        one_ms = []
        for step in range(args.settle + args.steps):
            for w in worlds:
                w.stages_before_solve()
            group.solve(ITERS)
            for c in ctxs:
                c.integrate_position(scenes.DT)
            one.integrate_velocity(scenes.DT, scenes.GRAVITY)
            one.update_broadphase()
            one.update_pairs()
            one.update_manifolds()
            one.pack_manifolds()
            one.refresh_contact_joints()
            st = one.solve_resident(iters=ITERS, schedule=capi.SCHEDULE_COLOUR)
            one.integrate_position(scenes.DT)
            if step >= args.settle:
                one_ms.append(st.ms_total)
        line["matches_in_process_group"] = digest(ctxs[0].download_bodies()) == digests[0]
        line["one_device_solve_ms"] = float(np.mean(one_ms))
        group.close()
    dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
