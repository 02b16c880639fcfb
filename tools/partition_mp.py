"""One world over N GPUs, one process per GPU (launch under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/partition_mp.py --scene pyramid_1m --steps 20 --settle 30 [--check]

Every rank steps its replica of the world (all stages but the solve redundantly) and calls the collective
phyx_b200_solve_partitioned; boundary rows travel between the GPUs by peer stores (CUDA IPC mapped buffers),
torch.distributed is used for the handle exchange, the barriers and the max-over-ranks only
(phyx_b200.partition.run_spanning).  --check: rank 0 repeats the run with all ranks as contexts on ITS device
and requires bit-identical bodies; it also times the one-device solve of the same states.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from phyx_b200 import partition  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="pyramid_1k")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--settle", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if args.backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    out = partition.run_spanning(args.scene, args.settle, args.steps, local, check=args.check)
    if dist.get_rank() == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
