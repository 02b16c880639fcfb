"""Island-parallel execution of ONE world over several GPUs.

Product path (device island builder, phyx_b200/csrc/islands.cu): every rank holds the whole world and runs the other stages
redundantly; Solver::SolveJoints is split BY ISLAND (reference src/Solver.cpp:73-92, 285-454): each rank relaxes a
contiguous run of island groups with about equal joint counts, with no communication inside the solve; afterwards the
ranks' results (velocity rows of their bodies, cached impulses of their joints) are merged by ONE integer-sum all-reduce
over NCCL per step (`IslandParallelWorld`, one process per GPU; `IslandGroup`, all ranks in one process, for tests).
Results are bit-identical to the one-device run of the same world whenever no manifold straddles a strip cut, which is
the case when the islands are separate piles (tests/test_gpu_islands.py).

The functions at the end (`find_islands`, `partition`, `rank_scene`) are the older geometric estimate of the islands made
from the scene description before any contact exists; they remain as host-side helpers (tests/test_islands_gloo.py).

Bodies of different islands never exchange impulses (reference src/Solver.cpp:285-454: union-find
over dynamic bodies joined by contact joints; static bodies do not merge islands), so each rank can
own a disjoint set of islands and step them with no data-path collective.  At scene-construction
time contacts do not exist yet, so islands are bounded conservatively from geometry: two dynamic
bodies whose AABBs, grown by `margin`, overlap on the x axis are put in the same island (a superset
of the contact graph's components for stacks/pyramids standing side by side).  Static bodies are
replicated on every rank.
"""
import numpy as np


def find_islands(scene, margin=1.0):
    """Returns (island_id per body, number of islands); static bodies get -1."""
    scene = np.asarray(scene, dtype=np.float32)
    static = scene[:, 5] != 0
    # conservative AABB half extent on x for a rotated box
    c, s = np.abs(np.cos(scene[:, 2])), np.abs(np.sin(scene[:, 2]))
    hx = c * scene[:, 3] + s * scene[:, 4]
    lo, hi = scene[:, 0] - hx - margin, scene[:, 0] + hx + margin
    ids = np.full(scene.shape[0], -1, dtype=np.int64)
    dyn = np.nonzero(~static)[0]
    if dyn.size == 0:
        return ids, 0
    order = dyn[np.argsort(lo[dyn], kind="stable")]
    reach = np.maximum.accumulate(hi[order])
    # a new island starts where a body's interval begins beyond everything seen so far
    starts = np.concatenate([[True], lo[order][1:] > reach[:-1]])
    ids[order] = np.cumsum(starts) - 1
    return ids, int(starts.sum())


def partition(scene, world_size, margin=1.0):
    """Contiguous (in x) runs of islands per rank, balanced by dynamic body count.
    Returns a list of index arrays (into `scene`), statics first on every rank."""
    scene = np.asarray(scene, dtype=np.float32)
    ids, count = find_islands(scene, margin)
    statics = np.nonzero(ids < 0)[0]
    sizes = np.bincount(ids[ids >= 0], minlength=count)
    bounds = np.searchsorted(np.cumsum(sizes), np.arange(1, world_size) * sizes.sum() / world_size, side="left") + 1
    owner = np.searchsorted(bounds, np.arange(count), side="right") if count else np.zeros(0, dtype=np.int64)
    parts = []
    for r in range(world_size):
        mine = np.nonzero((ids >= 0) & (owner[np.maximum(ids, 0)] == r))[0] if count else np.zeros(0, dtype=np.int64)
        parts.append(np.concatenate([statics, mine]))
    return parts


def rank_scene(scene, rank, world_size, margin=1.0):
    """The rows of `scene` that rank `rank` simulates (statics replicated) and their global indices."""
    idx = partition(scene, world_size, margin)[rank]
    return np.asarray(scene, dtype=np.float32)[idx], idx


# ---------------------------------------------------------------------------------------------------
# product path: islands from the contact graph (device), solve split by island, merged by an integer-sum all-reduce
def stages_before_solve(ctx):
    from . import scenes

    ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
    ctx.update_broadphase()
    bp = ctx.update_pairs()
    ctx.update_manifolds()
    ctx.pack_manifolds()
    ctx.refresh_contact_joints()
    return bp


def merge_model(words_per_rank, owner_of_word):
    """numpy model of the merge the ranks do on the device (k_island_pack / all-reduce / k_island_unpack): every rank
    contributes the 32-bit words it owns and zero elsewhere; the integer sum over the ranks is then the owned value of every
    word, exactly, for any bit pattern (-0.0, NaN payloads, denormals included), because each sum has one non-zero term."""
    words_per_rank = [np.asarray(w).view(np.int32) for w in words_per_rank]
    owner_of_word = np.asarray(owner_of_word)
    packed = [np.where(owner_of_word == r, w, 0).astype(np.int32) for r, w in enumerate(words_per_rank)]
    with np.errstate(over="ignore"):
        total = np.sum(np.stack(packed).astype(np.int64), axis=0).astype(np.int32)
    return packed, total


class IslandGroup:
    """`ranks` replicas of one world in ONE process (contexts may share a device): what the single-GPU tests drive.  The
    merge is the same integer sum the multi-process path does with NCCL, here with torch on the device."""

    def __init__(self, contexts, bodies, device=0):
        import torch

        self.torch = torch
        self.ctx = list(contexts)
        self.device = device
        for k, c in enumerate(self.ctx):
            c.upload_bodies(bodies)
            c.island_partition(k, len(self.ctx))
        self.buf = None

    def step(self, iters=(20, 20), schedule=0):
        from . import scenes

        torch = self.torch
        stats = []
        for c in self.ctx:
            stages_before_solve(c)
            stats.append(c.solve_resident(iters=iters, schedule=schedule))
        words = self.ctx[0].island_exchange_words()
        if self.buf is None or self.buf.shape[1] < words:
            width = (max(words, 1) + 63) // 64 * 64   # every rank's row starts 16-byte aligned (the pack kernel writes int4)
            self.buf = torch.zeros((len(self.ctx), width), dtype=torch.int32, device=f"cuda:{self.device}")
        for k, c in enumerate(self.ctx):
            c.island_pack(self.buf[k].data_ptr())
            c.synchronize()
        total = self.buf[:, :words].sum(dim=0, dtype=torch.int32).contiguous()
        torch.cuda.synchronize(self.device)
        for c in self.ctx:
            c.island_unpack(total.data_ptr())
            c.integrate_position(scenes.DT)
            c.synchronize()
        return stats


class IslandParallelWorld:
    """One process per GPU (torch.distributed, NCCL): this rank's replica of the world.  step() = World::Update with the
    solve split by island and one all-reduce of the results."""

    def __init__(self, ctx, bodies, device, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.ctx, self.device = ctx, device
        self.rank, self.ranks = dist.get_rank(group), dist.get_world_size(group)
        ctx.upload_bodies(bodies)
        ctx.island_partition(self.rank, self.ranks)
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
        self.buf = None
        self.events = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        self.timing = {"stages_before_solve_ms": 0.0, "solve_ms": 0.0, "exchange_ms": 0.0, "steps": 0}

    def step(self, iters=(20, 20)):
        from . import capi, scenes

        torch, dist = self.torch, self.dist
        e = self.events
        e[0].record(self.stream)
        bp = stages_before_solve(self.ctx)
        e[1].record(self.stream)
        st = self.ctx.solve_resident(iters=iters, schedule=capi.SCHEDULE_COLOUR)   # builds the islands, relaxes this rank's
        e[2].record(self.stream)
        words = self.ctx.island_exchange_words()
        with torch.cuda.stream(self.stream):   # everything stays ordered on the context's own stream
            if self.buf is None or self.buf.shape[0] < words:
                self.buf = torch.zeros(max(words + words // 8, 1), dtype=torch.int32, device=f"cuda:{self.device}")
            self.ctx.island_pack(self.buf.data_ptr())
            dist.all_reduce(self.buf[:words], op=dist.ReduceOp.SUM, group=self.group)
            self.ctx.island_unpack(self.buf.data_ptr())
        e[3].record(self.stream)
        self.ctx.integrate_position(scenes.DT)
        if self.timing["steps"] >= 0:
            e[3].synchronize()
            self.timing["stages_before_solve_ms"] += e[0].elapsed_time(e[1])
            self.timing["solve_ms"] += e[1].elapsed_time(e[2])
            self.timing["exchange_ms"] += e[2].elapsed_time(e[3])
            self.timing["steps"] += 1
        return bp, st

    def mean_timing(self):
        n = max(self.timing["steps"], 1)
        return {k: v / n for k, v in self.timing.items() if k != "steps"}


class ShardedWorld:
    """Island-parallel execution WITHOUT replication: every rank simulates only the bodies of its own islands (plus the static
    bodies) as a world of its own, so all eight stages scale with the rank count and a step needs no exchange of state at
    all.  This is the reference's island split (src/Solver.cpp:73-92: islands are solved independently) taken to the whole
    step.  It is valid while the ranks' islands stay apart, which is checked every `check_every` steps by an all-gather of
    each rank's x-extent (the only collective; a violation raises: re-partition with `IslandParallelWorld`, which handles any
    contact graph).  Shards are cut on the contact graph: the full world is stepped `probe_steps` steps on every rank, the
    device island builder assigns the island groups to the ranks, and each rank keeps its own bodies.  Results are NOT
    bit-identical to the one-device run of the full scene (body and manifold indices differ per shard, and with them the
    colouring priorities); use IslandParallelWorld where that is required."""

    def __init__(self, ctx, bodies, device, group=None, probe_steps=10, check_every=8, margin=25.0):
        import torch
        import torch.distributed as dist

        from . import capi, scenes

        self.torch, self.dist, self.group = torch, dist, group
        self.ctx, self.device = ctx, device
        self.rank, self.ranks = dist.get_rank(group), dist.get_world_size(group)
        self.check_every, self.margin, self.steps = check_every, margin, 0
        self.deferred_steps = 0
        self.graph_replays = 0
        # probe: the contact graph of the full scene after a few full steps, identical on every rank
        import ctypes as C

        ctx.upload_bodies(bodies)
        for _ in range(probe_steps):
            stages_before_solve(ctx)
            ctx.solve_resident(iters=(20, 20), schedule=capi.SCHEDULE_COLOUR)
            ctx.integrate_position(scenes.DT)
        probed = ctx.download_bodies()   # the state the shards start from: after the last complete step
        stages_before_solve(ctx)         # contacts of the next step: the graph the shards are cut on
        ctx.island_partition(self.rank, self.ranks)
        self.island_counts = ctx.build_islands()   # (groups, largest group, islands); also assigns the groups to the ranks
        owner = np.zeros(bodies.shape[0], np.uint8)
        ctx._check(ctx.l.phyx_b200_download_body_owners(ctx.h, owner.ctypes.data_as(C.c_void_p), owner.shape[0]))
        ctx.island_partition(0, 1)
        # this rank's world: the static bodies and the bodies of its own islands, in their original order
        static = (bodies["invMass"] == 0) & (bodies["invInertia"] == 0)
        self.global_index = np.nonzero(static | (owner == self.rank))[0]
        self.dynamic_owned = int((~static & (owner == self.rank)).sum())
        ctx.reset_collider()
        mine = np.array(probed[self.global_index], copy=True)
        mine["index"] = np.arange(mine.shape[0], dtype=np.uint32)
        ctx.upload_bodies(mine)
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
        ctx.update_broadphase()   # (fills the AABBs' sorted order; the extent check below only needs the uploaded AABBs)
        self.check_apart()        # a pile that was not one island yet when the shards were cut would straddle two ranks

    def step(self, iters=(20, 20)):
        from . import capi, scenes

        # World::Update of this rank's shard as one C-ABI call (counts stay on the device, one read-back per step)
        st, bp, info = self.ctx.world_step(scenes.DT, scenes.GRAVITY, iters=iters, schedule=capi.SCHEDULE_COLOUR)
        self.deferred_steps += int(info.deferred)
        self.graph_replays += int(info.graphReplay)
        self.steps += 1
        if self.check_every and self.steps % self.check_every == 0:
            self.check_apart()
        return bp, st

    def check_apart(self):
        """all-gather of the ranks' dynamic x-extents; raises if two ranks' islands could touch"""
        torch, dist = self.torch, self.dist
        e = self.ctx.dynamic_extent()   # reduced on the device: 16 bytes come back
        lo, hi = (float(e[0]), float(e[2])) if self.dynamic_owned else (float("inf"), float("-inf"))
        t = torch.tensor([lo, hi], dtype=torch.float64, device=f"cuda:{self.device}" if dist.get_backend(self.group) == "nccl" else "cpu")
        ext = [torch.zeros_like(t) for _ in range(self.ranks)]
        dist.all_gather(ext, t, group=self.group)
        spans = sorted((float(e[0]), float(e[1]), k) for k, e in enumerate(ext) if float(e[0]) <= float(e[1]))
        for (lo1, hi1, k1), (lo2, hi2, k2) in zip(spans[:-1], spans[1:]):
            if lo2 - hi1 < self.margin:
                raise RuntimeError(f"islands of ranks {k1} and {k2} are within {self.margin} of each other: the shards are no longer independent")
        return spans

    def gather_bodies(self, total):
        """the full scene's body records on every rank (each body from the rank that owns it; static bodies from rank 0)"""
        dist = self.dist
        mine = self.ctx.download_bodies()
        parts = [None] * self.ranks
        dist.all_gather_object(parts, (self.global_index, mine), group=self.group)
        out = np.zeros(total, dtype=mine.dtype)
        for idx, rec in reversed(parts):   # rank 0 last: its copy of the static bodies wins
            out[idx] = rec
        out["index"] = np.arange(total, dtype=np.uint32)
        return out
