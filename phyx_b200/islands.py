"""Island partition of a scene across ranks (host logic of the multi-GPU path).

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
