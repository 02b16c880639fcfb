"""Developer aid: world_step against the stage functions, step by step, printing what differs first."""
import sys
import numpy as np
sys.path.insert(0, "tests")
from phyx_b200 import capi, scenes, world
from test_gpu_world_step import stage_step

scene = sys.argv[1] if len(sys.argv) > 1 else "pyramid_1k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sc = scenes.make(scene)
wa, wb = world.World(sc, mirror_contents=False), world.World(sc, mirror_contents=False)
a, b = wa.context(), wb.context()
a.upload_bodies(wa.bodies()); b.upload_bodies(wb.bodies())
b.step_mode(mode)
for step in range(steps):
    sa = stage_step(a)
    sb, bp, info = b.world_step(scenes.DT, scenes.GRAVITY)
    print(step, info.as_dict())
    print("   A", sa.as_dict())
    print("   B", sb.as_dict())
    print("   plan A", a.strip_plan(), "\n   plan B", b.strip_plan())
    print("   counts", a.collider_counts(), b.collider_counts())
    for name, fa, fb in (("bodies", a.download_bodies, b.download_bodies), ("manifolds", a.download_manifolds, b.download_manifolds),
                         ("points", a.download_contact_points, b.download_contact_points), ("joints", a.download_joints, b.download_joints)):
        xa, xb = fa(), fb()
        if xa.shape != xb.shape:
            print("   ", name, "shapes", xa.shape, xb.shape)
            continue
        bad = [f for f in xa.dtype.names if not np.array_equal(np.ascontiguousarray(xa[f]).view(np.uint8), np.ascontiguousarray(xb[f]).view(np.uint8))]
        if bad:
            f = bad[0]
            idx = np.nonzero((xa[f] != xb[f]).reshape(len(xa), -1).any(axis=1))[0]
            print("   ", name, "differ in", bad, "first rows", idx[:8], "of", len(idx))
    sla, lva = a.get_schedule(); slb, lvb = b.get_schedule()
    print("   schedule equal:", np.array_equal(sla, slb), np.array_equal(lva, lvb), len(sla), len(slb))
