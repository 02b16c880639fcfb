import sys
sys.path.insert(0, "tests")
from phyx_b200 import capi, scenes, world
scene = sys.argv[1] if len(sys.argv) > 1 else "stack_10k"
w = world.World(scenes.make(scene), mirror_contents=False)
ctx = w.context(); ctx.upload_bodies(w.bodies())
for step in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    st, bp, info = ctx.world_step(scenes.DT, scenes.GRAVITY)
    print(step, info.deferred, info.graphReplay, info.graphStatus, info.stopStage, info.stopReason, info.manifolds, info.joints, ctx.l.phyx_b200_last_error().decode() if info.graphStatus == -1 else "")
