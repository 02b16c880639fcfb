"""Developer aid (GPU box): colour count / colouring rounds / solve kernel ms per step of a scene.
usage: python tools/colour_trace.py [scene] [steps]   (PHYX_COLOUR_HINTS=1, PHYX_COLOUR_DRIFT=n to compare)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from phyx_b200 import capi, partition, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pyramid_1m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
ctx = capi.Context(0)
w = partition.ReplicatedWorld(ctx, partition.body_records(scenes.make(name)))
out = []
for step in range(steps):
    w.stages_before_solve()
    st = ctx.solve_resident(schedule=capi.SCHEDULE_COLOUR)
    ctx.integrate_position(scenes.DT)
    _, levels = ctx.get_schedule() if step in (0, steps - 1) else (None, None)
    out.append((step, st.levels, st.colourRounds, round(st.ms_iterations, 3), st.kernelForm, int(st.activeJointIterations[0])))
    if levels is not None:
        print("level sizes at step", step, [(int(l["end"]) - int(l["start"])) // 2 for l in levels])
print(name, os.environ.get("PHYX_COLOUR_HINTS"), "step: levels rounds kernel_ms form active")
for o in out[:3] + out[3::6]:
    print("  ", o)
