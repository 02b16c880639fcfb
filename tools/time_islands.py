"""Developer aid: wall time of the device island builder (phyx_b200_build_islands) on a settled scene."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyx_b200 import scenes, world
scene = sys.argv[1] if len(sys.argv) > 1 else "islands_1m"
w = world.World(scenes.make(scene), mirror_contents=False)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    w.step(solve=world.SOLVE_B200)
ctx = w.context()
ctx.build_islands()
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    r = ctx.build_islands()
ctx.synchronize()
print(scene, "islands (groups, largest, before coalescing):", r, f"{(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per build")
