"""Developer aid (GPU box): World::Update through the C++ host mirror with and without the opt-in lazyBodies contract."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyx_b200 import scenes, world
sc = scenes.make("pyramid_1m")
for lazy in (False, True):
    w = world.World(sc, mirror_contents=False, lazy_bodies=lazy)
    for _ in range(33):
        w.step(solve=world.SOLVE_B200)
    w.context().synchronize()
    t = time.perf_counter()
    for _ in range(20):
        w.step(solve=world.SOLVE_B200)
    w.context().synchronize()
    ms = (time.perf_counter() - t) * 1e3 / 20
    t = time.perf_counter(); b = w.bodies(); sync_ms = (time.perf_counter() - t) * 1e3
    print(f"lazy_bodies={lazy}: {ms:.3f} ms per World::Update (host mirror), bodies() afterwards {sync_ms:.2f} ms", flush=True)
    w.close()
