"""Developer probe: step a scene with the host mirror on the GPU and print per-stage times.
usage: python tools/probe.py <scene> <steps> [solve_mode]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyx_b200 import scenes, world  # noqa: E402

scene, steps = sys.argv[1], int(sys.argv[2])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else world.SOLVE_B200
t0 = time.time()
sc = scenes.make(scene)
w = world.World(sc)
print(f"scene {scene}: {sc.shape[0]} bodies, built in {time.time() - t0:.2f}s", flush=True)
for s in range(steps):
    w.reset_stage_ms()
    t0 = time.time()
    w.step(solve=mode)
    dt = time.time() - t0
    st = w.solve_stats()
    bp = w.broadphase_stats()
    ms = {k: round(v, 2) for k, v in w.stage_ms().items()}
    print(json.dumps({"step": s, "wall_ms": round(dt * 1e3, 1), "stage_ms": ms, "solve": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.as_dict().items()},
                      "pairs": bp.pairs, "tests": bp.tests}), flush=True)
