import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from phyx_b200 import capi, scenes, world
scene = sys.argv[1]
w = world.World(scenes.make(scene), mirror_contents=False)
for _ in range(int(sys.argv[2])):
    w.step(solve=world.SOLVE_B200)
ctx = w.context()
for form, strips in ((0, 0), (1, 0), (2, 0), (3, 0)):
    ctx.solve_tuning(kernel_form=form, strips=strips)
    ms = []
    for _ in range(4):
        w.step(solve=world.SOLVE_B200)
        st = w.solve_stats()
        ms.append((st.ms_iterations, st.ms_total, st.kernelForm, st.levels))
    print(scene, "forced form", form, "->", [(round(a, 2), round(b, 2), k, l) for a, b, k, l in ms[1:]], ctx.strip_plan()["strips"])
