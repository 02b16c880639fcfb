"""Developer aid: where a pass of the strip-local solve kernel spends its time (per CTA %globaltimer stamps).
usage: python tools/strip_trace.py [scene] [settle steps] [strips]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phyx_b200 import capi, scenes, world

scene = sys.argv[1] if len(sys.argv) > 1 else "pyramid_1m"
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 30
strips = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w = world.World(scenes.make(scene), mirror_contents=False)
ctx = w.context()
ctx.solve_tuning(strips=strips)
for _ in range(settle):
    w.step(solve=world.SOLVE_B200, iters=(20, 20))
ctx.strip_trace(passes=48)
w.step(solve=world.SOLVE_B200, iters=(20, 20))
st = w.solve_stats()
t = ctx.strip_trace(fetch=True)
print("plan", {k: v for k, v in ctx.strip_plan().items() if not hasattr(v, "shape")})
print(f"kernel {st.ms_iterations:.3f} ms, form {st.kernelForm}, ran {st.contactIterationsRun}+{st.penetrationIterationsRun}, active {list(st.activeJointIterations)}, wake {st.wakePasses}")
if t is None:
    sys.exit("no trace")
t = t.astype(np.int64)
S = t.shape[0]
t0 = t[:, 0, 0].min()
print("pass: start skew(us) | interior med/max | waitA med/max | cut med/max | waitB med/max | pass med/max   (us)")
for p in range(t.shape[1]):
    if not t[:, p, 0].any():
        continue
    a = t[:, p, :5]
    ok = a[:, 4] > 0
    st_ = a[ok, 0]
    interior = (a[ok, 1] - a[ok, 0]) * 1e-3
    has_cut = a[:, 3] > 0
    waitA = (a[has_cut & (a[:, 2] > 0), 2] - a[has_cut & (a[:, 2] > 0), 1]) * 1e-3
    cut = (a[has_cut & (a[:, 2] > 0), 3] - a[has_cut & (a[:, 2] > 0), 2]) * 1e-3
    last = np.where(a[:, 3] > 0, a[:, 3], a[:, 1])
    waitB = (a[ok, 4] - last[ok]) * 1e-3
    total = (a[ok, 4] - a[ok, 0]) * 1e-3
    f = lambda x: f"{np.median(x):6.1f}/{x.max():6.1f}" if x.size else "   -  /   -  "
    s1, s2 = t[ok, p, 5] / 1965.0, t[ok, p, 6] / 1965.0   # SM clocks -> us at 1965 MHz
    print(f"{p:3d}: {(st_.min() - t0) * 1e-3:8.1f} +{(st_.max() - st_.min()) * 1e-3:6.1f} | {f(interior)} | {f(waitA)} | {f(cut)} | {f(waitB)} | {f(total)} | interior step1 {f(s1)} step2 {f(s2)}")
