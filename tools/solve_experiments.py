"""Developer aid (GPU box): where does a level pass of the record-form solve kernel spend its time?

Settles the bench scene, then runs the resident solve on ONE frozen state (snapshot / restore) under
the timing-only switches of solve.cu (PHYX_SOLVE_EXPERIMENT; run with PHYX_SOLVE_PAIRS=2 so that the record
form is the kernel): 0 = real kernel, 5 = with the L2 prefetch of the records predicted active, 1 = no joint
passes the skip test (barriers + index words + row gathers), 2 = no row gathers either (barriers + index
words).  1 and 2 compute wrong results by design.  Prints kernel ms and microseconds per level pass.
usage: PHYX_SOLVE_PAIRS=2 python tools/solve_experiments.py [scene] [settle]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from phyx_b200 import capi, scenes, world  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "pyramid_1m"
    settle = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    w = world.World(scenes.make(name), device=0, mirror_contents=False)
    for _ in range(settle):
        w.step(solve=world.SOLVE_B200, iters=(20, 20))
    ctx = w.context()
    ctx.integrate_velocity(scenes.DT, scenes.GRAVITY)
    ctx.update_broadphase()
    ctx.update_pairs()
    ctx.update_manifolds()
    ctx.pack_manifolds()
    ctx.refresh_contact_joints()
    ctx.snapshot_bodies()
    for exp in (0, 0, 5, 1, 2, 0):
        os.environ["PHYX_SOLVE_EXPERIMENT"] = str(exp)
        ctx.restore_bodies()
        st = ctx.solve_resident(iters=(20, 20), schedule=capi.SCHEDULE_COLOUR)
        passes = st.levels * (1 + st.contactIterationsRun + st.penetrationIterationsRun)
        print(f"experiment {exp}: kernel {st.ms_iterations:.3f} ms, {st.levels} levels, iterations {st.contactIterationsRun}+{st.penetrationIterationsRun}, "
              f"{passes} level passes, {st.ms_iterations * 1e3 / passes:.2f} us per pass, active {st.activeJointIterations[0]}+{st.activeJointIterations[1]} of {st.joints} joints",
              flush=True)


if __name__ == "__main__":
    main()
